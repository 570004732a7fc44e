#!/usr/bin/env python
"""Aggregate the ncu source page by CUDA source line: python tools/ncu_lines.py rep.ncu-rep kernel_regex [launch_skip] [top]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
fname = "?"
hdr = None
rows = []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        s = int(r[hdr.index("# Samples")])
        ie = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    rows.append((s, ie, fname, r[0], r[1].strip()))
tot = sum(r[0] for r in rows)
print("total samples", tot, "total warp instr", sum(r[1] for r in rows))
for s, ie, f, ln, src in sorted(rows, key=lambda r: -r[0])[:top]:
    print(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}% {ie:9d}  {f}:{ln}  {src[:110]}")
