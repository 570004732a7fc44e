#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py gpurun_out/<x>_launches.csv [out.txt]"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
    a = agg.setdefault(row["Kernel Name"][:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# {sys.argv[1]}: total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)",
       "launches   total_us   us/launch  share  kernel"]
for k, v in agg.items():
    out.append(f"{v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:10.2f} {100 * v[1] / tot:6.1f}%  {k}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text, end="")
