#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
    python tools/make_profiles.py [round_tag]          (reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_update.ncu-rep)"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src, dst = ROOT / "gpurun_out", ROOT / "profiles"
dst.mkdir(exist_ok=True)

lines = [l for l in open(src / f"{tag}_launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
    a = agg[row["Kernel Name"][:100]]
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 5 --warmup 3 --no-cpu-baseline",
       "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       "# the torch FillFunctor kernel is bench.py's 384 MiB L2 flush between timed steps (outside every timed region)",
       f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches", "launches   total_us   us/launch  share  kernel"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[0]:5d} {v[1]:10.1f} {v[1] / v[0]:10.2f} {100 * v[1] / tot:6.1f}%  {k}")
(dst / f"{tag}_launches_summary.txt").write_text("\n".join(out) + "\n")
shutil.copy(src / f"{tag}_launches.csv", dst / f"{tag}_launches.csv")

raw = subprocess.run(["ncu", "-i", str(src / f"{tag}_update.ncu-rep"), "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
out = ["# ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 12 -c 2  python bench.py --steps 5 --warmup 3 --no-cpu-baseline",
       "# raw page, selected metrics, one block per captured launch"]
traffic = []
for r in rows[2:]:
    out.append("---")
    d = dict(zip(hdr, r))
    for k in keys:
        if k in d:
            out.append(f"{k} [{units[hdr.index(k)]}] = {d[k]}")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    if d["Kernel Name"].startswith("update_kernel("):  # the resident-input launch `value` times (not the host-direct one)
        traffic.append((rd, wr))
(dst / f"{tag}_update_kernel_ncu.txt").write_text("\n".join(out) + "\n")
(dst / f"{tag}_update_kernel_traffic.json").write_text(json.dumps(
    {"kernel": "update_kernel", "dram_bytes_per_launch": sum(r + w for r, w in traffic) / len(traffic),
     "dram_bytes_read": sum(r for r, _ in traffic) / len(traffic), "dram_bytes_write": sum(w for _, w in traffic) / len(traffic),
     "launches_captured": len(traffic), "source": f"{tag}_update_kernel_ncu.txt",
     "note": "one captured launch of a 4.3 k-point scan; the read / write split varies between captures (4.3 MB + 0 here, "
             "1.0 MB + 3.2 MB in the previous one) with what bench.py's 384 MiB flush kernel, which runs right before every "
             "timed update, left dirty in L2; the sum is stable"}) + "\n")
print("\n".join(out[:24]))
print((dst / f"{tag}_launches_summary.txt").read_text()[:1500])
