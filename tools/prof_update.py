#!/usr/bin/env python
"""Small driver for ncu: the bench workload (one pose), a few single passes and a few whole updates.
Usage: python tools/prof_update.py [--map-points N] [--reps K]"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from agi_lidar_slam_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=2_000_000)
    ap.add_argument("--rings", type=int, default=128)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--poses", type=int, default=1)
    a = ap.parse_args()
    a.workload = "os1_128_2m"
    wl = bench.make_workload(a, 0)
    mp = wl["map"]
    ctx = _cabi.Context(0, max_scan_points=max(1 << 18, a.rings * a.cols), max_down_points=100000,
                        max_map_points=max(1 << 21, int(len(mp) * 1.05)))
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    s = wl["scans"][0]
    body, _, _ = ctx.scan_preprocess(s["scan"], None, None, wl["leaf"])
    ctx.scan_upload(body)
    ctx.state_upload(s["x_prior"], wl["P"])
    for _ in range(a.reps):
        ctx.pass_only_enqueue(True, False)
        ctx.pass_only_enqueue(False, False)
    ctx.synchronize()
    for _ in range(a.reps):
        ctx.update_enqueue(0.001, 4, False, from_snapshot=True)
    ctx.synchronize()
    x, P, nv, npass = ctx.state_download()
    print("M", len(body), "valid", nv, "passes", npass)


if __name__ == "__main__":
    main()
