#!/usr/bin/env python
"""In-kernel phase timeline of the update kernels (LIO_TIMELINE=1): python tools/timeline.py [--poses K]"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

os.environ["LIO_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from agi_lidar_slam_b200 import _cabi  # noqa: E402

TAGS = {1: "pass entry", 2: "pass const ready", 3: "search done", 4: "finish done", 5: "tile reduced", 6: "partial written",
        7: "ticket taken", 8: "released", 30: "stage: queries ready", 31: "stage: cell set built", 32: "stage: probes done, copies issued", 33: "stage: copies landed", 10: "solver: start", 13: "solver: rows seen (warp 7)", 11: "solver: partials reduced", 12: "solver: solved", 20: "solve: start", 21: "solve: u = S^-1 v done", 22: "solve: dx done", 23: "solve: boxplus + ctrl done, state published"}


def show(name, tl):
    a, b = tl
    ev = sorted([(t, "blk0  " + TAGS.get(k, str(k))) for k, t in a] + [(t, "SOLVER " + TAGS.get(k, str(k))) for k, t in b])
    if not ev:
        print(name, ": no events")
        return
    t0 = ev[0][0]
    print(f"--- {name}: total {(ev[-1][0] - t0) / 1000.0:.2f} us")
    prev = t0
    for t, n in ev:
        print(f"   {(t - t0) / 1000.0:8.2f} us  (+{(t - prev) / 1000.0:6.2f})  {n}")
        prev = t


def spread(name, ctx, t_ref=None):
    """Spread of the worker blocks and of the solver's warps in the last pass (lio_debug_blocks)."""
    t = ctx.debug_blocks()
    if os.environ.get("LIO_TIMELINE_DUMP"):
        np.save(os.environ["LIO_TIMELINE_DUMP"] + "_" + name.replace(" ", "_") + ".npy", t)
    filed, left, seen, loaded = t[0:256], t[256:512], t[512:528], t[544:560]
    f = filed[filed > 0]
    t0 = f.min()
    q = lambda v: "min %+.2f  p50 %+.2f  p90 %+.2f  max %+.2f us" % tuple((np.percentile(v, p) - t0) / 1000.0 for p in (0, 50, 90, 100))
    print(f"   {name}: {len(f)} rows filed: {q(f)} (relative to the first row)")
    l = left[left > 0] if name.startswith("search") else left[:0]  # (the slots keep the last search pass's times)
    if len(l):
        print(f"      searches left: {q(l)}")
    if len(l):
        dur = (filed - left)[filed > 0] / 1000.0
        print("      by worker index (octiles), row filed after the first [us]:",
              " ".join("%.1f" % ((filed[filed > 0][k::8].mean() - t0) / 1000.0) for k in range(1)),
              "| mean over index ranges:", " ".join("%.1f" % ((c.mean() - t0) / 1000.0) for c in np.array_split(f, 8)))
    first = t[576:768]
    if len(l) and (first > 0).any():
        order = np.argsort(filed)[-10:]
        print("      first tile searched / all tiles searched / row filed of the ten last:",
              ", ".join("%d: %+.1f %+.1f %+.1f" % (b, (first[b] - t0) / 1000.0, (left[b] - t0) / 1000.0, (filed[b] - t0) / 1000.0) for b in order))
        order = np.argsort(filed)[:5]
        print("      ... of the five first:",
              ", ".join("%d: %+.1f %+.1f %+.1f" % (b, (first[b] - t0) / 1000.0, (left[b] - t0) / 1000.0, (filed[b] - t0) / 1000.0) for b in order if filed[b] > 0))
    print("      ten last workers:", ", ".join("%d:%+.2f" % (b, (filed[b] - t0) / 1000.0) for b in np.argsort(filed)[-10:]))
    print(f"      solver warps, rows seen:   {q(seen[seen > 0])}")
    print(f"      solver warps, rows loaded: {q(loaded[loaded > 0])}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=2_000_000)
    ap.add_argument("--rings", type=int, default=128)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--poses", type=int, default=4)
    ap.add_argument("--map-cell", type=float, default=1.5)
    ap.add_argument("--pose", type=int, default=0)
    a = ap.parse_args()
    a.workload = "os1_128_2m"
    wl = bench.make_workload(a, 0)
    mp = wl["map"]
    ctx = _cabi.Context(0, max_scan_points=max(1 << 18, a.rings * a.cols), max_down_points=100000,
                        max_map_points=max(1 << 21, int(len(mp) * 1.05)), map_cell=a.map_cell)
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    s = wl["scans"][a.pose]
    body, _, _ = ctx.scan_preprocess(s["scan"], None, None, wl["leaf"])
    print("M =", len(body))
    ctx.scan_upload(body)
    ctx.state_upload(s["x_prior"], wl["P"])
    for rep in range(3):
        ctx.pass_only_enqueue(True, False)
        tl = ctx.debug_timeline()
        if rep == 2:
            show("single search pass (warm)", tl)
            r = ctx.timeline_raw
            names = ["search_tile entry .. body->world done", "probes resolved", "bucket list scanned", "(ring rounds)",
                     "top-5 merged", "neighbours published"]
            print("   query 0 of block 0: " + ", ".join("%s +%.2f us" % (nm, (r[240 + k] - r[239 + k]) / 1000.0)
                                                         for k, nm in enumerate(names)))
            spread("search pass", ctx)
        ctx.pass_only_enqueue(False, False)
        tl = ctx.debug_timeline()
        if rep == 2:
            show("single cached pass (warm)", tl)
            spread("cached pass", ctx)
    for rep in range(3):
        ctx.update_enqueue(0.001, 4, False, from_snapshot=True)
        tl = ctx.debug_timeline()
        if rep == 2:
            show("whole update (warm)", tl)
    r = ctx.timeline_raw
    print("   last staged tile of block 0: %d distinct cells, %d points, overflow %d" % (int(r[220]), int(r[221]), int(r[222])))
    print(ctx.state_download()[2:])


if __name__ == "__main__":
    main()
