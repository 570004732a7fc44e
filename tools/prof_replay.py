#!/usr/bin/env python
"""Per-stage host time of the replayed main loop (agi_lidar_slam_b200.replay) on one sequence.
    python tools/prof_replay.py [vlp16|os1_64] [n_scans]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from agi_lidar_slam_b200 import _cabi, synth  # noqa: E402
from agi_lidar_slam_b200.replay import LASER_POINT_COV, LioReplay, MeasureGroup, ReplayConfig  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "vlp16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
kw = dict(rings=16, cols=1800, fov=(-15.0, 15.0), max_range=100.0) if kind == "vlp16" else \
    dict(rings=64, cols=1024, fov=(-22.5, 22.5), max_range=120.0)
seq = synth.sequence(n, 2002, **kw)
ctx = _cabi.Context(0, max_scan_points=1 << 17, max_down_points=1 << 16, max_map_points=1 << 21)
r = LioReplay(ctx, ReplayConfig(max_iteration=3))
acc = {}


def timed(name, f, *a, **k):
    t0 = time.perf_counter()
    out = f(*a, **k)
    acc.setdefault(name, []).append(time.perf_counter() - t0)
    return out


for j, m in enumerate(seq):
    mg = MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"])
    if j < 14:
        r.process(mg)
        continue
    c = r.cfg
    if j % 2 == 0:  # the product path: host stage + one lio_scan_step
        poses = timed("host_stage (imu.process + fov_segment)", r.host_stage, mg)
        rep = timed("lio_scan_step", ctx.scan_step, mg.lidar, poses, r.x, r.P, c.filter_size_surf, c.filter_size_map,
                    LASER_POINT_COV, c.max_iteration, c.extrinsic_est, True)
        r.adopt(rep)
        mm = rep.m
    else:  # the same scan in its pieces, synchronised after each: where the step's time goes
        poses = r.host_stage(mg)
        timed("  begin (H2D scan + preprocess + prior) + sync", lambda: (ctx.scan_step_begin(mg.lidar, poses, r.x, r.P,
                                                                 c.filter_size_surf), ctx.synchronize()))
        timed("  update kernel + sync", lambda: (ctx.update_enqueue(LASER_POINT_COV, c.max_iteration, c.extrinsic_est,
                                                                    from_snapshot=True), ctx.synchronize()))
        timed("  end (map growth + report) + finish", lambda: (ctx.scan_step_end(c.filter_size_map, True),
                                                               r.adopt(ctx.scan_step_finish(r.x, r.P))))
for k, v in acc.items():
    print("%-48s %8.1f us" % (k, 1e6 * np.mean(v)))
print("(%d scans, N %d, M %d)" % (len(seq) - 14, len(m["lidar"]), mm))
