#!/usr/bin/env python
"""The preprocessing of the bench's OS1-128 scan a few times (a target for ncu captures of the voxel-filter kernels)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from agi_lidar_slam_b200 import _cabi, synth  # noqa: E402

cfg = synth.config3_os1_128(n_map=int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with _cabi.Context(0, max_scan_points=1 << 18, max_down_points=100000, max_map_points=1 << 16) as ctx:
    for _ in range(reps):
        m = ctx.scan_preprocess(cfg["scan"], None, None, 0.5, resident=True)
    print("N", len(cfg["scan"]), "M", m)
    import os
    if os.environ.get("LIO_TIMELINE") == "1":
        import torch
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for cold in (False, True):
            rows = []
            for _ in range(8):
                if cold:
                    flush.fill_(1)
                    torch.cuda.synchronize()
                ctx.scan_preprocess(cfg["scan"], None, None, 0.5, resident=True)
                ctx.debug_timeline()
                t = ctx.timeline_raw[200:208]
                rows.append([int(v + t[0]) for v in (t[5], t[6], t[1], t[2], t[3], t[4])])  # ns since the first block started
            print("cold" if cold else "warm", "centroid kernel: span, sorted, long, mid, short, kernel end [ns]:", np.median(rows, 0))
