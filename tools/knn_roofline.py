#!/usr/bin/env python
"""kNN kernel against HBM: batched 5-NN (lio_knn5's kernel) on a map that does NOT fit L2.

    python tools/knn_roofline.py [--map-points 50000000] [--queries 4000000] [--reps 5]

BASELINE.json's north_star asks for the kNN kernel's achieved HBM bandwidth.  On the bench workload (2 M-point map =
32 MB) the map is L2-resident and the update is latency-bound, so DRAM counters say nothing there; this
tool builds the 50 M-point city map of config 5 (800 MB of points + 2 GB of hash table) and streams millions of
queries through the same search code.  Prints queries/s and the candidate-streaming model of SURVEY.md §8d
(bytes = queries x (16 + candidates x 16 + probes x 16)); run it under ncu for the measured dram__bytes."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from agi_lidar_slam_b200 import _cabi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=50_000_000)
    ap.add_argument("--queries", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch

    t0 = time.time()
    scene, mp = synth.city_map(a.map_points, 5005)
    print(f"map: {len(mp)} points, extent {np.ptp(mp[:, 0]):.0f} x {np.ptp(mp[:, 1]):.0f} m, generated in {time.time() - t0:.1f} s",
          flush=True)
    rng = np.random.default_rng(1)
    a.queries = min(a.queries, 1_000_000)  # one launch worth: the two query orders below use the same points
    q = mp[rng.integers(0, len(mp), a.queries)] + rng.normal(0, 0.15, (a.queries, 3)).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    # sorted: the same random points in voxel order (kz, ky, kx).  One query per ~7 m of surface: neighbours in the order
    # are still far apart, so this says little about a scan.
    cell = np.floor(q / 0.5).astype(np.int64)
    q_sorted = np.ascontiguousarray(q[np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))])
    chunk = min(a.queries, 1_000_000)  # queries per launch (they go through the scan-sized buffers)
    ctx = _cabi.Context(0, max_scan_points=1 << 18, max_down_points=chunk, max_map_points=int(len(mp) * 1.02))
    t0 = time.time()
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    ctx.synchronize()
    print(f"map built on the device in {time.time() - t0:.2f} s (includes the host-to-device copy)", flush=True)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # correctness spot check against brute force on a few queries
    idx, d2, _ = ctx.knn5(q[:64])
    for k in range(3):
        d = ((mp - q[k]) ** 2)
        dd = (d[:, 0] + d[:, 1]) + d[:, 2]
        o = np.argsort(dd, kind="stable")[:5]
        ok = dd[o] <= 5.0
        assert np.array_equal(idx[k][ok], o[ok].astype(np.int32)), (idx[k], o)
    found5 = float((idx[:, 4] >= 0).mean())
    # scans: what the update's searches look like -- OS1-128 scans taken at random poses in the same city, downsampled at
    # 0.5 m by the product's own voxel filter (voxel order), transformed to the world frame, back to back
    d, col = synth.spinning_dirs(128, 1024, -22.5, 22.5)
    ext = float(np.ptp(mp[:, 0])) / 2 - 150.0
    parts = []
    k = 0
    n_scan_q = 1 << 18  # queries of the `scans` launches (a raycast scan costs the host half a second)
    while sum(len(p) for p in parts) < n_scan_q:
        r2 = np.random.default_rng(100 + k)
        pos = np.array([r2.uniform(-ext, ext), r2.uniform(-ext, ext), 2.0])
        R = synth.rot_zyx(r2.uniform(-np.pi, np.pi), r2.normal(0, 0.02), r2.normal(0, 0.02))
        scan = synth.static_scan(scene, d, col / 1024 * 100.0, pos, R, 120.0, 200 + k)
        body, _, _ = ctx.scan_preprocess(scan, None, None, 0.5)
        parts.append((body[:, :3].astype(np.float64) @ R.T + pos).astype(np.float32))
        k += 1
    q_scans = np.ascontiguousarray(np.concatenate(parts)[:n_scan_q])
    print(f"scans: {k} OS1-128 scans, {len(q_scans)} downsampled points", flush=True)
    # per order: one lio_knn5 call that brings the queries to the device (1 launch) + reps + 1 resident launches
    for name, qq in (("random", q), ("sorted", q_sorted), ("scans", q_scans)):
        times = []
        chunk = len(qq)
        ctx._check(ctx._lib.lio_knn5(ctx._h, qq.ctypes.data, chunk, 5.0, None, None, None))
        for rep in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.knn5_resident(chunk)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            if rep:
                times.append(e0.elapsed_time(e1))
        ms = float(np.mean(times))
        print(f"[{name}] {chunk} queries per launch: {ms:.3f} ms -> {chunk / ms / 1e3:.1f} M queries/s; "
              f"algorithmic (116 B/query): {116.0 * chunk / (ms * 1e-3) / 1e9:.1f} GB/s")


if __name__ == "__main__":
    main()
