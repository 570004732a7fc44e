#!/usr/bin/env python
"""bench.py — scans/s of the S-FAST_LIO IESKF update on synthetic OS1-128 scans against a 2M-point map.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2]): Ouster OS1-128 scan (128 x 1024 rays, ~1.3e5 points)
against a 2,000,000-point "city" map, 0.5 m surf voxel, up to 4 IESKF iterations, 5 cm / 1 deg perturbed prior.
A "step" is one whole update_iterated_dyn_share_modified of one scan (all passes: kNN + plane fit + Jacobian +
HtH/Hth reduction + 24x24 Kalman step).

`value`   scans/s with the downsampled scan, the map and the prior already resident in HBM; L2 is flushed before
          every timed step (the 2M-point map is 32 MB and would otherwise sit in the 126 MB L2).
`e2e`     the same metric through the C-ABI call a host program makes (lio_scan_upload + lio_update_scan) with
          pinned HOST buffers: scan + prior go host->device and the posterior comes back inside the timed region.
`roofline`the dominant kernel (fused search pass), algorithmic bytes 116 B x M per launch over its
          CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` the CPU path on this box's host cores: the reference's ikd-Tree (compiled in place, oracle/_ref) under
          the oracle's restated h_share_model / update loop, on a bounded sample of the same workload.
With N > 1 (torchrun) every rank replays its own scan sequence against its own replica of the map (independent
sequences, no collective in the data path; SURVEY.md §8e case 1): weak scaling, value = total scans / max-rank time.
`--workload sharded` runs the spatially sharded-map variant instead (x-slabs + NCCL all-reduce of the 92-double blob).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "scans/sec for IESKF update (kNN+plane+HtH) on OS1-128 scans"
UNIT = "scans/s"
R_COV = 0.001  # LASER_POINT_COV (laserMapping.cpp:29)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="os1_128_2m",
                    choices=["os1_128_2m", "sharded", "avia_200k", "vlp16_traj", "os1_64_seqs"],
                    help="os1_128_2m: config 3 (the headline), one independent sequence per GPU (no collective); sharded: "
                         "config 5, one map cut into x-slabs over the GPUs, 92-double blob exchanged per pass; avia_200k: "
                         "config 1, the same update on a 24k-ray Avia scan vs a 200k-point map; vlp16_traj: config 2, the "
                         "whole per-scan main loop (propagation, undistort + voxel filter, update, map growth) along a "
                         "VLP-16 trajectory; os1_64_seqs: config 4, the same loop for --seqs-per-gpu OS1-64 sequences "
                         "per GPU, their updates in one cooperative launch per scan")
    ap.add_argument("--seqs-per-gpu", type=int, default=8)
    ap.add_argument("--map-points", type=int, default=2_000_000)
    ap.add_argument("--rings", type=int, default=128)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--poses", type=int, default=4, help="distinct scan poses cycled through the steps (SURVEY.md 8d suggests "
                    "100: builder lines with --poses 32 are under profiles/; the default stays at round 1's 4 so that the "
                    "lines of the rounds compare)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sequences", default="2,4,8",
                    help="extra leg: independent sequences per cooperative launch (lio_update_enqueue_multi); '' = skip")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="sharded workload: blob exchange inside the persistent kernel over NVLink peer memory (peer) "
                         "or pass kernel -> NCCL all-reduce -> solve kernel per pass (nccl)")
    ap.add_argument("--host-threads", type=int, default=0, help="trajectory workloads: host threads of lio_seq_process_many "
                    "(0: min(8, sequences, cores / ranks - 1))")
    ap.add_argument("--sync-growth", action="store_true",
                    help="trajectory workloads: wait for each scan's map growth inside its step (default: deferred, the "
                         "step returns with the posterior as the reference publishes before map_incremental)")
    ap.add_argument("--legs", default="auto", help="extra legs of the default workload's line: 'auto' = sharded + os1_64_seqs "
                    "when N > 1, knn_hbm + dense_scene when N = 1; 'none'; or a comma list of those names")
    ap.add_argument("--shard-stripe", type=float, default=8.0, help="sharded map: width [m] of the x stripes dealt round-robin "
                    "to the ranks (lio_set_shard_stripes); 0 = one contiguous x slab per rank")
    ap.add_argument("--sharded-map-points", type=int, default=50_000_000, help="map size of the `sharded` extra leg")
    ap.add_argument("--e2e-steps", type=int, default=200, help="the e2e leg times max(--steps, this) steps and reports the "
                    "median step (one OS hiccup on one rank does not set an 8-rank number)")
    ap.add_argument("--map-cell", type=float, default=0.0, help="kNN hash cell edge [m] (0: library default); no effect on results")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------ workload
def workload_name(args, n_map):
    if args.workload == "avia_200k":
        return "Avia-like 24,000-ray scan vs %d-point hall map, leaf 0.5, max_iter 4, 1 sequence per GPU" % n_map
    return "OS1-128 %dx%d scan vs %d-point city map, leaf 0.5, max_iter 4, 1 sequence per GPU" % (args.rings, args.cols,
                                                                                                 n_map)


def make_workload(args, rank):
    from agi_lidar_slam_b200 import synth

    if args.workload == "avia_200k":  # config 1
        scene = synth.hall_scene(half=110.0, height=10.0)
        mp = synth.sample_map(scene, 200_000, 1001)
        scans = []
        for k in range(args.poses):
            rng = np.random.default_rng(1001 + 17 * k + 1)
            pos = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), 3.0])
            R = synth.rot_zyx(rng.uniform(-np.pi, np.pi), 0.45 + rng.normal(0, 0.02), rng.normal(0, 0.02))
            d = synth.avia_dirs(24_000, 1002 + k)
            scan = synth.static_scan(scene, d, np.arange(24_000) / 240000.0 * 1000.0, pos, R, 100.0,
                                     1001 + 1000 * rank + k)
            x_true = synth.make_state(pos=pos, R=R)
            scans.append(dict(scan=scan, x_true=x_true,
                              x_prior=synth.perturbed_prior(x_true, 1001 + 1000 * rank + 31 * k)))
        return dict(scene=scene, map=mp, scans=scans, P=synth.init_P(), leaf=0.5, max_iter=4, ext=False)
    scene, mp = synth.city_map(args.map_points, 3003)
    scans = []
    d, col = synth.spinning_dirs(args.rings, args.cols, -22.5, 22.5)
    tms = col / args.cols * 100.0
    for k in range(args.poses):
        # every rank replays sequences of the same difficulty: same poses, rank-specific range noise and prior error
        rng = np.random.default_rng(3003 + 17 * k + 1)
        pos = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), 2.0])
        R = synth.rot_zyx(rng.uniform(-np.pi, np.pi), rng.normal(0, 0.02), rng.normal(0, 0.02))
        scan = synth.static_scan(scene, d, tms, pos, R, 120.0, 3003 + 1000 * rank + k)
        x_true = synth.make_state(pos=pos, R=R)
        scans.append(dict(scan=scan, x_true=x_true, x_prior=synth.perturbed_prior(x_true, 3003 + 1000 * rank + 31 * k)))
    return dict(scene=scene, map=mp, scans=scans, P=synth.init_P(), leaf=0.5, max_iter=4, ext=False)


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe), sampled through NVML
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; polling the
    library avoids the pipe buffering of a child nvidia-smi, which loses short regions)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period_s=0.01):
        self.period = period_s
        self.sm, self.bits, self.power = [], [], []
        self.stop_flag = threading.Event()
        self.h = None
        self.err = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits.append(int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        seen = set()
        for b in self.bits:
            for bit, name in self.REASONS.items():
                if b & bit:
                    seen.add(name)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(seen), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None}


def ncu_traffic():
    """DRAM bytes per update_kernel launch from the committed ncu capture of this same workload (profiles/)."""
    p = ROOT / "profiles" / "r2_update_kernel_traffic.json"
    if not p.exists():
        p = ROOT / "profiles" / "r1_update_kernel_traffic.json"
    try:
        return float(json.loads(p.read_text())["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ CPU arm
class stdout_to_stderr:
    """The reference ikd-Tree printf()s thread start/stop notices; stdout carries exactly one JSON line, so whatever the
    CPU leg prints at the C level goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:  # noqa: BLE001
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_arm(wl, steps, warmup, seconds_budget, threads=None):
    """The reference CPU path: reference ikd-Tree (oracle/_ref) + restated update loop.  Returns a dict with scans/s."""
    with stdout_to_stderr():
        r = cpu_arm_inner(wl, steps, warmup, seconds_budget, threads)
        import gc
        gc.collect()  # the tree's destructor prints too
    return r


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return "%s x %d" % (line.split(":", 1)[1].strip(), os.cpu_count() or 1)
    except OSError:
        pass
    return "unknown x %d" % (os.cpu_count() or 1)


def cpu_arm_inner(wl, steps, warmup, seconds_budget, threads=None):
    from oracle import pyoracle as orc

    threads = threads or (os.cpu_count() or 1)
    orc.build()
    use_ikd = orc.ikd_available()
    t0 = time.perf_counter()
    if use_ikd:
        tree = orc.IkdTree()
        tree.build(wl["map"])
    else:
        tree = orc.Map(1.0)
        tree.build(wl["map"])
    build_s = time.perf_counter() - t0
    backend = tree.knn_backend()
    bodies = []
    t_vox = []
    for s in wl["scans"]:
        raw = s["scan"]
        pts5 = np.concatenate([raw[:, :3], np.zeros((len(raw), 1), np.float32), raw[:, 3:4]], 1)
        t0 = time.perf_counter()
        cen = orc.voxel_grid(pts5, wl["leaf"])[0]
        t_vox.append(time.perf_counter() - t0)
        bodies.append(np.ascontiguousarray(cen[:, :3]))
    # one search pass alone (Nearest_Search of every downsampled point at the prior, all threads): SURVEY.md 8d asks for
    # the stages of the CPU path, not only their sum
    t_knn = []
    for j in range(min(len(bodies), 4)):
        q = orc.body_to_world(wl["scans"][j]["x_prior"], bodies[j])
        t0 = time.perf_counter()
        tree.knn(q, 5, threads=threads) if use_ikd else tree.knn(q, 5, 5.0, threads=threads)
        t_knn.append(time.perf_counter() - t0)
    times, nvalid, npass = [], [], []
    k = 0
    t_start = time.perf_counter()
    while True:
        s = wl["scans"][k % len(bodies)]
        sc = orc.Scan(bodies[k % len(bodies)])
        t0 = time.perf_counter()
        x, P, trace, nv = sc.update(s["x_prior"], wl["P"], backend, R_COV, wl["max_iter"], wl["ext"], threads=threads)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
            nvalid.append(nv)
            npass.append(len(trace))
        k += 1
        if len(times) >= steps or (len(times) >= 3 and time.perf_counter() - t_start > seconds_budget):
            break
    tot = float(np.sum(times))
    return dict(value=len(times) / tot, ms_per_step=1000 * tot / len(times), steps=len(times), cores=threads,
                kind="reference" if use_ikd else "port",
                kind_detail=("reference ikd-Tree (oracle/_ref, compiled in place) for Build/Nearest_Search + restated "
                             "h_share_model/update loop (Eigen/PCL/Sophus absent)") if use_ikd else
                "oracle port (hashed-grid kNN + restated update loop)", build_s=build_s,
                matched_pts_per_s=float(np.sum(nvalid)) / tot, m=int(np.mean([len(b) for b in bodies])),
                passes=float(np.mean(npass)), cpu_model=cpu_model(),
                stages_ms={"voxel_grid_1_thread": 1000 * float(np.mean(t_vox)),
                           "nearest_search_one_pass": 1000 * float(np.mean(t_knn)),
                           "update_all_passes": 1000 * tot / len(times),
                           "note": "update_all_passes = searches (about half of the passes search again) + plane fits + "
                                   "Jacobian rows + HtH/Hth products + 24x24 steps; voxel_grid is the oracle's restated "
                                   "pcl::VoxelGrid (single-threaded, as PCL's)"})


# ------------------------------------------------------------------------------------------ trajectories (configs 2, 4)
TRAJ = {"vlp16_traj": dict(rings=16, cols=1800, fov=(-15.0, 15.0), max_range=100.0, seed=2002, sensor="VLP-16"),
        "os1_64_seqs": dict(rings=64, cols=1024, fov=(-22.5, 22.5), max_range=120.0, seed=4000, sensor="OS1-64")}
LEAD = 4  # scans the main loop spends on first-scan bookkeeping, IMU initialisation and the first-scan map build


def traj_sequences(args, rank, n_seq, n_scans):
    from agi_lidar_slam_b200 import synth

    t = TRAJ[args.workload]
    return [synth.sequence(n_scans, t["seed"] + rank * n_seq + k, rings=t["rings"], cols=t["cols"], fov=t["fov"],
                           max_range=t["max_range"]) for k in range(n_seq)]


def traj_cpu_arm(seq, warmup, steps, seconds_budget, threads=None):
    """The same main loop on the CPU oracle (tests/replay_oracle.py: restated stages, hashed-grid map port)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import pyoracle as orc
    from replay_oracle import OracleReplay

    orc.build()
    threads = threads or (os.cpu_count() or 1)
    rep = OracleReplay(orc, max_iteration=3, threads=threads)
    t_sum, n = 0.0, 0
    t_start = time.perf_counter()
    for j, m in enumerate(seq):
        t0 = time.perf_counter()
        r = rep.process(m)
        dt = time.perf_counter() - t0
        if j >= LEAD + warmup and r is not None:
            t_sum += dt
            n += 1
        if n >= steps or (n >= 3 and time.perf_counter() - t_start > seconds_budget):
            break
    return dict(value=n / t_sum, ms_per_step=1000 * t_sum / n, steps=n, cores=threads, kind="port",
                kind_detail="oracle port of the whole main loop (restated IMU propagation, undistort, VoxelGrid, update, "
                            "map_incremental on the hashed-grid map port)")


def traj_main(args, rank, world, local, dev, torch, dist, _cabi):
    line = traj_leg(args, rank, world, local, dev, torch, dist, _cabi, cpu_baseline=not args.no_cpu_baseline)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def traj_leg(args, rank, world, local, dev, torch, dist, _cabi, cpu_baseline=False):
    """Configs 2 and 4: one step = one scan of every sequence through the whole main loop of laserMapping.cpp:702-800.
    The raw scan and the IMU samples come from host memory every step by nature, so `value` is host-timed end to end
    (= `e2e`); max over ranks; the sequences of a GPU share one cooperative update launch per scan."""
    t = TRAJ[args.workload]
    n_seq = 1 if args.workload == "vlp16_traj" else max(1, args.seqs_per_gpu)
    n_scans = LEAD + args.warmup + args.steps
    seqs = traj_sequences(args, rank, n_seq, n_scans)
    ctxs = [_cabi.Context(local, max_scan_points=1 << 17, max_down_points=1 << 16, max_map_points=1 << 21)
            for _ in range(n_seq)]
    # the native main loop (lio_seq_process / lio_seq_process_many): one C-ABI call per step
    for c in ctxs:
        c.set_deferred_growth(not args.sync_growth)
    # one core per rank stays free for the rank's main / NCCL / sampler threads: with every core in an OpenMP team the
    # spinning teams starve each other (8 ranks on 32 cores: 4 threads per rank 32.7 k scans/s, 3 threads 48.1 k)
    host_threads = args.host_threads or max(1, min(8, n_seq, (os.cpu_count() or 1) // max(1, world) - 1))
    _cabi.set_host_threads(host_threads)
    runs = [_cabi.Sequence(c, max_iteration=3) for c in ctxs]
    # scans wait in pinned host memory (what a driver's receive buffer is), as the e2e rule of the bench contract asks
    for s in seqs:
        for m in s:
            pinned = torch.from_numpy(np.ascontiguousarray(m["lidar"], np.float32)).pin_memory()
            m["lidar_pinned"] = pinned  # keeps the allocation alive
            m["lidar"] = pinned.numpy()
    inputs = [[runs[k].input(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"]) for m in seqs[k]]
              for k in range(n_seq)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    launches0 = 0
    results = []
    res1 = _cabi.SeqResult()
    lib = _cabi.load_library()
    import ctypes
    for j in range(n_scans):
        if j == LEAD + args.warmup:
            barrier()
            sampler.start()
            launches0 = sum(c.launch_count for c in ctxs)
            t0 = time.perf_counter()
        if n_seq == 1:
            rc = lib.lio_seq_process(runs[0]._h, ctypes.byref(inputs[0][j]), ctypes.byref(res1))
            if rc:
                raise _cabi.LioError(rc, "lio_seq_process")
            if j >= LEAD + args.warmup:
                results.append((res1.status, res1.m, res1.n_valid, res1.n_passes))
        else:
            rs = _cabi.seq_process_many(runs, [inputs[k][j] for k in range(n_seq)])
            if j >= LEAD + args.warmup:
                results += [(r.status, r.m, r.n_valid, r.n_passes) for r in rs]
    for c in ctxs:
        c.scan_step_settle()  # the last scan's map growth belongs to the timed region
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    launches = sum(c.launch_count for c in ctxs) - launches0
    barrier()
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    ok = [dict(m=r[1], n_valid=r[2], n_passes=r[3]) for r in results if r[0] == _cabi.SEQ_UPDATED]
    n_upd = len(ok)
    M = float(np.mean([e["m"] for e in ok]))
    passes = float(np.mean([e["n_passes"] for e in ok]))
    n_raw = float(np.mean([len(m["lidar"]) for s in seqs for m in s[-args.steps:]]))
    value = world * n_seq * args.steps / dt
    b_scan = 16.0 * n_raw + 16.0 * M + 116.0 * M * passes
    peak, peak_src = measured_peak()
    achieved = b_scan * (n_seq * args.steps / dt) / 1e9
    n_imu = float(np.mean([len(m["imu"]) for s in seqs for m in s[-args.steps:]]))
    line = {
        "metric": "scans/sec through the whole per-scan main loop (propagation + undistort + voxel + IESKF update + map growth)",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": "%s %dx%d trajectory (10 Hz scans, 200 Hz IMU), %d sequence(s) per GPU, growing map, leaf "
                   "0.5/0.5, max_iter 3" % (t["sensor"], t["rings"], t["cols"], n_seq), "N_raw": n_raw, "M": M,
                   "passes_per_scan": passes, "updates_in_timed_region": n_upd,
                   "l2": "not flushed: every scan is new data from the host (pinned memory), the map is the sequence's own "
                   "growing map",
                   "timing": "host clock around the loop (host stages are on the path), device synchronised both sides",
                   "host_threads": host_threads,
                   "map_growth": "inside the step" if args.sync_growth else
                   "deferred: runs under the next scan's host stage and upload (lio_set_deferred_growth)"},
        "matched_pts_per_s": float(np.sum([e["n_valid"] for e in ok])) * world / dt,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(n_seq * (16 * n_raw + 200 * (n_imu + 1) + 602 * 8)),
                "d2h_bytes_per_step": int(n_seq * (607 * 8 + 8 + 12))},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "whole scan: undistort/voxel kernels + update_kernel + map growth",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_scan": b_scan,
                     "note": "B_scan = 16N + 16M + 116*M*I (SURVEY.md §8d); latency- and host-bound"},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and cpu_baseline:
        r = traj_cpu_arm(seqs[0], args.warmup, args.steps, args.cpu_seconds)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                "kind_detail": r["kind_detail"],
                                "sample": "%d scans of sequence 0 through the same loop" % r["steps"]}
        r3 = traj_cpu_arm(seqs[0], args.warmup, args.steps, min(8.0, args.cpu_seconds), threads=3)
        line["cpu_baseline_3_threads"] = {"value": r3["value"], "unit": UNIT, "cores": 3}
    for r in runs:
        r.close()
    for c in ctxs:
        c.close()
    return line


# ------------------------------------------------------------------------------------------ sharded map (config 5)
def sharded_main(args, rank, world, local, dev, torch, dist, _cabi):
    line = sharded_leg(args, rank, world, local, dev, torch, dist, _cabi)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _world_x(body, x):
    """x coordinate of p_world = rot * (R_LI * p + t_LI) + pos for the points of a downsampled scan (host, FP64)."""
    def rot(q):
        w, a, b, c = q
        return np.array([[1 - 2 * (b * b + c * c), 2 * (a * b - w * c), 2 * (a * c + w * b)],
                         [2 * (a * b + w * c), 1 - 2 * (a * a + c * c), 2 * (b * c - w * a)],
                         [2 * (a * c - w * b), 2 * (b * c + w * a), 1 - 2 * (a * a + b * b)]])
    x = np.asarray(x, np.float64)
    pi = np.asarray(body[:, :3], np.float64) @ rot(x[7:11]).T + x[11:14]
    return (pi @ rot(x[3:7]).T + x[0:3])[:, 0]


def sharded_leg(args, rank, world, local, dev, torch, dist, _cabi):
    """One map in x-slabs over the ranks; every rank runs every pass on the whole scan against its slab (+ halo) and
    contributes the rows of the queries it owns; one NCCL all-reduce of 92 doubles per pass; identical Kalman step on
    every rank.  Strong scaling: the same scans, whatever the number of GPUs."""
    from agi_lidar_slam_b200 import sharded

    wl = make_workload(args, 0)  # the SAME map and scans on every rank
    mp = wl["map"]
    bounds = sharded.slab_bounds(mp[:, 0], world)
    stripe = float(args.shard_stripe) if world > 1 else 0.0
    x_org = float(np.floor(mp[:, 0].min()))
    if stripe > 0:
        keep = sharded.stripe_indices(mp[:, 0], x_org, stripe, world, rank)
    else:
        keep = sharded.shard_indices(mp[:, 0], bounds, rank)
    ctx = _cabi.Context(local, max_scan_points=max(1 << 18, args.rings * args.cols), max_down_points=150000,
                        max_map_points=max(1 << 20, int(len(keep) * 1.05)))
    if stripe > 0:
        ctx.set_shard_stripes(x_org, stripe, world, rank)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_build(np.concatenate([mp[keep], np.zeros((len(keep), 1), np.float32)], 1))
    bodies, dense_bodies = [], []
    for s in wl["scans"]:
        body, _, _ = ctx.scan_preprocess(s["scan"], None, None, wl["leaf"])
        bodies.append(np.ascontiguousarray(body))
        body, _, _ = ctx.scan_preprocess(s["scan"], None, None, 0.15)  # the dense-scene variant (M ~ 40k)
        dense_bodies.append(np.ascontiguousarray(body))
    sparse_bodies = bodies
    M = int(np.mean([len(b) for b in bodies]))
    fused = world > 1 and args.exchange == "peer"
    if fused:
        sharded.connect_peers(ctx, rank, world)
        reduce = None
    elif world > 1:
        reduce, blob_t = sharded.nccl_reduce(ctx, dev)
    else:
        reduce = lambda: None  # noqa: E731
    own = (float(bounds[rank]), float(bounds[rank + 1]))
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # pinned copies: the uploads of a step are then truly asynchronous and the host runs ahead of the stream
    pin = lambda b: torch.from_numpy(b).pin_memory()  # noqa: E731
    keepalive = [[pin(b) for b in sparse_bodies], [pin(b) for b in dense_bodies]]
    sparse_bodies = bodies = [t.numpy() for t in keepalive[0]]
    dense_bodies = [t.numpy() for t in keepalive[1]]
    passes_of = {}

    def run(steps, warmup, bodies=None, aligned=False, queued=True):
        """Timed steps.  queued: nothing is read back between the steps, so the host runs ahead and every rank's stream
        holds the next step's kernel when the current one ends -- the ranks stay aligned through the exchange itself
        instead of paying the host's launch jitter (tens of microseconds at 8 processes) in every step.  The steps of the
        warm-up are read back one by one (pass counts per scan).  aligned: diagnostic, NCCL barrier before each step."""
        bodies = sparse_bodies if bodies is None else bodies
        evs, nvalid, npass, x = [], 0, 0, None
        for k in range(warmup + steps):
            j = k % len(bodies)
            ctx.scan_upload(bodies[j])
            ctx.state_upload(wl["scans"][j]["x_prior"], wl["P"])
            flush.fill_(1)
            if aligned and world > 1:
                dist.barrier()  # an NCCL kernel on this stream: the update kernels of all ranks start together
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if fused:
                ctx.update_enqueue_sharded(R_COV, wl["max_iter"], wl["ext"], True, own[0], own[1])
            else:
                ctx.update_begin(wl["max_iter"], wl["ext"], True)
                for _ in range(wl["max_iter"] + 1):
                    ctx.update_pass_enqueue(wl["ext"], own[0], own[1])
                    reduce()
                    ctx.update_step_enqueue(R_COV, wl["ext"])
            e1.record(stream)
            if k < warmup or not queued:
                x, P, nv, npz = ctx.state_download()
                passes_of[(id(bodies), j)] = (nv, npz)
            if k >= warmup:
                evs.append((e0, e1))
                nv, npz = passes_of.get((id(bodies), j), (0, 0))
                nvalid += nv
                npass += npz
        if queued and steps > 0:
            x, P, _, _ = ctx.state_download()  # the last step's posterior; synchronises
        torch.cuda.synchronize(dev)
        return sum(a.elapsed_time(b) for a, b in evs), nvalid, npass, x

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    run(0, max(len(bodies), 3, args.warmup))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count
    ms, nvalid, npass, x_last = run(args.steps, 0)
    launches = ctx.launch_count - l0
    barrier()
    # The sharded update is a collective: its kernels wait for each other in every pass, so a rank that starts late is
    # paid for by all of them.  `value` times it the way collectives are timed -- the ranks' streams aligned by an NCCL
    # barrier kernel in front of every step's first event (outside the timed bracket), as when one scan reaches all ranks
    # at once; the unaligned figure (every rank launches when its own host and its own L2 flush get there: +40 us of start
    # skew per step at 8 processes) is filed beside it.
    ums = ms
    ms, nvalid, npass, x_last = run(args.steps, 0, aligned=True)
    barrier()
    clocks = sampler.stop()
    # the same on the dense variant of the scans: enough rows per block that a rank's share of them matters
    dsteps = min(args.steps, 20)
    run(0, max(3, len(dense_bodies)), dense_bodies)
    barrier()
    dms, _, dnpass, _ = run(dsteps, 0, dense_bodies, aligned=True)
    barrier()
    if fused and ctx.peer_timed_out():
        raise SystemExit("peer exchange timed out")
    if world > 1:
        t = torch.tensor([ms, dms, ums], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, dms, ums = float(t[0].item()), float(t[1].item()), float(t[2].item())
        xs = torch.tensor(x_last, dtype=torch.float64, device=dev)
        x0 = xs.clone()
        dist.broadcast(x0, 0)
        same = bool(torch.equal(xs, x0))
    else:
        same = True
    # how the scan splits over the ranks (points whose p_world.x lies in the rank's window, at the last posterior)
    j_last = (args.steps - 1) % len(bodies)  # run(steps, 0): step k replays scan k % len(bodies)
    pw_x = _world_x(bodies[j_last], x_last)
    if stripe > 0:
        owned = int((np.mod(sharded.stripe_of(pw_x, x_org, stripe), world) == rank).sum())
    else:
        owned = int(((pw_x >= own[0]) & (pw_x < own[1])).sum())
    owned_max = owned
    vs_single = None
    if world > 1:
        t = torch.tensor([owned], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        owned_max = int(t.item())
        if rank == 0:
            # the same scan against the WHOLE map in one context on this GPU: what a single GPU would have answered
            c1 = _cabi.Context(local, max_scan_points=1 << 12, max_down_points=100000,
                               max_map_points=max(1 << 20, int(len(mp) * 1.02)))
            c1.set_stream(stream.cuda_stream)
            c1.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
            c1.scan_upload(bodies[j_last])
            c1.state_upload(wl["scans"][j_last]["x_prior"], wl["P"])
            c1.update_enqueue(R_COV, wl["max_iter"], wl["ext"], from_snapshot=True)
            x1, P1, nv1, np1 = c1.state_download()
            c1.close()
            vs_single = {"max_abs_state_diff": float(np.abs(np.asarray(x_last) - np.asarray(x1)).max()),
                         "pos_diff_m": float(np.abs(np.asarray(x_last)[0:3] - np.asarray(x1)[0:3]).max()),
                         "passes_single": int(np1)}
        barrier()
    peak, peak_src = measured_peak()
    passes = npass / args.steps
    alg = 116.0 * M * passes
    line = {
        "metric": METRIC, "value": args.steps / (ms / 1000.0), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": "sharded map: %d-point city map in %s (+ sqrt(5)+0.5 m halo), OS1-128 %dx%d scans, "
                   "%s" % (len(mp), ("%g m x-stripes dealt round-robin to %d ranks" % (stripe, world)) if stripe > 0 else
                           "%d x-slabs" % world, args.rings, args.cols,
                           "blobs exchanged inside the persistent kernel over NVLink peer memory" if fused else
                           "NCCL all-reduce of 92 doubles per pass between pass and solve kernels"), "M": M,
                   "passes_per_scan": passes, "local_map_points": int(len(keep)), "owned_points_max_rank": owned_max,
                   "scan_points_last": int(len(bodies[j_last])),
                   "value_unaligned": args.steps / (ums / 1000.0),
                   "timing": "CUDA events around every step's update on its stream, max over ranks of their sum; an NCCL barrier "
                             "kernel on the same stream in front of every step's first event aligns the ranks' starts (the "
                             "update is a collective); value_unaligned: without it",
                   "dense": {"value": dsteps / (dms / 1000.0), "unit": UNIT, "ms_per_step": dms / dsteps, "steps": dsteps,
                             "M": int(np.mean([len(b) for b in dense_bodies])), "passes_per_scan": dnpass / dsteps,
                             "what": "the same scans downsampled at 0.15 m"},
                   "vs_single_gpu": vs_single,
                   "l2": "flushed (384 MiB write) before every timed step", "states_identical_across_ranks": same},
        "matched_pts_per_s": nvalid / (ms / 1000.0), "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "update_kernel with in-kernel peer exchange" if fused else
                     "pass_kernel + solve_kernel per pass (host-driven, all-reduce between)",
                     "achieved": alg / (ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms / args.steps * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
        "clocks": clocks,
    }
    ctx.close()
    return line



# ------------------------------------------------------------------------------------------ extra legs at N = 1
def cell_count_grid(mp, cell=1.5):
    """Points per kNN cell of a map as a dense array over its bounding grid (None when that would not fit)."""
    try:
        inv = np.float32(1.0 / cell)
        ck = np.floor(mp[:, :3] * inv).astype(np.int32)
        lo = ck.min(0) - 1
        dim = (ck.max(0) + 2 - lo).astype(np.int64)
        ncell = int(dim[0] * dim[1] * dim[2])
        if ncell > 400_000_000:
            return None
        lin = (ck[:, 0] - lo[0]).astype(np.int64) + dim[0] * ((ck[:, 1] - lo[1]).astype(np.int64) + dim[1] * (ck[:, 2] - lo[2]).astype(np.int64))
        return {"counts": np.bincount(lin, minlength=ncell).astype(np.int32), "lo": lo, "dim": dim, "inv": inv, "cell": cell}
    except (MemoryError, ValueError):
        return None


def knn_stream_model(grid, queries, sample=20000):
    """SURVEY.md 8(d)'s candidate-streaming model next to the measured bytes: B_knn_stream = 16 + 16 C + 20 bytes per query,
    C = mean number of candidates a query examines = the points in the 27 cells of its 3x3x3 block (host arithmetic on the
    same map and a sample of the same queries)."""
    if grid is None:
        return None
    counts, lo, dim = grid["counts"], grid["lo"], grid["dim"]
    qs = queries[:: max(1, len(queries) // sample)]
    qc = np.floor(qs[:, :3] * grid["inv"]).astype(np.int64) - lo
    qc = qc[np.all((qc >= 1) & (qc < dim - 1), axis=1)]
    if len(qc) == 0:
        return None
    cand = np.zeros(len(qc), np.int64)
    occ = np.zeros(len(qc), np.int64)
    sect = np.zeros(len(qc), np.int64)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                c = counts[(qc[:, 0] + dx) + dim[0] * ((qc[:, 1] + dy) + dim[1] * (qc[:, 2] + dz))].astype(np.int64)
                cand += c
                occ += c > 0
                sect += np.where(c > 0, (c * 16 + 63) // 64 + 1, 0)  # 64-byte DRAM bursts a bucket touches (unaligned: one more)
    C = float(cand.mean())
    return {"cell_m": grid["cell"], "queries_sampled": int(len(qc)), "candidates_per_query_mean": C,
            "occupied_cells_of_27_mean": float(occ.mean()), "B_knn_stream_bytes_per_query": 16 + 16 * C + 20,
            "B_bursts_bytes_per_query": float(27 * 64 + 64 * sect.mean()),
            "what": "B_knn_stream = 16 (query) + 16 C (candidates streamed) + 20 (ids out), C = points in the query's 27 cells; "
                    "B_bursts = what a query without any sharing must move at the 64-byte granularity of a DRAM burst: 27 hash "
                    "probes (16 bytes used of each) + the bursts its occupied buckets touch"}


def knn_hbm_leg(args, local, dev, torch, _cabi):
    """The search kernel against HBM (north_star: achieved bandwidth of the kNN kernel): batched 5-NN on a 50M-point map
    (800 MB of points + the cell table: far beyond L2), 1M queries per launch, CUDA events.  queries/s is measured here;
    DRAM bytes per query come from the committed ncu capture of the same kernel on the same map and query order."""
    from agi_lidar_slam_b200 import synth

    n_map, nq = args.sharded_map_points, 1_000_000
    scene, mp = synth.city_map(n_map, 5005)
    rng = np.random.default_rng(1)
    # random: map points + noise in random order, no locality at all (the worst case)
    sel = rng.integers(0, len(mp), nq)
    q = np.ascontiguousarray(mp[sel] + rng.normal(0, 0.15, (nq, 3)).astype(np.float32), np.float32)
    ctx = _cabi.Context(local, max_scan_points=1 << 18, max_down_points=nq, max_map_points=int(len(mp) * 1.02))
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    ctx.synchronize()
    out = {"map_points": int(len(mp)), "queries_per_launch": nq, "unit": "queries/s",
           "algorithmic_bytes_per_query": 116}
    cap = {}
    f = ROOT / "profiles" / "r2_knn_hbm.json"
    if f.exists():
        cap = json.loads(f.read_text())
    peak, peak_src = measured_peak()
    # scans: what the update's searches look like -- OS1-128 scans at random poses in the same city, downsampled at 0.5 m
    # by the product's own voxel filter (voxel order), in the world frame, back to back until the launch is full
    d, col = synth.spinning_dirs(128, 1024, -22.5, 22.5)
    ext = float(np.ptp(mp[:, 0])) / 2 - 150.0
    parts, k = [], 0
    nqs = 1 << 18  # queries of the `scans` launches (a raycast scan costs the host half a second)
    while sum(len(p) for p in parts) < nqs:
        r2 = np.random.default_rng(100 + k)
        pos = np.array([r2.uniform(-ext, ext), r2.uniform(-ext, ext), 2.0])
        R = synth.rot_zyx(r2.uniform(-np.pi, np.pi), r2.normal(0, 0.02), r2.normal(0, 0.02))
        scan = synth.static_scan(scene, d, col / 1024 * 100.0, pos, R, 120.0, 200 + k)
        body, _, _ = ctx.scan_preprocess(scan, None, None, 0.5)
        parts.append((body[:, :3].astype(np.float64) @ R.T + pos).astype(np.float32))
        k += 1
    q_scans = np.ascontiguousarray(np.concatenate(parts)[:nqs])
    out["scans_in_launch"] = k
    grid = cell_count_grid(mp)
    for name, qq in (("random", q), ("scans", q_scans)):
        n = len(qq)
        ctx._check(ctx._lib.lio_knn5(ctx._h, qq.ctypes.data, n, 5.0, None, None, None))  # queries -> device
        ts = []
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.knn5_resident(n)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[1:]))
        leg = {"queries_per_launch": n, "ms_per_launch": ms, "queries_per_s": n / (ms * 1e-3)}
        c = cap.get(name)
        if c:  # DRAM bytes of the ncu capture applied to the duration measured here
            leg["dram_bytes_per_query"] = c["dram_bytes_per_launch"] / c["queries_per_launch"]
            leg["dram_GBps"] = leg["dram_bytes_per_query"] * n / (ms * 1e-3) / 1e9
            leg["frac_of_hbm_peak"] = leg["dram_GBps"] / peak
            leg["ncu_ms_per_launch"] = c["ms_per_launch"]
            leg["lts_hit_pct"], leg["l1tex_hit_pct"] = c.get("lts_hit_pct"), c.get("l1tex_hit_pct")
        m = knn_stream_model(grid, qq)
        if m:
            leg["stream_model"] = m
            if "dram_bytes_per_query" in leg:
                leg["dram_bytes_over_model"] = leg["dram_bytes_per_query"] / m["B_knn_stream_bytes_per_query"]
        out[name] = leg
    out["traffic_source"] = "profiles/r2_knn_hbm.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of knn_batch_kernel)" if cap else None
    out["peak"], out["peak_source"] = peak, peak_src
    ctx.close()
    return out


def dense_leg(args, wl, map4, local, dev, torch, _cabi, stream):
    """The same update on a dense scan: the OS1-128 scans downsampled at 0.15 m instead of 0.5 m (M ~ 40k, the size
    SURVEY.md 8d's worked example assumes): multi-tile chunks and 8-lane search groups.  Device-timed, L2 flushed."""
    n_map = len(map4)
    ctx = _cabi.Context(local, max_scan_points=max(1 << 18, args.rings * args.cols), max_down_points=150000,
                        max_map_points=max(1 << 21, int(n_map * 1.05)))
    ctx.set_stream(stream.cuda_stream)
    ctx.map_build(map4)
    bodies = [np.ascontiguousarray(ctx.scan_preprocess(s["scan"], None, None, 0.15)[0]) for s in wl["scans"]]
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)
    evs, npass, nvalid = [], 0, 0
    steps = min(args.steps, 20)
    for k in range(3 + steps):
        j = k % len(bodies)
        ctx.scan_upload(bodies[j])
        ctx.state_upload(wl["scans"][j]["x_prior"], wl["P"])
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.update_enqueue(R_COV, wl["max_iter"], wl["ext"], from_snapshot=True)
        e1.record(stream)
        x, P, nv, npz = ctx.state_download()
        if k >= 3:
            evs.append((e0, e1))
            npass += npz
            nvalid += nv
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    M = int(np.mean([len(b) for b in bodies]))
    passes = npass / steps
    peak, _ = measured_peak()
    ctx.close()
    return {"value": 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "M": M, "passes_per_scan": passes,
            "matched_pts_per_s": nvalid / steps / (ms * 1e-3), "surf_leaf": 0.15,
            "roofline_frac": 116.0 * M * passes / (ms * 1e-3) / 1e9 / peak,
            "what": "update of the same OS1-128 scans downsampled at 0.15 m (dense scene), device-resident, cold L2"}


# ------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    rank, world, local = dist_env()
    n_gpus = args.gpus
    if world > 1:
        n_gpus = world

    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.workload in TRAJ:
            seq = traj_sequences(args, 0, 1, LEAD + args.warmup + args.steps)[0]
            r = traj_cpu_arm(seq, args.warmup, args.steps, seconds_budget=150.0)
            print(json.dumps({
                "impl": "reference", "metric": "scans/sec through the whole per-scan main loop (propagation + undistort + "
                "voxel + IESKF update + map growth)", "value": r["value"], "unit": UNIT, "n_gpus": n_gpus,
                "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": {"workload": "%s trajectory, 1 sequence, CPU" % TRAJ[args.workload]["sensor"]},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "kind_detail": r["kind_detail"], "sample": "%d scans" % r["steps"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        wl = make_workload(args, 0)
        r = cpu_arm(wl, args.steps, args.warmup, seconds_budget=150.0)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": n_gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": workload_name(args, len(wl["map"])), "M": r["m"], "passes_per_scan": r["passes"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                             "kind_detail": r["kind_detail"], "cpu_model": r["cpu_model"], "stages_ms": r["stages_ms"],
                             "sample": "%d whole updates of the bench scans (tree build %.1f s untimed)" %
                             (r["steps"], r["build_s"])},
            "matched_pts_per_s": r["matched_pts_per_s"],
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    from agi_lidar_slam_b200 import _cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    if args.workload == "sharded":
        return sharded_main(args, rank, world, local, dev, torch, dist, _cabi)
    if args.workload in TRAJ:
        return traj_main(args, rank, world, local, dev, torch, dist, _cabi)

    wl = make_workload(args, rank)
    n_map = len(wl["map"])
    extra = {"map_cell": args.map_cell} if args.map_cell > 0 else {}
    ctx = _cabi.Context(local, max_scan_points=max(1 << 18, args.rings * args.cols), max_down_points=100000,
                        max_map_points=max(1 << 21, int(n_map * 1.05)), **extra)
    stream = torch.cuda.Stream(dev)  # a real (non-legacy) stream: events, L2 flush and our kernels all run on it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    map4 = np.concatenate([wl["map"], np.zeros((n_map, 1), np.float32)], 1)
    t0 = time.perf_counter()
    ctx.map_build(map4)
    ctx.synchronize()
    map_build_s = time.perf_counter() - t0

    # downsample every scan on the device once (lio_scan_preprocess) and keep host copies for the e2e leg
    bodies, raws = [], []
    for s in wl["scans"]:
        body, _, _ = ctx.scan_preprocess(s["scan"], None, None, wl["leaf"])
        bodies.append(torch.from_numpy(np.ascontiguousarray(body)).pin_memory())
        raws.append(torch.from_numpy(np.ascontiguousarray(s["scan"])).pin_memory())
    M = int(np.mean([b.shape[0] for b in bodies]))
    P0 = wl["P"]

    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def l2_flush():
        flush.fill_(1)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------------------------------------------------------- value: device-resident update, cold L2
    def resident_run(steps, warmup, cold):
        total_ms, nvalid, npass = 0.0, 0, 0
        evs = []
        for k in range(warmup + steps):
            j = k % len(bodies)
            ctx.scan_upload(bodies[j].numpy())
            ctx.state_upload(wl["scans"][j]["x_prior"], P0)
            if cold:
                l2_flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.update_enqueue(R_COV, wl["max_iter"], wl["ext"], from_snapshot=True)
            e1.record(stream)
            if k >= warmup:
                evs.append((e0, e1))
                x, P, nv, npz = ctx.state_download()
                nvalid += nv
                npass += npz
        torch.cuda.synchronize(dev)
        per_step = [a.elapsed_time(b) for a, b in evs]
        resident_run.median_ms = float(np.median(per_step)) if per_step else None
        total_ms = sum(per_step)
        return total_ms, nvalid, npass

    resident_run(0, max(3, args.warmup), True)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ms_cold, nvalid, npass = resident_run(args.steps, 0, True)
    med_cold = max_over_ranks(resident_run.median_ms)
    launches = ctx.launch_count - launches0
    barrier()
    ms_cold = max_over_ranks(ms_cold)
    ms_warm, _, _ = resident_run(args.steps, 0, False)
    med_warm = max_over_ranks(resident_run.median_ms)
    barrier()
    ms_warm = max_over_ranks(ms_warm)
    value = n_gpus * args.steps / (ms_cold / 1000.0)
    value_warm = n_gpus * args.steps / (ms_warm / 1000.0)
    matched = sum_over_ranks(float(nvalid)) / (ms_cold / 1000.0)
    passes_per_scan = npass / args.steps

    # ---------------------------------------------------------------- e2e: host buffers through the C-ABI
    def e2e_run(steps, warmup):
        ts = []
        for k in range(warmup + steps):
            j = k % len(bodies)
            l2_flush()
            torch.cuda.synchronize(dev)
            np.copyto(x_io, wl["scans"][j]["x_prior"])
            np.copyto(P_io, P0)
            t0 = time.perf_counter()
            # one C-ABI call: H2D of the scan (M x 16 B, pinned), prior in the kernel parameters, one kernel, posterior
            # written by the kernel into mapped pinned memory
            ctx.update_scan_host(bodies_np[j], x_io, P_io, R_COV, wl["max_iter"], wl["ext"])
            dt = time.perf_counter() - t0  # returns after the D2H of the posterior (host sync)
            if k >= warmup:
                ts.append(dt * 1000.0)
        return np.asarray(ts)

    bodies_np = [b.numpy() for b in bodies]  # views of the pinned tensors
    x_io, P_io = np.zeros(26), np.zeros((24, 24))
    e2e_steps = max(args.steps, args.e2e_steps)
    e2e_run(0, 3)
    barrier()
    e2e_ts = e2e_run(e2e_steps, 0)
    barrier()
    # the step a rank typically takes (median over >= 200 steps), max over ranks; the mean over the same window beside it
    e2e_med_ms = max_over_ranks(float(np.median(e2e_ts)))
    e2e_mean_ms = max_over_ranks(float(np.mean(e2e_ts)))
    e2e_value = n_gpus / (e2e_med_ms / 1000.0)
    clocks = sampler.stop()  # sampled over both timed regions (device-resident and end-to-end)
    h2d = M * 16 + 602 * 8  # scan (one copy) + prior {x, P} (kernel parameters)
    d2h = 607 * 8  # posterior {x, P, loop state} + sequence word, written by the kernel into mapped pinned memory

    # full scan (undistort-free preprocess + update), raw scan from pinned host memory: second line of SURVEY §8d
    def full_run(steps, warmup):
        t_ms = 0.0
        for k in range(warmup + steps):
            j = k % len(raws)
            l2_flush()
            torch.cuda.synchronize(dev)
            xs, Ps = wl["scans"][j]["x_prior"].copy(), P0.copy()
            t0 = time.perf_counter()
            # one C call: upload, voxel filter, update, posterior back; leaf_map = 0 keeps the map static
            # (the relocalisation loop of laserMapping_re.cpp, which is the shape of this config)
            ctx.scan_step(raws[j].numpy(), None, xs, Ps, wl["leaf"], 0.0, R_COV, wl["max_iter"], wl["ext"], True)
            dt = time.perf_counter() - t0
            if k >= warmup:
                t_ms += dt * 1000.0
        return t_ms

    full_run(0, 2)
    full_ms = max_over_ranks(full_run(max(5, args.steps // 2), 0))
    full_value = n_gpus * max(5, args.steps // 2) / (full_ms / 1000.0)

    # preprocessing alone (raw scan H2D + undistort/key + voxel filter), device-timed, cold L2; the upload alone beside it
    def prep_run(reps, with_kernels):
        ts = []
        dst = [torch.empty_like(r, device=dev) for r in raws]
        for k in range(reps + 3):
            j = k % len(raws)
            l2_flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if with_kernels:
                ctx.scan_preprocess(raws[j].numpy(), None, None, wl["leaf"], resident=True, want_m=False)
            else:
                dst[j].copy_(raws[j], non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts[3:]))

    prep_ms, prep_h2d_ms = prep_run(20, True), prep_run(20, False)
    preprocess = {"ms": prep_ms, "h2d_only_ms": prep_h2d_ms, "kernels_ms": prep_ms - prep_h2d_ms,
                  "what": "lio_scan_preprocess_resident from pinned host memory: H2D of the raw scan + undistort/leaf-hash "
                          "+ leaf ranks + placement + ordered centroids (4 kernels, no library kernel); CUDA events, L2 "
                          "flushed, median of 20; kernels_ms = ms - the same H2D alone"}

    # ---------------------------------------------------------------- several sequences per launch (config 4 shape)
    # Independent sequences (own map copy, own scan, own filter) sliced over ONE cooperative launch.  Not the headline:
    # `value` stays one sequence per GPU, the latency a robot sees.  Same device timing rules (events, L2 flushed).
    multi = None
    seqs = [int(v) for v in args.sequences.split(",") if v.strip()]
    if seqs:
        others = []
        for _ in range(max(seqs) - 1):
            c = _cabi.Context(local, max_scan_points=1 << 12, max_down_points=100000,
                              max_map_points=max(1 << 21, int(n_map * 1.05)), **extra)
            c.set_stream(stream.cuda_stream)
            c.map_build(map4)
            others.append(c)
        pool = [ctx] + others
        multi = {"unit": UNIT, "what": "n independent sequences (own map, scan, filter) in one cooperative launch, "
                 "device-timed, L2 flushed; scans/s summed over the sequences", "by_sequences": {}}
        for n in seqs:
            evs = []
            for k in range(3 + args.steps):
                for q in range(n):
                    j = (k + q) % len(bodies)
                    pool[q].scan_upload(bodies_np[j])
                    pool[q].state_upload(wl["scans"][j]["x_prior"], P0)
                l2_flush()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _cabi.update_enqueue_multi(pool[:n], R_COV, wl["max_iter"], wl["ext"], from_snapshot=True)
                e1.record(stream)
                if k >= 3:
                    evs.append((e0, e1))
                    pool[n - 1].state_download()
            torch.cuda.synchronize(dev)
            ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
            multi["by_sequences"][str(n)] = {"value": n_gpus * n * args.steps / (ms / 1000.0),
                                             "ms_per_launch": ms / args.steps}
        best = max(multi["by_sequences"], key=lambda k: multi["by_sequences"][k]["value"])
        multi["value"] = multi["by_sequences"][best]["value"]
        multi["sequences_per_launch"] = int(best)
        for c in others:
            c.close()

    # ---------------------------------------------------------------- roofline of the dominant kernel
    # update_kernel = the whole update (all passes) in one launch; algorithmic bytes per launch = 116 B x M per
    # h_share_model pass x passes.  Its duration IS the timed region of `value` (CUDA events around
    # the launch on its stream, L2 flushed before).  The single-pass kernel is timed alone as supporting evidence.
    peak, peak_src = measured_peak()
    ctx.scan_upload(bodies[0].numpy())
    ctx.state_upload(wl["scans"][0]["x_prior"], P0)
    m0 = bodies[0].shape[0]

    def time_pass(search, cold, reps=20):
        ts = []
        for _ in range(reps + 3):
            if cold:
                l2_flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.pass_only_enqueue(search, wl["ext"])
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts[3:]))

    t_search_cold = time_pass(True, True)
    t_search_warm = time_pass(True, False)
    t_cached_warm = time_pass(False, False)
    launch_ms = ms_cold / args.steps
    alg_bytes = 116.0 * M * passes_per_scan
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "update_kernel (persistent: all h_share_model passes + Kalman steps of one scan)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                "traffic_source": "profiles/r2_update_kernel_traffic.json (ncu --set full, dram__bytes_read+write.sum)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms_cold_l2": launch_ms,
                "note": "latency-bound by construction: 116 B x M x passes is ~4 MB per scan (SURVEY.md §8d)",
                "single_pass_kernel": {"search_ms_cold_l2": t_search_cold, "search_ms_warm_l2": t_search_warm,
                                       "cached_ms_warm_l2": t_cached_warm,
                                       "search_GBps_cold": 116.0 * m0 / (t_search_cold * 1e-3) / 1e9}}

    line = {
        "metric": METRIC if args.workload != "avia_200k" else METRIC.replace("OS1-128", "Avia"), "value": value,
        "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_cold / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": workload_name(args, n_map), "N_raw": int(np.mean([len(r) for r in raws])), "M": M,
                   "passes_per_scan": passes_per_scan, "l2": "flushed (384 MiB write) before every timed step",
                   "map_build_s": map_build_s},
        "value_l2_warm": value_warm, "ms_per_step_median": med_cold, "ms_per_step_median_l2_warm": med_warm,
        "matched_pts_per_s": matched,
        "full_scan": {"value": full_value, "unit": UNIT,
                      "what": "raw scan H2D + voxel downsample + update + posterior D2H through one lio_scan_step call (static map), host-timed"},
        "preprocess": preprocess,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step_median": e2e_med_ms, "ms_per_step_mean": e2e_mean_ms,
                "value_from_mean": n_gpus / (e2e_mean_ms / 1000.0),
                "how": "n_gpus / (max over ranks of the median host-timed step over `steps` steps)"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if multi:
        line["multi_sequence"] = multi

    ctx.close()
    del flush
    torch.cuda.empty_cache()
    legs = args.legs
    if legs == "auto":
        legs = "sharded,os1_64_seqs" if world > 1 else "knn_hbm,dense_scene"
    legs = [] if legs == "none" else [v.strip() for v in legs.split(",") if v.strip()]
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup = min(args.steps, 30), max(3, min(args.warmup, 5))
    if "sharded" in legs and world > 1:
        # config 5: ONE map cut into x-slabs over the ranks, the same scans on every rank, strong scaling
        sub.map_points, sub.workload = args.sharded_map_points, "sharded"
        r = sharded_leg(sub, rank, world, local, dev, torch, dist, _cabi)
        line["sharded"] = {"value": r["value"], "unit": UNIT, "scaling": "strong", "ms_per_step": r["ms_per_step"],
                           "steps": sub.steps, "exchange": args.exchange, "map_points": args.sharded_map_points,
                           "local_map_points": r["config"]["local_map_points"], "M": r["config"]["M"],
                           "owned_points_max_rank": r["config"].get("owned_points_max_rank"),
                           "passes_per_scan": r["config"]["passes_per_scan"],
                           "states_identical": r["config"]["states_identical_across_ranks"],
                           "vs_single_gpu": r["config"].get("vs_single_gpu"), "dense": r["config"].get("dense"),
                           "value_unaligned": r["config"].get("value_unaligned"), "timing": r["config"].get("timing"),
                           "scan_points_last": r["config"].get("scan_points_last"),
                           "what": r["config"]["workload"]}
    if "os1_64_seqs" in legs and world > 1:
        # config 4: 8 independent OS1-64 sequences per GPU through the whole main loop
        sub.workload, sub.seqs_per_gpu = "os1_64_seqs", 8
        r = traj_leg(sub, rank, world, local, dev, torch, dist, _cabi)
        line["os1_64_seqs"] = {"value": r["value"], "unit": UNIT, "scaling": "weak", "ms_per_step": r["ms_per_step"],
                               "steps": sub.steps, "sequences": 8 * world, "host_threads": r["config"]["host_threads"],
                               "M": r["config"]["M"], "N_raw": r["config"]["N_raw"], "what": r["config"]["workload"]}
    if "knn_hbm" in legs and world == 1:
        line["knn_hbm"] = knn_hbm_leg(args, local, dev, torch, _cabi)
    if "dense_scene" in legs and world == 1:
        line["dense_scene"] = dense_leg(args, wl, map4, local, dev, torch, _cabi, stream)

    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        r = cpu_arm(wl, steps=1000, warmup=1, seconds_budget=args.cpu_seconds)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                "kind_detail": r["kind_detail"],
                                "sample": "%d whole updates of the same scans in %.1f s (tree build %.1f s untimed)" %
                                (r["steps"], r["steps"] / r["value"], r["build_s"]),
                                "matched_pts_per_s": r["matched_pts_per_s"], "cpu_model": r["cpu_model"],
                                "stages_ms": r["stages_ms"]}
        r3 = cpu_arm(wl, steps=1000, warmup=1, seconds_budget=min(8.0, args.cpu_seconds), threads=3)
        line["cpu_baseline_3_threads"] = {"value": r3["value"], "unit": UNIT, "cores": 3,
                                          "note": "MP_PROC_NUM=3, the reference's own setting (CMakeLists.txt:23-26)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
