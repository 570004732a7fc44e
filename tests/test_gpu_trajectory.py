"""End-to-end parity on BASELINE.json config 2 (SURVEY.md App. C last rows): a 200-scan synthetic VLP-16 trajectory
with IMU goes through the product's replay of the reference main loop (IMU init -> forward propagation -> fused
undistort + voxel filter -> first-scan map build -> IESKF update -> map_incremental) and through the same loop on the
CPU oracle.  north_star: per-scan pose within 1e-4 m and 1e-4 rad.

The comparison is PER SCAN from a common starting point: after every scan the product continues from the oracle's
posterior (and, if the two maps ever differ, from the oracle's map).  A free-running comparison cannot hold 1e-4 for
ANY two implementations that are not bit-identical: the reference algorithm itself (growing map, discrete insert and
validity decisions) amplifies a 1e-9 state perturbation to ~5e-3 m within ten scans -- measured on the oracle against
itself in tests/test_oracle_pipeline.py::test_reference_loop_is_chaotic."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_SCANS = 200
TOL_POS = 1e-4  # m    (BASELINE.json north_star)
TOL_ROT = 1e-4  # rad  (BASELINE.json north_star)


def _p4(xyz):
    return np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1)


@pytest.mark.parametrize("name,n_scans,seed,rings,cols,fov,max_range,cube_len,det_range,extent,ref_tree", [
    ("config 2 (VLP-16, 200 scans)", N_SCANS, 2002, 16, 1800, (-15.0, 15.0), 100.0, 1000.0, 300.0, 120.0, False),
    ("config 4 shape (OS1-64, one of the 64 sequences, 40 scans)", 40, 4000, 64, 1024, (-22.5, 22.5), 120.0, 1000.0, 300.0,
     120.0, False),
    # a 40 m local-map cube: lasermap_fov_segment slides it and box-deletes the map behind the sensor
    ("moving local map (VLP-16, 80 scans, cube 40 m)", 80, 2002, 16, 1800, (-15.0, 15.0), 100.0, 40.0, 10.0, 120.0, False),
    # sparse scene, and the checker's live map is the REFERENCE ikd-Tree (Build / unbounded Nearest_Search / Add_Points):
    # ~6 % of the points fail gate 1 and map_incremental classifies them from neighbours beyond sqrt(5) m
    # (tests/test_oracle_reference_map.py is the CPU twin of this case)
    ("sparse 240 m yard on the reference ikd-Tree (36 scans)", 36, 2002, 16, 900, (-15.0, 15.0), 200.0, 1000.0, 300.0, 240.0,
     True),
])
def test_trajectory_pose_parity(orc, name, n_scans, seed, rings, cols, fov, max_range, cube_len, det_range, extent,
                                ref_tree):
    from agi_lidar_slam_b200 import _cabi, synth
    from agi_lidar_slam_b200.replay import MeasureGroup, NativeReplay as LioReplay, ReplayConfig
    from replay_oracle import OracleReplay

    if ref_tree and not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not present")
    N_SCANS = n_scans
    seq = synth.sequence(n_scans, seed, rings=rings, cols=cols, fov=fov, max_range=max_range,
                         scene=synth.block_scene(extent=extent))  # 10 Hz, IMU 200 Hz
    traj = synth.RampedTrajectory(synth.Trajectory())
    R0, p0 = traj.rot(0.0), traj.pos(0.0)
    with _cabi.Context(0, max_scan_points=1 << 17, max_down_points=1 << 16, max_map_points=1 << 21) as ctx, \
            _cabi.Context(0, max_scan_points=1 << 17, max_down_points=1 << 16, max_map_points=1 << 21) as ctx_free:
        gpu = LioReplay(ctx, ReplayConfig(max_iteration=3, cube_len=cube_len, det_range=det_range))
        free = LioReplay(ctx_free, ReplayConfig(max_iteration=3, cube_len=cube_len, det_range=det_range))
        cpu = OracleReplay(orc, max_iteration=3, cube_len=cube_len, det_range=det_range, use_ikd=ref_tree)
        worst_pos, worst_rot, worst_P, n_upd, exact_scans, map_resync = 0.0, 0.0, 0.0, 0, 0, 0
        gate1_fail = rows = 0
        horizon, free_worst = None, 0.0  # free-running product (never re-synchronised) against the oracle
        for k, m in enumerate(seq):
            mg = MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"])
            a = gpu.process(mg)
            f = free.process(mg)
            b = cpu.process(m)
            assert (a is None) == (b is None)  # same skip decisions (first scan, IMU init, map build)
            if gpu.map_built:
                gx, gi = ctx.map_dump()
                ox, oi = cpu.map.dump()
                same_map = gx.shape == ox.shape and np.array_equal(gi, oi) and \
                    np.array_equal(gx.view(np.uint32), ox.view(np.uint32))
            else:
                same_map = True
            if a is not None:
                n_upd += 1
                worst_pos = max(worst_pos, float(np.abs(a[0:3] - b[0:3]).max()))
                worst_rot = max(worst_rot, float(np.linalg.norm(orc.boxminus(a, b)[3:6])))
                worst_P = max(worst_P, float(np.abs(gpu.P - cpu.P).max() / np.abs(cpu.P).max()))
                ga, gb = gpu.log[-1], cpu.log[-1]
                exact = (ga["m"] == gb["m"] and ga["n_valid"] == gb["n_valid"] and ga["n_passes"] == gb["n_passes"]
                         and ga["counts"] == gb["counts"] and same_map)
                assert exact, (k, ga, gb, same_map)
                exact_scans += exact
                rows += len(cpu.last["d2"])
                gate1_fail += int((cpu.last["d2"][:, 4] > 5).sum())
                if f is not None:
                    df = max(float(np.abs(f[0:3] - b[0:3]).max()), float(np.linalg.norm(orc.boxminus(f, b)[3:6])))
                    free_worst = max(free_worst, df)
                    if horizon is None and df > TOL_POS:
                        horizon = n_upd
            # next scan starts from the oracle's posterior on both sides
            gpu.x, gpu.P = cpu.x.copy(), cpu.P.copy()
            if not same_map:
                map_resync += 1
                ctx.map_build(_p4(ox))
        map_total, map_valid = ctx.map_size()
    assert n_upd >= N_SCANS - 4
    assert worst_pos < TOL_POS and worst_rot < TOL_ROT, (worst_pos, worst_rot)
    assert worst_P < 1e-6
    # discrete outcomes (M, matched-point count, pass count, map_incremental classes, the map itself) identical in
    # EVERY scan (asserted per scan above), hence nothing was ever re-synchronised
    assert exact_scans == n_upd and map_resync == 0
    assert gpu.n_box_deleted == cpu.n_box_deleted and (cube_len > 100 or gpu.n_box_deleted > 0)
    if ref_tree:
        assert gate1_fail >= 0.05 * rows, (gate1_fail, rows)
    # the odometry follows the motion (scan-to-map registration on a 16-ring sensor drifts; both sides agree on it)
    pr = R0.T @ (seq[-1]["truth_pos"] - p0)
    assert np.linalg.norm(b[0:3] - pr) < 0.25 * max(1.0, np.linalg.norm(pr))
    # the reference loop amplifies any rounding difference (test_reference_loop_is_chaotic): how long a product that is
    # never re-synchronised stays within the tolerance is reported, not required
    print(f"{name}: {n_upd} updates, worst |dpos| {worst_pos:.2e} m, worst |drot| {worst_rot:.2e} rad, worst rel |dP| "
          f"{worst_P:.2e}, bit-identical discrete outcomes in {exact_scans}/{n_upd} scans, map resyncs {map_resync}, "
          f"map {map_valid} pts, {gpu.n_box_deleted} points box-deleted, gate-1 failures {100.0 * gate1_fail / max(rows, 1):.1f} %; "
          f"free-running: within 1e-4 for {'all ' + str(n_upd) if horizon is None else 'the first ' + str(horizon - 1)} "
          f"updates (worst {free_worst:.2e})")


def test_process_many_equals_independent_replays():
    """Config 4 shape on one GPU: three independent sequences stepped together (updates in one cooperative launch per
    scan, replay.process_many) follow the same trajectories as the same sequences replayed one by one.  The reference
    loop amplifies rounding differences (~5x per scan), so the horizon is short and the tolerance is 1e-6."""
    from agi_lidar_slam_b200 import _cabi, synth
    from agi_lidar_slam_b200.replay import (LioReplay, MeasureGroup, NativeReplay, ReplayConfig, native_process_many,
                                            process_many)

    n_seq, n_scans = 3, 12
    seqs = [synth.sequence(n_scans, 4000 + k, rings=32, cols=512, fov=(-22.5, 22.5), max_range=120.0) for k in range(n_seq)]
    mg = [[MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"]) for m in s] for s in seqs]

    def contexts():
        return [_cabi.Context(0, max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 19)
                for _ in range(n_seq)]

    solo = []
    for k, c in enumerate(contexts()):
        r = LioReplay(c, ReplayConfig(max_iteration=3))
        solo.append([r.process(m) for m in mg[k]])
        c.close()
    ctxs = contexts()
    reps = [LioReplay(c, ReplayConfig(max_iteration=3)) for c in ctxs]
    n_upd = 0
    for j in range(n_scans):
        res = process_many(reps, [mg[k][j] for k in range(n_seq)])
        for k in range(n_seq):
            assert (res[k] is None) == (solo[k][j] is None)
            if res[k] is not None:
                n_upd += 1
                assert np.abs(res[k] - solo[k][j]).max() < 1e-6
    assert n_upd >= n_seq * (n_scans - 4)
    for c in ctxs:
        c.close()
    # the native loop (lio_seq_process_many) takes the same steps
    ctxs = contexts()
    reps = [NativeReplay(c, ReplayConfig(max_iteration=3)) for c in ctxs]
    native = []
    for j in range(n_scans):
        res = native_process_many(reps, [mg[k][j] for k in range(n_seq)])
        native.append(res)
        for k in range(n_seq):
            assert (res[k] is None) == (solo[k][j] is None)
            if res[k] is not None:
                assert np.abs(res[k] - solo[k][j]).max() < 1e-6
    maps = [c.map_dump() for c in ctxs]
    for c in ctxs:
        c.close()
    # host threads and deferred map growth change nothing: same bits with one thread / growth inside the step ...
    for threads, deferred in ((1, False), (3, True)):
        _cabi.set_host_threads(threads)
        ctxs = contexts()
        for c in ctxs:
            c.set_deferred_growth(deferred)
        reps = [NativeReplay(c, ReplayConfig(max_iteration=3)) for c in ctxs]
        for j in range(n_scans):
            res = native_process_many(reps, [mg[k][j] for k in range(n_seq)])
            for k in range(n_seq):
                assert (res[k] is None) == (native[j][k] is None)
                if res[k] is not None:
                    assert np.array_equal(res[k], native[j][k]), (threads, deferred, j, k)
        for k, c in enumerate(ctxs):
            xyz, ids = c.map_dump()
            assert np.array_equal(ids, maps[k][1]) and np.array_equal(xyz.view(np.uint32), maps[k][0].view(np.uint32))
            c.close()
    _cabi.set_host_threads(0)
