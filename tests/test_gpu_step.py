"""lio_scan_step: one main-loop iteration (laserMapping.cpp:737-785) enqueued with a single host synchronisation must
give exactly what the same stages give when the host drives them one call at a time."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx():
    from agi_lidar_slam_b200 import _cabi

    return _cabi.Context(0, max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 19)


def _staged(rep, meas, poses):
    """The device part of one iteration through the per-stage entry points (what lio_scan_step fuses)."""
    c, ctx = rep.cfg, rep.ctx
    m = ctx.scan_preprocess(meas.lidar, poses, rep.x, c.filter_size_surf, resident=True)
    if m < 5:
        return dict(status="few-points", m=m)
    if not rep.map_built:
        ctx.map_build_scan(rep.x)
        rep.map_built = True
        return dict(status="map-built", m=m)
    rep.x, rep.P, nv, npass = ctx.update_scan(rep.x, rep.P, 0.001, c.max_iteration, c.extrinsic_est)
    counts = ctx.map_incremental(rep.x, c.filter_size_map, rep._ekf_inited)
    return dict(status="ok", m=m, n_valid=nv, n_passes=npass, counts=counts.tolist())


@pytest.mark.parametrize("ext", [False, True])
def test_scan_step_equals_staged_calls(ext):
    from agi_lidar_slam_b200 import synth
    from agi_lidar_slam_b200.replay import LioReplay, MeasureGroup, ReplayConfig

    seq = synth.sequence(16, 2002, rings=16, cols=900)
    a_ctx, b_ctx = _ctx(), _ctx()
    try:
        A = LioReplay(a_ctx, ReplayConfig(max_iteration=3, extrinsic_est=ext))
        B = LioReplay(b_ctx, ReplayConfig(max_iteration=3, extrinsic_est=ext))
        n_ok = 0
        few = np.array([[1.0, 0.0, 0.0, 0.0], [2.0, 0.1, 0.0, 1.0], [3.0, 0.0, 0.2, 2.0]], np.float32)
        for j, m in enumerate(seq):
            lidar = few if j == 9 else m["lidar"]  # one scan of three points in the middle: skipped by both
            mg = MeasureGroup(lidar, m["imu"], m["lidar_beg_time"], m["lidar_end_time"])
            pa = A.host_stage(mg)
            rb = B.process(mg)
            if pa is None:
                assert rb is None
                continue
            la = _staged(A, mg, pa)
            lb = B.log[-1]
            assert la == lb, (j, la, lb)
            assert np.array_equal(A.x, B.x) and np.array_equal(A.P, B.P), j
            if la["status"] == "ok":
                n_ok += 1
                assert np.array_equal(rb, A.x)
            if j == 9:
                assert la["status"] == "few-points" and la["m"] == 3
        assert n_ok >= 8
        ax, ai = a_ctx.map_dump()
        bx, bi = b_ctx.map_dump()
        assert np.array_equal(ai, bi) and np.array_equal(ax.view(np.uint32), bx.view(np.uint32))
        assert a_ctx.map_size() == b_ctx.map_size()
    finally:
        a_ctx.close()
        b_ctx.close()


def test_scan_step_empty_scan_and_errors():
    from agi_lidar_slam_b200 import _cabi, synth
    from agi_lidar_slam_b200.replay import default_state

    cfg = synth.small_config()
    ctx = _ctx()
    try:
        x, P = default_state(), np.eye(24) * 0.01
        x[0:3] = cfg["x_true"][0:3]
        x[3:7] = cfg["x_true"][3:7]
        # first call on an empty map: builds it from the scan
        rep = ctx.scan_step(cfg["scan"], None, x, P, 0.5, 0.5)
        assert rep.status == _cabi.SCAN_MAP_BUILT and rep.m > 100
        total0 = ctx.map_size()
        x0, P0 = x.copy(), P.copy()
        # an empty scan and a 4-point scan: skipped, nothing changes
        for n in (0, 4):
            rep = ctx.scan_step(cfg["scan"][:n].copy(), None, x, P, 0.5, 0.5)
            assert rep.status == _cabi.SCAN_FEW_POINTS and rep.m <= n
            assert np.array_equal(x, x0) and np.array_equal(P, P0) and ctx.map_size() == total0
        # and the next full scan updates as usual
        rep = ctx.scan_step(cfg["scan"], None, x, P, 0.5, 0.5)
        assert rep.status == _cabi.SCAN_UPDATED and rep.n_valid > 100 and rep.n_passes >= 2
        assert not np.array_equal(P, P0)
        # end / finish without begin
        with pytest.raises(_cabi.LioError):
            ctx.scan_step_end(0.5)
        with pytest.raises(_cabi.LioError):
            ctx.scan_step_finish(x, P)
    finally:
        ctx.close()


def test_native_loop_equals_python_loop():
    """lio_seq_process (the main loop in C++) against replay.LioReplay (the same loop spelled out in Python over the same
    entry points): bit-identical states, logs, map and local-map box, including a sliding 40 m local-map cube."""
    from agi_lidar_slam_b200 import synth
    from agi_lidar_slam_b200.replay import LioReplay, MeasureGroup, NativeReplay, ReplayConfig

    seq = synth.sequence(60, 2002, rings=16, cols=600)
    a_ctx, b_ctx = _ctx(), _ctx()
    try:
        cfg = ReplayConfig(max_iteration=3, cube_len=40.0, det_range=10.0)
        A, B = LioReplay(a_ctx, cfg), NativeReplay(b_ctx, cfg)
        for j, m in enumerate(seq):
            mg = MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"])
            ra, rb = A.process(mg), B.process(mg)
            assert (ra is None) == (rb is None), j
            assert A.log[-1] == B.log[-1], (j, A.log[-1], B.log[-1])
            assert np.array_equal(A.x, B.x) and np.array_equal(A.P, B.P), j
            if A.local_map is not None:
                assert np.array_equal(A.local_map, B.local_map)
        assert A.n_box_deleted == B.n_box_deleted and A.n_box_deleted > 0
        assert sum(e["status"] == "ok" for e in B.log) >= 50
        ax, ai = a_ctx.map_dump()
        bx, bi = b_ctx.map_dump()
        assert np.array_equal(ai, bi) and np.array_equal(ax.view(np.uint32), bx.view(np.uint32))
    finally:
        a_ctx.close()
        b_ctx.close()


def test_deferred_map_growth_changes_nothing():
    """lio_set_deferred_growth: the step returns with the posterior while the scan's map growth still runs (the reference
    publishes before map_incremental, laserMapping.cpp:776-785).  Same states, same counts (collected by
    lio_scan_step_settle), same map, same box deletes -- with a sliding local-map cube that touches the map between scans."""
    from agi_lidar_slam_b200 import synth
    from agi_lidar_slam_b200.replay import MeasureGroup, NativeReplay, ReplayConfig

    seq = synth.sequence(50, 2002, rings=16, cols=600)
    a_ctx, b_ctx = _ctx(), _ctx()
    try:
        cfg = ReplayConfig(max_iteration=3, cube_len=40.0, det_range=10.0)
        b_ctx.set_deferred_growth(True)
        A, B = NativeReplay(a_ctx, cfg), NativeReplay(b_ctx, cfg)
        n_ok = 0
        for j, m in enumerate(seq):
            mg = MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"])
            ra, rb = A.process(mg), B.process(mg)
            assert (ra is None) == (rb is None), j
            la, lb = dict(A.log[-1]), dict(B.log[-1])
            if la["status"] == "ok":
                n_ok += 1
                assert lb.pop("counts") == [-1, -1, -1]
                ca = la.pop("counts")
                if j % 3 == 0:  # settle explicitly now and then; otherwise the next step settles on its own
                    assert b_ctx.scan_step_settle() == ca, j
            assert la == lb, (j, la, lb)
            assert np.array_equal(A.x, B.x) and np.array_equal(A.P, B.P), j
        assert n_ok >= 40
        assert b_ctx.scan_step_settle() == A.log[-1]["counts"]
        assert A.seq.local_map()[1] == B.seq.local_map()[1] and A.seq.local_map()[1] > 0
        ax, ai = a_ctx.map_dump()
        bx, bi = b_ctx.map_dump()
        assert np.array_equal(ai, bi) and np.array_equal(ax.view(np.uint32), bx.view(np.uint32))
        assert a_ctx.map_size() == b_ctx.map_size()
    finally:
        a_ctx.close()
        b_ctx.close()


def test_deferred_growth_reports_capacity_errors_one_call_later():
    """A map that outgrows lio_caps.max_map_points: the synchronous loop fails in the scan whose growth overflowed, the
    deferred loop in the next call that settles that growth -- never silently."""
    from agi_lidar_slam_b200 import _cabi, synth
    from agi_lidar_slam_b200.replay import MeasureGroup, NativeReplay, ReplayConfig

    seq = synth.sequence(40, 2002, rings=16, cols=600)
    failed_at = {}
    for deferred in (False, True):
        ctx = _cabi.Context(0, max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=6000)
        try:
            ctx.set_deferred_growth(deferred)
            rep = NativeReplay(ctx, ReplayConfig(max_iteration=3))
            for j, m in enumerate(seq):
                try:
                    rep.process(MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"]))
                except _cabi.LioError as e:
                    assert e.code == _cabi.LIO_E_CAPACITY, e
                    failed_at[deferred] = j
                    break
        finally:
            ctx.close()
    assert False in failed_at and True in failed_at, failed_at
    assert failed_at[True] in (failed_at[False], failed_at[False] + 1), failed_at


def test_static_map_step_is_preprocess_plus_update():
    """leaf_map = 0: the relocalisation loop (laserMapping_re.cpp, map_incremental() commented out at :676) -- one call
    gives what lio_scan_preprocess + lio_update_scan give, and the map stays as it was."""
    from agi_lidar_slam_b200 import _cabi, synth

    cfg = synth.small_config()
    a, b = _ctx(), _ctx()
    try:
        mp = np.concatenate([cfg["map"], np.zeros((len(cfg["map"]), 1), np.float32)], 1)
        a.map_build(mp)
        b.map_build(mp)
        size0 = b.map_size()
        m = a.scan_preprocess(cfg["scan"], None, None, 0.5, resident=True)
        xa, Pa, nv, npass = a.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, False)
        xb = np.ascontiguousarray(cfg["x_prior"], np.float64).copy()
        Pb = np.ascontiguousarray(cfg["P"], np.float64).reshape(24, 24).copy()
        rep = b.scan_step(cfg["scan"], None, xb, Pb, 0.5, 0.0, 0.001, 4, False, True)
        assert rep.status == _cabi.SCAN_UPDATED and rep.m == m and rep.n_valid == nv and rep.n_passes == npass
        assert list(rep.counts) == [0, 0, 0]
        assert np.array_equal(xa, xb) and np.array_equal(np.asarray(Pa).reshape(24, 24), Pb)
        assert b.map_size() == size0
        xyz_a, ids_a = a.map_dump()
        xyz_b, ids_b = b.map_dump()
        assert np.array_equal(ids_a, ids_b) and np.array_equal(xyz_a.view(np.uint32), xyz_b.view(np.uint32))
    finally:
        a.close()
        b.close()
