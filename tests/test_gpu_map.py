"""Map maintenance parity: Build / Add_Points (with and without downsample) / Delete_Point_Boxes / flatten and the
map_incremental policy, against the oracle map (itself pinned to the reference ikd-Tree in test_oracle_map.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _p4(xyz):
    return np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1)


def _same_map(ctx, om):
    gx, gi = ctx.map_dump()
    ox, oi = om.dump()
    assert np.array_equal(gi, oi)
    assert np.array_equal(gx.view(np.uint32), ox.view(np.uint32))


def test_build_and_dump(ctx, orc, small_cfg):
    mp = small_cfg["map"]
    ctx.map_build(_p4(mp))
    om = orc.Map(1.0)
    om.build(mp)
    _same_map(ctx, om)
    assert ctx.map_size() == (len(mp), len(mp))


def test_add_points_sequences(ctx, orc):
    rng = np.random.default_rng(2)
    base = rng.uniform(-6, 6, (3000, 3)).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
    ctx.map_build(_p4(base))
    om = orc.Map(1.0)
    om.build(base)
    for it in range(6):
        new = rng.uniform(-7, 7, (2500, 3)).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
        if it == 2:
            k = len(new[1::3])
            new[::3][:k] = new[1::3]  # duplicates inside one batch
        ds = it % 3 != 2
        a = ctx.map_add(_p4(new), ds)
        b = om.add(new, ds, 0.5)
        assert a == b
        _same_map(ctx, om)
        q = rng.uniform(-7, 7, (2000, 3)).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
        gi, gd, _ = ctx.knn5(q)
        oi, od, _ = om.knn(q)
        assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    t, v = ctx.map_size()
    assert v == om.size()


def test_bucket_growth_dense_cell(ctx, orc):
    """Hundreds of un-thinned points in one cell force repeated bucket doubling."""
    rng = np.random.default_rng(3)
    base = rng.uniform(0, 1, (40, 3)).astype(np.float32)
    ctx.map_build(_p4(base))
    om = orc.Map(1.0)
    om.build(base)
    for _ in range(5):
        new = rng.uniform(0, 1, (300, 3)).astype(np.float32)
        ctx.map_add(_p4(new), False)
        om.add(new, False)
    _same_map(ctx, om)
    q = rng.uniform(0, 1, (500, 3)).astype(np.float32)
    gi, gd, _ = ctx.knn5(q)
    oi, od, _ = om.knn(q)
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))


def test_delete_boxes(ctx, orc, small_cfg):
    mp = small_cfg["map"]
    ctx.map_build(_p4(mp))
    om = orc.Map(1.0)
    om.build(mp)
    boxes = np.array([[-5, -5, -1, 5, 5, 3], [10, -30, -1, 30, 0, 10]], np.float32)
    assert ctx.map_delete_boxes(boxes) == om.delete_boxes(boxes) > 0
    _same_map(ctx, om)


def test_errors(ctx):
    from agi_lidar_slam_b200 import _cabi

    with pytest.raises(_cabi.LioError) as e:
        ctx.map_add(np.zeros((3, 4), np.float32), True)
    assert e.value.code == _cabi.LIO_E_EMPTY_MAP
    ctx.map_build(np.zeros((0, 4), np.float32))  # Build([]) leaves the map empty (ikd_Tree.cpp:359)
    with pytest.raises(_cabi.LioError):
        ctx.update_scan(np.zeros(26), np.eye(24))
    with pytest.raises(_cabi.LioError) as e:
        ctx.map_build(np.zeros((int(ctx.caps.max_map_points) + 1, 4), np.float32))
    assert e.value.code == _cabi.LIO_E_CAPACITY


def test_map_incremental_matches_oracle(ctx, orc, small_cfg):
    cfg = small_cfg
    mp = cfg["map"][::2]  # thinner map so that some scan voxels are new
    ctx.map_build(_p4(mp))
    om = orc.Map(1.0)
    om.build(mp)
    s = cfg["scan"]
    pts5 = np.concatenate([s[:, :3], np.zeros((len(s), 1), np.float32), s[:, 3:4]], 1)
    body = np.ascontiguousarray(orc.voxel_grid(pts5, 0.5)[0][:, :4])
    ctx.scan_upload(body)
    x, P, nv, npass = ctx.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, False)
    sc = orc.Scan(body[:, :3])
    xr, Pr, trace, _ = sc.update(cfg["x_prior"], cfg["P"], om.knn_backend(), 0.001, 4, False, threads=8)
    ref = sc.get()
    world = orc.body_to_world(xr, body[:, :3])
    cls = orc.map_incremental_classify(world, ref["near_raw"], ref["cnt"], True, 0.5)
    na = om.add(world[cls == 1], True, 0.5)
    om.add(world[cls == 2], False)
    counts = ctx.map_incremental(xr, 0.5, True)  # same state on both sides: stage-wise comparison
    assert counts.tolist() == [int((cls == 1).sum()), int((cls == 2).sum()), na]
    _same_map(ctx, om)


@pytest.mark.parametrize("ds", [0.2, 0.4, 0.75])
def test_add_points_any_downsample_size_equals_reference_tree(ctx, orc, ds):
    """set_downsample_param with a voxel that does not divide the 1.5 m kNN cell: the survivors are the
    reference tree's, every one of them filed where the searches look for it (kNN against the tree afterwards)."""
    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not present")
    rng = np.random.default_rng(11)
    sc = np.array([1, 1, 0.15], np.float32)
    base = rng.uniform(-9, 9, (6000, 3)).astype(np.float32) * sc
    ctx.map_set_downsample(ds)
    ctx.map_build(_p4(base))
    ref = orc.IkdLiveMap(ds)
    ref.build(base)
    for it in range(5):
        new = rng.uniform(-10, 10, (5000, 3)).astype(np.float32) * sc
        a = ctx.map_add(_p4(new), True)
        b = ref.add(new, True, ds)
        assert a == b, (it, a, b)
        gx, gi = ctx.map_dump()
        rx, ri = ref.dump()
        assert np.array_equal(gi, ri) and np.array_equal(gx.view(np.uint32), rx.view(np.uint32)), it
        q = rng.uniform(-11, 11, (4000, 3)).astype(np.float32) * sc
        idx, d2, _ = ctx.knn5(q, max_d2=np.inf)
        ti, td, _ = ref.knn(q, 5)
        assert np.array_equal(d2.view(np.uint32), td.view(np.uint32)) and (idx == ti).mean() > 0.999, it


def test_map_incremental_large_filter_size(ctx, orc, small_cfg):
    """filter_size_map so large (3 fs^2 > 5) that neighbours beyond sqrt(5) m can decide need_add: the rows are
    completed to the reference's five unbounded neighbours before the classification."""
    cfg = small_cfg
    mp = cfg["map"][::20]  # thin map: many rows are short after the bounded search
    ctx.map_build(_p4(mp))
    om = orc.Map(1.0)
    om.build(mp)
    s = cfg["scan"]
    pts5 = np.concatenate([s[:, :3], np.zeros((len(s), 1), np.float32), s[:, 3:4]], 1)
    body = np.ascontiguousarray(orc.voxel_grid(pts5, 0.5)[0][:, :4])
    ctx.scan_upload(body)
    ctx.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, False)
    sc = orc.Scan(body[:, :3])
    xr, Pr, trace, _ = sc.update(cfg["x_prior"], cfg["P"], om.knn_backend(), 0.001, 4, False, threads=8)
    ref = sc.get()
    assert (ref["d2"][:, 4] > 5).mean() > 0.05
    world = orc.body_to_world(xr, body[:, :3])
    for fs in (1.5, 2.0):
        cls = orc.map_incremental_classify(world, ref["near_raw"], ref["cnt"], True, fs)
        na = om.add(world[cls == 1], True, fs)
        om.add(world[cls == 2], False)
        counts = ctx.map_incremental(xr, fs, True)
        assert counts.tolist() == [int((cls == 1).sum()), int((cls == 2).sum()), na], fs
        _same_map(ctx, om)


def test_removed_points_log(orc, small_cfg):
    """lio_map_removed_points ≙ KD_TREE::acquire_removed_points: box deletes and downsample replacements are handed out
    once, then the log is empty; Build starts a new one."""
    from agi_lidar_slam_b200 import _cabi

    mp = small_cfg["map"]
    p4 = lambda a: np.concatenate([a, np.zeros((len(a), 1), np.float32)], 1)  # noqa: E731
    with _cabi.Context(0, max_scan_points=1 << 12, max_down_points=1 << 15, max_map_points=1 << 17) as ctx:
        ctx.map_build(p4(mp))
        assert len(ctx.map_removed_points()) == 0
        box = np.array([[-5, -5, -10, 5, 5, 10]], np.float32)
        inside = (mp[:, 0] >= -5) & (mp[:, 0] < 5) & (mp[:, 1] >= -5) & (mp[:, 1] < 5)
        n_del = ctx.map_delete_boxes(box)
        assert n_del == int(inside.sum()) > 0
        got = ctx.map_removed_points()
        key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]  # noqa: E731
        assert np.array_equal(key(got).view(np.uint32), key(mp[inside]).view(np.uint32))
        assert len(ctx.map_removed_points()) == 0  # taken
        # a downsampling Add_Points whose new point beats the old one of its box retires the old one
        live_xyz, _ = ctx.map_dump()
        victim = live_xyz[len(live_xyz) // 2]
        centre = (np.floor(victim / 0.5) + 0.5) * 0.5
        ctx.map_set_downsample(0.5)
        added = ctx.map_add(p4(centre[None].astype(np.float32)), True)
        got = ctx.map_removed_points()
        if added == 1:  # the box centre itself is closer than any sampled point unless one sits exactly there
            assert len(got) >= 1 and (np.abs(got - victim).max(1) < 1e-6).any()
        ctx.map_build(p4(mp))
        assert len(ctx.map_removed_points()) == 0
