"""The oracle's map semantics pinned against the REFERENCE ikd-Tree compiled in place (oracle/_ref/libikd_ref.so):
Build + Nearest_Search, Add_Points with/without downsample, Delete_Point_Boxes, flatten."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def need_ikd(orc):
    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not built (needs /root/reference)")


def _rows_sorted(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


def _bounded(ti, td, max_d2=5.0):
    td = np.where(td <= max_d2, td, np.inf).astype(np.float32)
    return np.where(np.isfinite(td), ti, -1), td


def test_knn_equals_reference_tree(orc, need_ikd, small_cfg):
    mp = small_cfg["map"]
    q = orc.body_to_world(small_cfg["x_prior"], small_cfg["scan"][:, :3])
    om = orc.Map(1.0)
    om.build(mp)
    t = orc.IkdTree()
    t.build(mp)
    oi, od, on = om.knn(q)
    ti, td = _bounded(*t.knn(q)[:2])
    assert np.array_equal(od.view(np.uint32), td.view(np.uint32))
    assert np.array_equal(oi, ti)


def test_knn_independent_of_cell_size(orc, small_cfg):
    mp = small_cfg["map"]
    q = orc.body_to_world(small_cfg["x_prior"], small_cfg["scan"][::7, :3])
    ref = None
    for cell in (0.5, 1.0, 2.5):
        om = orc.Map(cell)
        om.build(mp)
        r = om.knn(q)
        if ref is not None:
            assert np.array_equal(r[0], ref[0]) and np.array_equal(r[1], ref[1])
        ref = r


def test_knn_brute_force(orc):
    rng = np.random.default_rng(0)
    mp = rng.uniform(-4, 4, (1500, 3)).astype(np.float32)
    q = rng.uniform(-4, 4, (200, 3)).astype(np.float32)
    om = orc.Map(1.0)
    om.build(mp)
    oi, od, _ = om.knn(q)
    d = ((q[:, None, :] - mp[None, :, :]) ** 2)
    d2 = (d[..., 0] + d[..., 1]) + d[..., 2]  # same FP32 association as the reference
    order = np.lexsort((np.broadcast_to(np.arange(len(mp)), d2.shape), d2), axis=1)[:, :5]
    bd = np.take_along_axis(d2, order, 1)
    assert np.array_equal(np.where(bd <= 5, order, -1), oi)
    assert np.array_equal(np.where(bd <= 5, bd, np.inf).astype(np.float32), od)


def test_add_points_downsample_equals_reference_tree(orc, need_ikd):
    """After every Add_Points call the live point sets of the oracle map and the reference tree coincide."""
    rng = np.random.default_rng(1)
    scale = np.array([1, 1, 0.15], np.float32)
    base = rng.uniform(-5, 5, (2000, 3)).astype(np.float32) * scale
    om = orc.Map(1.0)
    om.build(base)
    t = orc.IkdTree()
    t.set_downsample_param(0.5)
    t.build(base)
    for it in range(8):
        new = rng.uniform(-6, 6, (1500, 3)).astype(np.float32) * scale
        ds = it % 4 != 3
        om.add(new, ds, 0.5)
        t.add_points(new, ds)
        a = _rows_sorted(om.dump()[0])
        b = _rows_sorted(t.flatten()[0])
        assert a.shape == b.shape, (it, a.shape, b.shape)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), it
        assert t.validnum() == om.size()
    # and searches on the grown structures agree
    q = rng.uniform(-6, 6, (1500, 3)).astype(np.float32) * scale
    oi, od, on = om.knn(q)
    ti, td, tn = t.knn(q)
    _, tdb = _bounded(ti, td)
    assert np.array_equal(od.view(np.uint32), tdb.view(np.uint32))
    ok = np.isfinite(od)
    assert np.array_equal(on[ok].view(np.uint32), tn[ok].view(np.uint32))


def test_downsample_rule_details(orc, need_ikd):
    """New point wins ties; a single existing point that is closer survives; >1 existing points collapse."""
    for case in range(3):
        om = orc.Map(1.0)
        t = orc.IkdTree()
        t.set_downsample_param(0.5)
        if case == 0:  # tie: same distance to the centre (0.25,0.25,0.25)
            base = np.array([[0.15, 0.25, 0.25], [3, 3, 3]], np.float32)
            new = np.array([[0.35, 0.25, 0.25]], np.float32)
        elif case == 1:  # existing closer
            base = np.array([[0.24, 0.25, 0.25], [3, 3, 3]], np.float32)
            new = np.array([[0.45, 0.05, 0.25]], np.float32)
        else:  # two existing in the voxel, new farther than both
            base = np.array([[0.24, 0.25, 0.25], [0.20, 0.2, 0.2], [3, 3, 3]], np.float32)
            new = np.array([[0.45, 0.05, 0.45]], np.float32)
        om.build(base)
        t.build(base)
        om.add(new, True, 0.5)
        t.add_points(new, True)
        a, b = _rows_sorted(om.dump()[0]), _rows_sorted(t.flatten()[0])
        assert np.array_equal(a, b), (case, a, b)
        if case == 0:
            assert any((a == new[0]).all(1))
        if case == 2:
            assert len(a) == 2


def test_delete_boxes_equals_reference_tree(orc, need_ikd, small_cfg):
    mp = small_cfg["map"]
    om = orc.Map(1.0)
    om.build(mp)
    t = orc.IkdTree()
    t.build(mp)
    boxes = np.array([[-5, -5, -1, 5, 5, 3], [10, -30, -1, 30, 0, 10]], np.float32)
    assert om.delete_boxes(boxes) == t.delete_boxes(boxes)
    assert np.array_equal(_rows_sorted(om.dump()[0]), _rows_sorted(t.flatten()[0]))
    q = mp[::50]
    oi, od, _ = om.knn(q)
    _, td = _bounded(*t.knn(q)[:2])
    assert np.array_equal(od.view(np.uint32), td.view(np.uint32))


def test_update_same_with_either_knn_backend(orc, need_ikd, small_cfg):
    """The whole oracle update gives identical bits with its own map or with the reference tree underneath."""
    cfg = small_cfg
    s = cfg["scan"]
    pts5 = np.concatenate([s[:, :3], np.zeros((len(s), 1), np.float32), s[:, 3:4]], 1)
    body = orc.voxel_grid(pts5, 0.5)[0][:, :3]
    om = orc.Map(1.0)
    om.build(cfg["map"])
    t = orc.IkdTree()
    t.build(cfg["map"])
    a = orc.Scan(body).update(cfg["x_prior"], cfg["P"], om.knn_backend(), 0.001, 4, False)
    b = orc.Scan(body).update(cfg["x_prior"], cfg["P"], t.knn_backend(), 0.001, 4, False, threads=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]
