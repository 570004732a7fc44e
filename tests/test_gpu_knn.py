"""kNN stage parity (SURVEY.md App. C row "kNN"): ids, d2 and gate-1 mask bit-exact against the oracle, whose own
kNN is pinned against the reference ikd-Tree in test_oracle_map.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _queries(cfg, orc):
    return orc.body_to_world(cfg["x_prior"], cfg["scan"][:, :3])


@pytest.mark.parametrize("which", ["small", "avia"])
def test_knn5_matches_oracle_bit_exact(ctx, orc, small_cfg, avia_cfg, which):
    cfg = small_cfg if which == "small" else avia_cfg
    mp = cfg["map"]
    q = _queries(cfg, orc)
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    idx, d2, nbr = ctx.knn5(q)
    om = orc.Map(1.0)
    om.build(mp)
    oi, od, on = om.knn(q, 5, 5.0, threads=8)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(nbr.view(np.uint32), on.view(np.uint32))
    # gate 1 (esekfom.hpp:144-147)
    assert np.array_equal(idx[:, 4] >= 0, oi[:, 4] >= 0)
    assert (idx[:, 4] >= 0).mean() > 0.5


def test_knn5_matches_reference_ikd_tree(ctx, orc, small_cfg):
    """Directly against the reference's own tree (unbounded search cut at d2 <= 5), ties aside."""
    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not present")
    cfg = small_cfg
    mp = cfg["map"]
    q = _queries(cfg, orc)
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    idx, d2, _ = ctx.knn5(q)
    t = orc.IkdTree()
    t.build(mp)
    ti, td, _ = t.knn(q, 5, threads=4)
    td = np.where(td <= 5.0, td, np.inf).astype(np.float32)
    ti = np.where(np.isfinite(td), ti, -1)
    assert np.array_equal(d2.view(np.uint32), td.view(np.uint32))
    assert np.array_equal(idx, ti)


def test_knn_edge_cases(ctx, orc):
    rng = np.random.default_rng(5)
    # tiny map (< 5 points): fewer than k results, like the reference on a 3-point tree
    mp = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
    ctx.map_build(np.concatenate([mp, np.zeros((3, 1), np.float32)], 1))
    q = rng.uniform(-1, 1, (7, 3)).astype(np.float32)
    idx, d2, _ = ctx.knn5(q)
    om = orc.Map(1.0)
    om.build(mp)
    oi, od, _ = om.knn(q)
    assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert (idx[:, 3:] == -1).all()
    # exact ties: lattice points at identical distances are ordered by id on both sides
    g = np.stack(np.meshgrid(*[np.arange(-2, 3)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.5
    ctx.map_build(np.concatenate([g, np.zeros((len(g), 1), np.float32)], 1))
    q = np.zeros((1, 3), np.float32)
    idx, d2, _ = ctx.knn5(q)
    om = orc.Map(1.0)
    om.build(g)
    oi, od, _ = om.knn(q)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # far queries: nothing within sqrt(5) m
    q = np.full((4, 3), 100.0, np.float32)
    idx, d2, _ = ctx.knn5(q)
    assert (idx == -1).all() and np.isinf(d2).all()
    # empty query batch
    idx, d2, _ = ctx.knn5(np.zeros((0, 3), np.float32))
    assert idx.shape == (0, 5)


def test_sparse_map_needs_ring_expansion(ctx, orc):
    """Neighbours between 1 m and sqrt(5) m away are found by the widened search (stage 2)."""
    rng = np.random.default_rng(11)
    mp = (rng.uniform(-20, 20, (4000, 3)) * np.array([1, 1, 0.05])).astype(np.float32)  # ~2.5 pts / m^2 sheet
    mp = mp[::6]  # sparse: typical 5th neighbour > 1 m
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    q = (rng.uniform(-18, 18, (3000, 3)) * np.array([1, 1, 0.05])).astype(np.float32)
    idx, d2, _ = ctx.knn5(q)
    om = orc.Map(1.0)
    om.build(mp)
    oi, od, _ = om.knn(q, threads=8)
    assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    far = np.isfinite(d2[:, 4]) & (d2[:, 4] > 1.0)
    assert far.sum() > 100


def _p4(xyz):
    return np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1)


def test_knn5_unbounded_matches_reference_tree(ctx, orc, small_cfg):
    """Nearest_Search with its default max_dist = INFINITY (ikd_Tree.h:285; what esekfom.hpp:140-141 gets): five
    neighbours for every query however far from the map, bit-equal to the reference's own tree."""
    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not present")
    mp = small_cfg["map"]
    ctx.map_build(_p4(mp))
    tree = orc.IkdTree()
    tree.build(mp)
    rng = np.random.default_rng(5)
    lo, hi = mp.min(0), mp.max(0)
    q = np.concatenate([
        rng.uniform(lo - 30, hi + 30, (3000, 3)),           # around and outside the map
        rng.uniform(lo, hi, (1000, 3)) + [0, 0, 40.0],       # high above it
        mp[rng.integers(0, len(mp), 500)] + rng.normal(0, 0.3, (500, 3)),
        np.array([[900.0, -700.0, 300.0], [1e4, -2e4, 3e3]]),  # hundreds of cells away: the box-sweep path
    ]).astype(np.float32)
    gi, gd, gx = ctx.knn5(q, max_d2=np.inf)
    ti, td, tx = tree.knn(q, 5, np.inf, threads=4)
    assert np.all(gi >= 0) and np.all(np.isfinite(gd))
    assert np.array_equal(gd.view(np.uint32), td.view(np.uint32))
    same = gi == ti
    assert same.mean() > 0.999  # the tree's order among exactly equal distances is traversal-dependent
    assert np.array_equal(gx[same].view(np.uint32), tx[same].view(np.uint32))
    # an intermediate bound: the prefix of the unbounded list, cut with `dist <= max_dist^2` (ikd_Tree.cpp:980)
    bi, bd, _ = ctx.knn5(q, max_d2=40.0)
    inside = gd <= 40.0
    assert np.array_equal(bi[inside], gi[inside]) and np.all(bi[~inside] == -1) and np.all(np.isinf(bd[~inside]))
    # after a box delete the cell box of the map does not shrink; the search must still be exact
    box = np.array([[lo[0] - 1, lo[1] - 1, lo[2] - 1, (lo[0] + hi[0]) / 2, hi[1] + 1, hi[2] + 1]], np.float32)
    assert ctx.map_delete_boxes(box) == tree.delete_boxes(box) > 0
    gi, gd, _ = ctx.knn5(q, max_d2=np.inf)
    ti, td, _ = tree.knn(q, 5, np.inf, threads=4)
    assert np.array_equal(gd.view(np.uint32), td.view(np.uint32)) and (gi == ti).mean() > 0.999


def test_knn5_unbounded_tiny_map(ctx, orc):
    """Fewer than five live points: the reference returns what there is, so does the unbounded search."""
    mp = np.array([[0, 0, 0], [10, 0, 0], [0, 30, 1]], np.float32)
    ctx.map_build(_p4(mp))
    q = np.array([[1, 1, 1], [100, 100, 100], [-50, 3, 0]], np.float32)
    gi, gd, _ = ctx.knn5(q, max_d2=np.inf)
    om = orc.Map(1.0)
    om.build(mp)
    oi, od, _ = om.knn(q, 5, np.inf)
    assert np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert np.all(gi[:, :3] >= 0) and np.all(gi[:, 3:] == -1)
