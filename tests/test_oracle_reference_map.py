"""The oracle's main loop with the REFERENCE ikd-Tree as its live map.

The reference calls Nearest_Search without a distance bound (esekfom.hpp:140-141; default max_dist = INFINITY,
ikd_Tree.h:285): Nearest_Points[i] always holds min(5, #live points) neighbours, and map_incremental
(laserMapping.cpp:382-433) reads them for points that failed the d2[4] <= 5 gate as well -- points_near[0] decides
PointNoNeedDownsample, all five decide need_add.  These tests pin the oracle's restated loop (hashed-grid Map port,
unbounded search) to the loop run on the reference's own tree -- Build / unbounded Nearest_Search / both Add_Points
calls / flatten -- on a SPARSE scene where a good part of every scan sits on the map frontier, scan by scan:
map_incremental classes, the whole map after the insert, and the filter state."""
import numpy as np
import pytest

N_SCANS = 36


def sparse_sequence(n_scans=N_SCANS, seed=2002):
    """A 240 m yard seen by a 16-ring sensor with 900 columns and 200 m range: far walls are hit sparsely, so several
    per cent of the downsampled points have fewer than five map points within sqrt(5) m, many of them none."""
    from agi_lidar_slam_b200 import synth

    return synth.sequence(n_scans, seed, rings=16, cols=900, scene=synth.block_scene(extent=240.0), max_range=200.0)


def _same_map(a, b):
    ax, ai = a.dump()
    bx, bi = b.dump()
    return ai.shape == bi.shape and np.array_equal(ai, bi) and np.array_equal(ax.view(np.uint32), bx.view(np.uint32))


def test_loop_on_reference_tree_sparse_scene(orc):
    from replay_oracle import OracleReplay

    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not built (needs /root/reference once)")
    seq = sparse_sequence()
    port = OracleReplay(orc, max_iteration=3, threads=8)
    ref = OracleReplay(orc, max_iteration=3, threads=8, use_ikd=True)
    rows = gate1_fail = none_near = n_upd = bounded_would_differ = 0
    for k, m in enumerate(seq):
        xa, xb = port.process(m), ref.process(m)
        assert (xa is None) == (xb is None), k
        if xa is None:
            continue
        n_upd += 1
        a, b = port.last, ref.last
        # the unbounded neighbour rows themselves (ids are compared through d2 and the classes; the tree's order among
        # exactly equal distances is traversal-dependent, the port's is by id)
        assert np.array_equal(a["cnt"], b["cnt"]) and np.array_equal(a["d2"].view(np.uint32), b["d2"].view(np.uint32)), k
        assert np.all(a["cnt"] == 5)  # the map holds >= 5 points: every row is full, however far the neighbours are
        assert np.array_equal(a["cls"], b["cls"]), k
        assert port.log[-1]["counts"] == ref.log[-1]["counts"], k
        assert _same_map(port.map, ref.map), k
        assert np.array_equal(xa, xb) and np.array_equal(port.P, ref.P), k
        d2 = a["d2"]
        rows += len(d2)
        gate1_fail += int((d2[:, 4] > 5).sum())
        none_near += int((d2[:, 0] > 5).sum())
        # what a neighbour cache cut at d2 <= 5 would have classified (round 1's deviation): it must differ somewhere
        # on this scene, or the scene does not exercise the rule
        cut = (d2 <= 5).sum(1).astype(np.int32)
        cls_cut = orc.map_incremental_classify(a["world"], a["near_raw"], cut, True, 0.5)
        bounded_would_differ += int((cls_cut != a["cls"]).sum())
    assert n_upd >= N_SCANS - 4
    assert gate1_fail >= 0.05 * rows, (gate1_fail, rows)  # >= 5 % of the points fail gate 1 ...
    assert none_near >= 0.02 * rows, (none_near, rows)     # ... and many have no map point within sqrt(5) m at all
    assert bounded_would_differ > 0
    print(f"{n_upd} updates on the reference tree: {rows} rows, {gate1_fail} fail gate 1 ({100 * gate1_fail / rows:.1f} %), "
          f"{none_near} without a neighbour within sqrt(5) m; a cache cut at d2 <= 5 would misclassify "
          f"{bounded_would_differ} points; map {port.map.size()} points")


def test_unbounded_knn_port_equals_reference_tree(orc, small_cfg):
    """Map.knn(max_d2 = inf) against Nearest_Search(max_dist = INFINITY) on far-away and frontier queries."""
    if not orc.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not built")
    mp = small_cfg["map"]
    om = orc.Map(1.5)
    om.build(mp)
    tree = orc.IkdTree()
    tree.build(mp)
    rng = np.random.default_rng(5)
    lo, hi = mp.min(0), mp.max(0)
    q = np.concatenate([
        rng.uniform(lo - 30, hi + 30, (400, 3)),          # around and outside the map
        rng.uniform(lo, hi, (200, 3)) + [0, 0, 40.0],      # high above it
        mp[rng.integers(0, len(mp), 100)] + rng.normal(0, 0.3, (100, 3)),
        np.array([[1e4, -2e4, 3e3]]),                      # absurdly far: still five neighbours
    ]).astype(np.float32)
    gi, gd, _ = om.knn(q, 5, np.inf, threads=4)
    ti, td, _ = tree.knn(q, 5, np.inf, threads=4)
    assert np.all(gi >= 0) and np.all(np.isfinite(gd))
    assert np.array_equal(gd.view(np.uint32), td.view(np.uint32))
    assert (gi == ti).mean() > 0.999  # exact d2 ties are the only way the id order can differ
    # a bounded search returns the prefix of the unbounded one
    bi, bd, _ = om.knn(q, 5, 5.0, threads=4)
    for r in range(5):
        inside = gd[:, r] <= 5
        assert np.array_equal(bi[inside, r], gi[inside, r]) and np.all(bi[~inside, r] == -1)
