"""Golden vectors produced by the UNMODIFIED reference ikd-Tree (tests/golden/make_golden.py, run where
/root/reference exists): they pin the oracle (CPU suite) and the CUDA path (gpu suite) on machines without the
reference.  Ties aside (none in this data), neighbour sets are compared through (d2 bits, coordinates)."""
from pathlib import Path

import numpy as np
import pytest

G = np.load(Path(__file__).resolve().parent / "golden" / "ikd_reference.npz")


def _rows_sorted(a):
    return a[np.lexsort(a.T[::-1])]


def _p4(xyz):
    return np.concatenate([xyz, np.zeros((len(xyz), 1), np.float32)], 1)


def _check_knn(d2, xyz):
    assert np.array_equal(d2.view(np.uint32), G["knn_d2"].view(np.uint32))
    ok = np.isfinite(G["knn_d2"])
    assert np.array_equal(xyz[ok].view(np.uint32), G["knn_xyz"][ok].view(np.uint32))
    assert np.isfinite(G["knn_d2"][:, 4]).sum() > 300 and np.isinf(G["knn_d2"][-100:]).all()


def _replay_map_ops(m_add, m_delete, m_dump):
    # The map CONTENTS after each call are what the path depends on.  (Add_Points' return value is a diagnostic the
    # reference only prints: it counts the insert operations of its sequential loop, re-insertions of surviving old
    # points included -- ikd_Tree.cpp:455-470 -- whereas lio_map_add reports the new points that stayed.)
    for b, ds in zip(G["add_batches"], G["add_downsample"]):
        m_add(b, bool(ds))
    assert np.array_equal(_rows_sorted(m_dump()).view(np.uint32), G["flat_after_add"].view(np.uint32))
    assert m_delete(G["del_boxes"]) == int(G["del_returned"])
    assert np.array_equal(_rows_sorted(m_dump()).view(np.uint32), G["flat_after_delete"].view(np.uint32))


def test_oracle_matches_reference_golden(orc):
    om = orc.Map(1.0)
    om.build(G["map"])
    _, d2, xyz = om.knn(G["queries"])
    _check_knn(d2, xyz)
    _replay_map_ops(lambda b, ds: om.add(b, ds, 0.5), om.delete_boxes, lambda: om.dump()[0])


@pytest.mark.gpu
def test_cuda_matches_reference_golden(ctx):
    ctx.map_build(_p4(G["map"]))
    _, d2, xyz = ctx.knn5(G["queries"])
    _check_knn(d2, xyz)
    _replay_map_ops(lambda b, ds: ctx.map_add(_p4(b), ds), ctx.map_delete_boxes, lambda: ctx.map_dump()[0])
