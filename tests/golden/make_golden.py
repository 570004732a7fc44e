#!/usr/bin/env python
"""Generates tests/golden/ikd_reference.npz by RUNNING THE REFERENCE'S OWN ikd-Tree (compiled in place from
/root/reference into oracle/_ref/libikd_ref.so — this only works in the authoring container).  The fixture pins the
oracle and the CUDA path to outputs of the unmodified reference on machines where /root/reference does not exist.

    python tests/golden/make_golden.py

Contents (all inputs are included, so the fixture is self-contained):
  map, queries                  float32 (4000,3), (600,3)
  knn_d2, knn_xyz               reference Nearest_Search(k=5): ascending squared distances (inf-padded beyond
                                d2 > 5, the bound the path uses) and the neighbours' coordinates
  add_batches / add_downsample  three Add_Points batches with the downsample flag of each
  add_returned                  what Add_Points returned for each batch
  flat_after_add                flatten() after the three batches, rows sorted lexicographically
  del_boxes, del_returned, flat_after_delete   Delete_Point_Boxes then flatten()
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as orc  # noqa: E402


def rows_sorted(a):
    return a[np.lexsort(a.T[::-1])]


def main():
    orc.build()
    assert orc.ikd_available(), "needs /root/reference (reference ikd-Tree compiled in place)"
    rng = np.random.default_rng(20261018)
    # a 20 m x 20 m floor with a wall, ~1 point per 0.3 m: plenty of 5-NN within sqrt(5) m, some queries far away
    floor = np.c_[rng.uniform(-10, 10, (3000, 2)), rng.normal(0, 0.01, 3000)]
    wall = np.c_[np.full(1000, 6.0) + rng.normal(0, 0.01, 1000), rng.uniform(-10, 10, 1000), rng.uniform(0, 4, 1000)]
    mp = np.concatenate([floor, wall]).astype(np.float32)
    q = np.concatenate([np.c_[rng.uniform(-11, 11, (500, 2)), rng.uniform(-0.2, 3, 500)],
                        rng.uniform(20, 30, (100, 3))]).astype(np.float32)
    t = orc.IkdTree()
    t.set_downsample_param(0.5)
    t.build(mp)
    ti, td, tn = t.knn(q, 5, threads=1)
    td = np.where(td <= 5.0, td, np.inf).astype(np.float32)
    tn = np.where(np.isfinite(td)[..., None], tn, 0).astype(np.float32)
    batches, flags, returned = [], [], []
    for k, ds in enumerate([True, False, True]):
        b = np.c_[rng.uniform(-12, 12, (800, 2)), rng.normal(0, 0.02, 800)].astype(np.float32)
        batches.append(b)
        flags.append(ds)
        returned.append(t.add_points(b, ds))
    flat_add = rows_sorted(t.flatten()[0])
    boxes = np.array([[-3, -3, -1, 2, 2, 1], [5.5, -10, 0, 6.5, 0, 4]], np.float32)
    ndel = t.delete_boxes(boxes)
    flat_del = rows_sorted(t.flatten()[0])
    out = Path(__file__).resolve().parent / "ikd_reference.npz"
    np.savez_compressed(out, map=mp, queries=q, knn_d2=td, knn_xyz=tn, add_batches=np.stack(batches),
                        add_downsample=np.array(flags), add_returned=np.array(returned, np.int64), flat_after_add=flat_add,
                        del_boxes=boxes, del_returned=np.int64(ndel), flat_after_delete=flat_del)
    print("wrote", out, out.stat().st_size, "bytes; returned", returned, "deleted", ndel, "valid 5-NN queries",
          int(np.isfinite(td[:, 4]).sum()))


if __name__ == "__main__":
    main()
