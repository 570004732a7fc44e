"""PCD map hand-over and odometry message: round trips and layout checks on CPU; map reload into the
CUDA map on the GPU."""
import numpy as np
import pytest

from agi_lidar_slam_b200 import formats


def test_pcd_binary_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    rec = np.zeros((1000, 12), np.float32)
    rec[:, :3] = rng.normal(0, 20, (1000, 3))
    rec[:, 3] = 1.0
    rec[:, 8] = rng.uniform(0, 255, 1000)
    rec[:, 9] = rng.uniform(0, 100, 1000)
    p = tmp_path / "GlobalMap_ikdtree.pcd"
    formats.write_pcd_binary(p, rec)
    raw = p.read_bytes()
    head, _, body = raw.partition(b"DATA binary\n")
    assert b"FIELDS x y z intensity normal_x normal_y normal_z curvature" in head and b"POINTS 1000" in head
    assert len(body) == 1000 * 32  # packed fields, no struct padding
    assert np.array_equal(np.frombuffer(body, "<f4").reshape(1000, 8)[:, 3], rec[:, 8])
    back = formats.read_pcd(p)
    assert np.array_equal(back, rec)
    # xyz-only input, and an ascii file with a padding field as PCL writes for other point types
    formats.write_pcd_binary(p, rec[:, :3])
    assert np.array_equal(formats.read_pcd(p)[:, :3], rec[:, :3])
    q = tmp_path / "a.pcd"
    q.write_text("VERSION 0.7\nFIELDS x y z _ intensity\nSIZE 4 4 4 1 4\nTYPE F F F U F\nCOUNT 1 1 1 4 1\nWIDTH 2\nHEIGHT 1\n"
                 "POINTS 2\nDATA ascii\n1 2 3 0 0 0 0 7\n4 5 6 0 0 0 0 8\n")
    a = formats.read_pcd(q)
    assert a[:, :3].tolist() == [[1, 2, 3], [4, 5, 6]] and a[:, 8].tolist() == [7, 8]


def test_odometry_message_index_swap():
    P = np.arange(576, dtype=np.float64).reshape(24, 24)
    x = np.zeros(26)
    x[0:3] = [1, 2, 3]
    x[3:7] = [0.5, 0.1, 0.2, 0.3]  # w x y z
    m = formats.odometry_message(x, P)
    assert m["position"].tolist() == [1, 2, 3] and m["orientation_xyzw"].tolist() == [0.1, 0.2, 0.3, 0.5]
    c = m["covariance"].reshape(6, 6)
    # ROS row 0 (x) takes filter row 3 (rot x) exactly as the reference's k = i < 3 ? i + 3 : i - 3 does
    assert c[0].tolist() == [P[3, 3], P[3, 4], P[3, 5], P[3, 0], P[3, 1], P[3, 2]]
    assert c[4].tolist() == [P[1, 3], P[1, 4], P[1, 5], P[1, 0], P[1, 1], P[1, 2]]


@pytest.mark.gpu
def test_map_survives_pcd_hand_over(ctx, orc, small_cfg, tmp_path):
    """flatten -> GlobalMap_ikdtree.pcd -> loadPCDFile -> Build (relocalisation mode): same searches afterwards."""
    mp = small_cfg["map"]
    ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    q = orc.body_to_world(small_cfg["x_prior"], small_cfg["scan"][::5, :3])
    a = ctx.knn5(q)
    xyz, _ = ctx.map_dump()
    formats.write_pcd_binary(tmp_path / "GlobalMap_ikdtree.pcd", xyz)
    ctx.map_build(formats.read_pcd(tmp_path / "GlobalMap_ikdtree.pcd"))
    b = ctx.knn5(q)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
