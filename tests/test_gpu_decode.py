"""Sensor decoding on the device: PointCloud2 records of an Ouster and of a Velodyne driver go
through lio_scan_preprocess_cloud2 and come out as the cloud oust64_handler / velodyne_handler would have produced
(same points, same order, same FP32 times), then through the same voxel filter as host-decoded points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ouster_records(scan, rng):
    """ouster_ros::Point: x y z (pad) intensity t(u32 ns) reflectivity(u16) ring(u8) ambient(u16) range(u32): 48 bytes."""
    n = len(scan)
    rec = np.zeros((n, 48), np.uint8)
    rec[:, 0:12] = scan[:, :3].astype("<f4").view(np.uint8).reshape(n, 12)
    rec[:, 16:20] = rng.uniform(0, 255, n).astype("<f4").view(np.uint8).reshape(n, 4)
    rec[:, 20:24] = (scan[:, 3].astype(np.float64) * 1e6).astype("<u4").view(np.uint8).reshape(n, 4)  # ms -> ns
    return rec


def _velodyne_records(scan, rng):
    """velodyne_ros::Point: x y z (pad) intensity time(f32 s) ring(u16): 32 bytes."""
    n = len(scan)
    rec = np.zeros((n, 32), np.uint8)
    rec[:, 0:12] = scan[:, :3].astype("<f4").view(np.uint8).reshape(n, 12)
    rec[:, 16:20] = rng.uniform(0, 255, n).astype("<f4").view(np.uint8).reshape(n, 4)
    rec[:, 20:24] = (scan[:, 3] / np.float32(1000)).astype("<f4").view(np.uint8).reshape(n, 4)  # ms -> s
    return rec


@pytest.mark.parametrize("sensor", ["ouster", "velodyne"])
def test_decode_matches_reference_handlers(ctx, orc, small_cfg, sensor):
    from agi_lidar_slam_b200 import _cabi

    rng = np.random.default_rng(4)
    scan = small_cfg["scan"].copy()
    scan[::37, :3] *= np.float32(0.02)  # some returns inside the blind zone
    if sensor == "ouster":
        rec = _ouster_records(scan, rng)
        lay = _cabi.CloudLayout(48, 0, 4, 8, 16, 20, 1, 3, 1, 1e-6, 2.0)  # t in ns -> ms, point_filter_num 3, blind 2 m
        ref, ref_i = orc.decode_cloud2(rec, 48, 0, 4, 8, 16, 20, "<u4", 1e-6, 2.0, 3, 1)
    else:
        rec = _velodyne_records(scan, rng)
        lay = _cabi.CloudLayout(32, 0, 4, 8, 16, 20, 0, 2, 2, 1e3, 2.0)  # time in s -> ms, point_filter_num 2
        ref, ref_i = orc.decode_cloud2(rec, 32, 0, 4, 8, 16, 20, "<f4", 1e3, 2.0, 2, 2)
    nd, m = ctx.scan_preprocess_cloud2(rec, lay, None, None, 0.5)
    got, got_i = ctx.scan_decoded()
    assert nd == len(ref) and 0 < nd < len(scan) // lay.point_filter_num + 1
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))  # same points, same order, same FP32 time
    assert np.array_equal(got_i, ref_i)
    # and the downsampled cloud equals the one obtained from host-decoded points
    rec12 = np.zeros((len(ref), 12), np.float32)
    rec12[:, :3] = ref[:, :3]
    rec12[:, 3] = 1.0
    rec12[:, 8] = ref_i
    rec12[:, 9] = ref[:, 3]
    out, _, _ = ctx.scan_preprocess(rec12, None, None, 0.5)
    assert m == len(out)
    cen, _, _ = orc.voxel_grid(np.concatenate([ref[:, :3], ref_i[:, None], ref[:, 3:4]], 1), 0.5)
    assert np.array_equal(out[:, :3].view(np.uint32), cen[:, :3].view(np.uint32))


def test_decode_rejects_bad_layouts(ctx):
    from agi_lidar_slam_b200 import _cabi

    rec = np.zeros((10, 32), np.uint8)
    for bad in (_cabi.CloudLayout(32, 0, 4, 30, -1, -1, 0, 1, 2, 1.0, 0.0),  # z field runs past the record
                _cabi.CloudLayout(32, 0, 4, 8, -1, -1, 0, 0, 2, 1.0, 0.0),   # point_filter_num 0
                _cabi.CloudLayout(32, 0, 4, 8, -1, -1, 0, 1, 7, 1.0, 0.0)):  # unknown rule
        with pytest.raises(_cabi.LioError) as e:
            ctx.scan_preprocess_cloud2(rec, bad)
        assert e.value.code == _cabi.LIO_E_INVALID
