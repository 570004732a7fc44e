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


def _livox_records(scan, rng):
    """livox_ros_driver/CustomPoint on the wire: offset_time(u32 ns) x y z reflectivity(u8) tag(u8) line(u8): 19 bytes."""
    n = len(scan)
    rec = np.zeros((n, 19), np.uint8)
    rec[:, 0:4] = (scan[:, 3].astype(np.float64) * 1e6).astype("<u4").view(np.uint8).reshape(n, 4)  # ms -> ns
    rec[:, 4:16] = scan[:, :3].astype("<f4").view(np.uint8).reshape(n, 12)
    rec[:, 16] = rng.integers(0, 256, n)
    rec[:, 17] = rng.choice([0x00, 0x10, 0x20, 0x30, 0x05, 0x14], n, p=[0.5, 0.3, 0.05, 0.05, 0.05, 0.05])
    rec[:, 18] = rng.integers(0, 8, n)  # lines 6 and 7 fall outside N_SCANS = 6
    return rec


@pytest.mark.parametrize("pfn", [1, 3])
def test_decode_livox_custom_points(ctx, orc, small_cfg, pfn):
    """avia_handler (preprocess.cpp:160-183): tag / line test, decimation over the VALID records, the `differs from
    pl_full[i - 1]` test with its zero-initialised predecessor and `a || b || c && d` precedence."""
    from agi_lidar_slam_b200 import _cabi

    rng = np.random.default_rng(11)
    scan = small_cfg["scan"].copy()
    scan[:, 3] = np.linspace(0.0, 99.9, len(scan), dtype=np.float32)  # offset_time, ms
    scan[::41, :3] *= np.float32(0.02)      # inside the blind zone
    dup = np.arange(5, len(scan), 29)
    scan[dup, :3] = scan[dup - 1, :3]       # exact repeats of the previous record
    dz = np.arange(7, len(scan), 53)
    scan[dz, :2] = scan[dz - 1, :2]         # differs in z only: kept only outside the blind zone
    rec = _livox_records(scan, rng)
    lay = _cabi.CloudLayout(19, 4, 8, 12, 16, 0, 1, pfn, 3, 0.0, 2.0, off_ring=18, ring_type=1, off_tag=17,
                            intensity_type=1, n_scans=6)
    ref, ref_i = orc.decode_avia(rec, 19, 4, 8, 12, 16, 0, 17, 18, 6, 2.0, pfn)
    nd, m = ctx.scan_preprocess_cloud2(rec, lay, None, None, 0.5)
    got, got_i = ctx.scan_decoded()
    assert nd == len(ref) and 100 < nd < len(scan)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(got_i, ref_i)
    assert 0 < m <= nd


def _ring_records(scan, rings, rng, kind):
    """velodyne_ros::Point (32 B: x y z pad intensity(f32) time(f32) ring(u16)) or rslidar_ros::Point (32 B: x y z pad
    intensity(u8) ring(u16 @18) timestamp(f64 @24)); time fields left at zero."""
    n = len(scan)
    rec = np.zeros((n, 32), np.uint8)
    rec[:, 0:12] = scan[:, :3].astype("<f4").view(np.uint8).reshape(n, 12)
    if kind == "velodyne":
        rec[:, 16:20] = rng.uniform(0, 255, n).astype("<f4").view(np.uint8).reshape(n, 4)
        rec[:, 24:26] = rings.astype("<u2").view(np.uint8).reshape(n, 2)
    else:
        rec[:, 16] = rng.integers(0, 256, n)
        rec[:, 18:20] = rings.astype("<u2").view(np.uint8).reshape(n, 2)
    return rec


@pytest.mark.parametrize("kind", ["velodyne", "rs"])
def test_decode_times_from_yaw_when_driver_gives_none(ctx, orc, kind):
    """given_offset_time == false (velodyne_handler :395-421, rs_handler :886-912): per-ring times from the azimuth.  The
    kept set, its order and the intensities are exact; yaw comes from atan2f, whose last bit differs between libm and
    CUDA, so the times agree to 1e-5 ms (1 ulp of yaw is ~3e-6 ms) instead of bitwise."""
    from agi_lidar_slam_b200 import _cabi

    rng = np.random.default_rng(21)
    n_rings, cols = 16, 900
    az = -(np.arange(cols) / cols) * 2 * np.pi * 1.08 + 0.3  # clockwise, a little more than one revolution
    el = np.deg2rad(np.linspace(-15, 15, n_rings))
    r = rng.uniform(0.5, 60.0, (cols, n_rings)).astype(np.float32)
    x = (r * np.cos(el)[None, :] * np.cos(az)[:, None]).astype(np.float32)
    y = (r * np.cos(el)[None, :] * np.sin(az)[:, None]).astype(np.float32)
    z = (r * np.sin(el)[None, :]).astype(np.float32)
    scan = np.stack([x.ravel(), y.ravel(), z.ravel()], 1)  # firing order: the rings cycle column by column
    rings = np.tile(np.arange(n_rings), cols)
    rec = _ring_records(scan, rings, rng, kind)
    if kind == "velodyne":
        lay = _cabi.CloudLayout(32, 0, 4, 8, 16, 20, 0, 2, 2, 1e3, 2.0, off_ring=24, ring_type=0, n_scans=n_rings,
                                scan_rate=10, yaw_time=1)
        ref, ref_i = orc.decode_yaw_times(rec, 32, 0, 4, 8, 16, "<f4", 24, "<u2", n_rings, 10, 2.0, 2)
    else:
        lay = _cabi.CloudLayout(32, 0, 4, 8, 16, 24, 3, 2, 4, 1.0, 2.0, off_ring=18, ring_type=0, intensity_type=1,
                                n_scans=n_rings, scan_rate=10, yaw_time=1)
        ref, ref_i = orc.decode_yaw_times(rec, 32, 0, 4, 8, 16, "u1", 18, "<u2", n_rings, 10, 2.0, 2)
    nd, _ = ctx.scan_preprocess_cloud2(rec, lay, None, None, 0.5)
    got, got_i = ctx.scan_decoded()
    assert nd == len(ref) and nd > 1000
    assert np.array_equal(got[:, :3].view(np.uint32), ref[:, :3].view(np.uint32))
    assert np.array_equal(got_i, ref_i)
    assert np.abs(got[:, 3] - ref[:, 3]).max() < 1e-5
    assert ref[:, 3].max() > 100.0  # more than one revolution: the +360 / omega_l branch was taken
    # a ring index the per-ring state cannot hold is an error, not an out-of-bounds access
    bad = rec.copy()
    bad[5, 24 if kind == "velodyne" else 18] = 200
    with pytest.raises(_cabi.LioError):
        ctx.scan_preprocess_cloud2(bad, lay, None, None, 0.5)


def test_decode_rs_with_timestamps(ctx, orc, small_cfg):
    """rs_handler with point timestamps: FP64 (t - t0) * 1000 rounded to FP32, uint8 intensity; yaw_time = 1 changes
    nothing because the last record's timestamp is positive."""
    from agi_lidar_slam_b200 import _cabi

    rng = np.random.default_rng(31)
    scan = small_cfg["scan"].copy()
    scan[::37, :3] *= np.float32(0.02)
    rec = _ring_records(scan, rng.integers(0, 32, len(scan)), rng, "rs")
    ts = 1.7e9 + scan[:, 3].astype(np.float64) / 1000.0 + 0.05
    rec[:, 24:32] = ts.astype("<f8").view(np.uint8).reshape(len(scan), 8)
    ref, ref_i = orc.decode_rs(rec, 32, 0, 4, 8, 16, 24, 2.0, 3)
    for yaw_time in (0, 1):
        lay = _cabi.CloudLayout(32, 0, 4, 8, 16, 24, 3, 3, 4, 1.0, 2.0, off_ring=18, ring_type=0, intensity_type=1,
                                n_scans=32, scan_rate=10, yaw_time=yaw_time)
        nd, _ = ctx.scan_preprocess_cloud2(rec, lay, None, None, 0.5)
        got, got_i = ctx.scan_decoded()
        assert nd == len(ref) and nd > 100
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(got_i, ref_i)
