"""Hand-checked cases for the oracle's restatement of the sensor handlers (src/preprocess.cpp) -- the checker of
tests/test_gpu_decode.py must itself follow the reference's control flow, quirks included."""
import numpy as np

from oracle import pyoracle as orc


def _livox(rows):
    rec = np.zeros((len(rows), 19), np.uint8)
    for i, (t_ns, x, y, z, refl, tag, line) in enumerate(rows):
        rec[i, 0:4] = np.array([t_ns], "<u4").view(np.uint8)
        rec[i, 4:16] = np.array([x, y, z], "<f4").view(np.uint8)
        rec[i, 16:19] = (refl, tag, line)
    return rec


def test_avia_handler_control_flow():
    rows = [
        (0, 9.0, 9.0, 9.0, 1, 0x00, 0),        # record 0: the loops start at 1, never looked at
        (1000000, 5.0, 0.0, 0.0, 10, 0x10, 1),  # valid #1, kept (differs from the zero point in x)
        (2000000, 5.0, 0.0, 0.0, 20, 0x00, 2),  # valid #2, exact repeat of record 1: dropped
        (3000000, 5.0, 0.0, 0.5, 30, 0x00, 9),  # line >= N_SCANS: not valid, pl_full[3] stays zero
        (4000000, 0.0, 0.0, 0.5, 40, 0x00, 3),  # valid #3: differs from zero only in z, range 0.25 < blind^2: dropped
        (5000000, 0.0, 0.0, 3.0, 50, 0x20, 3),  # tag 0x20: not valid
        (6000000, 0.0, 0.0, 3.0, 60, 0x14, 3),  # valid #4 (0x14 & 0x30 == 0x10): z differs from zero and range 9 > 4: kept
        (7000000, 0.1, 0.0, 0.0, 70, 0x00, 0),  # valid #5: x differs -> kept even inside the blind zone (a || b || c && d)
    ]
    xyzt, inten = orc.decode_avia(_livox(rows), 19, 4, 8, 12, 16, 0, 17, 18, 6, 2.0, 1)
    assert xyzt[:, :3].tolist() == [[5.0, 0.0, 0.0], [0.0, 0.0, 3.0], [np.float32(0.1), 0.0, 0.0]]
    assert xyzt[:, 3].tolist() == [1.0, 6.0, 7.0]  # offset_time / 1e6: ns -> ms
    assert inten.tolist() == [10.0, 60.0, 70.0]
    # point_filter_num 2 counts VALID records: #2 (repeat of an unwritten predecessor -> compared with zero -> kept), #4
    xyzt, inten = orc.decode_avia(_livox(rows), 19, 4, 8, 12, 16, 0, 17, 18, 6, 2.0, 2)
    assert inten.tolist() == [20.0, 60.0]


def test_yaw_times_control_flow():
    # two rings, clockwise rotation (yaw decreases), 10 Hz: omega_l = 3.61 deg/ms
    deg = np.deg2rad
    pts = [  # (ring, yaw_deg, range)
        (0, 10.0, 5.0),    # first of ring 0: fixes yaw_fp, skipped
        (1, 10.0, 5.0),    # first of ring 1: skipped
        (0, 0.0, 5.0),     # (10 - 0) / 3.61
        (1, 20.0, 5.0),    # yaw > yaw_fp: (10 - 20 + 360) / 3.61
        (0, -170.0, 1.0),  # inside the blind zone: time state advances, point dropped
        (1, 15.0, 5.0),    # (10 - 15 + 360) / 3.61 = 98.3 < previous 96.95? no: 98.3 > 96.95, no wrap
        (0, 5.0, 5.0),     # (10 - 5) / 3.61 = 1.385 < previous 49.86 -> + 360 / 3.61
    ]
    rec = np.zeros((len(pts), 32), np.uint8)
    for i, (ring, yaw, r) in enumerate(pts):
        rec[i, 0:12] = np.array([r * np.cos(deg(yaw)), r * np.sin(deg(yaw)), 0.0], "<f4").view(np.uint8)
        rec[i, 16:20] = np.array([float(i)], "<f4").view(np.uint8)
        rec[i, 24:26] = np.array([ring], "<u2").view(np.uint8)
    xyzt, inten = orc.decode_yaw_times(rec, 32, 0, 4, 8, 16, "<f4", 24, "<u2", 2, 10, 2.0, 1)
    assert inten.tolist() == [2.0, 3.0, 5.0, 6.0]
    k = 57.2957 / 57.29577951308232  # the handler's rounded rad -> deg factor
    want = [10 * k / 3.61, (-10 * k + 360) / 3.61, (-5 * k + 360) / 3.61, 5 * k / 3.61 + 360 / 3.61]
    assert np.allclose(xyzt[:, 3], want, rtol=0, atol=2e-4)
    # decimation is by RECORD index, after the per-ring state has advanced
    xyzt2, inten2 = orc.decode_yaw_times(rec, 32, 0, 4, 8, 16, "<f4", 24, "<u2", 2, 10, 2.0, 2)
    assert inten2.tolist() == [2.0, 6.0] and np.array_equal(xyzt2[:, 3], xyzt[[0, 3], 3])


def test_rs_timestamps():
    rec = np.zeros((4, 32), np.uint8)
    ts = np.array([1000.25, 1000.26, 1000.27, 1000.35])
    for i in range(4):
        rec[i, 0:12] = np.array([3.0 + i, 0.0, 0.0], "<f4").view(np.uint8)
        rec[i, 16] = 7 * i
        rec[i, 24:32] = ts[i:i + 1].astype("<f8").view(np.uint8)
    xyzt, inten = orc.decode_rs(rec, 32, 0, 4, 8, 16, 24, 2.0, 1)
    assert np.array_equal(xyzt[:, 3], ((ts - ts[0]) * 1000.0).astype(np.float32))
    assert inten.tolist() == [0.0, 7.0, 14.0, 21.0]
