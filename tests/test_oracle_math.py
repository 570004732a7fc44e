"""Pins for the oracle's restated third-party arithmetic (Eigen / Sophus are absent: "parity unpinned" against the
real libraries, so the restatement is cross-checked against NumPy / SciPy and against algebraic identities)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation


def test_so3_exp_log_against_scipy(orc):
    rng = np.random.default_rng(1)
    for _ in range(200):
        w = rng.normal(size=3)
        w *= rng.uniform(1e-12, 3.1) / np.linalg.norm(w)  # |w| < pi: log returns the principal rotation vector
        q = orc.so3_exp(w)  # w,x,y,z
        R = orc.quat_to_mat(q)
        Rs = Rotation.from_rotvec(w).as_matrix()
        assert np.abs(R - Rs).max() < 1e-14
        assert np.abs(orc.so3_log(q) - w).max() < 1e-12
        assert abs(np.linalg.norm(q) - 1) < 1e-15
        q2 = orc.mat_to_quat(R)
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-13
        v = rng.normal(size=3)
        assert np.abs(orc.quat_rotate(q, v) - Rs @ v).max() < 1e-14


def test_so3_small_angle_and_pi(orc):
    assert np.array_equal(orc.so3_exp(np.zeros(3)), np.array([1.0, 0, 0, 0]))
    w = np.array([1e-12, -2e-12, 3e-13])
    assert np.allclose(orc.so3_log(orc.so3_exp(w)), w, rtol=1e-9, atol=0)
    w = np.array([np.pi - 1e-9, 0, 0])
    assert np.allclose(orc.so3_log(orc.so3_exp(w)), w, atol=1e-8)
    # trace <= 0 branch of matrix -> quaternion
    for axis in np.eye(3):
        R = Rotation.from_rotvec(axis * 3.1).as_matrix()
        q = orc.mat_to_quat(R)
        assert np.abs(orc.quat_to_mat(q) - R).max() < 1e-14


def test_boxplus_boxminus_round_trip(orc):
    rng = np.random.default_rng(2)
    x = orc.default_state()
    x[3:7] = orc.so3_exp(rng.normal(size=3))
    x[7:11] = orc.so3_exp(rng.normal(size=3) * 0.2)
    for _ in range(50):
        d = rng.normal(size=24) * 0.3
        y = orc.boxplus(x, d)
        assert np.abs(orc.boxminus(y, x) - d).max() < 1e-12
    assert np.abs(orc.boxminus(x, x)).max() < 1e-15


def test_inverse_against_numpy(orc):
    rng = np.random.default_rng(3)
    for n in (3, 12, 24):
        A = rng.normal(size=(n, n))
        Ai = orc.inverse(A)
        assert np.abs(Ai @ A - np.eye(n)).max() < 1e-10
        assert np.abs(Ai - np.linalg.inv(A)).max() <= 1e-9 * np.abs(Ai).max()
    # the shapes the filter inverts: P (diag-dominant covariance) and HTH/R + P^-1
    P = np.diag(np.r_[np.ones(6), np.full(6, 1e-5), np.ones(3), np.full(3, 1e-4), np.full(3, 1e-3), np.full(3, 1e-5)])
    Pi = orc.inverse(P)
    assert np.allclose(np.diag(Pi), 1 / np.diag(P), rtol=1e-15)
    H = rng.normal(size=(500, 12))
    A = np.zeros((24, 24))
    A[:12, :12] = H.T @ H / 0.001
    A += Pi
    assert np.abs(orc.inverse(A) @ A - np.eye(24)).max() < 1e-8
    # needs a row swap
    A = np.array([[0.0, 2.0], [3.0, 1.0]])
    assert np.allclose(orc.inverse(A), np.linalg.inv(A))


def test_qr_solve_against_lstsq(orc):
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(500):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        d = rng.uniform(1, 30)
        # five points near the plane n.p + d = 0
        basis = np.linalg.svd(n[None, :])[2][1:]
        pts = (-d * n)[None, :] + rng.uniform(-1, 1, (5, 2)) @ basis + rng.normal(0, 0.01, (5, 1)) * n[None, :]
        pts = pts.astype(np.float32)
        x, full = orc.qr_solve_5x3(pts)
        ref = np.linalg.lstsq(pts.astype(np.float64), -np.ones(5), rcond=None)[0]
        assert full
        worst = max(worst, np.abs(x - ref).max() / np.abs(ref).max())
    assert worst < 5e-3  # FP32 QR of a mildly conditioned 5x3 system vs FP64 SVD


def test_esti_plane_recipe(orc):
    pts = np.array([[0, 0, 2], [1, 0, 2], [0, 1, 2], [1, 1, 2.0], [0.5, 0.5, 2]], np.float32)
    pabcd, ok = orc.esti_plane(pts, 0.1)
    assert ok
    assert np.allclose(np.abs(pabcd[:3]), [0, 0, 1], atol=1e-6) and abs(abs(pabcd[3]) - 2) < 1e-5
    assert abs(pabcd[:3] @ pts[0] + pabcd[3]) < 1e-5  # sign convention: n.p + d = 0
    # 0.1 m inlier test (common_lib.h:127-132)
    bad = pts.copy()
    bad[4, 2] += 0.6
    assert not orc.esti_plane(bad, 0.1)[1]
    # collinear neighbours: rank deficient -> dependent component zeroed, no crash, rejected or degenerate
    line = np.array([[i, 2 * i, 1.0] for i in range(5)], np.float32)
    x, full = orc.qr_solve_5x3(line)
    assert np.all(np.isfinite(x))


def test_predict_against_dense_numpy(orc):
    rng = np.random.default_rng(5)
    x = orc.default_state()
    x[3:7] = orc.so3_exp(rng.normal(size=3))
    x[14:17] = rng.normal(size=3)
    x[17:23] = rng.normal(size=6) * 0.01
    A = rng.normal(size=(24, 24))
    P = A @ A.T * 1e-3
    Q = np.diag(rng.uniform(1e-5, 1e-1, 12))
    acc, gyr, dt = rng.normal(size=3) + [0, 0, 9.8], rng.normal(size=3) * 0.2, 0.005
    xn, Pn = orc.predict(x, P, dt, Q, acc, gyr)
    R = orc.quat_to_mat(x[3:7])
    am = acc - x[20:23]
    hat = np.array([[0, -am[2], am[1]], [am[2], 0, -am[0]], [-am[1], am[0], 0]])
    F = np.zeros((24, 24))
    F[0:3, 12:15] = np.eye(3)
    F[12:15, 3:6] = -R @ hat
    F[12:15, 18:21] = -R
    F[12:15, 21:24] = np.eye(3)
    F[3:6, 15:18] = -np.eye(3)
    W = np.zeros((24, 12))
    W[12:15, 3:6] = -R
    W[3:6, 0:3] = -np.eye(3)
    W[15:18, 6:9] = np.eye(3)
    W[18:21, 9:12] = np.eye(3)
    Fd = np.eye(24) + F * dt
    Pref = Fd @ P @ Fd.T + (dt * W) @ Q @ (dt * W).T
    assert np.abs(Pn - Pref).max() <= 1e-13 * np.abs(Pref).max()
    assert np.allclose(xn[0:3], x[0:3] + x[14:17] * dt)
    assert np.allclose(xn[14:17], x[14:17] + (R @ am + x[23:26]) * dt)
    Rn = orc.quat_to_mat(xn[3:7])
    assert np.abs(Rn - R @ Rotation.from_rotvec((gyr - x[17:20]) * dt).as_matrix()).max() < 1e-14
