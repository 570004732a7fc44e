"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/build/liblio_oracle.so) and of the
reference's own ikd-Tree compiled in place (oracle/_ref/libikd_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (agi_lidar_slam_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
ORACLE_SO = _DIR / "build" / "liblio_oracle.so"
IKD_SO = _DIR / "_ref" / "libikd_ref.so"

STATE_DOUBLES = 26
POSE_DOUBLES = 22
CARRY_DOUBLES = 27

_orc = None
_ikd = None


def build(force: bool = False) -> None:
    """Compile the oracle (and, where /root/reference exists, oracle/_ref) via oracle/Makefile."""
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (_DIR / "lio_oracle.cpp").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "build/liblio_oracle.so"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(_DIR), "ref"], check=True, capture_output=True)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib() -> C.CDLL:
    global _orc
    if _orc is None:
        if not ORACLE_SO.exists():
            build()
        L = C.CDLL(str(ORACLE_SO))
        vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
        S = {
            "orc_abi_version": (C.c_int, []),
            "orc_so3_exp": (None, [vp, vp]), "orc_so3_log": (None, [vp, vp]),
            "orc_quat_to_mat": (None, [vp, vp]), "orc_mat_to_quat": (None, [vp, vp]),
            "orc_quat_rotate": (None, [vp, vp, vp]),
            "orc_boxplus": (None, [vp, vp, vp]), "orc_boxminus": (None, [vp, vp, vp]),
            "orc_inverse": (C.c_int, [vp, vp, C.c_int]),
            "orc_qr_solve_5x3": (C.c_int, [vp, vp]), "orc_esti_plane": (C.c_int, [vp, f32, vp]),
            "orc_esti_plane_batch": (None, [vp, i64, f32, vp, vp]),
            "orc_predict": (None, [vp, vp, f64, vp, vp, vp]),
            "orc_imu_forward": (C.c_int, [vp, C.c_int, f64, f64, vp, vp, vp, vp, C.c_int]),
            "orc_imu_init": (None, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
            "orc_undistort": (None, [vp, i64, vp, C.c_int, vp, vp]),
            "orc_voxel_grid": (i64, [vp, i64, f32, vp, vp, vp]),
            "orc_map_create": (vp, [f32]), "orc_map_destroy": (None, [vp]),
            "orc_map_build": (None, [vp, vp, i64]), "orc_map_add": (C.c_int, [vp, vp, i64, C.c_int, f32]),
            "orc_map_delete_boxes": (C.c_int, [vp, vp, C.c_int]), "orc_map_size": (i64, [vp]),
            "orc_map_dump": (i64, [vp, vp, vp, i64]),
            "orc_map_knn": (None, [vp, vp, i64, C.c_int, f32, vp, vp, vp, C.c_int]),
            "orc_map_knn_callback": (vp, []),
            "orc_scan_create": (vp, [vp, i64]), "orc_scan_destroy": (None, [vp]),
            "orc_h_share_model": (i64, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int]),
            "orc_scan_get": (None, [vp, vp, vp, vp, vp, vp, vp]),
            "orc_scan_get_rows": (None, [vp, vp, vp, vp]),
            "orc_update": (C.c_int, [vp, vp, vp, f64, C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int, C.POINTER(i32)]),
            "orc_sizeof_pass_trace": (C.c_int, []),
            "orc_body_to_world": (None, [vp, vp, i64, vp]),
            "orc_map_incremental_classify": (None, [vp, i64, vp, vp, C.c_int, f32, vp]),
        }  # fmt: skip
        for k, (r, a) in S.items():
            fn = getattr(L, k)
            fn.restype, fn.argtypes = r, a
        _orc = L
    return _orc


def ikd_available() -> bool:
    return IKD_SO.exists()


def ikd() -> C.CDLL:
    """The reference ikd-Tree (prebuilt in this container from /root/reference; travels to the GPU box)."""
    global _ikd
    if _ikd is None:
        if not IKD_SO.exists():
            build()
        if not IKD_SO.exists():
            raise FileNotFoundError(f"{IKD_SO} missing and /root/reference not available to build it")
        L = C.CDLL(str(IKD_SO))
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        S = {
            "ikd_create": (vp, [f32, f32, f32]), "ikd_destroy": (None, [vp]), "ikd_set_downsample": (None, [vp, f32]),
            "ikd_size": (C.c_int, [vp]), "ikd_validnum": (C.c_int, [vp]), "ikd_empty": (C.c_int, [vp]),
            "ikd_build": (None, [vp, vp, vp, i64]),
            "ikd_knn": (None, [vp, vp, i64, C.c_int, f32, vp, vp, vp, C.c_int]),
            "ikd_add_points": (C.c_int, [vp, vp, vp, i64, C.c_int]),
            "ikd_delete_boxes": (C.c_int, [vp, vp, C.c_int]),
            "ikd_flatten": (i64, [vp, vp, vp, i64]),
            "ikd_knn_callback": (vp, []),
        }  # fmt: skip
        for k, (r, a) in S.items():
            fn = getattr(L, k)
            fn.restype, fn.argtypes = r, a
        _ikd = L
    return _ikd


# --------------------------------------------------------------------------- math
def so3_exp(w):
    w = np.ascontiguousarray(w, np.float64)
    q = np.zeros(4)
    lib().orc_so3_exp(_p(w), _p(q))
    return q


def so3_log(q):
    q = np.ascontiguousarray(q, np.float64)
    w = np.zeros(3)
    lib().orc_so3_log(_p(q), _p(w))
    return w


def quat_to_mat(q):
    q = np.ascontiguousarray(q, np.float64)
    m = np.zeros(9)
    lib().orc_quat_to_mat(_p(q), _p(m))
    return m.reshape(3, 3)


def mat_to_quat(m):
    m = np.ascontiguousarray(m, np.float64).reshape(9)
    q = np.zeros(4)
    lib().orc_mat_to_quat(_p(m), _p(q))
    return q


def quat_rotate(q, v):
    q = np.ascontiguousarray(q, np.float64)
    v = np.ascontiguousarray(v, np.float64)
    o = np.zeros(3)
    lib().orc_quat_rotate(_p(q), _p(v), _p(o))
    return o


def boxplus(x, f):
    x = np.ascontiguousarray(x, np.float64)
    f = np.ascontiguousarray(f, np.float64)
    o = np.zeros(STATE_DOUBLES)
    lib().orc_boxplus(_p(x), _p(f), _p(o))
    return o


def boxminus(x1, x2):
    x1 = np.ascontiguousarray(x1, np.float64)
    x2 = np.ascontiguousarray(x2, np.float64)
    o = np.zeros(24)
    lib().orc_boxminus(_p(x1), _p(x2), _p(o))
    return o


def inverse(A):
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    o = np.zeros_like(A)
    lib().orc_inverse(_p(A), _p(o), n)
    return o


def qr_solve_5x3(pts):
    pts = np.ascontiguousarray(pts, np.float32).reshape(15)
    x = np.zeros(3, np.float32)
    full = lib().orc_qr_solve_5x3(_p(pts), _p(x))
    return x, bool(full)


def esti_plane(pts, thr=0.1):
    pts = np.ascontiguousarray(pts, np.float32).reshape(15)
    o = np.zeros(4, np.float32)
    ok = lib().orc_esti_plane(_p(pts), thr, _p(o))
    return o, bool(ok)


def esti_plane_batch(pts, thr=0.1):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 15)
    m = pts.shape[0]
    o = np.zeros((m, 4), np.float32)
    ok = np.zeros(m, np.uint8)
    lib().orc_esti_plane_batch(_p(pts), m, thr, _p(o), _p(ok))
    return o, ok


def predict(x, P, dt, Q, acc, gyro):
    x = np.ascontiguousarray(x, np.float64).copy()
    P = np.ascontiguousarray(P, np.float64).reshape(24, 24).copy()
    Q = np.ascontiguousarray(Q, np.float64).reshape(144)
    acc = np.ascontiguousarray(acc, np.float64)
    gyro = np.ascontiguousarray(gyro, np.float64)
    lib().orc_predict(_p(x), _p(P), float(dt), _p(Q), _p(acc), _p(gyro))
    return x, P


def default_state():
    """state_ikfom defaults (use-ikfom.hpp:18-27)."""
    x = np.zeros(STATE_DOUBLES)
    x[3] = 1.0
    x[7] = 1.0
    x[25] = -9.81
    return x


def imu_init_P():
    """init_P of IMU_init (IMU_Processing.hpp:233-238)."""
    P = np.eye(24)
    for i in range(6, 12):
        P[i, i] = 0.00001
    for i in range(15, 18):
        P[i, i] = 0.0001
    for i in range(18, 21):
        P[i, i] = 0.001
    for i in range(21, 24):
        P[i, i] = 0.00001
    return P


def new_carry(cov_gyr=0.1, cov_acc=0.1, cov_bg=1e-4, cov_ba=1e-4, mean_acc_norm=9.81):
    """ImuProcess members read by the forward pass (27 doubles, see orc::ImuCarry)."""
    c = np.zeros(CARRY_DOUBLES)
    c[0:3] = cov_gyr
    c[3:6] = cov_acc
    c[6:9] = cov_bg
    c[9:12] = cov_ba
    c[12] = mean_acc_norm
    return c


def imu_forward(imu, pcl_beg, pcl_end, x, P, carry, cap=256):
    """UndistortPcl forward half.  imu: (n,7) [t, acc3, gyr3].  Returns (poses (P,22), x, P); carry is updated."""
    imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
    x = np.ascontiguousarray(x, np.float64).copy()
    P = np.ascontiguousarray(P, np.float64).reshape(24, 24).copy()
    poses = np.zeros((cap, POSE_DOUBLES))
    n = lib().orc_imu_forward(_p(imu), imu.shape[0], float(pcl_beg), float(pcl_end), _p(x), _p(P), _p(carry),
                              _p(poses), cap)
    return poses[:n].copy(), x, P


def imu_init(imu, first_frame, stats, x, tli, rli_mat):
    imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
    x = np.ascontiguousarray(x, np.float64).copy()
    P = np.zeros((24, 24))
    tli = np.ascontiguousarray(tli, np.float64)
    rli = np.ascontiguousarray(rli_mat, np.float64).reshape(9)
    lib().orc_imu_init(_p(imu), imu.shape[0], int(first_frame), _p(stats), _p(x), _p(P), _p(tli), _p(rli))
    return x, P


def undistort(pts4, poses, end_state):
    """Sort by time (stable) + UndistortPcl back half.  Returns (sorted+compensated (n,4), order (n,))."""
    pts = np.ascontiguousarray(pts4, np.float32).copy()
    n = pts.shape[0]
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, POSE_DOUBLES)
    end = np.ascontiguousarray(end_state, np.float64)
    order = np.zeros(n, np.int64)
    lib().orc_undistort(_p(pts), n, _p(poses), poses.shape[0], _p(end), _p(order))
    return pts, order


def voxel_grid(pts5, leaf=0.5):
    """pcl::VoxelGrid restated.  pts5: (n,5) [x,y,z,intensity,curvature] -> (centroids (m,5), keys (m,3), point_keys (n,3))."""
    pts = np.ascontiguousarray(pts5, np.float32).reshape(-1, 5)
    n = pts.shape[0]
    out = np.zeros((max(n, 1), 5), np.float32)
    keys = np.zeros((max(n, 1), 3), np.int32)
    pk = np.zeros((max(n, 1), 3), np.int32)
    m = lib().orc_voxel_grid(_p(pts), n, leaf, _p(out), _p(keys), _p(pk))
    if m < 0:
        return None, None, pk[:n]
    return out[:m].copy(), keys[:m].copy(), pk[:n].copy()


# --------------------------------------------------------------------------- maps
class Map:
    """Restated map semantics (hashed grid, canonical (d2,id) order)."""

    def __init__(self, cell=1.0):
        self.h = lib().orc_map_create(cell)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def build(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        lib().orc_map_build(self.h, _p(xyz), xyz.shape[0])

    def add(self, xyz, downsample_on, ds=0.5):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        return lib().orc_map_add(self.h, _p(xyz), xyz.shape[0], int(downsample_on), ds)

    def delete_boxes(self, boxes6):
        b = np.ascontiguousarray(boxes6, np.float32).reshape(-1, 6)
        return lib().orc_map_delete_boxes(self.h, _p(b), b.shape[0])

    def size(self):
        return lib().orc_map_size(self.h)

    def dump(self):
        n = lib().orc_map_dump(self.h, None, None, 0)
        xyz = np.zeros((n, 3), np.float32)
        ids = np.zeros(n, np.int32)
        lib().orc_map_dump(self.h, _p(xyz), _p(ids), n)
        return xyz, ids

    def knn(self, q, k=5, max_d2=5.0, threads=1):
        """max_d2 = np.inf: the reference's unbounded Nearest_Search (ikd_Tree.h:285)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        m = q.shape[0]
        idx = np.zeros((m, k), np.int32)
        d2 = np.zeros((m, k), np.float32)
        nbr = np.zeros((m, k, 3), np.float32)
        lib().orc_map_knn(self.h, _p(q), m, k, max_d2, _p(idx), _p(d2), _p(nbr), threads)
        return idx, d2, nbr

    def knn_backend(self):
        return lib().orc_map_knn_callback(), self.h


class IkdTree:
    """The REFERENCE ikd-Tree (KD_TREE<pcl::PointXYZINormal>), unmodified, behind oracle/ikd_wrap.cpp."""

    def __init__(self, delete_param=0.5, balance_param=0.6, box_length=0.2):
        self.h = ikd().ikd_create(delete_param, balance_param, box_length)

    def __del__(self):
        if getattr(self, "h", None):
            ikd().ikd_destroy(self.h)
            self.h = None

    def set_downsample_param(self, ds):
        ikd().ikd_set_downsample(self.h, ds)

    def build(self, xyz, ids=None):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        if ids is not None:
            ids = np.ascontiguousarray(ids, np.int32)
        ikd().ikd_build(self.h, _p(xyz), _p(ids), xyz.shape[0])

    def knn(self, q, k=5, max_dist=np.inf, threads=1):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        m = q.shape[0]
        idx = np.zeros((m, k), np.int32)
        d2 = np.zeros((m, k), np.float32)
        nbr = np.zeros((m, k, 3), np.float32)
        ikd().ikd_knn(self.h, _p(q), m, k, max_dist, _p(idx), _p(d2), _p(nbr), threads)
        return idx, d2, nbr

    def add_points(self, xyz, downsample_on, ids=None):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        if ids is not None:
            ids = np.ascontiguousarray(ids, np.int32)
        return ikd().ikd_add_points(self.h, _p(xyz), _p(ids), xyz.shape[0], int(downsample_on))

    def delete_boxes(self, boxes6):
        b = np.ascontiguousarray(boxes6, np.float32).reshape(-1, 6)
        return ikd().ikd_delete_boxes(self.h, _p(b), b.shape[0])

    def flatten(self):
        n = ikd().ikd_flatten(self.h, None, None, 0)
        xyz = np.zeros((n, 3), np.float32)
        ids = np.zeros(n, np.int32)
        ikd().ikd_flatten(self.h, _p(xyz), _p(ids), n)
        return xyz, ids

    def size(self):
        return ikd().ikd_size(self.h)

    def validnum(self):
        return ikd().ikd_validnum(self.h)

    def knn_backend(self):
        return ikd().ikd_knn_callback(), self.h


class IkdLiveMap:
    """The REFERENCE ikd-Tree as the live map of a whole replay (Build at laserMapping.cpp:747-758, unbounded
    Nearest_Search at esekfom.hpp:140, both Add_Points calls of laserMapping.cpp:430-431, Delete_Point_Boxes at :361-364)
    behind the interface of `Map`.  Points carry an id (in normal_x, which the tree copies through verbatim) handed out
    as the product and the Map port hand them out: input index of Build, then next_id + batch index."""

    def __init__(self, downsample=0.5):
        self.tree = IkdTree()
        self.tree.set_downsample_param(downsample)  # laserMapping.cpp:748
        self.next_id = 0

    def build(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self.tree.build(xyz, np.arange(len(xyz), dtype=np.int32))
        self.next_id = len(xyz)

    def add(self, xyz, downsample_on, ds=0.5):
        """Returns the number of NEW points that are in the map afterwards (the tree's own return value counts the
        insert operations of its sequential loop instead; the reference only prints it)."""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        n = len(xyz)
        if n == 0:
            return 0
        ids = np.arange(self.next_id, self.next_id + n, dtype=np.int32)
        self.tree.add_points(xyz, downsample_on, ids)
        base = self.next_id
        self.next_id += n
        if not downsample_on:
            return n
        return int((self.tree.flatten()[1] >= base).sum())

    def delete_boxes(self, boxes6):
        return self.tree.delete_boxes(boxes6)

    def size(self):
        return self.tree.validnum()

    def dump(self):
        xyz, ids = self.tree.flatten()
        o = np.argsort(ids, kind="stable")
        return xyz[o], ids[o]

    def knn(self, q, k=5, max_d2=np.inf, threads=1):
        idx, d2, nbr = self.tree.knn(q, k, np.inf, threads)
        return idx, d2, nbr

    def knn_backend(self):
        return self.tree.knn_backend()


# --------------------------------------------------------------------------- h_share_model / update
def pass_trace_dtype():
    dt = np.dtype([("searched", np.int32), ("valid", np.int32), ("n_valid", np.int32), ("converged", np.int32),
                   ("blob", np.float64, 90), ("dx", np.float64, 24)])
    assert dt.itemsize == lib().orc_sizeof_pass_trace(), (dt.itemsize, lib().orc_sizeof_pass_trace())
    return dt


class Scan:
    """Per-scan persistent state of h_share_model (Nearest_Points, point_selected_surf, normvec)."""

    def __init__(self, body_xyz):
        b = np.ascontiguousarray(body_xyz, np.float32).reshape(-1, 3)
        self.m = b.shape[0]
        self.h = lib().orc_scan_create(_p(b), self.m)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scan_destroy(self.h)
            self.h = None

    def h_share_model(self, x, converge, extrinsic_est, backend, threads=1):
        x = np.ascontiguousarray(x, np.float64)
        cb, ctx = backend
        return lib().orc_h_share_model(self.h, _p(x), int(converge), int(extrinsic_est), cb, ctx, threads)

    def get(self):
        m = self.m
        world = np.zeros((m, 3), np.float32)
        near = np.zeros((m, 5, 4), np.float32)
        d2 = np.zeros((m, 5), np.float32)
        cnt = np.zeros(m, np.int32)
        sel = np.zeros(m, np.uint8)
        nv = np.zeros((m, 4), np.float32)
        lib().orc_scan_get(self.h, _p(world), _p(near), _p(d2), _p(cnt), _p(sel), _p(nv))
        idx = near[:, :, 3].copy().view(np.int32)
        return dict(world=world, nbr=near[:, :, :3].copy(), idx=idx, d2=d2, cnt=cnt, selected=sel, normvec=nv,
                    near_raw=near)

    def rows(self, v):
        hx = np.zeros((v, 12))
        h = np.zeros(v)
        vi = np.zeros(v, np.int32)
        lib().orc_scan_get_rows(self.h, _p(hx), _p(h), _p(vi))
        return hx, h, vi

    def update(self, x, P, backend, R=0.001, max_iter=4, extrinsic_est=False, threads=1):
        x = np.ascontiguousarray(x, np.float64).copy()
        P = np.ascontiguousarray(P, np.float64).reshape(24, 24).copy()
        trace = np.zeros(max_iter + 2, pass_trace_dtype())
        nv = C.c_int32(0)
        cb, ctx = backend
        n = lib().orc_update(self.h, _p(x), _p(P), R, max_iter, int(extrinsic_est), cb, ctx, threads, _p(trace),
                             trace.shape[0], C.byref(nv))
        return x, P, trace[:n].copy(), nv.value


def body_to_world(x, body):
    x = np.ascontiguousarray(x, np.float64)
    b = np.ascontiguousarray(body, np.float32).reshape(-1, 3)
    w = np.zeros_like(b)
    lib().orc_body_to_world(_p(x), _p(b), b.shape[0], _p(w))
    return w


def map_incremental_classify(world, near_raw, near_cnt, ekf_inited=True, filter_size_map=0.5):
    w = np.ascontiguousarray(world, np.float32).reshape(-1, 3)
    nr = np.ascontiguousarray(near_raw, np.float32).reshape(-1, 20)
    nc = np.ascontiguousarray(near_cnt, np.int32)
    cls = np.zeros(w.shape[0], np.uint8)
    lib().orc_map_incremental_classify(_p(w), w.shape[0], _p(nr), _p(nc), int(ekf_inited), filter_size_map, _p(cls))
    return cls


# --------------------------------------------------------------------------- sensor decoding (src/preprocess.cpp)
def decode_cloud2(data, point_step, off_x, off_y, off_z, off_intensity, off_time, time_dtype, time_scale, blind,
                  point_filter_num, rule):
    """oust64_handler (preprocess.cpp:243-268, rule 1) / velodyne_handler with point times (:380-428, rule 2), feature
    extraction off: every point_filter_num-th record, blind-zone test on the FP32 squared range compared in FP64 with
    the double `blind`, curvature = time field * time_unit_scale in FP32.  data: (n, point_step) uint8.
    Returns (xyzt (k,4) float32, intensity (k,) float32)."""
    data = np.ascontiguousarray(data, np.uint8).reshape(-1, point_step)
    n = data.shape[0]

    def field(off, dt):
        w = np.dtype(dt).itemsize
        return np.ascontiguousarray(data[:, off:off + w]).view(dt)[:, 0]

    x, y, z = field(off_x, "<f4"), field(off_y, "<f4"), field(off_z, "<f4")
    inten = field(off_intensity, "<f4") if off_intensity >= 0 else np.zeros(n, np.float32)
    if off_time >= 0:
        t = field(off_time, time_dtype)
        if np.dtype(time_dtype) == np.dtype("<f8"):
            tms = (t * np.float64(np.float32(time_scale))).astype(np.float32)
        else:
            tms = (t.astype(np.float32) * np.float32(time_scale)).astype(np.float32)
    else:
        tms = np.zeros(n, np.float32)
    rng = ((x * x + y * y) + z * z).astype(np.float32).astype(np.float64)
    b2 = np.float64(blind) * np.float64(blind)
    keep = (np.arange(n) % point_filter_num == 0) & (~(rng < b2) if rule == 1 else (rng > b2))
    return np.stack([x, y, z, tms], 1)[keep].astype(np.float32), inten[keep].astype(np.float32)


def _fields(data, point_step):
    data = np.ascontiguousarray(data, np.uint8).reshape(-1, point_step)

    def field(off, dt):
        w = np.dtype(dt).itemsize
        return np.ascontiguousarray(data[:, off:off + w]).view(dt)[:, 0]

    return data.shape[0], field


def decode_avia(data, point_step, off_x, off_y, off_z, off_refl, off_time, off_tag, off_line, n_scans, blind,
                point_filter_num):
    """avia_handler, feature extraction off (preprocess.cpp:160-183), over livox CustomPoint records, as written: the
    loop starts at record 1; a record is valid when line < N_SCANS and (tag & 0x30) is 0x10 or 0x00; every
    point_filter_num-th VALID record is copied into pl_full[i] with curvature = offset_time / float(1000000), and pushed
    when it differs from pl_full[i - 1] (zero-initialised unless record i - 1 was copied too) by more than 1e-7 in x, or
    in y, or (in z and the squared range exceeds blind^2) -- `a || b || c && d`.
    Returns (xyzt (k,4) float32, intensity (k,) float32)."""
    n, field = _fields(data, point_step)
    x, y, z = field(off_x, "<f4"), field(off_y, "<f4"), field(off_z, "<f4")
    refl, tag, line, t = field(off_refl, "u1"), field(off_tag, "u1"), field(off_line, "u1"), field(off_time, "<u4")
    pl_full = np.zeros((max(n, 1), 3), np.float32)
    b2 = np.float64(blind) * np.float64(blind)
    valid_num = 0
    out, inten = [], []
    for i in range(1, n):
        if line[i] < n_scans and ((tag[i] & 0x30) == 0x10 or (tag[i] & 0x30) == 0x00):
            valid_num += 1
            if valid_num % point_filter_num == 0:
                pl_full[i] = (x[i], y[i], z[i])
                cur = np.float32(t[i]) / np.float32(1000000)
                dx = np.abs(np.float32(x[i] - pl_full[i - 1, 0]))
                dy = np.abs(np.float32(y[i] - pl_full[i - 1, 1]))
                dz = np.abs(np.float32(z[i] - pl_full[i - 1, 2]))
                rng = np.float64(np.float32(np.float32(x[i] * x[i] + y[i] * y[i]) + z[i] * z[i]))
                if np.float64(dx) > 1e-7 or np.float64(dy) > 1e-7 or (np.float64(dz) > 1e-7 and rng > b2):
                    out.append((x[i], y[i], z[i], cur))
                    inten.append(np.float32(refl[i]))
    return np.array(out, np.float32).reshape(-1, 4), np.array(inten, np.float32)


def decode_yaw_times(data, point_step, off_x, off_y, off_z, off_intensity, intensity_dtype, off_ring, ring_dtype, n_scans,
                     scan_rate, blind, point_filter_num):
    """velodyne_handler (preprocess.cpp:380-428) / rs_handler (:872-921) when the driver gives no point times
    (given_offset_time == false), feature extraction off.  Per ring: the first record fixes yaw_fp and is skipped
    (`continue`), every later record gets curvature = (yaw_fp - yaw) / omega_l, or (yaw_fp - yaw + 360) / omega_l when
    yaw > yaw_fp, plus 360 / omega_l when that is smaller than the ring's previous curvature.  yaw = atan2(y, x) in FP32
    (std::atan2(float, float): preprocess.h:6 has `using namespace std`) times the double 57.2957; yaw_fp is double,
    time_last and curvature are float.  Then every point_filter_num-th RECORD (by index) outside the blind zone is kept."""
    n, field = _fields(data, point_step)
    x, y, z = field(off_x, "<f4"), field(off_y, "<f4"), field(off_z, "<f4")
    inten = field(off_intensity, intensity_dtype).astype(np.float32) if off_intensity >= 0 else np.zeros(n, np.float32)
    ring = field(off_ring, ring_dtype)
    omega_l = 0.361 * scan_rate
    yaw = np.arctan2(y, x).astype(np.float32).astype(np.float64) * 57.2957
    rng = ((x * x + y * y) + z * z).astype(np.float32).astype(np.float64)
    b2 = np.float64(blind) * np.float64(blind)
    is_first = [True] * n_scans
    yaw_fp = [0.0] * n_scans
    time_last = [np.float32(0)] * n_scans
    out, oi = [], []
    for i in range(n):
        layer = int(ring[i])
        ya = yaw[i]
        if is_first[layer]:
            yaw_fp[layer] = ya
            is_first[layer] = False
            time_last[layer] = np.float32(0)
            continue
        if ya <= yaw_fp[layer]:
            cur = np.float32((yaw_fp[layer] - ya) / omega_l)
        else:
            cur = np.float32((yaw_fp[layer] - ya + 360.0) / omega_l)
        if cur < time_last[layer]:
            cur = np.float32(np.float64(cur) + 360.0 / omega_l)
        time_last[layer] = cur
        if i % point_filter_num == 0 and rng[i] > b2:
            out.append((x[i], y[i], z[i], cur))
            oi.append(inten[i])
    return np.array(out, np.float32).reshape(-1, 4), np.array(oi, np.float32)


def decode_rs(data, point_step, off_x, off_y, off_z, off_intensity, off_time, blind, point_filter_num):
    """rs_handler with point timestamps (preprocess.cpp:872-921): curvature = (timestamp - points[0].timestamp) * 1000.0
    in FP64, stored as FP32; uint8 intensity; every point_filter_num-th record with squared range > blind^2."""
    n, field = _fields(data, point_step)
    x, y, z = field(off_x, "<f4"), field(off_y, "<f4"), field(off_z, "<f4")
    inten = field(off_intensity, "u1").astype(np.float32)
    ts = field(off_time, "<f8")
    tms = ((ts - ts[0]) * 1000.0).astype(np.float32) if n else np.zeros(0, np.float32)
    rng = ((x * x + y * y) + z * z).astype(np.float32).astype(np.float64)
    keep = (np.arange(n) % point_filter_num == 0) & (rng > np.float64(blind) * np.float64(blind))
    return np.stack([x, y, z, tms], 1)[keep].astype(np.float32), inten[keep]
