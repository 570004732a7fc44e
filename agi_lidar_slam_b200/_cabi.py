"""ctypes binding of liblio_b200.so (include/lio_b200.h).

This is the ONLY way Python reaches the hot path: there is no NumPy/torch fallback.  Importing this module
without the built library raises; creating a context without a B200-class GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LIO_LIB", str(_PKG / "liblio_b200.so")))  # LIO_LIB: kernel-variant experiments

LIO_OK = 0
LIO_E_INVALID, LIO_E_NO_DEVICE, LIO_E_CUDA, LIO_E_CAPACITY, LIO_E_EMPTY_MAP, LIO_E_VOXEL_RANGE = -1, -2, -3, -4, -5, -6
_ERR_NAMES = {
    -1: "LIO_E_INVALID",
    -2: "LIO_E_NO_DEVICE",
    -3: "LIO_E_CUDA",
    -4: "LIO_E_CAPACITY",
    -5: "LIO_E_EMPTY_MAP",
    -6: "LIO_E_VOXEL_RANGE",
}

STATE_DOUBLES = 26  # lio_state
POSE_DOUBLES = 22  # lio_pose6d
BLOB = 92


class LioError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class ScanReport(C.Structure):
    """lio_scan_report (include/lio_b200.h)."""
    _fields_ = [("m", C.c_int64), ("status", C.c_int32), ("n_valid", C.c_int32), ("n_passes", C.c_int32),
                ("counts", C.c_int32 * 3)]


SCAN_UPDATED, SCAN_FEW_POINTS, SCAN_MAP_BUILT = 0, 1, 2


class SeqConfig(C.Structure):
    """lio_seq_config (include/lio_b200.h)."""
    _fields_ = [("filter_size_surf", C.c_float), ("filter_size_map", C.c_float), ("max_iteration", C.c_int32),
                ("extrinsic_est", C.c_int32), ("extrinsic_T", C.c_double * 3), ("extrinsic_R", C.c_double * 9),
                ("gyr_cov", C.c_double), ("acc_cov", C.c_double), ("b_gyr_cov", C.c_double), ("b_acc_cov", C.c_double),
                ("cube_len", C.c_double), ("det_range", C.c_double), ("laser_point_cov", C.c_double)]


class SeqInput(C.Structure):
    """lio_seq_input: one MeasureGroup."""
    _fields_ = [("lidar", C.c_void_p), ("n", C.c_int64), ("stride_bytes", C.c_int32), ("n_imu", C.c_int32),
                ("imu", C.c_void_p), ("lidar_beg_time", C.c_double), ("lidar_end_time", C.c_double)]


class SeqResult(C.Structure):
    """lio_seq_result."""
    _fields_ = [("status", C.c_int32), ("n_valid", C.c_int32), ("n_passes", C.c_int32), ("counts", C.c_int32 * 3),
                ("m", C.c_int64), ("n_box_deleted", C.c_int64), ("x", C.c_double * 26)]


SEQ_UPDATED, SEQ_FEW_POINTS, SEQ_MAP_BUILT, SEQ_FIRST_SCAN, SEQ_NO_IMU, SEQ_IMU_INIT = range(6)
SEQ_STATUS_NAMES = {SEQ_UPDATED: "ok", SEQ_FEW_POINTS: "few-points", SEQ_MAP_BUILT: "map-built", SEQ_FIRST_SCAN: "first",
                    SEQ_NO_IMU: "no-imu", SEQ_IMU_INIT: "imu-init"}


class Caps(C.Structure):
    _fields_ = [
        ("max_scan_points", C.c_int64),
        ("max_down_points", C.c_int64),
        ("max_map_points", C.c_int64),
        ("map_cell", C.c_float),
        ("knn_max_d2", C.c_float),
        ("plane_thr", C.c_float),
        ("map_downsample", C.c_float),
    ]


class CloudLayout(C.Structure):
    """lio_cloud_layout: where x/y/z/intensity/time sit in a PointCloud2 record + the handler's decimation / blind rule."""

    _fields_ = [("point_step", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_time", C.c_int32), ("time_type", C.c_int32),
                ("point_filter_num", C.c_int32), ("rule", C.c_int32), ("time_scale", C.c_float), ("blind", C.c_double),
                ("off_ring", C.c_int32), ("ring_type", C.c_int32), ("off_tag", C.c_int32), ("intensity_type", C.c_int32),
                ("n_scans", C.c_int32), ("scan_rate", C.c_int32), ("yaw_time", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/lio_b200.h declares (tests/test_abi.py checks the library exports all of them)
EXPORTS = [
    "lio_abi_version", "lio_default_caps", "lio_create", "lio_destroy", "lio_set_stream", "lio_synchronize",
    "lio_last_error", "lio_launch_count", "lio_map_build", "lio_map_add", "lio_map_delete_boxes", "lio_map_size",
    "lio_map_dump", "lio_map_set_downsample", "lio_knn5", "lio_knn5_resident", "lio_scan_preprocess", "lio_scan_preprocess_resident", "lio_scan_preprocess_cloud2", "lio_scan_decoded",
    "lio_scan_upload",
    "lio_update_pass", "lio_update_scan", "lio_update_scan_host", "lio_state_upload", "lio_state_download", "lio_update_enqueue",
    "lio_update_enqueue_multi", "lio_scan_step", "lio_scan_step_begin", "lio_scan_step_end",
    "lio_scan_step_finish", "lio_scan_step_prefetch", "lio_set_deferred_growth", "lio_scan_step_settle", "lio_update_begin", "lio_update_pass_enqueue", "lio_update_step_enqueue", "lio_blob_device_ptr",
    "lio_blob_download", "lio_blob_upload", "lio_blob_bind", "lio_peer_handle", "lio_peer_connect",
    "lio_update_enqueue_sharded", "lio_peer_status", "lio_set_shard_stripes", "lio_map_removed_points", "lio_pass_only_enqueue", "lio_debug_timeline", "lio_debug_blocks",
    "lio_get_neighbors", "lio_map_incremental", "lio_map_build_scan", "lio_predict", "lio_boxplus", "lio_boxminus",
    "lio_imu_proc_init", "lio_imu_set_param", "lio_imu_process",
    "lio_seq_default_config", "lio_seq_create", "lio_seq_destroy", "lio_seq_process", "lio_seq_process_many",
    "lio_seq_get_state", "lio_seq_set_state", "lio_seq_local_map", "lio_set_host_threads",
]  # fmt: skip

_lib = None


def load_library() -> C.CDLL:
    """Load liblio_b200.so; fail loudly when it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C agi_lidar_slam_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
    P = C.POINTER
    sig = {
        "lio_abi_version": (C.c_int, []),
        "lio_default_caps": (None, [P(Caps)]),
        "lio_create": (C.c_int, [C.c_int, P(Caps), P(vp)]),
        "lio_destroy": (None, [vp]),
        "lio_set_stream": (C.c_int, [vp, vp]),
        "lio_synchronize": (C.c_int, [vp]),
        "lio_last_error": (C.c_char_p, [vp]),
        "lio_launch_count": (i64, [vp]),
        "lio_map_build": (C.c_int, [vp, vp, i64, C.c_int]),
        "lio_map_add": (C.c_int, [vp, vp, i64, C.c_int, C.c_int, P(i32)]),
        "lio_map_delete_boxes": (C.c_int, [vp, vp, C.c_int, P(i32)]),
        "lio_map_size": (C.c_int, [vp, P(i64), P(i64)]),
        "lio_map_dump": (C.c_int, [vp, vp, vp, i64, P(i64)]),
        "lio_knn5": (C.c_int, [vp, vp, i64, C.c_float, vp, vp, vp]),
        "lio_map_set_downsample": (C.c_int, [vp, C.c_float]),
        "lio_knn5_resident": (C.c_int, [vp, i64]),
        "lio_scan_preprocess": (C.c_int, [vp, vp, i64, C.c_int, vp, C.c_int, vp, f32, vp, P(i64), vp, vp]),
        "lio_scan_preprocess_resident": (C.c_int, [vp, vp, i64, C.c_int, vp, C.c_int, vp, f32, P(i64)]),
        "lio_scan_upload": (C.c_int, [vp, vp, i64, C.c_int]),
        "lio_scan_preprocess_cloud2": (C.c_int, [vp, vp, i64, P(CloudLayout), vp, C.c_int, vp, f32, P(i64), P(i64)]),
        "lio_scan_decoded": (C.c_int, [vp, vp, vp, i64, P(i64)]),
        "lio_update_pass": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, P(i32)]),
        "lio_update_scan": (C.c_int, [vp, vp, vp, f64, C.c_int, C.c_int, P(i32), P(i32)]),
        "lio_update_scan_host": (C.c_int, [vp, vp, i64, C.c_int, vp, vp, f64, C.c_int, C.c_int, P(i32), P(i32)]),
        "lio_state_upload": (C.c_int, [vp, vp, vp]),
        "lio_state_download": (C.c_int, [vp, vp, vp, P(i32), P(i32)]),
        "lio_update_enqueue": (C.c_int, [vp, f64, C.c_int, C.c_int, C.c_int]),
        "lio_update_enqueue_multi": (C.c_int, [vp, C.c_int, f64, C.c_int, C.c_int, C.c_int]),
        "lio_scan_step": (C.c_int, [vp, vp, i64, C.c_int, vp, C.c_int, vp, vp, C.c_float, C.c_float, f64, C.c_int,
                                    C.c_int, C.c_int, vp]),
        "lio_scan_step_begin": (C.c_int, [vp, vp, i64, C.c_int, vp, C.c_int, vp, vp, C.c_float, vp]),
        "lio_scan_step_end": (C.c_int, [vp, C.c_float, C.c_int]),
        "lio_scan_step_finish": (C.c_int, [vp, vp, vp, vp]),
        "lio_scan_step_prefetch": (C.c_int, [vp, vp, i64, C.c_int]),
        "lio_set_deferred_growth": (C.c_int, [vp, C.c_int]),
        "lio_scan_step_settle": (C.c_int, [vp, vp]),
        "lio_update_begin": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "lio_update_pass_enqueue": (C.c_int, [vp, C.c_int, f32, f32]),
        "lio_update_step_enqueue": (C.c_int, [vp, f64, C.c_int]),
        "lio_blob_device_ptr": (vp, [vp]),
        "lio_blob_download": (C.c_int, [vp, vp]),
        "lio_blob_upload": (C.c_int, [vp, vp]),
        "lio_blob_bind": (C.c_int, [vp, vp]),
        "lio_peer_handle": (C.c_int, [vp, vp]),
        "lio_peer_connect": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "lio_update_enqueue_sharded": (C.c_int, [vp, f64, C.c_int, C.c_int, C.c_int, f32, f32]),
        "lio_peer_status": (C.c_int, [vp, P(i32)]),
        "lio_pass_only_enqueue": (C.c_int, [vp, C.c_int, C.c_int]),
        "lio_set_shard_stripes": (C.c_int, [vp, f32, f32, C.c_int, C.c_int]),
        "lio_map_removed_points": (C.c_int, [vp, vp, C.c_int64, vp]),
        "lio_debug_timeline": (C.c_int, [vp, vp]),
        "lio_debug_blocks": (C.c_int, [vp, vp]),
        "lio_get_neighbors": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "lio_map_incremental": (C.c_int, [vp, vp, f32, C.c_int, vp]),
        "lio_map_build_scan": (C.c_int, [vp, vp]),
        "lio_predict": (C.c_int, [vp, vp, f64, vp, vp, vp]),
        "lio_boxplus": (C.c_int, [vp, vp, vp]),
        "lio_boxminus": (C.c_int, [vp, vp, vp]),
        "lio_imu_proc_init": (None, [vp]),
        "lio_imu_set_param": (None, [vp, vp, vp, vp, vp, vp, vp]),
        "lio_imu_process": (C.c_int, [vp, vp, C.c_int, f64, f64, vp, vp, vp, C.c_int, P(C.c_int), P(C.c_int)]),
        "lio_seq_default_config": (None, [P(SeqConfig)]),
        "lio_seq_create": (C.c_int, [vp, P(SeqConfig), P(vp)]),
        "lio_seq_destroy": (None, [vp]),
        "lio_seq_process": (C.c_int, [vp, P(SeqInput), P(SeqResult)]),
        "lio_seq_process_many": (C.c_int, [vp, C.c_int, vp, vp]),
        "lio_seq_get_state": (C.c_int, [vp, vp, vp]),
        "lio_seq_set_state": (C.c_int, [vp, vp, vp]),
        "lio_seq_local_map": (C.c_int, [vp, vp, P(i64)]),
        "lio_set_host_threads": (C.c_int, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def default_caps() -> Caps:
    caps = Caps()
    load_library().lio_default_caps(C.byref(caps))
    return caps


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim != 2 or a.shape[1] != shape_last):
        raise ValueError(f"expected an (n, {shape_last}) float32 array, got {a.shape}")
    return a


def _state(x):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    if x.size != STATE_DOUBLES:
        raise ValueError("state must hold 26 doubles (pos3, rot4 wxyz, R_LI4 wxyz, t_LI3, vel3, bg3, ba3, grav3)")
    return x


def _points(pts):
    """(n,4) float32 -> stride 16; (n,12) float32 -> stride 48 (pcl::PointXYZINormal)."""
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] not in (4, 12):
        raise ValueError("points must be (n,4) [x,y,z,w] or (n,12) [PointXYZINormal] float32")
    return pts, pts.shape[1] * 4


class Context:
    """One lio_ctx (one GPU).  Not re-entrant, like the reference's single application thread."""

    def __init__(self, device: int = 0, caps: Caps | None = None, **cap_overrides):
        self._lib = load_library()
        self._h = C.c_void_p()
        caps = caps or default_caps()
        for k, v in cap_overrides.items():
            if not hasattr(caps, k):
                raise TypeError(f"unknown lio_caps field {k}")
            setattr(caps, k, v)
        self.caps = caps
        rc = self._lib.lio_create(device, C.byref(caps), C.byref(self._h))
        if rc != LIO_OK or not self._h:
            self._h = C.c_void_p()
            raise LioError(rc, "lio_create failed (needs a visible sm_100 GPU; there is no CPU fallback)")

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.lio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != LIO_OK:
            raise LioError(rc, (self._lib.lio_last_error(self._h) or b"").decode())

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.lio_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._lib.lio_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.lio_launch_count(self._h))

    # -- map
    def map_build(self, pts):
        pts, stride = _points(pts)
        self._check(self._lib.lio_map_build(self._h, _ptr(pts), pts.shape[0], stride))

    def map_add(self, pts, downsample_on: bool) -> int:
        pts, stride = _points(pts)
        n = C.c_int32(0)
        self._check(self._lib.lio_map_add(self._h, _ptr(pts), pts.shape[0], stride, int(downsample_on), C.byref(n)))
        return n.value

    def map_delete_boxes(self, boxes6) -> int:
        b = _f32(boxes6).reshape(-1, 6)
        n = C.c_int32(0)
        self._check(self._lib.lio_map_delete_boxes(self._h, _ptr(b), b.shape[0], C.byref(n)))
        return n.value

    def map_size(self):
        t, v = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.lio_map_size(self._h, C.byref(t), C.byref(v)))
        return t.value, v.value

    def map_dump(self):
        n = C.c_int64(0)
        self._check(self._lib.lio_map_dump(self._h, None, None, 0, C.byref(n)))
        xyz = np.zeros((n.value, 3), np.float32)
        ids = np.zeros(n.value, np.int32)
        if n.value:
            self._check(self._lib.lio_map_dump(self._h, _ptr(xyz), _ptr(ids), n.value, C.byref(n)))
        return xyz, ids

    def map_removed_points(self) -> np.ndarray:
        """The points deleted since the last call (fetch and clear), (n, 3) float32, order unspecified."""
        n = C.c_int64(0)
        self._check(self._lib.lio_map_removed_points(self._h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.float32)
        self._check(self._lib.lio_map_removed_points(self._h, _ptr(out), n.value, C.byref(n)))
        return out[: n.value].copy()

    def map_set_downsample(self, downsample_size: float):
        self._check(self._lib.lio_map_set_downsample(self._h, float(downsample_size)))

    def knn5(self, q_xyz, want_xyz=True, max_d2=5.0):
        """max_d2 = max_dist ** 2 of KD_TREE::Nearest_Search; np.inf is the reference's default (unbounded)."""
        q = _f32(q_xyz, 3)
        m = q.shape[0]
        idx = np.full((m, 5), -1, np.int32)
        d2 = np.full((m, 5), np.inf, np.float32)
        nbr = np.zeros((m, 5, 3), np.float32) if want_xyz else None
        self._check(self._lib.lio_knn5(self._h, _ptr(q), m, float(max_d2), _ptr(idx), _ptr(d2), _ptr(nbr)))
        return idx, d2, nbr

    def knn5_resident(self, m: int):
        self._check(self._lib.lio_knn5_resident(self._h, m))

    # -- scan
    def scan_preprocess(self, raw_pts, poses=None, end_state=None, leaf=0.5, want_undistorted=False, want_keys=False,
                        resident=False, want_m=True):
        raw, stride = _points(raw_pts)
        n = raw.shape[0]
        if poses is not None:
            poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, POSE_DOUBLES)
            n_poses = poses.shape[0]
            end = _state(end_state)
        else:
            n_poses, end = 0, None
        m = C.c_int64(0)
        if resident:
            self._check(self._lib.lio_scan_preprocess_resident(self._h, _ptr(raw), n, stride, _ptr(poses), n_poses,
                                                               _ptr(end), leaf, C.byref(m) if want_m else None))
            return m.value if want_m else None
        out = np.zeros((int(self.caps.max_down_points), raw.shape[1]), np.float32)
        und = np.zeros((n, 4), np.float32) if want_undistorted else None
        keys = np.zeros((n, 3), np.int32) if want_keys else None
        self._check(self._lib.lio_scan_preprocess(self._h, _ptr(raw), n, stride, _ptr(poses), n_poses, _ptr(end), leaf,
                                                  _ptr(out), C.byref(m), _ptr(und), _ptr(keys)))
        return out[: m.value].copy(), und, keys

    def scan_preprocess_cloud2(self, data: np.ndarray, layout: CloudLayout, poses=None, end_state=None, leaf=0.5):
        """data: (n, point_step) uint8 PointCloud2 records.  Returns (n_decoded, m)."""
        data = np.ascontiguousarray(data, np.uint8)
        if poses is not None:
            poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, POSE_DOUBLES)
            n_poses, end = poses.shape[0], _state(end_state)
        else:
            n_poses, end = 0, None
        nd, m = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.lio_scan_preprocess_cloud2(self._h, _ptr(data), data.shape[0], C.byref(layout), _ptr(poses),
                                                         n_poses, _ptr(end), leaf, C.byref(nd), C.byref(m)))
        return nd.value, m.value

    def scan_decoded(self):
        n = C.c_int64(0)
        self._check(self._lib.lio_scan_decoded(self._h, None, None, 0, C.byref(n)))
        xyzt, inten = np.zeros((n.value, 4), np.float32), np.zeros(n.value, np.float32)
        if n.value:
            self._check(self._lib.lio_scan_decoded(self._h, _ptr(xyzt), _ptr(inten), n.value, C.byref(n)))
        return xyzt, inten

    def scan_upload(self, down_pts):
        pts, stride = _points(down_pts)
        self._check(self._lib.lio_scan_upload(self._h, _ptr(pts), pts.shape[0], stride))

    # -- one main-loop iteration per call
    @staticmethod
    def _step_inputs(raw_pts, poses):
        pts, stride = _points(raw_pts)
        if poses is None or len(poses) < 2:
            return pts, stride, None, 0
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, POSE_DOUBLES)
        return pts, stride, poses, poses.shape[0]

    def scan_step(self, raw_pts, poses, x, P, leaf_surf, leaf_map, R=0.001, max_iter=4, extrinsic_est=False,
                  ekf_inited=True):
        """lio_scan_step.  x (26,) and P (24,24) float64 contiguous are updated IN PLACE; returns the ScanReport."""
        pts, stride, poses, n_poses = self._step_inputs(raw_pts, poses)
        rep = ScanReport()
        self._check(self._lib.lio_scan_step(self._h, _ptr(pts), pts.shape[0], stride,
                                            _ptr(poses) if n_poses else None, n_poses, x.ctypes.data, P.ctypes.data,
                                            leaf_surf, leaf_map, R, max_iter, int(extrinsic_est), int(ekf_inited),
                                            C.byref(rep)))
        return rep

    def scan_step_begin(self, raw_pts, poses, x, P, leaf_surf) -> bool:
        pts, stride, poses, n_poses = self._step_inputs(raw_pts, poses)
        due = C.c_int32(0)
        self._check(self._lib.lio_scan_step_begin(self._h, _ptr(pts), pts.shape[0], stride,
                                                  _ptr(poses) if n_poses else None, n_poses, x.ctypes.data,
                                                  P.ctypes.data, leaf_surf, C.byref(due)))
        return bool(due.value)

    def scan_step_end(self, leaf_map, ekf_inited=True):
        self._check(self._lib.lio_scan_step_end(self._h, leaf_map, int(ekf_inited)))

    def scan_step_finish(self, x, P):
        rep = ScanReport()
        self._check(self._lib.lio_scan_step_finish(self._h, x.ctypes.data, P.ctypes.data, C.byref(rep)))
        return rep

    def set_deferred_growth(self, on=True):
        """lio_set_deferred_growth: steps return with the posterior while the scan's map growth still runs."""
        self._check(self._lib.lio_set_deferred_growth(self._h, int(on)))

    def scan_step_settle(self):
        """lio_scan_step_settle: waits for a deferred map growth; returns its three counts."""
        counts = (C.c_int32 * 3)()
        self._check(self._lib.lio_scan_step_settle(self._h, counts))
        return list(counts)

    # -- update
    def update_pass(self, x, do_search: bool, extrinsic_est: bool):
        x = _state(x)
        blob = np.zeros(90, np.float64)
        nv = C.c_int32(0)
        self._check(self._lib.lio_update_pass(self._h, _ptr(x), int(do_search), int(extrinsic_est), _ptr(blob),
                                              C.byref(nv)))
        return blob, nv.value

    def update_scan(self, x, P, R=0.001, max_iter=4, extrinsic_est=False):
        x = _state(x).copy()
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(24, 24).copy()
        nv, npass = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.lio_update_scan(self._h, _ptr(x), _ptr(P), R, max_iter, int(extrinsic_est),
                                              C.byref(nv), C.byref(npass)))
        return x, P, nv.value, npass.value

    def update_scan_host(self, down_pts4, x, P, R=0.001, max_iter=4, extrinsic_est=False):
        """Per-scan call with the downsampled cloud (n,4) float32 in host memory.  x (26,) and P (24,24) float64 are
        updated IN PLACE (no conversions on the hot call); returns (n_valid, n_passes)."""
        nv, npass = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.lio_update_scan_host(self._h, down_pts4.ctypes.data, down_pts4.shape[0], 16, x.ctypes.data,
                                                   P.ctypes.data, R, max_iter, int(extrinsic_est), C.byref(nv),
                                                   C.byref(npass)))
        return nv.value, npass.value

    def state_upload(self, x, P):
        x = _state(x)
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(576)
        self._check(self._lib.lio_state_upload(self._h, _ptr(x), _ptr(P)))

    def state_download(self):
        x = np.zeros(STATE_DOUBLES, np.float64)
        P = np.zeros((24, 24), np.float64)
        nv, npass = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.lio_state_download(self._h, _ptr(x), _ptr(P), C.byref(nv), C.byref(npass)))
        return x, P, nv.value, npass.value

    def update_enqueue(self, R=0.001, max_iter=4, extrinsic_est=False, from_snapshot=False):
        self._check(self._lib.lio_update_enqueue(self._h, R, max_iter, int(extrinsic_est), int(from_snapshot)))

    def update_begin(self, max_iter=4, extrinsic_est=False, from_snapshot=False):
        self._check(self._lib.lio_update_begin(self._h, max_iter, int(extrinsic_est), int(from_snapshot)))

    def update_pass_enqueue(self, extrinsic_est=False, x_own_min=-np.inf, x_own_max=np.inf):
        self._check(self._lib.lio_update_pass_enqueue(self._h, int(extrinsic_est), x_own_min, x_own_max))

    def update_step_enqueue(self, R=0.001, extrinsic_est=False):
        self._check(self._lib.lio_update_step_enqueue(self._h, R, int(extrinsic_est)))

    @property
    def blob_device_ptr(self) -> int:
        return int(self._lib.lio_blob_device_ptr(self._h) or 0)

    def pass_only_enqueue(self, do_search: bool, extrinsic_est=False):
        self._check(self._lib.lio_pass_only_enqueue(self._h, int(do_search), int(extrinsic_est)))

    def blob_upload(self, blob92):
        b = np.ascontiguousarray(blob92, np.float64).reshape(BLOB)
        self._check(self._lib.lio_blob_upload(self._h, _ptr(b)))

    def blob_bind(self, device_ptr: int | None):
        self._check(self._lib.lio_blob_bind(self._h, C.c_void_p(device_ptr or 0)))

    def peer_handle(self) -> np.ndarray:
        h = np.zeros(64, np.uint8)
        self._check(self._lib.lio_peer_handle(self._h, _ptr(h)))
        return h

    def peer_connect(self, rank: int, world: int, handles):
        h = np.ascontiguousarray(handles, np.uint8).reshape(world, 64)
        self._check(self._lib.lio_peer_connect(self._h, rank, world, _ptr(h)))

    def update_enqueue_sharded(self, R=0.001, max_iter=4, extrinsic_est=False, from_snapshot=True, x_own_min=-np.inf,
                               x_own_max=np.inf):
        self._check(self._lib.lio_update_enqueue_sharded(self._h, R, max_iter, int(extrinsic_est), int(from_snapshot),
                                                         x_own_min, x_own_max))

    def set_shard_stripes(self, x_origin: float, stripe_width: float, world: int, rank: int):
        self._check(self._lib.lio_set_shard_stripes(self._h, x_origin, stripe_width, world, rank))

    def peer_timed_out(self) -> bool:
        v = C.c_int32(0)
        self._check(self._lib.lio_peer_status(self._h, C.byref(v)))
        return bool(v.value)

    def debug_timeline(self):
        """[(tag, ns)] of block 0 and of the solving block (needs LIO_TIMELINE=1 at context creation)."""
        t = np.zeros(256, np.int64)
        self._check(self._lib.lio_debug_timeline(self._h, _ptr(t)))
        a = [(int(t[1 + 2 * k]), int(t[2 + 2 * k])) for k in range(int(t[0]))]
        b = [(int(t[129 + 2 * k]), int(t[130 + 2 * k])) for k in range(int(t[128]))]
        self.timeline_raw = t
        return a, b

    def debug_blocks(self):
        """Per-block times of the last pass (lio_debug_blocks): workers' rows filed, searches left, solver warps."""
        t = np.zeros(768, np.int64)
        self._check(self._lib.lio_debug_blocks(self._h, _ptr(t)))
        return t

    def blob_download(self):
        b = np.zeros(BLOB, np.float64)
        self._check(self._lib.lio_blob_download(self._h, _ptr(b)))
        return b

    def get_neighbors(self, m: int):
        idx = np.full((m, 5), -1, np.int32)
        d2 = np.full((m, 5), np.inf, np.float32)
        nbr = np.zeros((m, 5, 3), np.float32)
        world = np.zeros((m, 3), np.float32)
        sel = np.zeros(m, np.uint8)
        nv = np.zeros((m, 4), np.float32)
        self._check(self._lib.lio_get_neighbors(self._h, _ptr(idx), _ptr(d2), _ptr(nbr), _ptr(world), _ptr(sel),
                                                _ptr(nv)))
        return dict(idx=idx, d2=d2, nbr=nbr, world=world, selected=sel, normvec=nv)

    def map_build_scan(self, x):
        x = _state(x)
        self._check(self._lib.lio_map_build_scan(self._h, _ptr(x)))

    def map_incremental(self, x, filter_size_map=0.5, ekf_inited=True):
        x = _state(x)
        counts = np.zeros(3, np.int32)
        self._check(self._lib.lio_map_incremental(self._h, _ptr(x), filter_size_map, int(ekf_inited), _ptr(counts)))
        return counts


class Sequence:
    """One lio_seq: the per-scan main loop (laserMapping.cpp:702-800) of one sequence on one Context, in native code."""

    def __init__(self, ctx: Context, **cfg):
        self._lib = load_library()
        self.ctx = ctx
        self.cfg = SeqConfig()
        self._lib.lio_seq_default_config(C.byref(self.cfg))
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise TypeError(f"unknown lio_seq_config field {k}")
            if k in ("extrinsic_T", "extrinsic_R"):
                v = np.asarray(v, np.float64).ravel()
                v = (C.c_double * len(v))(*v)
            setattr(self.cfg, k, v)
        self._h = C.c_void_p()
        rc = self._lib.lio_seq_create(ctx._h, C.byref(self.cfg), C.byref(self._h))
        if rc:
            raise LioError(rc, "lio_seq_create")
        self._keep = None

    def close(self):
        if self._h:
            self._lib.lio_seq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def input(self, lidar, imu, lidar_beg_time, lidar_end_time) -> SeqInput:
        """Pack one MeasureGroup; the arrays must stay alive until the call that consumes the SeqInput returns."""
        pts, stride = _points(lidar)
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
        si = SeqInput(pts.ctypes.data, pts.shape[0], stride, imu.shape[0], imu.ctypes.data if imu.shape[0] else None,
                      float(lidar_beg_time), float(lidar_end_time))
        si._keep = (pts, imu)
        return si

    def process(self, lidar, imu, lidar_beg_time, lidar_end_time) -> SeqResult:
        si = self.input(lidar, imu, lidar_beg_time, lidar_end_time)
        res = SeqResult()
        rc = self._lib.lio_seq_process(self._h, C.byref(si), C.byref(res))
        if rc:
            raise LioError(rc, (self._lib.lio_last_error(self.ctx._h) or b"").decode())
        return res

    def get_state(self):
        x, P = np.zeros(STATE_DOUBLES), np.zeros((24, 24))
        self._lib.lio_seq_get_state(self._h, _ptr(x), _ptr(P))
        return x, P

    def set_state(self, x, P):
        x = _state(x)
        P = np.ascontiguousarray(P, np.float64).reshape(576)
        self._lib.lio_seq_set_state(self._h, _ptr(x), _ptr(P))

    def local_map(self):
        box = np.zeros(6, np.float32)
        nd = C.c_int64(0)
        rc = self._lib.lio_seq_local_map(self._h, _ptr(box), C.byref(nd))
        return (box.reshape(2, 3) if rc == 0 else None), int(nd.value)


def set_host_threads(n: int):
    """lio_set_host_threads: host threads of lio_seq_process_many (0 = default)."""
    if load_library().lio_set_host_threads(int(n)):
        raise ValueError("host threads must be in [0, 64]")


def seq_process_many(seqs, inputs):
    """lio_seq_process_many: one main-loop iteration of several independent sequences (one SeqInput each)."""
    n = len(seqs)
    hs = (C.c_void_p * n)(*[s._h for s in seqs])
    ins = (SeqInput * n)(*inputs)
    res = (SeqResult * n)()
    rc = load_library().lio_seq_process_many(hs, n, ins, res)
    if rc:
        raise LioError(rc, "; ".join((s._lib.lio_last_error(s.ctx._h) or b"").decode() for s in seqs))
    return list(res)


def update_enqueue_multi(ctxs, R=0.001, max_iter=4, extrinsic_est=False, from_snapshot=True):
    """n <= 8 independent updates (one per Context, same device) in one cooperative launch."""
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    rc = load_library().lio_update_enqueue_multi(arr, len(ctxs), R, max_iter, int(extrinsic_est), int(from_snapshot))
    if rc:
        raise LioError(rc, (ctxs[0]._lib.lio_last_error(ctxs[0]._h) or b"").decode())


# -- host-side sequential pieces (no context needed)
def predict(x, P, dt, Q, acc, gyro):
    lib = load_library()
    x = _state(x).copy()
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(24, 24).copy()
    Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(144)
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    gyro = np.ascontiguousarray(gyro, dtype=np.float64)
    rc = lib.lio_predict(_ptr(x), _ptr(P), float(dt), _ptr(Q), _ptr(acc), _ptr(gyro))
    if rc:
        raise LioError(rc, "lio_predict")
    return x, P


def boxplus(x, f):
    lib = load_library()
    x = _state(x)
    f = np.ascontiguousarray(f, dtype=np.float64).reshape(24)
    out = np.zeros(STATE_DOUBLES)
    rc = lib.lio_boxplus(_ptr(x), _ptr(f), _ptr(out))
    if rc:
        raise LioError(rc, "lio_boxplus")
    return out


def boxminus(x1, x2):
    lib = load_library()
    x1, x2 = _state(x1), _state(x2)
    out = np.zeros(24)
    rc = lib.lio_boxminus(_ptr(x1), _ptr(x2), _ptr(out))
    if rc:
        raise LioError(rc, "lio_boxminus")
    return out


IMU_PROC_BYTES = 51 * 8 + 16  # lio_imu_proc: 51 doubles + 4 int32


class ImuProc:
    """Host-side ImuProcess mirror (lio_imu_proc): IMU_init + the forward half of UndistortPcl."""

    def __init__(self):
        self._lib = load_library()
        self._buf = np.zeros(IMU_PROC_BYTES // 8, np.float64)
        self._lib.lio_imu_proc_init(_ptr(self._buf))

    def set_param(self, transl=(0, 0, 0), rot=None, gyr=(0.1,) * 3, acc=(0.1,) * 3, gyr_bias=(1e-4,) * 3,
                  acc_bias=(1e-4,) * 3):
        a = [np.ascontiguousarray(v, np.float64) for v in (transl, np.eye(3) if rot is None else rot, gyr, acc, gyr_bias,
                                                            acc_bias)]
        self._lib.lio_imu_set_param(_ptr(self._buf), *[_ptr(v) for v in a])

    @property
    def need_init(self) -> bool:
        return bool(self._buf[51:].view(np.int32)[1])

    def process(self, imu, lidar_beg_time, lidar_end_time, x, P):
        """imu: (n,7) [stamp, acc3, gyr3].  Returns (x, P, poses (n_poses,22), initialising)."""
        imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
        x = _state(x).copy()
        P = np.ascontiguousarray(P, np.float64).reshape(24, 24).copy()
        cap = imu.shape[0] + 2
        poses = np.zeros((cap, POSE_DOUBLES))
        npos, ini = C.c_int(0), C.c_int(0)
        rc = self._lib.lio_imu_process(_ptr(self._buf), _ptr(imu), imu.shape[0], float(lidar_beg_time),
                                       float(lidar_end_time), _ptr(x), _ptr(P), _ptr(poses), cap, C.byref(npos),
                                       C.byref(ini))
        if rc:
            raise LioError(rc, "lio_imu_process")
        return x, P, poses[: npos.value].copy(), bool(ini.value)
