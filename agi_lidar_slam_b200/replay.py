"""ROS-free replay harness: the per-scan main loop of S-FAST_LIO (src/laserMapping.cpp:702-800) on the C-ABI.

One `LioReplay.process(meas)` call is one iteration of that loop for one synchronised MeasureGroup (common_lib.h:40-49:
a scan, its IMU samples, lidar_beg_time / lidar_end_time, as sync_packages (:218-275) hands them over):

    first scan: remember first_lidar_time, continue                       (:711-716)
    ImuProcess::Process  -> IMU_init | forward propagation (host)         (:719)        lio_imu_process
    skip while the filter initialises (empty undistorted cloud)           (:722-725)
    flg_EKF_inited                                                        (:731-733)
    UndistortPcl per-point loop + VoxelGrid surf filter (fused, device)   (:719,737-738) lio_scan_preprocess_resident
    fewer than 5 points -> skip                                           (:741-744)
    empty map -> pointBodyToWorld + Build, continue                       (:747-758)    lio_map_build_scan
    update_iterated_dyn_share_modified                                    (:772-774)    lio_update_scan
    map_incremental                                                       (:785)        lio_map_incremental

    lasermap_fov_segment: slide the local-map box, box-delete behind it  (:736, :309-365)  lio_map_delete_boxes
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _cabi

INIT_TIME = 0.1  # laserMapping.cpp:28
LASER_POINT_COV = 0.001  # laserMapping.cpp:29


@dataclass
class MeasureGroup:
    """common_lib.h:40-49.  lidar: (n,4) float32 [x,y,z,t_ms] or (n,12) PointXYZINormal; imu: (k,7) [stamp, acc3, gyr3]."""

    lidar: np.ndarray
    imu: np.ndarray
    lidar_beg_time: float
    lidar_end_time: float


@dataclass
class ReplayConfig:
    filter_size_surf: float = 0.5  # launch files (e.g. mapping_velodyne.launch)
    filter_size_map: float = 0.5
    max_iteration: int = 3  # mapping_velodyne.launch:10
    extrinsic_est: bool = False
    extrinsic_T: tuple = (0.0, 0.0, 0.0)
    extrinsic_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    gyr_cov: float = 0.1
    acc_cov: float = 0.1
    b_gyr_cov: float = 0.0001
    b_acc_cov: float = 0.0001
    cube_len: float = 1000.0  # cube_side_length (launch files)
    det_range: float = 300.0  # DET_RANGE (laserMapping.cpp:39; mapping/det_range)

MOV_THRESHOLD = 1.5  # laserMapping.cpp:40


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def default_state() -> np.ndarray:
    """state_ikfom defaults (use-ikfom.hpp:18-27)."""
    x = np.zeros(_cabi.STATE_DOUBLES)
    x[3] = 1.0
    x[7] = 1.0
    x[25] = -9.81
    return x


class LioReplay:
    def __init__(self, ctx: _cabi.Context, cfg: ReplayConfig | None = None):
        self.ctx = ctx
        self.cfg = cfg or ReplayConfig()
        self.imu = _cabi.ImuProc()
        c = self.cfg
        self.imu.set_param(c.extrinsic_T, c.extrinsic_R, (c.gyr_cov,) * 3, (c.acc_cov,) * 3, (c.b_gyr_cov,) * 3,
                           (c.b_acc_cov,) * 3)
        self.x = default_state()
        self.P = np.eye(24)
        self.first_scan = True
        self.first_lidar_time = 0.0
        self.map_built = False
        self.local_map = None  # LocalMap_Points: (2,3) float32 [vertex_min, vertex_max]
        self.n_box_deleted = 0
        self.log = []  # per scan: dict(status, m, n_valid, n_passes, counts)

    def _lasermap_fov_segment(self):
        """laserMapping.cpp:309-365: keep the sensor MOV_THRESHOLD * DET_RANGE away from the faces of the local-map
        cube; when it comes closer, shift the cube by mov_dist and delete the slab of map that falls out."""
        c = self.cfg
        R = _quat_to_mat(self.x[3:7])
        pos_lid = self.x[0:3] + R @ self.x[11:14]  # pos + rot * offset_T_L_I (:728-729)
        if self.local_map is None:
            self.local_map = np.stack([pos_lid - c.cube_len / 2.0, pos_lid + c.cube_len / 2.0]).astype(np.float32)
            return 0
        lm = self.local_map
        edge = np.float32(MOV_THRESHOLD * c.det_range)
        d_min = np.abs(pos_lid - lm[0]).astype(np.float32)  # float dist_to_map_edge[3][2]
        d_max = np.abs(pos_lid - lm[1]).astype(np.float32)
        if not (np.any(d_min <= edge) or np.any(d_max <= edge)):
            return 0
        mov = np.float32(max((c.cube_len - 2.0 * MOV_THRESHOLD * c.det_range) * 0.5 * 0.9,
                             float(np.float32(c.det_range) * np.float32(MOV_THRESHOLD - 1))))
        new = lm.copy()
        boxes = []
        for i in range(3):
            box = lm.copy()
            if d_min[i] <= edge:
                new[1, i] -= mov
                new[0, i] -= mov
                box[0, i] = lm[1, i] - mov
                boxes.append(box.reshape(6))
            elif d_max[i] <= edge:
                new[1, i] += mov
                new[0, i] += mov
                box[1, i] = lm[0, i] + mov
                boxes.append(box.reshape(6))
        self.local_map = new
        n = 0
        if boxes and self.map_built:
            n = self.ctx.map_delete_boxes(np.stack(boxes))
        self.n_box_deleted += n
        return n

    def process(self, meas: MeasureGroup):
        """Returns the state after this scan (the odometry the reference publishes), or None when the scan is skipped."""
        c = self.cfg
        if self.first_scan:
            self.first_lidar_time = meas.lidar_beg_time
            self.first_scan = False
            self.log.append(dict(status="first"))
            return None
        if len(meas.imu) == 0:
            self.log.append(dict(status="no-imu"))
            return None
        self.x, self.P, poses, initialising = self.imu.process(meas.imu, meas.lidar_beg_time, meas.lidar_end_time,
                                                               self.x, self.P)
        if initialising:
            self.log.append(dict(status="imu-init"))
            return None
        ekf_inited = not ((meas.lidar_beg_time - self.first_lidar_time) < INIT_TIME)
        self._lasermap_fov_segment()
        m = self.ctx.scan_preprocess(meas.lidar, poses, self.x, c.filter_size_surf, resident=True)
        if m < 5:
            self.log.append(dict(status="few-points", m=m))
            return None
        if not self.map_built:
            self.ctx.map_build_scan(self.x)
            self.map_built = True
            self.log.append(dict(status="map-built", m=m))
            return None
        self.x, self.P, nv, npass = self.ctx.update_scan(self.x, self.P, LASER_POINT_COV, c.max_iteration,
                                                         c.extrinsic_est)
        counts = self.ctx.map_incremental(self.x, c.filter_size_map, ekf_inited)
        self.log.append(dict(status="ok", m=m, n_valid=nv, n_passes=npass, counts=counts.tolist()))
        return self.x.copy()
