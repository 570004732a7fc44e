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

lasermap_fov_segment (:309-365) is not replayed: with the launch files' cube_side_length = 1000 m the local-map box
never moves inside the synthetic scenes.
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
        self.log = []  # per scan: dict(status, m, n_valid, n_passes, counts)

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
