"""ROS-free replay harness: the per-scan main loop of S-FAST_LIO (src/laserMapping.cpp:702-800) on the C-ABI.

One `LioReplay.process(meas)` call is one iteration of that loop for one synchronised MeasureGroup (common_lib.h:40-49:
a scan, its IMU samples, lidar_beg_time / lidar_end_time, as sync_packages (:218-275) hands them over):

    first scan: remember first_lidar_time, continue                       (:711-716)    host
    ImuProcess::Process  -> IMU_init | forward propagation                (:719)        host: lio_imu_process
    skip while the filter initialises (empty undistorted cloud)           (:722-725)    host
    flg_EKF_inited                                                        (:731-733)    host
    lasermap_fov_segment: slide the local-map box, box-delete behind it   (:736, :309-365)  lio_map_delete_boxes
    -- the rest is ONE call with one host synchronisation: lio_scan_step -------------------------------------------
    UndistortPcl per-point loop + VoxelGrid surf filter (fused)           (:719,737-738)
    fewer than 5 points -> skip                                           (:741-744)
    empty map -> pointBodyToWorld + Build, continue                       (:747-758)
    update_iterated_dyn_share_modified                                    (:772-774)
    map_incremental                                                       (:785)
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
        self.x = default_state()  # contiguous float64: the step calls update x and P in place
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

    def host_stage(self, meas: MeasureGroup):
        """The host part of one main-loop iteration: bookkeeping, IMU initialisation / forward propagation, the sliding
        local-map box.  Returns the IMU poses of the scan when the device part is due, else None."""
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
        self._ekf_inited = not ((meas.lidar_beg_time - self.first_lidar_time) < INIT_TIME)
        self._lasermap_fov_segment()
        return poses

    def adopt(self, rep):
        """Book a lio_scan_report: log it, return the odometry (or None when the scan was skipped)."""
        if rep.status == _cabi.SCAN_FEW_POINTS:
            self.log.append(dict(status="few-points", m=int(rep.m)))
            return None
        if rep.status == _cabi.SCAN_MAP_BUILT:
            self.map_built = True
            self.log.append(dict(status="map-built", m=int(rep.m)))
            return None
        self.log.append(dict(status="ok", m=int(rep.m), n_valid=int(rep.n_valid), n_passes=int(rep.n_passes),
                             counts=list(rep.counts)))
        return self.x.copy()

    def process(self, meas: MeasureGroup):
        """Returns the state after this scan (the odometry the reference publishes), or None when the scan is skipped.
        The device part is ONE C-ABI call (lio_scan_step) with one host synchronisation."""
        poses = self.host_stage(meas)
        if poses is None:
            return None
        c = self.cfg
        rep = self.ctx.scan_step(meas.lidar, poses, self.x, self.P, c.filter_size_surf, c.filter_size_map,
                                 LASER_POINT_COV, c.max_iteration, c.extrinsic_est, self._ekf_inited)
        return self.adopt(rep)


def process_many(replays, meass):
    """One main-loop iteration of several INDEPENDENT sequences on one GPU (BASELINE.json config 4): host stage and
    enqueue of the preprocessing per sequence (each on its context's stream), then the updates of all sequences that are
    due in ONE cooperative launch (lio_update_enqueue_multi, <= 8 per launch), then map growth per sequence, and one
    synchronisation per sequence at the very end.  All replays must share max_iteration and extrinsic_est.  Returns the
    list of per-sequence results of LioReplay.process."""
    out = [None] * len(replays)
    begun, due = [], []
    for k, (r, m) in enumerate(zip(replays, meass)):
        poses = r.host_stage(m)
        if poses is None:
            continue
        begun.append(k)
        if r.ctx.scan_step_begin(m.lidar, poses, r.x, r.P, r.cfg.filter_size_surf):
            due.append(k)
    if due:
        cfg = replays[due[0]].cfg
        for a in range(0, len(due), 8):
            grp = [replays[k].ctx for k in due[a:a + 8]]
            if len(grp) == 1:
                grp[0].update_enqueue(LASER_POINT_COV, cfg.max_iteration, cfg.extrinsic_est, from_snapshot=True)
            else:
                _cabi.update_enqueue_multi(grp, LASER_POINT_COV, cfg.max_iteration, cfg.extrinsic_est,
                                           from_snapshot=True)
        for k in due:
            replays[k].ctx.scan_step_end(replays[k].cfg.filter_size_map, replays[k]._ekf_inited)
    for k in begun:
        r = replays[k]
        out[k] = r.adopt(r.ctx.scan_step_finish(r.x, r.P))
    return out


class NativeReplay:
    """LioReplay's interface on the native loop (lio_seq_process, csrc/lio_seq.cu): the host stage, the bookkeeping and
    the local-map box live in C++; one C-ABI call per MeasureGroup.  `x` / `P` are pushed before and pulled after every
    call, so a caller may overwrite them between scans (relocalisation, teacher-forced parity tests)."""

    def __init__(self, ctx: _cabi.Context, cfg: ReplayConfig | None = None):
        self.ctx = ctx
        self.cfg = c = cfg or ReplayConfig()
        self.seq = _cabi.Sequence(ctx, filter_size_surf=c.filter_size_surf, filter_size_map=c.filter_size_map,
                                  max_iteration=c.max_iteration, extrinsic_est=int(c.extrinsic_est),
                                  extrinsic_T=c.extrinsic_T, extrinsic_R=np.asarray(c.extrinsic_R, np.float64),
                                  gyr_cov=c.gyr_cov, acc_cov=c.acc_cov, b_gyr_cov=c.b_gyr_cov, b_acc_cov=c.b_acc_cov,
                                  cube_len=c.cube_len, det_range=c.det_range, laser_point_cov=LASER_POINT_COV)
        self.x, self.P = self.seq.get_state()
        self.map_built = False
        self.n_box_deleted = 0
        self.log = []

    def book(self, res):
        """Log a lio_seq_result; returns the odometry or None when the scan was skipped."""
        name = _cabi.SEQ_STATUS_NAMES[res.status]
        self.n_box_deleted = int(res.n_box_deleted)
        if res.status == _cabi.SEQ_UPDATED:
            self.log.append(dict(status=name, m=int(res.m), n_valid=int(res.n_valid), n_passes=int(res.n_passes),
                                 counts=list(res.counts)))
            return self.x.copy()
        if res.status == _cabi.SEQ_MAP_BUILT:
            self.map_built = True
        self.log.append(dict(status=name, m=int(res.m)) if res.status in (_cabi.SEQ_MAP_BUILT, _cabi.SEQ_FEW_POINTS)
                        else dict(status=name))
        return None

    @property
    def local_map(self):
        return self.seq.local_map()[0]

    def process(self, meas: MeasureGroup):
        self.seq.set_state(self.x, self.P)
        res = self.seq.process(meas.lidar, meas.imu, meas.lidar_beg_time, meas.lidar_end_time)
        self.x, self.P = self.seq.get_state()
        return self.book(res)


def native_process_many(replays, meass):
    """process_many on the native loop: lio_seq_process_many."""
    for r in replays:
        r.seq.set_state(r.x, r.P)
    ins = [r.seq.input(m.lidar, m.imu, m.lidar_beg_time, m.lidar_end_time) for r, m in zip(replays, meass)]
    res = _cabi.seq_process_many([r.seq for r in replays], ins)
    out = []
    for r, e in zip(replays, res):
        r.x, r.P = r.seq.get_state()
        out.append(r.book(e))
    return out
