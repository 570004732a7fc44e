"""TEST INFRASTRUCTURE — the reference main loop (src/laserMapping.cpp:702-800) replayed on the CPU oracle: the
checker of the product's agi_lidar_slam_b200.replay.LioReplay.  Same MeasureGroup in, same skip rules, oracle stages."""
import numpy as np

INIT_TIME = 0.1
MAX_INI_COUNT = 10


class OracleReplay:
    def __init__(self, orc, filter_size_surf=0.5, filter_size_map=0.5, max_iteration=3, extrinsic_est=False, threads=8,
                 use_ikd=False):
        self.orc = orc
        self.fs, self.fm, self.max_iter, self.ext, self.threads = filter_size_surf, filter_size_map, max_iteration, \
            extrinsic_est, threads
        self.x = orc.default_state()
        self.P = np.eye(24)
        self.first_scan = True
        self.first_lidar_time = 0.0
        self.need_init = True
        self.first_frame = True
        self.stats = np.zeros(13)
        self.stats[2] = -1.0  # mean_acc = (0,0,-1)
        self.stats[6:12] = 0.1  # cov_acc, cov_gyr ctor defaults
        self.stats[12] = 1.0
        self.carry = orc.new_carry()
        self.carry[19] = 0.0  # last_lidar_end_time_ (uninitialised in the reference; 0 here and in the product)
        self.map = None
        self.use_ikd = use_ikd
        self.log = []

    def process(self, meas):
        orc = self.orc
        if self.first_scan:
            self.first_lidar_time = meas["lidar_beg_time"]
            self.first_scan = False
            return None
        imu = meas["imu"]
        if len(imu) == 0:
            return None
        if self.need_init:
            self.x, self.P = orc.imu_init(imu, self.first_frame, self.stats, self.x, np.zeros(3), np.eye(3))
            self.first_frame = False
            self.carry[20:27] = imu[-1]  # last_imu_
            if self.stats[12] > MAX_INI_COUNT:
                self.need_init = False
                self.carry[0:3] = 0.1  # cov_gyr = cov_gyr_scale
                self.carry[3:6] = 0.1  # cov_acc = cov_acc_scale
            self.carry[12] = np.linalg.norm(self.stats[0:3])  # mean_acc.norm()
            return None
        poses, self.x, self.P = orc.imu_forward(imu, meas["lidar_beg_time"], meas["lidar_end_time"], self.x, self.P,
                                                self.carry)
        ekf_inited = not ((meas["lidar_beg_time"] - self.first_lidar_time) < INIT_TIME)
        und, order = orc.undistort(meas["lidar"], poses, self.x)
        pts5 = np.concatenate([und[:, :3], np.zeros((len(und), 1), np.float32), und[:, 3:4]], 1)
        cen, _, _ = orc.voxel_grid(pts5, self.fs)
        body = np.ascontiguousarray(cen[:, :3])
        if len(body) < 5:
            return None
        if self.map is None:
            world = orc.body_to_world(self.x, body)
            self.map = orc.Map(1.0)
            self.map.build(world)
            self.log.append(dict(status="map-built", m=len(body)))
            return None
        sc = orc.Scan(body)
        self.x, self.P, trace, nv = sc.update(self.x, self.P, self.map.knn_backend(), 0.001, self.max_iter, self.ext,
                                              threads=self.threads)
        ref = sc.get()
        world = orc.body_to_world(self.x, body)
        cls = orc.map_incremental_classify(world, ref["near_raw"], ref["cnt"], ekf_inited, self.fm)
        na = self.map.add(world[cls == 1], True, self.fm)
        self.map.add(world[cls == 2], False)
        self.log.append(dict(status="ok", m=len(body), n_valid=nv, n_passes=len(trace),
                             counts=[int((cls == 1).sum()), int((cls == 2).sum()), int(na)]))
        return self.x.copy()
