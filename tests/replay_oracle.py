"""TEST INFRASTRUCTURE — the reference main loop (src/laserMapping.cpp:702-800) replayed on the CPU oracle: the
checker of the product's agi_lidar_slam_b200.replay.LioReplay.  Same MeasureGroup in, same skip rules, oracle stages."""
import numpy as np

INIT_TIME = 0.1
MAX_INI_COUNT = 10


class OracleReplay:
    def __init__(self, orc, filter_size_surf=0.5, filter_size_map=0.5, max_iteration=3, extrinsic_est=False, threads=8,
                 use_ikd=False, cube_len=1000.0, det_range=300.0):
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
        self.cube_len, self.det_range = cube_len, det_range
        self.local_map = None
        self.n_box_deleted = 0
        self.log = []

    def fov_segment(self):
        """lasermap_fov_segment (src/laserMapping.cpp:309-365), float arithmetic as declared there."""
        f32 = np.float32
        R = self.orc.quat_to_mat(self.x[3:7])
        pos = self.x[0:3] + R @ self.x[11:14]
        if self.local_map is None:
            self.local_map = np.array([pos - self.cube_len / 2.0, pos + self.cube_len / 2.0]).astype(f32)
            return
        thr = f32(1.5 * self.det_range)
        dist = np.zeros((3, 2), f32)
        need = False
        for i in range(3):
            dist[i, 0] = abs(pos[i] - self.local_map[0, i])
            dist[i, 1] = abs(pos[i] - self.local_map[1, i])
            if dist[i, 0] <= thr or dist[i, 1] <= thr:
                need = True
        if not need:
            return
        mov = f32(max((self.cube_len - 2.0 * 1.5 * self.det_range) * 0.5 * 0.9, float(f32(self.det_range) * f32(0.5))))
        new = self.local_map.copy()
        rm = []
        for i in range(3):
            tmp = self.local_map.copy()
            if dist[i, 0] <= thr:
                new[1, i] -= mov
                new[0, i] -= mov
                tmp[0, i] = self.local_map[1, i] - mov
                rm.append(tmp.reshape(6))
            elif dist[i, 1] <= thr:
                new[1, i] += mov
                new[0, i] += mov
                tmp[1, i] = self.local_map[0, i] + mov
                rm.append(tmp.reshape(6))
        self.local_map = new
        if rm and self.map is not None:
            self.n_box_deleted += self.map.delete_boxes(np.stack(rm))

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
        self.fov_segment()
        und, order = orc.undistort(meas["lidar"], poses, self.x)
        pts5 = np.concatenate([und[:, :3], np.zeros((len(und), 1), np.float32), und[:, 3:4]], 1)
        cen, _, _ = orc.voxel_grid(pts5, self.fs)
        body = np.ascontiguousarray(cen[:, :3])
        if len(body) < 5:
            return None
        if self.map is None:
            world = orc.body_to_world(self.x, body)
            # use_ikd: the reference's own ikd-Tree is the live map (Build / unbounded Nearest_Search / Add_Points /
            # Delete_Point_Boxes), not the hashed-grid port
            self.map = orc.IkdLiveMap(self.fm) if self.use_ikd else orc.Map(1.0)
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
        self.last = dict(cls=cls, cnt=ref["cnt"].copy(), d2=ref["d2"].copy(), world=world, near_raw=ref["near_raw"].copy(),
                         idx=ref["idx"].copy(), body=body)
        self.log.append(dict(status="ok", m=len(body), n_valid=nv, n_passes=len(trace),
                             counts=[int((cls == 1).sum()), int((cls == 2).sum()), int(na)]))
        return self.x.copy()
