"""Synthetic scenes, scans and trajectories of the shapes BASELINE.json names.

Pure NumPy, seeded (np.random.default_rng(seed) -> PCG64, bit-reproducible across machines), no file or network
access.  Everything is produced once on the host and handed, as the same arrays, to the CUDA path and to the
oracle, so both sides see identical bits.

Scene model: a set of axis-aligned rectangles (ground, walls, roofs, pillars).  Maps are point samples of the
rectangles at one point per 0.5 m surface voxel; scans are ray casts against the same rectangles.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

G = 9.81


# ------------------------------------------------------------------------------------------- scenes
@dataclass
class Scene:
    """rects: (P, 7) float64 rows [axis, c, u0, u1, v0, v1, _]: the plane {x_axis = c}, spanning [u0,u1]x[v0,v1] on
    the two remaining axes taken in cyclic order (axis+1, axis+2)."""

    rects: np.ndarray
    name: str = "scene"


def _box_rects(x0, x1, y0, y1, z0, z1, top=True):
    r = [
        [0, x0, y0, y1, z0, z1, 0], [0, x1, y0, y1, z0, z1, 0],
        [1, y0, z0, z1, x0, x1, 0], [1, y1, z0, z1, x0, x1, 0],
    ]  # fmt: skip
    if top:
        r.append([2, z1, x0, x1, y0, y1, 0])
    return r


def hall_scene(half=40.0, height=8.0) -> Scene:
    """Config 1 'hall': ground z=0, ceiling-less, 4 walls, 6 box pillars."""
    r = [[2, 0.0, -half, half, -half, half, 0]]
    r += _box_rects(-half, half, -half, half, 0.0, height, top=False)
    for k, (px, py) in enumerate([(-20, -15), (-20, 15), (0, -22), (0, 22), (20, -15), (20, 15)]):
        r += _box_rects(px - 1.5, px + 1.5, py - 1.5, py + 1.5, 0.0, height - 1.0, top=True)
    return Scene(np.asarray(r, np.float64), "hall")


def city_scene(extent=500.0, pitch=40.0, footprint=20.0, height=15.0) -> Scene:
    """Configs 3/5 'city': ground plus a grid of box buildings (footprint every `pitch` metres)."""
    h = extent / 2
    r = [[2, 0.0, -h, h, -h, h, 0]]
    n = int(extent // pitch)
    off = -(n - 1) * pitch / 2
    for i in range(n):
        for j in range(n):
            cx, cy = off + i * pitch + pitch / 2 * 0.5, off + j * pitch + pitch / 2 * 0.5
            r += _box_rects(cx - footprint / 2, cx + footprint / 2, cy - footprint / 2, cy + footprint / 2, 0.0, height)
    return Scene(np.asarray(r, np.float64), "city")


def block_scene(extent=120.0) -> Scene:
    """Config 2/4 'block': 120 m x 120 m yard, perimeter walls, a handful of buildings."""
    h = extent / 2
    r = [[2, 0.0, -h, h, -h, h, 0]]
    r += _box_rects(-h, h, -h, h, 0.0, 10.0, top=False)
    for (cx, cy, sx, sy, hz) in [(-30, -30, 16, 12, 9), (28, -25, 12, 20, 12), (-25, 30, 20, 10, 7), (30, 32, 14, 14, 10),
                                 (0, 0, 6, 6, 5)]:
        r += _box_rects(cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, 0.0, hz)
    return Scene(np.asarray(r, np.float64), "block")


def sample_map(scene: Scene, n_points: int | None, seed: int, voxel=0.5, sigma_n=0.01) -> np.ndarray:
    """One point per `voxel` x `voxel` surface cell (uniform jitter in the cell, N(0, sigma_n) along the normal).
    Returns (n,3) float32.  If n_points is given the sample is cut (uniformly at random) or must already fit."""
    rng = np.random.default_rng(seed)
    out = []
    for axis, c, u0, u1, v0, v1, _ in scene.rects:
        axis = int(axis)
        nu, nv = max(1, int(np.ceil((u1 - u0) / voxel))), max(1, int(np.ceil((v1 - v0) / voxel)))
        uu, vv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
        u = u0 + (uu.ravel() + rng.random(nu * nv)) * voxel
        v = v0 + (vv.ravel() + rng.random(nu * nv)) * voxel
        keep = (u <= u1) & (v <= v1)
        w = c + rng.normal(0.0, sigma_n, nu * nv)
        p = np.empty((nu * nv, 3))
        p[:, axis] = w
        p[:, (axis + 1) % 3] = u
        p[:, (axis + 2) % 3] = v
        out.append(p[keep])
    pts = np.concatenate(out)
    if n_points is not None:
        if pts.shape[0] < n_points:
            raise ValueError(f"scene yields {pts.shape[0]} points < requested {n_points}; enlarge the scene")
        sel = rng.permutation(pts.shape[0])[:n_points]
        sel.sort()
        pts = pts[sel]
    return pts.astype(np.float32)


def city_map(n_points: int, seed: int):
    """A city scene sized so that it yields >= n_points surface voxels, cut to exactly n_points."""
    # area per 40 m tile: 1600 ground + 1600 building skin -> ~12800 pts
    extent = max(120.0, 40.0 * np.ceil(np.sqrt(n_points / 12000.0) + 1))
    scene = city_scene(extent=float(extent))
    return scene, sample_map(scene, n_points, seed)


# ------------------------------------------------------------------------------------------- ray casting
def raycast(scene: Scene, origins: np.ndarray, dirs: np.ndarray, max_range: float, chunk=16384) -> np.ndarray:
    """Nearest hit distance per ray (inf when none within max_range).  origins/dirs: (N,3) float64, world frame."""
    N = dirs.shape[0]
    origins = np.broadcast_to(origins, dirs.shape)
    rects = scene.rects
    # cull rectangles that cannot be reached from any origin
    o_min, o_max = origins.min(0) - max_range, origins.max(0) + max_range
    keep = np.ones(len(rects), bool)
    for k, (axis, c, u0, u1, v0, v1, _) in enumerate(rects):
        a = int(axis)
        b, d = (a + 1) % 3, (a + 2) % 3
        keep[k] = (o_min[a] <= c <= o_max[a]) and u1 >= o_min[b] and u0 <= o_max[b] and v1 >= o_min[d] and v0 <= o_max[d]
    rects = rects[keep]
    best = np.full(N, np.inf)
    for s in range(0, N, chunk):
        o, d = origins[s:s + chunk], dirs[s:s + chunk]
        tb = np.full(o.shape[0], np.inf)
        for axis, c, u0, u1, v0, v1, _ in rects:
            a = int(axis)
            b, e = (a + 1) % 3, (a + 2) % 3
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (c - o[:, a]) / d[:, a]
            hu = o[:, b] + t * d[:, b]
            hv = o[:, e] + t * d[:, e]
            ok = (t > 0.3) & (t < tb) & (hu >= u0) & (hu <= u1) & (hv >= v0) & (hv <= v1)
            tb = np.where(ok, t, tb)
        best[s:s + chunk] = tb
    best[best > max_range] = np.inf
    return best


def rot_zyx(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def mat_to_quat_wxyz(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        v = np.zeros(3)
        v[i] = 0.25 * s
        v[j] = (R[j, i] + R[i, j]) / s
        v[k] = (R[k, i] + R[i, k]) / s
        q = np.array([(R[k, j] - R[j, k]) / s, *v])
    return q / np.linalg.norm(q)


def make_state(pos=(0, 0, 0), R=None, vel=(0, 0, 0), bg=(0, 0, 0), ba=(0, 0, 0), grav=(0, 0, -G), R_LI=None,
               t_LI=(0, 0, 0)) -> np.ndarray:
    """Flat lio_state (26 doubles): pos3, rot wxyz, R_LI wxyz, t_LI3, vel3, bg3, ba3, grav3."""
    x = np.zeros(26)
    x[0:3] = pos
    x[3:7] = mat_to_quat_wxyz(np.eye(3) if R is None else R)
    x[7:11] = mat_to_quat_wxyz(np.eye(3) if R_LI is None else R_LI)
    x[11:14] = t_LI
    x[14:17] = vel
    x[17:20] = bg
    x[20:23] = ba
    x[23:26] = grav
    return x


def spinning_dirs(rings: int, cols: int, fov_down_deg: float, fov_up_deg: float):
    """Unit ray directions in the sensor frame, column-major (all rings of column 0 first), and column index."""
    el = np.deg2rad(np.linspace(fov_down_deg, fov_up_deg, rings))
    az = 2 * np.pi * np.arange(cols) / cols
    azg, elg = np.meshgrid(az, el, indexing="ij")  # (cols, rings)
    d = np.stack([np.cos(elg) * np.cos(azg), np.cos(elg) * np.sin(azg), np.sin(elg)], -1).reshape(-1, 3)
    col = np.repeat(np.arange(cols), rings)
    return d, col


def avia_dirs(n: int, seed: int, hfov_deg=70.4, vfov_deg=77.2):
    """Livox-Avia-like non-repetitive rosette inside an hfov x vfov window looking along +x."""
    rng = np.random.default_rng(seed)
    i = np.arange(n)
    # two incommensurate rotors + small dither
    a = 2 * np.pi * (i * 0.61803398875 % 1.0)
    r = np.sqrt((i * 0.7548776662 % 1.0))
    y = r * np.cos(a) * np.tan(np.deg2rad(hfov_deg / 2)) + rng.normal(0, 1e-3, n)
    z = r * np.sin(a) * np.tan(np.deg2rad(vfov_deg / 2)) + rng.normal(0, 1e-3, n)
    d = np.stack([np.ones(n), y, z], -1)
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def static_scan(scene: Scene, dirs_sensor: np.ndarray, times_ms: np.ndarray, pos, R, max_range, seed,
                sigma_r=0.02) -> np.ndarray:
    """Scan from a fixed sensor pose (world R, pos).  Returns (n,4) float32 [x,y,z,t_ms] in the SENSOR frame, rays
    without a hit dropped, order preserved (time ascending when times_ms is)."""
    rng = np.random.default_rng(seed)
    dw = dirs_sensor @ np.asarray(R).T
    t = raycast(scene, np.asarray(pos, float)[None, :], dw, max_range)
    ok = np.isfinite(t)
    rr = t[ok] + rng.normal(0, sigma_r, ok.sum())
    p = dirs_sensor[ok] * rr[:, None]
    return np.concatenate([p, times_ms[ok, None]], 1).astype(np.float32)


# ------------------------------------------------------------------------------------------- trajectories
@dataclass
class Trajectory:
    """Smooth analytic 6-DoF motion: Lissajous position, gently varying attitude."""

    ax: float = 20.0
    ay: float = 14.0
    az: float = 0.3
    wx: float = 0.21
    wy: float = 0.14
    wz: float = 0.33
    yaw_amp: float = 1.2
    yaw_w: float = 0.17
    tilt_amp: float = 0.06
    tilt_w: float = 0.41
    z0: float = 1.8

    def pos(self, t):
        t = np.asarray(t, float)
        return np.stack([self.ax * np.sin(self.wx * t), self.ay * np.sin(self.wy * t + 0.7),
                         self.z0 + self.az * np.sin(self.wz * t)], -1)

    def vel(self, t):
        t = np.asarray(t, float)
        return np.stack([self.ax * self.wx * np.cos(self.wx * t), self.ay * self.wy * np.cos(self.wy * t + 0.7),
                         self.az * self.wz * np.cos(self.wz * t)], -1)

    def acc(self, t):
        t = np.asarray(t, float)
        return np.stack([-self.ax * self.wx ** 2 * np.sin(self.wx * t),
                         -self.ay * self.wy ** 2 * np.sin(self.wy * t + 0.7),
                         -self.az * self.wz ** 2 * np.sin(self.wz * t)], -1)

    def rot(self, t):
        yaw = self.yaw_amp * np.sin(self.yaw_w * t)
        pitch = self.tilt_amp * np.sin(self.tilt_w * t)
        roll = self.tilt_amp * np.cos(self.tilt_w * t * 0.7)
        return rot_zyx(yaw, pitch, roll)

    def omega_body(self, t, h=1e-5):
        Rm, Rp = self.rot(t - h), self.rot(t + h)
        dR = Rm.T @ Rp
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2.0
        return w / (2 * h)


def imu_stream(traj: Trajectory, t0: float, t1: float, rate_hz: float, seed: int, sigma_g=1e-3, sigma_a=1e-2,
               bg=(0.002, 0.002, 0.002), ba=(0.02, 0.02, 0.02)) -> np.ndarray:
    """(n,7) [stamp, acc3 (specific force, body), gyr3 (body)] with white noise and constant biases."""
    rng = np.random.default_rng(seed)
    ts = np.arange(int(np.floor(t0 * rate_hz)) + 1, int(np.floor(t1 * rate_hz)) + 1) / rate_hz
    out = np.zeros((ts.size, 7))
    gvec = np.array([0, 0, -G])
    for k, t in enumerate(ts):
        R = traj.rot(t)
        f = R.T @ (traj.acc(t) - gvec)
        out[k, 0] = t
        out[k, 1:4] = f + np.asarray(ba) + rng.normal(0, sigma_a, 3)
        out[k, 4:7] = traj.omega_body(t) + np.asarray(bg) + rng.normal(0, sigma_g, 3)
    return out


def moving_scan(scene: Scene, traj: Trajectory, t_beg: float, period: float, rings: int, cols: int, fov_down: float,
                fov_up: float, max_range: float, seed: int, sigma_r=0.02) -> np.ndarray:
    """One motion-distorted revolution starting at t_beg: column c is fired at t_beg + c/cols*period from the pose
    at that instant (lidar frame == IMU frame).  Returns (n,4) float32 [x,y,z,t_ms], time ascending."""
    rng = np.random.default_rng(seed)
    d, col = spinning_dirs(rings, cols, fov_down, fov_up)
    tc = t_beg + np.arange(cols) / cols * period
    Rc = np.stack([traj.rot(t) for t in tc])  # (cols,3,3)
    pc = traj.pos(tc)
    dw = np.einsum("nij,nj->ni", Rc[col], d)
    t = raycast(scene, pc[col], dw, max_range)
    ok = np.isfinite(t)
    rr = t[ok] + rng.normal(0, sigma_r, ok.sum())
    p = d[ok] * rr[:, None]
    tms = (col[ok] / cols * period * 1000.0)
    return np.concatenate([p, tms[:, None]], 1).astype(np.float32)


@dataclass
class RampedTrajectory:
    """`base` replayed on a warped clock that stands still until t_start and reaches unit rate T seconds later (rate =
    smoothstep): the platform is at rest while the filter initialises from the first IMU samples, then moves off."""

    base: Trajectory
    t_start: float = 0.45
    T: float = 1.5

    def _tau(self, t):
        u = np.clip((np.asarray(t, float) - self.t_start) / self.T, 0.0, None)
        uc = np.minimum(u, 1.0)
        tau = self.T * (uc ** 3 - 0.5 * uc ** 4) + self.T * np.maximum(u - 1.0, 0.0)
        d1 = np.where(u < 1.0, 3 * uc ** 2 - 2 * uc ** 3, 1.0)
        d2 = np.where(u < 1.0, (6 * uc - 6 * uc ** 2) / self.T, 0.0)
        return tau, d1, d2

    def pos(self, t):
        return self.base.pos(self._tau(t)[0])

    def vel(self, t):
        tau, d1, _ = self._tau(t)
        return self.base.vel(tau) * np.asarray(d1)[..., None]

    def acc(self, t):
        tau, d1, d2 = self._tau(t)
        return self.base.acc(tau) * np.asarray(d1 ** 2)[..., None] + self.base.vel(tau) * np.asarray(d2)[..., None]

    def rot(self, t):
        return self.base.rot(float(self._tau(t)[0]))

    def omega_body(self, t, h=1e-5):
        Rm, Rp = self.rot(t - h), self.rot(t + h)
        dR = Rm.T @ Rp
        w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2.0
        return w / (2 * h)


def sequence(n_scans: int, seed: int, rings=16, cols=1800, fov=(-15.0, 15.0), max_range=100.0, scan_hz=10.0,
             imu_hz=200.0, scene: Scene | None = None, traj=None):
    """Config 2 / 4 shape: `n_scans` contiguous motion-distorted revolutions with their IMU samples, packaged the way
    sync_packages (src/laserMapping.cpp:218-275) hands them to the main loop: list of dicts
    {lidar (n,4) [x,y,z,t_ms], imu (k,7) [stamp, acc3, gyr3], lidar_beg_time, lidar_end_time} plus ground truth."""
    scene = scene or block_scene()
    traj = traj or RampedTrajectory(Trajectory())
    period = 1.0 / scan_hz
    t_first = 0.05
    imu = imu_stream(traj, 0.0, t_first + n_scans * period + 0.05, imu_hz, seed + 1)
    out = []
    k_imu = 0
    mean_scantime, scan_num = 0.0, 0
    for k in range(n_scans):
        t_beg = t_first + k * period
        pts = moving_scan(scene, traj, t_beg, period, rings, cols, fov[0], fov[1], max_range, seed + 10 + k)
        last = float(pts[-1, 3]) / 1000.0 if len(pts) else 0.0
        if len(pts) <= 5 or last < 0.5 * mean_scantime:
            t_end = t_beg + mean_scantime
        else:
            scan_num += 1
            t_end = t_beg + last
            mean_scantime += (last - mean_scantime) / scan_num
        j = k_imu
        while j < len(imu) and imu[j, 0] <= t_end:
            j += 1
        out.append(dict(lidar=pts, imu=imu[k_imu:j].copy(), lidar_beg_time=t_beg, lidar_end_time=t_end,
                        truth_pos=traj.pos(t_end), truth_R=traj.rot(t_end)))
        k_imu = j
    return out


# ------------------------------------------------------------------------------------------- named configs
def perturbed_prior(x_true: np.ndarray, seed: int, dpos=0.05, drot_deg=1.0) -> np.ndarray:
    """Truth pose + (5 cm, 1 deg) perturbation: the propagated prior of a single-scan update."""
    rng = np.random.default_rng(seed)
    x = x_true.copy()
    dp = rng.normal(size=3)
    x[0:3] += dpos * dp / np.linalg.norm(dp)
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(drot_deg)
    dq = np.array([np.cos(ang / 2), *(np.sin(ang / 2) * ax)])
    w0, x0, y0, z0 = x[3:7]
    w1, x1, y1, z1 = dq
    q = np.array([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                  w0 * y1 + y0 * w1 + z0 * x1 - x0 * z1, w0 * z1 + z0 * w1 + x0 * y1 - y0 * x1])
    x[3:7] = q / np.linalg.norm(q)
    return x


def init_P() -> np.ndarray:
    """init_P of ImuProcess::IMU_init (src/IMU_Processing.hpp:233-238)."""
    P = np.eye(24)
    P[6:12, 6:12] = np.eye(6) * 1e-5
    P[15:18, 15:18] = np.eye(3) * 1e-4
    P[18:21, 18:21] = np.eye(3) * 1e-3
    P[21:24, 21:24] = np.eye(3) * 1e-5
    return P


def config1_avia(seed=1001, n_map=200_000, n_rays=24_000):
    """Config 1: 24k-ray Avia-style scan vs a 200k-point hall map."""
    scene = hall_scene(half=110.0, height=10.0)
    mp = sample_map(scene, n_map, seed)
    R = rot_zyx(0.3, 0.45, -0.01)  # nose down: the 70 x 77 deg window looks at the floor and the far wall
    pos = np.array([3.0, -2.0, 3.0])
    d = avia_dirs(n_rays, seed + 1)
    tms = (np.arange(n_rays) / 240000.0 * 1000.0)
    scan = static_scan(scene, d, tms, pos, R, 100.0, seed + 2)
    x_true = make_state(pos=pos, R=R)
    return dict(scene=scene, map=mp, scan=scan, x_true=x_true, x_prior=perturbed_prior(x_true, seed + 3), P=init_P(),
                leaf=0.5, max_iter=4, extrinsic_est=False)


def config3_os1_128(seed=3003, n_map=2_000_000, rings=128, cols=1024, pose_index=0):
    """Config 3: OS1-128 (131,072 rays) vs a 2M-point city map, single-scan update."""
    scene, mp = city_map(n_map, seed)
    rng = np.random.default_rng(seed + 17 * pose_index + 1)
    pos = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), 2.0])
    R = rot_zyx(rng.uniform(-np.pi, np.pi), rng.normal(0, 0.02), rng.normal(0, 0.02))
    d, col = spinning_dirs(rings, cols, -22.5, 22.5)
    tms = col / cols * 100.0
    scan = static_scan(scene, d, tms, pos, R, 120.0, seed + 2 + pose_index)
    x_true = make_state(pos=pos, R=R)
    return dict(scene=scene, map=mp, scan=scan, x_true=x_true, x_prior=perturbed_prior(x_true, seed + 3 + pose_index),
                P=init_P(), leaf=0.5, max_iter=4, extrinsic_est=False)


def small_config(seed=7, n_map=20_000, rings=16, cols=256):
    """A seconds-scale case for CPU-side tests and smoke()."""
    scene = hall_scene(half=25.0, height=6.0)
    mp = sample_map(scene, None, seed)
    if n_map is not None and mp.shape[0] > n_map:
        mp = mp[np.sort(np.random.default_rng(seed).permutation(mp.shape[0])[:n_map])]
    R = rot_zyx(0.4, 0.01, 0.02)
    pos = np.array([1.0, 2.0, 1.2])
    d, col = spinning_dirs(rings, cols, -20.0, 20.0)
    tms = col / cols * 100.0
    scan = static_scan(scene, d, tms, pos, R, 60.0, seed + 2)
    x_true = make_state(pos=pos, R=R)
    return dict(scene=scene, map=mp, scan=scan, x_true=x_true, x_prior=perturbed_prior(x_true, seed + 3), P=init_P(),
                leaf=0.5, max_iter=4, extrinsic_est=False)
