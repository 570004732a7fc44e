"""End-to-end parity on BASELINE.json config 2 (SURVEY.md App. C last rows): a 200-scan synthetic VLP-16 trajectory
with IMU goes through the product's replay of the reference main loop (IMU init -> forward propagation -> fused
undistort + voxel filter -> first-scan map build -> IESKF update -> map_incremental) and through the same loop on
the CPU oracle.  north_star: per-scan pose within 1e-4 m and 1e-4 rad."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_SCANS = 200
TOL_POS = 1e-4  # m    (BASELINE.json north_star)
TOL_ROT = 1e-4  # rad  (BASELINE.json north_star)


def test_config2_trajectory_pose_parity(orc):
    from agi_lidar_slam_b200 import _cabi, synth
    from agi_lidar_slam_b200.replay import LioReplay, MeasureGroup, ReplayConfig
    from replay_oracle import OracleReplay

    seq = synth.sequence(N_SCANS, 2002)  # 16 rings x 1800 columns, 10 Hz, IMU 200 Hz
    traj = synth.RampedTrajectory(synth.Trajectory())
    R0, p0 = traj.rot(0.0), traj.pos(0.0)
    with _cabi.Context(0, max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 21) as ctx:
        gpu = LioReplay(ctx, ReplayConfig(max_iteration=3))
        cpu = OracleReplay(orc, max_iteration=3)
        worst_pos, worst_rot, n_upd, same_valid = 0.0, 0.0, 0, 0
        for m in seq:
            a = gpu.process(MeasureGroup(m["lidar"], m["imu"], m["lidar_beg_time"], m["lidar_end_time"]))
            b = cpu.process(m)
            assert (a is None) == (b is None)  # same skip decisions (first scan, IMU init, map build)
            if a is None:
                continue
            n_upd += 1
            worst_pos = max(worst_pos, float(np.abs(a[0:3] - b[0:3]).max()))
            worst_rot = max(worst_rot, float(np.linalg.norm(orc.boxminus(a, b)[3:6])))
            ga, gb = gpu.log[-1], cpu.log[-1]
            assert ga["n_passes"] == gb["n_passes"]  # same search schedule / convergence decisions
            same_valid += ga["n_valid"] == gb["n_valid"]
        map_total, map_valid = ctx.map_size()
    assert n_upd >= N_SCANS - 4
    assert worst_pos < TOL_POS and worst_rot < TOL_ROT, (worst_pos, worst_rot)
    assert map_valid == cpu.map.size()  # the grown maps hold the same number of live points
    assert same_valid >= 0.9 * n_upd
    # the odometry follows the motion (scan-to-map registration on a 16-ring sensor drifts a little; both sides agree)
    pr = R0.T @ (seq[-1]["truth_pos"] - p0)
    assert np.linalg.norm(a[0:3] - pr) < 0.2 * max(1.0, np.linalg.norm(pr))
    print(f"config 2: {n_upd} updates, worst |dpos| {worst_pos:.2e} m, worst |drot| {worst_rot:.2e} rad, "
          f"map {map_valid} pts, valid-count equal in {same_valid}/{n_upd} scans")
