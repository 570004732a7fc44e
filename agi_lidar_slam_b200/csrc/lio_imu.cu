// Host-side, sequential half of ImuProcess (src/IMU_Processing.hpp): constructor defaults, set_param, IMU_init and the
// forward propagation of UndistortPcl.  <= 50 tiny 24x24 steps per scan: it stays on the host, exactly where the
// reference runs it; its outputs (IMUpose list, state at scan end) feed the fused per-point kernel
// (lio_scan_preprocess).  No device work in this file.
#include <math.h>
#include <string.h>

#include "lio_common.cuh"

using namespace lio;

namespace {
const double kG = 9.81;  // G_m_s2 (common_lib.h:17)
const int kMaxIniCount = 10;  // MAX_INI_COUNT (IMU_Processing.hpp:33)

inline double norm3(const double v[3]) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

// one IMUpose entry (set_pose6d, common_lib.h:63-78)
inline void put_pose(lio_pose6d* out, int& n, int cap, double t, const double a[3], const double g[3], const StateD& s) {
  if (n >= cap) return;
  lio_pose6d& p = out[n++];
  p.offset_time = t;
  for (int i = 0; i < 3; ++i) {
    p.acc[i] = a[i];
    p.gyr[i] = g[i];
    p.vel[i] = s.vel[i];
    p.pos[i] = s.pos[i];
  }
  quat_to_mat(s.rot, p.rot);
}

// IMU_init (IMU_Processing.hpp:180-244)
void imu_init(lio_imu_proc* ip, const lio_imu_sample* imu, int n, double lidar_beg_time, lio_state* xs, double P[576]) {
  int& N = ip->init_iter_num;
  if (ip->b_first_frame) {
    // Reset() (:157-167)
    ip->mean_acc[0] = ip->mean_acc[1] = 0.0;
    ip->mean_acc[2] = -1.0;
    ip->mean_gyr[0] = ip->mean_gyr[1] = ip->mean_gyr[2] = 0.0;
    ip->angvel_last[0] = ip->angvel_last[1] = ip->angvel_last[2] = 0.0;
    ip->imu_need_init = 1;
    memset(&ip->last_imu, 0, sizeof(ip->last_imu));
    N = 1;
    ip->b_first_frame = 0;
    for (int i = 0; i < 3; ++i) {
      ip->mean_acc[i] = imu[0].acc[i];
      ip->mean_gyr[i] = imu[0].gyr[i];
    }
    ip->first_lidar_time = lidar_beg_time;
  }
  for (int k = 0; k < n; ++k) {
    const double Nd = (double)N;
    for (int i = 0; i < 3; ++i) {
      const double ca = imu[k].acc[i], cg = imu[k].gyr[i];
      ip->mean_acc[i] += (ca - ip->mean_acc[i]) / Nd;
      ip->mean_gyr[i] += (cg - ip->mean_gyr[i]) / Nd;
      const double da = ca - ip->mean_acc[i], dg = cg - ip->mean_gyr[i];
      ip->cov_acc[i] = ip->cov_acc[i] * (Nd - 1.0) / Nd + da * da / Nd;
      ip->cov_gyr[i] = ip->cov_gyr[i] * (Nd - 1.0) / Nd + dg * dg / Nd / Nd * (double)(N - 1);
    }
    N++;
  }
  StateD& x = *reinterpret_cast<StateD*>(xs);
  const double na = norm3(ip->mean_acc);
  for (int i = 0; i < 3; ++i) {
    x.grav[i] = -ip->mean_acc[i] / na * kG;
    x.bg[i] = ip->mean_gyr[i];
    x.tli[i] = ip->lidar_T_wrt_imu[i];
  }
  x.rli = mat_to_quat(ip->lidar_R_wrt_imu);
  for (int i = 0; i < 576; ++i) P[i] = 0.0;
  for (int i = 0; i < 24; ++i) P[i * 24 + i] = 1.0;
  for (int i = 6; i < 12; ++i) P[i * 24 + i] = 0.00001;
  for (int i = 15; i < 18; ++i) P[i * 24 + i] = 0.0001;
  for (int i = 18; i < 21; ++i) P[i * 24 + i] = 0.001;
  for (int i = 21; i < 24; ++i) P[i * 24 + i] = 0.00001;
  ip->last_imu = imu[n - 1];
}

// forward half of UndistortPcl (IMU_Processing.hpp:258-358)
int imu_forward(lio_imu_proc* ip, const lio_imu_sample* imu, int n, double pcl_beg, double pcl_end, lio_state* xs,
                double P[576], lio_pose6d* poses, int cap) {
  StateD& x = *reinterpret_cast<StateD*>(xs);
  int np = 0;
  put_pose(poses, np, cap, 0.0, ip->acc_s_last, ip->angvel_last, x);
  // process_noise_cov() (use-ikfom.hpp:40-48); the diagonal is overwritten before every step (:319-322)
  double Q[144];
  memset(Q, 0, sizeof(Q));
  for (int i = 0; i < 3; ++i) {
    Q[i * 13] = 0.0001;
    Q[(3 + i) * 13] = 0.0001;
    Q[(6 + i) * 13] = 0.00001;
    Q[(9 + i) * 13] = 0.00001;
  }
  double in_acc[3] = {0, 0, 0}, in_gyr[3] = {0, 0, 0};
  const double last_end = ip->last_lidar_end_time;
  double imu_end_time = ip->last_imu.stamp;
  // v_imu = {last_imu_, meas.imu...}; walk consecutive pairs
  for (int k = 0; k < n; ++k) {
    const lio_imu_sample& head = k == 0 ? ip->last_imu : imu[k - 1];
    const lio_imu_sample& tail = imu[k];
    imu_end_time = tail.stamp;
    if (tail.stamp < last_end) continue;
    for (int i = 0; i < 3; ++i) {
      in_gyr[i] = 0.5 * (head.gyr[i] + tail.gyr[i]);
      in_acc[i] = 0.5 * (head.acc[i] + tail.acc[i]) * kG / norm3(ip->mean_acc);
      Q[i * 13] = ip->cov_gyr[i];
      Q[(3 + i) * 13] = ip->cov_acc[i];
      Q[(6 + i) * 13] = ip->cov_bias_gyr[i];
      Q[(9 + i) * 13] = ip->cov_bias_acc[i];
    }
    const double dt = head.stamp < last_end ? tail.stamp - last_end : tail.stamp - head.stamp;
    lio_predict(xs, P, dt, Q, in_acc, in_gyr);
    double am[3], aw[3];
    for (int i = 0; i < 3; ++i) {
      ip->angvel_last[i] = tail.gyr[i] - x.bg[i];
      am[i] = tail.acc[i] * kG / norm3(ip->mean_acc) - x.ba[i];
    }
    quat_rotate(x.rot, am, aw);
    for (int i = 0; i < 3; ++i) ip->acc_s_last[i] = aw[i] + x.grav[i];
    put_pose(poses, np, cap, tail.stamp - pcl_beg, ip->acc_s_last, ip->angvel_last, x);
  }
  lio_predict(xs, P, fabs(pcl_end - imu_end_time), Q, in_acc, in_gyr);  // :353-354
  if (n > 0) ip->last_imu = imu[n - 1];
  ip->last_lidar_end_time = pcl_end;
  return np;
}
}  // namespace

extern "C" {

void lio_imu_proc_init(lio_imu_proc* ip) {
  if (!ip) return;
  memset(ip, 0, sizeof(*ip));
  for (int i = 0; i < 3; ++i) {
    ip->cov_acc[i] = ip->cov_gyr[i] = 0.1;
    ip->cov_bias_gyr[i] = ip->cov_bias_acc[i] = 0.0001;
    ip->cov_acc_scale[i] = ip->cov_gyr_scale[i] = 0.1;
  }
  ip->mean_acc[2] = -1.0;
  ip->lidar_R_wrt_imu[0] = ip->lidar_R_wrt_imu[4] = ip->lidar_R_wrt_imu[8] = 1.0;
  ip->init_iter_num = 1;
  ip->imu_need_init = 1;
  ip->b_first_frame = 1;
}

void lio_imu_set_param(lio_imu_proc* ip, const double transl[3], const double rot[9], const double gyr[3],
                       const double acc[3], const double gyr_bias[3], const double acc_bias[3]) {
  if (!ip) return;
  for (int i = 0; i < 3; ++i) {
    ip->lidar_T_wrt_imu[i] = transl[i];
    ip->cov_gyr_scale[i] = gyr[i];
    ip->cov_acc_scale[i] = acc[i];
    ip->cov_bias_gyr[i] = gyr_bias[i];
    ip->cov_bias_acc[i] = acc_bias[i];
  }
  for (int i = 0; i < 9; ++i) ip->lidar_R_wrt_imu[i] = rot[i];
}

int lio_imu_process(lio_imu_proc* ip, const lio_imu_sample* imu, int n_imu, double lidar_beg_time, double lidar_end_time,
                    lio_state* x, double P[576], lio_pose6d* poses, int cap, int* n_poses, int* initialising) {
  if (!ip || !x || !P || n_imu < 0 || (n_imu > 0 && !imu) || !n_poses || !initialising) return LIO_E_INVALID;
  *n_poses = 0;
  *initialising = ip->imu_need_init ? 1 : 0;
  if (n_imu == 0) return LIO_OK;  // `if (meas.imu.empty()) return;` (:408)
  if (ip->imu_need_init) {
    imu_init(ip, imu, n_imu, lidar_beg_time, x, P);
    ip->last_imu = imu[n_imu - 1];
    if (ip->init_iter_num > kMaxIniCount) {
      ip->imu_need_init = 0;
      for (int i = 0; i < 3; ++i) {
        ip->cov_acc[i] = ip->cov_acc_scale[i];
        ip->cov_gyr[i] = ip->cov_gyr_scale[i];
      }
    }
    return LIO_OK;
  }
  if (!poses || cap < n_imu + 1) return LIO_E_CAPACITY;
  *n_poses = imu_forward(ip, imu, n_imu, lidar_beg_time, lidar_end_time, x, P, poses, cap);
  return LIO_OK;
}

}  // extern "C"
