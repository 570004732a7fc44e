// lio_seq.cu — the per-scan main loop of S-FAST_LIO (src/laserMapping.cpp:702-800) as native host code over the C-ABI:
// one lio_seq = one sequence (one robot / one bag): ImuProcess + the loop's bookkeeping + the sliding local-map box, on
// one context.  lio_seq_process is one loop iteration for one synchronised MeasureGroup; lio_seq_process_many steps
// several independent sequences together and runs their updates in one cooperative launch per scan.
//
// Host code only (no kernels here): everything the device does goes through the lio_* entry points.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/lio_b200.h"

namespace {
constexpr double kInitTime = 0.1;        // INIT_TIME (laserMapping.cpp:28)
constexpr double kMovThreshold = 1.5;    // MOV_THRESHOLD (laserMapping.cpp:40)
std::atomic<int> g_host_threads{0};      // lio_set_host_threads; 0 = default
int host_threads() {
  int n = g_host_threads.load(std::memory_order_relaxed);
  if (n <= 0) {
    const char* e = getenv("LIO_HOST_THREADS");
    n = e ? atoi(e) : 0;
    if (n <= 0) n = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
  }
  return std::max(1, std::min(n, 64));
}
}  // namespace

struct lio_seq {
  lio_ctx* ctx = nullptr;
  lio_seq_config cfg;
  lio_imu_proc imu;
  lio_state x;
  double P[576];
  bool first_scan = true;           // flg_first_scan (:711-716)
  double first_lidar_time = 0.0;
  bool ekf_inited = false;          // flg_EKF_inited (:731-733)
  bool local_map_init = false;      // Localmap_Initialized (:313-323)
  float lm_min[3], lm_max[3];       // LocalMap_Points
  int64_t n_box_deleted = 0;
  std::vector<lio_pose6d> poses;
  int n_poses = 0;
  int phase = 0;                    // 0 idle, 1 device part begun
};

namespace {

void quat_to_mat(const double q[4], double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// lasermap_fov_segment (laserMapping.cpp:309-365): keep the sensor MOV_THRESHOLD * DET_RANGE away from the faces of the
// local-map cube; when it comes closer, shift the cube by mov_dist and box-delete the slab of map that falls out.
// float arithmetic where the reference declares floats (dist_to_map_edge, mov_dist, BoxPointType).
int fov_segment(lio_seq* s) {
  double R[9];
  quat_to_mat(s->x.rot, R);
  double pos_lid[3];
  for (int i = 0; i < 3; ++i)
    pos_lid[i] = s->x.pos[i] + R[3 * i] * s->x.offset_T_L_I[0] + R[3 * i + 1] * s->x.offset_T_L_I[1] +
                 R[3 * i + 2] * s->x.offset_T_L_I[2];  // pos + rot * offset_T_L_I (:728-729)
  if (!s->local_map_init) {
    for (int i = 0; i < 3; ++i) {
      s->lm_min[i] = (float)(pos_lid[i] - s->cfg.cube_len / 2.0);
      s->lm_max[i] = (float)(pos_lid[i] + s->cfg.cube_len / 2.0);
    }
    s->local_map_init = true;
    return LIO_OK;
  }
  const float edge = (float)(kMovThreshold * s->cfg.det_range);
  float d_min[3], d_max[3];
  bool need_move = false;
  for (int i = 0; i < 3; ++i) {
    d_min[i] = (float)std::fabs(pos_lid[i] - (double)s->lm_min[i]);
    d_max[i] = (float)std::fabs(pos_lid[i] - (double)s->lm_max[i]);
    if (d_min[i] <= edge || d_max[i] <= edge) need_move = true;
  }
  if (!need_move) return LIO_OK;
  const float mov = (float)std::max((s->cfg.cube_len - 2.0 * kMovThreshold * s->cfg.det_range) * 0.5 * 0.9,
                                    (double)((float)s->cfg.det_range * (float)(kMovThreshold - 1)));
  float boxes[3][6];
  int nb = 0;
  float new_min[3], new_max[3];
  memcpy(new_min, s->lm_min, sizeof(new_min));
  memcpy(new_max, s->lm_max, sizeof(new_max));
  for (int i = 0; i < 3; ++i) {
    float* b = boxes[nb];
    memcpy(b, s->lm_min, 12);
    memcpy(b + 3, s->lm_max, 12);
    if (d_min[i] <= edge) {
      new_max[i] -= mov;
      new_min[i] -= mov;
      b[i] = s->lm_max[i] - mov;
      ++nb;
    } else if (d_max[i] <= edge) {
      new_max[i] += mov;
      new_min[i] += mov;
      b[3 + i] = s->lm_min[i] + mov;
      ++nb;
    }
  }
  memcpy(s->lm_min, new_min, sizeof(new_min));
  memcpy(s->lm_max, new_max, sizeof(new_max));
  int64_t total = 0, valid = 0;
  if (nb > 0 && lio_map_size(s->ctx, &total, &valid) == LIO_OK && total > 0) {
    int32_t deleted = 0;
    const int rc = lio_map_delete_boxes(s->ctx, &boxes[0][0], nb, &deleted);
    if (rc) return rc;
    s->n_box_deleted += deleted;
  }
  return LIO_OK;
}

// The host part of one iteration.  Returns 1 when the device part is due, 0 when the scan is done (res->status says
// why), < 0 on error.
int host_stage(lio_seq* s, const lio_seq_input* in, lio_seq_result* res) {
  memset(res, 0, sizeof(*res));
  if (s->first_scan) {  // :711-716
    s->first_lidar_time = in->lidar_beg_time;
    s->first_scan = false;
    res->status = LIO_SEQ_FIRST_SCAN;
    return 0;
  }
  if (in->n_imu <= 0) {  // ImuProcess::Process: `if (meas.imu.empty()) return;` -> empty cloud -> skip
    res->status = LIO_SEQ_NO_IMU;
    return 0;
  }
  if ((int)s->poses.size() < in->n_imu + 2) s->poses.resize((size_t)in->n_imu + 2);
  int initialising = 0;
  s->n_poses = 0;
  const int rc = lio_imu_process(&s->imu, in->imu, in->n_imu, in->lidar_beg_time, in->lidar_end_time, &s->x, s->P,
                                 s->poses.data(), (int)s->poses.size(), &s->n_poses, &initialising);
  if (rc) return rc;
  if (initialising) {  // :722-725
    res->status = LIO_SEQ_IMU_INIT;
    return 0;
  }
  s->ekf_inited = !((in->lidar_beg_time - s->first_lidar_time) < kInitTime);  // :731-733
  const int rf = fov_segment(s);                                             // :736
  if (rf) return rf;
  return 1;
}

void adopt(lio_seq* s, const lio_scan_report& rep, lio_seq_result* res) {
  res->m = rep.m;
  res->n_valid = rep.n_valid;
  res->n_passes = rep.n_passes;
  memcpy(res->counts, rep.counts, sizeof(rep.counts));
  res->status = rep.status == LIO_SCAN_UPDATED     ? LIO_SEQ_UPDATED
                : rep.status == LIO_SCAN_MAP_BUILT ? LIO_SEQ_MAP_BUILT
                                                   : LIO_SEQ_FEW_POINTS;
  res->x = s->x;
  res->n_box_deleted = s->n_box_deleted;
}

}  // namespace

extern "C" {

void lio_seq_default_config(lio_seq_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->filter_size_surf = 0.5f;  // launch files
  cfg->filter_size_map = 0.5f;
  cfg->max_iteration = 3;        // mapping_velodyne.launch:10
  cfg->extrinsic_est = 0;
  cfg->extrinsic_R[0] = cfg->extrinsic_R[4] = cfg->extrinsic_R[8] = 1.0;
  cfg->gyr_cov = cfg->acc_cov = 0.1;
  cfg->b_gyr_cov = cfg->b_acc_cov = 0.0001;
  cfg->cube_len = 1000.0;        // cube_side_length
  cfg->det_range = 300.0;        // DET_RANGE (laserMapping.cpp:39)
  cfg->laser_point_cov = 0.001;  // LASER_POINT_COV (laserMapping.cpp:29)
}

int lio_seq_create(lio_ctx* ctx, const lio_seq_config* cfg, lio_seq** out) {
  if (!ctx || !cfg || !out || !(cfg->filter_size_surf > 0.f) || !(cfg->filter_size_map > 0.f) ||
      cfg->max_iteration < 0 || cfg->max_iteration > 32)
    return LIO_E_INVALID;
  lio_seq* s = new (std::nothrow) lio_seq();
  if (!s) return LIO_E_INVALID;
  s->ctx = ctx;
  s->cfg = *cfg;
  lio_imu_proc_init(&s->imu);
  const double g[3] = {cfg->gyr_cov, cfg->gyr_cov, cfg->gyr_cov}, a[3] = {cfg->acc_cov, cfg->acc_cov, cfg->acc_cov};
  const double bg[3] = {cfg->b_gyr_cov, cfg->b_gyr_cov, cfg->b_gyr_cov};
  const double ba[3] = {cfg->b_acc_cov, cfg->b_acc_cov, cfg->b_acc_cov};
  lio_imu_set_param(&s->imu, cfg->extrinsic_T, cfg->extrinsic_R, g, a, bg, ba);  // laserMapping.cpp:692-695
  memset(&s->x, 0, sizeof(s->x));  // state_ikfom defaults (use-ikfom.hpp:18-27)
  s->x.rot[0] = 1.0;
  s->x.offset_R_L_I[0] = 1.0;
  s->x.grav[2] = -9.81;
  memset(s->P, 0, sizeof(s->P));
  for (int i = 0; i < 24; ++i) s->P[i * 24 + i] = 1.0;
  *out = s;
  return LIO_OK;
}

void lio_seq_destroy(lio_seq* s) { delete s; }

int lio_set_host_threads(int n) {
  if (n < 0 || n > 64) return LIO_E_INVALID;
  g_host_threads.store(n, std::memory_order_relaxed);
  return LIO_OK;
}

int lio_seq_get_state(const lio_seq* s, lio_state* x, double P[576]) {
  if (!s) return LIO_E_INVALID;
  if (x) *x = s->x;
  if (P) memcpy(P, s->P, sizeof(s->P));
  return LIO_OK;
}

int lio_seq_set_state(lio_seq* s, const lio_state* x, const double P[576]) {
  if (!s) return LIO_E_INVALID;
  if (x) s->x = *x;
  if (P) memcpy(s->P, P, sizeof(s->P));
  return LIO_OK;
}

int lio_seq_local_map(const lio_seq* s, float box6[6], int64_t* n_box_deleted) {
  if (!s) return LIO_E_INVALID;
  if (box6) {
    memcpy(box6, s->lm_min, 12);
    memcpy(box6 + 3, s->lm_max, 12);
  }
  if (n_box_deleted) *n_box_deleted = s->n_box_deleted;
  return s->local_map_init ? LIO_OK : LIO_E_EMPTY_MAP;
}

int lio_seq_process(lio_seq* s, const lio_seq_input* in, lio_seq_result* res) {
  if (!s || !in || !res || in->n < 0 || in->n_imu < 0 || (in->n_imu > 0 && !in->imu)) return LIO_E_INVALID;
  // the records go up while the host propagates the IMU samples
  if (!s->first_scan && in->n_imu > 0) {
    const int rp = lio_scan_step_prefetch(s->ctx, in->lidar, in->n, in->stride_bytes);
    if (rp) return rp;
  }
  const int h = host_stage(s, in, res);
  if (h <= 0) {
    lio_scan_step_prefetch(s->ctx, nullptr, 0, 16);  // scan skipped: drop the staged copy
    return h;
  }
  lio_scan_report rep;
  const int rc = lio_scan_step(s->ctx, in->lidar, in->n, in->stride_bytes, s->poses.data(), s->n_poses, &s->x, s->P,
                               s->cfg.filter_size_surf, s->cfg.filter_size_map, s->cfg.laser_point_cov,
                               s->cfg.max_iteration, s->cfg.extrinsic_est, s->ekf_inited ? 1 : 0, &rep);
  if (rc) return rc;
  adopt(s, rep, res);
  return LIO_OK;
}

int lio_seq_process_many(lio_seq* const* seqs, int n_seq, const lio_seq_input* in, lio_seq_result* res) {
  if (!seqs || n_seq < 1 || n_seq > 64 || !in || !res) return LIO_E_INVALID;
  for (int k = 0; k < n_seq; ++k) {
    if (!seqs[k] || in[k].n < 0 || in[k].n_imu < 0 || (in[k].n_imu > 0 && !in[k].imu)) return LIO_E_INVALID;
    // one context per sequence (a context is not re-entrant, and the host stages below run concurrently), and one set
    // of filter settings for the shared update launch
    for (int j = 0; j < k; ++j)
      if (seqs[j] == seqs[k] || seqs[j]->ctx == seqs[k]->ctx) return LIO_E_INVALID;
    const lio_seq_config &a = seqs[0]->cfg, &b = seqs[k]->cfg;
    if (a.max_iteration != b.max_iteration || a.extrinsic_est != b.extrinsic_est ||
        a.laser_point_cov != b.laser_point_cov)
      return LIO_E_INVALID;
  }
  lio_seq* begun[64];
  lio_ctx* due[64];
  int idx_of[64], n_begun = 0, n_due = 0;
  int first_err = LIO_OK;
  // host stage (IMU propagation) + upload + preprocessing enqueue of every sequence: independent contexts and streams,
  // so the host side of the sequences runs on a few threads instead of queueing behind one
  int h_of[64], rc_of[64];
  int32_t d_of[64];
  const int nt = std::min(host_threads(), n_seq);
#pragma omp parallel for num_threads(nt) schedule(static, 1) if (nt > 1)
  for (int k = 0; k < n_seq; ++k) {
    lio_seq* s = seqs[k];
    d_of[k] = 0;
    rc_of[k] = LIO_OK;
    memset(&res[k], 0, sizeof(res[k]));
    if (!s->first_scan && in[k].n_imu > 0) {
      rc_of[k] = lio_scan_step_prefetch(s->ctx, in[k].lidar, in[k].n, in[k].stride_bytes);
      if (rc_of[k]) {
        h_of[k] = 1;  // reported through rc_of below
        continue;
      }
    }
    h_of[k] = host_stage(s, &in[k], &res[k]);
    if (h_of[k] <= 0) {
      lio_scan_step_prefetch(s->ctx, nullptr, 0, 16);
      continue;
    }
    rc_of[k] = lio_scan_step_begin(s->ctx, in[k].lidar, in[k].n, in[k].stride_bytes, s->poses.data(), s->n_poses,
                                   &s->x, s->P, s->cfg.filter_size_surf, &d_of[k]);
  }
  for (int k = 0; k < n_seq; ++k) {
    lio_seq* s = seqs[k];
    if (h_of[k] < 0 && first_err == LIO_OK) first_err = h_of[k];
    if (h_of[k] <= 0) continue;
    if (rc_of[k]) {
      if (first_err == LIO_OK) first_err = rc_of[k];
      continue;
    }
    idx_of[n_begun] = k;
    begun[n_begun++] = s;
    if (d_of[k]) due[n_due++] = s->ctx;
    s->phase = d_of[k] ? 1 : 0;
  }
  // the updates that are due: one cooperative launch per group of <= 8 (all sequences share the filter settings of
  // the first one, as BASELINE.json config 4 has them)
  if (n_due > 0) {
    const lio_seq_config& c0 = begun[0]->cfg;
    for (int a = 0; a < n_due; a += 8) {
      const int g = std::min(8, n_due - a);
      const int rc = g == 1 ? lio_update_enqueue(due[a], c0.laser_point_cov, c0.max_iteration, c0.extrinsic_est, 1)
                            : lio_update_enqueue_multi(due + a, g, c0.laser_point_cov, c0.max_iteration,
                                                       c0.extrinsic_est, 1);
      if (rc && first_err == LIO_OK) first_err = rc;
    }
  }
#pragma omp parallel for num_threads(nt) schedule(static, 1) if (nt > 1)
  for (int j = 0; j < n_begun; ++j) {
    lio_seq* s = begun[j];
    rc_of[j] = s->phase == 1 ? lio_scan_step_end(s->ctx, s->cfg.filter_size_map, s->ekf_inited ? 1 : 0) : LIO_OK;
  }
  for (int j = 0; j < n_begun; ++j)
    if (rc_of[j] && first_err == LIO_OK) first_err = rc_of[j];
  for (int j = 0; j < n_begun; ++j) {
    lio_seq* s = begun[j];
    s->phase = 0;
    lio_scan_report rep;
    const int rc = lio_scan_step_finish(s->ctx, &s->x, s->P, &rep);
    if (rc) {
      if (first_err == LIO_OK) first_err = rc;
      continue;
    }
    adopt(s, rep, &res[idx_of[j]]);
  }
  return first_err;
}

}  // extern "C"
