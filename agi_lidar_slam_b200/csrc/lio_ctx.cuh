// Context layout shared by the translation units of liblio_b200.so.
#pragma once
#include <math.h>

#include <string>

#include "lio_common.cuh"

struct lio_ctx {
  int device = 0;
  int sm_count = 148;
  lio_caps caps{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t capture_stream = nullptr;  // graphs are captured here, launched on `stream`
  std::string err;
  int64_t launches = 0;

  // ---- map
  lio::MapView map{};
  uint32_t hash_cap = 0;
  int32_t next_id = 0;     // id given to the next inserted point (host mirror)
  bool map_built = false;  // ≙ ikdtree.Root_Node != nullptr
  int knn_rings = 3;
  bool batched_finish = true;   // LIO_FINISH_BATCHED=0: finish every search tile right behind its search
  bool interleave = false;      // LIO_INTERLEAVE=1: scan points dealt to the update's blocks in runs of 8 (load balance)
  bool stage_search = false;    // LIO_STAGE_SEARCH=1: the update's searches stage their cells in shared memory
  float map_downsample = 0.5f;  // ≙ KD_TREE::downsample_size (set_downsample_param, laserMapping.cpp:748)
  // batch scratch (sized for max(max_down_points, build chunk))
  int64_t batch_cap = 0;
  float4* d_batch_pts = nullptr;    // staged points of the current insert batch
  uint32_t* d_batch_slot = nullptr; // hash slot per point
  uint32_t* d_batch_rank = nullptr; // rank within its cell for this batch
  unsigned long long* d_vox_best = nullptr;  // downsample-add: per batch-voxel best (dist bits << 32 | ~idx)
  unsigned long long* d_vox_key = nullptr;   //                  per batch-voxel key (bmin bit patterns hashed)
  uint32_t vox_cap = 0;
  uint8_t* d_batch_flag = nullptr;  // 1 = append this point

  // ---- current scan (feats_down_body) and per-point persistent state
  int64_t scan_m = 0;               // host mirror (-1: unknown, device-resident only)
  int* d_scan_m = nullptr;          // device copy read by the kernels
  float4* d_body = nullptr;         // M x (x,y,z,intensity)
  float4* d_world = nullptr;        // M x FP32 p_world of the last pass
  float4* d_near = nullptr;         // M x 5 cached Nearest_Points (x,y,z,id bits)
  float* d_near_d2 = nullptr;       // M x 5
  int* d_near_cnt = nullptr;        // M: length of the row's known prefix of the unbounded neighbour list
  float4* d_near_q = nullptr;       // M: FP32 p_world the row was searched at
  uint8_t* d_selected = nullptr;    // M point_selected_surf
  float4* d_normvec = nullptr;      // M x (a,b,c,pd2)
  float4* d_plane = nullptr;        // M x pabcd fitted by the last search pass
  unsigned long long* d_partials = nullptr;  // pass_grid x 2 LIO_BLOB stamped words: per-block partial sums of one pass
  int pass_grid = 0;                // blocks of the persistent update grid (all co-resident)
  double* d_blob = nullptr;         // LIO_BLOB (own buffer, or the one given to lio_blob_bind)
  double* d_blob_own = nullptr;
  double* d_prior = nullptr;        // 288: P11^-1 and P21 P11^-1 of the current update
  unsigned* d_sync = nullptr;       // grid barrier words {arrivals, release}
  // sharded map: peer mailboxes (own one allocated here, the peers' mapped through cudaIpc)
  void* d_mailbox = nullptr;        // {stamped words [8 ranks][2 slots][LIO_BLOB][2]; int err}
  int peer_world = 0, peer_rank = 0;
  unsigned peer_epoch = 0;          // advanced by sharded launches only: stays in lockstep across the ranks
  // striped ownership of the sharded calls (lio_set_shard_stripes); stripe_world == 0: the calls' x windows
  int stripe_world = 0, stripe_rank = 0;
  float stripe_origin = 0.f, stripe_width = 0.f;
  unsigned long long* peer_mbox[8] = {nullptr};
  void* peer_base[8] = {nullptr};   // cudaIpcOpenMemHandle results (closed in lio_destroy)
  int* d_peer_err = nullptr;
  unsigned epoch = 0;               // stamp base of the next update_kernel launch
  unsigned long long* d_pub = nullptr;  // 29 stamped words published by the solving block after every Kalman step
  long long* d_dbg = nullptr;       // in-kernel timeline (only with LIO_TIMELINE=1)
  uint8_t* d_cls = nullptr;         // map_incremental class per point
  float4* d_add_a = nullptr;        // compacted PointToAdd, PointNoNeedDownsample right behind it

  // ---- filter state
  double* d_state_blk = nullptr;    // one allocation: {x 26, P 576, ctrl 4} {x0 26, P0 576} {xprop 26} {dx 24}
  lio::StateD* d_x = nullptr;       // current state
  lio::StateD* d_xprop = nullptr;   // x_propagated
  double* d_P = nullptr;            // 24x24
  lio::StateD* d_x0 = nullptr;      // prior snapshot (lio_state_upload)
  double* d_P0 = nullptr;
  lio::Ctrl* d_ctrl = nullptr;
  double* d_dx = nullptr;           // last dx (24)
  void* h_pinned = nullptr;         // pinned staging: [0,606) download area, [640, 1242) upload area, [1280, ..) blob
  size_t h_pinned_bytes = 0;
  cudaEvent_t upload_done = nullptr;  // guards reuse of the upload staging area
  cudaEvent_t multi_evt = nullptr;    // orders this context's stream with a multi-sequence launch on another stream
  double* h_out = nullptr;            // mapped pinned: posterior {x, P, ctrl} + sequence word (host-direct path)
  double* h_out_dev = nullptr;        // its device address
  unsigned long long host_seq = 0;
  bool no_zero_copy = true;

  // ---- preprocess
  lio::VoxelFilter vf{};            // working set of the surf voxel filter (lio_preprocess.cu)
  unsigned char* d_cloud = nullptr; // PointCloud2 bytes of the scan being decoded (allocated on first use)
  size_t cloud_bytes = 0;
  int64_t n_decoded = 0;
  float4* d_raw = nullptr;          // N raw points (x,y,z,t_ms)
  float* d_raw_aux = nullptr;       // N intensity (stride-48 input)
  float4* d_undist = nullptr;       // N undistorted
  int* d_vkeys = nullptr;           // N x 3
  lio_pose6d* d_poses = nullptr;    // up to 256
  uint32_t* d_sort_keys_in = nullptr;  // ring keys of the sensor decoders (the scan's voxel filter sorts nothing)
  uint32_t* d_sort_keys_out = nullptr;
  uint32_t* d_sort_vals_in = nullptr;
  uint32_t* d_sort_vals_out = nullptr;
  void* d_cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  int* d_prep_counters = nullptr;   // working set (reset by centroid_kernel): [0] M, [1..6] key min/max, [7] error; [16..23] = the same
                                    // as filed for the host at the end of the preprocessing; [8..9] map_incremental class counts, [11] their sum,
                                    // [10] voxel runs, [12] decoded points

  // lio_scan_step: one main-loop iteration enqueued without intermediate host synchronisation
  int min_m = 0;               // update / map growth treat a scan with fewer points as empty (5 inside a step)
  int64_t scan_m_bound = 0;    // upper bound of M while only the device knows it (launch bounds)
  int step_phase = 0;          // 0 idle, 1 begun (update due), 2 map growth + report enqueued, 3 first-scan branch done
  int step_status = 0;         // LIO_SCAN_* of the first-scan branch
  int64_t step_m = 0;
  // deferred map growth (lio_set_deferred_growth): lio_scan_step_finish returns once the posterior is on the host;
  // the growth counts of that scan are collected by settle_growth() before the next host use of the map
  bool deferred_growth = false;
  bool growth_pending = false;
  cudaEvent_t ev_post = nullptr;    // posterior + preprocess counters are in h_pinned
  cudaEvent_t ev_growth = nullptr;  // map growth done, its counts are in h_pinned[620..626)
  cudaStream_t prep_stream = nullptr;  // the next scan's upload + undistort + sort run here next to the pending growth
  cudaEvent_t ev_prep = nullptr;       // preprocessing on prep_stream done
  cudaEvent_t centroid_wait = nullptr; // preprocess(): make the kernel that writes d_body / d_scan_m wait for this
  const void* staged_ptr = nullptr;    // lio_scan_step_prefetch: this scan's records are already in d_raw / d_raw_aux
  int64_t staged_n = 0;
  int staged_stride = 0;
  bool staged_on_prep = false;         // ... copied on prep_stream
  int32_t last_counts[3] = {0, 0, 0};
  int growth_rc = 0;                // error of a deferred growth, reported by the next call that settles it
};

#define LIO_CHECK(ctx, call)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
      return LIO_E_CUDA;                                                                       \
    }                                                                                          \
  } while (0)

namespace lio {
// what lio_update_scan_host hands to the launcher (host-direct path: no copies around the kernel)
struct HostDirect {
  const double* x0;              // prior, host memory
  const double* P0;
  int m;                         // scan size
  const float4* body_src;        // device-accessible address of the caller's pinned scan, or nullptr (scan already in d_body)
  double* out_dev;               // device address of the mapped pinned result buffer
  unsigned long long* flag_dev;  // ... of its sequence word
  unsigned long long seq;
};
// launchers implemented in lio_pass.cu / lio_map.cu / lio_preprocess.cu
int ensure_tables(lio_ctx* c);
int pass_grid_blocks(lio_ctx* c);
int launch_update(lio_ctx* c, double R, int max_iter, int extrinsic_est, int from_snapshot, float own_min = -INFINITY,
                  float own_max = INFINITY, bool sharded = false, const HostDirect* hd = nullptr);
int launch_update_multi(lio_ctx* const* cs, int n, double R, int max_iter, int extrinsic_est, int from_snapshot);
int launch_pass(lio_ctx* c, int mode, int extrinsic_est, float own_min, float own_max, bool sharded_call = false);
int launch_solve(lio_ctx* c, double R, int extrinsic_est);
int launch_begin(lio_ctx* c, int max_iter, int extrinsic_est, int from_snapshot);
int launch_knn_batch(lio_ctx* c, const float4* d_q, int64_t m);
int launch_far_complete(lio_ctx* c, const float4* d_q, int64_t m, int min_m, int64_t bound, int need);

int map_reset(lio_ctx* c);
int map_append_batch(lio_ctx* c, const float4* d_pts, int64_t n, int32_t id_base, const uint8_t* d_flag);
int map_add_downsample(lio_ctx* c, const float4* d_pts, int64_t n, int32_t id_base, int32_t* n_added);
int map_delete_boxes(lio_ctx* c, const float* h_boxes6, int nb, int32_t* n_deleted);
int map_dump(lio_ctx* c, float* xyz, int32_t* ids, int64_t cap, int64_t* n);
int map_removed_points(lio_ctx* c, float* xyz, int64_t cap, int64_t* n);
int map_incremental(lio_ctx* c, const lio_state* x, float fsm, int ekf_inited, int32_t counts[3]);
int map_incremental_enqueue(lio_ctx* c, float fsm, int ekf_inited, int min_m, int64_t bound);
int map_build_scan(lio_ctx* c, const lio_state* x);
int check_id_space(lio_ctx* c, int64_t n_new);  // lio_api.cu: LIO_E_CAPACITY when n_new more ids would wrap int32
int settle_growth(lio_ctx* c);  // lio_api.cu: waits for a deferred map growth and books its counts (no-op otherwise)

int preprocess(lio_ctx* c, int64_t n, int n_poses, const lio_state* end_state, float leaf, bool has_aux);
int decode_cloud2(lio_ctx* c, int64_t n, const lio_cloud_layout& L, bool yaw_times, int64_t* n_out);
}  // namespace lio
