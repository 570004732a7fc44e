/*
 * lio_b200.h — C-ABI of the B200-native S-FAST_LIO hot path (liblio_b200.so).
 *
 * The reference (zhan994/agi_lidar_slam, src/S-FAST_LIO) has NO plugin/FFI boundary for this path:
 * everything is C++ templates compiled into one ROS node.  The boundary kept here is
 * the reference's C++ CALL SURFACE; each entry point below names the reference call it stands under
 * (paths relative to src/S-FAST_LIO/).  include/lio_facade.hpp re-creates that C++ surface
 * (KD_TREE<PointType>, esekfom::esekf, ImuProcess, VoxelGrid) on top of these functions, and
 * INTEGRATION.md shows the lines a maintainer changes in laserMapping.cpp.
 *
 * Conventions: extern "C"; plain pointers and sizes; int return = 0 (LIO_OK) or a negative LIO_E_*;
 * no exceptions cross the boundary; caller owns every host buffer; the library owns all device memory;
 * one context per GPU; a context is NOT re-entrant (the reference drives the path from one thread).
 * There is no CPU fallback: lio_create fails with LIO_E_NO_DEVICE when no sm_100 device is usable.
 *
 * Point buffers are arrays of records `stride_bytes` apart whose first three floats are x,y,z:
 *   stride 16 : {x,y,z,w}           w = per-point time [ms] for raw scans, intensity otherwise
 *   stride 48 : pcl::PointXYZINormal {x,y,z,_}{nx,ny,nz,_}{intensity,curvature,_,_}  (common_lib.h:26);
 *               curvature = per-point time offset in ms (preprocess.cpp:165-168)
 */
#ifndef LIO_B200_H
#define LIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIO_ABI_VERSION 2
#define LIO_NUM_MATCH_POINTS 5 /* common_lib.h:18 */

enum {
  LIO_OK = 0,
  LIO_E_INVALID = -1,     /* bad argument */
  LIO_E_NO_DEVICE = -2,   /* no usable CUDA device (product never falls back to CPU) */
  LIO_E_CUDA = -3,        /* CUDA runtime error; see lio_last_error() */
  LIO_E_CAPACITY = -4,    /* scan / map / hash capacity given in lio_caps exceeded */
  LIO_E_EMPTY_MAP = -5,   /* operation needs a built map (ikdtree.Root_Node == nullptr) */
  LIO_E_VOXEL_RANGE = -6  /* pcl::VoxelGrid "leaf size too small" overflow: output == input */
};

/* state_ikfom (include/use-ikfom.hpp:18-27).  Rotations are unit quaternions (w,x,y,z), as the old
 * Sophus::SO3 the reference links stores them.  Error-state order (24): pos, rot, R_LI, t_LI, vel, bg, ba, grav. */
typedef struct lio_state {
  double pos[3];
  double rot[4];
  double offset_R_L_I[4];
  double offset_T_L_I[3];
  double vel[3];
  double bg[3];
  double ba[3];
  double grav[3];
} lio_state; /* 26 doubles */

/* sfast_lio::Pose6D (msg/Pose6D.msg), element of ImuProcess::IMUpose (src/IMU_Processing.hpp:127). */
typedef struct lio_pose6d {
  double offset_time;
  double acc[3];
  double gyr[3];
  double vel[3];
  double pos[3];
  double rot[9]; /* row-major */
} lio_pose6d; /* 22 doubles */

/* One IMU sample (sensor_msgs::Imu fields the path reads): stamp [s], linear_acceleration, angular_velocity. */
typedef struct lio_imu_sample {
  double stamp;
  double acc[3];
  double gyr[3];
} lio_imu_sample; /* 7 doubles */

typedef struct lio_caps {
  int64_t max_scan_points; /* N cap of a raw scan                                   (default 262144)  */
  int64_t max_down_points; /* M cap after the surf voxel filter; reference: 100000  (esekfom.hpp:23-29) */
  int64_t max_map_points;  /* live + garbage slots of the map point pool            (default 4194304) */
  float map_cell;          /* edge of a kNN hash cell [m]; does not change results  (default 1.5)     */
  float knn_max_d2;        /* search bound on squared distance; the path needs 5    (esekfom.hpp:147) */
  float plane_thr;         /* esti_plane inlier threshold                           (esekfom.hpp:157: 0.1f) */
  float map_downsample;    /* ikdtree.set_downsample_param(filter_size_map_min)     (laserMapping.cpp:748: 0.5) */
} lio_caps;

typedef struct lio_ctx lio_ctx;

/* ---- context ---------------------------------------------------------------------------------- */
int lio_abi_version(void);
void lio_default_caps(lio_caps* caps);
/* Creates a context on CUDA device `device`; *out = NULL on failure. */
int lio_create(int device, const lio_caps* caps, lio_ctx** out);
void lio_destroy(lio_ctx* ctx);
/* Work is enqueued on this cudaStream_t (pass torch.cuda.current_stream().cuda_stream); default: own stream. */
int lio_set_stream(lio_ctx* ctx, void* cuda_stream);
int lio_synchronize(lio_ctx* ctx);
const char* lio_last_error(lio_ctx* ctx);
/* Kernels launched by this context since creation (bench.py's "gpu_launches"). */
int64_t lio_launch_count(lio_ctx* ctx);

/* ---- map: KD_TREE<PointType> surface (include/ikd-Tree/ikd_Tree.h:264-299) ----------------------- */
/* ≙ KD_TREE::Build (ikd_Tree.cpp:355-367; laserMapping.cpp:756).  Replaces any existing map. ids = 0..n-1. */
int lio_map_build(lio_ctx* ctx, const void* pts, int64_t n, int stride_bytes);
/* ≙ KD_TREE::Add_Points (ikd_Tree.cpp:419-512; laserMapping.cpp:430-431).  With downsample_on the per-voxel
 * "keep the point nearest the voxel centre, new point wins ties" rule is applied per batch.
 * *n_added = new points that ended up in the map (the reference's return value instead counts the insert operations
 * of its sequential loop, re-insertions of surviving old points included; it is only printed).  New point i gets
 * id = next_id + i. */
int lio_map_add(lio_ctx* ctx, const void* pts, int64_t n, int stride_bytes, int downsample_on, int32_t* n_added);
/* ≙ KD_TREE::set_downsample_param (ikd_Tree.h:274; laserMapping.cpp:748): the voxel edge lio_map_add downsamples with
 * (initially caps.map_downsample).  Any size > 0; it need not divide the kNN cell.  lio_map_incremental / lio_scan_step
 * downsample with their own filter_size_map argument, the one value the reference uses for both (laserMapping.cpp:748). */
int lio_map_set_downsample(lio_ctx* ctx, float downsample_size);
/* ≙ KD_TREE::Delete_Point_Boxes (ikd_Tree.cpp:559-579; laserMapping.cpp:361-364). boxes6 = nb x {min xyz, max xyz}, half-open. */
int lio_map_delete_boxes(lio_ctx* ctx, const float* boxes6, int nb, int32_t* n_deleted);
/* ≙ KD_TREE::size() / validnum() (ikd_Tree.cpp:66-128): total = slots ever used, valid = live points. */
int lio_map_size(lio_ctx* ctx, int64_t* total, int64_t* valid);
/* ≙ KD_TREE::flatten(Root_Node, PCL_Storage, NOT_RECORD) (ikd_Tree.cpp:1490-1516): live points, ascending id.
 * xyz (cap x 3) and ids (cap) may be NULL; *n = live count. */
int lio_map_dump(lio_ctx* ctx, float* xyz, int32_t* ids, int64_t cap, int64_t* n);
/* ≙ KD_TREE::acquire_removed_points (ikd_Tree.h:294, ikd_Tree.cpp:582-594; laserMapping.cpp:361-362): the points that
 * left the map since the last call that took them -- box deletes (Delete_Point_Boxes) and the points a downsampling
 * Add_Points replaced -- x,y,z each, order unspecified.  *n = points logged; with xyz != NULL the first min(*n, cap) are
 * copied and the log is cleared, with xyz == NULL only the count is returned.  (The reference files a deleted point
 * when the subtree holding it is next rebuilt; here it is filed when it is deleted.)  The log holds
 * min(max_map_points, 2^22) points: LIO_E_CAPACITY if more were deleted between two calls. */
int lio_map_removed_points(lio_ctx* ctx, float* xyz, int64_t cap, int64_t* n);
/* ≙ a batch of KD_TREE::Nearest_Search(point, 5, Nearest_Points, Point_Distance, max_dist) (ikd_Tree.cpp:370-402;
 * esekfom.hpp:140).  max_d2 = max_dist * max_dist, as the tree squares it (ikd_Tree.cpp:965) and tests `dist <= max_d2`
 * (:980); INFINITY is the reference's default (ikd_Tree.h:285) and what esekfom.hpp:140-141 gets.  Up to
 * caps.knn_max_d2 the call is one launch of the hot group search; beyond it the rows that search leaves short are
 * completed by the unbounded shell search (exact; slower).  q_xyz: m x 3 world-frame points.  Outputs (each may be
 * NULL), per query ascending by (d2 FP32, id): idx5 m x 5 point ids (-1 pad), d2_5 m x 5 (+inf pad), nbr_xyz m x 5 x 3. */
int lio_knn5(lio_ctx* ctx, const float* q_xyz, int64_t m, float max_d2, int32_t* idx5, float* d2_5, float* nbr_xyz);

/* Instrumentation: re-runs the search kernel on the (at most max_down_points) queries the last lio_knn5 call left on
 * the device; no copies, no synchronisation (tools/knn_roofline.py brackets it with CUDA events). */
int lio_knn5_resident(lio_ctx* ctx, int64_t m);

/* ---- scan preprocessing: ImuProcess::UndistortPcl back half + pcl::VoxelGrid ------------------------ */
/* ≙ IMU_Processing.hpp:361-401 followed by laserMapping.cpp:737-738 (leaf from :683), fused in one pass.
 * raw_pts: n records (time in w / curvature, ms).  imu_poses/n_poses = ImuProcess::IMUpose; end_state = state
 * after forward propagation (imu_state at :355).  n_poses < 2 => no motion compensation (voxel filter only).
 * Outputs: out_pts (max_down_points records of `stride_bytes`; centroids ascending by (kz,ky,kx)), *m = count;
 * optional undistorted (n x 4 floats, INPUT order) and voxel_key_xyz (n x 3 absolute voxel indices).
 * The downsampled cloud stays resident on the device as the current scan (feats_down_body). */
int lio_scan_preprocess(lio_ctx* ctx, const void* raw_pts, int64_t n, int stride_bytes, const lio_pose6d* imu_poses,
                        int n_poses, const lio_state* end_state, float leaf, void* out_pts, int64_t* m,
                        float* undistorted, int32_t* voxel_key_xyz);
/* Same work without host output copies (device-resident result only); *m may be NULL (no sync). */
int lio_scan_preprocess_resident(lio_ctx* ctx, const void* raw_pts, int64_t n, int stride_bytes,
                                 const lio_pose6d* imu_poses, int n_poses, const lio_state* end_state, float leaf,
                                 int64_t* m);
/* Sets the current scan (feats_down_body, laserMapping.cpp:738) from host memory: m records. */
int lio_scan_upload(lio_ctx* ctx, const void* down_pts, int64_t m, int stride_bytes);

/* ---- sensor decoding (next row of SURVEY.md §8f): sensor_msgs::PointCloud2 bytes -> the path's input cloud ------ */
/* Where the fields of one point record sit in a PointCloud2 `data` blob, and the decimation / blind-zone rule of the
 * handler that would have read it (src/preprocess.cpp, feature extraction off as in every launch file). */
typedef struct lio_cloud_layout {
  int32_t point_step;       /* bytes per point record                                                  */
  int32_t off_x, off_y, off_z;        /* float32 fields                                                */
  int32_t off_intensity;    /* intensity field (see intensity_type), or -1                             */
  int32_t off_time;         /* per-point time field, or -1 (all points at t = 0)                       */
  int32_t time_type;        /* 0 = float32 (velodyne_ros::Point::time), 1 = uint32 (ouster_ros::Point::t,
                               livox offset_time), 2 = float64, 3 = float64 relative to record 0, in seconds
                               (rslidar_ros::Point::timestamp: (t - t0) * 1000.0, :880-882)            */
  int32_t point_filter_num; /* keep every point_filter_num-th record (i % n == 0; rule 3: every n-th VALID record) */
  int32_t rule;             /* 1 = oust64_handler (:243-268): drop when range^2 < blind^2;
                               2 = velodyne_handler (:380-428): keep when range^2 > blind^2
                               3 = avia_handler (:160-183) on livox CustomPoint records: line / tag test, decimation
                                   over the valid records, "differs from the previous record" test, time =
                                   offset_time / float(1000000) (time_scale unused)
                               4 = rs_handler (:872-921): as rule 2
                               (range^2 = FP32 x*x + y*y + z*z, compared in FP64 with the double blind) */
  float time_scale;         /* time_unit_scale: field units -> ms (preprocess.cpp:55-66)               */
  double blind;             /* blind radius [m] (preprocess.h:161)                                     */
  /* zero-initialise what a sensor does not have */
  int32_t off_ring;         /* ring (rules 2, 4 with yaw_time) / line (rule 3) field, or -1            */
  int32_t ring_type;        /* 0 = uint16, 1 = uint8                                                   */
  int32_t off_tag;          /* rule 3: livox tag (uint8)                                               */
  int32_t intensity_type;   /* 0 = float32, 1 = uint8 (livox reflectivity)                             */
  int32_t n_scans;          /* N_SCANS: rule 3 keeps line < n_scans; yaw_time: rings must be < n_scans */
  int32_t scan_rate;        /* SCAN_RATE [Hz] (omega_l = 0.361 * SCAN_RATE deg/ms)                     */
  int32_t yaw_time;         /* rules 2, 4: 1 = do as the handlers do when the LAST record's time field is not > 0
                               (given_offset_time == false, :296-310): per ring, the first record fixes yaw_fp and is
                               dropped, later ones get (yaw_fp - yaw [+ 360]) / omega_l, plus a revolution when that
                               falls behind the ring's previous time; 0 = always read the time field  */
  int32_t reserved;
} lio_cloud_layout;
/* ≙ Preprocess::process (preprocess.cpp:48-86) for a PointCloud2 lidar followed by lio_scan_preprocess_resident: the
 * raw message bytes go up once, are decoded, decimated and blind-filtered on the device (order preserved) and run
 * through undistortion + voxel filter.  *n_decoded = points that survived decoding, *m = feats_down_size. */
int lio_scan_preprocess_cloud2(lio_ctx* ctx, const void* data, int64_t n_records, const lio_cloud_layout* layout,
                               const lio_pose6d* imu_poses, int n_poses, const lio_state* end_state, float leaf,
                               int64_t* n_decoded, int64_t* m);
/* The decoded cloud of the last lio_scan_preprocess_cloud2 call: n x {x,y,z,t_ms} and n intensities (either may be NULL). */
int lio_scan_decoded(lio_ctx* ctx, float* xyzt, float* intensity, int64_t cap, int64_t* n);

/* ---- IESKF update: esekfom::esekf surface (include/esekfom.hpp) --------------------------------------- */
/* ≙ one esekf::h_share_model call (esekfom.hpp:106-227) + the H^T H / H^T h products of :306-319 for the
 * current scan at state x.  do_search ≙ dyn_share.converge.  blob90 = upper triangle of H^T H (12x12, row-major,
 * 78) followed by H^T h (12); *n_valid = effct_feat_num. */
int lio_update_pass(lio_ctx* ctx, const lio_state* x, int do_search, int extrinsic_est, double blob90[90],
                    int32_t* n_valid);
/* ≙ esekf::update_iterated_dyn_share_modified(R, feats_down_body, ikdtree, Nearest_Points, maximum_iter,
 * extrinsic_est) (esekfom.hpp:270-346), whole loop on the device.  x_io / P_io (24x24 row-major) are the
 * propagated prior in, the posterior out.  *n_valid_last = effct_feat_num of the last pass, *n_passes = number of
 * h_share_model calls made. */
int lio_update_scan(lio_ctx* ctx, lio_state* x_io, double P_io[576], double R, int max_iter, int extrinsic_est,
                    int32_t* n_valid_last, int32_t* n_passes);
/* The same for a downsampled cloud in HOST memory (m records of stride 16: x,y,z,intensity): one copy for the scan, the
 * prior in the kernel parameters, one kernel, the posterior written by the kernel into mapped pinned memory (the call
 * spins on a sequence word: no copy and no stream synchronisation after the kernel).  This is the per-scan call of a
 * host that runs its own voxel filter. */
int lio_update_scan_host(lio_ctx* ctx, const void* down_pts, int64_t m, int stride_bytes, lio_state* x_io,
                         double P_io[576], double R, int max_iter, int extrinsic_est, int32_t* n_valid_last,
                         int32_t* n_passes);
/* Device-resident pieces of the above (used by bench.py and by the multi-GPU sharded-map driver):        */
int lio_state_upload(lio_ctx* ctx, const lio_state* x, const double P[576]); /* also stored as the prior snapshot */
int lio_state_download(lio_ctx* ctx, lio_state* x, double P[576], int32_t* n_valid_last, int32_t* n_passes);
/* Enqueue the whole update (one persistent cooperative kernel) on the context stream; no host sync.  from_snapshot != 0 first restores the state
 * uploaded by lio_state_upload (so a benchmark can repeat the same update). */
int lio_update_enqueue(lio_ctx* ctx, double R, int max_iter, int extrinsic_est, int from_snapshot);
/* One iteration of the reference main loop for one scan (src/laserMapping.cpp:737-785): VoxelGrid(UndistortPcl back half),
 * `feats_down_size < 5 -> continue` (:741-744), first-scan Build (:747-758), update_iterated_dyn_share_modified (:772-774),
 * map_incremental (:785) -- enqueued back to back on the context's stream.  M, the map_incremental class counts and the
 * insert sizes stay on the device; the host synchronises ONCE, at the end, and gets everything in the report.
 * x_io / P_io: the propagated state at scan end (what lio_imu_process returned) in, the posterior out (untouched when
 * the scan is skipped).  raw_pts / poses as for lio_scan_preprocess.  leaf_map = filter_size_map_min; leaf_map == 0 keeps
 * the map static -- the relocalisation loop, where map_incremental() is commented out (src/laserMapping_re.cpp:676). */
typedef struct lio_scan_report {
  int64_t m;          /* feats_down_size */
  int32_t status;     /* LIO_SCAN_* */
  int32_t n_valid;    /* effct_feat_num of the last pass */
  int32_t n_passes;
  int32_t counts[3];  /* map_incremental: PointToAdd, PointNoNeedDownsample, points Add_Points(downsample) inserted */
} lio_scan_report;
enum { LIO_SCAN_UPDATED = 0, LIO_SCAN_FEW_POINTS = 1, LIO_SCAN_MAP_BUILT = 2 };
int lio_scan_step(lio_ctx*, const void* raw_pts, int64_t n, int stride_bytes, const lio_pose6d* imu_poses, int n_poses,
                  lio_state* x_io, double P_io[576], float leaf_surf, float leaf_map, double R, int max_iter,
                  int extrinsic_est, int ekf_inited, lio_scan_report* report);
/* The same in three enqueue-only pieces plus the synchronising one, for hosts that step several sequences together:
 *   begin (preprocessing + prior; *update_due = 0 on the first-scan branch, which is complete after begin)
 *   -> lio_update_enqueue / lio_update_enqueue_multi with from_snapshot = 1 for the contexts that are due
 *   -> end (map growth from the posterior + report copies)  -> finish (synchronise, bookkeeping, report). */
int lio_scan_step_begin(lio_ctx*, const void* raw_pts, int64_t n, int stride_bytes, const lio_pose6d* imu_poses,
                        int n_poses, const lio_state* x, const double P[576], float leaf_surf, int32_t* update_due);
int lio_scan_step_end(lio_ctx*, float leaf_map, int ekf_inited);
/* Optional, before the host's IMU propagation of the same scan: starts the upload of the raw records (they do not depend
 * on the filter state).  A following lio_scan_step / lio_scan_step_begin with the SAME pointer, n and stride uses the
 * staged copy instead of copying again; any other call, or n == 0, drops it.  The buffer must stay untouched in between. */
int lio_scan_step_prefetch(lio_ctx*, const void* raw_pts, int64_t n, int stride_bytes);
int lio_scan_step_finish(lio_ctx*, lio_state* x_out, double P_out[576], lio_scan_report* report);
/* Deferred map growth.  The reference publishes a scan's odometry BEFORE map_incremental runs (laserMapping.cpp:776-785).
 * With on != 0 the step does the same: lio_scan_step / lio_scan_step_finish (and lio_seq_process*) return as soon as the
 * posterior is on the host, while the map growth of that scan still runs on the context's stream underneath the host's
 * IMU propagation and the upload of the next scan (stream order keeps the next update behind it).  report.counts is
 * then {-1,-1,-1}; lio_scan_step_settle waits for the growth and returns its counts (and a capacity error, if any).
 * Every other entry point that needs the host's view of the map settles implicitly.  Results are unchanged. */
int lio_set_deferred_growth(lio_ctx*, int on);
int lio_scan_step_settle(lio_ctx*, int32_t counts[3]);

/* n <= 8 INDEPENDENT updates -- different sequences, each context with its own map, scan and filter state (BASELINE.json
 * config 4) -- in ONE cooperative launch on ctxs[0]'s stream: the persistent grid is cut into n slices, each with its own
 * workers and solver block.  A single update is latency-bound and leaves most of the GPU idle; slices fill it.  All
 * contexts must live on the same device; states go in with lio_state_upload and come out with lio_state_download per
 * context (stream order between the contexts' streams and the launch is taken care of). */
int lio_update_enqueue_multi(lio_ctx* const* ctxs, int n, double R, int max_iter, int extrinsic_est, int from_snapshot);
/* Sharded-map driver: begin, then per pass {pass_enqueue -> all-reduce 92 doubles at
 * lio_blob_device_ptr -> step_enqueue}.  x_own_min/max say which rows this rank owns: those whose p_world.x AT THEIR
 * LAST SEARCH PASS lies in [x_own_min, x_own_max) -- the same FP32 bits on every rank, fixed until the next search pass,
 * so a row and its cached neighbours stay with one rank (use -inf/+inf for a single GPU; lio_set_shard_stripes replaces
 * the window by stripes).  A rank searches only the rows it owns; after a sharded update the per-point outputs
 * (lio_get_neighbors) are meaningful for owned rows only.  begin resets the loop state (x_propagated = x,
 * converge = true, ...) and forms the per-update constants of the Kalman step. */
int lio_update_begin(lio_ctx* ctx, int max_iter, int extrinsic_est, int from_snapshot);
int lio_update_pass_enqueue(lio_ctx* ctx, int extrinsic_est, float x_own_min, float x_own_max);
int lio_update_step_enqueue(lio_ctx* ctx, double R, int extrinsic_est);
/* The same exchange INSIDE the persistent kernel, over NVLink peer memory instead of NCCL + launches: every rank's
 * solver block stores its blob as self-validating stamped words into a mailbox of every peer (mapped with cudaIpc; no
 * fence, no flag), polls its own mailbox until every rank's words carry the pass's stamp and adds
 * the blobs in rank order; the whole sharded update is one launch per rank.  Setup once: lio_peer_handle on every rank,
 * all-gather the 64-byte handles, lio_peer_connect.  All ranks must then enqueue the same sequence of sharded updates.
 * lio_peer_status reports whether a peer failed to show up (the wait is bounded; the update is then garbage). */
int lio_peer_handle(lio_ctx* ctx, unsigned char handle[64]);
int lio_peer_connect(lio_ctx* ctx, int rank, int world, const unsigned char* handles /* world x 64 */);
int lio_update_enqueue_sharded(lio_ctx* ctx, double R, int max_iter, int extrinsic_est, int from_snapshot,
                               float x_own_min, float x_own_max);
int lio_peer_status(lio_ctx* ctx, int32_t* timed_out);
/* Striped ownership for the two sharded calls above (lio_update_pass_enqueue, lio_update_enqueue_sharded): instead of
 * one x window per rank, rank r owns the rows whose search position falls into the stripes
 * s = floor((p_world.x - x_origin) / stripe_width) with s mod world == r, and holds the map points of those stripes plus
 * the halo.  A scan is a few hundred metres wide, so with stripes of a few tens of metres every rank gets ~1/world of
 * its rows wherever the robot is (one window per rank leaves all of them with the one or two ranks around it).  The
 * calls' x_own_min / x_own_max are then ignored.  world = 0 switches back to windows.  No reference counterpart. */
int lio_set_shard_stripes(lio_ctx* ctx, float x_origin, float stripe_width, int world, int rank);
/* Instrumentation: enqueue ONE h_share_model pass (search or cached) at the device-resident state with no Kalman
 * step behind it, so that bench.py can bracket exactly that kernel with CUDA events. */
int lio_pass_only_enqueue(lio_ctx* ctx, int do_search, int extrinsic_est);
/* Instrumentation: with LIO_TIMELINE=1 in the environment at lio_create, the update kernels record %globaltimer at
 * their phase boundaries; out[0] = pairs written by block 0 from out[1] on as (tag, ns), out[128] = pairs written by
 * the block that ran the Kalman step from out[129] on.  Tags are listed in csrc/lio_pass.cu. */
int lio_debug_timeline(lio_ctx* ctx, int64_t out[256]);
/* Instrumentation (LIO_TIMELINE=1), the spread over the blocks in the last pass that ran: out[b] = ns at which worker
 * block b filed its partial row (b < 256), out[256 + b] = ns at which it left its search tiles (a cached pass leaves the
 * last search pass's value),
 * out[512 + w] = ns at which warp w of the solving block had seen its rows, out[544 + w] = had loaded them. */
int lio_debug_blocks(lio_ctx* ctx, int64_t out[768]);
/* Device pointer of the 92-double reduction blob {HtH 78, Hth 12, n_valid, n_searched} written by pass_enqueue. */
void* lio_blob_device_ptr(lio_ctx* ctx);
/* Makes the context reduce into / solve from a caller-owned device buffer of 92 doubles (e.g. the storage of a tensor
 * that an NCCL all-reduce works on in place); NULL restores the context's own buffer. */
int lio_blob_bind(lio_ctx* ctx, void* device_buffer);
/* Replaces the blob from host memory (a driver that sums the ranks' blobs on the host). */
int lio_blob_upload(lio_ctx* ctx, const double blob92[92]);
/* Synchronises and copies that blob to the host (tests, single-rank drivers). */
int lio_blob_download(lio_ctx* ctx, double blob92[92]);

/* ≙ the global Nearest_Points the update fills (esekfom.hpp:136; laserMapping.cpp:62,771) plus the per-point
 * intermediates of the last pass.  All outputs optional (NULL): idx5 m x 5, d2_5 m x 5, nbr_xyz m x 5 x 3,
 * world_xyz m x 3 (FP32 p_world), selected m (point_selected_surf), normvec m x 4 (a,b,c,pd2).
 * The neighbour rows are the reference's: Nearest_Search is called WITHOUT a distance bound (esekfom.hpp:140-141), so a
 * row holds min(5, live map points) neighbours of the point's last search pass however far away they are.  The update
 * itself only searches within d2 <= caps.knn_max_d2 (its gate, esekfom.hpp:144-147, needs no more); asking for the rows
 * here completes the short ones with the unbounded search first (one extra pass over those rows only). */
int lio_get_neighbors(lio_ctx* ctx, int32_t* idx5, float* d2_5, float* nbr_xyz, float* world_xyz, uint8_t* selected,
                      float* normvec);

/* ---- map maintenance: map_incremental (src/laserMapping.cpp:382-433) ----------------------------------- */
/* Classifies the current scan with the cached neighbours at state x and performs both Add_Points calls
 * (:430-431).  counts[0] = |PointToAdd|, counts[1] = |PointNoNeedDownsample|, counts[2] = added by the first call.
 * The classification is the reference's on its unbounded Nearest_Points: rows without a neighbour within
 * sqrt(caps.knn_max_d2) get their nearest map point from the unbounded search first (points_near[0] decides
 * PointNoNeedDownsample, :408-414).  filter_size_map is also the downsample size of the first Add_Points call. */
int lio_map_incremental(lio_ctx* ctx, const lio_state* x, float filter_size_map, int ekf_inited, int32_t counts[3]);

/* ≙ the first-scan branch of the main loop (laserMapping.cpp:747-758): every point of the current scan goes through
 * pointBodyToWorld (:277-288) at state x and the result is KD_TREE::Build'ed (ids 0..M-1, scan order). */
int lio_map_build_scan(lio_ctx* ctx, const lio_state* x);

/* ---- host-side, sequential (<= 50 steps / scan): ImuProcess forward half ------------------------------- */
/* ImuProcess members that persist between scans (src/IMU_Processing.hpp:95-136). */
typedef struct lio_imu_proc {
  double cov_gyr[3], cov_acc[3], cov_bias_gyr[3], cov_bias_acc[3];
  double cov_gyr_scale[3], cov_acc_scale[3];
  double mean_acc[3], mean_gyr[3];
  double acc_s_last[3], angvel_last[3];
  double lidar_T_wrt_imu[3], lidar_R_wrt_imu[9]; /* row-major */
  double last_lidar_end_time, first_lidar_time;
  lio_imu_sample last_imu;
  int32_t init_iter_num, imu_need_init, b_first_frame, pad_;
} lio_imu_proc; /* 51 doubles + 4 int32 */
/* ≙ ImuProcess::ImuProcess() (IMU_Processing.hpp:139-153). */
void lio_imu_proc_init(lio_imu_proc* ip);
/* ≙ ImuProcess::set_param (IMU_Processing.hpp:169-178; laserMapping.cpp:692-695). */
void lio_imu_set_param(lio_imu_proc* ip, const double transl[3], const double rot[9], const double gyr[3],
                       const double acc[3], const double gyr_bias[3], const double acc_bias[3]);
/* ≙ ImuProcess::Process (IMU_Processing.hpp:405-441) without the per-point loop: while the filter is initialising
 * (IMU_init, :180-244) it updates x / P from the IMU statistics and returns *initialising = 1, *n_poses = 0;
 * afterwards it runs the forward half of UndistortPcl (:258-358): x, P propagated to the scan end, IMUpose list
 * (n_imu + 1 entries) for lio_scan_preprocess.  imu = meas.imu of this scan. */
int lio_imu_process(lio_imu_proc* ip, const lio_imu_sample* imu, int n_imu, double lidar_beg_time, double lidar_end_time,
                    lio_state* x, double P[576], lio_pose6d* poses, int cap, int* n_poses, int* initialising);
/* ≙ esekf::predict (esekfom.hpp:82-95; use-ikfom.hpp:57-123).  Q is 12x12 row-major. */
int lio_predict(lio_state* x, double P[576], double dt, const double Q[144], const double acc[3], const double gyro[3]);
/* ≙ esekf::boxplus / boxminus (esekfom.hpp:59-73, 236-258). */
int lio_boxplus(const lio_state* x, const double f[24], lio_state* out);
int lio_boxminus(const lio_state* x1, const lio_state* x2, double out[24]);


/* ---- the per-scan main loop as one object (src/laserMapping.cpp:702-800) ---------------------------------- */
/* One lio_seq = one sequence (one robot / one bag) on one context: ImuProcess (IMU_Processing.hpp:79-136), the loop's
 * bookkeeping (flg_first_scan, first_lidar_time, flg_EKF_inited), the sliding local-map box of lasermap_fov_segment
 * (:309-365) and the filter state.  Host code; the device work of an iteration is one lio_scan_step. */
typedef struct lio_seq lio_seq;
typedef struct lio_seq_config {
  float filter_size_surf, filter_size_map; /* filter_size_surf_min / filter_size_map_min (launch files) */
  int32_t max_iteration;                   /* NUM_MAX_ITERATIONS */
  int32_t extrinsic_est;                   /* extrinsic_est_en */
  double extrinsic_T[3], extrinsic_R[9];   /* mapping/extrinsic_T, extrinsic_R (row-major) */
  double gyr_cov, acc_cov, b_gyr_cov, b_acc_cov;
  double cube_len, det_range;              /* cube_side_length, DET_RANGE */
  double laser_point_cov;                  /* LASER_POINT_COV */
} lio_seq_config;
/* One synchronised MeasureGroup (common_lib.h:40-49) as sync_packages (laserMapping.cpp:218-275) hands it over. */
typedef struct lio_seq_input {
  const void* lidar;      /* n records of stride_bytes (16: x,y,z,t_ms; 48: PointXYZINormal, time in curvature) */
  int64_t n;
  int32_t stride_bytes;
  int32_t n_imu;
  const lio_imu_sample* imu;
  double lidar_beg_time, lidar_end_time;
} lio_seq_input;
enum {
  LIO_SEQ_UPDATED = 0,    /* the odometry in result.x is new */
  LIO_SEQ_FEW_POINTS = 1, /* feats_down_size < 5 (:741-744) */
  LIO_SEQ_MAP_BUILT = 2,  /* first scan with points: ikdtree.Build (:747-758) */
  LIO_SEQ_FIRST_SCAN = 3, /* flg_first_scan (:711-716) */
  LIO_SEQ_NO_IMU = 4,     /* meas.imu.empty() */
  LIO_SEQ_IMU_INIT = 5    /* IMU_init still collecting (:722-725) */
};
typedef struct lio_seq_result {
  int32_t status; /* LIO_SEQ_* */
  int32_t n_valid, n_passes;
  int32_t counts[3];
  int64_t m;
  int64_t n_box_deleted; /* running total of points removed by the sliding local-map box */
  lio_state x;           /* state after this scan (what the reference publishes as odometry) */
} lio_seq_result;
void lio_seq_default_config(lio_seq_config* cfg);
int lio_seq_create(lio_ctx*, const lio_seq_config* cfg, lio_seq** out);
void lio_seq_destroy(lio_seq*);
/* One iteration of the main loop. */
int lio_seq_process(lio_seq*, const lio_seq_input* in, lio_seq_result* result);
/* One iteration for n_seq <= 64 INDEPENDENT sequences (each with its own context on the same device; BASELINE.json
 * config 4): host stages and preprocessing per sequence on its context's stream, the updates that are due in one
 * cooperative launch per group of 8 (lio_update_enqueue_multi), map growth per sequence, one synchronisation per
 * sequence at the end.  in / result: arrays of n_seq.  All sequences use the filter settings of the first one. */
int lio_seq_process_many(lio_seq* const* seqs, int n_seq, const lio_seq_input* in, lio_seq_result* results);
/* Host threads lio_seq_process_many spreads the per-sequence host work over (IMU propagation, uploads, kernel enqueues:
 * independent contexts and streams).  0 = default: LIO_HOST_THREADS from the environment, else min(8, cores / 2).
 * Process-wide.  Results do not depend on it. */
int lio_set_host_threads(int n);
/* Filter state in / out between scans (relocalisation as in laserMapping_re.cpp:590-600, teacher-forced tests). */
int lio_seq_get_state(const lio_seq*, lio_state* x, double P[576]);
int lio_seq_set_state(lio_seq*, const lio_state* x, const double P[576]);
/* LocalMap_Points {min, max} and the running box-delete count; LIO_E_EMPTY_MAP before the box exists. */
int lio_seq_local_map(const lio_seq*, float box6[6], int64_t* n_box_deleted);

#ifdef __cplusplus
}
#endif
#endif /* LIO_B200_H */
