// lio_facade.hpp — the reference's C++ call surface for the hot path, re-created on top of the C-ABI (lio_b200.h).
//
// S-FAST_LIO's host code talks to three objects on this path (paths relative to src/S-FAST_LIO/):
//   KD_TREE<PointType> ikdtree     include/ikd-Tree/ikd_Tree.h:264-299   Build / Nearest_Search / Add_Points / ...
//   esekfom::esekf kf              include/esekfom.hpp:36-346             predict / update_iterated_dyn_share_modified
//   ImuProcess + pcl::VoxelGrid    src/IMU_Processing.hpp:79-90, src/laserMapping.cpp:683,737-738
// The classes below keep those names, argument orders and (absent) error behaviour, so that laserMapping.cpp keeps
// compiling with the include lines swapped (INTEGRATION.md shows the diff).  They are header-only C++14 and need
// nothing but the C library: PointType is any struct whose first three floats are x, y, z and whose size is 16 or
// 48 bytes (pcl::PointXYZINormal qualifies).  Types that come from Eigen / Sophus in the reference (state_ikfom, cov)
// are the plain lio_state / double[576] here; INTEGRATION.md has the two 10-line converters.
//
// Error behaviour: the reference calls return void and fail silently (ROS_WARN + skip).  The facade does the same but
// keeps the last LIO_E_* code in `last_error` so that a caller who cares can look.
#ifndef LIO_FACADE_HPP
#define LIO_FACADE_HPP

#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "lio_b200.h"

namespace lio_b200 {

struct BoxPointType {  // ikd_Tree.h:30-33
  float vertex_min[3];
  float vertex_max[3];
};
enum delete_point_storage_set { NOT_RECORD, DELETE_POINTS_REC, MULTI_THREAD_REC };  // ikd_Tree.h:39-43

// One GPU context shared by the map, the filter and the preprocessing of one node (the reference is one process).
class Context {
 public:
  explicit Context(int device = 0, const lio_caps* caps = nullptr) { last_error = lio_create(device, caps, &ctx_); }
  ~Context() { lio_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  lio_ctx* get() const { return ctx_; }
  bool ok() const { return ctx_ != nullptr; }
  int last_error = LIO_OK;

 private:
  lio_ctx* ctx_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------
// KD_TREE<PointType> (ikd_Tree.h:46-299): the calls the hot path and its callers make.
// ---------------------------------------------------------------------------------------------------------
template <typename PointType, typename Alloc = std::allocator<PointType>>
class KD_TREE {
  static_assert(sizeof(PointType) == 16 || sizeof(PointType) == 48, "PointType must be 16 or 48 bytes (x,y,z first)");

 public:
  using PointVector = std::vector<PointType, Alloc>;
  using Ptr = std::shared_ptr<KD_TREE<PointType, Alloc>>;
  struct KD_TREE_NODE {};  // opaque: host code only tests Root_Node against nullptr (laserMapping.cpp:747)

  explicit KD_TREE(std::shared_ptr<Context> ctx, float = 0.5f, float = 0.6f, float = 0.2f) : ctx_(std::move(ctx)) {}
  void Set_delete_criterion_param(float) {}   // a hash map has no balance / lazy-delete criteria
  void Set_balance_criterion_param(float) {}
  void set_downsample_param(float downsample_param) {
    downsample_size = downsample_param;
    last_error = lio_map_set_downsample(ctx_->get(), downsample_param);
  }
  void InitializeKDTree(float = 0.5f, float = 0.7f, float = 0.2f) {}

  int size() {
    int64_t t = 0, v = 0;
    last_error = lio_map_size(ctx_->get(), &t, &v);
    return (int)t;
  }
  int validnum() {
    int64_t t = 0, v = 0;
    last_error = lio_map_size(ctx_->get(), &t, &v);
    return (int)v;
  }
  // ikd_Tree.cpp:355-367 — by value, like the reference
  void Build(PointVector point_cloud) {
    last_error = lio_map_build(ctx_->get(), point_cloud.data(), (int64_t)point_cloud.size(), (int)sizeof(PointType));
    Root_Node = (last_error == LIO_OK && !point_cloud.empty()) ? &root_token_ : nullptr;
  }
  // ikd_Tree.cpp:370-402.  k_nearest <= 5 is served (NUM_MATCH_POINTS, what the path asks for); outputs ascending by
  // distance; fewer than k results only where the map holds fewer points within max_dist, exactly as the reference.
  void Nearest_Search(PointType point, int k_nearest, PointVector& Nearest_Points, std::vector<float>& Point_Distance,
                      float max_dist = INFINITY) {
    PointVector().swap(Nearest_Points);
    std::vector<float>().swap(Point_Distance);
    const float* q = reinterpret_cast<const float*>(&point);
    int32_t idx[5];
    float d2[5], xyz[15];
    last_error = lio_knn5(ctx_->get(), q, 1, max_dist * max_dist, idx, d2, xyz);  // ikd_Tree.cpp:965
    if (last_error != LIO_OK) return;
    for (int r = 0; r < 5 && r < k_nearest; ++r) {
      if (idx[r] < 0) break;
      PointType p;
      std::memset(&p, 0, sizeof(p));
      std::memcpy(&p, xyz + 3 * r, 12);
      Nearest_Points.push_back(p);
      Point_Distance.push_back(d2[r]);
    }
  }
  // ikd_Tree.cpp:419-512
  int Add_Points(PointVector& PointToAdd, bool downsample_on) {
    int32_t added = 0;
    last_error = lio_map_add(ctx_->get(), PointToAdd.data(), (int64_t)PointToAdd.size(), (int)sizeof(PointType),
                             downsample_on ? 1 : 0, &added);
    return added;
  }
  // ikd_Tree.cpp:559-579
  int Delete_Point_Boxes(std::vector<BoxPointType>& BoxPoints) {
    int32_t deleted = 0;
    last_error = lio_map_delete_boxes(ctx_->get(), reinterpret_cast<const float*>(BoxPoints.data()),
                                      (int)BoxPoints.size(), &deleted);
    return deleted;
  }
  // ikd_Tree.cpp:1490-1516 (root ignored: there is one tree)
  void flatten(KD_TREE_NODE*, PointVector& Storage, delete_point_storage_set) {
    int64_t n = 0;
    last_error = lio_map_dump(ctx_->get(), nullptr, nullptr, 0, &n);
    std::vector<float> xyz((size_t)n * 3);
    if (n > 0) last_error = lio_map_dump(ctx_->get(), xyz.data(), nullptr, n, &n);
    Storage.clear();
    Storage.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      PointType p;
      std::memset(&p, 0, sizeof(p));
      std::memcpy(&p, &xyz[3 * i], 12);
      Storage.push_back(p);
    }
  }
  // ikd_Tree.cpp:582-594: APPENDS the points deleted since the last call (box deletes, downsample replacements) and
  // clears the log.  Order unspecified, as in the reference (it files them rebuild by rebuild).
  void acquire_removed_points(PointVector& removed_points) {
    int64_t n = 0;
    last_error = lio_map_removed_points(ctx_->get(), nullptr, 0, &n);
    if (last_error != LIO_OK || n == 0) return;
    std::vector<float> xyz((size_t)n * 3);
    last_error = lio_map_removed_points(ctx_->get(), xyz.data(), n, &n);
    for (int64_t i = 0; i < n; ++i) {
      PointType p;
      std::memset(&p, 0, sizeof(p));
      std::memcpy(&p, &xyz[3 * i], 12);
      removed_points.push_back(p);
    }
  }

  PointVector PCL_Storage;
  KD_TREE_NODE* Root_Node = nullptr;
  float downsample_size = 0.5f;
  int last_error = LIO_OK;
  const std::shared_ptr<Context>& context() const { return ctx_; }

 private:
  std::shared_ptr<Context> ctx_;
  KD_TREE_NODE root_token_;
};

// ---------------------------------------------------------------------------------------------------------
// esekfom::esekf (esekfom.hpp:36-346)
// ---------------------------------------------------------------------------------------------------------
struct input_ikfom {  // use-ikfom.hpp:29-33
  double acc[3];
  double gyro[3];
};

// esekfom.hpp:15-22 without Eigen: h is effct_feat_num long, h_x is effct_feat_num x 12, row-major.
struct dyn_share_datastruct {
  bool valid = true;
  bool converge = true;
  std::vector<double> h;
  std::vector<double> h_x;
};

class esekf {
 public:
  esekf() {
    std::memset(&x_, 0, sizeof(x_));
    x_.rot[0] = 1.0;
    x_.offset_R_L_I[0] = 1.0;
    x_.grav[2] = -9.81;  // state_ikfom defaults (use-ikfom.hpp:18-27)
    std::memset(P_, 0, sizeof(P_));
    for (int i = 0; i < 24; ++i) P_[i * 24 + i] = 1.0;
  }
  lio_state get_x() const { return x_; }
  const double* get_P() const { return P_; }
  void change_x(const lio_state& s) { x_ = s; }
  void change_P(const double P[576]) { std::memcpy(P_, P, sizeof(P_)); }
  // esekfom.hpp:82-95
  void predict(double& dt, const double Q[144], const input_ikfom& i_in) {
    last_error = lio_predict(&x_, P_, dt, Q, i_in.acc, i_in.gyro);
  }
  // esekfom.hpp:270-346.  feats_down_body: the downsampled scan; when it is the cloud lio_scan_preprocess just left on
  // the device pass nullptr / 0 and nothing is copied.  Nearest_Points stays on the device: map_incremental should use
  // lio_map_incremental (device side); the overload below / fetch_neighbors() copy the cache for host code that wants it.
  template <typename PointType, typename Alloc>
  void update_iterated_dyn_share_modified(double R, const PointType* feats_down_body, int64_t feats_down_size,
                                          KD_TREE<PointType, Alloc>& ikdtree, int maximum_iter, bool extrinsic_est) {
    lio_ctx* c = ikdtree.context()->get();
    if (feats_down_body != nullptr) {
      last_error = lio_scan_upload(c, feats_down_body, feats_down_size, (int)sizeof(PointType));
      if (last_error != LIO_OK) return;
    }
    last_error = lio_update_scan(c, &x_, P_, R, maximum_iter, extrinsic_est ? 1 : 0, &effct_feat_num, &n_passes);
  }
  // The reference's own argument list (esekfom.hpp:270-275): feats_down_body is anything that points at an object with a
  // `points` vector (PointCloudXYZI::Ptr qualifies), Nearest_Points comes back filled as the reference leaves it -- the
  // neighbours of the last SEARCH pass from an UNBOUNDED search (esekfom.hpp:140-141), ascending by distance, five per
  // point wherever the map holds five points -- for host code that still runs map_incremental
  // (laserMapping.cpp:382-433) itself.
  template <typename CloudPtr, typename PointType, typename Alloc>
  void update_iterated_dyn_share_modified(double R, CloudPtr& feats_down_body, KD_TREE<PointType, Alloc>& ikdtree,
                                          std::vector<std::vector<PointType, Alloc>>& Nearest_Points, int maximum_iter,
                                          bool extrinsic_est) {
    const auto& pts = feats_down_body->points;
    update_iterated_dyn_share_modified(R, pts.data(), (int64_t)pts.size(), ikdtree, maximum_iter, extrinsic_est);
    if (last_error == LIO_OK) fetch_neighbors(ikdtree, pts.size(), Nearest_Points);
  }
  // Copies the device-side Nearest_Points cache of the last update (m = feats_down_size) into the reference's container.
  template <typename PointType, typename Alloc>
  void fetch_neighbors(KD_TREE<PointType, Alloc>& ikdtree, size_t m,
                       std::vector<std::vector<PointType, Alloc>>& Nearest_Points) {
    std::vector<int32_t> idx(5 * m);
    std::vector<float> xyz(15 * m);
    last_error = lio_get_neighbors(ikdtree.context()->get(), idx.data(), nullptr, xyz.data(), nullptr, nullptr, nullptr);
    if (last_error != LIO_OK) return;
    Nearest_Points.resize(m);  // laserMapping.cpp:771
    for (size_t i = 0; i < m; ++i) {
      Nearest_Points[i].clear();
      for (int r = 0; r < 5 && idx[5 * i + r] >= 0; ++r) {
        PointType p;
        std::memset(&p, 0, sizeof(p));
        std::memcpy(&p, &xyz[15 * i + 3 * r], 12);
        Nearest_Points[i].push_back(p);
      }
    }
  }
  // esekfom.hpp:106-227: ONE measurement-model pass at the filter's current state x_.  ekfom_data.converge says whether
  // the neighbours are searched again (:136-141); on return valid / h / h_x are the reference's: one row per matched
  // point, in scan order (laserCloudOri is filled in index order, :176-187), h_x = [n^T, A^T, B^T, C^T] (:197-226) with
  // B = C = 0 unless extrinsic_est, h = -pd2.  The pass itself runs on the device (lio_update_pass); the rows are formed
  // here from its per-point outputs (normal, residual) and the state, in the reference's FP64 arithmetic.
  template <typename CloudPtr, typename PointType, typename Alloc>
  void h_share_model(dyn_share_datastruct& ekfom_data, CloudPtr& feats_down_body, KD_TREE<PointType, Alloc>& ikdtree,
                     std::vector<std::vector<PointType, Alloc>>& Nearest_Points, bool extrinsic_est) {
    lio_ctx* c = ikdtree.context()->get();
    const auto& pts = feats_down_body->points;
    const size_t m = pts.size();
    last_error = lio_scan_upload(c, pts.data(), (int64_t)m, (int)sizeof(PointType));
    if (last_error != LIO_OK) return;
    double blob[90];
    last_error = lio_update_pass(c, &x_, ekfom_data.converge ? 1 : 0, extrinsic_est ? 1 : 0, blob, &effct_feat_num);
    if (last_error != LIO_OK) return;
    if (ekfom_data.converge) fetch_neighbors(ikdtree, m, Nearest_Points);
    std::vector<uint8_t> sel(m);
    std::vector<float> nv(4 * m);
    last_error = lio_get_neighbors(c, nullptr, nullptr, nullptr, nullptr, sel.data(), nv.data());
    if (last_error != LIO_OK) return;
    ekfom_data.h.clear();
    ekfom_data.h_x.clear();
    if (effct_feat_num < 1) {  // :189-194
      ekfom_data.valid = false;
      return;
    }
    double R[9], Rli[9];
    quat_to_mat(x_.rot, R);
    quat_to_mat(x_.offset_R_L_I, Rli);
    for (size_t i = 0; i < m; ++i) {
      if (!sel[i]) continue;
      const float* pf = reinterpret_cast<const float*>(&pts[i]);
      const double pb[3] = {pf[0], pf[1], pf[2]};
      double pI[3], C[3], A[3], B[3] = {0.0, 0.0, 0.0};
      for (int r = 0; r < 3; ++r)
        pI[r] = (Rli[3 * r] * pb[0] + Rli[3 * r + 1] * pb[1]) + Rli[3 * r + 2] * pb[2] + x_.offset_T_L_I[r];  // :206
      const double n[3] = {nv[4 * i], nv[4 * i + 1], nv[4 * i + 2]};
      for (int r = 0; r < 3; ++r) C[r] = (R[r] * n[0] + R[3 + r] * n[1]) + R[6 + r] * n[2];  // rot^T * norm_vec (:214)
      A[0] = pI[1] * C[2] - pI[2] * C[1];                                                       // [pI]x * C (:215)
      A[1] = pI[2] * C[0] - pI[0] * C[2];
      A[2] = pI[0] * C[1] - pI[1] * C[0];
      double row[12] = {n[0], n[1], n[2], A[0], A[1], A[2], 0, 0, 0, 0, 0, 0};
      if (extrinsic_est) {  // :216-221
        double D[3];  // offset_R_L_I^T * C
        for (int r = 0; r < 3; ++r) D[r] = (Rli[r] * C[0] + Rli[3 + r] * C[1]) + Rli[6 + r] * C[2];
        B[0] = pb[1] * D[2] - pb[2] * D[1];
        B[1] = pb[2] * D[0] - pb[0] * D[2];
        B[2] = pb[0] * D[1] - pb[1] * D[0];
        for (int r = 0; r < 3; ++r) {
          row[6 + r] = B[r];
          row[9 + r] = C[r];
        }
      }
      ekfom_data.h_x.insert(ekfom_data.h_x.end(), row, row + 12);
      ekfom_data.h.push_back(-(double)nv[4 * i + 3]);  // :225
    }
  }
  int32_t effct_feat_num = 0;  // esekfom.hpp:26
  int32_t n_passes = 0;
  int last_error = LIO_OK;

 private:
  static void quat_to_mat(const double q[4], double m[9]) {  // w, x, y, z -> row-major rotation matrix
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = 1 - 2 * (y * y + z * z);
    m[1] = 2 * (x * y - w * z);
    m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z);
    m[4] = 1 - 2 * (x * x + z * z);
    m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y);
    m[7] = 2 * (y * z + w * x);
    m[8] = 1 - 2 * (x * x + y * y);
  }
  lio_state x_;
  double P_[576];
};

// ---------------------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointType> as laserMapping.cpp uses it (:66-67 declaration, :683-686 setLeafSize, :737-738
// setInputCloud + filter): the centroid of every occupied leaf, in ascending leaf-index order (PCL's own order).
// Cloud types: anything with a `points` vector (and optionally width / height / is_dense, which are not touched).
// ---------------------------------------------------------------------------------------------------------
template <typename PointType>
class VoxelGrid {
  static_assert(sizeof(PointType) == 16 || sizeof(PointType) == 48, "PointType must be 16 or 48 bytes (x,y,z first)");

 public:
  explicit VoxelGrid(std::shared_ptr<Context> ctx) : ctx_(std::move(ctx)) {}
  void setLeafSize(float lx, float ly, float lz) {
    leaf_ = lx;
    cubic_ = (lx == ly && ly == lz);  // the path only ever sets cubic leaves; anything else is refused in filter()
  }
  template <typename CloudPtr>
  void setInputCloud(const CloudPtr& cloud) {
    in_ = cloud->points.data();
    n_ = (int64_t)cloud->points.size();
  }
  template <typename Cloud>
  void filter(Cloud& output) {
    output.points.clear();
    if (!cubic_ || !(leaf_ > 0.f) || in_ == nullptr) {
      last_error = LIO_E_INVALID;
      return;
    }
    std::vector<PointType> out((size_t)(n_ < 1 ? 1 : n_));  // at most one centroid per input point
    int64_t m = 0;
    last_error = lio_scan_preprocess(ctx_->get(), in_, n_, (int)sizeof(PointType), nullptr, 0, nullptr, leaf_, out.data(),
                                     &m, nullptr, nullptr);
    if (last_error != LIO_OK) return;
    out.resize((size_t)m);
    output.points.assign(out.begin(), out.end());
  }
  int last_error = LIO_OK;

 private:
  std::shared_ptr<Context> ctx_;
  const PointType* in_ = nullptr;
  int64_t n_ = 0;
  float leaf_ = 0.f;
  bool cubic_ = false;
};

// ---------------------------------------------------------------------------------------------------------
// ImuProcess (IMU_Processing.hpp:79-136) + pcl::VoxelGrid (laserMapping.cpp:683,737-738), fused on the device
// ---------------------------------------------------------------------------------------------------------
class ImuProcess {
 public:
  ImuProcess() { lio_imu_proc_init(&ip_); }
  void set_param(const double transl[3], const double rot[9], const double gyr[3], const double acc[3],
                 const double gyr_bias[3], const double acc_bias[3]) {
    lio_imu_set_param(&ip_, transl, rot, gyr, acc, gyr_bias, acc_bias);
  }
  // Process (:405-441) followed by downSizeFilterSurf.filter (laserMapping.cpp:737-738).  Returns feats_down_size,
  // 0 while the filter is initialising (the reference leaves the cloud empty and the caller skips the scan).
  template <typename PointType>
  int64_t Process(Context& ctx, const PointType* lidar, int64_t n, const lio_imu_sample* imu, int n_imu,
                  double lidar_beg_time, double lidar_end_time, esekf& kf_state, float filter_size_surf) {
    lio_state x = kf_state.get_x();
    double P[576];
    std::memcpy(P, kf_state.get_P(), sizeof(P));
    poses_.resize((size_t)n_imu + 2);
    int n_poses = 0, initialising = 0;
    last_error = lio_imu_process(&ip_, imu, n_imu, lidar_beg_time, lidar_end_time, &x, P, poses_.data(),
                                 (int)poses_.size(), &n_poses, &initialising);
    if (last_error != LIO_OK) return 0;
    kf_state.change_x(x);
    kf_state.change_P(P);
    if (initialising || n_imu == 0) return 0;
    int64_t m = 0;
    last_error = lio_scan_preprocess_resident(ctx.get(), lidar, n, (int)sizeof(PointType), poses_.data(), n_poses, &x,
                                              filter_size_surf, &m);
    return last_error == LIO_OK ? m : 0;
  }
  double first_lidar_time = 0.0;
  int last_error = LIO_OK;

 private:
  lio_imu_proc ip_;
  std::vector<lio_pose6d> poses_;
};

}  // namespace lio_b200
#endif  // LIO_FACADE_HPP
