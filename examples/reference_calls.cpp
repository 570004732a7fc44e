// The call expressions laserMapping.cpp makes on the hot path, verbatim, compiled against include/lio_facade.hpp:
//   :361-364  acquire_removed_points / Delete_Point_Boxes        :683      downSizeFilterSurf.setLeafSize
//   :430-431  the two Add_Points calls of map_incremental        :737-738  setInputCloud + filter
//   :747-756  Root_Node / set_downsample_param / Build           :771-774  update_iterated_dyn_share_modified
// and esekfom.hpp:300's h_share_model(dyn_share, feats_down_body, ikdtree, Nearest_Points, extrinsic_est).
// Only the declarations around them differ from the reference (no ROS / PCL / Eigen here): the objects are the facade's.
//   g++ -std=c++14 -O2 -Iinclude examples/reference_calls.cpp -Lagi_lidar_slam_b200 -llio_b200 -Wl,-rpath,... -o ...
// Exits 0 when every call did what the reference's would (checked against values computed on the host below).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <vector>

#include "lio_facade.hpp"

using namespace std;
using namespace lio_b200;

struct PointType {  // pcl::PointXYZINormal's layout (common_lib.h:26)
  float x, y, z, pad0;
  float normal_x, normal_y, normal_z, pad1;
  float intensity, curvature, pad2, pad3;
};
struct PointCloudXYZI {  // what the path touches of pcl::PointCloud<PointType>
  typedef shared_ptr<PointCloudXYZI> Ptr;
  vector<PointType> points;
  void resize(size_t n) { points.resize(n); }
  void clear() { points.clear(); }
};
typedef vector<PointType> PointVector;
typedef lio_state state_ikfom;

#define LASER_POINT_COV (0.001)
static const int NUM_MAX_ITERATIONS = 4;
static const bool extrinsic_est_en = false;
static const float filter_size_surf_min = 0.5f, filter_size_map_min = 0.5f;

static PointType pt(float x, float y, float z) {
  PointType p;
  memset(&p, 0, sizeof(p));
  p.x = x;
  p.y = y;
  p.z = z;
  return p;
}

#define CHECK(cond)                                                  \
  do {                                                               \
    if (!(cond)) {                                                   \
      fprintf(stderr, "line %d: %s does not hold\n", __LINE__, #cond); \
      return 1;                                                      \
    }                                                                \
  } while (0)

int main() {
  auto ctx = make_shared<Context>(0);
  if (!ctx->ok()) {
    fprintf(stderr, "no usable B200-class device (error %d): the library has no CPU fallback\n", ctx->last_error);
    return 2;
  }
  KD_TREE<PointType> ikdtree(ctx);
  esekf kf;
  VoxelGrid<PointType> downSizeFilterSurf(ctx);
  PointCloudXYZI::Ptr feats_undistort(new PointCloudXYZI());
  PointCloudXYZI::Ptr feats_down_body(new PointCloudXYZI());
  PointCloudXYZI::Ptr feats_down_world(new PointCloudXYZI());
  vector<PointVector> Nearest_Points;
  vector<BoxPointType> cub_needrm;
  int kdtree_delete_counter = 0, add_point_size = 0, feats_down_size = 0;
  state_ikfom state_point;

  // a ground plane and two walls, several raw points per 0.5 m leaf; the sensor sits at the origin, body = world
  mt19937 rng(11);
  uniform_real_distribution<float> u(0.f, 1.f);
  normal_distribution<float> g(0.f, 0.005f);
  for (int i = -40; i < 40; ++i)
    for (int j = -40; j < 40; ++j)
      for (int r = 0; r < 3; ++r) feats_undistort->points.push_back(pt(0.5f * (i + u(rng)), 0.5f * (j + u(rng)), -1.3f + g(rng)));
  for (int j = -40; j < 40; ++j)
    for (int k = 0; k < 10; ++k)
      for (int r = 0; r < 3; ++r) {
        feats_undistort->points.push_back(pt(15.2f + g(rng), 0.5f * (j + u(rng)), -1.5f + 0.5f * (k + u(rng))));
        feats_undistort->points.push_back(pt(0.5f * (j + u(rng)), -12.2f + g(rng), -1.5f + 0.5f * (k + u(rng))));
      }

  // ---- laserMapping.cpp:683
  downSizeFilterSurf.setLeafSize(filter_size_surf_min, filter_size_surf_min, filter_size_surf_min);
  // ---- :737-739
  downSizeFilterSurf.setInputCloud(feats_undistort);
  downSizeFilterSurf.filter(*feats_down_body);
  feats_down_size = feats_down_body->points.size();
  CHECK(downSizeFilterSurf.last_error == LIO_OK);
  CHECK(feats_down_size > 5000 && feats_down_size < (int)feats_undistort->points.size() / 2);
  {  // every output is the centroid of the input points of its leaf: check one leaf by hand
    const PointType c = feats_down_body->points[feats_down_size / 2];
    const int kx = (int)floorf(c.x / 0.5f), ky = (int)floorf(c.y / 0.5f), kz = (int)floorf(c.z / 0.5f);
    float sx = 0, sy = 0, sz = 0;
    int n = 0;
    for (const PointType& p : feats_undistort->points)
      if ((int)floorf(p.x / 0.5f) == kx && (int)floorf(p.y / 0.5f) == ky && (int)floorf(p.z / 0.5f) == kz) {
        sx += p.x;
        sy += p.y;
        sz += p.z;
        ++n;
      }
    CHECK(n > 0 && c.x == sx / n && c.y == sy / n && c.z == sz / n);
  }

  // ---- :747-758 (first scan: the map is built from the scan; the state is the identity, so world = body)
  if (ikdtree.Root_Node == nullptr) {
    ikdtree.set_downsample_param(filter_size_map_min);
    feats_down_world->resize(feats_down_size);
    for (int i = 0; i < feats_down_size; i++) feats_down_world->points[i] = feats_down_body->points[i];
    ikdtree.Build(feats_down_world->points);
  }
  CHECK(ikdtree.Root_Node != nullptr && ikdtree.validnum() == feats_down_size);

  // ---- :771-776 (second scan = the same surfaces seen from a pose 5 cm / 0.2 deg off; the update must find it)
  {
    lio_state x = kf.get_x();
    x.pos[0] = 0.05;
    x.pos[1] = -0.03;
    x.rot[0] = cos(0.0035 / 2);
    x.rot[3] = sin(0.0035 / 2);
    kf.change_x(x);
    double P[576] = {0};
    for (int i = 0; i < 24; ++i) P[i * 24 + i] = i < 6 ? 1e-2 : 1e-6;
    kf.change_P(P);
  }
  Nearest_Points.resize(feats_down_size);
  kf.update_iterated_dyn_share_modified(LASER_POINT_COV, feats_down_body, ikdtree, Nearest_Points, NUM_MAX_ITERATIONS,
                                        extrinsic_est_en);
  state_point = kf.get_x();
  CHECK(kf.last_error == LIO_OK && kf.effct_feat_num > feats_down_size / 2);
  CHECK(fabs(state_point.pos[0]) < 5e-3 && fabs(state_point.pos[1]) < 5e-3 && fabs(state_point.rot[3]) < 5e-4);
  CHECK((int)Nearest_Points.size() == feats_down_size && Nearest_Points[0].size() == 5);

  // ---- esekfom.hpp:300 (one measurement-model pass at the posterior): rows = matched points, residuals small
  {
    dyn_share_datastruct dyn_share;
    dyn_share.valid = true;
    dyn_share.converge = true;
    kf.h_share_model(dyn_share, feats_down_body, ikdtree, Nearest_Points, extrinsic_est_en);
    CHECK(kf.last_error == LIO_OK && dyn_share.valid);
    CHECK((int)dyn_share.h.size() == kf.effct_feat_num && dyn_share.h_x.size() == 12 * dyn_share.h.size());
    double worst = 0, nrm = 0;
    for (size_t i = 0; i < dyn_share.h.size(); ++i) {
      worst = fmax(worst, fabs(dyn_share.h[i]));
      const double* r = &dyn_share.h_x[12 * i];
      nrm = fmax(nrm, fabs(sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) - 1.0));
      CHECK(r[6] == 0 && r[11] == 0);  // extrinsic_est off (:221-223)
    }
    CHECK(worst < 0.1 && nrm < 1e-3);  // point-to-plane distances at the solution; unit normals
  }

  // ---- :430-431 (map_incremental's two calls): a strip of new ground beyond the old one, and a few loose points
  PointVector PointToAdd, PointNoNeedDownsample;
  for (int i = 40; i < 50; ++i)
    for (int j = -40; j < 40; ++j) PointToAdd.push_back(pt(0.5f * (i + u(rng)), 0.5f * (j + u(rng)), -1.3f));
  for (int j = 0; j < 7; ++j) PointNoNeedDownsample.push_back(pt(40.f + j, 40.f, 3.f));
  const int before = ikdtree.validnum();
  add_point_size = ikdtree.Add_Points(PointToAdd, true);
  ikdtree.Add_Points(PointNoNeedDownsample, false);
  CHECK(add_point_size == (int)PointToAdd.size());  // one point per empty 0.5 m box: all of them stay
  CHECK(ikdtree.validnum() == before + add_point_size + (int)PointNoNeedDownsample.size());
  add_point_size = PointToAdd.size() + PointNoNeedDownsample.size();

  // ---- :359-364 (lasermap_fov_segment: the map behind the sensor is cut away)
  BoxPointType box;
  box.vertex_min[0] = -30.f; box.vertex_min[1] = -30.f; box.vertex_min[2] = -5.f;
  box.vertex_max[0] = -10.f; box.vertex_max[1] = 30.f;  box.vertex_max[2] = 5.f;
  cub_needrm.push_back(box);
  PointVector points_history;
  ikdtree.acquire_removed_points(points_history);
  const size_t history_before = points_history.size();  // (downsample replacements of the Add_Points above: none here)
  if (cub_needrm.size() > 0) kdtree_delete_counter = ikdtree.Delete_Point_Boxes(cub_needrm);
  CHECK(kdtree_delete_counter > 1000);
  ikdtree.acquire_removed_points(points_history);
  CHECK((int)(points_history.size() - history_before) == kdtree_delete_counter);
  for (size_t i = history_before; i < points_history.size(); ++i) {
    const PointType& p = points_history[i];
    CHECK(p.x >= -30.f && p.x < -10.f && p.y >= -30.f && p.y < 30.f);
  }
  ikdtree.acquire_removed_points(points_history);  // taken: nothing new
  CHECK((int)(points_history.size() - history_before) == kdtree_delete_counter);
  printf("reference call expressions: feats_down_size %d, effct_feat_num %d, |pos| %.1e, added %d, box-deleted %d\n",
         feats_down_size, kf.effct_feat_num, hypot(state_point.pos[0], state_point.pos[1]), add_point_size,
         kdtree_delete_counter);
  return 0;
}
