// Minimal ROS-free C++ host on the reference's call surface (include/lio_facade.hpp): build a map with
// ikdtree.Build, run kf.update_iterated_dyn_share_modified on one scan, grow the map with Add_Points.
//   g++ -std=c++14 -O2 -Iinclude examples/replay_main.cpp -Lagi_lidar_slam_b200 -llio_b200
//       -Wl,-rpath,$PWD/agi_lidar_slam_b200 -o examples/replay_main        (one command line)
// Prints the recovered pose error; exits non-zero when the update does not pull the perturbed prior to the truth.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "lio_facade.hpp"

struct PointXYZINormal {  // pcl::PointXYZINormal layout (common_lib.h:26), 48 bytes
  float x, y, z, pad0;
  float normal_x, normal_y, normal_z, pad1;
  float intensity, curvature, pad2, pad3;
};
using PointType = PointXYZINormal;
using Tree = lio_b200::KD_TREE<PointType>;

static PointType make_point(float x, float y, float z) {
  PointType p;
  std::memset(&p, 0, sizeof(p));
  p.x = x;
  p.y = y;
  p.z = z;
  p.pad0 = 1.0f;
  return p;
}

int main() {
  auto ctx = std::make_shared<lio_b200::Context>(0);
  if (!ctx->ok()) {
    std::fprintf(stderr, "no usable B200-class device (error %d): the library has no CPU fallback\n", ctx->last_error);
    return 2;
  }
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  std::normal_distribution<float> g(0.f, 0.01f);
  // scene: ground z = 0 and two walls x = 12, y = -9, one point per 0.5 m surface voxel
  Tree::PointVector map;
  for (int i = -40; i < 40; ++i)
    for (int j = -40; j < 40; ++j) map.push_back(make_point(0.5f * (i + u(rng)), 0.5f * (j + u(rng)), g(rng)));
  for (int j = -40; j < 40; ++j)
    for (int k = 0; k < 12; ++k) {
      map.push_back(make_point(12.f + g(rng), 0.5f * (j + u(rng)), 0.5f * (k + u(rng))));
      map.push_back(make_point(0.5f * (j + u(rng)), -9.f + g(rng), 0.5f * (k + u(rng))));
    }
  Tree ikdtree(ctx);
  ikdtree.set_downsample_param(0.5f);
  ikdtree.Build(map);
  if (ikdtree.Root_Node == nullptr || ikdtree.validnum() != (int)map.size()) return 3;

  // a scan taken at the true pose (1, 2, 1.5), yaw 0.3: fresh samples of the same surfaces in the body frame
  const double yaw = 0.3, tx = 1.0, ty = 2.0, tz = 1.5;
  const double c = std::cos(yaw), s = std::sin(yaw);
  Tree::PointVector feats_down_body;
  for (int n = 0; n < 4000; ++n) {
    float wx, wy, wz;
    const int which = n % 3;
    if (which == 0) { wx = 30.f * u(rng) - 15.f; wy = 30.f * u(rng) - 15.f; wz = g(rng); }
    else if (which == 1) { wx = 12.f + g(rng); wy = 30.f * u(rng) - 15.f; wz = 5.5f * u(rng); }
    else { wx = 30.f * u(rng) - 15.f; wy = -9.f + g(rng); wz = 5.5f * u(rng); }
    const double dx = wx - tx, dy = wy - ty, dz = wz - tz;  // body = R^T (world - t)
    feats_down_body.push_back(make_point((float)(c * dx + s * dy), (float)(-s * dx + c * dy), (float)dz));
  }

  lio_b200::esekf kf;
  lio_state x = kf.get_x();
  x.pos[0] = tx + 0.04;  // 5 cm / 1 degree perturbed prior
  x.pos[1] = ty - 0.03;
  x.pos[2] = tz + 0.02;
  x.rot[0] = std::cos((yaw + 0.0175) / 2);
  x.rot[3] = std::sin((yaw + 0.0175) / 2);
  kf.change_x(x);
  // the reference's call, argument for argument (laserMapping.cpp:771-774): a cloud pointer and the Nearest_Points
  // container that map_incremental reads afterwards
  struct Cloud {
    Tree::PointVector points;
  };
  auto feats_down = std::make_shared<Cloud>();
  feats_down->points = feats_down_body;
  std::vector<Tree::PointVector> Nearest_Points;
  kf.update_iterated_dyn_share_modified(0.001 /*LASER_POINT_COV*/, feats_down, ikdtree, Nearest_Points, 4, false);
  if (kf.last_error != LIO_OK) return 4;
  size_t with5 = 0;
  for (const auto& v : Nearest_Points) with5 += v.size() == 5;
  if (Nearest_Points.size() != feats_down_body.size() || with5 < (size_t)kf.effct_feat_num) return 5;
  const lio_state xs = kf.get_x();
  const double ex = xs.pos[0] - tx, ey = xs.pos[1] - ty, ez = xs.pos[2] - tz;
  const double eyaw = 2.0 * std::atan2(xs.rot[3], xs.rot[0]) - yaw;
  std::printf("matched %d points in %d passes; pose error (%.4f, %.4f, %.4f) m, yaw %.5f rad\n", kf.effct_feat_num,
              kf.n_passes, ex, ey, ez, eyaw);

  // map_incremental's two calls (laserMapping.cpp:430-431) through the tree facade
  Tree::PointVector to_add(feats_down_body.begin(), feats_down_body.begin() + 100);
  for (auto& p : to_add) {  // world frame
    const double bx = p.x, by = p.y;
    p.x = (float)(c * bx - s * by + tx);
    p.y = (float)(s * bx + c * by + ty);
    p.z = (float)(p.z + tz);
  }
  const int added = ikdtree.Add_Points(to_add, true);
  Tree::PointVector nearest;
  std::vector<float> dist;
  ikdtree.Nearest_Search(to_add[0], 5, nearest, dist);
  std::printf("Add_Points kept %d of %zu; map holds %d points; 5-NN of the first one: %zu found, d2[0] = %.4g\n", added,
              to_add.size(), ikdtree.validnum(), nearest.size(), dist.empty() ? -1.0 : dist[0]);
  const bool ok = std::fabs(ex) < 0.01 && std::fabs(ey) < 0.01 && std::fabs(ez) < 0.01 && std::fabs(eyaw) < 2e-3 &&
                  nearest.size() == 5;
  return ok ? 0 : 1;
}
