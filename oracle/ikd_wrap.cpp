// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this.
//
// Thin extern "C" wrapper around the REFERENCE's own ikd-Tree, compiled in place
// from /root/reference/src/S-FAST_LIO/include/ikd-Tree/ikd_Tree.{h,cpp} (sources are
// never copied into this repo; see oracle/Makefile, target _ref/libikd_ref.so).
// It exposes exactly the calls the hot path makes on the tree:
//   Build            ikd_Tree.cpp:355-367   (laserMapping.cpp:756)
//   Nearest_Search   ikd_Tree.cpp:370-402   (esekfom.hpp:140)
//   Add_Points       ikd_Tree.cpp:419-512   (laserMapping.cpp:430-431)
//   Delete_Point_Boxes ikd_Tree.cpp:559-579 (laserMapping.cpp:361-364)
//   flatten          ikd_Tree.cpp:1490-1516
// Point ids ride in the bits of `normal_x`, which the tree copies through verbatim.
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

#include <ikd-Tree/ikd_Tree.h>

using PointT = pcl::PointXYZINormal;
using Tree = KD_TREE<PointT>;
using PointVector = Tree::PointVector;

static inline PointT make_point(const float* xyz, int32_t id) {
  PointT p;
  std::memset(&p, 0, sizeof(p));
  p.x = xyz[0];
  p.y = xyz[1];
  p.z = xyz[2];
  std::memcpy(&p.normal_x, &id, 4);
  return p;
}
static inline int32_t point_id(const PointT& p) {
  int32_t id;
  std::memcpy(&id, &p.normal_x, 4);
  return id;
}

extern "C" {

void* ikd_create(float delete_param, float balance_param, float box_length) {
  // sizeof(KD_TREE<PointXYZINormal>) is ~80 MB (op-log ring, ikd_Tree.h:18,173) -> heap.
  return new Tree(delete_param, balance_param, box_length);
}
void ikd_destroy(void* h) { delete static_cast<Tree*>(h); }
void ikd_set_downsample(void* h, float ds) { static_cast<Tree*>(h)->set_downsample_param(ds); }
int ikd_size(void* h) { return static_cast<Tree*>(h)->size(); }
int ikd_validnum(void* h) { return static_cast<Tree*>(h)->validnum(); }
int ikd_empty(void* h) { return static_cast<Tree*>(h)->Root_Node == nullptr; }

void ikd_build(void* h, const float* xyz, const int32_t* ids, int64_t n) {
  PointVector pts((size_t)n);
  for (int64_t i = 0; i < n; ++i) pts[i] = make_point(xyz + 3 * i, ids ? ids[i] : (int32_t)i);
  static_cast<Tree*>(h)->Build(pts);
}

// Batch of Nearest_Search calls (the reference issues them from an OpenMP loop of
// MP_PROC_NUM threads, esekfom.hpp:114-117).  Outputs are in the tree's own order
// (ascending distance, ties traversal-dependent); -1 / +inf pad when fewer than k found.
void ikd_knn(void* h, const float* q, int64_t m, int k, float max_dist, int32_t* idx, float* d2,
             float* nbr_xyz, int nthreads) {
  Tree* t = static_cast<Tree*>(h);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    PointT p = make_point(q + 3 * i, -1);
    PointVector near;
    std::vector<float> dist;
    t->Nearest_Search(p, k, near, dist, max_dist);
    for (int j = 0; j < k; ++j) {
      bool ok = j < (int)near.size();
      if (idx) idx[i * k + j] = ok ? point_id(near[j]) : -1;
      if (d2) d2[i * k + j] = ok ? dist[j] : INFINITY;
      if (nbr_xyz) {
        nbr_xyz[(i * k + j) * 3 + 0] = ok ? near[j].x : 0.f;
        nbr_xyz[(i * k + j) * 3 + 1] = ok ? near[j].y : 0.f;
        nbr_xyz[(i * k + j) * 3 + 2] = ok ? near[j].z : 0.f;
      }
    }
  }
}

int ikd_add_points(void* h, const float* xyz, const int32_t* ids, int64_t n, int downsample_on) {
  PointVector pts((size_t)n);
  for (int64_t i = 0; i < n; ++i) pts[i] = make_point(xyz + 3 * i, ids ? ids[i] : -1);
  return static_cast<Tree*>(h)->Add_Points(pts, downsample_on != 0);
}

int ikd_delete_boxes(void* h, const float* boxes6, int nb) {
  std::vector<BoxPointType> boxes((size_t)nb);
  for (int b = 0; b < nb; ++b)
    for (int a = 0; a < 3; ++a) {
      boxes[b].vertex_min[a] = boxes6[6 * b + a];
      boxes[b].vertex_max[a] = boxes6[6 * b + 3 + a];
    }
  return static_cast<Tree*>(h)->Delete_Point_Boxes(boxes);
}

// kNN back-end callback with the signature oracle/lio_oracle.cpp expects (orc::knn_fn):
// the reference's unbounded Nearest_Search (esekfom.hpp:140 passes no max_dist), canonicalised to
// ascending (d2, id), then cut at d2 <= max_d2.  out_pts: k x 4 floats (x,y,z,id bits).
int ikd_knn_cb(void* h, const float* q, int k, float max_d2, float* out_pts, float* out_d2) {
  Tree* t = static_cast<Tree*>(h);
  PointT p = make_point(q, -1);
  PointVector near;
  std::vector<float> dist;
  t->Nearest_Search(p, k, near, dist);
  struct C { float d2; int32_t id; int j; };
  C c[16];
  int n = (int)near.size();
  if (n > 16) n = 16;
  for (int j = 0; j < n; ++j) c[j] = C{dist[j], point_id(near[j]), j};
  for (int a = 1; a < n; ++a)
    for (int b = a; b > 0 && (c[b].d2 < c[b - 1].d2 || (c[b].d2 == c[b - 1].d2 && c[b].id < c[b - 1].id)); --b) {
      C tmp = c[b]; c[b] = c[b - 1]; c[b - 1] = tmp;
    }
  int cnt = 0;
  for (int j = 0; j < n; ++j) {
    if (!(c[j].d2 <= max_d2)) break;
    const PointT& np = near[c[j].j];
    out_pts[4 * cnt] = np.x;
    out_pts[4 * cnt + 1] = np.y;
    out_pts[4 * cnt + 2] = np.z;
    std::memcpy(&out_pts[4 * cnt + 3], &c[j].id, 4);
    out_d2[cnt] = c[j].d2;
    ++cnt;
  }
  return cnt;
}
void* ikd_knn_callback() { return (void*)&ikd_knn_cb; }

// Live points, pre-order DFS (ikd_Tree.cpp:1490-1516).  Returns the count; fills up to cap.
int64_t ikd_flatten(void* h, float* xyz_out, int32_t* ids_out, int64_t cap) {
  Tree* t = static_cast<Tree*>(h);
  PointVector st;
  if (t->Root_Node != nullptr) t->flatten(t->Root_Node, st, NOT_RECORD);
  int64_t n = (int64_t)st.size();
  for (int64_t i = 0; i < n && i < cap; ++i) {
    if (xyz_out) {
      xyz_out[3 * i + 0] = st[i].x;
      xyz_out[3 * i + 1] = st[i].y;
      xyz_out[3 * i + 2] = st[i].z;
    }
    if (ids_out) ids_out[i] = point_id(st[i]);
  }
  return n;
}

}  // extern "C"
