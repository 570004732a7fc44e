// Shared device/host definitions of the B200 S-FAST_LIO hot path.
// Every translation unit that includes this is compiled with -fmad=false: FP32/FP64 expressions are
// evaluated one IEEE rounding at a time in the order written (the bit-exact parts of the path — p_world,
// d2, the plane fit and its two gates — depend on it).  Fused multiply-adds are written explicitly
// (fma()) only where the result is tolerance-compared (the FP64 normal-equation accumulation).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/lio_b200.h"

#define LIO_K 5
#define LIO_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define LIO_BLOB 92  // 78 HtH + 12 Hth + n_valid + n_searched (as doubles, exact integers)

namespace lio {

// ---------------------------------------------------------------- map (voxel-hash with float4 buckets)
// One hash entry per occupied kNN cell.  16 bytes => one vector load per probe.
struct __align__(16) CellEntry {
  unsigned long long key;  // packed cell index or LIO_EMPTY_KEY
  uint32_t start;          // first slot of this cell's bucket in the point pool
  uint32_t count;          // slots in use (live + dead)
};

struct MapView {
  CellEntry* table;     // [hash_cap]
  uint32_t* cell_cap;   // [hash_cap] bucket capacity (slots)
  uint32_t* cell_pend;  // [hash_cap] appends pending in the current batch
  uint32_t* cell_base;  // [hash_cap] append base of the current batch
  float4* pool;         // [pool_cap] x,y,z, id bits (id < 0: dead slot)
  uint32_t hash_mask;
  uint32_t pool_cap;
  float4* removed;      // [removed_cap] log of the points taken out of the map (box deletes, downsample replacements)
  uint32_t removed_cap;
  uint32_t* counters;  // [0] pool_top  [1] n_cells  [2] n_live  [3] error flag  [4] scratch count  [5] slots used
                       // [6] removed points logged  [7] ... that did not fit the log
                       // [8..10] / [11..13] smallest / largest cell index (as int) that ever held a point: the box an
                       // unbounded search has to cover (warp_knn_far)
  float inv_cell;
  float cell;
};

// Files a point that leaves the map (≙ KD_TREE::Points_deleted, ikd_Tree.cpp:582-594: what acquire_removed_points hands out).
__device__ __forceinline__ void log_removed(const MapView& m, const float4& q) {
  const uint32_t k = atomicAdd(&m.counters[6], 1u);
  if (k < m.removed_cap)
    m.removed[k] = q;
  else
    atomicAdd(&m.counters[7], 1u);
}

__host__ __device__ __forceinline__ unsigned long long pack_cell(int x, int y, int z) {
  const unsigned long long B = 1u << 20;
  return (((unsigned long long)(z + (int)B) & 0x1FFFFF) << 42) | (((unsigned long long)(y + (int)B) & 0x1FFFFF) << 21) |
         ((unsigned long long)(x + (int)B) & 0x1FFFFF);
}
__host__ __device__ __forceinline__ uint32_t hash64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}
__device__ __forceinline__ int cell_coord(float v, float inv_cell) { return (int)floorf(v * inv_cell); }

// Un-fused FP32 squared distance, the expression of ikd_Tree.cpp:1539-1544 / common_lib.h:86-90.
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Read-only probe.  Returns entry index or -1.
__device__ __forceinline__ int map_find(const MapView& m, unsigned long long key, uint32_t& start, uint32_t& count) {
  uint32_t h = hash64(key) & m.hash_mask;
  for (;;) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(m.table + h));
    const unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
    if (k == key) {
      start = raw.z;
      count = raw.w;
      return (int)h;
    }
    if (k == LIO_EMPTY_KEY) {
      start = 0;
      count = 0;
      return -1;
    }
    h = (h + 1) & m.hash_mask;
  }
}

// ---------------------------------------------------------------- surf voxel filter (pcl::VoxelGrid) working set
// Hash of the occupied leaves of the current scan (key = packed absolute leaf index) with, per leaf, its point count,
// its segment in the point-index array and its output rank; the rank comes from a bitmap over the scan's bounding grid
// (one bit per leaf, a count per 1024-leaf superblock).  Everything is left clean by the last kernel of a scan.
struct VoxelFilter {
  unsigned long long* key;  // [hash_mask + 1]
  uint32_t* cnt;            // points of the leaf
  uint32_t* off;            // first entry of the leaf's segment in `seg`
  uint32_t* rank;           // output position
  uint32_t* lin;            // PCL's linear leaf index relative to the scan's minimum
  uint32_t* list;           // [hash_mask + 1] hash slots of the occupied leaves, in order of creation
  uint4* big;               // {hash slot, points, segment offset} of the leaves a whole block sums (more than VF_MID points)
  uint4* mid;               // ... a warp sums (VF_SMALL < points <= VF_MID)
  uint32_t* slot;           // [max_scan_points] hash slot of every point's leaf
  uint32_t* pos;            // [max_scan_points] the point's place in its leaf's segment (arrival order)
  uint32_t* seg;            // [max_scan_points] point indices, leaf by leaf
  uint32_t* seg2;           // [max_scan_points] the segments of the long leaves in ascending point order
  uint32_t* bitmap;         // [bitmap_bits / 32]
  uint32_t* sbcount;        // [bitmap_bits / 1024] occupied leaves per superblock
  uint32_t* sbprefix;       // ... exclusive scan
  int* ctr;                 // [0] leaves  [1] segment cursor  [2] long leaves  [3] centroid block ticket  [4] bits block
                            // ticket  [5] mid leaves  [6] mid ticket  [7] short-leaf ticket
  uint32_t hash_mask;
  unsigned long long bitmap_bits;
};

// ---------------------------------------------------------------- FP64 rotation helpers (oracle op order)
struct Quatd {
  double w, x, y, z;
};
__host__ __device__ __forceinline__ void quat_to_mat(const Quatd& q, double m[9]) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0] = 1.0 - (tyy + tzz);
  m[1] = txy - twz;
  m[2] = txz + twy;
  m[3] = txy + twz;
  m[4] = 1.0 - (txx + tzz);
  m[5] = tyz - twx;
  m[6] = txz - twy;
  m[7] = tyz + twx;
  m[8] = 1.0 - (txx + tyy);
}
// v + w*uv + qv x uv with uv = 2 (qv x v): how `Sophus::SO3 * Vector3d` rotates (esekfom.hpp:128).
__host__ __device__ __forceinline__ void quat_rotate(const Quatd& q, const double v[3], double o[3]) {
  double uv0 = q.y * v[2] - q.z * v[1];
  double uv1 = q.z * v[0] - q.x * v[2];
  double uv2 = q.x * v[1] - q.y * v[0];
  uv0 += uv0;
  uv1 += uv1;
  uv2 += uv2;
  const double c0 = q.y * uv2 - q.z * uv1;
  const double c1 = q.z * uv0 - q.x * uv2;
  const double c2 = q.x * uv1 - q.y * uv0;
  o[0] = (v[0] + q.w * uv0) + c0;
  o[1] = (v[1] + q.w * uv1) + c1;
  o[2] = (v[2] + q.w * uv2) + c2;
}
__host__ __device__ __forceinline__ void mat3_vec(const double m[9], const double v[3], double o[3]) {
  o[0] = (m[0] * v[0] + m[1] * v[1]) + m[2] * v[2];
  o[1] = (m[3] * v[0] + m[4] * v[1]) + m[5] * v[2];
  o[2] = (m[6] * v[0] + m[7] * v[1]) + m[8] * v[2];
}
__host__ __device__ __forceinline__ void mat3T_vec(const double m[9], const double v[3], double o[3]) {
  o[0] = (m[0] * v[0] + m[3] * v[1]) + m[6] * v[2];
  o[1] = (m[1] * v[0] + m[4] * v[1]) + m[7] * v[2];
  o[2] = (m[2] * v[0] + m[5] * v[1]) + m[8] * v[2];
}
__host__ __device__ __forceinline__ void mat3_mul(const double a[9], const double b[9], double o[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}
__host__ __device__ __forceinline__ void mat3T_mul(const double a[9], const double b[9], double o[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[3 * i + j] = (a[i] * b[j] + a[3 + i] * b[3 + j]) + a[6 + i] * b[6 + j];
}
__host__ __device__ __forceinline__ Quatd quat_normalized(Quatd q) {
  const double n = sqrt(((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z);
  return Quatd{q.w / n, q.x / n, q.y / n, q.z / n};
}
__host__ __device__ __noinline__ inline Quatd quat_mul(const Quatd& a, const Quatd& b) {
  Quatd r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return quat_normalized(r);
}
__host__ __device__ __noinline__ inline Quatd mat_to_quat(const double m[9]) {
  Quatd q;
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t;
    q.y = (m[2] - m[6]) * t;
    q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[3 * k + j] - m[3 * j + k]) * t;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q.x = v[0];
    q.y = v[1];
    q.z = v[2];
  }
  return quat_normalized(q);
}
// Sophus::SO3::exp / log (old, quaternion-backed)
__host__ __device__ __noinline__ inline Quatd so3_exp(const double w[3]) {
  const double theta = sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
  const double half = 0.5 * theta;
  double imag, real;
  if (theta < 1e-10) {
    const double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    real = 1.0 - 0.125 * t2 + 0.00260417 * t4;
  } else {
    imag = sin(half) / theta;
    real = cos(half);
  }
  return quat_normalized(Quatd{real, imag * w[0], imag * w[1], imag * w[2]});
}
__host__ __device__ __noinline__ inline void so3_log(const Quatd& q, double o[3]) {
  const double n = sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);
  const double w = q.w;
  double f;
  if (n < 1e-10) {
    f = 2.0 / w - 2.0 * (n * n) / (w * (w * w));
  } else if (fabs(w) < 1e-10) {
    f = (w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
  } else {
    f = 2.0 * atan(n / w) / n;
  }
  o[0] = f * q.x;
  o[1] = f * q.y;
  o[2] = f * q.z;
}

// Flat state, same layout as lio_state (26 doubles).
struct StateD {
  double pos[3];
  Quatd rot;
  Quatd rli;
  double tli[3];
  double vel[3];
  double bg[3];
  double ba[3];
  double grav[3];
};
static_assert(sizeof(StateD) == sizeof(lio_state), "state layout");

__host__ __device__ __noinline__ inline void boxplus(const StateD& x, const double f[24], StateD& r) {
  r = x;
  for (int i = 0; i < 3; ++i) r.pos[i] = x.pos[i] + f[i];
  r.rot = quat_mul(x.rot, so3_exp(f + 3));
  r.rli = quat_mul(x.rli, so3_exp(f + 6));
  for (int i = 0; i < 3; ++i) {
    r.tli[i] = x.tli[i] + f[9 + i];
    r.vel[i] = x.vel[i] + f[12 + i];
    r.bg[i] = x.bg[i] + f[15 + i];
    r.ba[i] = x.ba[i] + f[18 + i];
    r.grav[i] = x.grav[i] + f[21 + i];
  }
}
__host__ __device__ __noinline__ inline void boxminus(const StateD& x1, const StateD& x2, double o[24]) {
  double R1[9], R2[9], D[9];
  for (int i = 0; i < 3; ++i) o[i] = x1.pos[i] - x2.pos[i];
  quat_to_mat(x1.rot, R1);
  quat_to_mat(x2.rot, R2);
  mat3T_mul(R2, R1, D);
  so3_log(mat_to_quat(D), o + 3);
  quat_to_mat(x1.rli, R1);
  quat_to_mat(x2.rli, R2);
  mat3T_mul(R2, R1, D);
  so3_log(mat_to_quat(D), o + 6);
  for (int i = 0; i < 3; ++i) {
    o[9 + i] = x1.tli[i] - x2.tli[i];
    o[12 + i] = x1.vel[i] - x2.vel[i];
    o[15 + i] = x1.bg[i] - x2.bg[i];
    o[18 + i] = x1.ba[i] - x2.ba[i];
    o[21 + i] = x1.grav[i] - x2.grav[i];
  }
}

// ---------------------------------------------------------------- update-loop control block (device resident)
struct Ctrl {
  int iter;          // loop variable i of esekfom.hpp:292 (starts at -1)
  int converge;      // dyn_share.converge
  int t;             // converged-pass counter
  int done;          // loop has returned
  int n_passes;      // h_share_model calls made
  int n_valid_last;  // effct_feat_num of the last call
  int max_iter;
  int pad;
};

}  // namespace lio
