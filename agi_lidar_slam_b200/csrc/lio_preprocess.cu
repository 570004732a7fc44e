// Scan preprocessing in one pass over the raw scan:
//   ImuProcess::UndistortPcl back half (src/IMU_Processing.hpp:361-401): per-point motion compensation, FP64
//   pcl::VoxelGrid<PointType>::filter (src/laserMapping.cpp:737-738, leaf :683): centroid per occupied leaf
// fused: each raw point is compensated in registers and binned straight into a scan-voxel hash.  Centroids are
// accumulated as fixed-point int64 offsets from the voxel origin, so the sums are exact and independent of the
// order in which the atomics land (no run-to-run jitter in downstream neighbour sets).  Output order is PCL's:
// ascending (kz, ky, kx).  The only library call is cub::DeviceRadixSort for that final ordering of the M voxels.
#include <cub/device/device_radix_sort.cuh>

#include "lio_ctx.cuh"

namespace lio {

#define MAX_POSES 128
constexpr double FX_POS = 68719476736.0;    // 2^36 per metre (offset inside the leaf)
constexpr double FX_INT = 1048576.0;        // 2^20 per intensity unit
constexpr double FX_TIME = 4294967296.0;    // 2^32 per millisecond

struct PrepArgs {
  const float4* raw;      // x,y,z,t_ms
  const float* aux;       // intensity or nullptr
  int n;
  const lio_pose6d* poses;
  int n_poses;
  StateD end;
  float leaf, inv_leaf;
  float4* undist;         // n (x,y,z,t_ms)
  int* vkeys;             // n x 3 or nullptr
  unsigned long long* svox_key;
  long long* svox_acc;
  uint32_t* svox_cnt;
  uint32_t smask;
  int* counters;          // [0] M  [1..3] key min  [4..6] key max  [7] error
};

__global__ void __launch_bounds__(256) undistort_voxel_kernel(const PrepArgs a) {
  __shared__ lio_pose6d s_pose[MAX_POSES];
  for (int k = threadIdx.x; k < a.n_poses * 22; k += blockDim.x)
    reinterpret_cast<double*>(s_pose)[k] = reinterpret_cast<const double*>(a.poses)[k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int kx = 0, ky = 0, kz = 0;
  const bool act = i < a.n;
  if (act) {
    const float4 r = __ldg(a.raw + i);
    float ox = r.x, oy = r.y, oz = r.z;
    if (a.n_poses >= 2) {
      const double t = (double)r.w / double(1000);
      // head = last pose with offset_time < t among poses[0 .. n_poses-2]  (:361-373)
      int lo = 0, hi = a.n_poses - 2, head = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (s_pose[mid].offset_time < t) {
          head = mid;
          lo = mid + 1;
        } else {
          hi = mid - 1;
        }
      }
      if (head >= 0) {
        const lio_pose6d& H = s_pose[head];
        const lio_pose6d& T = s_pose[head + 1];
        const double dt = t - H.offset_time;
        const double wdt[3] = {T.gyr[0] * dt, T.gyr[1] * dt, T.gyr[2] * dt};
        double E[9], Ri[9], Rli[9], Rend[9];
        quat_to_mat(so3_exp(wdt), E);
        mat3_mul(H.rot, E, Ri);  // R_i = R_head * Exp(w_tail dt)   (:382)
        quat_to_mat(a.end.rli, Rli);
        quat_to_mat(a.end.rot, Rend);
        const double Pi[3] = {r.x, r.y, r.z};
        double T_ei[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
          T_ei[k] = ((H.pos[k] + H.vel[k] * dt) + ((0.5 * T.acc[k]) * dt) * dt) - a.end.pos[k];  // (:386-387)
        double v1[3], v2[3], v3[3], v4[3];
        mat3_vec(Rli, Pi, v1);
        v1[0] += a.end.tli[0];
        v1[1] += a.end.tli[1];
        v1[2] += a.end.tli[2];
        mat3_vec(Ri, v1, v2);
        v2[0] += T_ei[0];
        v2[1] += T_ei[1];
        v2[2] += T_ei[2];
        mat3T_vec(Rend, v2, v3);
        v3[0] -= a.end.tli[0];
        v3[1] -= a.end.tli[1];
        v3[2] -= a.end.tli[2];
        mat3T_vec(Rli, v3, v4);  // (:388-393)
        ox = (float)v4[0];
        oy = (float)v4[1];
        oz = (float)v4[2];
      }
    }
    if (a.undist) a.undist[i] = make_float4(ox, oy, oz, r.w);
    // voxel index: floor(p * inverse_leaf_size) in FP32, as PCL computes it
    kx = (int)floorf(ox * a.inv_leaf);
    ky = (int)floorf(oy * a.inv_leaf);
    kz = (int)floorf(oz * a.inv_leaf);
    if (a.vkeys) {
      a.vkeys[3 * i] = kx;
      a.vkeys[3 * i + 1] = ky;
      a.vkeys[3 * i + 2] = kz;
    }
    const unsigned long long key = pack_cell(kx, ky, kz);
    uint32_t h = hash64(key) & a.smask;
    uint32_t probes = 0;
    bool ok = true;
    for (;;) {
      const unsigned long long prev = atomicCAS(&a.svox_key[h], LIO_EMPTY_KEY, key);
      if (prev == LIO_EMPTY_KEY || prev == key) break;
      h = (h + 1) & a.smask;
      if (++probes > a.smask) {
        ok = false;
        atomicExch(&a.counters[7], 1);
        break;
      }
    }
    if (ok) {
      const double offx = (double)ox - (double)kx * (double)a.leaf;
      const double offy = (double)oy - (double)ky * (double)a.leaf;
      const double offz = (double)oz - (double)kz * (double)a.leaf;
      unsigned long long* acc = reinterpret_cast<unsigned long long*>(a.svox_acc + (size_t)h * 5);
      atomicAdd(acc + 0, (unsigned long long)__double2ll_rn(offx * FX_POS));
      atomicAdd(acc + 1, (unsigned long long)__double2ll_rn(offy * FX_POS));
      atomicAdd(acc + 2, (unsigned long long)__double2ll_rn(offz * FX_POS));
      if (a.aux) atomicAdd(acc + 3, (unsigned long long)__double2ll_rn((double)__ldg(a.aux + i) * FX_INT));
      atomicAdd(acc + 4, (unsigned long long)__double2ll_rn((double)r.w * FX_TIME));
      atomicAdd(&a.svox_cnt[h], 1u);
    }
  }
  // voxel index bounds for PCL's "leaf size too small" check: warp-reduce, then 6 atomics per warp
  const unsigned FULL = 0xffffffffu;
  const int big = 0x7fffffff;
  int mnx = act ? kx : big, mny = act ? ky : big, mnz = act ? kz : big;
  int mxx = act ? kx : -big, mxy = act ? ky : -big, mxz = act ? kz : -big;
  mnx = __reduce_min_sync(FULL, mnx);
  mny = __reduce_min_sync(FULL, mny);
  mnz = __reduce_min_sync(FULL, mnz);
  mxx = __reduce_max_sync(FULL, mxx);
  mxy = __reduce_max_sync(FULL, mxy);
  mxz = __reduce_max_sync(FULL, mxz);
  if ((threadIdx.x & 31) == 0 && mnx != big) {
    atomicMin(&a.counters[1], mnx);
    atomicMin(&a.counters[2], mny);
    atomicMin(&a.counters[3], mnz);
    atomicMax(&a.counters[4], mxx);
    atomicMax(&a.counters[5], mxy);
    atomicMax(&a.counters[6], mxz);
  }
}

__global__ void svox_init_kernel(unsigned long long* key, long long* acc, uint32_t* cnt, uint32_t cap) {
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += gridDim.x * blockDim.x) {
    key[h] = LIO_EMPTY_KEY;
    cnt[h] = 0;
#pragma unroll
    for (int f = 0; f < 5; ++f) acc[(size_t)h * 5 + f] = 0;
  }
}

__global__ void prep_reset_kernel(int* counters, unsigned long long* sort_keys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sort_keys[i] = ~0ull;
  if (i == 0) {
    counters[0] = 0;
    counters[1] = counters[2] = counters[3] = 0x7fffffff;
    counters[4] = counters[5] = counters[6] = -0x7fffffff;
    counters[7] = 0;
  }
}

// occupied slots -> (key, slot) list (unordered; sorted next)
__global__ void svox_compact_kernel(const unsigned long long* key, uint32_t cap, unsigned long long* out_key,
                                    uint32_t* out_slot, int* counters, int max_out) {
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += gridDim.x * blockDim.x) {
    const unsigned long long k = key[h];
    if (k != LIO_EMPTY_KEY) {
      const int j = atomicAdd(&counters[0], 1);
      if (j < max_out) {
        out_key[j] = k;
        out_slot[j] = h;
      }
    }
  }
}

// centroid per voxel in sorted order; the consumed hash slots are cleared for the next scan
__global__ void svox_finalize_kernel(const unsigned long long* sorted_key, const uint32_t* sorted_slot, int* counters,
                                     int max_m, float leaf, unsigned long long* key, long long* acc, uint32_t* cnt,
                                     float4* body, float* body_time, int* scan_m) {
  const int Mtot = counters[0];
  const int M = Mtot > max_m ? max_m : Mtot;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) {
    *scan_m = M;
    if (Mtot > max_m) counters[7] = 2;  // more voxels than lio_caps.max_down_points
  }
  if (j >= Mtot) return;
  const unsigned long long k = sorted_key[j];
  const uint32_t h = sorted_slot[j];
  const int B = 1 << 20;
  const int kx = (int)(k & 0x1FFFFF) - B, ky = (int)((k >> 21) & 0x1FFFFF) - B, kz = (int)((k >> 42) & 0x1FFFFF) - B;
  const double n = (double)cnt[h];
  const long long* s = acc + (size_t)h * 5;
  const double cx = (double)kx * (double)leaf + ((double)s[0] / n) / FX_POS;
  const double cy = (double)ky * (double)leaf + ((double)s[1] / n) / FX_POS;
  const double cz = (double)kz * (double)leaf + ((double)s[2] / n) / FX_POS;
  const double ci = ((double)s[3] / n) / FX_INT;
  const double ct = ((double)s[4] / n) / FX_TIME;
  if (j < M) {
    body[j] = make_float4((float)cx, (float)cy, (float)cz, (float)ci);
    if (body_time) body_time[j] = (float)ct;
  }
  key[h] = LIO_EMPTY_KEY;
  cnt[h] = 0;
#pragma unroll
  for (int f = 0; f < 5; ++f) acc[(size_t)h * 5 + f] = 0;
}

size_t preprocess_sort_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 63);
  return bytes;
}

int preprocess_init_tables(lio_ctx* c) {
  svox_init_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_svox_key, c->d_svox_acc, c->d_svox_cnt, c->svox_cap);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

// raw points already staged in c->d_raw (and c->d_raw_aux when has_aux), poses in c->d_poses
int preprocess(lio_ctx* c, int64_t n, int n_poses, const lio_state* end_state, float leaf, bool has_aux) {
  if (n_poses > MAX_POSES) {
    c->err = "more than 128 IMU poses in one scan";
    return LIO_E_CAPACITY;
  }
  PrepArgs a;
  a.raw = c->d_raw;
  a.aux = has_aux ? c->d_raw_aux : nullptr;
  a.n = (int)n;
  a.poses = c->d_poses;
  a.n_poses = n_poses;
  if (end_state)
    memcpy(&a.end, end_state, sizeof(lio_state));
  else
    memset(&a.end, 0, sizeof(a.end));
  a.leaf = leaf;
  a.inv_leaf = 1.0f / leaf;
  a.undist = c->d_undist;
  a.vkeys = c->d_vkeys;
  a.svox_key = c->d_svox_key;
  a.svox_acc = c->d_svox_acc;
  a.svox_cnt = c->d_svox_cnt;
  a.smask = c->svox_cap - 1;
  a.counters = c->d_prep_counters;
  const int grid = (int)((n + 255) / 256);
  prep_reset_kernel<<<grid > 0 ? grid : 1, 256, 0, c->stream>>>(c->d_prep_counters, c->d_sort_keys_in, (int)n);
  if (n > 0) undistort_voxel_kernel<<<grid, 256, 0, c->stream>>>(a);
  svox_compact_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->d_svox_key, c->svox_cap, c->d_sort_keys_in,
                                                              c->d_sort_vals_in, c->d_prep_counters, (int)n);
  c->launches += 3;
  if (n > 0) {
    size_t bytes = c->cub_tmp_bytes;
    LIO_CHECK(c, cub::DeviceRadixSort::SortPairs(c->d_cub_tmp, bytes, c->d_sort_keys_in, c->d_sort_keys_out,
                                                 c->d_sort_vals_in, c->d_sort_vals_out, (int)n, 0, 63, c->stream));
  }
  svox_finalize_kernel<<<grid > 0 ? grid : 1, 256, 0, c->stream>>>(
      c->d_sort_keys_out, c->d_sort_vals_out, c->d_prep_counters, (int)c->caps.max_down_points, leaf, c->d_svox_key,
      c->d_svox_acc, c->d_svox_cnt, c->d_body, reinterpret_cast<float*>(c->d_normvec) /*scratch: mean time*/,
      c->d_scan_m);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  c->scan_m = -1;  // known on the device only until someone asks
  return LIO_OK;
}

}  // namespace lio
