// Scan preprocessing:
//   ImuProcess::UndistortPcl back half (src/IMU_Processing.hpp:361-401): per-point motion compensation, FP64
//   pcl::VoxelGrid<PointType>::filter (src/laserMapping.cpp:737-738, leaf :683): centroid per occupied leaf
// The raw scan is read once: each point is compensated in registers, its voxel index is taken on the compensated
// coordinates and the index bounds are reduced on the fly (one kernel).  The voxel filter then follows PCL's own
// algorithm: linear leaf index relative to the cloud minimum, STABLE sort of (index, point)
// pairs, one centroid per run of equal indices — with the FP32 sums taken in ascending point order, the order the
// oracle defines (PCL's std::sort is unstable, so PCL itself leaves it open).  Sequential FP32 sums in a defined
// order make the centroids BIT-EXACT against the oracle, and run-to-run deterministic, which the downstream
// neighbour sets and validity gates need (the filter loop amplifies 1e-9 input differences to millimetres within
// ten scans).  The only library call on the scan path is cub::DeviceRadixSort (the sensor decoders also
// use cub::DeviceSelect / DeviceScan).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "lio_ctx.cuh"

namespace lio {

#define MAX_POSES 128

struct PrepArgs {
  const float4* raw;      // x,y,z,t_ms
  int n;
  const lio_pose6d* poses;
  int n_poses;
  StateD end;
  float inv_leaf;
  float4* undist;         // n (x,y,z,t_ms)
  int* vkeys;             // n x 3 absolute voxel indices
  int* counters;          // [0] M  [1..3] key min  [4..6] key max  [7] error
};

__global__ void __launch_bounds__(256) undistort_key_kernel(const PrepArgs a) {
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
      // head = last pose with offset_time < t among poses[0 .. n_poses-2]: the first segment the reference's walk from the
      // back (:361-398) compensates the point with.  poses[1 ..] carry IMU stamps and ascend; poses[0] is the previous
      // scan end with offset_time hard-coded to 0.0 (:273-276) and may exceed poses[1] when an IMU sample falls between
      // the two scans (A.9) -- so the search runs over [1, n-2] and falls back to 0.
      int lo = 1, hi = a.n_poses - 2, head = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (s_pose[mid].offset_time < t) {
          head = mid;
          lo = mid + 1;
        } else {
          hi = mid - 1;
        }
      }
      if (head < 0 && s_pose[0].offset_time < t) head = 0;
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
    a.undist[i] = make_float4(ox, oy, oz, r.w);
    // voxel index: floor(p * inverse_leaf_size) in FP32, as PCL computes it
    kx = (int)floorf(ox * a.inv_leaf);
    ky = (int)floorf(oy * a.inv_leaf);
    kz = (int)floorf(oz * a.inv_leaf);
    a.vkeys[3 * i] = kx;
    a.vkeys[3 * i + 1] = ky;
    a.vkeys[3 * i + 2] = kz;
  }
  // voxel index bounds (PCL: min_b / max_b from getMinMax3D): warp-reduce, block-reduce, 6 atomics per block
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
  // Six words of one L2 sector are the target: 4,096 warps sending their own atomics (or even just reading the words to
  // filter them) queue up on that sector -- 75 % of this kernel's samples at 131 k points.  So the block reduces first
  // and six of its threads send one atomic each: 3 k requests instead of 25 k.
  __shared__ int s_red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_red[0][warp] = mnx;
    s_red[1][warp] = mny;
    s_red[2][warp] = mnz;
    s_red[3][warp] = mxx;
    s_red[4][warp] = mxy;
    s_red[5][warp] = mxz;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_min = threadIdx.x < 3;
    int v = s_red[threadIdx.x][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) v = is_min ? min(v, s_red[threadIdx.x][w]) : max(v, s_red[threadIdx.x][w]);
    if (is_min) {
      if (v != big) atomicMin(&a.counters[1 + threadIdx.x], v);
    } else {
      if (v != -big) atomicMax(&a.counters[1 + threadIdx.x], v);
    }
  }
}

// working counters of the first scan; afterwards centroid_kernel leaves them reset
__global__ void prep_reset_kernel(int* counters) {
  counters[0] = 0;
  counters[1] = counters[2] = counters[3] = 0x7fffffff;
  counters[4] = counters[5] = counters[6] = -0x7fffffff;
  counters[7] = 0;
}

// PCL's linear leaf index ijk . (1, div_x, div_x div_y) relative to min_b; values = point index
__global__ void linear_index_kernel(const int* vkeys, int n, int* counters, uint32_t* keys, uint32_t* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long dx = (long long)counters[4] - counters[1] + 1, dy = (long long)counters[5] - counters[2] + 1,
                  dz = (long long)counters[6] - counters[3] + 1;
  if (dx * dy * dz > 0x7fffffffLL) {  // "leaf size too small": PCL returns the input unchanged
    if (i == 0) counters[7] = 3;
    keys[i] = 0;
    vals[i] = i;
    return;
  }
  const long long lin = (long long)(vkeys[3 * i] - counters[1]) +
                        dx * ((long long)(vkeys[3 * i + 1] - counters[2]) + dy * (long long)(vkeys[3 * i + 2] - counters[3]));
  keys[i] = (uint32_t)lin;
  vals[i] = (uint32_t)i;
}

// Run heads + gather in one pass over the sorted order (what cub::DeviceSelect + a gather kernel did in three launches).
// A tile of RT sorted positions per block: head flags (leaf index differs from its predecessor's), block scan, then the
// tile's exclusive prefix by a decoupled look-back over the tiles before it -- each tile publishes {launch tag, state,
// count} in one 64-bit word as soon as it knows its own count, a warp sums the published counts of up to 32 predecessors
// per round until it meets one that already carries its inclusive prefix.  Tiles take their number from a ticket, so a
// tile only ever waits for tiles that have already started.  Nothing is reset by the host between launches: words are
// recognised by the launch tag, and the tile with the highest ticket puts the ticket counter back to zero.
constexpr int RT = 2048;
__global__ void __launch_bounds__(256) runs_gather_kernel(const uint32_t* keys, const uint32_t* vals, int n,
                                                          const float4* undist, const float* aux, float4* sorted_pts,
                                                          float* sorted_aux, int* heads, int* n_runs,
                                                          unsigned long long* status, unsigned* ticket,
                                                          unsigned tag, int* stalled) {
  __shared__ int s_tile, s_prefix, s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int j0 = tile * RT + tid * 8;
  uint32_t k[8], prev = 0;
  bool f[8];
  int cnt = 0;
  if (j0 < n) {
    if (j0 + 8 <= n) {
      const uint4 a = *reinterpret_cast<const uint4*>(keys + j0), b = *reinterpret_cast<const uint4*>(keys + j0 + 4);
      k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) k[u] = j0 + u < n ? keys[j0 + u] : 0u;
    }
    if (j0 > 0) prev = keys[j0 - 1];
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = j0 + u;
    f[u] = j < n && (j == 0 || k[u] != (u ? k[u - 1] : prev));
    cnt += f[u];
  }
  // block-exclusive scan of the per-thread counts
  int inc = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int warp_base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < warp) warp_base += s_warp[w];
    total += s_warp[w];
  }
  int local = warp_base + inc - cnt;
  // word = tag << 32 | state << 30 | count, state 1 = the tile's own count, 2 = inclusive prefix.  The own count goes out
  // first, then the tile does its gather (independent of the prefix) while the other tiles publish theirs.
  const unsigned long long mine = ((unsigned long long)tag << 32);
  if (tid == 0)
    *reinterpret_cast<volatile unsigned long long*>(status + tile) =
        mine | ((tile == 0 ? 2ull : 1ull) << 30) | (unsigned)total;
  // gather into sorted order: the runs become contiguous, so the sequential sums of centroid_kernel stream memory
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = j0 + u;
    if (j < n) {
      const uint32_t i = vals[j];
      sorted_pts[j] = __ldg(undist + i);
      if (aux) sorted_aux[j] = __ldg(aux + i);
    }
  }
  // look-back (warp 0)
  if (warp == 0) {
    int prefix = 0;
    for (int p = tile - 1;; p -= 32) {
      const int idx = p - lane;
      unsigned st = 2, val = 0;  // before tile 0: an inclusive prefix of nothing
      if (idx >= 0) {
        unsigned long long w;
        unsigned spins = 0;
        do {
          w = *reinterpret_cast<volatile unsigned long long*>(status + idx);
        } while (((unsigned)(w >> 32) != tag || (((unsigned)w >> 30) & 3u) == 0) && ++spins < (1u << 20));
        if (spins >= (1u << 20)) {  // cannot happen with a healthy launch; never hang the device over it
          *stalled = 1;
          w = ((unsigned long long)tag << 32) | (2ull << 30);
        }
        st = ((unsigned)w >> 30) & 3u;
        val = (unsigned)w & 0x3fffffffu;
      }
      const unsigned pm = __ballot_sync(0xffffffffu, st == 2);
      const int first = pm ? __ffs(pm) - 1 : 31;  // nearest predecessor that already knows its inclusive prefix
      int v = lane <= first ? (int)val : 0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      prefix += v;
      if (pm) break;
    }
    if (lane == 0) {
      if (tile > 0)
        *reinterpret_cast<volatile unsigned long long*>(status + tile) = mine | (2ull << 30) | (unsigned)(prefix + total);
      s_prefix = prefix;
    }
  }
  __syncthreads();
  const int base = s_prefix;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (f[u]) heads[base + local++] = j0 + u;
  if (tid == 0 && tile == (n + RT - 1) / RT - 1) {
    *n_runs = base + total;
    *ticket = 0;  // every ticket of this launch has been handed out (this is the highest): ready for the next launch
  }
}

// One run of equal leaf indices per thread: FP32 sums in sorted (= ascending point) order, then / count.  The order
// is sequential by definition (bit parity with the oracle), so a run cannot be split; what can be done is to keep its
// adds fed and to keep long runs out of each other's way.  Runs are dealt to the threads warp-first (run m -> warp
// m % n_warps, lane m / n_warps): neighbouring leaves -- a wall next to the sensor puts thousands of points into each of
// a few dozen consecutive leaves -- land in different warps instead of queueing up in one.  Short runs: the owning
// thread streams its points eight loads at a time.  Long runs: the WARP takes them one after the other -- all lanes
// stage 256 points into shared memory with coalesced loads (the next stage's loads are in flight during the sums), then
// lanes 0-4 run the five component sums (x, y, z, time, intensity) as five independent sequential chains.
constexpr int LONG_RUN = 96;
constexpr int STAGE = 256;
__global__ void __launch_bounds__(128) centroid_kernel(const float4* sorted_pts, const float* sorted_aux,
                                                       const int* heads, const int* n_runs, int n, int max_m,
                                                       float4* body, float* body_time, int* scan_m, int* counters) {
  __shared__ __align__(16) float s_stage[4][5][STAGE];
  const int Mtot = *n_runs;
  const int M = Mtot > max_m ? max_m : Mtot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int m = lane * n_warps + blockIdx.x * (blockDim.x >> 5) + warp;
  if (m == 0) {
    // the last kernel of the preprocessing files this scan's counters as the report the host reads (counters[16..23])
    // and leaves the working set reset for the next scan -- what a one-thread reset kernel used to do before every scan
    int err = counters[7];
    if (counters[14]) {  // runs_gather_kernel gave up waiting for a tile: the run heads are not to be trusted
      err = 4;
      counters[14] = 0;
    }
    *scan_m = err >= 3 ? 0 : M;
    if (Mtot > max_m && err == 0) err = 2;  // more voxels than lio_caps.max_down_points
    counters[16] = Mtot;
#pragma unroll
    for (int k = 1; k < 7; ++k) counters[16 + k] = counters[k];
    counters[23] = err;
    counters[0] = 0;
    counters[1] = counters[2] = counters[3] = 0x7fffffff;
    counters[4] = counters[5] = counters[6] = -0x7fffffff;
    counters[7] = 0;
  }
  int beg = 0, end = 0;
  if (m < M) {
    beg = heads[m];
    end = (m + 1 < Mtot) ? heads[m + 1] : n;
  }
  const bool is_long = (end - beg) >= LONG_RUN;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f, st = 0.f;
  if (m < M && !is_long) {
    int j = beg;
    for (; j + 8 <= end; j += 8) {
      float4 p[8];
      float q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p[u] = __ldg(sorted_pts + j + u);
        q[u] = sorted_aux ? __ldg(sorted_aux + j + u) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        sx = sx + p[u].x;
        sy = sy + p[u].y;
        sz = sz + p[u].z;
        st = st + p[u].w;
        si = si + q[u];
      }
    }
    for (; j < end; ++j) {
      const float4 p = __ldg(sorted_pts + j);
      sx = sx + p.x;
      sy = sy + p.y;
      sz = sz + p.z;
      st = st + p.w;
      if (sorted_aux) si = si + __ldg(sorted_aux + j);
    }
  }
  // long runs of this warp, one at a time
  unsigned todo = __ballot_sync(0xffffffffu, m < M && is_long);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int rb = __shfl_sync(0xffffffffu, beg, src), re = __shfl_sync(0xffffffffu, end, src);
    float acc = 0.f;  // lane c < 5: running sum of component c
    float4 pre[STAGE / 32];
    float preq[STAGE / 32];
#pragma unroll
    for (int u = 0; u < STAGE / 32; ++u) {
      const int j = rb + lane + 32 * u;
      pre[u] = j < re ? __ldg(sorted_pts + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      preq[u] = (sorted_aux && j < re) ? __ldg(sorted_aux + j) : 0.f;
    }
    for (int base = rb; base < re; base += STAGE) {
      const int cnt = min(STAGE, re - base);
      __syncwarp();
#pragma unroll
      for (int u = 0; u < STAGE / 32; ++u) {
        const int t = lane + 32 * u;
        s_stage[warp][0][t] = pre[u].x;
        s_stage[warp][1][t] = pre[u].y;
        s_stage[warp][2][t] = pre[u].z;
        s_stage[warp][3][t] = pre[u].w;
        s_stage[warp][4][t] = preq[u];
      }
      __syncwarp();
      if (base + STAGE < re) {
#pragma unroll
        for (int u = 0; u < STAGE / 32; ++u) {
          const int j = base + STAGE + lane + 32 * u;
          pre[u] = j < re ? __ldg(sorted_pts + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          preq[u] = (sorted_aux && j < re) ? __ldg(sorted_aux + j) : 0.f;
        }
      }
      if (lane < 5) {
        const float* v = s_stage[warp][lane];
        int t = 0;
        for (; t + 16 <= cnt; t += 16) {
          const float4 w0 = *reinterpret_cast<const float4*>(v + t), w1 = *reinterpret_cast<const float4*>(v + t + 4),
                       w2 = *reinterpret_cast<const float4*>(v + t + 8), w3 = *reinterpret_cast<const float4*>(v + t + 12);
          acc = acc + w0.x; acc = acc + w0.y; acc = acc + w0.z; acc = acc + w0.w;
          acc = acc + w1.x; acc = acc + w1.y; acc = acc + w1.z; acc = acc + w1.w;
          acc = acc + w2.x; acc = acc + w2.y; acc = acc + w2.z; acc = acc + w2.w;
          acc = acc + w3.x; acc = acc + w3.y; acc = acc + w3.z; acc = acc + w3.w;
        }
        for (; t < cnt; ++t) acc = acc + v[t];
      }
    }
    const float rx = __shfl_sync(0xffffffffu, acc, 0), ry = __shfl_sync(0xffffffffu, acc, 1),
                rz = __shfl_sync(0xffffffffu, acc, 2), rt = __shfl_sync(0xffffffffu, acc, 3),
                ri = __shfl_sync(0xffffffffu, acc, 4);
    if (lane == src) {
      sx = rx;
      sy = ry;
      sz = rz;
      st = rt;
      si = ri;
    }
  }
  if (m >= M) return;
  const float cnt = (float)(end - beg);
  body[m] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
  if (body_time) body_time[m] = st / cnt;
}

// ---------------------------------------------------------------------------------------------------------
// PointCloud2 decoding (src/preprocess.cpp, oust64_handler :243-268 / velodyne_handler :380-428 with point times)
// ---------------------------------------------------------------------------------------------------------
// avia_handler (:160-183, feature extraction off).  A record is `valid` when its line is below N_SCANS and its tag says
// single / strongest return; valid_num counts the valid records from index 1 on, every point_filter_num-th of them is
// written into pl_full[i] and kept when it differs from pl_full[i - 1] -- which is the previous RECORD if that one was
// written too, else the zero-initialised point of pl_full.resize() -- by more than 1e-7 in x or in y, or in z while
// outside the blind zone (the reference's `a || b || c && d`).
__device__ __forceinline__ bool avia_valid(const unsigned char* data, const lio_cloud_layout& L, int i) {
  if (i < 1) return false;  // the handler's loops start at 1
  const unsigned char* r = data + (size_t)i * L.point_step;
  const unsigned line = r[L.off_ring], tag = r[L.off_tag];
  return line < (unsigned)L.n_scans && ((tag & 0x30) == 0x10 || (tag & 0x30) == 0x00);
}
struct AviaValid {
  const unsigned char* data;
  lio_cloud_layout L;
  __device__ __forceinline__ int operator()(const int& i) const { return avia_valid(data, L, i) ? 1 : 0; }
};

struct DecodeKeep {
  const unsigned char* data;
  lio_cloud_layout L;
  const int* valid_num;  // rule 3: inclusive count of valid records
  const float* tms;      // yaw-derived times per record (< 0: first point of its ring, dropped), or nullptr
  __device__ __forceinline__ bool operator()(const int& i) const {
    const unsigned char* r = data + (size_t)i * L.point_step;
    float x, y, z;
    memcpy(&x, r + L.off_x, 4);
    memcpy(&y, r + L.off_y, 4);
    memcpy(&z, r + L.off_z, 4);
    // float products and sums (float * float stays float in C++), then compared in double with the double `blind`
    const double range = (double)__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    const double b2 = __dmul_rn(L.blind, L.blind);
    if (L.rule == 3) {
      if (!avia_valid(data, L, i) || valid_num[i] % L.point_filter_num != 0) return false;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (avia_valid(data, L, i - 1) && valid_num[i - 1] % L.point_filter_num == 0) {
        memcpy(&px, r - L.point_step + L.off_x, 4);
        memcpy(&py, r - L.point_step + L.off_y, 4);
        memcpy(&pz, r - L.point_step + L.off_z, 4);
      }
      return (double)fabsf(__fsub_rn(x, px)) > 1e-7 || (double)fabsf(__fsub_rn(y, py)) > 1e-7 ||
             ((double)fabsf(__fsub_rn(z, pz)) > 1e-7 && range > b2);
    }
    if (tms && tms[i] < 0.f) return false;  // first point of its ring: `continue` before the decimation test
    if (i % L.point_filter_num != 0) return false;
    if (L.rule == 1) return !(range < b2);  // oust64_handler: `if (range < (blind * blind)) continue;`
    return range > b2;                      // velodyne_handler / rs_handler: `if (x*x + y*y + z*z > (blind * blind)) push_back`
  }
};

__global__ void decode_gather_kernel(const unsigned char* data, lio_cloud_layout L, const int* kept, const int* n_kept,
                                     const float* tms, float4* raw, float* aux) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *n_kept) return;
  const unsigned char* r = data + (size_t)kept[j] * L.point_step;
  float x, y, z, inten = 0.f, t_ms = 0.f;
  memcpy(&x, r + L.off_x, 4);
  memcpy(&y, r + L.off_y, 4);
  memcpy(&z, r + L.off_z, 4);
  if (L.off_intensity >= 0) {
    if (L.intensity_type == 1)
      inten = (float)r[L.off_intensity];  // livox reflectivity (uint8)
    else
      memcpy(&inten, r + L.off_intensity, 4);
  }
  if (tms) {
    t_ms = tms[kept[j]];
  } else if (L.off_time >= 0) {
    if (L.rule == 3) {
      uint32_t t;
      memcpy(&t, r + L.off_time, 4);
      t_ms = __fdiv_rn((float)t, 1000000.0f);  // offset_time / float(1000000)  (:165-167)
    } else if (L.time_type == 0) {
      float t;
      memcpy(&t, r + L.off_time, 4);
      t_ms = __fmul_rn(t, L.time_scale);  // time * time_unit_scale (float * float)
    } else if (L.time_type == 1) {
      uint32_t t;
      memcpy(&t, r + L.off_time, 4);
      t_ms = __fmul_rn((float)t, L.time_scale);  // uint32 t * float time_unit_scale -> float
    } else if (L.time_type == 2) {
      double t;
      memcpy(&t, r + L.off_time, 8);
      t_ms = (float)__dmul_rn(t, (double)L.time_scale);
    } else {  // rs_handler (:880-882): (timestamp - points[0].timestamp) * 1000.0
      double t, t0;
      memcpy(&t, r + L.off_time, 8);
      memcpy(&t0, data + L.off_time, 8);
      t_ms = (float)__dmul_rn(__dsub_rn(t, t0), 1000.0);
    }
  }
  raw[j] = make_float4(x, y, z, t_ms);
  aux[j] = inten;
}

// ---- point times from the azimuth, for drivers that give none (velodyne_handler :395-421, rs_handler :886-912)
__device__ __forceinline__ unsigned record_ring(const unsigned char* r, const lio_cloud_layout& L) {
  if (L.ring_type == 1) return r[L.off_ring];
  uint16_t v;
  memcpy(&v, r + L.off_ring, 2);
  return v;
}
__global__ void ring_key_kernel(const unsigned char* data, lio_cloud_layout L, int n, uint32_t* keys, uint32_t* vals,
                                int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned ring = record_ring(data + (size_t)i * L.point_step, L);
  if (ring >= (unsigned)L.n_scans) {  // the reference indexes its per-ring vectors out of bounds here
    *err = 1;
    ring = (unsigned)L.n_scans - 1;
  }
  keys[i] = ring;
  vals[i] = (uint32_t)i;
}
// yaw of every record, in ring-sorted order: atan2(float, float) resolves to the FP32 overload (preprocess.h:6 has
// `using namespace std`), the product with 57.2957 is FP64
__global__ void yaw_sorted_kernel(const unsigned char* data, lio_cloud_layout L, const uint32_t* vals_sorted, int n,
                                  double* yaw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned char* r = data + (size_t)vals_sorted[j] * L.point_step;
  float x, y;
  memcpy(&x, r + L.off_x, 4);
  memcpy(&y, r + L.off_y, 4);
  yaw[j] = __dmul_rn((double)atan2f(y, x), 57.2957);
}
// One thread per ring walks its records in message order: the first fixes yaw_fp and is dropped, every later one gets
// (yaw_fp - yaw [+ 360]) / omega_l, plus one revolution when that falls behind the ring's previous time.
__global__ void yaw_chain_kernel(const uint32_t* keys_sorted, const uint32_t* vals_sorted, const double* yaw, int n,
                                 int n_scans, double omega_l, float* tms) {
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (ring >= n_scans) return;
  int lo = 0, hi = n;
  while (lo < hi) {  // first record of this ring
    const int mid = (lo + hi) >> 1;
    if (keys_sorted[mid] < (uint32_t)ring) lo = mid + 1; else hi = mid;
  }
  const int beg = lo;
  hi = n;
  while (lo < hi) {  // first record of the next ring
    const int mid = (lo + hi) >> 1;
    if (keys_sorted[mid] <= (uint32_t)ring) lo = mid + 1; else hi = mid;
  }
  const int end = lo;
  if (beg >= end) return;
  const double yaw_fp = yaw[beg];
  const double rev = __ddiv_rn(360.0, omega_l);
  tms[vals_sorted[beg]] = -1.f;
  float time_last = 0.f;
  for (int j = beg + 1; j < end; ++j) {
    const double ya = yaw[j];
    float cur = ya <= yaw_fp ? (float)__ddiv_rn(__dsub_rn(yaw_fp, ya), omega_l)
                             : (float)__ddiv_rn(__dadd_rn(__dsub_rn(yaw_fp, ya), 360.0), omega_l);
    if (cur < time_last) cur = (float)__dadd_rn((double)cur, rev);
    time_last = cur;
    tms[vals_sorted[j]] = cur;
  }
}

// data already in c->d_cloud; leaves the decoded cloud in c->d_raw / c->d_raw_aux and its size in *n_out (host sync).
// yaw_times: the driver gave no point times (the handlers' given_offset_time == false).
int decode_cloud2(lio_ctx* c, int64_t n, const lio_cloud_layout& L, bool yaw_times, int64_t* n_out) {
  int* kept = reinterpret_cast<int*>(c->d_sort_vals_in);
  int* n_kept = c->d_prep_counters + 12;
  int* d_err = c->d_prep_counters + 13;
  *n_out = 0;
  if (n <= 0) {
    LIO_CHECK(c, cudaMemsetAsync(n_kept, 0, sizeof(int), c->stream));
    return LIO_OK;
  }
  const int grid = (int)((n + 255) / 256);
  LIO_CHECK(c, cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  int* valid_num = nullptr;
  float* tms = nullptr;
  cub::CountingInputIterator<int> it(0);
  if (L.rule == 3) {
    valid_num = reinterpret_cast<int*>(c->d_sort_vals_out);
    AviaValid av{c->d_cloud, L};
    cub::TransformInputIterator<int, AviaValid, cub::CountingInputIterator<int>> flags(it, av);
    size_t bytes = c->cub_tmp_bytes;
    LIO_CHECK(c, cub::DeviceScan::InclusiveSum(c->d_cub_tmp, bytes, flags, valid_num, (int)n, c->stream));
    c->launches += 2;
  } else if (yaw_times) {
    uint32_t* keys_in = c->d_sort_keys_in;
    uint32_t* keys_out = c->d_sort_keys_out;
    uint32_t* vals_in = reinterpret_cast<uint32_t*>(c->d_vkeys);
    uint32_t* vals_out = vals_in + n;
    double* yaw = reinterpret_cast<double*>(c->d_undist);
    tms = reinterpret_cast<float*>(c->d_sort_vals_out);
    ring_key_kernel<<<grid, 256, 0, c->stream>>>(c->d_cloud, L, (int)n, keys_in, vals_in, d_err);
    int bits = 1;
    while ((1 << bits) < L.n_scans) ++bits;
    size_t bytes = c->cub_tmp_bytes;
    LIO_CHECK(c, cub::DeviceRadixSort::SortPairs(c->d_cub_tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                                 bits, c->stream));  // stable: message order within a ring
    yaw_sorted_kernel<<<grid, 256, 0, c->stream>>>(c->d_cloud, L, vals_out, (int)n, yaw);
    yaw_chain_kernel<<<(L.n_scans + 31) / 32, 32, 0, c->stream>>>(keys_out, vals_out, yaw, (int)n, L.n_scans,
                                                                  0.361 * (double)L.scan_rate, tms);
    c->launches += 6;
  }
  size_t bytes = c->cub_tmp_bytes;
  DecodeKeep keep{c->d_cloud, L, valid_num, tms};
  LIO_CHECK(c, cub::DeviceSelect::If(c->d_cub_tmp, bytes, it, kept, n_kept, (int)n, keep, c->stream));
  decode_gather_kernel<<<grid, 256, 0, c->stream>>>(c->d_cloud, L, kept, n_kept, tms, c->d_raw, c->d_raw_aux);
  c->launches += 3;
  LIO_CHECK(c, cudaGetLastError());
  int h[2] = {0, 0};
  LIO_CHECK(c, cudaMemcpyAsync(h, n_kept, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  if (h[1]) {
    c->err = "a ring index is >= n_scans (the reference indexes its per-ring state out of bounds)";
    return LIO_E_INVALID;
  }
  *n_out = h[0];
  return LIO_OK;
}

int preprocess_init_counters(lio_ctx* c) {
  prep_reset_kernel<<<1, 1, 0, c->stream>>>(c->d_prep_counters);
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

size_t preprocess_sort_bytes(int64_t n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n, 0, 31);
  cub::CountingInputIterator<int> it(0);
  DecodeKeep keep{};  // the decoders' compaction (lio_scan_preprocess_cloud2)
  cub::DeviceSelect::If(nullptr, b, it, (int*)nullptr, (int*)nullptr, (int)n, keep);
  size_t d = 0;
  cub::DeviceScan::InclusiveSum(nullptr, d, (const int*)nullptr, (int*)nullptr, (int)n);
  a = a > d ? a : d;
  return a > b ? a : b;
}

int preprocess_init_tables(lio_ctx*) { return LIO_OK; }

// raw points already staged in c->d_raw (and c->d_raw_aux when has_aux), poses in c->d_poses
int preprocess(lio_ctx* c, int64_t n, int n_poses, const lio_state* end_state, float leaf, bool has_aux) {
  if (n_poses > MAX_POSES) {
    c->err = "more than 128 IMU poses in one scan";
    return LIO_E_CAPACITY;
  }
  PrepArgs a;
  a.raw = c->d_raw;
  a.n = (int)n;
  a.poses = c->d_poses;
  a.n_poses = n_poses;
  if (end_state)
    memcpy(&a.end, end_state, sizeof(lio_state));
  else
    memset(&a.end, 0, sizeof(a.end));
  a.inv_leaf = 1.0f / leaf;
  a.undist = c->d_undist;
  a.vkeys = c->d_vkeys;
  a.counters = c->d_prep_counters;
  const int grid = (int)((n + 255) / 256);
  uint32_t* keys_in = c->d_sort_keys_in;
  uint32_t* keys_out = c->d_sort_keys_out;
  int* heads = c->d_run_heads;
  int* n_runs = c->d_prep_counters + 10;
  if (n > 0) {
    undistort_key_kernel<<<grid, 256, 0, c->stream>>>(a);
    linear_index_kernel<<<grid, 256, 0, c->stream>>>(c->d_vkeys, (int)n, c->d_prep_counters, keys_in,
                                                      c->d_sort_vals_in);
    c->launches += 2;
    size_t bytes = c->cub_tmp_bytes;
    LIO_CHECK(c, cub::DeviceRadixSort::SortPairs(c->d_cub_tmp, bytes, keys_in, keys_out, c->d_sort_vals_in,
                                                 c->d_sort_vals_out, (int)n, 0, 31, c->stream));
  } else {
    LIO_CHECK(c, cudaMemsetAsync(n_runs, 0, sizeof(int), c->stream));
  }
  const int max_m = (int)c->caps.max_down_points;
  const int cgrid = (int)((std::min<int64_t>(n, max_m) + 127) / 128);
  float4* sorted_pts = c->d_raw;  // the raw scan has been consumed by now
  float* sorted_aux = has_aux ? c->d_sorted_aux : nullptr;
  if (n > 0) {
    const unsigned ntiles = (unsigned)((n + RT - 1) / RT);
    runs_gather_kernel<<<ntiles, 256, 0, c->stream>>>(keys_out, c->d_sort_vals_out, (int)n, c->d_undist,
                                                      has_aux ? c->d_raw_aux : nullptr, sorted_pts, sorted_aux, heads,
                                                      n_runs, c->d_runs_status, c->d_runs_ticket, ++c->runs_tag,
                                                      c->d_prep_counters + 14);
    c->launches++;
  }
  if (c->centroid_wait) LIO_CHECK(c, cudaStreamWaitEvent(c->stream, c->centroid_wait, 0));
  centroid_kernel<<<cgrid > 0 ? cgrid : 1, 128, 0, c->stream>>>(
      sorted_pts, sorted_aux, heads, n_runs, (int)n, max_m, c->d_body,
      reinterpret_cast<float*>(c->d_normvec) /*scratch: mean time*/, c->d_scan_m, c->d_prep_counters);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  c->scan_m = -1;  // known on the device only until someone asks
  return LIO_OK;
}

}  // namespace lio
