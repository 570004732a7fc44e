// Scan preprocessing:
//   ImuProcess::UndistortPcl back half (src/IMU_Processing.hpp:361-401): per-point motion compensation, FP64
//   pcl::VoxelGrid<PointType>::filter (src/laserMapping.cpp:737-738, leaf :683): centroid per occupied leaf
// The raw scan is read once: each point is compensated in registers, its voxel index is taken on the compensated
// coordinates, the index bounds are reduced on the fly and the point is entered into a hash of occupied leaves (one
// kernel).  What PCL gets from sorting all points by their linear leaf index is obtained without
// a sort of the points:
//   * the ORDER of the output (ascending leaf index = ascending (kz,ky,kx)) is a rank over the occupied leaves only: one
//     bit per leaf of the scan's bounding grid, a count per 1024-leaf superblock, an exclusive scan over the superblocks;
//     rank = superblock prefix + population count of the bits below;
//   * the SUM ORDER inside a leaf is the oracle's, ascending point index (PCL's std::sort is unstable, so PCL itself
//     leaves it open): the points of a leaf are placed into the leaf's segment in the order the atomics of the first
//     kernel gave, and the segment is put in index order right before the sequential FP32 sums -- by a sorting network in the
//     registers of one thread for up to 16 points, by a rank sort in one warp for up to 128, and for more by a bitmap over
//     the point indices in the shared memory of one block (sorting distinct integers of a bounded universe is marking
//     and counting: no comparison network), whose five sequential sums then run under the loads of the next 512 points.
// Sequential FP32 sums in a defined order make the centroids BIT-EXACT against the oracle and run-to-run deterministic,
// which the downstream neighbour sets and validity gates need (the filter loop amplifies 1e-9 input differences to
// millimetres within ten scans).  No library kernel on this path; the sensor decoders further down still use
// cub::DeviceSelect / DeviceScan / DeviceRadixSort.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "lio_ctx.cuh"

namespace lio {

#define MAX_POSES 128

// ---- the leaf hash and the rank structures (device buffers of lio_ctx, see VoxelFilter in lio_ctx.cuh)
constexpr uint32_t VF_SB_BITS = 10;     // leaves per superblock: 1024
constexpr int VF_SMALL = 16;            // leaves up to this many points are ordered and summed by one thread,
constexpr int VF_MID = 128;             // ... up to this many by one warp, longer ones by one block
constexpr int VF_WIN_WORDS = 4096;      // block path: words of the point-index bitmap (one window = 131,072 indices)
constexpr int VF_CHUNK = 896;           // block path: points staged per step of the sequential sums (224 loaders x 4)

// Optional timeline (LIO_TIMELINE=1, lio_debug_timeline): slot k of dbg[200..] takes the maximum over the stamping
// threads of the global timer in ns (the start of a phase is filed negated, so that its maximum is the earliest start).
__device__ __forceinline__ void vf_stamp(long long* dbg, int k, bool negate = false) {
  if (!dbg) return;
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  atomicMax(dbg + 200 + k, negate ? -t : t);
}

struct PrepArgs {
  const float4* raw;      // x,y,z,t_ms
  int n;
  const lio_pose6d* poses;
  int n_poses;
  StateD end;
  float inv_leaf;
  float4* undist;         // n (x,y,z,t_ms)
  int* vkeys;             // n x 3 absolute voxel indices
  int* counters;          // [0] unused  [1..3] key min  [4..6] key max  [7] error
  VoxelFilter vf;
  int max_m;
};

// One thread per raw point: motion compensation, leaf index, index bounds, and the point's leaf in the leaf hash
// (key = the packed absolute index; the threads of a warp that hit the same leaf send ONE probe and ONE count update,
// whose return value is also their place in the leaf's segment: no second round of atomics when the points are placed).
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
  // the leaf of this point: the lanes that share a leaf elect one of them to probe the hash and to add their number
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool in_range = kx > -(1 << 20) && kx < (1 << 20) && ky > -(1 << 20) && ky < (1 << 20) && kz > -(1 << 20) &&
                        kz < (1 << 20);  // 21 bits per axis in the packed key
  const unsigned long long key = (act && in_range) ? pack_cell(kx, ky, kz) : (LIO_EMPTY_KEY - 1ull - (unsigned)lane);
  const unsigned peers = __match_any_sync(FULL, key);
  const int leader = __ffs(peers) - 1;
  uint32_t h = 0xFFFFFFFFu, base = 0;
  if (act && in_range && lane == leader) {
    h = hash64(key) & a.vf.hash_mask;
    uint32_t probes = 0;
    for (;;) {
      const unsigned long long prev = atomicCAS(&a.vf.key[h], LIO_EMPTY_KEY, key);
      if (prev == LIO_EMPTY_KEY) {
        a.vf.list[atomicAdd(&a.vf.ctr[0], 1)] = h;  // a new leaf (the list is as long as the hash)
        break;
      }
      if (prev == key) break;
      h = (h + 1) & a.vf.hash_mask;
      if (++probes > a.vf.hash_mask) {
        h = 0xFFFFFFFFu;
        break;
      }
    }
    if (h != 0xFFFFFFFFu) base = atomicAdd(&a.vf.cnt[h], (uint32_t)__popc(peers));
  }
  h = __shfl_sync(FULL, h, leader);
  base = __shfl_sync(FULL, base, leader);
  if (act) {
    a.vf.slot[i] = h;
    a.vf.pos[i] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));  // arrival order; the centroid kernel orders it
    if (!in_range) atomicMax(&a.counters[7], 3);   // beyond any grid PCL could index: "leaf size too small"
    else if (h == 0xFFFFFFFFu) atomicMax(&a.counters[7], 2);  // more leaves than the hash holds (> max_down_points)
  }
  // Six words of one L2 sector are the target of the bounds: 4,096 warps sending their own atomics (or even just reading
  // the words to filter them) queue up on that sector -- 75 % of this kernel's samples at 131 k points.  So the block
  // reduces first and six of its threads send one atomic each: 3 k requests instead of 25 k.
  __shared__ int s_red[6][8];
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

// working counters of the first scan; afterwards the last kernel of the filter leaves them reset
__global__ void prep_reset_kernel(int* counters, int* vf_ctr) {
  counters[0] = 0;
  counters[1] = counters[2] = counters[3] = 0x7fffffff;
  counters[4] = counters[5] = counters[6] = -0x7fffffff;
  counters[7] = 0;
  for (int k = 0; k < 8; ++k) vf_ctr[k] = 0;
}

__device__ __forceinline__ void unpack_cell(unsigned long long key, int& x, int& y, int& z) {
  x = (int)(key & 0x1FFFFF) - (1 << 20);
  y = (int)((key >> 21) & 0x1FFFFF) - (1 << 20);
  z = (int)((key >> 42) & 0x1FFFFF) - (1 << 20);
}

// One thread per occupied leaf: PCL's linear leaf index ijk . (1, div_x, div_x div_y) relative to min_b, its bit in the
// grid bitmap, its superblock's count, its segment in the point-index array and -- by its length -- who will sum it.
// The guards: PCL's own ("leaf size too small": the index would overflow an int -> output = input, reported as
// LIO_E_VOXEL_RANGE) and this implementation's bitmap capacity.  The last block to finish turns the superblock counts
// into their exclusive scan (the grid of a scan has a few ten thousand superblocks).
__global__ void __launch_bounds__(256) voxel_bits_kernel(VoxelFilter vf, int* counters) {
  __shared__ uint32_t s_w[8];
  __shared__ int s_last;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_leaf = vf.ctr[0];
  const long long dx = (long long)counters[4] - counters[1] + 1, dy = (long long)counters[5] - counters[2] + 1,
                  dz = (long long)counters[6] - counters[3] + 1;
  if (counters[7] >= 2) return;
  if (dx * dy * dz > 0x7fffffffLL) {
    if (e == 0) counters[7] = 3;
    return;
  }
  if (dx * dy * dz > (long long)vf.bitmap_bits) {
    if (e == 0) counters[7] = 5;
    return;
  }
  if (e < n_leaf) {
    const uint32_t h = vf.list[e];
    int kx, ky, kz;
    unpack_cell(vf.key[h], kx, ky, kz);
    const uint32_t lin =
        (uint32_t)((long long)(kx - counters[1]) + dx * ((long long)(ky - counters[2]) + dy * (long long)(kz - counters[3])));
    vf.lin[h] = lin;
    atomicOr(&vf.bitmap[lin >> 5], 1u << (lin & 31));
    atomicAdd(&vf.sbcount[lin >> VF_SB_BITS], 1u);
    const uint32_t c = vf.cnt[h];
    const uint32_t off = (uint32_t)atomicAdd(&vf.ctr[1], (int)c);
    vf.off[h] = off;
    if (c > (uint32_t)VF_MID)
      vf.big[atomicAdd(&vf.ctr[2], 1)] = make_uint4(h, c, off, 0u);
    else if (c > (uint32_t)VF_SMALL)
      vf.mid[atomicAdd(&vf.ctr[5], 1)] = make_uint4(h, c, off, 0u);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&vf.ctr[4], 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // every thread scans `per` consecutive counts (a multiple of four: 16-byte loads; the counts behind the grid's last
  // superblock are zero and the array is padded, so nothing is masked)
  const int n_sb = n_leaf > 0 ? (int)((dx * dy * dz + (1 << VF_SB_BITS) - 1) >> VF_SB_BITS) : 0;
  const int per = (((n_sb + 255) / 256) + 3) & ~3;
  const uint4* src = reinterpret_cast<const uint4*>(vf.sbcount + (size_t)threadIdx.x * per);
  uint4* dst = reinterpret_cast<uint4*>(vf.sbprefix + (size_t)threadIdx.x * per);
  uint32_t sum = 0;
  for (int k = 0; k < per / 4; ++k) {
    const uint4 v = __ldcg(src + k);
    sum += (v.x + v.y) + (v.z + v.w);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
    inc += lane >= off ? t : 0u;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t run = inc - sum;
  for (int w = 0; w < warp; ++w) run += s_w[w];
  for (int k = 0; k < per / 4; ++k) {
    const uint4 v = __ldcg(src + k);
    uint4 o;
    o.x = run;
    o.y = o.x + v.x;
    o.z = o.y + v.y;
    o.w = o.z + v.z;
    run = o.w + v.w;
    dst[k] = o;
  }
}

// One thread per point: its index goes to its place in its leaf's segment.  The first threads also take one occupied
// leaf each: its output position = rank among the occupied leaves in index order (superblock prefix + bits below).
__global__ void __launch_bounds__(256) voxel_place_kernel(VoxelFilter vf, const int* counters, int n) {
  if (counters[7] >= 2) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vf.seg[vf.off[vf.slot[i]] + vf.pos[i]] = (uint32_t)i;
  if (i < vf.ctr[0]) {
    const uint32_t h = vf.list[i];
    const uint32_t lin = vf.lin[h];
    uint32_t rank = vf.sbprefix[lin >> VF_SB_BITS];
    // the superblock's 32 words are one 128-byte line: eight independent loads, the bits at and above `lin` masked off
    const uint4* row = reinterpret_cast<const uint4*>(vf.bitmap + ((lin >> VF_SB_BITS) << (VF_SB_BITS - 5)));
    const uint32_t wl = (lin >> 5) & 31u, below = (1u << (lin & 31)) - 1u;
#pragma unroll
    for (uint32_t q = 0; q < 8; ++q) {
      const uint4 v = row[q];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t wi = 4 * q + j;
        rank += (uint32_t)__popc(w[j] & (wi < wl ? 0xFFFFFFFFu : (wi == wl ? below : 0u)));
      }
    }
    vf.rank[h] = rank | (vf.cnt[h] > (uint32_t)VF_SMALL ? 0x80000000u : 0u);  // the flag: a warp or a block sums it
  }
}

// acc + v[0] + v[1] + ... + v[m - 1], strictly in that order (the FP32 chain the oracle defines).  The additions are
// the critical path (one dependent FADD per point); the shared-memory reads run one block of 16 ahead of them.  v is
// 16-byte aligned and readable up to the next multiple of 16 behind m.
__device__ __forceinline__ float chain_sum(float acc, const float* v, uint32_t m) {
  const float4* p = reinterpret_cast<const float4*>(v);
  // the buffer is readable up to the next multiple of 16 behind m and the loads run a block ahead WITHOUT a condition (a
  // predicated load waits for its predicate, i.e. for the loop counter): stale lanes are loaded and never added
  float4 c0 = p[0], c1 = p[1], c2 = p[2], c3 = p[3];
  uint32_t t = 0;
  const uint32_t last = (m + 15u) / 16u * 4u;  // float4s that may be read
  for (; t + 16 <= m; t += 16) {
    p += 4;
    const uint32_t q = min((uint32_t)(t / 4 + 4), last - 4);  // clamped: re-reads the last block instead of running over
    const float4* pn = reinterpret_cast<const float4*>(v) + q;
    const float4 n0 = pn[0], n1 = pn[1], n2 = pn[2], n3 = pn[3];
    acc = acc + c0.x; acc = acc + c0.y; acc = acc + c0.z; acc = acc + c0.w;
    acc = acc + c1.x; acc = acc + c1.y; acc = acc + c1.z; acc = acc + c1.w;
    acc = acc + c2.x; acc = acc + c2.y; acc = acc + c2.z; acc = acc + c2.w;
    acc = acc + c3.x; acc = acc + c3.y; acc = acc + c3.z; acc = acc + c3.w;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  const float r[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
#pragma unroll
  for (uint32_t k = 0; k < 16; ++k)
    if (t + k < m) acc = acc + r[k];
  return acc;
}

struct CentroidArgs {
  VoxelFilter vf;
  const float4* undist;   // x,y,z,t_ms in input order
  const float* aux;       // intensity in input order, or nullptr
  int n, max_m;
  float4* body;           // centroid x,y,z + mean intensity, in leaf-index order
  float* body_time;       // mean time
  int* scan_m;
  int* counters;
  long long* dbg;
};

// Leaves a leaf's hash entry, bit and counts as the next scan expects to find them.
__device__ __forceinline__ void leaf_cleanup(const VoxelFilter& vf, uint32_t h, bool has_lin) {
  if (has_lin) {
    const uint32_t lin = vf.lin[h];
    vf.bitmap[lin >> 5] = 0u;
    vf.sbcount[lin >> VF_SB_BITS] = 0u;
  }
  vf.key[h] = LIO_EMPTY_KEY;
  vf.cnt[h] = 0u;
}

__device__ __forceinline__ void leaf_write(const CentroidArgs& a, uint32_t h, uint32_t c, float sx, float sy, float sz,
                                           float st, float si) {
  const float fc = (float)c;
  const uint32_t r = a.vf.rank[h] & 0x7fffffffu;
  if (r < (uint32_t)a.max_m) {
    a.body[r] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
    if (a.body_time) a.body_time[r] = st / fc;
  }
}

// Centroids: every leaf's five sums are sequential FP32 chains in ascending point index.
//   long leaves (> VF_MID points), one block each, taken first: the leaf's indices are marked in a bitmap over point
//     indices in shared memory, a count over the bitmap gives every marked index its place, the ordered indices go to
//     vf.seg2; then five lanes run the five chains over points the block stages 512 at a time, the loads of the next 512
//     in flight meanwhile;
//   mid leaves, one warp each (ticket): rank of every index among the leaf's indices, points staged at their rank;
//   short leaves, one thread each (ticket per 256): the <= 16 indices through a sorting network in registers.
// Blocks that have a long leaf join the ticketed work when they are through with it.  The last block to finish files
// the scan's counters for the host and resets the working set.
constexpr int VF_STAGE_ROW = VF_CHUNK + 4;  // rows 16 bytes askew: the five chain lanes read five different bank groups
constexpr int VF_ROW = VF_MID + 4;
constexpr int VF_CENTROID_SMEM = (2 * VF_WIN_WORDS + 2 * 5 * VF_STAGE_ROW) * 4;  // dynamic: bitmap, counts, two point stages
static_assert(8 * (VF_MID + 5 * VF_ROW) <= 2 * VF_WIN_WORDS, "the warps' areas overlay the bitmap and its counts");
__global__ void __launch_bounds__(256) voxel_centroid_kernel(const CentroidArgs a) {
  extern __shared__ __align__(16) unsigned char vf_smem[];
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(vf_smem);                                           // [VF_WIN_WORDS]
  uint32_t* s_pre = s_bits + VF_WIN_WORDS;                                                            // [VF_WIN_WORDS]
  float(*s_stage)[5][VF_STAGE_ROW] = reinterpret_cast<float(*)[5][VF_STAGE_ROW]>(vf_smem + 8 * VF_WIN_WORDS);  // [2][5][.]
  __shared__ int s_flag;
  __shared__ uint32_t s_wtot[8];
  __shared__ uint32_t s_lohi[2];
  __shared__ float s_sum[5];
  const VoxelFilter& vf = a.vf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the counters of this launch, fetched together (they sit at the head of three dependent phases otherwise)
  const int4 ctr_lo = __ldcg(reinterpret_cast<const int4*>(vf.ctr)), ctr_hi = __ldcg(reinterpret_cast<const int4*>(vf.ctr) + 1);
  const int err = __ldcg(a.counters + 7);
  const bool ok = err < 2;            // 2: too many leaves, 3: PCL's overflow guard, 5: bitmap capacity
  const int n_leaf_all = ctr_lo.x;   // every leaf is in the list (and has to be cleaned up), at most max_m come out
  const int n_leaf = min(n_leaf_all, a.max_m);
  if (tid == 0) vf_stamp(a.dbg, 0, true);
  if (ok) {
    // ---- long leaves
    const int n_big = ctr_lo.z;
    for (int e = blockIdx.x; e < n_big; e += gridDim.x) {
      const uint4 leaf = __ldcg(vf.big + e);
      const uint32_t h = leaf.x, c = leaf.y, off = leaf.z;
      // span of the leaf's point indices
      uint32_t lo = 0xFFFFFFFFu, hi = 0u;
      // the first VF_KEEP x 256 indices stay in registers for the marking below: one round trip for all of them (the
      // longest leaf of an OS1-128 scan next to a wall has ~4,600 points), no reload
      constexpr int VF_KEEP = 24;
      uint32_t keep[VF_KEEP];
#pragma unroll
      for (int j = 0; j < VF_KEEP; ++j) keep[j] = __ldcg(vf.seg + off + min((uint32_t)tid + 256u * j, c - 1));  // (clamped: a repeat)
#pragma unroll
      for (int j = 0; j < VF_KEEP; ++j) {
        lo = min(lo, keep[j]);
        hi = max(hi, keep[j]);
      }
      for (uint32_t k0 = tid + 256u * VF_KEEP; k0 < c; k0 += 8 * 256) {
        uint32_t idx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = __ldcg(vf.seg + off + min(k0 + 256u * j, c - 1));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lo = min(lo, idx[j]);
          hi = max(hi, idx[j]);
        }
      }
      lo = __reduce_min_sync(0xffffffffu, lo);
      hi = __reduce_max_sync(0xffffffffu, hi);
      if (tid == 0) {
        s_lohi[0] = 0xFFFFFFFFu;
        s_lohi[1] = 0u;
      }
      __syncthreads();
      if (lane == 0) {
        atomicMin(&s_lohi[0], lo);
        atomicMax(&s_lohi[1], hi);
      }
      __syncthreads();
      const uint32_t wbase = s_lohi[0] >> 5, nwords = (s_lohi[1] >> 5) - wbase + 1;
      if (tid == 0) vf_stamp(a.dbg, 5);
      uint32_t out = 0;
      for (uint32_t w0 = 0; w0 < nwords; w0 += VF_WIN_WORDS) {
        const uint32_t nw = min((uint32_t)VF_WIN_WORDS, nwords - w0);
        for (uint32_t w = tid; w < nw; w += 256) s_bits[w] = 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < VF_KEEP; ++j) {
          const uint32_t w = (keep[j] >> 5) - wbase - w0;  // wraps to a large number below the window
          if ((uint32_t)tid + 256u * j < c && w < nw) atomicOr(&s_bits[w], 1u << (keep[j] & 31));
        }
        for (uint32_t k0 = tid + 256u * VF_KEEP; k0 < c; k0 += 8 * 256) {  // eight loads in flight, then their marks
          uint32_t idx[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) idx[j] = __ldcg(vf.seg + off + min(k0 + 256u * j, c - 1));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t w = (idx[j] >> 5) - wbase - w0;  // wraps to a large number below the window
            if (k0 + 256u * j < c && w < nw) atomicOr(&s_bits[w], 1u << (idx[j] & 31));
          }
        }
        __syncthreads();
        // exclusive count of the marks before every word: each thread counts a run of g consecutive words, the runs'
        // totals go through a block scan
        const uint32_t g = (nw + 255) / 256;
        uint32_t sum = 0;
        for (uint32_t k = 0; k < g; ++k) {
          const uint32_t w = tid * g + k;
          sum += w < nw ? (uint32_t)__popc(s_bits[w]) : 0u;
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
          inc += lane >= d ? t : 0u;
        }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t run = out + inc - sum;
        for (int w = 0; w < warp; ++w) run += s_wtot[w];
        for (uint32_t k = 0; k < g; ++k) {
          const uint32_t w = tid * g + k;
          if (w < nw) {
            s_pre[w] = run;
            run += (uint32_t)__popc(s_bits[w]);
          }
        }
        for (int w = 0; w < 8; ++w) out += s_wtot[w];
        __syncthreads();
        // the marked indices in ascending order (words interleaved over the threads: a dense stretch of the leaf is
        // spread over a whole warp)
        for (uint32_t w = tid; w < nw; w += 256) {
          uint32_t v = s_bits[w];
          uint32_t q = s_pre[w];
          const uint32_t first = (wbase + w0 + w) << 5;
          while (v) {
            vf.seg2[off + q++] = first + (uint32_t)(__ffs(v) - 1);
            v &= v - 1u;
          }
        }
        __syncthreads();
      }
      if (tid == 0) vf_stamp(a.dbg, 6);
      // ---- the five chains: warp 0 sums (lane k < 5: component k), warps 1-7 stage VF_CHUNK points per step; the points
      // run one step ahead of the sums and their indices one step ahead of the points, and every load is unconditional
      // (indices clamped, not branched around), so none of them makes a warp wait where it is issued
      float acc = 0.f;
      {
        constexpr int NL = 224, PER = VF_CHUNK / NL;
        static_assert(PER * NL == VF_CHUNK, "chunk = loader threads x points per thread");
        const int lt = tid - 32;
        float4 pp[PER];
        float xx[PER];
        uint32_t ii[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) xx[j] = 0.f;
        auto load_idx = [&](uint32_t base) {
#pragma unroll
          for (int j = 0; j < PER; ++j) ii[j] = __ldcg(vf.seg2 + off + min(base + (uint32_t)(lt + NL * j), c - 1));
        };
        auto load_pts = [&]() {
#pragma unroll
          for (int j = 0; j < PER; ++j) {
            pp[j] = __ldg(a.undist + ii[j]);
            if (a.aux) xx[j] = __ldg(a.aux + ii[j]);
          }
        };
        auto stash = [&](int buf) {
          float(*st)[VF_STAGE_ROW] = s_stage[buf];
#pragma unroll
          for (int j = 0; j < PER; ++j) {
            st[0][lt + NL * j] = pp[j].x;
            st[1][lt + NL * j] = pp[j].y;
            st[2][lt + NL * j] = pp[j].z;
            st[3][lt + NL * j] = pp[j].w;
            st[4][lt + NL * j] = xx[j];
          }
        };
        if (warp > 0) {
          load_idx(0);
          load_pts();
          load_idx(VF_CHUNK);
          stash(0);
        }
        __syncthreads();
        int buf = 0;
        for (uint32_t base = 0; base < c; base += VF_CHUNK) {
          const uint32_t m = min((uint32_t)VF_CHUNK, c - base);
          const bool next = base + VF_CHUNK < c;
          if (warp > 0) {
            if (next) {
              load_pts();
              load_idx(base + 2 * VF_CHUNK);
              stash(buf ^ 1);
            }
          } else if (lane < 5) {
            acc = chain_sum(acc, s_stage[buf][lane], m);
          }
          __syncthreads();
          buf ^= 1;
        }
      }
      if (tid < 5) s_sum[tid] = acc;
      __syncthreads();
      if (tid == 0) {
        leaf_write(a, h, c, s_sum[0], s_sum[1], s_sum[2], s_sum[3], s_sum[4]);
        leaf_cleanup(vf, h, true);
      }
      __syncthreads();
    }
    if (tid == 0) vf_stamp(a.dbg, 1);
    // ---- mid leaves: one warp per leaf
    {
      uint32_t* w_idx = reinterpret_cast<uint32_t*>(vf_smem) + warp * (VF_MID + 5 * VF_ROW);  // [VF_MID] indices
      float* w_pt = reinterpret_cast<float*>(w_idx + VF_MID);                                 // [5][VF_ROW] points at their rank
      const int n_mid = ctr_hi.y;
      for (;;) {
        int e = 0;
        if (lane == 0) e = atomicAdd(&vf.ctr[6], 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= n_mid) break;
        const uint4 leaf = __ldcg(vf.mid + e);
        const uint32_t h = leaf.x, c = leaf.y, off = leaf.z;
        uint32_t v[VF_MID / 32], r[VF_MID / 32];
#pragma unroll
        for (int q = 0; q < VF_MID / 32; ++q) {
          const uint32_t k = (uint32_t)(lane + 32 * q);
          v[q] = k < c ? __ldcg(vf.seg + off + k) : 0xFFFFFFFFu;
          w_idx[k] = v[q];
          r[q] = 0;
        }
        __syncwarp();
        for (uint32_t j = 0; j < c; ++j) {
          const uint32_t x = w_idx[j];
#pragma unroll
          for (int q = 0; q < VF_MID / 32; ++q) r[q] += x < v[q] ? 1u : 0u;
        }
#pragma unroll
        for (int q = 0; q < VF_MID / 32; ++q)
          if (v[q] != 0xFFFFFFFFu) {
            const float4 p = __ldg(a.undist + v[q]);
            w_pt[0 * VF_ROW + r[q]] = p.x;
            w_pt[1 * VF_ROW + r[q]] = p.y;
            w_pt[2 * VF_ROW + r[q]] = p.z;
            w_pt[3 * VF_ROW + r[q]] = p.w;
            w_pt[4 * VF_ROW + r[q]] = a.aux ? __ldg(a.aux + v[q]) : 0.f;
          }
        __syncwarp();
        float acc = 0.f;
        if (lane < 5) acc = chain_sum(0.f, w_pt + lane * VF_ROW, c);
        const float sx = __shfl_sync(0xffffffffu, acc, 0), sy = __shfl_sync(0xffffffffu, acc, 1),
                    sz = __shfl_sync(0xffffffffu, acc, 2), st = __shfl_sync(0xffffffffu, acc, 3),
                    si = __shfl_sync(0xffffffffu, acc, 4);
        if (lane == 0) {
          leaf_write(a, h, c, sx, sy, sz, st, si);
          leaf_cleanup(vf, h, true);
        }
        __syncwarp();
      }
    }
  }
  if (lane == 0) vf_stamp(a.dbg, 2);
  // ---- short leaves (and, after an error, the clean-up of every leaf): one thread per leaf, 256 leaves per ticket
  for (;;) {
    __syncthreads();
    if (tid == 0) s_flag = atomicAdd(&vf.ctr[7], 256);
    __syncthreads();
    const int e0 = s_flag;
    if (e0 >= n_leaf_all) break;
    const int e = e0 + tid;
    if (e >= n_leaf_all) continue;
    const uint32_t h = vf.list[e];
    if (ok && (vf.rank[h] & 0x80000000u)) continue;  // a warp or a block takes it (and may have cleared its count already)
    const uint32_t c = vf.cnt[h];
    if (ok) {
      const uint32_t off = vf.off[h];
      // The leaf's indices go through a sorting network in registers (fixed sequence of min/max, padding sorts to the
      // end), then every point load is issued before the first addition: two memory latencies per leaf, whatever its
      // length, instead of one per point (issue is in order: an addition that waits for its load holds up the rest).
      uint32_t v[VF_SMALL];
#pragma unroll
      for (int j = 0; j < VF_SMALL; ++j) v[j] = (uint32_t)j < c ? __ldcg(vf.seg + off + j) : 0xFFFFFFFFu;
#pragma unroll
      for (int size = 2; size <= VF_SMALL; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
#pragma unroll
          for (int i = 0; i < VF_SMALL; ++i) {
            const int j = i ^ stride;
            if (j > i) {
              const uint32_t lo = min(v[i], v[j]), hi = max(v[i], v[j]);
              const bool up = (i & size) == 0;
              v[i] = up ? lo : hi;
              v[j] = up ? hi : lo;
            }
          }
      float4 p[VF_SMALL];
      float x[VF_SMALL];
#pragma unroll
      for (int j = 0; j < VF_SMALL; ++j) {
        p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        x[j] = 0.f;
        if ((uint32_t)j < c) {
          p[j] = __ldg(a.undist + v[j]);
          if (a.aux) x[j] = __ldg(a.aux + v[j]);
        }
      }
      float sx = 0.f, sy = 0.f, sz = 0.f, st = 0.f, si = 0.f;
#pragma unroll
      for (int j = 0; j < VF_SMALL; ++j)
        if ((uint32_t)j < c) {
          sx = sx + p[j].x;
          sy = sy + p[j].y;
          sz = sz + p[j].z;
          st = st + p[j].w;
          si = si + x[j];
        }
      leaf_write(a, h, c, sx, sy, sz, st, si);
    }
    leaf_cleanup(vf, h, ok);
  }
  if (tid == 0) vf_stamp(a.dbg, 3);
  // ---- the last block files the scan's counters as the report the host reads (counters[16..23]) and leaves the working
  // set reset for the next scan
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(&vf.ctr[3], 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_flag && tid == 0) {
    int e2 = err;
    if (n_leaf_all > a.max_m && e2 == 0) e2 = 2;  // more voxels than lio_caps.max_down_points
    *a.scan_m = err >= 2 ? 0 : n_leaf;
    a.counters[16] = n_leaf_all;
#pragma unroll
    for (int k = 1; k < 7; ++k) a.counters[16 + k] = a.counters[k];
    a.counters[23] = e2;
    a.counters[0] = 0;
    a.counters[1] = a.counters[2] = a.counters[3] = 0x7fffffff;
    a.counters[4] = a.counters[5] = a.counters[6] = -0x7fffffff;
    a.counters[7] = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) vf.ctr[k] = 0;
    vf_stamp(a.dbg, 4);
  }
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
  prep_reset_kernel<<<1, 1, 0, c->stream>>>(c->d_prep_counters, c->vf.ctr);
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

// raw points already staged in c->d_raw (and c->d_raw_aux when has_aux), poses in c->d_poses
int preprocess(lio_ctx* c, int64_t n, int n_poses, const lio_state* end_state, float leaf, bool has_aux) {
  if (n_poses > MAX_POSES) {
    c->err = "more than 128 IMU poses in one scan";
    return LIO_E_CAPACITY;
  }
  const int max_m = (int)c->caps.max_down_points;
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
  a.vf = c->vf;
  a.max_m = max_m;
  if (n > 0) {
    const int grid = (int)((n + 255) / 256);
    // leaves: at most one per point and never more than the hash holds
    const int lgrid = (int)((std::min<int64_t>(n, (int64_t)c->vf.hash_mask + 1) + 255) / 256);
    undistort_key_kernel<<<grid, 256, 0, c->stream>>>(a);
    voxel_bits_kernel<<<lgrid, 256, 0, c->stream>>>(c->vf, c->d_prep_counters);
    voxel_place_kernel<<<grid, 256, 0, c->stream>>>(c->vf, c->d_prep_counters, (int)n);
    c->launches += 3;
  }
  CentroidArgs ca;
  ca.vf = c->vf;
  ca.undist = c->d_undist;
  ca.aux = has_aux ? c->d_raw_aux : nullptr;
  ca.n = (int)n;
  ca.max_m = max_m;
  ca.body = c->d_body;
  ca.body_time = reinterpret_cast<float*>(c->d_normvec);  // scratch: mean time
  ca.scan_m = c->d_scan_m;
  ca.counters = c->d_prep_counters;
  ca.dbg = c->d_dbg;
  if (c->d_dbg) LIO_CHECK(c, cudaMemsetAsync(c->d_dbg + 200, 0x80, 8 * 8, c->stream));
  // the one kernel that overwrites d_body / d_scan_m, which a pending map growth still reads
  if (c->centroid_wait) LIO_CHECK(c, cudaStreamWaitEvent(c->stream, c->centroid_wait, 0));
  static bool attr_set[64] = {false};
  if (c->device >= 64 || !attr_set[c->device]) {
    LIO_CHECK(c, cudaFuncSetAttribute((const void*)voxel_centroid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      VF_CENTROID_SMEM));
    if (c->device < 64) attr_set[c->device] = true;
  }
  // long leaves first (one block each, strided), then tickets for the mid and the short ones
  const int64_t leaves_max = std::min<int64_t>(n, (int64_t)c->vf.hash_mask + 1);
  const int cgrid = std::max(1, std::min((int)((leaves_max + 63) / 64), c->sm_count * 2));
  voxel_centroid_kernel<<<cgrid, 256, VF_CENTROID_SMEM, c->stream>>>(ca);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  c->scan_m = -1;  // known on the device only until someone asks
  return LIO_OK;
}

}  // namespace lio
