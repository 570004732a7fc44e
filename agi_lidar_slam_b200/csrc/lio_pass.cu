// IESKF measurement update on the device.
//   pass   ≙ esekf::h_share_model (esekfom.hpp:106-227) fused with the H^T H / H^T h products of
//            update_iterated_dyn_share_modified (esekfom.hpp:306-319): body->world, 5-NN on the voxel hash,
//            validity gates, 5x3 plane fit, residual, 1x12 Jacobian row, block-level FP64 reduction.
//   solve  ≙ the rest of one loop iteration (esekfom.hpp:297-345): grid-level reduction, boxminus, Kalman step,
//            boxplus, convergence state machine, covariance update.
// update_kernel runs the WHOLE update_iterated_dyn_share_modified loop as one persistent cooperative launch.  Blocks
// 0 .. grid-2 are workers (one h_share_model pass after the other: search tile -> finish tile -> block reduction ->
// one partial row of self-validating stamped words); the last block is the solver: it keeps the filter in shared
// memory, adds the workers' rows once their stamps show up, performs the Kalman step and publishes the new pose the
// same way.  There is no host round trip, no launch and no fence between passes.  The sharded-map variant (ownership
// lists, runs dealt over the blocks, peer exchange) lives in its own instantiation (block_pass<true>, update_kernel_sh)
// so that the plain update does not carry its code or its registers.  pass_kernel / solve_kernel / begin_kernel run the same device code
// as single steps (host-driven passes, and the NCCL variant of the sharded-map driver); update_kernel_host adds the
// host-direct prologue / epilogue; block_exchange moves the sharded-map blobs through NVLink peer mailboxes.
// No floating-point atomics anywhere: every sum has a fixed order.
//
// Kalman step.  The reference forms K_front = (H^T H / R + P^-1)^-1 with two 24x24 inverses per pass
// (esekfom.hpp:311).  H has n = 6 (12 with extrinsic_est) non-zero columns, so with P11 = P[:n,:n], P21 = P[n:,:n]
//     K_front[:, :n] = [ I ; P21 P11^-1 ] (H^T H / R + P11^-1)^-1
// (block inversion; exact algebra).  P is constant during the loop, so P11^-1 and P21 P11^-1 are formed once per
// update and each pass inverts one n x n SPD matrix (Gauss-Jordan in shared memory, one element per thread).
#include <stdio.h>

#include <algorithm>

#include "lio_ctx.cuh"
#include "lio_knn.cuh"

namespace lio {

// One block of 512 threads per SM.  The grid reduction costs the solving block a load of every worker row through its
// one SM (~1 us per 100 KB), so half as many, fatter blocks halve it; and a lone block per SM keeps its instruction
// stream to itself.  Measured on the bench workload: 2 x 256 threads per SM 146 us, 1 x 512 threads 124 us per update.
#ifndef LIO_THREADS
#define LIO_THREADS 512
#endif
constexpr int THREADS = LIO_THREADS;  // threads per block of every kernel in this file
#ifndef LIO_BLOCKS_PER_SM
#define LIO_BLOCKS_PER_SM 1
#endif
constexpr int ROWS_MAX = 256;    // Jacobian rows staged per tile (cached passes: one thread per point)
constexpr int RS = 14;           // row stride: 12 Jacobian columns, residual, 1.0 (row counter)
constexpr int NOUT_EXT = 91;     // 78 HtH + 12 Hth + count
constexpr int NOUT_NOEXT = 28;   // 21 HtH (6x6 upper) + 6 Hth + count

// compact output o -> (row column a, row column b): the accumulated quantity is sum_rows row[a] * row[b]
__constant__ unsigned char c_oa_ext[NOUT_EXT], c_ob_ext[NOUT_EXT], c_oe_ext[NOUT_EXT];
__constant__ unsigned char c_oa_no[NOUT_NOEXT], c_ob_no[NOUT_NOEXT], c_oe_no[NOUT_NOEXT];
__constant__ unsigned long long c_is_no[2];  // bit e set where HtH entry e (< 78) is one of the 21 kept without extrinsic estimation
                                             // (a mask, not a table: read with a uniform address, the constant bank serves it in one go)

struct PassArgs {
  const float4* body;
  const int* scan_m;
  int min_m;  // a scan with fewer points is skipped by the main loop (laserMapping.cpp:741-744): treated as empty
  int m_value;             // >= 0: the scan size by value (host-direct path), else *scan_m
  const float4* body_src;  // non-null: pass 0 reads the scan from here (pinned host memory) and leaves a copy in `body`
  MapView map;
  float4* near_pts;
  float* near_d2;
  int* near_cnt;
  float4* near_q;    // the FP32 p_world every row was searched at (its last search pass): what an unbounded completion of
                     // the row (far_search_kernel) has to use as the query
  uint8_t* selected;
  float4* normvec;
  float4* plane;     // pabcd of the last search pass (the fit depends on the neighbours only)
  float4* world;
  int extrinsic_est;
  float max_d2, plane_thr;
  int rings;
  float own_min, own_max;
  int sharded;  // rows are shared out over the ranks (a finite window, or stripes)
  int batched_finish;  // 0 (LIO_FINISH_BATCHED=0): blocks with several tiles finish each tile right after its search
  int interleave;  // LIO_INTERLEAVE=1: runs of 8 points dealt round-robin over the blocks also on one GPU (load balance)
  // striped ownership (lio_set_shard_stripes): the rank owns the stripes s = floor((x - origin) / width) with
  // s mod stripe_mod == stripe_rank; stripe_mod == 0: the window above
  float stripe_origin, stripe_inv_w;
  int stripe_mod, stripe_rank;
  int stage;  // 1: the searches of a tile stage its neighbour cells in shared memory first (stage_cells)
  unsigned long long* partials;  // [workers][ROW_WORDS]: one row of stamped words per worker block (st_stamped)
  long long* dbg;    // optional timeline (LIO_TIMELINE=1): [0] = entries used by block 0, [1..] = (tag, globaltimer ns)
};

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Timeline instrumentation (LIO_TIMELINE=1).  The LAST thread of block 0 appends (tag, time) pairs from slot 1, the last
// thread of the solving block from slot 129.  The last thread, because a lane that branches off on its own leaves its
// warp split for the code that follows (see warp_gauss_jordan), and warp 0 is where the solver runs its collectives.
__device__ __forceinline__ void stamp(long long* dbg, int base, int tag) {
  if (dbg == nullptr || threadIdx.x != blockDim.x - 1) return;
  if (base == 0 && blockIdx.x != 0) return;
  const long long n = dbg[base];
  if (n < 62) {
    dbg[base + 1 + 2 * n] = tag;
    dbg[base + 2 + 2 * n] = global_ns();
    dbg[base] = n + 1;
  }
}

// per-block / per-warp times of the last pass (lio_debug_blocks): slot 256 + k of the timeline buffer
__device__ __forceinline__ void stamp_slot(long long* dbg, int k) {
  if (dbg != nullptr) dbg[256 + k] = global_ns();
}

struct SolveArgs {
  StateD* x;
  StateD* xprop;
  double* P;
  const StateD* x0;  // prior snapshot (lio_state_upload)
  const double* P0;
  Ctrl* ctrl;
  double* dx_out;
  double* blob;      // LIO_BLOB: reduced blob of the last pass (also the all-reduce buffer of the sharded driver)
  double* prior;     // 144 (P11^-1) + 144 (P21 P11^-1, (24-n) x n)
  unsigned* sync;    // [0] arrival counter, [1] release flag
  unsigned long long* pub;  // PUB_WORDS stamped words: x[0..13] of the new state and {converge, done} (publish_state)
  double R;
  int max_iter;
  int from_snapshot;
  long long* dbg;
};

// per-pass constants shared by the block
struct PassConst {
  Quatd rot, rli;
  double pos[3], tli[3];
  double Rt[9], Rli[9];
};

__device__ __forceinline__ int scan_size(const PassArgs& a) {
  const int M = a.m_value >= 0 ? a.m_value : *a.scan_m;
  return M < a.min_m ? 0 : M;
}
// Ownership of the scan points: worker w owns the contiguous chunk [w * C, (w + 1) * C) in EVERY pass, search or cached.
// The per-point arrays (neighbour cache, plane, selected, normvec, ...) are then private to one block for the whole
// update, so no pass has to wait for another block's writes to become visible -- the only traffic between blocks is the
// partial rows up and the new state down, both made of self-validating words (st_stamped / ld_stamped): no fences.
// (Tiles dealt round-robin instead of contiguous chunks were tried for load balance -- tile costs differ by 3x between
// open ground and cluttered corners -- and lost: 137 us against 124 us per update on the bench workload.)
__device__ __forceinline__ int chunk_points(int M, int nworkers) {
  int c = (M + nworkers - 1) / nworkers;
  c = (c + 7) & ~7;
  return c < 8 ? 8 : c;
}
__device__ __forceinline__ int workers_used(int M, int nworkers) {
  const int c = chunk_points(M, nworkers);
  return (M + c - 1) / c;  // <= nworkers
}
// lanes per query in a search: all queries of the chunk in flight at once when it is small (the pass is latency-bound)
__device__ __forceinline__ int pick_group(int C) { return C <= THREADS / 32 ? 32 : (C <= THREADS / 16 ? 16 : 8); }

// A double as two 64-bit words {stamp << 32 | low half, stamp << 32 | high half}: each word is written by one 8-byte
// store, so a reader that finds the expected stamp in both holds the value that was written with it -- no fence, no flag,
// no second round trip.  Stamps are epoch + pass + 1 and never repeat between launches.
__device__ __forceinline__ void st_stamped(unsigned long long* p, double v, unsigned stamp) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long hs = (unsigned long long)stamp << 32;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(hs | (b & 0xffffffffull)), "l"(hs | (b >> 32))
               : "memory");
}
__device__ __forceinline__ bool ld_stamped(const unsigned long long* p, unsigned stamp, double& v) {
  unsigned long long lo, hi;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
  return (unsigned)(lo >> 32) == stamp && (unsigned)(hi >> 32) == stamp;
}
constexpr int ROW_WORDS = 2 * LIO_BLOB;  // 64-bit words of one worker's partial row

__device__ __forceinline__ void load_pass_const(const StateD* x, PassConst& pc) {
  const double* s = reinterpret_cast<const double*>(x);
  double v[26];
#pragma unroll
  for (int k = 0; k < 14; ++k) v[k] = __ldcg(s + k);
  pc.pos[0] = v[0]; pc.pos[1] = v[1]; pc.pos[2] = v[2];
  pc.rot = Quatd{v[3], v[4], v[5], v[6]};
  pc.rli = Quatd{v[7], v[8], v[9], v[10]};
  pc.tli[0] = v[11]; pc.tli[1] = v[12]; pc.tli[2] = v[13];
  quat_to_mat(pc.rot, pc.Rt);
  quat_to_mat(pc.rli, pc.Rli);
}

// step 1.1-1.2 (esekfom.hpp:123-133): p_world = rot * (R_LI * p + t_LI) + pos, FP64 -> FP32
__device__ __forceinline__ void body_to_world(const PassConst& pc, const double pb[3], float& pwx, float& pwy,
                                              float& pwz) {
  double pi[3], pg[3];
  quat_rotate(pc.rli, pb, pi);
  pi[0] += pc.tli[0];
  pi[1] += pc.tli[1];
  pi[2] += pc.tli[2];
  quat_rotate(pc.rot, pi, pg);
  pwx = (float)(pg[0] + pc.pos[0]);
  pwy = (float)(pg[1] + pc.pos[1]);
  pwz = (float)(pg[2] + pc.pos[2]);
}

constexpr int SROWS_MAX = THREADS / 8;  // rows of a search tile at the smallest group size

// Does this rank own a row searched at world x = qx?  The same FP32 operations on every rank: the answers partition the rows.
__device__ __forceinline__ bool owns_row(const PassArgs& a, float qx) {
  if (a.stripe_mod > 0) {
    const int st = (int)floorf((qx - a.stripe_origin) * a.stripe_inv_w);
    int r = st % a.stripe_mod;
    if (r < 0) r += a.stripe_mod;
    return r == a.stripe_rank;
  }
  return (qx >= a.own_min) && (qx < a.own_max);
}

// A block works through positions t = 0 .. n-1 of its chunk of the scan: all of its points in order, or -- sharded map --
// only the ones this rank owns (list = their offsets in the chunk, ascending).
constexpr int OWN_MAX = 2048;  // longest chunk a list is kept for (beyond: every point is visited and the rows are masked)
// Which points a block visits.  Unsharded: its contiguous chunk [beg, beg + C) (neighbouring queries share cells).
// Sharded: runs of 8 points dealt round-robin over the blocks (run g = r * nworkers + wid), so that every block holds a
// uniform sample of the scan and owns ~1/world of it whatever the shape of the ranks' regions -- with contiguous chunks
// (28 m of scan each) whole blocks fall to one rank and the slowest block of a pass is as slow as on one GPU.
template <bool SH>
struct Chunk {
  int beg, wid, nworkers;
  const unsigned short* list;  // sharded: offsets k of the owned points, ascending; else nullptr (k = t)
  bool dealt;                  // runs of 8 dealt round-robin (always with a list)
  __device__ __forceinline__ int point(int k) const {
    return (SH && dealt) ? ((((k >> 3) * nworkers + wid) << 3) + (k & 7)) : beg + k;
  }
  __device__ __forceinline__ int at(int t) const { return point((SH && list) ? (int)list[t] : t); }
};
// SH = false is the instantiation every unsharded update runs: lists, dealing and ownership are compiled out of it (a
// pass runs from a cold instruction cache and registers are at the limit: the feature must cost nothing where it is off)
__device__ __forceinline__ bool sharded_lists(const PassArgs& a, int C) { return a.sharded && C <= OWN_MAX; }
// runs dealt round-robin: every worker files a row (the reduction waits for all of them)
__device__ __forceinline__ bool dealt_runs(const PassArgs& a, int C) { return sharded_lists(a, C) || a.interleave; }
// number of points of worker `wid` under the dealing: its runs r = 0 .. C/8-1 clipped to the scan (a prefix of k)
__device__ __noinline__ int dealt_count(int M, int C, int nworkers, int wid) {
  int n = 0;
  for (int r = 0; r < (C >> 3); ++r) n += max(0, min(8, M - (((r * nworkers) + wid) << 3)));
  return n;
}

// Search phase of one tile: its queries (at most THREADS / G), one per G-lane group (esekfom.hpp:140).  The 5 neighbours
// go to the cache the later passes read (a.near_*) and to shared memory for the finish phase of this tile.  With staging
// on the caller has staged the tile's neighbour cells (stage_cells); use_stage says whether the tile fitted.
template <int G, bool SH, bool STG>
__device__ __noinline__ void search_tile(const PassArgs& a, const PassConst& pc, int t0, int n, const Chunk<SH> ch,
                                         float4* s_nb, int* s_cnt, float4* s_body, uint2* s_cells,
                                         const float4* body, StageSmem* st, bool use_stage, bool copy_body) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int row = threadIdx.x / G;
  const bool act = t0 + row < n;  // group-uniform
  const int qi = act ? row : n - 1 - t0;  // idle groups redo the last query: the whole warp stays together for the shuffles
  const int i = ch.at(t0 + qi);
  const float4 b = body[i];  // device copy, or the caller's pinned host buffer in pass 0 of the host-direct path
  if (gl == 0 && act) {
    s_body[row] = b;
    if (copy_body) const_cast<float4*>(a.body)[i] = b;  // host-direct pass 0: the finish phase reads the device copy
  }
  const double pb[3] = {b.x, b.y, b.z};
  float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
  body_to_world(pc, pb, qv.x, qv.y, qv.z);
  unsigned long long key[LIO_K];
  uint32_t slot[LIO_K];
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[239] = global_ns();
  const int cnt = group_knn5<G>(a.map, qv.x, qv.y, qv.z, a.max_d2, a.rings, gl, s_cells + row * KNN_CELLS, key, slot, a.dbg,
                                (STG && use_stage) ? st : nullptr, qi);
  if (!act) return;
  // lanes 0..4 of the group fetch and publish one neighbour each
  if (gl < LIO_K) {
    unsigned long long k = key[0];
    uint32_t sl = slot[0];
#pragma unroll
    for (int r = 1; r < LIO_K; ++r)
      if (gl == r) {
        k = key[r];
        sl = slot[r];
      }
    float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    float d = CUDART_INF_F;
    if (gl < cnt) {
      v = (STG && (sl & ST_FLAG)) ? st->pts[sl & ~ST_FLAG] : __ldg(a.map.pool + sl);
      d = __uint_as_float((uint32_t)(k >> 32));
    }
    s_nb[row * LIO_K + gl] = v;
    a.near_pts[(size_t)i * LIO_K + gl] = v;
    a.near_d2[(size_t)i * LIO_K + gl] = d;
  }
  if (gl == 0) {
    // step 1.4 (esekfom.hpp:144-147): gate on the count and on the 5th squared distance
    const float d4 = __uint_as_float((uint32_t)(key[LIO_K - 1] >> 32));
    const bool sel = (cnt < LIO_K) ? false : (d4 > 5.0f ? false : true);
    s_cnt[row] = sel ? 1 : 0;
    a.near_cnt[i] = cnt;
    a.near_q[i] = make_float4(qv.x, qv.y, qv.z, 0.f);
  }
  if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[245] = global_ns();
}

// Finish phase of one tile, one thread per point: steps 1.1-1.2 and 1.5-3 of h_share_model (esekfom.hpp:123-133,
// 153-226) from the point's 5 neighbours (just found: shared memory; cached: a.near_pts with the sticky mask).
template <bool SH>
__device__ __noinline__ void finish_tile(const PassArgs& a, const PassConst& pc, int t0, int n, const Chunk<SH> ch,
                                         int rows, int mode, const float4* s_nb,
                                         const int* s_cnt, double* s_rows, unsigned char* s_valid, const float4* s_body,
                                         bool copy_body) {
  // mode 0: cached pass (plane of the last search pass); 1: the search tile just done (neighbours, gate and points in
  // shared memory); 2: search pass of a block with several tiles, all searched before (neighbours and gate from the cache
  // they were filed in, one thread per row for ROWS_MAX rows at a time instead of a quarter of that per search tile)
  const bool search = mode != 0;
  const int row = threadIdx.x;
  if (row >= rows) return;
  if (t0 + row >= n) {
    s_valid[row] = 0;
    return;
  }
  const int i = ch.at(t0 + row);
  const unsigned short* list = SH ? ch.list : nullptr;
  const float4 b = mode == 1 ? s_body[row] : __ldcg(a.body + i);  // the search phase left it in shared memory
  if (copy_body) const_cast<float4*>(a.body)[i] = b;               // host-direct path, pass 0: keep a device copy
  const double pb[3] = {b.x, b.y, b.z};
  float pwx, pwy, pwz;
  body_to_world(pc, pb, pwx, pwy, pwz);
  a.world[i] = make_float4(pwx, pwy, pwz, b.w);
  // esti_plane (common_lib.h:102-134) depends on the 5 neighbours only, and they change only in search passes: a
  // cached pass re-reads the plane fitted by the last search pass instead of repeating the QR (same bits).  A point
  // still selected at that time necessarily passed the fit (esekfom.hpp:150-173).
  float pabcd[4] = {0.f, 0.f, 0.f, 0.f};
  float pd2 = 0.f;
  bool sel;
  if (search) {
    float4 nb[LIO_K];
    if (mode == 1) {
#pragma unroll
      for (int r = 0; r < LIO_K; ++r) nb[r] = s_nb[row * LIO_K + r];
      sel = s_cnt[row] != 0;
    } else {
#pragma unroll
      for (int r = 0; r < LIO_K; ++r) nb[r] = __ldcg(a.near_pts + (size_t)i * LIO_K + r);
      // gate 1 as search_tile takes it (esekfom.hpp:144-147): five neighbours, the fifth within sqrt(5) m
      sel = __ldcg(a.near_cnt + i) >= LIO_K && !(__ldcg(a.near_d2 + (size_t)i * LIO_K + (LIO_K - 1)) > 5.0f);
    }
    if (sel) {
      sel = esti_plane(nb, a.plane_thr, pabcd);
      a.plane[i] = make_float4(pabcd[0], pabcd[1], pabcd[2], pabcd[3]);
    }
  } else {
    sel = __ldcg(a.selected + i) != 0;  // sticky between search passes (esekfom.hpp:150)
    if (sel) {
      const float4 pl = __ldcg(a.plane + i);
      pabcd[0] = pl.x;
      pabcd[1] = pl.y;
      pabcd[2] = pl.z;
      pabcd[3] = pl.w;
    }
  }
  if (sel) {
    sel = false;
    pd2 = ((pabcd[0] * pwx + pabcd[1] * pwy) + pabcd[2] * pwz) + pabcd[3];
    const double nrm = sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
    const float sc = (float)(1.0 - 0.9 * fabs((double)pd2) / sqrt(nrm));
    if ((double)sc > 0.9) sel = true;
  }
  a.selected[i] = sel ? 1 : 0;
  if (sel) a.normvec[i] = make_float4(pabcd[0], pabcd[1], pabcd[2], pd2);
  // Ownership (sharded map): by the x of the position the row was SEARCHED at -- the same on every rank and fixed until
  // the next search pass, so a row and its cached neighbours stay with one rank.  Listed rows are owned by construction.
  bool valid = sel;
  if (SH) {
    float qx = pwx;
    if (!list && !search && a.sharded) qx = __ldcg(&a.near_q[i]).x;
    valid = sel && (list != nullptr || owns_row(a, qx));
  } else {
    valid = sel && !(pwx != pwx);  // (an unsharded update owns every row; a NaN position contributes nothing, as before)
  }
  s_valid[row] = valid ? 1 : 0;
  if (valid) {
    // step 3 (esekfom.hpp:197-226): Jacobian row and residual
    double* rowp = s_rows + row * RS;
    double pI[3], C[3], A[3];
    quat_rotate(pc.rli, pb, pI);
    pI[0] += pc.tli[0];
    pI[1] += pc.tli[1];
    pI[2] += pc.tli[2];
    const double nv[3] = {(double)pabcd[0], (double)pabcd[1], (double)pabcd[2]};
    mat3T_vec(pc.Rt, nv, C);
    const double pIx[9] = {0.0, -pI[2], pI[1], pI[2], 0.0, -pI[0], -pI[1], pI[0], 0.0};
    mat3_vec(pIx, C, A);
    rowp[0] = nv[0];
    rowp[1] = nv[1];
    rowp[2] = nv[2];
    rowp[3] = A[0];
    rowp[4] = A[1];
    rowp[5] = A[2];
    if (a.extrinsic_est) {
      double M1[9], B[3];
      const double px[9] = {0.0, -pb[2], pb[1], pb[2], 0.0, -pb[0], -pb[1], pb[0], 0.0};
      for (int r = 0; r < 3; ++r)
        for (int c2 = 0; c2 < 3; ++c2)
          M1[3 * r + c2] =
              (px[3 * r] * pc.Rli[3 * c2] + px[3 * r + 1] * pc.Rli[3 * c2 + 1]) + px[3 * r + 2] * pc.Rli[3 * c2 + 2];
      mat3_vec(M1, C, B);
      rowp[6] = B[0];
      rowp[7] = B[1];
      rowp[8] = B[2];
      rowp[9] = C[0];
      rowp[10] = C[1];
      rowp[11] = C[2];
    } else {
      rowp[6] = rowp[7] = rowp[8] = rowp[9] = rowp[10] = rowp[11] = 0.0;
    }
    rowp[12] = -(double)pd2;  // esekfom.hpp:225
    rowp[13] = 1.0;
  }
}

// One h_share_model pass of this block: its tiles (tile = blockIdx.x, += gridDim.x), Jacobian rows staged in shared
// memory, products accumulated by thread (output o, segment seg) over rows seg, seg + nseg, ... of every tile, then
// the segments are combined in order and the block's partial blob is written to a.partials[blockIdx.x].
struct PassSmem;
template <bool SH, bool STG>
__device__ void block_pass(const PassArgs& a, bool search, PassSmem* ps, StageSmem* st, int nworkers, int wid,
                           bool first_pass, unsigned target);

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Sum of the workers' partial rows into s_warp[warp][LIO_BLOB] in a FIXED order: warp w adds the rows of workers
// w, w + nwarps, ... in ascending order (lane = output, one coalesced 16-byte load per output and row, CH rows in
// flight).  A row is taken when every word of it carries `target` (ld_stamped); until then the chunk is simply loaded
// again, so the reduction runs WHILE the slower workers are still busy and only the last chunk is left when the last
// one is done.  The waiting does not change the order, hence not the bits.
template <int KPL, int CH>
__device__ __forceinline__ void reduce_rows(const PassArgs& a, int nb, int nout, unsigned target, double* s_warp) {
  static_assert(CH <= 32, "one lane watches one row of the chunk");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  // NO lane-dependent branch in here: a warp that splits can stay split, and every vote behind the split costs a
  // rendezvous of ~1,000 cycles instead of ~30 (measured).  Lanes without an output load output 0 and drop it; rows
  // beyond the last one are clamped to it and dropped.
  // A worker's row holds its nout outputs back to back (2 words each); s_warp is indexed by blob position.
  int e[KPL], o[KPL];
  bool use[KPL];
  double acc[KPL];
#pragma unroll
  for (int k = 0; k < KPL; ++k) {
    o[k] = lane + 32 * k;
    use[k] = o[k] < nout;
    if (!use[k]) o[k] = 0;
    e[k] = a.extrinsic_est ? c_oe_ext[o[k]] : c_oe_no[o[k]];
    acc[k] = 0.0;
  }
  const unsigned long long* a_partials = a.partials;  // (read once: the asm statements below clobber nothing)
#pragma unroll 1
  for (int base = warp; base < nb; base += NW * CH) {
    // 1. wait: lane j watches ONE word of row j of the chunk until all of them have shown up.  Cheap (a few dozen
    //    requests per round trip for the whole block), and the bulk load below -- all the rows through one SM: ~100 KB,
    //    a microsecond -- is then issued once, not over and over while the slowest worker is still busy.
    {
      const int jj = lane < CH ? lane : CH - 1;
      const int b = base + NW * jj;
      const unsigned long long* w0 = a.partials + (size_t)(b < nb ? b : nb - 1) * ROW_WORDS;
      bool seen;
      do {
        unsigned long long w;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(w0) : "memory");
        seen = (unsigned)(w >> 32) == target;
      } while (!__all_sync(0xffffffffu, seen));
    }
    stamp(a.dbg, 128, 13);
    if (lane == 0) stamp_slot(a.dbg, 512 + warp);
    // 2. load: one coalesced 16-byte load per output and row.  ALL the loads of the chunk are issued before the first
    //    stamp is looked at (rows beyond the last one are clamped to it and dropped, so no load is conditional): with
    //    the check behind every load the compiler waited for each pair of them, and the ten rows of a warp cost five
    //    L2 round trips instead of one (2.3 us of every pass between "rows seen" and "rows loaded").  Every word
    //    validates itself (the other words of a row may lag a little behind the watched one: then simply once more).
    unsigned long long lo[CH][KPL], hi[CH][KPL];
    bool ready;
    do {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int b = base + NW * j;
        const unsigned long long* rowp = a_partials + (size_t)(b < nb ? b : nb - 1) * ROW_WORDS;
#pragma unroll
        for (int k = 0; k < KPL; ++k)
          asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
                       : "=l"(lo[j][k]), "=l"(hi[j][k])
                       : "l"(rowp + 2 * o[k]));
      }
      ready = true;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const bool in = base + NW * j < nb;  // the same for the whole warp
#pragma unroll
        for (int k = 0; k < KPL; ++k)
          ready = ready && (!in || ((unsigned)(lo[j][k] >> 32) == target && (unsigned)(hi[j][k] >> 32) == target));
      }
    } while (!__all_sync(0xffffffffu, ready));
    if (lane == 0) stamp_slot(a.dbg, 544 + warp);
    // 3. add, in row order
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const bool in = base + NW * j < nb;
#pragma unroll
      for (int k = 0; k < KPL; ++k) {
        const double x = __longlong_as_double((long long)((hi[j][k] << 32) | (lo[j][k] & 0xffffffffull)));
        acc[k] += (in && use[k]) ? x : 0.0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KPL; ++k)
    if (use[k]) s_warp[warp * LIO_BLOB + e[k]] = acc[k];
}

__device__ __noinline__ void block_reduce_partials(const PassArgs& a, bool search, int nworkers, unsigned target,
                                                   double* s_blob, double* s_warp) {
  const int tid = threadIdx.x;
  const int Ms = scan_size(a);
  // (sharded lists: the runs of the scan are dealt over ALL the workers, every one of them files a row)
  const int nb = dealt_runs(a, chunk_points(Ms, nworkers)) ? nworkers : workers_used(Ms, nworkers);
  const int nout = a.extrinsic_est ? NOUT_EXT : NOUT_NOEXT;
  if (a.extrinsic_est)
    reduce_rows<3, 5>(a, nb, nout, target, s_warp);
  else
    reduce_rows<1, 10>(a, nb, nout, target, s_warp);  // 16 warps x 10 rows: every worker row of a 148-SM grid in one chunk
  __syncthreads();
  if (tid < LIO_BLOB) {
    double sum = 0.0;
    const int o_valid = a.extrinsic_est ? 1 : ((tid < 78 && ((c_is_no[tid >> 6] >> (tid & 63)) & 1ull)) || (tid >= 78 && tid < 84) || tid == 90);
    if (tid < 91 && o_valid) {
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) sum += s_warp[w * LIO_BLOB + tid];
    }
    if (tid == 91) sum = search ? 1.0 : 0.0;
    s_blob[tid] = sum;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// n x n FP64 inverse (n <= 12) of a symmetric positive definite matrix: Gauss-Jordan on [A | I] held in shared
// memory (W: n rows of WS doubles), ONE ELEMENT PER THREAD.  Both matrices inverted here are SPD by construction (P11
// is a covariance block, S = HtH / R + P11^-1), so no pivot search is needed: step k subtracts
// (a_ik / a_kk) * row k from every other row.  An element (i != k, c > k) reads a_ik, a_kk, a_kc -- none of which is
// written in step k -- so a step is: four shared loads, one division, one FMA, one store, one barrier.  The first
// `nthr` threads of the block call this (nthr = inv_threads(n), whole warps) and meet at named barrier 1; the other
// warps are free to do something else meanwhile.  A LOOP on purpose: the solve runs from a cold instruction cache,
// so its cost is its code size.
// ---------------------------------------------------------------------------------------------------------
constexpr int WS = 25;  // odd row stride: the rows of one column land in distinct banks
// 1/p to double precision without the IEEE division sequence: FP32 reciprocal seed + two Newton steps in FP64
// (1e-7 -> 1e-14 -> 1e-28 relative before rounding).  Pivots of an SPD matrix: positive, far from the FP32 range limits.
__device__ __forceinline__ double fast_rcp(double p) {
  double r = (double)__frcp_rn((float)p);
  r = fma(r, fma(-p, r, 1.0), r);
  r = fma(r, fma(-p, r, 1.0), r);
  return r;
}
__device__ __forceinline__ int inv_threads(int n) { return n <= 6 ? 96 : 192; }
__device__ __forceinline__ void inv_barrier(int nthr) { asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); }
__device__ __noinline__ void block_inverse_spd(const double* A, double* Ainv, int n, double* W) {
  const int tid = threadIdx.x;
  const int nthr = inv_threads(n);
  const int w = 2 * n, total = n * w;
#pragma unroll 1
  for (int idx = tid; idx < total; idx += nthr) {
    const int r = idx / w, c = idx - r * w;
    W[r * WS + c] = c < n ? A[r * n + c] : (c - n == r ? 1.0 : 0.0);
  }
  // this thread's (at most two) elements
  const int i0 = tid / w, c0 = tid - i0 * w;
  const int idx1 = tid + nthr;
  const int i1 = idx1 / w, c1 = idx1 - i1 * w;
  const bool has0 = tid < total, has1 = idx1 < total;
  inv_barrier(nthr);
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    const double pinv = fast_rcp(W[k * WS + k]);
    if (has0 && i0 != k && c0 > k) W[i0 * WS + c0] = fma(-(W[i0 * WS + k] * pinv), W[k * WS + c0], W[i0 * WS + c0]);
    if (has1 && i1 != k && c1 > k) W[i1 * WS + c1] = fma(-(W[i1 * WS + k] * pinv), W[k * WS + c1], W[i1 * WS + c1]);
    inv_barrier(nthr);
  }
  if (has0 && c0 >= n) Ainv[i0 * n + (c0 - n)] = W[i0 * WS + c0] * fast_rcp(W[i0 * WS + i0]);
  if (has1 && c1 >= n) Ainv[i1 * n + (c1 - n)] = W[i1 * WS + c1] * fast_rcp(W[i1 * WS + i1]);
  inv_barrier(nthr);
}

// ---------------------------------------------------------------------------------------------------------
// SO(3) pieces of boxplus / boxminus (esekfom.hpp:59-73, 236-258) in quaternion form, one rotation per thread so the
// two rotations of the state (rot, offset_R_L_I) run side by side.  Same formulas as Sophus::SO3::exp / log
// (lio_common.cuh); the relative rotation of boxminus is formed as conj(q2) * q1 instead of through 3x3 matrices,
// and unit norm is restored with the series of 1/sqrt(1+e) (e ~ 1e-16) instead of sqrt + 4 divisions.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Quatd qmul_raw(const Quatd& a, const Quatd& b) {
  Quatd r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ Quatd renorm(const Quatd& q) {
  const double n2 = ((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z;
  const double e = n2 - 1.0;
  const double sc = fabs(e) < 1e-6 ? (1.0 - 0.5 * e) + 0.375 * (e * e) : 1.0 / sqrt(n2);
  return Quatd{q.w * sc, q.x * sc, q.y * sc, q.z * sc};
}
// q * Exp(w)
__device__ __noinline__ void rot_plus(const Quatd* q, const double* w, Quatd* out) {
  const double th2 = (w[0] * w[0] + w[1] * w[1]) + w[2] * w[2];
  double imag, real;
  if (th2 < 2.5e-3) {
    // |w| < 0.05 rad (every step after the first few centimetres of correction): sin(t/2)/t and cos(t/2) as series in
    // t^2 -- no square root, no sincos, no division on the critical path; truncation < 1e-17
    imag = 0.5 + th2 * (-1.0 / 48.0 + th2 * (1.0 / 3840.0 + th2 * (-1.0 / 645120.0)));
    real = 1.0 + th2 * (-1.0 / 8.0 + th2 * (1.0 / 384.0 + th2 * (-1.0 / 46080.0)));
  } else {
    const double theta = sqrt(th2);
    double sn, cs;
    sincos(0.5 * theta, &sn, &cs);
    imag = sn / theta;
    real = cs;
  }
  *out = renorm(qmul_raw(*q, Quatd{real, imag * w[0], imag * w[1], imag * w[2]}));
}
// Log(q2^-1 * q1)
__device__ __noinline__ void rot_minus(const Quatd* q1, const Quatd* q2, double* o) {
  const Quatd r = qmul_raw(Quatd{q2->w, -q2->x, -q2->y, -q2->z}, *q1);
  const double n = sqrt((r.x * r.x + r.y * r.y) + r.z * r.z);
  double f;
  if (n < 1e-10) {
    f = 2.0 / r.w - 2.0 * (n * n) / (r.w * (r.w * r.w));
  } else if (fabs(r.w) < 1e-10) {
    f = (r.w > 0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
  } else {
    f = 2.0 * atan(n / r.w) / n;
  }
  o[0] = f * r.x;
  o[1] = f * r.y;
  o[2] = f * r.z;
}

// shared state of the solving block: the filter (x, x_propagated, P, loop state) and the per-update constants stay
// here for the whole update; the rest is scratch of one step
struct SolveSmem {
  double blob[LIO_BLOB];
  double warp_part[(THREADS / 32) * LIO_BLOB];  // per-warp sums of the grid reduction
  Ctrl sc;                                      // loop state (mirrored to SolveArgs::ctrl after every step)
  double S[144], Sinv[144], Kf[288], KH[288], dxn[24], dx[24];
  double W[12 * WS];   // [A | I] of block_inverse_spd
  double prior[288];   // P11^-1 (n x n) and P21 P11^-1 ((24-n) x n) of this update
  double P[576];
  double xa[26], xb[26], xn[26];  // x, x_propagated, x [+] dx
  int fin;
};

// Once per update: restore the prior (from_snapshot), x_propagated = x, loop state, P11^-1 and P21 P11^-1.
__device__ __noinline__ void block_prior(const SolveArgs& s, int n, SolveSmem* sm, const double* xsrc, const double* Psrc,
                                         bool restore) {
  const int tid = threadIdx.x;
#pragma unroll 1
  for (int k = tid; k < 576; k += THREADS) sm->P[k] = Psrc[k];
  if (tid < 26) sm->xa[tid] = xsrc[tid];
  __syncthreads();
  if (restore) {
    if (tid < 26) reinterpret_cast<double*>(s.x)[tid] = sm->xa[tid];
#pragma unroll 1
    for (int k = tid; k < 576; k += THREADS) s.P[k] = sm->P[k];
  }
  if (tid < 26) {
    reinterpret_cast<double*>(s.xprop)[tid] = sm->xa[tid];  // esekfom.hpp:287
    sm->xb[tid] = sm->xa[tid];
  }
  if (tid < 24) sm->dxn[tid] = 0.0;  // x [-] x_propagated of the first step: x IS x_propagated
  if (tid == 0) {
    Ctrl c0;
    c0.iter = -1;
    c0.converge = 1;
    c0.t = 0;
    c0.done = 0;
    c0.n_passes = 0;
    c0.n_valid_last = 0;
    c0.max_iter = s.max_iter;
    c0.pad = 0;
    sm->sc = c0;
    *s.ctrl = c0;
  }
#pragma unroll 1
  for (int k = tid; k < n * n; k += THREADS) sm->S[k] = sm->P[(k / n) * 24 + (k % n)];
  __syncthreads();
  if (tid < inv_threads(n)) block_inverse_spd(sm->S, sm->Sinv, n, sm->W);
  __syncthreads();
#pragma unroll 1
  for (int k = tid; k < n * n; k += THREADS) {
    s.prior[k] = sm->Sinv[k];
    sm->prior[k] = sm->Sinv[k];
  }
#pragma unroll 1
  for (int k = tid; k < (24 - n) * n; k += THREADS) {
    const int r = k / n, c = k % n;
    double acc = 0.0;
#pragma unroll 1
    for (int j = 0; j < n; ++j) acc = fma(sm->P[(n + r) * 24 + j], sm->Sinv[j * n + c], acc);
    s.prior[144 + k] = acc;
    sm->prior[144 + k] = acc;
  }
  __syncthreads();
}

// The stepwise driver keeps nothing in shared memory between its launches: filter state and constants come back from
// global memory in ONE round of loads (all issued before the first use); x [-] x_propagated is formed again from them
// (same function, same inputs, same bits as the copy the persistent kernel carries along).
__device__ __forceinline__ void step_boxminus(SolveSmem* sm);
__device__ __forceinline__ void solve_load_inputs(const SolveArgs& s, SolveSmem* sm) {
  const int tid = threadIdx.x;
  double vp[(576 + THREADS - 1) / THREADS], vq[(288 + THREADS - 1) / THREADS], va = 0.0, vb = 0.0;
#pragma unroll
  for (int u = 0; u < (576 + THREADS - 1) / THREADS; ++u)
    vp[u] = (tid + u * THREADS < 576) ? __ldcg(s.P + tid + u * THREADS) : 0.0;
#pragma unroll
  for (int u = 0; u < (288 + THREADS - 1) / THREADS; ++u)
    vq[u] = (tid + u * THREADS < 288) ? __ldcg(s.prior + tid + u * THREADS) : 0.0;
  if (tid < 26) {
    va = __ldcg(reinterpret_cast<const double*>(s.x) + tid);
    vb = __ldcg(reinterpret_cast<const double*>(s.xprop) + tid);
  }
  int cv = 0;
  if (tid < 8) cv = __ldcg(reinterpret_cast<const int*>(s.ctrl) + tid);
#pragma unroll
  for (int u = 0; u < (576 + THREADS - 1) / THREADS; ++u)
    if (tid + u * THREADS < 576) sm->P[tid + u * THREADS] = vp[u];
#pragma unroll
  for (int u = 0; u < (288 + THREADS - 1) / THREADS; ++u)
    if (tid + u * THREADS < 288) sm->prior[tid + u * THREADS] = vq[u];
  if (tid < 26) {
    sm->xa[tid] = va;
    sm->xb[tid] = vb;
  }
  if (tid < 8) reinterpret_cast<int*>(&sm->sc)[tid] = cv;
  __syncthreads();
  step_boxminus(sm);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// The Kalman step (esekfom.hpp:297-345).  It sits on the critical path of every pass (all workers wait for it), so what
// counts is the length of its dependent chain:
//   S = HtH[:n,:n] / R + P11^-1,  v = Hth / R + (HtH / R) dx_new[:n],  u = S^-1 v     one warp, registers + shuffles
//   dx[:n] = u - dx_new[:n],  dx[n:] = (P21 P11^-1) u - dx_new[n:]                       (esekfom.hpp:319, regrouped)
//   x = x [+] dx (the two rotations in two warps), loop state, publication of the new state to the workers;
//   x [-] x_propagated for the NEXT step and (after the last step) S^-1 for the covariance come after the publication.
// S is SPD (HtH / R is PSD, P11^-1 is PD): elimination needs no pivot search.
//
// Warp collectives and divergence.  A lane-dependent branch can leave a warp split for a long stretch of the code
// behind it (measured here: __activemask() == 1 some hundred instructions after an `if (lane == 0)`), and a split warp
// pays a rendezvous of ~750-1,400 cycles for every shuffle, vote or __syncwarp instead of ~30: the 6 x 6 elimination
// took 15,400 cycles that way and 980 in a converged warp.  So: collectives only in code that is reached straight
// from a block barrier, with no lane-dependent branch in between (selects instead); lane-specific work (one rotation,
// one flag) goes to OTHER warps between barriers.
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;
// lane i < N holds row i of A in a[] and of B in b[]; on return b = A^-1 B (row i in lane i).  Lanes >= N carry zero
// rows and only take part in the shuffles.  Branch-free.
template <int N, int NR>
__device__ __forceinline__ void warp_gauss_jordan(double (&a)[N], double (&b)[NR], int lane) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double pinv = 1.0 / __shfl_sync(FULL, a[k], k);
    const bool piv = lane == k;
    const double f = a[k];  // a_ik: this row's entry in the pivot column
#pragma unroll
    for (int c = k + 1; c < N; ++c) {
      const double t = a[c] * pinv;  // what the pivot row becomes (normalised); only lane k's is taken
      const double pc = __shfl_sync(FULL, t, k);
      a[c] = piv ? t : fma(-f, pc, a[c]);
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      const double t = b[c] * pinv;
      const double pc = __shfl_sync(FULL, t, k);
      b[c] = piv ? t : fma(-f, pc, b[c]);
    }
  }
}
__device__ __forceinline__ int blob_index(int r, int c) {  // upper triangle of the 12 x 12 HtH, row-major
  const int lo = r < c ? r : c, hi = r < c ? c : r;
  return lo * 12 - (lo * (lo - 1)) / 2 + (hi - lo);
}
// row `lane` of S = HtH[:N,:N] / R + P11^-1; lanes >= N read row 0 and get zeros (no branch)
template <int N>
__device__ __forceinline__ void load_S_row(const SolveSmem* sm, double inv_R, int lane, double (&a)[N]) {
  const int r = lane < N ? lane : 0;
  const double keep = lane < N ? 1.0 : 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = keep * fma(sm->blob[blob_index(r, j)], inv_R, sm->prior[r * N + j]);
}

// What the workers need of a step: x[0..13] = pos, rot, R_LI, t_LI as 28 stamped halves, and {converge, done} in word 28.
// PUB_COPIES copies 1 KB apart (different L2 slices): worker w polls copy w % PUB_COPIES, so that ~300 spinning blocks
// do not queue up on one slice.  Called by one converged warp.
constexpr int PUB_WORDS = 29;
constexpr int PUB_COPIES = 8;
constexpr int PUB_STRIDE = 128;  // 64-bit words between two copies
__device__ __forceinline__ void publish_state(unsigned long long* pub, const double* x14, int flags, unsigned target,
                                              int lane) {
  const int l = lane < PUB_WORDS ? lane : PUB_WORDS - 1;  // lanes 29..31 repeat word 28 (same value, same address)
  const unsigned long long b = (unsigned long long)__double_as_longlong(x14[l < 28 ? (l >> 1) : 0]);
  const unsigned half = l < 28 ? ((l & 1) ? (unsigned)(b >> 32) : (unsigned)b) : (unsigned)flags;
  const unsigned long long w = ((unsigned long long)target << 32) | half;
#pragma unroll
  for (int c = 0; c < PUB_COPIES; ++c)
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(pub + c * PUB_STRIDE + l), "l"(w) : "memory");
}

// u = S^-1 v into sm->S[0..N) (scratch), by warp 0, reached straight from a block barrier
template <int N>
__device__ __forceinline__ void warp_solve(SolveSmem* sm, double inv_R, int lane) {
  double a[N], v[1];
  load_S_row<N>(sm, inv_R, lane, a);
  {
    const int r = lane < N ? lane : 0;
    double acc = sm->blob[78 + r];
#pragma unroll
    for (int j = 0; j < N; ++j) acc = fma(sm->blob[blob_index(r, j)], sm->dxn[j], acc);
    v[0] = (lane < N ? 1.0 : 0.0) * (acc * inv_R);
  }
  warp_gauss_jordan<N, 1>(a, v, lane);
  if (lane < N) sm->S[lane] = v[0];  // (no collective behind this branch)
}
// S^-1 into sm->Sinv, by one warp reached straight from a block barrier (only after the last step: the covariance)
template <int N>
__device__ __forceinline__ void warp_invert(SolveSmem* sm, double inv_R, int lane) {
  double a[N], b[N];
  load_S_row<N>(sm, inv_R, lane, a);
#pragma unroll
  for (int j = 0; j < N; ++j) b[j] = (lane == j) ? 1.0 : 0.0;
  warp_gauss_jordan<N, N>(a, b, lane);
  if (lane < N) {
#pragma unroll
    for (int j = 0; j < N; ++j) sm->Sinv[lane * N + j] = b[j];
  }
}

// dx_new = x [-] x_propagated (esekfom.hpp:303) for the state in sm->xa: the two rotations in two warps, the vector parts
// in a third.  Whole block; the caller synchronises.
__device__ __forceinline__ void step_boxminus(SolveSmem* sm) {
  const int tid = threadIdx.x;
  if (tid == 0) rot_minus(reinterpret_cast<const Quatd*>(sm->xa + 3), reinterpret_cast<const Quatd*>(sm->xb + 3), sm->dxn + 3);
  if (tid == 32) rot_minus(reinterpret_cast<const Quatd*>(sm->xa + 7), reinterpret_cast<const Quatd*>(sm->xb + 7), sm->dxn + 6);
  if (tid >= 64 && tid < 64 + 24) {
    const int j = tid - 64;  // error-state index; state index: pos 0-2 | t_LI 11-13 | vel.. 14-25
    if (j < 3) sm->dxn[j] = sm->xa[j] - sm->xb[j];
    if (j >= 9) sm->dxn[j] = sm->xa[j + 2] - sm->xb[j + 2];
  }
}


// P = (I - K H) P (esekfom.hpp:342), K H = K_front[:, :n] HtH / R with K_front[:, :n] = [Sinv ; (P21 P11^-1) Sinv]; whole
// block, once per update (after the step that set sm->fin; sm->Sinv is in place).
__device__ __noinline__ void block_final_P(const SolveArgs& s, int n, SolveSmem* sm) {
  const int tid = threadIdx.x;
  const double inv_R = 1.0 / s.R;
#pragma unroll 1
  for (int k = tid; k < 24 * n; k += THREADS) {
    const int r = k / n, c = k % n;
    double acc;
    if (r < n) {
      acc = sm->Sinv[r * n + c];
    } else {
      acc = 0.0;
#pragma unroll 1
      for (int j = 0; j < n; ++j) acc = fma(sm->prior[144 + (r - n) * n + j], sm->Sinv[j * n + c], acc);
    }
    sm->Kf[k] = acc;
  }
  __syncthreads();
#pragma unroll 1
  for (int k = tid; k < 24 * n; k += THREADS) {
    const int r = k / n, c = k % n;
    double acc = 0.0;
#pragma unroll 1
    for (int j = 0; j < n; ++j) acc = fma(sm->Kf[r * n + j], sm->blob[blob_index(j, c)], acc);
    sm->KH[k] = acc * inv_R;
  }
  __syncthreads();
  double pnew[(576 + THREADS - 1) / THREADS];
#pragma unroll
  for (int u = 0; u < (576 + THREADS - 1) / THREADS; ++u) {
    const int k = tid + u * THREADS;
    pnew[u] = 0.0;
    if (k < 576) {
      const int r = k / 24, c = k % 24;
      double acc = sm->P[k];
#pragma unroll 1
      for (int j = 0; j < n; ++j) acc = fma(-sm->KH[r * n + j], sm->P[j * 24 + c], acc);
      pnew[u] = acc;
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < (576 + THREADS - 1) / THREADS; ++u) {
    const int k = tid + u * THREADS;
    if (k < 576) {
      s.P[k] = pnew[u];
      sm->P[k] = pnew[u];  // the shared copy stays current (host-direct path reads the posterior from it)
    }
  }
  __syncthreads();
}

// One Kalman step of the solving block on the reduced blob in sm->blob (esekfom.hpp:297-345); the caller has just
// synchronised the block.  On return the new state has been published (stamp `target`), mirrored to global memory, and
// sm->dxn holds x [-] x_propagated for the next step; after the last step P has been updated.
__device__ __noinline__ void block_step(const SolveArgs& s, int n, SolveSmem* sm, unsigned target) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int iter = sm->sc.iter, max_iter = sm->sc.max_iter, t_old = sm->sc.t, np_old = sm->sc.n_passes;
  const int n_valid = (int)sm->blob[90];
  const double inv_R = 1.0 / s.R;
  if (n_valid < 1) {
    // `if (!dyn_share.valid) continue;` (esekfom.hpp:297-299): nothing changes, the loop counter advances
    const int done = (iter + 1 >= max_iter) ? 1 : sm->sc.done;
    if (tid < 32) publish_state(s.pub, sm->xa, (sm->sc.converge ? 1 : 0) | (done ? 2 : 0), target, lane);
    if (tid >= 32 && tid < 32 + LIO_BLOB) s.blob[tid - 32] = sm->blob[tid - 32];
    __syncthreads();
    if (tid == 0) {
      sm->sc.n_valid_last = 0;
      sm->sc.n_passes = np_old + 1;
      sm->sc.iter = iter + 1;
      sm->sc.done = done;
      *s.ctrl = sm->sc;
      sm->fin = 0;
    }
    __syncthreads();
    return;
  }
  stamp(s.dbg, 128, 20);
  // ---- u = S^-1 v: warp 0, converged (straight from the caller's barrier)
  if (tid < 32) {
    if (n == 6)
      warp_solve<6>(sm, inv_R, lane);
    else
      warp_solve<12>(sm, inv_R, lane);
  }
  __syncthreads();
  stamp(s.dbg, 128, 21);
  // ---- dx (esekfom.hpp:319)
  if (tid < 24) {
    double acc;
    if (tid < n) {
      acc = sm->S[tid];
    } else {
      acc = 0.0;
#pragma unroll 1
      for (int j = 0; j < n; ++j) acc = fma(sm->prior[144 + (tid - n) * n + j], sm->S[j], acc);
    }
    sm->dx[tid] = acc - sm->dxn[tid];
  }
  __syncthreads();
  stamp(s.dbg, 128, 22);
  // ---- x = x [+] dx (esekfom.hpp:321): one rotation per warp, the vector parts and the loop state in two more
  if (tid == 0) rot_plus(reinterpret_cast<const Quatd*>(sm->xa + 3), sm->dx + 3, reinterpret_cast<Quatd*>(sm->xn + 3));
  if (tid == 32) rot_plus(reinterpret_cast<const Quatd*>(sm->xa + 7), sm->dx + 6, reinterpret_cast<Quatd*>(sm->xn + 7));
  if (tid >= 64 && tid < 64 + 24) {
    const int j = tid - 64;
    if (j < 3) sm->xn[j] = sm->xa[j] + sm->dx[j];
    if (j >= 9) sm->xn[j + 2] = sm->xa[j + 2] + sm->dx[j];
  }
  if (tid == 96) {
    bool converge = true;  // esekfom.hpp:324-330: every one of the 24 components
#pragma unroll 1
    for (int jj = 0; jj < 24; ++jj)
      if (fabs(sm->dx[jj]) > 0.001) {
        converge = false;
        break;
      }
    int t = t_old;
    if (converge) t++;
    if (!t && iter == max_iter - 2) converge = true;  // forced re-search (esekfom.hpp:333)
    const int fin = (t > 1 || iter == max_iter - 1) ? 1 : 0;
    sm->sc.converge = converge ? 1 : 0;
    sm->sc.t = t;
    sm->sc.n_valid_last = n_valid;
    sm->sc.n_passes = np_old + 1;
    sm->sc.iter = iter + 1;
    if (fin) sm->sc.done = 1;
    sm->fin = fin;
  }
  __syncthreads();
  // ---- the workers are waiting for exactly this
  if (tid < 32) publish_state(s.pub, sm->xn, (sm->sc.converge ? 1 : 0) | (sm->fin ? 2 : 0), target, lane);
  stamp(s.dbg, 128, 23);
  // ---- off the workers' critical path from here on
  if (tid >= 32 && tid < 32 + 26) {
    const double xv = sm->xn[tid - 32];
    sm->xa[tid - 32] = xv;  // the next step starts here
    reinterpret_cast<double*>(s.x)[tid - 32] = xv;
  }
  if (tid >= 64 && tid < 64 + 24 && s.dx_out) s.dx_out[tid - 64] = sm->dx[tid - 64];
  if (tid >= 96 && tid < 96 + LIO_BLOB) s.blob[tid - 96] = sm->blob[tid - 96];
  if (tid == 0) *s.ctrl = sm->sc;
  __syncthreads();
  if (sm->fin) {
    // S^-1 itself is only needed for the covariance, once, after the last step.  (Forming it in the idle warp next to
    // every step's solve was tried: the six right-hand sides take longer than the one of the solve and the barrier
    // behind it waits for both: +0.5 us per pass.)
    if (tid < 32) {
      if (n == 6)
        warp_invert<6>(sm, inv_R, lane);
      else
        warp_invert<12>(sm, inv_R, lane);
    }
    __syncthreads();
    block_final_P(s, n, sm);
  } else {
    step_boxminus(sm);  // dx_new of the next step, while the workers run their pass
    __syncthreads();
  }
}

struct __align__(16) PassSmem {
  double rows[ROWS_MAX * RS];
  double acc[THREADS];
  float4 nb[SROWS_MAX * LIO_K];  // neighbours found by the search phase of the current tile
  PassConst pc;
  int cnt[SROWS_MAX];            // gate 1 of the search phase
  float4 body_row[SROWS_MAX];    // the search tile's scan points, for its finish phase
  float4 q[SROWS_MAX];           // ... and their FP32 p_world: the queries
  uint2 cells[SROWS_MAX * KNN_CELLS];  // group_knn5: {end, base} of the non-empty buckets of every group's round
  unsigned short oab[THREADS];         // thread (output o, segment): row columns (a | b << 8) of its product (fill_oab)
  unsigned char valid[ROWS_MAX];
  double xrecv[15];  // worker_receive: the published x[0..13] and, in the low half of [14], the flags
  int flag;
  int n_own;                      // sharded map: length of `own` (-1: not built yet)
  int wcnt[THREADS / 32];
  unsigned short own[OWN_MAX];    // offsets (in the chunk) of the points this rank owns, ascending
};

static_assert(SROWS_MAX == ST_TILE, "a search tile is a staging tile");
// dynamic shared memory of the pass / update kernels: {PassSmem | SolveSmem} then the staging area of the searches
constexpr size_t SMEM_STAGE_OFF =
    ((sizeof(PassSmem) > sizeof(SolveSmem) ? sizeof(PassSmem) : sizeof(SolveSmem)) + 127) & ~(size_t)127;
constexpr size_t SMEM_PASS_BYTES = SMEM_STAGE_OFF + sizeof(StageSmem);  // with staging; SMEM_STAGE_OFF without
static size_t pass_smem_bytes(const lio_ctx* c) { return c->stage_search ? SMEM_PASS_BYTES : SMEM_STAGE_OFF; }
// Sharded map: the points of this block the rank owns = whose p_world.x AT THE LAST SEARCH PASS lies in its region.
// Every rank computes the same bits for every point, so the ranks' lists partition the scan.  The list is rebuilt in
// search passes (when the positions change) and kept across the cached passes that follow (the persistent kernel keeps
// it in shared memory; a per-pass launch rebuilds it from a.near_q).  Out of line: the unsharded update never fetches
// this code (a pass runs from a cold instruction cache, its cost is its code size).
__device__ __noinline__ void build_own_list(const PassArgs& a, PassSmem* ps, const Chunk<true> ch, const float4* body, int C,
                                            int M, bool search) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int built = 0;
#pragma unroll 1
  for (int k0 = 0; k0 < C; k0 += THREADS) {
    const int k = k0 + tid, i = ch.point(k);
    const bool in = k < C && i < M;
    float qx = 0.f;
    if (in) {
      if (search) {
        const float4 b = body[i];
        const double pb[3] = {b.x, b.y, b.z};
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        body_to_world(ps->pc, pb, q.x, q.y, q.z);
        a.near_q[i] = q;
        qx = q.x;
      } else {
        qx = __ldcg(&a.near_q[i]).x;
      }
    }
    const bool own = in && owns_row(a, qx);
    if (in && !own && search) {  // another rank's row: nothing cached here
      a.selected[i] = 0;
      a.near_cnt[i] = 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, own);
    if (lane == 0) ps->wcnt[warp] = __popc(bal);
    __syncthreads();
    int base = built, tot = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
      if (w < warp) base += ps->wcnt[w];
      tot += ps->wcnt[w];
    }
    if (own) ps->own[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)k;
    built += tot;
    __syncthreads();
  }
  if (tid == 0) ps->n_own = built;
  __syncthreads();
}

// ps->pc holds the per-pass constants (the caller loads or computes them and synchronises the block).
// Worker `wid` of `nworkers` runs its chunk of the scan (chunk_points) and leaves its partial blob, stamped `target`, in
// row wid of a.partials; a worker whose chunk is empty writes nothing (the reduction knows: workers_used).
// once per kernel, before the first pass (the caller synchronises the block)
__device__ __forceinline__ void fill_oab(const PassArgs& a, PassSmem* ps) {
  const int nout = a.extrinsic_est ? NOUT_EXT : NOUT_NOEXT;
  const int o = threadIdx.x % nout;
  ps->oab[threadIdx.x] = a.extrinsic_est ? (unsigned short)(c_oa_ext[o] | (c_ob_ext[o] << 8))
                                          : (unsigned short)(c_oa_no[o] | (c_ob_no[o] << 8));
}

template <bool SH, bool STG>
__device__ void block_pass(const PassArgs& a, bool search, PassSmem* ps, StageSmem* st, int nworkers, int wid,
                           bool first_pass, unsigned target) {
  const int tid = threadIdx.x;
  const bool from_host = first_pass && a.body_src != nullptr;  // pass 0 always searches
  const float4* body = from_host ? a.body_src : a.body;
  const int M = scan_size(a);
  stamp(a.dbg, 0, 2);
  const int C = chunk_points(M, nworkers);
  const int beg = wid * C;
  const bool compact = SH && sharded_lists(a, C);
  const bool dealt = SH && dealt_runs(a, C);
  if (!dealt && beg >= M) return;  // block-uniform
  const int end = min(M, beg + C);
  Chunk<SH> ch{beg, wid, nworkers, compact ? ps->own : nullptr, dealt};
  if (SH) {
    if (compact && (search || ps->n_own < 0))
      build_own_list(a, ps, Chunk<true>{beg, wid, nworkers, ps->own, dealt}, body, C, M, search);
  }
  const int n = compact ? ps->n_own : (dealt ? dealt_count(M, C, nworkers, wid) : end - beg);
  const int G = pick_group(n);
  const int nout = a.extrinsic_est ? NOUT_EXT : NOUT_NOEXT;
  const int nseg = THREADS / nout;
  const int o = tid % nout, seg = tid / nout;
  int ca = 0, cb = 0;
  if (seg < nseg) {
    // (from shared memory, filed once per kernel: indexed by thread, the constant bank serves a warp's 28 addresses one
    // after the other -- 5 % of a pass's stall samples sat on these two loads)
    ca = ps->oab[tid] & 0xff;
    cb = ps->oab[tid] >> 8;
  }
  double acc = 0.0;
  // A block whose share is one search tile (the latency-bound bench case) searches and finishes it out of shared memory.
  // With several tiles (dense scans, many sequences per launch) all of them are searched first and the rows are then
  // finished ROWS_MAX at a time from the neighbour cache: one thread per row keeps 256 threads busy with the plane fits
  // instead of THREADS / G (64) between every two searches.
  const int sstep = THREADS / G;
  const int mode = !search ? 0 : ((n > sstep && a.batched_finish) ? 2 : 1);
  const int step = mode == 1 ? sstep : (n < ROWS_MAX ? max(n, 1) : ROWS_MAX);
#pragma unroll 1
  for (int phase = (mode == 2 ? 0 : 1); phase < 2; ++phase) {
    int pstep = phase == 0 ? sstep : step;
#pragma unroll 1
    for (int t0 = 0; t0 < n; t0 += pstep) {
      // the last search tile of a block with several (mode 2) takes the widest groups that fit what is left: 7 queries
      // behind two full tiles of 64 are searched by 32 lanes each, not by 8
      const int Gt = phase == 0 ? pick_group(n - t0) : G;
      if (phase == 0) pstep = THREADS / Gt;
      const int rows = pstep;
      if (phase == 0 || mode == 1) {
        const int tend = min(n, t0 + THREADS / Gt);
        // the tile's queries (p_world, FP64 -> FP32) and, when staging, an empty cell set; the previous tile is done with
        // the staging area
        if (STG && a.stage) {
          for (int h = tid; h < ST_HASH; h += THREADS) st->key[h] = LIO_EMPTY_KEY;
          if (tid == 0) {
            st->n_list = 0;
            st->n_pts = 0;
            st->overflow = 0;
          }
        }
        bool use_stage = false;
        if (STG && a.stage) {
          if (tid < tend - t0) {  // the staging needs the queries first; search_tile recomputes the same bits
            const float4 b = body[ch.at(t0 + tid)];
            const double pb[3] = {b.x, b.y, b.z};
            float pwx, pwy, pwz;
            body_to_world(ps->pc, pb, pwx, pwy, pwz);
            ps->q[tid] = make_float4(pwx, pwy, pwz, 0.f);
          }
          __syncthreads();
          stage_cells<THREADS>(a.map, st, ps->q, tend - t0, a.dbg);
          use_stage = st->overflow == 0;
        }
        const bool cp = from_host && mode == 2;
        if (Gt == 32)
          search_tile<32, SH, STG>(a, ps->pc, t0, n, ch, ps->nb, ps->cnt, ps->body_row, ps->cells, body, st, use_stage, cp);
        else if (Gt == 16)
          search_tile<16, SH, STG>(a, ps->pc, t0, n, ch, ps->nb, ps->cnt, ps->body_row, ps->cells, body, st, use_stage, cp);
        else
          search_tile<8, SH, STG>(a, ps->pc, t0, n, ch, ps->nb, ps->cnt, ps->body_row, ps->cells, body, st, use_stage, cp);
        __syncthreads();
        if (STG && tid == 0 && a.stage) st->phase ^= 1;  // the mbarrier's next phase (read again only behind the next tile's barriers)
        stamp(a.dbg, 0, 3);
        if (tid == THREADS - 1 && wid < 256) stamp_slot(a.dbg, 256 + wid);
      }
      if (phase == 0) continue;
      finish_tile<SH>(a, ps->pc, t0, n, ch, rows, mode, ps->nb, ps->cnt, ps->rows, ps->valid, ps->body_row,
                      from_host && mode == 1);
      __syncthreads();
      stamp(a.dbg, 0, 4);
      if (seg < nseg) {
#pragma unroll 1
        for (int r = seg; r < rows; r += nseg)
          if (ps->valid[r]) acc = fma(ps->rows[r * RS + ca], ps->rows[r * RS + cb], acc);
      }
      __syncthreads();
      stamp(a.dbg, 0, 5);
    }
  }
  if (seg < nseg) ps->acc[seg * nout + o] = acc;
  __syncthreads();
  if (tid < nout) {
    double sum = 0.0;
    for (int g = 0; g < nseg; ++g) sum += ps->acc[g * nout + tid];
    st_stamped(a.partials + (size_t)wid * ROW_WORDS + 2 * tid, sum, target);  // outputs back to back
  }
  if (tid == THREADS - 1 && wid < 256) stamp_slot(a.dbg, wid);
  stamp(a.dbg, 0, 6);
}

// ---------------------------------------------------------------------------------------------------------
// Sharded map (SURVEY.md 8e case 2): the ranks' blobs are exchanged INSIDE the persistent kernel over NVLink peer
// memory.  Every rank owns a mailbox of stamped words {words[world][2][LIO_BLOB][2]} that its peers have mapped
// (cudaIpc).  After its local reduction the solver block stores its 92 doubles into slot [rank][pass & 1] of every
// rank's mailbox as self-validating words (st_stamped at system scope: value and stamp travel in the same 8-byte store),
// then polls its own mailbox until every word of every rank carries this pass's stamp and adds the blobs in RANK ORDER,
// so that every rank forms the same bits and performs the identical Kalman step.  No fence, no flag, no second hop: the
// exchange costs one NVLink store latency plus the skew between the ranks.  Two slots suffice: a peer can only write
// pass p + 2 after it has received this rank's pass p + 1 blob, which is sent after pass p has been read.  No NCCL
// call, no launch, no host on the path.
// ---------------------------------------------------------------------------------------------------------
constexpr int LIO_MAX_RANKS = 8;
struct ShardArgs {
  int world, rank;
  unsigned long long* mbox[LIO_MAX_RANKS];  // mailbox words of every rank as mapped in THIS process ([rank] = own)
  int* err;                                 // set to 1 when a peer did not show up in time
  unsigned epoch;                           // stamp base of this sharded launch (advances in lockstep on all ranks)
};

__device__ __forceinline__ void st_stamped_sys(unsigned long long* p, double v, unsigned stamp) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long hs = (unsigned long long)stamp << 32;
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(hs | (b & 0xffffffffull)), "l"(hs | (b >> 32))
               : "memory");
}
__device__ __forceinline__ bool ld_stamped_sys(const unsigned long long* p, unsigned stamp, double& v) {
  unsigned long long lo, hi;
  asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
  return (unsigned)(lo >> 32) == stamp && (unsigned)(hi >> 32) == stamp;
}

__device__ __noinline__ void block_exchange(const ShardArgs& sh, unsigned target, int pass_no, double* s_blob,
                                            double* s_in) {
  const int tid = threadIdx.x;
  const int slot = pass_no & 1;
  const int total = sh.world * LIO_BLOB;
  // 1. this rank's blob into everybody's mailbox (own included), the remote ones first
#pragma unroll 1
  for (int k = tid; k < total; k += THREADS) {
    const int pp = k / LIO_BLOB, e = k - pp * LIO_BLOB;
    const int p = (sh.rank + 1 + pp) % sh.world;
    st_stamped_sys(sh.mbox[p] + (((size_t)sh.rank * 2 + slot) * LIO_BLOB + e) * 2, s_blob[e], target);
  }
  // 2. every rank's blob out of the own mailbox, each word polled until it carries the stamp (bounded: a missing peer
  //    must not hang the GPU).  Uniform control flow per warp (see block_step on split warps).
  const unsigned long long* mine = sh.mbox[sh.rank];
#pragma unroll 1
  for (int k0 = 0; k0 < total; k0 += THREADS) {
    const int k = min(k0 + tid, total - 1);
    const int r = k / LIO_BLOB, e = k - r * LIO_BLOB;
    const unsigned long long* w = mine + (((size_t)r * 2 + slot) * LIO_BLOB + e) * 2;
    double v;
    bool ok;
    long long spins = 0;
    do {
      ok = ld_stamped_sys(w, target, v);
      if (!ok && ++spins > 200000000LL) {
        *sh.err = 1;
        ok = true;
      }
    } while (!__all_sync(0xffffffffu, ok));
    if (k0 + tid < total) s_in[k] = v;
  }
  __syncthreads();
  // 3. sum in rank order
  if (tid < LIO_BLOB) {
    double acc = 0.0;
#pragma unroll 1
    for (int r = 0; r < sh.world; ++r) acc += s_in[r * LIO_BLOB + tid];
    if (tid == 91) acc = s_blob[91];
    s_blob[tid] = acc;
  }
  __syncthreads();
}

// Host-direct path (lio_update_scan_host): nothing is copied before or after the kernel.  The prior travels in the
// kernel parameters, the scan is read by pass 0 straight from the caller's pinned buffer, and the posterior is written
// to mapped pinned memory followed by a sequence word the host spins on.
struct HostPath {
  int use_param_prior;      // 1: the prior is x0 / P0 below
  double* host_out;         // mapped pinned {x 26, P 576, ctrl 4 (as 8 ints)} or nullptr
  unsigned long long* host_flag;
  unsigned long long seq;
  double x0[26];
  double P0[576];
};

// Workers: wait for the publication of step `target` (warp 0: one stamped word per lane, polled until its stamp shows
// up -- data and flag arrive in the same word, one L2 round trip after the solver's store), rebuild the per-pass
// constants from the 14 doubles, hand {converge, done} to the block.
__device__ __forceinline__ void worker_receive(const SolveArgs& s, PassSmem* ps, unsigned target, int* flags_out,
                                               int wid);

// The whole update_iterated_dyn_share_modified loop (esekfom.hpp:270-346).  Cooperative launch: all blocks resident.
// Blocks 0 .. gridDim-2 are WORKERS (h_share_model passes over their own chunk of the scan); the last block is the
// SOLVER: it keeps the filter state in shared memory for the whole update, sums the workers' partial rows as they
// arrive, performs the Kalman step (one warp) and publishes the new state.  Rows and publication are stamped words
// (target = epoch + pass + 1): nothing is zeroed between launches and there is no fence on the per-pass path.
// nblk / bid: size of the (sub-)grid that works on this update and this block's index in it -- the whole grid for a
// single update, a slice of it when several independent sequences share one launch (update_kernel_multi).
template <bool HOST, bool SH, bool STG>
__device__ __forceinline__ void update_body(const PassArgs& a, const SolveArgs& s, const unsigned epoch,
                                            const ShardArgs& sh, const HostPath& hp, const int nblk, const int bid) {
  extern __shared__ __align__(128) unsigned char smem_raw[];  // SMEM_PASS_BYTES
  const int tid = threadIdx.x;
  const int n = a.extrinsic_est ? 12 : 6;
  const int nworkers = nblk - 1;
  if (bid == nworkers) {
    // ---------------------------------------------------------------- solver
    SolveSmem& ss = *reinterpret_cast<SolveSmem*>(smem_raw);
    if (HOST && hp.use_param_prior)
      block_prior(s, n, &ss, hp.x0, hp.P0, true);
    else
      block_prior(s, n, &ss, reinterpret_cast<const double*>(s.from_snapshot ? s.x0 : s.x),
                  s.from_snapshot ? s.P0 : s.P, s.from_snapshot != 0);
    if (tid == 0 && a.m_value >= 0) *const_cast<int*>(a.scan_m) = a.m_value;  // later calls read the device copy
    bool search = true;
    for (int pass_no = 0; pass_no <= s.max_iter; ++pass_no) {
      const unsigned target = epoch + (unsigned)pass_no + 1u;
      stamp(a.dbg, 128, 10);
      block_reduce_partials(a, search, nworkers, target, ss.blob, ss.warp_part);
      if (sh.world > 1) block_exchange(sh, sh.epoch + (unsigned)pass_no + 1u, pass_no, ss.blob, ss.warp_part);
      stamp(a.dbg, 128, 11);
      block_step(s, n, &ss, target);
      stamp(a.dbg, 128, 12);
      search = ss.sc.converge != 0;
      if (ss.sc.done) break;
      __syncthreads();
    }
    if (HOST && hp.host_out != nullptr) {
      // posterior straight into the host's mapped buffer, then the sequence word it spins on
      __syncthreads();
      for (int k = tid; k < 26; k += THREADS) hp.host_out[k] = ss.xa[k];
      for (int k = tid; k < 576; k += THREADS) hp.host_out[26 + k] = ss.P[k];
      if (tid < 8) reinterpret_cast<int*>(hp.host_out + 602)[tid] = reinterpret_cast<const int*>(&ss.sc)[tid];
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        *reinterpret_cast<volatile unsigned long long*>(hp.host_flag) = hp.seq;
        __threadfence_system();
      }
    }
    return;
  }
  // ------------------------------------------------------------------ workers
  PassSmem& ps = *reinterpret_cast<PassSmem*>(smem_raw);
  StageSmem* st = reinterpret_cast<StageSmem*>(smem_raw + SMEM_STAGE_OFF);
  const int wid = bid;
  if (STG && tid == 0 && a.stage) {  // (without staging the area is not even allocated)
    mbar_init(&st->mbar, THREADS);
    st->phase = 0;
  }
  const StateD* x_first =
      (HOST && hp.use_param_prior) ? reinterpret_cast<const StateD*>(hp.x0) : (s.from_snapshot ? s.x0 : s.x);
  stamp(a.dbg, 0, 1);
  if (tid == 0) {
    load_pass_const(x_first, ps.pc);  // pass 0 searches at the prior
    ps.n_own = -1;
  }
  fill_oab(a, &ps);
  __syncthreads();
  bool search = true;
  for (int pass_no = 0; pass_no <= s.max_iter; ++pass_no) {
    const unsigned target = epoch + (unsigned)pass_no + 1u;
    block_pass<SH, STG>(a, search, &ps, st, nworkers, wid, pass_no == 0, target);
    stamp(a.dbg, 0, 7);
    worker_receive(s, &ps, target, &ps.flag, wid);
    __syncthreads();
    stamp(a.dbg, 0, 8);
    const int fl = ps.flag;
    search = (fl & 1) != 0;
    __syncthreads();  // ps.flag / ps.pc are rewritten by the next round
    if (fl & 2) break;
  }
}

__device__ __forceinline__ void worker_receive(const SolveArgs& s, PassSmem* ps, unsigned target, int* flags_out,
                                               int wid) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const unsigned long long* pub = s.pub + (wid % PUB_COPIES) * PUB_STRIDE;
  // Uniform control flow only (see block_step on split warps): every lane watches its own word (232 bytes per round
  // trip and block) and the warp leaves the loop together once all of them carry the stamp -- each word validates
  // itself, so the order in which the solver's stores become visible does not matter.
  unsigned long long w;
  const int l = lane < PUB_WORDS ? lane : PUB_WORDS - 1;
  bool ok;
  do {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(pub + l) : "memory");
    ok = (unsigned)(w >> 32) == target;
  } while (!__all_sync(FULL, ok));
  reinterpret_cast<unsigned*>(ps->xrecv)[l] = (unsigned)w;  // halves 0..27 are x[0..13]; word 28 lands in xrecv[14]'s low half
  __syncwarp();
  const double* x14 = ps->xrecv;
  PassConst& pc = ps->pc;
  if (lane == 0) {
    pc.pos[0] = x14[0]; pc.pos[1] = x14[1]; pc.pos[2] = x14[2];
    pc.rot = Quatd{x14[3], x14[4], x14[5], x14[6]};
    quat_to_mat(pc.rot, pc.Rt);
    *flags_out = (int)reinterpret_cast<const unsigned*>(ps->xrecv)[28];
  }
  if (lane == 1) {
    pc.rli = Quatd{x14[7], x14[8], x14[9], x14[10]};
    pc.tli[0] = x14[11]; pc.tli[1] = x14[12]; pc.tli[2] = x14[13];
    quat_to_mat(pc.rli, pc.Rli);
  }
}

__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s,
                                                                            const unsigned epoch, const __grid_constant__ ShardArgs sh) {
  update_body<false, false, false>(a, s, epoch, sh, *reinterpret_cast<const HostPath*>(&a), (int)gridDim.x,
                                   (int)blockIdx.x);  // the host path is compiled out
}
// the instantiation with ownership lists and dealt runs: sharded map (windows or stripes), or LIO_INTERLEAVE=1
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel_sh(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s,
                                                                               const unsigned epoch, const __grid_constant__ ShardArgs sh) {
  update_body<false, true, false>(a, s, epoch, sh, *reinterpret_cast<const HostPath*>(&a), (int)gridDim.x, (int)blockIdx.x);
}
// the same loop with the host-direct prologue / epilogue; its 4.9 KB of extra parameters are only paid by that path
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel_host(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s,
                                                                                 const unsigned epoch, const __grid_constant__ ShardArgs sh,
                                                                                 const __grid_constant__ HostPath hp) {
  update_body<true, false, false>(a, s, epoch, sh, hp, (int)gridDim.x, (int)blockIdx.x);
}
// LIO_STAGE_SEARCH=1: the same two kernels with the shared-memory staging of a search tile's neighbour cells compiled in
// (stage_cells; measured slower than the direct search, DESIGN.md 4.1, so the default instantiations do not carry its
// code: a pass runs straight-line code from a cold instruction cache and its registers are the kernel's)
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel_stage(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s,
                                                                                  const unsigned epoch, const __grid_constant__ ShardArgs sh) {
  update_body<false, false, true>(a, s, epoch, sh, *reinterpret_cast<const HostPath*>(&a), (int)gridDim.x, (int)blockIdx.x);
}
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel_host_stage(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s,
                                                                                       const unsigned epoch,
                                                                                       const __grid_constant__ ShardArgs sh,
                                                                                       const __grid_constant__ HostPath hp) {
  update_body<true, false, true>(a, s, epoch, sh, hp, (int)gridDim.x, (int)blockIdx.x);
}

// Several INDEPENDENT updates (different sequences, each with its own map, scan and filter: BASELINE.json config 4) in
// one cooperative launch: the grid is cut into equal slices, one per update, each with its own workers and solver.
// A single update is latency-bound and leaves most issue slots of the GPU idle; slices fill them.
constexpr int LIO_MAX_MULTI = 8;
struct MultiArgs {
  int n;
  unsigned epoch[LIO_MAX_MULTI];
  PassArgs a[LIO_MAX_MULTI];
  SolveArgs s[LIO_MAX_MULTI];
};
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) update_kernel_multi(const __grid_constant__ MultiArgs m) {
  const int per = (int)gridDim.x / m.n;
  const int q = (int)blockIdx.x / per;
  if (q >= m.n) return;  // left-over blocks when the grid does not divide evenly
  ShardArgs sh;
  sh.world = 1;
  update_body<false, false, false>(m.a[q], m.s[q], m.epoch[q], sh, *reinterpret_cast<const HostPath*>(&m), per,
                                   (int)blockIdx.x - q * per);
}

// One pass at the state in s.x; the last block to finish reduces the partials into s.blob (same worker split and
// the same summation order as update_kernel, so the stepwise driver reproduces its bits).
// mode: 0 cached, 1 search, -1 as the loop state says (sharded driver).
template <bool SH>
__global__ void __launch_bounds__(THREADS, LIO_BLOCKS_PER_SM) pass_kernel(const __grid_constant__ PassArgs a, const __grid_constant__ SolveArgs s, int mode,
                                                                          unsigned target) {
  extern __shared__ __align__(128) unsigned char smem_raw[];  // SMEM_PASS_BYTES
  PassSmem& ps = *reinterpret_cast<PassSmem*>(smem_raw);
  StageSmem* st = reinterpret_cast<StageSmem*>(smem_raw + SMEM_STAGE_OFF);
  // the reduction of the last block runs after its pass: its scratch lies over the row staging area
  static_assert(sizeof(ps.rows) >= sizeof(double) * (THREADS / 32) * LIO_BLOB && THREADS >= LIO_BLOB, "reduce scratch");
  double* s_warp = ps.rows;
  double* s_blob = ps.acc;
  const int tid = threadIdx.x;
  const int nworkers = (int)gridDim.x - 1;
  if (mode < 0 && s.ctrl->done) return;
  const bool search = mode < 0 ? (s.ctrl->converge != 0) : (mode != 0);
  if ((int)blockIdx.x < nworkers) {
    if (tid == 0) {
      load_pass_const(s.x, ps.pc);
      ps.n_own = -1;  // (a launch per pass: the ownership list is rebuilt from a.near_q)
    }
    fill_oab(a, &ps);
    __syncthreads();
    block_pass<SH, false>(a, search, &ps, st, nworkers, (int)blockIdx.x, false, target);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(&s.sync[0], 1u);
    ps.flag = (ticket == gridDim.x - 1u) ? 1 : 0;
  }
  __syncthreads();
  stamp(a.dbg, 0, 7);
  if (ps.flag) {
    __threadfence();
    stamp(a.dbg, 128, 10);
    block_reduce_partials(a, search, nworkers, target, s_blob, s_warp);
    stamp(a.dbg, 128, 11);
    if (tid < LIO_BLOB) s.blob[tid] = s_blob[tid];
    if (tid == 0) s.sync[0] = 0;  // ready for the next pass launch
  }
}

// Kalman step from the blob in s.blob (already summed over ranks by the sharded driver).
__global__ void __launch_bounds__(THREADS) solve_kernel(const __grid_constant__ SolveArgs s, int extrinsic_est, unsigned target) {
  __shared__ SolveSmem ss;
  if (s.ctrl->done) return;
  solve_load_inputs(s, &ss);
  if (threadIdx.x < LIO_BLOB) ss.blob[threadIdx.x] = s.blob[threadIdx.x];
  __syncthreads();
  block_step(s, extrinsic_est ? 12 : 6, &ss, target);
}

__global__ void __launch_bounds__(THREADS) begin_kernel(const __grid_constant__ SolveArgs s, int extrinsic_est) {
  __shared__ SolveSmem ss;
  block_prior(s, extrinsic_est ? 12 : 6, &ss, reinterpret_cast<const double*>(s.from_snapshot ? s.x0 : s.x),
              s.from_snapshot ? s.P0 : s.P, s.from_snapshot != 0);
}

// Stand-alone batch of Nearest_Search calls (lio_knn5): one 8-lane group per query.
__global__ void __launch_bounds__(256) knn_batch_kernel(MapView map, const float4* q, int m, float max_d2, int rings,
                                                        float4* near_pts, float* near_d2, int* near_cnt) {
  constexpr int G = 8;
  __shared__ uint2 s_cells[(256 / G) * KNN_CELLS];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int gglobal = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int ngroups = (gridDim.x * blockDim.x) / G;
  const int iters = (m + ngroups - 1) / ngroups;  // uniform trip count: the warp stays together for the shuffles
  for (int it = 0; it < iters; ++it) {
    const int i_raw = gglobal + it * ngroups;
    const bool act = i_raw < m;
    const int i = act ? i_raw : m - 1;
    const float4 p = __ldg(q + i);
    unsigned long long ok[LIO_K];
    uint32_t os[LIO_K];
    const int f = group_knn5<G>(map, p.x, p.y, p.z, max_d2, rings, gl, s_cells + (threadIdx.x / G) * KNN_CELLS, ok, os);
    if (!act) continue;
    if (gl < LIO_K) {
      unsigned long long k = ok[0];
      uint32_t sl = os[0];
#pragma unroll
      for (int r = 1; r < LIO_K; ++r)
        if (gl == r) {
          k = ok[r];
          sl = os[r];
        }
      float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      float d = CUDART_INF_F;
      if (gl < f) {
        v = __ldg(map.pool + sl);
        d = __uint_as_float((uint32_t)(k >> 32));
      }
      near_pts[(size_t)i * LIO_K + gl] = v;
      near_d2[(size_t)i * LIO_K + gl] = d;
    }
    if (gl == 0) near_cnt[i] = f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Unbounded completion of neighbour rows (esekfom.hpp:140-141: Nearest_Search gets no max_dist, so the reference's
// Nearest_Points[i] holds min(5, #live points) neighbours however far away).  near_cnt[i] is the length of the row's
// KNOWN PREFIX of that list: the bounded search of the update leaves the neighbours within d2 <= 5 (all of them when
// it found five), far_search_kernel extends the short rows to `need` entries.
// ---------------------------------------------------------------------------------------------------------
// One warp per block, rows dealt to the warps with the grid's stride: a warp looks at the count of each of its rows and
// completes the short ones (they cluster in index -- a frontier is contiguous in voxel order -- and the stride spreads them
// over the warps; no list, no second launch).
__global__ void __launch_bounds__(32) far_search_kernel(const __grid_constant__ MapView map, const float4* q, const int* scan_m, int m_value,
                                                        int min_m, int need, float4* near_pts, float* near_d2,
                                                        int* near_cnt) {
  const int lane = threadIdx.x;  // (warp_knn_far synchronises the block)
  int M = m_value >= 0 ? m_value : *scan_m;
  if (M < min_m) M = 0;
#pragma unroll 1
  for (int i = blockIdx.x; i < M; i += gridDim.x) {
    if (__ldcg(near_cnt + i) >= need) continue;  // block-uniform
    const float4 p = __ldg(q + i);
    unsigned long long key[LIO_K];
    uint32_t slot[LIO_K];
    int found = warp_knn_far(map, p.x, p.y, p.z, need, key, slot);
    if (found > need) found = need;  // only the first `need` entries are the true nearest ones
    if (lane < LIO_K) {
      unsigned long long k = key[0];
      uint32_t sl = slot[0];
#pragma unroll
      for (int r = 1; r < LIO_K; ++r)
        if (lane == r) {
          k = key[r];
          sl = slot[r];
        }
      float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      float d = CUDART_INF_F;
      if (lane < found) {
        v = __ldg(map.pool + sl);
        d = __uint_as_float((uint32_t)(k >> 32));
      }
      near_pts[(size_t)i * LIO_K + lane] = v;
      near_d2[(size_t)i * LIO_K + lane] = d;
    }
    if (lane == 0) near_cnt[i] = found;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
static bool g_tables_ready[64] = {false};
int ensure_tables(lio_ctx* c) {
  if (c->device < 64 && g_tables_ready[c->device]) return LIO_OK;
  unsigned char oa[NOUT_EXT], ob[NOUT_EXT], oe[NOUT_EXT], na[NOUT_NOEXT], nb[NOUT_NOEXT], ne[NOUT_NOEXT];
  int e = 0, k = 0;
  for (int a = 0; a < 12; ++a)
    for (int b = a; b < 12; ++b) {
      oa[e] = (unsigned char)a;
      ob[e] = (unsigned char)b;
      oe[e] = (unsigned char)e;
      if (b < 6) {
        na[k] = (unsigned char)a;
        nb[k] = (unsigned char)b;
        ne[k] = (unsigned char)e;
        ++k;
      }
      ++e;
    }
  for (int a = 0; a < 12; ++a) {
    oa[78 + a] = (unsigned char)a;
    ob[78 + a] = 12;
    oe[78 + a] = (unsigned char)(78 + a);
    if (a < 6) {
      na[k] = (unsigned char)a;
      nb[k] = 12;
      ne[k] = (unsigned char)(78 + a);
      ++k;
    }
  }
  oa[90] = ob[90] = 13;
  oe[90] = 90;
  na[k] = nb[k] = 13;
  ne[k] = 90;
  ++k;
  if (k != NOUT_NOEXT) {
    c->err = "internal: compact output table size";
    return LIO_E_INVALID;
  }
  LIO_CHECK(c, cudaMemcpyToSymbol(c_oa_ext, oa, NOUT_EXT));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_ob_ext, ob, NOUT_EXT));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_oe_ext, oe, NOUT_EXT));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_oa_no, na, NOUT_NOEXT));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_ob_no, nb, NOUT_NOEXT));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_oe_no, ne, NOUT_NOEXT));
  {
    unsigned long long is_no[2] = {0ull, 0ull};
    for (int i = 0; i < 21; ++i) is_no[ne[i] >> 6] |= 1ull << (ne[i] & 63);
    LIO_CHECK(c, cudaMemcpyToSymbol(c_is_no, is_no, sizeof(is_no)));
  }
  // the pass / update kernels carry {pass or solve state | staging area of the searches} in dynamic shared memory
  const void* big[] = {(const void*)update_kernel,       (const void*)update_kernel_sh,   (const void*)update_kernel_host,
                       (const void*)update_kernel_multi, (const void*)pass_kernel<false>, (const void*)pass_kernel<true>,
                       (const void*)update_kernel_stage, (const void*)update_kernel_host_stage};
  for (const void* f : big)
    LIO_CHECK(c, cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PASS_BYTES));
  if (c->device < 64) g_tables_ready[c->device] = true;
  return LIO_OK;
}

// blocks of the persistent grid: everything co-resident (cooperative launch), 2 blocks per SM at most
int pass_grid_blocks(lio_ctx* c) {
  if (c->pass_grid > 0) return c->pass_grid;
  int per_sm = 0;
  cudaFuncSetAttribute((const void*)update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PASS_BYTES);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, update_kernel, THREADS, SMEM_PASS_BYTES) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  if (per_sm > LIO_BLOCKS_PER_SM) per_sm = LIO_BLOCKS_PER_SM;
  c->pass_grid = per_sm * c->sm_count;
  return c->pass_grid;
}

static PassArgs make_pass_args(lio_ctx* c, int ext, float own_min, float own_max, bool sharded_call = false) {
  PassArgs a;
  a.body = c->d_body;
  a.scan_m = c->d_scan_m;
  a.m_value = -1;
  a.min_m = c->min_m;
  a.body_src = nullptr;
  a.map = c->map;
  a.near_pts = c->d_near;
  a.near_d2 = c->d_near_d2;
  a.near_cnt = c->d_near_cnt;
  a.near_q = c->d_near_q;
  a.selected = c->d_selected;
  a.normvec = c->d_normvec;
  a.plane = c->d_plane;
  a.world = c->d_world;
  a.extrinsic_est = ext;
  a.max_d2 = c->caps.knn_max_d2;
  a.plane_thr = c->caps.plane_thr;
  a.rings = c->knn_rings;
  a.own_min = own_min;
  a.own_max = own_max;
  a.stripe_mod = sharded_call ? c->stripe_world : 0;
  a.stripe_rank = c->stripe_rank;
  a.stripe_origin = c->stripe_origin;
  a.stripe_inv_w = c->stripe_width > 0.f ? 1.0f / c->stripe_width : 0.f;
  a.sharded = (a.stripe_mod > 0 || !(own_min == -INFINITY && own_max == INFINITY)) ? 1 : 0;
  a.stage = c->stage_search ? 1 : 0;
  a.interleave = c->interleave ? 1 : 0;
  a.batched_finish = c->batched_finish ? 1 : 0;
  a.partials = c->d_partials;
  a.dbg = c->d_dbg;
  if (c->d_dbg) cudaMemsetAsync(c->d_dbg, 0, 256 * sizeof(long long), c->stream);
  return a;
}

static SolveArgs make_solve_args(lio_ctx* c, double R, int max_iter, int from_snapshot) {
  SolveArgs s;
  s.x = c->d_x;
  s.xprop = c->d_xprop;
  s.P = c->d_P;
  s.x0 = c->d_x0;
  s.P0 = c->d_P0;
  s.ctrl = c->d_ctrl;
  s.dx_out = c->d_dx;
  s.blob = c->d_blob;
  s.prior = c->d_prior;
  s.sync = c->d_sync;
  s.pub = c->d_pub;
  s.R = R;
  s.max_iter = max_iter;
  s.from_snapshot = from_snapshot;
  s.dbg = c->d_dbg;
  return s;
}

// Stamps are epoch + pass + 1 and only ever grow; long before they wrap (once in ~10^8 updates) every stamped word is
// cleared (on `on`'s stream: the one the next launch of k's buffers goes to) and the count starts again.
static int epoch_guard(lio_ctx* k, lio_ctx* on) {
  if (k->epoch <= 0xF0000000u) return LIO_OK;
  LIO_CHECK(on, cudaMemsetAsync(k->d_partials, 0, 8 * (size_t)ROW_WORDS * pass_grid_blocks(k), on->stream));
  LIO_CHECK(on, cudaMemsetAsync(k->d_pub, 0, 8 * PUB_COPIES * PUB_STRIDE, on->stream));
  LIO_CHECK(on, cudaMemsetAsync(k->d_sync, 0, 2 * sizeof(unsigned), on->stream));
  k->epoch = 0;
  return LIO_OK;
}

// The whole update as ONE cooperative launch.  own_min/own_max: ownership window of this rank (sharded map).
int launch_update(lio_ctx* c, double R, int max_iter, int ext, int from_snapshot, float own_min, float own_max,
                  bool sharded, const HostDirect* hd) {
  int rc = ensure_tables(c);
  if (rc) return rc;
  PassArgs a = make_pass_args(c, ext, own_min, own_max, sharded);
  if (hd != nullptr) a.interleave = 0;  // the host-direct kernel is the plain instantiation
  static thread_local HostPath hp;  // 4.9 KB: kept off the stack frame of every call
  hp.use_param_prior = 0;
  hp.host_out = nullptr;
  hp.host_flag = nullptr;
  hp.seq = 0;
  if (hd != nullptr) {
    hp.use_param_prior = 1;
    memcpy(hp.x0, hd->x0, sizeof(hp.x0));
    memcpy(hp.P0, hd->P0, sizeof(hp.P0));
    hp.host_out = hd->out_dev;
    hp.host_flag = hd->flag_dev;
    hp.seq = hd->seq;
    a.m_value = hd->m;
    a.body_src = hd->body_src;
  }
  ShardArgs sh;
  memset(&sh, 0, sizeof(sh));
  sh.world = 1;
  if (sharded) {
    if (c->peer_world < 2) {
      c->err = "sharded update without connected peers (lio_peer_connect)";
      return LIO_E_INVALID;
    }
    sh.world = c->peer_world;
    sh.rank = c->peer_rank;
    for (int r = 0; r < c->peer_world; ++r) {
      sh.mbox[r] = c->peer_mbox[r];
    }
    sh.err = c->d_peer_err;
    sh.epoch = c->peer_epoch;
    c->peer_epoch += 40;
  }
  SolveArgs s = make_solve_args(c, R, max_iter, from_snapshot);
  // arrival / release flags are epoch stamps: nothing to zero between launches (wrap-around once in ~10^8 updates)
  if (const int re = epoch_guard(c, c)) return re;
  unsigned epoch = c->epoch;
  c->epoch += 40;  // > max_iter + 2
  void* args[] = {&a, &s, &epoch, &sh, &hp};
  const bool sh_code = a.sharded || a.interleave;  // (never with the host-direct path: launch_update's callers)
  // staging (LIO_STAGE_SEARCH=1) exists for the unsharded single-sequence update only
  const bool stg = a.stage && !sh_code;
  a.stage = stg ? 1 : 0;
  const void* fn = hd != nullptr ? (stg ? (const void*)update_kernel_host_stage : (const void*)update_kernel_host)
                                 : (sh_code ? (const void*)update_kernel_sh
                                            : (stg ? (const void*)update_kernel_stage : (const void*)update_kernel));
  LIO_CHECK(c, cudaLaunchCooperativeKernel(fn, dim3(pass_grid_blocks(c)), dim3(THREADS), args, pass_smem_bytes(c), c->stream));
  c->launches++;
  return LIO_OK;
}

// n independent updates (one per context, all on the same device) in one cooperative launch on cs[0]'s stream.
int launch_update_multi(lio_ctx* const* cs, int n, double R, int max_iter, int ext, int from_snapshot) {
  lio_ctx* c = cs[0];
  int rc = ensure_tables(c);
  if (rc) return rc;
  static thread_local MultiArgs m;
  m.n = n;
  for (int q = 0; q < n; ++q) {
    lio_ctx* k = cs[q];
    m.a[q] = make_pass_args(k, ext, -INFINITY, INFINITY);
    m.a[q].interleave = 0;  // (the multi-sequence kernel is the plain instantiation)
    m.s[q] = make_solve_args(k, R, max_iter, from_snapshot);
    if (const int re = epoch_guard(k, c)) return re;
    m.epoch[q] = k->epoch;
    k->epoch += 40;
  }
  void* args[] = {&m};
  LIO_CHECK(c, cudaLaunchCooperativeKernel((const void*)update_kernel_multi, dim3(pass_grid_blocks(c)), dim3(THREADS),
                                           args, pass_smem_bytes(c), c->stream));
  c->launches++;
  return LIO_OK;
}

// mode: -1 = as the device loop state says, 0 = cached, 1 = search.  The reduced blob lands in c->d_blob.
int launch_pass(lio_ctx* c, int mode, int extrinsic_est, float own_min, float own_max, bool sharded_call) {
  int rc = ensure_tables(c);
  if (rc) return rc;
  PassArgs a = make_pass_args(c, extrinsic_est, own_min, own_max, sharded_call);
  SolveArgs s = make_solve_args(c, 0.0, 0, 0);
  LIO_CHECK(c, cudaMemsetAsync(c->d_sync, 0, 2 * sizeof(unsigned), c->stream));
  if (const int re = epoch_guard(c, c)) return re;
  if (a.sharded || a.interleave)
    pass_kernel<true><<<pass_grid_blocks(c), THREADS, pass_smem_bytes(c), c->stream>>>(a, s, mode, ++c->epoch);
  else
    pass_kernel<false><<<pass_grid_blocks(c), THREADS, pass_smem_bytes(c), c->stream>>>(a, s, mode, ++c->epoch);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int launch_solve(lio_ctx* c, double R, int extrinsic_est) {
  SolveArgs s = make_solve_args(c, R, 0, 0);
  solve_kernel<<<1, THREADS, 0, c->stream>>>(s, extrinsic_est, ++c->epoch);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int launch_begin(lio_ctx* c, int max_iter, int extrinsic_est, int from_snapshot) {
  int rc = ensure_tables(c);
  if (rc) return rc;
  SolveArgs s = make_solve_args(c, 0.0, max_iter, from_snapshot);
  LIO_CHECK(c, cudaMemsetAsync(c->d_sync, 0, 2 * sizeof(unsigned), c->stream));
  begin_kernel<<<1, THREADS, 0, c->stream>>>(s, extrinsic_est);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

// Extends the rows of the neighbour cache (c->d_near*) whose known prefix is shorter than `need` (1 ... 5) to `need`
// entries of the UNBOUNDED nearest-neighbour list of their query q[i].  m >= 0: exact row count; m < 0: the device-side
// scan size (a scan of fewer than min_m points counts as empty), `bound` rows at most.  Enqueue only.
int launch_far_complete(lio_ctx* c, const float4* d_q, int64_t m, int min_m, int64_t bound, int need) {
  if (bound <= 0) return LIO_OK;
  const int grid = (int)std::min<int64_t>(bound, (int64_t)c->sm_count * 16);
  far_search_kernel<<<grid, 32, 0, c->stream>>>(c->map, d_q, c->d_scan_m, m >= 0 ? (int)m : -1, m >= 0 ? 0 : min_m, need,
                                                 c->d_near, c->d_near_d2, c->d_near_cnt);
  c->launches += 1;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int launch_knn_batch(lio_ctx* c, const float4* d_q, int64_t m) {
  const int grid = c->sm_count * 8;
  knn_batch_kernel<<<grid, 256, 0, c->stream>>>(c->map, d_q, (int)m, c->caps.knn_max_d2, c->knn_rings, c->d_near,
                                                c->d_near_d2, c->d_near_cnt);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

}  // namespace lio
