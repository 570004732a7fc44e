// IESKF measurement update on the device.
//   pass   ≙ esekf::h_share_model (esekfom.hpp:106-227) fused with the H^T H / H^T h products of
//            update_iterated_dyn_share_modified (esekfom.hpp:306-319): body->world, 5-NN on the voxel hash,
//            validity gates, 5x3 plane fit, residual, 1x12 Jacobian row, block-level FP64 reduction.
//   solve  ≙ the rest of one loop iteration (esekfom.hpp:297-345): grid-level reduction, boxminus, Kalman step,
//            boxplus, convergence state machine, covariance update.
// update_kernel runs the WHOLE update_iterated_dyn_share_modified loop as one persistent cooperative launch: every
// pass ends in a grid barrier; the block that arrives last reduces the per-block partial sums in block order and
// performs the 24-state Kalman step while the others wait for the new state.  The loop state (lio::Ctrl) lives in
// device memory, so there is no host round trip and no launch between passes.  pass_kernel / solve_kernel run the
// same device code as single steps (host-driven passes, and the sharded-map driver that all-reduces the blob
// between them).  No floating-point atomics anywhere: every sum has a fixed order.
//
// Kalman step.  The reference forms K_front = (H^T H / R + P^-1)^-1 with two 24x24 inverses per pass
// (esekfom.hpp:311).  H has n = 6 (12 with extrinsic_est) non-zero columns, so with P11 = P[:n,:n], P21 = P[n:,:n]
//     K_front[:, :n] = [ I ; P21 P11^-1 ] (H^T H / R + P11^-1)^-1
// (block inversion; exact algebra).  P is constant during the loop, so P11^-1 and P21 P11^-1 are formed once per
// update and each pass inverts one n x n matrix in a single warp (rows in registers, Gauss-Jordan, partial pivoting).
#include <cooperative_groups.h>
#include <stdio.h>

#include "lio_ctx.cuh"
#include "lio_knn.cuh"

namespace lio {

constexpr int THREADS = 256;     // threads per block of every kernel in this file
constexpr int ROWS_MAX = 256;    // Jacobian rows staged per tile (cached passes: one thread per point)
constexpr int RS = 14;           // row stride: 12 Jacobian columns, residual, 1.0 (row counter)
constexpr int NOUT_EXT = 91;     // 78 HtH + 12 Hth + count
constexpr int NOUT_NOEXT = 28;   // 21 HtH (6x6 upper) + 6 Hth + count

// compact output o -> (row column a, row column b): the accumulated quantity is sum_rows row[a] * row[b]
__constant__ unsigned char c_oa_ext[NOUT_EXT], c_ob_ext[NOUT_EXT], c_oe_ext[NOUT_EXT];
__constant__ unsigned char c_oa_no[NOUT_NOEXT], c_ob_no[NOUT_NOEXT], c_oe_no[NOUT_NOEXT];

struct PassArgs {
  const float4* body;
  const int* scan_m;
  MapView map;
  float4* near_pts;
  float* near_d2;
  int* near_cnt;
  uint8_t* selected;
  float4* normvec;
  float4* world;
  int extrinsic_est;
  float max_d2, plane_thr;
  int rings;
  float own_min, own_max;
  double* partials;  // [gridDim][LIO_BLOB]
};

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
  double R;
  int max_iter;
  int from_snapshot;
};

// per-pass constants shared by the block
struct PassConst {
  Quatd rot, rli;
  double pos[3], tli[3];
  double Rt[9], Rli[9];
};

__device__ __forceinline__ int pick_group(int M, int nblocks) {
  const long long lanes = (long long)nblocks * THREADS;
  if ((long long)M * 32 <= lanes) return 32;
  if ((long long)M * 16 <= lanes) return 16;
  return 8;
}
__device__ __forceinline__ int pick_rows_cached(int M, int nblocks) {
  int r = (M + nblocks - 1) / nblocks;
  r = (r + 31) & ~31;
  return r < 32 ? 32 : (r > ROWS_MAX ? ROWS_MAX : r);
}
__device__ __forceinline__ int tiles_of(int M, int rows) { return (M + rows - 1) / rows; }

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

// steps 1.4-3 of h_share_model for one point whose neighbours are known (esekfom.hpp:144-226)
__device__ __forceinline__ void finish_point(const PassArgs& a, const PassConst& pc, int i, const double pb[3], float pwx,
                                             float pwy, float pwz, const float4 nb[LIO_K], bool sel, double* row,
                                             unsigned char* valid_out) {
  float pabcd[4] = {0.f, 0.f, 0.f, 0.f};
  float pd2 = 0.f;
  if (sel) {
    sel = false;
    if (esti_plane(nb, a.plane_thr, pabcd)) {
      pd2 = ((pabcd[0] * pwx + pabcd[1] * pwy) + pabcd[2] * pwz) + pabcd[3];
      const double nrm = sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
      const float sc = (float)(1.0 - 0.9 * fabs((double)pd2) / sqrt(nrm));
      if ((double)sc > 0.9) sel = true;
    }
  }
  a.selected[i] = sel ? 1 : 0;
  if (sel) a.normvec[i] = make_float4(pabcd[0], pabcd[1], pabcd[2], pd2);
  const bool valid = sel && (pwx >= a.own_min) && (pwx < a.own_max);
  *valid_out = valid ? 1 : 0;
  if (valid) {
    double pI[3], C[3], A[3];
    quat_rotate(pc.rli, pb, pI);
    pI[0] += pc.tli[0];
    pI[1] += pc.tli[1];
    pI[2] += pc.tli[2];
    const double nv[3] = {(double)pabcd[0], (double)pabcd[1], (double)pabcd[2]};
    mat3T_vec(pc.Rt, nv, C);
    const double pIx[9] = {0.0, -pI[2], pI[1], pI[2], 0.0, -pI[0], -pI[1], pI[0], 0.0};
    mat3_vec(pIx, C, A);
    row[0] = nv[0];
    row[1] = nv[1];
    row[2] = nv[2];
    row[3] = A[0];
    row[4] = A[1];
    row[5] = A[2];
    if (a.extrinsic_est) {
      double M1[9], B[3];
      const double px[9] = {0.0, -pb[2], pb[1], pb[2], 0.0, -pb[0], -pb[1], pb[0], 0.0};
      for (int r = 0; r < 3; ++r)
        for (int c2 = 0; c2 < 3; ++c2)
          M1[3 * r + c2] =
              (px[3 * r] * pc.Rli[3 * c2] + px[3 * r + 1] * pc.Rli[3 * c2 + 1]) + px[3 * r + 2] * pc.Rli[3 * c2 + 2];
      mat3_vec(M1, C, B);
      row[6] = B[0];
      row[7] = B[1];
      row[8] = B[2];
      row[9] = C[0];
      row[10] = C[1];
      row[11] = C[2];
    } else {
      row[6] = row[7] = row[8] = row[9] = row[10] = row[11] = 0.0;
    }
    row[12] = -(double)pd2;  // esekfom.hpp:225
    row[13] = 1.0;
  }
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

// One tile of a search pass: THREADS / G queries, one per G-lane group.
template <int G>
__device__ __noinline__ void search_tile(const PassArgs& a, const PassConst& pc, int M, int tile, double* s_rows,
                                            unsigned char* s_valid) {
  constexpr int ROWS = THREADS / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const int row = threadIdx.x / G;
  const int i = tile * ROWS + row;
  if (i >= M) {  // group-uniform
    if (gl == 0) s_valid[row] = 0;
    return;
  }
  const float4 b = __ldg(a.body + i);
  const double pb[3] = {b.x, b.y, b.z};
  float pwx, pwy, pwz;
  body_to_world(pc, pb, pwx, pwy, pwz);
  unsigned long long key[LIO_K];
  uint32_t slot[LIO_K];
  const int cnt = group_knn5<G>(a.map, pwx, pwy, pwz, a.max_d2, a.rings, gmask, gl, key, slot);  // esekfom.hpp:140
  if (gl != 0) return;
  a.world[i] = make_float4(pwx, pwy, pwz, b.w);
  float4 nb[LIO_K];
  float d2[LIO_K];
#pragma unroll
  for (int r = 0; r < LIO_K; ++r) {
    if (r < cnt) {
      nb[r] = __ldg(a.map.pool + slot[r]);
      d2[r] = __uint_as_float((uint32_t)(key[r] >> 32));
    } else {
      nb[r] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      d2[r] = CUDART_INF_F;
    }
  }
#pragma unroll
  for (int r = 0; r < LIO_K; ++r) {
    a.near_pts[(size_t)i * LIO_K + r] = nb[r];
    a.near_d2[(size_t)i * LIO_K + r] = d2[r];
  }
  a.near_cnt[i] = cnt;
  // step 1.4 (esekfom.hpp:144-147)
  const bool sel = (cnt < LIO_K) ? false : (d2[LIO_K - 1] > 5.0f ? false : true);
  finish_point(a, pc, i, pb, pwx, pwy, pwz, nb, sel, s_rows + row * RS, s_valid + row);
}

// One tile of a non-search pass: `rows` points, one per thread, neighbours and the sticky mask from the last search.
__device__ __noinline__ void cached_tile(const PassArgs& a, const PassConst& pc, int M, int tile, int rows,
                                            double* s_rows, unsigned char* s_valid) {
  const int row = threadIdx.x;
  if (row >= rows) return;
  const int i = tile * rows + row;
  if (i >= M) {
    s_valid[row] = 0;
    return;
  }
  const float4 b = __ldg(a.body + i);
  const double pb[3] = {b.x, b.y, b.z};
  float pwx, pwy, pwz;
  body_to_world(pc, pb, pwx, pwy, pwz);
  a.world[i] = make_float4(pwx, pwy, pwz, b.w);
  float4 nb[LIO_K];
#pragma unroll
  for (int r = 0; r < LIO_K; ++r) nb[r] = __ldcg(a.near_pts + (size_t)i * LIO_K + r);
  const bool sel = __ldcg(a.selected + i) != 0;  // sticky between search passes (esekfom.hpp:150)
  finish_point(a, pc, i, pb, pwx, pwy, pwz, nb, sel, s_rows + row * RS, s_valid + row);
}

// One h_share_model pass of this block: its tiles (tile = blockIdx.x, += gridDim.x), Jacobian rows staged in shared
// memory, products accumulated by thread (output o, segment seg) over rows seg, seg + nseg, ... of every tile, then
// the segments are combined in order and the block's partial blob is written to a.partials[blockIdx.x].
__device__ void block_pass(const PassArgs& a, const StateD* x, bool search, double* s_rows, unsigned char* s_valid,
                           PassConst* s_pc, double* s_acc) {
  const int tid = threadIdx.x;
  const int M = *a.scan_m;
  if (tid == 0) load_pass_const(x, *s_pc);
  __syncthreads();
  const PassConst& pc = *s_pc;
  const int G = pick_group(M, gridDim.x);
  const int rows = search ? THREADS / G : pick_rows_cached(M, gridDim.x);
  const int ntiles = tiles_of(M, rows);
  const int nout = a.extrinsic_est ? NOUT_EXT : NOUT_NOEXT;
  const int nseg = THREADS / nout;
  const int o = tid % nout, seg = tid / nout;
  int ca = 0, cb = 0;
  if (seg < nseg) {
    ca = a.extrinsic_est ? c_oa_ext[o] : c_oa_no[o];
    cb = a.extrinsic_est ? c_ob_ext[o] : c_ob_no[o];
  }
  double acc = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (search) {
      if (G == 32)
        search_tile<32>(a, pc, M, tile, s_rows, s_valid);
      else if (G == 16)
        search_tile<16>(a, pc, M, tile, s_rows, s_valid);
      else
        search_tile<8>(a, pc, M, tile, s_rows, s_valid);
    } else {
      cached_tile(a, pc, M, tile, rows, s_rows, s_valid);
    }
    __syncthreads();
    if (seg < nseg) {
      for (int r = seg; r < rows; r += nseg)
        if (s_valid[r]) acc = fma(s_rows[r * RS + ca], s_rows[r * RS + cb], acc);
    }
    __syncthreads();
  }
  if (blockIdx.x < ntiles) {
    if (seg < nseg) s_acc[seg * nout + o] = acc;
    __syncthreads();
    double* out = a.partials + (size_t)blockIdx.x * LIO_BLOB;
    if (tid < nout) {
      double s = 0.0;
      for (int g = 0; g < nseg; ++g) s += s_acc[g * nout + tid];
      out[a.extrinsic_est ? c_oe_ext[tid] : c_oe_no[tid]] = s;
    }
  }
}

// Sum of the per-block partials in block order into s_blob[LIO_BLOB] (all threads of the block take part).
__device__ __noinline__ void block_reduce_partials(const PassArgs& a, bool search, double* s_blob, double* s_acc) {
  const int tid = threadIdx.x;
  const int M = *a.scan_m;
  const int G = pick_group(M, gridDim.x);
  const int rows = search ? THREADS / G : pick_rows_cached(M, gridDim.x);
  const int ntiles = tiles_of(M, rows);
  const int nb = ntiles < (int)gridDim.x ? ntiles : (int)gridDim.x;
  const int nout = a.extrinsic_est ? NOUT_EXT : NOUT_NOEXT;
  const int nsub = THREADS / nout;
  const int o = tid % nout, sub = tid / nout;
  if (tid < LIO_BLOB) s_blob[tid] = 0.0;
  if (sub < nsub) {
    const int e = a.extrinsic_est ? c_oe_ext[o] : c_oe_no[o];
    double acc = 0.0;
    for (int b = sub; b < nb; b += nsub) acc += __ldcg(a.partials + (size_t)b * LIO_BLOB + e);
    s_acc[sub * nout + o] = acc;
  }
  __syncthreads();
  if (tid < nout) {
    double s = 0.0;
    for (int g = 0; g < nsub; ++g) s += s_acc[g * nout + tid];
    s_blob[a.extrinsic_est ? c_oe_ext[tid] : c_oe_no[tid]] = s;
  }
  if (tid == 0) s_blob[91] = search ? 1.0 : 0.0;
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// N x N FP64 inverse in one warp: lane r holds row r of [A | I] in registers; Gauss-Jordan with partial pivoting
// (first maximum of |a_rk|, r >= k), elimination of all other rows per step, one division per row at the end.
// A, Ainv: shared memory, row-major N x N.  Called by all 32 lanes of one warp.
// ---------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void warp_inverse(const double* A, double* Ainv) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  double a[N], b[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    a[j] = lane < N ? A[lane * N + j] : 0.0;
    b[j] = (lane == j) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const bool cand0 = lane >= k && lane < N;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(a[k]));
    const uint32_t hi = cand0 ? (uint32_t)(bits >> 32) : 0u, lo32 = (uint32_t)bits;
    const uint32_t mhi = __reduce_max_sync(FULL, hi);
    const bool cand = cand0 && hi == mhi;
    const uint32_t mlo = __reduce_max_sync(FULL, cand ? lo32 : 0u);
    const unsigned wm = __ballot_sync(FULL, cand && lo32 == mlo);
    const int piv = wm ? __ffs(wm) - 1 : k;
    if (piv != k) {  // warp-uniform
      const int src = lane == k ? piv : (lane == piv ? k : lane);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        a[j] = __shfl_sync(FULL, a[j], src);
        b[j] = __shfl_sync(FULL, b[j], src);
      }
    }
    const double pivot = __shfl_sync(FULL, a[k], k);
    const double f = a[k] / pivot;
    const bool upd = lane != k && lane < N;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j > k) {
        const double pk = __shfl_sync(FULL, a[j], k);
        if (upd) a[j] = fma(-f, pk, a[j]);
      }
      const double pbj = __shfl_sync(FULL, b[j], k);
      if (upd) b[j] = fma(-f, pbj, b[j]);
    }
    if (upd) a[k] = 0.0;
  }
  double diag = 1.0;
#pragma unroll
  for (int j = 0; j < N; ++j)
    if (lane == j) diag = a[j];
  if (lane < N) {
#pragma unroll
    for (int j = 0; j < N; ++j) Ainv[lane * N + j] = b[j] / diag;
  }
  __syncwarp();
}

// The same algorithm for a run-time n <= 12 with [A | I] in shared memory (W: n rows of WS doubles); used for
// n = 12, where the unrolled register version would not fit the 128-register budget of the persistent kernel.
constexpr int WS = 25;  // odd row stride: lanes (= rows) hit distinct banks
__device__ __noinline__ void warp_inverse_smem(const double* A, double* Ainv, int n, double* W) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int w = 2 * n;
  for (int idx = lane; idx < n * w; idx += 32) {
    const int r = idx / w, c = idx % w;
    W[r * WS + c] = c < n ? A[r * n + c] : (c - n == r ? 1.0 : 0.0);
  }
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    const bool cand0 = lane >= k && lane < n;
    const unsigned long long bits = cand0 ? (unsigned long long)__double_as_longlong(fabs(W[lane * WS + k])) : 0ull;
    const uint32_t hi = (uint32_t)(bits >> 32), lo32 = (uint32_t)bits;
    const uint32_t mhi = __reduce_max_sync(FULL, hi);
    const bool cand = cand0 && hi == mhi;
    const uint32_t mlo = __reduce_max_sync(FULL, cand ? lo32 : 0u);
    const unsigned wm = __ballot_sync(FULL, cand && lo32 == mlo);
    const int piv = wm ? __ffs(wm) - 1 : k;
    if (piv != k) {
      if (lane < w) {
        const double t = W[k * WS + lane];
        W[k * WS + lane] = W[piv * WS + lane];
        W[piv * WS + lane] = t;
      }
      __syncwarp();
    }
    if (lane < n && lane != k) {
      const double f = W[lane * WS + k] / W[k * WS + k];
      for (int c = k + 1; c < w; ++c) W[lane * WS + c] = fma(-f, W[k * WS + c], W[lane * WS + c]);
      W[lane * WS + k] = 0.0;
    }
    __syncwarp();
  }
  if (lane < n) {
    const double d = W[lane * WS + lane];
    for (int c = 0; c < n; ++c) Ainv[lane * n + c] = W[lane * WS + n + c] / d;
  }
  __syncwarp();
}

// shared scratch of the solve step (doubles)
struct SolveSmem {
  double blob[LIO_BLOB];
  double S[144], Sinv[144], Kf[288], KH[288], Kh[24], dxn[24], dx[24];
  double W[12 * WS];  // [A | I] of warp_inverse_smem
  int fin, skip;
};

// Once per update: restore the prior (from_snapshot), x_propagated = x, loop state, P11^-1 and P21 P11^-1.
__device__ __noinline__ void block_prior(const SolveArgs& s, int n, SolveSmem* sm) {
  const int tid = threadIdx.x;
  if (s.from_snapshot) {
    if (tid < 26) reinterpret_cast<double*>(s.x)[tid] = reinterpret_cast<const double*>(s.x0)[tid];
    for (int k = tid; k < 576; k += THREADS) s.P[k] = s.P0[k];
  }
  const double* Psrc = s.from_snapshot ? s.P0 : s.P;
  const double* xsrc = reinterpret_cast<const double*>(s.from_snapshot ? s.x0 : s.x);
  if (tid < 26) reinterpret_cast<double*>(s.xprop)[tid] = xsrc[tid];  // esekfom.hpp:287
  if (tid == 0) {
    s.ctrl->iter = -1;
    s.ctrl->converge = 1;
    s.ctrl->t = 0;
    s.ctrl->done = 0;
    s.ctrl->n_passes = 0;
    s.ctrl->n_valid_last = 0;
    s.ctrl->max_iter = s.max_iter;
  }
  for (int k = tid; k < n * n; k += THREADS) sm->S[k] = Psrc[(k / n) * 24 + (k % n)];
  __syncthreads();
  if (tid < 32) {
    if (n == 6)
      warp_inverse<6>(sm->S, sm->Sinv);
    else
      warp_inverse_smem(sm->S, sm->Sinv, n, sm->W);
  }
  __syncthreads();
  for (int k = tid; k < n * n; k += THREADS) s.prior[k] = sm->Sinv[k];
  for (int k = tid; k < (24 - n) * n; k += THREADS) {
    const int r = k / n, c = k % n;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc = fma(Psrc[(n + r) * 24 + j], sm->Sinv[j * n + c], acc);
    s.prior[144 + k] = acc;
  }
  __syncthreads();
}

// One Kalman step from the reduced blob in sm->blob (esekfom.hpp:297-345).
__device__ __noinline__ void block_solve(const SolveArgs& s, int n, SolveSmem* sm) {
  const int tid = threadIdx.x;
  Ctrl* ctrl = s.ctrl;
  const int iter = __ldcg(&ctrl->iter), max_iter = __ldcg(&ctrl->max_iter);
  if (tid < LIO_BLOB) s.blob[tid] = sm->blob[tid];
  const int n_valid = (int)sm->blob[90];
  if (n_valid < 1) {
    // `if (!dyn_share.valid) continue;` (esekfom.hpp:297-299): nothing changes, the loop counter advances
    if (tid == 0) {
      ctrl->n_valid_last = 0;
      ctrl->n_passes = __ldcg(&ctrl->n_passes) + 1;
      ctrl->iter = iter + 1;
      if (iter + 1 >= max_iter) ctrl->done = 1;
    }
    __syncthreads();
    return;
  }
  // S = HtH[:n,:n] / R + P11^-1 ; dx_new = x [-] x_propagated (esekfom.hpp:303), in another warp
  for (int k = tid; k < n * n; k += THREADS) {
    const int r = k / n, c = k % n;
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    sm->S[k] = sm->blob[lo * 12 - (lo * (lo - 1)) / 2 + (hi - lo)] / s.R + __ldcg(s.prior + k);
  }
  if (tid == 32) {
    StateD xa, xb;
#pragma unroll
    for (int k = 0; k < 26; ++k) {
      reinterpret_cast<double*>(&xa)[k] = __ldcg(reinterpret_cast<const double*>(s.x) + k);
      reinterpret_cast<double*>(&xb)[k] = __ldcg(reinterpret_cast<const double*>(s.xprop) + k);
    }
    boxminus(xa, xb, sm->dxn);
  }
  __syncthreads();
  if (tid < 32) {
    if (n == 6)
      warp_inverse<6>(sm->S, sm->Sinv);
    else
      warp_inverse_smem(sm->S, sm->Sinv, n, sm->W);
  }
  __syncthreads();
  // K_front[:, :n] = [Sinv ; (P21 P11^-1) Sinv]
  for (int k = tid; k < 24 * n; k += THREADS) {
    const int r = k / n, c = k % n;
    double acc;
    if (r < n) {
      acc = sm->Sinv[r * n + c];
    } else {
      acc = 0.0;
      for (int j = 0; j < n; ++j) acc = fma(__ldcg(s.prior + 144 + (r - n) * n + j), sm->Sinv[j * n + c], acc);
    }
    sm->Kf[k] = acc;
  }
  __syncthreads();
  // KH[:, :n] = K_front[:, :n] HtH / R ;  K h = K_front[:, :n] Hth / R   (esekfom.hpp:314-319, regrouped)
  for (int k = tid; k < 24 * n + 24; k += THREADS) {
    if (k < 24 * n) {
      const int r = k / n, c = k % n;
      double acc = 0.0;
      for (int j = 0; j < n; ++j) {
        const int lo = j < c ? j : c, hi = j < c ? c : j;
        acc = fma(sm->Kf[r * n + j], sm->blob[lo * 12 - (lo * (lo - 1)) / 2 + (hi - lo)], acc);
      }
      sm->KH[k] = acc / s.R;
    } else {
      const int r = k - 24 * n;
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc = fma(sm->Kf[r * n + j], sm->blob[78 + j], acc);
      sm->Kh[r] = acc / s.R;
    }
  }
  __syncthreads();
  if (tid < 24) {
    // dx = K h + (K H - I) dx_new   (esekfom.hpp:319)
    double acc = 0.0;
    for (int c = 0; c < 24; ++c) {
      const double khc = (c < n) ? sm->KH[tid * n + c] : 0.0;
      acc = fma(khc - (tid == c ? 1.0 : 0.0), sm->dxn[c], acc);
    }
    sm->dx[tid] = sm->Kh[tid] + acc;
  }
  __syncthreads();
  if (tid == 0) {
    StateD xa, xn;
#pragma unroll
    for (int k = 0; k < 26; ++k)
      reinterpret_cast<double*>(&xa)[k] = __ldcg(reinterpret_cast<const double*>(s.x) + k);
    boxplus(xa, sm->dx, xn);  // esekfom.hpp:321
    *s.x = xn;
    bool converge = true;
    for (int jj = 0; jj < 24; ++jj)
      if (fabs(sm->dx[jj]) > 0.001) {
        converge = false;
        break;
      }
    int t = __ldcg(&ctrl->t);
    if (converge) t++;
    if (!t && iter == max_iter - 2) converge = true;
    const int fin = (t > 1 || iter == max_iter - 1) ? 1 : 0;
    ctrl->converge = converge ? 1 : 0;
    ctrl->t = t;
    ctrl->n_valid_last = n_valid;
    ctrl->n_passes = __ldcg(&ctrl->n_passes) + 1;
    ctrl->iter = iter + 1;
    if (fin) ctrl->done = 1;
    sm->fin = fin;
  }
  if (tid < 24 && s.dx_out) s.dx_out[tid] = sm->dx[tid];
  __syncthreads();
  if (sm->fin) {
    // P = (I - K H) P   (esekfom.hpp:342); K H has n non-zero columns.  Rows are independent: read all, then write.
    double out[3];
    int cnt = 0;
    for (int k = tid; k < 576; k += THREADS) {
      const int r = k / 24, c = k % 24;
      double acc = __ldcg(s.P + r * 24 + c);
      for (int j = 0; j < n; ++j) acc = fma(-sm->KH[r * n + j], __ldcg(s.P + j * 24 + c), acc);
      out[cnt++] = acc;
    }
    __syncthreads();
    cnt = 0;
    for (int k = tid; k < 576; k += THREADS) s.P[k] = out[cnt++];
  }
  __syncthreads();
}

struct __align__(16) PassSmem {
  double rows[ROWS_MAX * RS];
  double acc[THREADS];
  PassConst pc;
  unsigned char valid[ROWS_MAX];
  int flag;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The whole update_iterated_dyn_share_modified loop (esekfom.hpp:270-346).  Cooperative launch: all blocks resident.
__global__ void __launch_bounds__(THREADS, 2) update_kernel(const PassArgs a, const SolveArgs s) {
  __shared__ PassSmem ps;
  __shared__ SolveSmem ss;
  const int tid = threadIdx.x;
  const int n = a.extrinsic_est ? 12 : 6;
  const unsigned nblk = gridDim.x;
  // the last block prepares the prior while the others already search (pass 0 always searches at the prior state)
  if (blockIdx.x == nblk - 1) block_prior(s, n, &ss);
  const StateD* x_first = s.from_snapshot ? s.x0 : s.x;
  for (int pass_no = 0; pass_no <= s.max_iter; ++pass_no) {
    bool search = true;
    if (pass_no > 0) search = __ldcg(&s.ctrl->converge) != 0;
    block_pass(a, pass_no == 0 ? x_first : s.x, search, ps.rows, ps.valid, &ps.pc, ps.acc);
    // grid barrier; the last block to arrive reduces and solves
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned ticket = atomicAdd(&s.sync[0], 1u);
      ps.flag = (ticket == (unsigned)(pass_no + 1) * nblk - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (ps.flag) {
      __threadfence();
      block_reduce_partials(a, search, ss.blob, ps.acc);
      block_solve(s, n, &ss);
      __threadfence();
      if (tid == 0) st_release(&s.sync[1], (unsigned)(pass_no + 1));
    } else if (tid == 0) {
      while (ld_acquire(&s.sync[1]) < (unsigned)(pass_no + 1)) {
      }
    }
    __syncthreads();
    if (__ldcg(&s.ctrl->done)) break;
  }
}

// One pass at the state in s.x (or its snapshot); the last block to finish reduces the partials into s.blob.
// mode: 0 cached, 1 search, -1 as the loop state says (sharded driver).
__global__ void __launch_bounds__(THREADS, 2) pass_kernel(const PassArgs a, const SolveArgs s, int mode) {
  __shared__ PassSmem ps;
  __shared__ double s_blob[LIO_BLOB];
  const int tid = threadIdx.x;
  if (mode < 0 && s.ctrl->done) return;
  const bool search = mode < 0 ? (s.ctrl->converge != 0) : (mode != 0);
  block_pass(a, s.x, search, ps.rows, ps.valid, &ps.pc, ps.acc);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(&s.sync[0], 1u);
    ps.flag = (ticket == gridDim.x - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (ps.flag) {
    __threadfence();
    block_reduce_partials(a, search, s_blob, ps.acc);
    if (tid < LIO_BLOB) s.blob[tid] = s_blob[tid];
    if (tid == 0) s.sync[0] = 0;  // ready for the next pass launch
  }
}

// Kalman step from the blob in s.blob (already summed over ranks by the sharded driver).
__global__ void __launch_bounds__(THREADS) solve_kernel(const SolveArgs s, int extrinsic_est) {
  __shared__ SolveSmem ss;
  if (s.ctrl->done) return;
  if (threadIdx.x < LIO_BLOB) ss.blob[threadIdx.x] = s.blob[threadIdx.x];
  __syncthreads();
  block_solve(s, extrinsic_est ? 12 : 6, &ss);
}

__global__ void __launch_bounds__(THREADS) begin_kernel(const SolveArgs s, int extrinsic_est) {
  __shared__ SolveSmem ss;
  block_prior(s, extrinsic_est ? 12 : 6, &ss);
}

// Stand-alone batch of Nearest_Search calls (lio_knn5): one 8-lane group per query.
__global__ void __launch_bounds__(256) knn_batch_kernel(MapView map, const float4* q, int m, float max_d2, int rings,
                                                        float4* near_pts, float* near_d2, int* near_cnt) {
  constexpr int G = 8;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const int gglobal = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int ngroups = (gridDim.x * blockDim.x) / G;
  for (int i = gglobal; i < m; i += ngroups) {
    const float4 p = __ldg(q + i);
    unsigned long long ok[LIO_K];
    uint32_t os[LIO_K];
    const int f = group_knn5<G>(map, p.x, p.y, p.z, max_d2, rings, gmask, gl, ok, os);
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
  if (c->device < 64) g_tables_ready[c->device] = true;
  return LIO_OK;
}

// blocks of the persistent grid: everything co-resident (cooperative launch), 2 blocks per SM at most
int pass_grid_blocks(lio_ctx* c) {
  if (c->pass_grid > 0) return c->pass_grid;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, update_kernel, THREADS, 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  c->pass_grid = per_sm * c->sm_count;
  return c->pass_grid;
}

static PassArgs make_pass_args(lio_ctx* c, int ext, float own_min, float own_max) {
  PassArgs a;
  a.body = c->d_body;
  a.scan_m = c->d_scan_m;
  a.map = c->map;
  a.near_pts = c->d_near;
  a.near_d2 = c->d_near_d2;
  a.near_cnt = c->d_near_cnt;
  a.selected = c->d_selected;
  a.normvec = c->d_normvec;
  a.world = c->d_world;
  a.extrinsic_est = ext;
  a.max_d2 = c->caps.knn_max_d2;
  a.plane_thr = c->caps.plane_thr;
  a.rings = c->knn_rings;
  a.own_min = own_min;
  a.own_max = own_max;
  a.partials = c->d_partials;
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
  s.R = R;
  s.max_iter = max_iter;
  s.from_snapshot = from_snapshot;
  return s;
}

// The whole update as ONE cooperative launch (plus the 8-byte reset of the barrier words).
int launch_update(lio_ctx* c, double R, int max_iter, int ext, int from_snapshot) {
  int rc = ensure_tables(c);
  if (rc) return rc;
  PassArgs a = make_pass_args(c, ext, -INFINITY, INFINITY);
  SolveArgs s = make_solve_args(c, R, max_iter, from_snapshot);
  LIO_CHECK(c, cudaMemsetAsync(c->d_sync, 0, 2 * sizeof(unsigned), c->stream));
  void* args[] = {&a, &s};
  LIO_CHECK(c, cudaLaunchCooperativeKernel((const void*)update_kernel, dim3(pass_grid_blocks(c)), dim3(THREADS), args,
                                           0, c->stream));
  c->launches++;
  return LIO_OK;
}

// mode: -1 = as the device loop state says, 0 = cached, 1 = search.  The reduced blob lands in c->d_blob.
int launch_pass(lio_ctx* c, int mode, int extrinsic_est, float own_min, float own_max) {
  int rc = ensure_tables(c);
  if (rc) return rc;
  PassArgs a = make_pass_args(c, extrinsic_est, own_min, own_max);
  SolveArgs s = make_solve_args(c, 0.0, 0, 0);
  LIO_CHECK(c, cudaMemsetAsync(c->d_sync, 0, 2 * sizeof(unsigned), c->stream));
  pass_kernel<<<pass_grid_blocks(c), THREADS, 0, c->stream>>>(a, s, mode);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int launch_solve(lio_ctx* c, double R, int extrinsic_est) {
  SolveArgs s = make_solve_args(c, R, 0, 0);
  solve_kernel<<<1, THREADS, 0, c->stream>>>(s, extrinsic_est);
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

int launch_knn_batch(lio_ctx* c, const float4* d_q, int64_t m) {
  const int grid = c->sm_count * 8;
  knn_batch_kernel<<<grid, 256, 0, c->stream>>>(c->map, d_q, (int)m, c->caps.knn_max_d2, c->knn_rings, c->d_near,
                                                c->d_near_d2, c->d_near_cnt);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

}  // namespace lio
