// IESKF measurement update on the device.
//   pass_kernel  ≙ esekf::h_share_model (esekfom.hpp:106-227) fused with the H^T H / H^T h products of
//                  update_iterated_dyn_share_modified (esekfom.hpp:306-319): body->world, 5-NN on the voxel hash,
//                  validity gates, 5x3 plane fit, residual, 1x12 Jacobian row, block-level FP64 reduction.
//   solve_kernel ≙ the rest of one loop iteration (esekfom.hpp:297-345): grid-level reduction, boxminus, two
//                  24x24 inverses, Kalman step, boxplus, convergence state machine, covariance update.
// Both read the loop state (lio::Ctrl) from device memory, so the whole update is one CUDA graph with no host
// round trip.  No atomics anywhere on the reduction path: every sum has a fixed order.
#include <stdio.h>

#include "lio_ctx.cuh"
#include "lio_knn.cuh"

namespace lio {

// (a,b) of the e-th element of the row-major upper triangle of a 12x12 matrix
__constant__ unsigned char c_pair_a[78];
__constant__ unsigned char c_pair_b[78];

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
  const StateD* x;
  const Ctrl* ctrl;  // nullptr: single pass driven by the host (always runs)
  int extrinsic_est;
  float max_d2, plane_thr;
  int rings;
  float own_min, own_max;
  double* partials;
};

constexpr int ROW_STRIDE = 13;  // 12 Jacobian columns + residual

template <int QPW, int WARPS, bool SEARCH>
__global__ void __launch_bounds__(WARPS * 32) pass_kernel(const PassArgs a) {
  constexpr int ROWS = QPW * WARPS;
  __shared__ double s_rows[ROWS][ROW_STRIDE];
  __shared__ unsigned char s_valid[ROWS];
  if (a.ctrl != nullptr) {
    if (a.ctrl->done) return;
    if ((a.ctrl->converge != 0) != SEARCH) return;
  }
  const unsigned FULL = 0xffffffffu;
  const int M = *a.scan_m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (M + ROWS - 1) / ROWS;

  // state (uniform; broadcast loads)
  const Quatd rot = a.x->rot, rli = a.x->rli;
  const double pos[3] = {a.x->pos[0], a.x->pos[1], a.x->pos[2]};
  const double tli[3] = {a.x->tli[0], a.x->tli[1], a.x->tli[2]};

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int qbase = tile * ROWS + warp * QPW;
    const int i = qbase + lane;
    const bool owner = (lane < QPW) && (i < M);

    // step 1.1-1.2 (esekfom.hpp:123-133): p_world = rot * (R_LI * p + t_LI) + pos, FP64 -> FP32
    double pb[3] = {0, 0, 0};
    float pwx = 0.f, pwy = 0.f, pwz = 0.f;
    if (owner) {
      const float4 b = __ldg(a.body + i);
      pb[0] = b.x;
      pb[1] = b.y;
      pb[2] = b.z;
      double pi[3], pg[3];
      quat_rotate(rli, pb, pi);
      pi[0] += tli[0];
      pi[1] += tli[1];
      pi[2] += tli[2];
      quat_rotate(rot, pi, pg);
      pwx = (float)(pg[0] + pos[0]);
      pwy = (float)(pg[1] + pos[1]);
      pwz = (float)(pg[2] + pos[2]);
      if (a.world) a.world[i] = make_float4(pwx, pwy, pwz, b.w);
    }

    float4 nb[LIO_K];
    int cnt = 0;
    bool sel = false;
    if (SEARCH) {
      // step 1.3 (esekfom.hpp:140): the warp searches for its QPW queries one after another
      unsigned long long mykey[LIO_K];
      uint32_t myslot[LIO_K];
#pragma unroll
      for (int r = 0; r < LIO_K; ++r) {
        mykey[r] = ~0ull;
        myslot[r] = 0;
      }
      for (int qi = 0; qi < QPW; ++qi) {
        if (qbase + qi >= M) break;
        const float qx = __shfl_sync(FULL, pwx, qi), qy = __shfl_sync(FULL, pwy, qi), qz = __shfl_sync(FULL, pwz, qi);
        unsigned long long ok[LIO_K];
        uint32_t os[LIO_K];
        const int f = warp_knn5(a.map, qx, qy, qz, a.max_d2, a.rings, ok, os);
        if (lane == qi) {
#pragma unroll
          for (int r = 0; r < LIO_K; ++r) {
            mykey[r] = ok[r];
            myslot[r] = os[r];
          }
          cnt = f;
        }
      }
      if (owner) {
        float d2[LIO_K];
#pragma unroll
        for (int r = 0; r < LIO_K; ++r) {
          if (r < cnt) {
            nb[r] = __ldg(a.map.pool + myslot[r]);
            d2[r] = __uint_as_float((uint32_t)(mykey[r] >> 32));
          } else {
            nb[r] = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            d2[r] = CUDART_INF_F;
          }
          a.near_pts[(size_t)i * LIO_K + r] = nb[r];
          a.near_d2[(size_t)i * LIO_K + r] = d2[r];
        }
        a.near_cnt[i] = cnt;
        // step 1.4 (esekfom.hpp:144-147)
        sel = (cnt < LIO_K) ? false : (d2[LIO_K - 1] > 5.0f ? false : true);
      }
    } else if (owner) {
#pragma unroll
      for (int r = 0; r < LIO_K; ++r) nb[r] = __ldg(a.near_pts + (size_t)i * LIO_K + r);
      sel = a.selected[i] != 0;  // sticky between search passes (esekfom.hpp:150)
    }

    // step 1.5 (esekfom.hpp:153-173)
    float pabcd[4] = {0.f, 0.f, 0.f, 0.f};
    float pd2 = 0.f;
    if (owner && sel) {
      sel = false;
      if (esti_plane(nb, a.plane_thr, pabcd)) {
        pd2 = ((pabcd[0] * pwx + pabcd[1] * pwy) + pabcd[2] * pwz) + pabcd[3];
        const double nrm = sqrt((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
        const float sc = (float)(1.0 - 0.9 * fabs((double)pd2) / sqrt(nrm));
        if ((double)sc > 0.9) sel = true;
      }
    }
    if (owner) {
      a.selected[i] = sel ? 1 : 0;
      if (a.normvec && sel) a.normvec[i] = make_float4(pabcd[0], pabcd[1], pabcd[2], pd2);
    }

    // step 3 (esekfom.hpp:194-226): Jacobian row for valid points this rank owns
    const bool valid = owner && sel && (pwx >= a.own_min) && (pwx < a.own_max);
    const int rix = warp * QPW + lane;
    if (lane < QPW) {
      s_valid[rix] = valid ? 1 : 0;
      if (valid) {
        double Rt[9], pI[3], C[3], A[3];
        quat_to_mat(rot, Rt);
        quat_rotate(rli, pb, pI);
        pI[0] += tli[0];
        pI[1] += tli[1];
        pI[2] += tli[2];
        const double nv[3] = {(double)pabcd[0], (double)pabcd[1], (double)pabcd[2]};
        mat3T_vec(Rt, nv, C);
        const double pIx[9] = {0.0, -pI[2], pI[1], pI[2], 0.0, -pI[0], -pI[1], pI[0], 0.0};
        mat3_vec(pIx, C, A);
        double* row = s_rows[rix];
        row[0] = nv[0];
        row[1] = nv[1];
        row[2] = nv[2];
        row[3] = A[0];
        row[4] = A[1];
        row[5] = A[2];
        if (a.extrinsic_est) {
          double Rli[9], M1[9], B[3];
          quat_to_mat(rli, Rli);
          const double px[9] = {0.0, -pb[2], pb[1], pb[2], 0.0, -pb[0], -pb[1], pb[0], 0.0};
          for (int r = 0; r < 3; ++r)
            for (int c2 = 0; c2 < 3; ++c2)
              M1[3 * r + c2] = (px[3 * r] * Rli[3 * c2] + px[3 * r + 1] * Rli[3 * c2 + 1]) + px[3 * r + 2] * Rli[3 * c2 + 2];
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
      }
    }
    __syncthreads();

    // block-level segmented reduction: thread e owns output element e and walks the tile's rows in index order
    for (int e = threadIdx.x; e < LIO_BLOB; e += WARPS * 32) {
      double acc = 0.0;
      if (e < 78) {
        const int ca = c_pair_a[e], cb = c_pair_b[e];
        if (a.extrinsic_est || cb < 6) {
          for (int r = 0; r < ROWS; ++r)
            if (s_valid[r]) acc = fma(s_rows[r][ca], s_rows[r][cb], acc);
        }
      } else if (e < 90) {
        const int ca = e - 78;
        if (a.extrinsic_est || ca < 6) {
          for (int r = 0; r < ROWS; ++r)
            if (s_valid[r]) acc = fma(s_rows[r][ca], s_rows[r][12], acc);
        }
      } else if (e == 90) {
        int n = 0;
        for (int r = 0; r < ROWS; ++r) n += s_valid[r];
        acc = (double)n;
      } else {
        acc = (tile == 0 && SEARCH) ? 1.0 : 0.0;
      }
      a.partials[(size_t)tile * LIO_BLOB + e] = acc;
    }
    __syncthreads();
  }
}

// Grid-level reduction of the per-tile partials in a fixed order: 8 lanes per output element take the tiles
// sub, sub+8, ... in ascending order, then a fixed shuffle tree combines the 8 partial sums.
__device__ __forceinline__ void reduce_partials(const double* partials, int ntiles, double* out /*LIO_BLOB*/) {
  const int e = threadIdx.x >> 3, sub = threadIdx.x & 7;
  double acc = 0.0;
  if (e < LIO_BLOB)
    for (int t = sub; t < ntiles; t += 8) acc += partials[(size_t)t * LIO_BLOB + e];
  acc += __shfl_down_sync(0xffffffffu, acc, 4, 8);
  acc += __shfl_down_sync(0xffffffffu, acc, 2, 8);
  acc += __shfl_down_sync(0xffffffffu, acc, 1, 8);
  if (e < LIO_BLOB && sub == 0) out[e] = acc;
}

__global__ void __launch_bounds__(768) reduce_blob_kernel(const double* partials, const int* scan_m, int rows,
                                                          const Ctrl* ctrl, int rows_search, int rows_cached,
                                                          double* blob) {
  if (ctrl != nullptr) {
    if (ctrl->done) return;
    rows = ctrl->converge ? rows_search : rows_cached;
  }
  const int ntiles = (*scan_m + rows - 1) / rows;
  reduce_partials(partials, ntiles, blob);
}

// ---------------------------------------------------------------------------------------------------------
// 24x24 FP64 inverse by partial-pivot LU + triangular solves against the identity (what Eigen's inverse() does
// for n > 4), cooperative over a 768-thread block, in the per-element operation order of the
// oracle (elimination a_ij -= l_i * u_j for k ascending; forward sums j ascending; backward sums j descending).
// A, Ainv, lu: shared 24x24 row-major.  Thread layout: LU phase thread (i,j) = (tid / 24, tid % 24) for tid < 576;
// solve phase warp c = column c of the identity, lane i = row i.
// ---------------------------------------------------------------------------------------------------------
__device__ void block_inverse24(const double* A, double* Ainv, double* lu, int* perm, int* s_piv) {
  const int tid = threadIdx.x;
  const int i = tid / 24, j = tid % 24;
  if (tid < 576) lu[tid] = A[tid];
  if (tid < 24) perm[tid] = tid;
  __syncthreads();
  for (int k = 0; k < 24; ++k) {
    if (tid < 32) {
      // first maximum of |lu[r][k]|, r = k..23
      double v = (tid >= k && tid < 24) ? fabs(lu[tid * 24 + k]) : -1.0;
      int r = tid;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, off);
        const int orow = __shfl_down_sync(0xffffffffu, r, off);
        if (ov > v || (ov == v && orow < r)) {
          v = ov;
          r = orow;
        }
      }
      if (tid == 0) *s_piv = r;
    }
    __syncthreads();
    const int piv = *s_piv;
    if (piv != k) {
      if (tid < 24) {
        const double t0 = lu[k * 24 + tid];
        lu[k * 24 + tid] = lu[piv * 24 + tid];
        lu[piv * 24 + tid] = t0;
      } else if (tid == 24) {
        const int t0 = perm[k];
        perm[k] = perm[piv];
        perm[piv] = t0;
      }
    }
    __syncthreads();
    if (tid < 576 && j == k && i > k) lu[i * 24 + k] = lu[i * 24 + k] / lu[k * 24 + k];
    __syncthreads();
    if (tid < 576 && i > k && j > k) lu[i * 24 + j] = lu[i * 24 + j] - lu[i * 24 + k] * lu[k * 24 + j];
    __syncthreads();
  }
  // solves: warp c handles column c of P*I; lane r holds element r
  {
    const int c = tid >> 5, r = tid & 31;
    if (c < 24) {
      double s = (r < 24 && perm[r] == c) ? 1.0 : 0.0;
      // forward (unit lower): s_r -= lu[r][jj] * y_jj, jj ascending
      for (int jj = 0; jj < 24; ++jj) {
        const double yj = __shfl_sync(0xffffffffu, s, jj);
        if (r > jj && r < 24) s = s - lu[r * 24 + jj] * yj;
      }
      // backward (upper): x_r = (s_r - sum_{jj>r} lu[r][jj] x_jj) / lu[r][r], jj descending
      for (int jj = 23; jj >= 0; --jj) {
        if (r == jj) s = s / lu[r * 24 + r];
        const double xj = __shfl_sync(0xffffffffu, s, jj);
        if (r < jj) s = s - lu[r * 24 + jj] * xj;
      }
      if (r < 24) Ainv[r * 24 + c] = s;
    }
  }
  __syncthreads();
}

struct SolveArgs {
  const double* partials;
  const int* scan_m;
  int rows_search, rows_cached;
  const double* blob_in;  // non-null: already reduced (and all-reduced across ranks)
  double* blob_out;       // reduced blob of this pass (for lio_get / tests)
  StateD* x;
  const StateD* xprop;
  double* P;
  Ctrl* ctrl;
  double* dx_out;
  double R;
};

__global__ void __launch_bounds__(768, 1) solve_kernel(const SolveArgs a) {
  __shared__ double sP[576], sInv[576], sA[576], sKf[576], sLU[576], sKH[24 * 12];
  __shared__ double sHTH[144], sHth[12], sblob[LIO_BLOB + 4], sdxn[24], sdx[24];
  __shared__ int sperm[24], spiv, s_final;
  const int tid = threadIdx.x;
  Ctrl* ctrl = a.ctrl;
  if (ctrl->done) return;
  const int iter = ctrl->iter, max_iter = ctrl->max_iter;

  if (a.blob_in != nullptr) {
    if (tid < LIO_BLOB) sblob[tid] = a.blob_in[tid];
  } else {
    const int rows = ctrl->converge ? a.rows_search : a.rows_cached;
    reduce_partials(a.partials, (*a.scan_m + rows - 1) / rows, sblob);
  }
  __syncthreads();
  if (a.blob_in == nullptr && a.blob_out != nullptr && tid < LIO_BLOB) a.blob_out[tid] = sblob[tid];
  const int n_valid = (int)sblob[90];
  __syncthreads();
  if (n_valid < 1) {
    // `if (!dyn_share.valid) continue;` (esekfom.hpp:297-299): nothing changes, the loop counter advances
    if (tid == 0) {
      ctrl->n_valid_last = 0;
      ctrl->n_passes += 1;
      ctrl->iter = iter + 1;
      if (iter + 1 >= max_iter) ctrl->done = 1;
    }
    return;
  }

  if (tid < 576) sP[tid] = a.P[tid];
  if (tid < 144) {
    const int r = tid / 12, c = tid % 12;
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    sHTH[tid] = sblob[lo * 12 - (lo * (lo - 1)) / 2 + (hi - lo)];
  }
  if (tid < 12) sHth[tid] = sblob[78 + tid];
  if (tid == 767) boxminus(*a.x, *a.xprop, sdxn);  // dx_new (esekfom.hpp:303)
  __syncthreads();

  block_inverse24(sP, sInv, sLU, sperm, &spiv);
  if (tid < 576) {
    const int r = tid / 24, c = tid % 24;
    const double hth = (r < 12 && c < 12) ? sHTH[r * 12 + c] : 0.0;
    sA[tid] = hth / a.R + sInv[tid];
  }
  __syncthreads();
  block_inverse24(sA, sKf, sLU, sperm, &spiv);

  // KH[:, :12] = K_front[:, :12] * HTH / R ;  K h = K_front[:, :12] * Hth / R   (esekfom.hpp:314-319, regrouped)
  if (tid < 288) {
    const int r = tid / 12, c = tid % 12;
    double acc = 0.0;
    for (int k = 0; k < 12; ++k) acc += sKf[r * 24 + k] * sHTH[k * 12 + c];
    sKH[tid] = acc / a.R;
  }
  __syncthreads();
  if (tid < 24) {
    double kh = 0.0;
    for (int k = 0; k < 12; ++k) kh += sKf[tid * 24 + k] * sHth[k];
    kh = kh / a.R;
    double acc = 0.0;
    for (int c = 0; c < 24; ++c) {
      const double khc = (c < 12) ? sKH[tid * 12 + c] : 0.0;
      acc += (khc - (tid == c ? 1.0 : 0.0)) * sdxn[c];
    }
    sdx[tid] = kh + acc;
  }
  __syncthreads();
  if (tid == 0) {
    StateD xn;
    boxplus(*a.x, sdx, xn);  // esekfom.hpp:321
    *a.x = xn;
    bool converge = true;
    for (int jj = 0; jj < 24; ++jj)
      if (fabs(sdx[jj]) > 0.001) {
        converge = false;
        break;
      }
    int t = ctrl->t;
    if (converge) t++;
    if (!t && iter == max_iter - 2) converge = true;
    const int fin = (t > 1 || iter == max_iter - 1) ? 1 : 0;
    ctrl->converge = converge ? 1 : 0;
    ctrl->t = t;
    ctrl->n_valid_last = n_valid;
    ctrl->n_passes += 1;
    ctrl->iter = iter + 1;
    if (fin) ctrl->done = 1;
    s_final = fin;
  }
  if (tid < 24 && a.dx_out) a.dx_out[tid] = sdx[tid];
  __syncthreads();
  if (s_final && tid < 576) {
    // P = (I - KH) P   (esekfom.hpp:342)
    const int r = tid / 24, c = tid % 24;
    double acc = 0.0;
    for (int k = 0; k < 24; ++k) {
      const double khk = (k < 12) ? sKH[r * 12 + k] : 0.0;
      acc += ((r == k ? 1.0 : 0.0) - khk) * sP[k * 24 + c];
    }
    a.P[tid] = acc;
  }
}

__global__ void begin_kernel(Ctrl* ctrl, StateD* x, StateD* xprop, double* P, const StateD* x0, const double* P0,
                             int max_iter, int from_snapshot) {
  const int tid = threadIdx.x;
  if (from_snapshot) {
    if (tid < 26) reinterpret_cast<double*>(x)[tid] = reinterpret_cast<const double*>(x0)[tid];
    for (int k = tid; k < 576; k += blockDim.x) P[k] = P0[k];
  }
  __syncthreads();
  if (tid < 26) reinterpret_cast<double*>(xprop)[tid] = reinterpret_cast<const double*>(x)[tid];
  if (tid == 0) {
    ctrl->iter = -1;
    ctrl->converge = 1;
    ctrl->t = 0;
    ctrl->done = 0;
    ctrl->n_passes = 0;
    ctrl->n_valid_last = 0;
    ctrl->max_iter = max_iter;
  }
}

// Stand-alone batch of Nearest_Search calls (lio_knn5): one warp per query.
__global__ void __launch_bounds__(256) knn_batch_kernel(MapView map, const float4* q, int m, float max_d2, int rings,
                                                        float4* near_pts, float* near_d2, int* near_cnt) {
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = wglobal; i < m; i += nwarps) {
    const float4 p = __ldg(q + i);
    unsigned long long ok[LIO_K];
    uint32_t os[LIO_K];
    const int f = warp_knn5(map, p.x, p.y, p.z, max_d2, rings, ok, os);
    if (lane < LIO_K) {
      unsigned long long k = ok[0];
      uint32_t s = os[0];
#pragma unroll
      for (int r = 1; r < LIO_K; ++r)
        if (lane == r) {
          k = ok[r];
          s = os[r];
        }
      float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      float d = CUDART_INF_F;
      if (lane < f) {
        v = __ldg(map.pool + s);
        d = __uint_as_float((uint32_t)(k >> 32));
      }
      near_pts[(size_t)i * LIO_K + lane] = v;
      near_d2[(size_t)i * LIO_K + lane] = d;
    }
    if (lane == 0) near_cnt[i] = f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
constexpr int QPW_CACHED = 32, WARPS_CACHED = 4;
constexpr int WARPS_SEARCH = 4;

static bool g_pairs_ready[64] = {false};
static int ensure_pairs(lio_ctx* c) {
  if (c->device < 64 && g_pairs_ready[c->device]) return LIO_OK;
  unsigned char pa[78], pb[78];
  int e = 0;
  for (int a = 0; a < 12; ++a)
    for (int b = a; b < 12; ++b) {
      pa[e] = (unsigned char)a;
      pb[e] = (unsigned char)b;
      ++e;
    }
  LIO_CHECK(c, cudaMemcpyToSymbol(c_pair_a, pa, 78));
  LIO_CHECK(c, cudaMemcpyToSymbol(c_pair_b, pb, 78));
  if (c->device < 64) g_pairs_ready[c->device] = true;
  return LIO_OK;
}

int ensure_tables(lio_ctx* c) { return ensure_pairs(c); }

static PassArgs make_args(lio_ctx* c, bool with_ctrl, int ext, float own_min, float own_max) {
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
  a.x = c->d_x;
  a.ctrl = with_ctrl ? c->d_ctrl : nullptr;
  a.extrinsic_est = ext;
  a.max_d2 = c->caps.knn_max_d2;
  a.plane_thr = c->caps.plane_thr;
  a.rings = c->knn_rings;
  a.own_min = own_min;
  a.own_max = own_max;
  a.partials = c->d_partials;
  return a;
}

static int rows_search(const lio_ctx* c) { return c->qpw_search * WARPS_SEARCH; }
static int rows_cached() { return QPW_CACHED * WARPS_CACHED; }

template <int QPW>
static void launch_search_variant(lio_ctx* c, const PassArgs& a, int grid) {
  pass_kernel<QPW, WARPS_SEARCH, true><<<grid, WARPS_SEARCH * 32, 0, c->stream>>>(a);
}

// force_search: -1 = device-controlled (both variants are enqueued; the one whose turn it is not exits at once),
// 0 / 1 = host-driven single pass.
int launch_pass(lio_ctx* c, int force_search, int extrinsic_est, float own_min, float own_max) {
  int rc = ensure_pairs(c);
  if (rc) return rc;
  const bool with_ctrl = force_search < 0;
  PassArgs a = make_args(c, with_ctrl, extrinsic_est, own_min, own_max);
  const int grid = c->sm_count * 8;
  if (with_ctrl || force_search == 1) {
    switch (c->qpw_search) {
      case 4: launch_search_variant<4>(c, a, grid); break;
      case 16: launch_search_variant<16>(c, a, grid); break;
      case 32: launch_search_variant<32>(c, a, grid); break;
      default: launch_search_variant<8>(c, a, grid); break;
    }
    c->launches++;
  }
  if (with_ctrl || force_search == 0) {
    pass_kernel<QPW_CACHED, WARPS_CACHED, false><<<grid, WARPS_CACHED * 32, 0, c->stream>>>(a);
    c->launches++;
  }
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

// Reduce the partials of the pass just enqueued into c->d_blob.  host_search: -1 device-controlled.
int launch_reduce_blob_mode(lio_ctx* c, int host_search) {
  const int rows = host_search == 1 ? rows_search(c) : rows_cached();
  reduce_blob_kernel<<<1, 768, 0, c->stream>>>(c->d_partials, c->d_scan_m, rows, host_search < 0 ? c->d_ctrl : nullptr,
                                               rows_search(c), rows_cached(), c->d_blob);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}
int launch_reduce_blob(lio_ctx* c) { return launch_reduce_blob_mode(c, -1); }

int launch_solve(lio_ctx* c, double R, int external_blob) {
  SolveArgs s;
  s.partials = c->d_partials;
  s.scan_m = c->d_scan_m;
  s.rows_search = rows_search(c);
  s.rows_cached = rows_cached();
  s.blob_in = external_blob ? c->d_blob : nullptr;
  s.blob_out = c->d_blob;
  s.x = c->d_x;
  s.xprop = c->d_xprop;
  s.P = c->d_P;
  s.ctrl = c->d_ctrl;
  s.dx_out = c->d_dx;
  s.R = R;
  solve_kernel<<<1, 768, 0, c->stream>>>(s);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int launch_begin(lio_ctx* c, int max_iter, int from_snapshot) {
  begin_kernel<<<1, 128, 0, c->stream>>>(c->d_ctrl, c->d_x, c->d_xprop, c->d_P, c->d_x0, c->d_P0, max_iter,
                                         from_snapshot);
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
