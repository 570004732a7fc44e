// GPU voxel-hash map: the observable semantics of the reference ikd-Tree calls on the hot path
//   Build (ikd_Tree.cpp:355-367) / Add_Points (:419-512) / Delete_Point_Boxes (:559-579) / flatten (:1490-1516)
// on a hash of kNN cells, each owning one contiguous float4 bucket in a point pool (x,y,z,id bits; id < 0 = dead).
// Inserts are batched in three phases (reserve -> grow -> fill) so no thread ever waits on another.
// map_incremental (laserMapping.cpp:382-433) is classified and applied on the device as well.
#include "lio_ctx.cuh"

namespace lio {

#define BASE_SENTINEL 0xFFFFFFFFu

__global__ void map_init_kernel(MapView m, uint32_t hash_cap) {
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < hash_cap; h += gridDim.x * blockDim.x) {
    m.table[h].key = LIO_EMPTY_KEY;
    m.table[h].start = 0;
    m.table[h].count = 0;
    m.cell_cap[h] = 0;
    m.cell_pend[h] = 0;
    m.cell_base[h] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x < 16)
    m.counters[threadIdx.x] = threadIdx.x >= 8 && threadIdx.x < 11    ? 0x7fffffffu               // cell box: empty
                              : threadIdx.x >= 11 && threadIdx.x < 14 ? (uint32_t)(-0x7fffffff)
                                                                      : 0u;
}

// Batch size and id base of an insert: host values, or -- when the host enqueues a whole scan without waiting for the
// counts (lio_scan_step) -- values other kernels of the same stream left on the device.  n is then the launch bound.
struct BatchDev {
  const int* n_dev;       // nullptr: n is exact
  const int* id_off_dev;  // nullptr: ids start at id_base
};
__device__ __forceinline__ int batch_n(int n, const BatchDev& b) {
  if (b.n_dev == nullptr) return n;
  const int v = *b.n_dev;
  return v < n ? v : n;
}
__device__ __forceinline__ int batch_id(int id_base, const BatchDev& b) {
  return id_base + (b.id_off_dev ? *b.id_off_dev : 0);
}

// phase 1: find-or-create the cell of every point, take a rank inside this batch; the box of cells ever used
// (counters[8..13], what an unbounded search has to cover) grows by a block-level reduction and at most six atomics
__global__ void __launch_bounds__(256) map_reserve_kernel(MapView m, const float4* pts, int n, BatchDev bd,
                                                          const uint8_t* flag, uint32_t* slot, uint32_t* rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  n = batch_n(n, bd);
  const bool act = i < n && (!flag || flag[i]);
  int cc[3] = {0, 0, 0};
  if (act) {
    const float4 p = pts[i];
    cc[0] = cell_coord(p.x, m.inv_cell);
    cc[1] = cell_coord(p.y, m.inv_cell);
    cc[2] = cell_coord(p.z, m.inv_cell);
    const unsigned long long key = pack_cell(cc[0], cc[1], cc[2]);
    uint32_t h = hash64(key) & m.hash_mask;
    uint32_t probes = 0;
    bool full = false;
    for (;;) {
      const unsigned long long prev = atomicCAS(&m.table[h].key, LIO_EMPTY_KEY, key);
      if (prev == LIO_EMPTY_KEY) {
        atomicAdd(&m.counters[1], 1u);
        break;
      }
      if (prev == key) break;
      h = (h + 1) & m.hash_mask;
      if (++probes > m.hash_mask) {
        full = true;
        break;
      }
    }
    if (full) {
      atomicExch(&m.counters[3], 1u);
      slot[i] = BASE_SENTINEL;
    } else {
      slot[i] = h;
      rank[i] = atomicAdd(&m.cell_pend[h], 1u);
    }
  }
  __shared__ int s_red[6][8];
  const int big = 0x7fffffff;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int mn = __reduce_min_sync(0xffffffffu, act ? cc[a] : big);
    const int mx = __reduce_max_sync(0xffffffffu, act ? cc[a] : -big);
    if (lane == 0) {
      s_red[a][warp] = mn;
      s_red[3 + a][warp] = mx;
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_min = threadIdx.x < 3;
    int v = s_red[threadIdx.x][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) v = is_min ? min(v, s_red[threadIdx.x][w]) : max(v, s_red[threadIdx.x][w]);
    int* g = reinterpret_cast<int*>(m.counters) + 8 + threadIdx.x;
    const int cur = *reinterpret_cast<volatile int*>(g);  // most blocks do not extend the box: a load, no atomic
    if (is_min) {
      if (v < cur) atomicMin(g, v);
    } else {
      if (v > cur) atomicMax(g, v);
    }
  }
}

// phase 2: one thread per touched cell makes room (amortised doubling; the live points move, the old run is retired)
__global__ void map_grow_kernel(MapView m, int n, BatchDev bd, const uint8_t* flag, const uint32_t* slot,
                                const uint32_t* rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  n = batch_n(n, bd);
  if (i >= n) return;
  if (flag && !flag[i]) return;
  const uint32_t h = slot[i];
  if (h == BASE_SENTINEL || rank[i] != 0) return;
  uint32_t cnt = m.table[h].count;
  const uint32_t need = cnt + m.cell_pend[h];
  if (need > m.cell_cap[h]) {
    uint32_t newcap = 4;
    while (newcap < need) newcap <<= 1;
    const uint32_t newstart = atomicAdd(&m.counters[0], newcap);
    if ((unsigned long long)newstart + newcap > m.pool_cap) {
      atomicExch(&m.counters[3], 2u);
      m.cell_base[h] = BASE_SENTINEL;
      return;
    }
    const uint32_t old = m.table[h].start;
    uint32_t live = 0;
    for (uint32_t t = 0; t < cnt; ++t) {
      float4 p = m.pool[old + t];
      if (__float_as_int(p.w) >= 0) {
        m.pool[newstart + live++] = p;
        p.w = __int_as_float(-1);
        m.pool[old + t] = p;  // retire the old slot so pool sweeps see each live point once
      }
    }
    m.table[h].start = newstart;
    m.table[h].count = live;
    m.cell_cap[h] = newcap;
    cnt = live;
  }
  m.cell_base[h] = cnt;
}

// phase 3: write the points
__global__ void map_fill_kernel(MapView m, const float4* pts, int n, BatchDev bd, const uint8_t* flag,
                                const uint32_t* slot, const uint32_t* rank, int id_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  n = batch_n(n, bd);
  id_base = batch_id(id_base, bd);
  bool wrote = false;
  if (i < n && (!flag || flag[i])) {
    const uint32_t h = slot[i];
    if (h != BASE_SENTINEL) {
      const uint32_t base = m.cell_base[h];
      if (base != BASE_SENTINEL) {
        float4 p = pts[i];
        p.w = __int_as_float(id_base + i);
        m.pool[m.table[h].start + base + rank[i]] = p;
        wrote = true;
        if (rank[i] == 0) {
          m.table[h].count = base + m.cell_pend[h];
          m.cell_pend[h] = 0;
        }
      }
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, wrote);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&m.counters[2], (uint32_t)__popc(b));
}

// ---- Add_Points with downsample --------------------------------------------------------------------------------
struct VoxGeom {
  float bmin[3], bmax[3], mid[3];
  int k[3];
};
__device__ __forceinline__ VoxGeom vox_geom(const float4 p, float ds) {
  VoxGeom g;
  const float v[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float f = floorf(v[a] / ds);
    g.k[a] = (int)f;
    g.bmin[a] = f * ds + 0.0f;  // ikd_Tree.cpp:431-439 (FP32); +0 folds -0 into +0
    g.bmax[a] = g.bmin[a] + ds;
    g.mid[a] = (float)((double)g.bmin[a] + (double)(g.bmax[a] - g.bmin[a]) / 2.0);  // :440-448
  }
  return g;
}

__global__ void vox_init_kernel(unsigned long long* vkey, unsigned long long* vbest, uint32_t cap) {
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += gridDim.x * blockDim.x) {
    vkey[h] = LIO_EMPTY_KEY;
    vbest[h] = ~0ull;
  }
}

// per batch voxel: nearest-to-centre new point; ties go to the LATER point (a new point wins ties, :456-462)
__global__ void vox_best_kernel(const float4* pts, int n, BatchDev bd, float ds, unsigned long long* vkey,
                                unsigned long long* vbest, uint32_t vmask, uint32_t* vslot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  n = batch_n(n, bd);
  if (i >= n) return;
  const float4 p = pts[i];
  const VoxGeom g = vox_geom(p, ds);
  const float d = dist2(p.x, p.y, p.z, g.mid[0], g.mid[1], g.mid[2]);
  const unsigned long long key = pack_cell(g.k[0], g.k[1], g.k[2]);
  uint32_t h = hash64(key) & vmask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&vkey[h], LIO_EMPTY_KEY, key);
    if (prev == LIO_EMPTY_KEY || prev == key) break;
    h = (h + 1) & vmask;
  }
  vslot[i] = h;
  atomicMin(&vbest[h], ((unsigned long long)__float_as_uint(d) << 32) | (0xFFFFFFFFu - (uint32_t)i));
}

// the winner of each batch voxel settles it against the points already in the map
__global__ void vox_apply_kernel(MapView m, const float4* pts, int n, BatchDev bd, float ds,
                                 const unsigned long long* vbest, const uint32_t* vslot, int id_base,
                                 uint8_t* append_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  n = batch_n(n, bd);
  id_base = batch_id(id_base, bd);
  if (i >= n) return;
  append_flag[i] = 0;
  const unsigned long long best = vbest[vslot[i]];
  if ((uint32_t)best != 0xFFFFFFFFu - (uint32_t)i) return;
  const float4 p = pts[i];
  const VoxGeom g = vox_geom(p, ds);
  const float dnew = __uint_as_float((uint32_t)(best >> 32));
  int c0[3], c1[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // the box is half open: its last cell is the one of the largest float below bmax (cell_coord is monotone), not the
    // one of bmax itself -- with 0.5 m boxes in 1.5 m cells that is one cell instead of up to eight
    c0[a] = cell_coord(g.bmin[a], m.inv_cell);
    c1[a] = cell_coord(nextafterf(g.bmax[a], -INFINITY), m.inv_cell);
  }
  // pass 1: existing points inside the half-open box
  int E = 0;
  float best_d = dnew;
  int best_id = -1;
  uint32_t best_slot = 0;
  for (int cz = c0[2]; cz <= c1[2]; ++cz)
    for (int cy = c0[1]; cy <= c1[1]; ++cy)
      for (int cx = c0[0]; cx <= c1[0]; ++cx) {
        uint32_t start, count;
        if (map_find(m, pack_cell(cx, cy, cz), start, count) < 0) continue;
        for (uint32_t t = 0; t < count; ++t) {
          const float4 q = m.pool[start + t];
          const int id = __float_as_int(q.w);
          if (id < 0) continue;
          if (g.bmin[0] <= q.x && g.bmax[0] > q.x && g.bmin[1] <= q.y && g.bmax[1] > q.y && g.bmin[2] <= q.z &&
              g.bmax[2] > q.z) {
            ++E;
            const float d = dist2(q.x, q.y, q.z, g.mid[0], g.mid[1], g.mid[2]);
            if (d < best_d || (best_id >= 0 && d == best_d && id < best_id)) {
              best_d = d;
              best_id = id;
              best_slot = start + t;
            }
          }
        }
      }
  if (best_id >= 0 && E <= 1) return;  // the single existing point stays (:463-465)
  if (E == 0) {
    append_flag[i] = 1;  // new voxel: plain insert
    atomicAdd(&m.counters[4], 1u);
    return;
  }
  // collapse the voxel to the winner: an existing winner keeps its slot; a new winner retires every point of the box
  // and goes through the normal append (reserve / grow / fill) so that it is filed in the bucket of ITS OWN cell -- the
  // voxel need not lie inside one kNN cell (any downsample size is allowed), and a point filed under a neighbour's
  // cell would be invisible to the searches that rely on "the 3x3x3 block holds everything within one cell edge"
  for (int cz = c0[2]; cz <= c1[2]; ++cz)
    for (int cy = c0[1]; cy <= c1[1]; ++cy)
      for (int cx = c0[0]; cx <= c1[0]; ++cx) {
        uint32_t start, count;
        if (map_find(m, pack_cell(cx, cy, cz), start, count) < 0) continue;
        for (uint32_t t = 0; t < count; ++t) {
          float4 q = m.pool[start + t];
          if (__float_as_int(q.w) < 0) continue;
          if (g.bmin[0] <= q.x && g.bmax[0] > q.x && g.bmin[1] <= q.y && g.bmax[1] > q.y && g.bmin[2] <= q.z &&
              g.bmax[2] > q.z && (best_id < 0 || start + t != best_slot)) {
            log_removed(m, q);
            q.w = __int_as_float(-1);
            m.pool[start + t] = q;
          }
        }
      }
  if (best_id < 0) {
    append_flag[i] = 1;
    atomicAdd(&m.counters[4], 1u);
    atomicSub(&m.counters[2], (uint32_t)E);
  } else {
    atomicSub(&m.counters[2], (uint32_t)(E - 1));
  }
}

// ---- Delete_Point_Boxes / flatten: pool sweeps --------------------------------------------------------------------
__global__ void map_delete_kernel(MapView m, const float* boxes6, int nb, uint32_t* n_deleted) {
  const uint32_t top = m.counters[0];
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < top; s += gridDim.x * blockDim.x) {
    float4 q = m.pool[s];
    if (__float_as_int(q.w) < 0) continue;
    bool hit = false;
    for (int b = 0; b < nb && !hit; ++b) {
      const float* mn = boxes6 + 6 * b;
      hit = mn[0] <= q.x && mn[3] > q.x && mn[1] <= q.y && mn[4] > q.y && mn[2] <= q.z && mn[5] > q.z;
    }
    if (hit) {
      log_removed(m, q);
      q.w = __int_as_float(-1);
      m.pool[s] = q;
      atomicAdd(n_deleted, 1u);
      atomicSub(&m.counters[2], 1u);
    }
  }
}

__global__ void map_dump_kernel(MapView m, float4* out, uint32_t cap, uint32_t* n_out) {
  const uint32_t top = m.counters[0];
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < top; s += gridDim.x * blockDim.x) {
    const float4 q = m.pool[s];
    if (__float_as_int(q.w) < 0) continue;
    const uint32_t j = atomicAdd(n_out, 1u);
    if (j < cap) out[j] = q;
  }
}

// ---- map_incremental (laserMapping.cpp:382-433) -------------------------------------------------------------------
// class 0 = skip, 1 = PointToAdd (Add_Points with downsample), 2 = PointNoNeedDownsample.
// The reference reads Nearest_Points[i] as its UNBOUNDED search left it (esekfom.hpp:140-141: min(5, #live points)
// neighbours, however far).  Here a row holds a known PREFIX of that list, near_cnt[i] long: the neighbours within
// d2 <= 5 from the update's search, extended by far_search_kernel to at least one entry (map_incremental_enqueue).
// That decides every row exactly as the full list would:
//   * `!Nearest_Points[i].empty()` and `points_near.size() < NUM_MATCH_POINTS` depend on the map only (n_live);
//   * points_near[0] is in the prefix;
//   * a neighbour beyond the prefix is farther than sqrt(5) m from the point, which itself lies within sqrt(3)/2 * fs
//     of its voxel centre: for 3 fs^2 <= 5 it cannot be closer to the centre than the point is (:416-422 stays false).
//     For larger filter_size_map the caller completes the rows to all five entries first.
__global__ void map_incr_classify_kernel(const float4* body, const int* scan_m, const StateD* xs, const float4* near_pts,
                                         const int* near_cnt, const uint32_t* map_counters, int ekf_inited, float fsm,
                                         int min_m, float4* world, uint8_t* cls) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = *scan_m;
  if (i >= M || M < min_m) return;  // fewer than min_m points: the main loop skips the scan (laserMapping.cpp:741-744)
  // pointBodyToWorld (laserMapping.cpp:277-288): rotation MATRICES here, FP64 -> FP32
  double R[9], Rli[9];
  quat_to_mat(xs->rot, R);
  quat_to_mat(xs->rli, Rli);
  const float4 b = body[i];
  const double p[3] = {b.x, b.y, b.z};
  double a[3], w[3];
  mat3_vec(Rli, p, a);
  a[0] += xs->tli[0];
  a[1] += xs->tli[1];
  a[2] += xs->tli[2];
  mat3_vec(R, a, w);
  const float pw[3] = {(float)(w[0] + xs->pos[0]), (float)(w[1] + xs->pos[1]), (float)(w[2] + xs->pos[2])};
  world[i] = make_float4(pw[0], pw[1], pw[2], b.w);
  uint8_t c = 1;
  const int cnt = near_cnt[i];
  const uint32_t n_live = map_counters[2];  // size of the reference's list = min(5, n_live)
  if (cnt > 0 && ekf_inited) {
    const double fs = (double)fsm;
    float mid[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) mid[k] = (float)(floor((double)pw[k] / fs) * fs + 0.5 * fs);
    const float dist = dist2(pw[0], pw[1], pw[2], mid[0], mid[1], mid[2]);
    const float4 n0 = near_pts[(size_t)i * LIO_K];
    if (fabs((double)(n0.x - mid[0])) > 0.5 * fs && fabs((double)(n0.y - mid[1])) > 0.5 * fs &&
        fabs((double)(n0.z - mid[2])) > 0.5 * fs) {
      c = 2;
    } else {
      bool need_add = true;
      if (n_live >= LIO_K) {  // `if (points_near.size() < NUM_MATCH_POINTS) break;`
        for (int j = 0; j < cnt; ++j) {
          const float4 q = near_pts[(size_t)i * LIO_K + j];
          if (dist2(q.x, q.y, q.z, mid[0], mid[1], mid[2]) < dist) {
            need_add = false;
            break;
          }
        }
      }
      c = need_add ? 1 : 0;
    }
  }
  cls[i] = c;
}

// pointBodyToWorld (laserMapping.cpp:277-288) for the whole current scan: the first-scan map build (:747-758)
__global__ void scan_to_world_kernel(const float4* body, const int* scan_m, const StateD* xs, float4* world) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *scan_m) return;
  double R[9], Rli[9];
  quat_to_mat(xs->rot, R);
  quat_to_mat(xs->rli, Rli);
  const float4 b = body[i];
  const double p[3] = {b.x, b.y, b.z};
  double a[3], w[3];
  mat3_vec(Rli, p, a);
  a[0] += xs->tli[0];
  a[1] += xs->tli[1];
  a[2] += xs->tli[2];
  mat3_vec(R, a, w);
  world[i] = make_float4((float)(w[0] + xs->pos[0]), (float)(w[1] + xs->pos[1]), (float)(w[2] + xs->pos[2]), b.w);
}

// Order-preserving compaction of the two classes by a single block (M <= ~1e5): PointToAdd first, then
// PointNoNeedDownsample right behind it in the SAME array, so that both Add_Points calls of laserMapping.cpp:430-431 end in
// one reserve / grow / fill sequence: vox_apply_kernel decides append_flag for the first n_a points, the others are
// appended unconditionally (flag 1, written here).  The id of a point is id_base + its position, exactly what the two
// separate calls hand out (the second call's ids start after all n_a points of the first).
__global__ void __launch_bounds__(1024) map_incr_compact_kernel(const float4* world, const uint8_t* cls,
                                                                const int* scan_m, int min_m, float4* out,
                                                                uint8_t* append_flag, int* counts /*[0] n_a [1] n_b [3] sum*/) {
  __shared__ int wa[32], wb[32];
  const int M = *scan_m < min_m ? 0 : *scan_m;
  const int chunk = (M + 1023) / 1024;
  const int lo = min(M, (int)threadIdx.x * chunk), hi = min(M, lo + chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int na = 0, nb = 0;
  for (int i = lo; i < hi; ++i) {
    na += cls[i] == 1;
    nb += cls[i] == 2;
  }
  int ia = na, ib = nb;  // inclusive scans: warp, then the 32 warp totals
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int ta = __shfl_up_sync(0xffffffffu, ia, off), tb = __shfl_up_sync(0xffffffffu, ib, off);
    if (lane >= off) {
      ia += ta;
      ib += tb;
    }
  }
  if (lane == 31) {
    wa[warp] = ia;
    wb[warp] = ib;
  }
  __syncthreads();
  if (warp == 0) {
    int va = wa[lane], vb = wb[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int ta = __shfl_up_sync(0xffffffffu, va, off), tb = __shfl_up_sync(0xffffffffu, vb, off);
      if (lane >= off) {
        va += ta;
        vb += tb;
      }
    }
    wa[lane] = va;
    wb[lane] = vb;
  }
  __syncthreads();
  const int tot_a = wa[31], tot_b = wb[31];
  int oa = (warp ? wa[warp - 1] : 0) + ia - na;
  int ob = tot_a + (warp ? wb[warp - 1] : 0) + ib - nb;
  for (int i = lo; i < hi; ++i) {
    if (cls[i] == 1) out[oa++] = world[i];
    if (cls[i] == 2) {
      append_flag[ob] = 1;
      out[ob++] = world[i];
    }
  }
  if (threadIdx.x == 0) {
    counts[0] = tot_a;
    counts[1] = tot_b;
    counts[3] = tot_a + tot_b;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int check_map_error(lio_ctx* c) {
  uint32_t h[8];
  LIO_CHECK(c, cudaMemcpyAsync(h, c->map.counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  if (h[3] != 0) {
    c->err = h[3] == 1 ? "map hash table full (raise lio_caps.max_map_points)"
                       : "map point pool full (raise lio_caps.max_map_points)";
    return LIO_E_CAPACITY;
  }
  return LIO_OK;
}

int map_reset(lio_ctx* c) {
  // every pool slot starts dead (id bits 0xFFFFFFFF < 0) so that pool sweeps (dump / delete) never see garbage
  LIO_CHECK(c, cudaMemsetAsync(c->map.pool, 0xFF, sizeof(float4) * (size_t)c->map.pool_cap, c->stream));
  map_init_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->map, c->hash_cap);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  c->next_id = 0;
  c->map_built = false;
  return LIO_OK;
}

// Enqueue only.  n is exact (n_dev == nullptr) or the launch bound of a count that lives on the device.
int map_append_batch_enqueue(lio_ctx* c, const float4* d_pts, int64_t n, const int* n_dev, int32_t id_base,
                             const int* id_off_dev, const uint8_t* d_flag) {
  if (n <= 0) return LIO_OK;
  if (n > c->batch_cap) {
    c->err = "insert batch larger than the context's batch capacity";
    return LIO_E_CAPACITY;
  }
  const int grid = (int)((n + 255) / 256);
  const BatchDev bd{n_dev, id_off_dev};
  map_reserve_kernel<<<grid, 256, 0, c->stream>>>(c->map, d_pts, (int)n, bd, d_flag, c->d_batch_slot, c->d_batch_rank);
  map_grow_kernel<<<grid, 256, 0, c->stream>>>(c->map, (int)n, bd, d_flag, c->d_batch_slot, c->d_batch_rank);
  map_fill_kernel<<<grid, 256, 0, c->stream>>>(c->map, d_pts, (int)n, bd, d_flag, c->d_batch_slot, c->d_batch_rank,
                                               id_base);
  c->launches += 3;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int map_append_batch(lio_ctx* c, const float4* d_pts, int64_t n, int32_t id_base, const uint8_t* d_flag) {
  if (n <= 0) return LIO_OK;
  const int rc = map_append_batch_enqueue(c, d_pts, n, nullptr, id_base, nullptr, d_flag);
  if (rc) return rc;
  return check_map_error(c);
}

// Add_Points(downsample_on = true), first half, enqueue only: per batch voxel the winner settles it against the map and
// append_flag says which points remain to be appended; the number of points added ends up in map.counters[4].
static int vox_phase_enqueue(lio_ctx* c, const float4* d_pts, int64_t n, const int* n_dev, int32_t id_base, float ds) {
  if (n > c->batch_cap || (uint64_t)n * 2 > c->vox_cap) {
    c->err = "Add_Points batch larger than the context's batch capacity";
    return LIO_E_CAPACITY;
  }
  const int grid = (int)((n + 255) / 256);
  const BatchDev bd{n_dev, nullptr};
  LIO_CHECK(c, cudaMemsetAsync(c->map.counters + 4, 0, sizeof(uint32_t), c->stream));
  vox_init_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(c->d_vox_key, c->d_vox_best, c->vox_cap);
  vox_best_kernel<<<grid, 256, 0, c->stream>>>(d_pts, (int)n, bd, ds, c->d_vox_key, c->d_vox_best, c->vox_cap - 1,
                                               c->d_batch_rank /*reused as vslot*/);
  vox_apply_kernel<<<grid, 256, 0, c->stream>>>(c->map, d_pts, (int)n, bd, ds, c->d_vox_best, c->d_batch_rank, id_base,
                                                c->d_batch_flag);
  c->launches += 3;
  LIO_CHECK(c, cudaGetLastError());
  return LIO_OK;
}

int map_add_downsample_enqueue(lio_ctx* c, const float4* d_pts, int64_t n, const int* n_dev, int32_t id_base, float ds) {
  if (n <= 0) return LIO_OK;
  const int rc = vox_phase_enqueue(c, d_pts, n, n_dev, id_base, ds);
  if (rc) return rc;
  return map_append_batch_enqueue(c, d_pts, n, n_dev, id_base, nullptr, c->d_batch_flag);
}

int map_add_downsample(lio_ctx* c, const float4* d_pts, int64_t n, int32_t id_base, int32_t* n_added) {
  if (n_added) *n_added = 0;
  if (n <= 0) return LIO_OK;
  int rc = map_add_downsample_enqueue(c, d_pts, n, nullptr, id_base, c->map_downsample);
  if (rc) return rc;
  rc = check_map_error(c);
  if (rc) return rc;
  if (n_added) {
    uint32_t v = 0;
    LIO_CHECK(c, cudaMemcpyAsync(&v, c->map.counters + 4, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    LIO_CHECK(c, cudaStreamSynchronize(c->stream));
    *n_added = (int32_t)v;
  }
  return LIO_OK;
}

int map_delete_boxes(lio_ctx* c, const float* h_boxes6, int nb, int32_t* n_deleted) {
  if (n_deleted) *n_deleted = 0;
  if (nb <= 0) return LIO_OK;
  float* d_boxes = nullptr;
  uint32_t* d_n = nullptr;
  LIO_CHECK(c, cudaMallocAsync(&d_boxes, sizeof(float) * 6 * nb, c->stream));
  LIO_CHECK(c, cudaMallocAsync(&d_n, sizeof(uint32_t), c->stream));
  LIO_CHECK(c, cudaMemcpyAsync(d_boxes, h_boxes6, sizeof(float) * 6 * nb, cudaMemcpyHostToDevice, c->stream));
  LIO_CHECK(c, cudaMemsetAsync(d_n, 0, sizeof(uint32_t), c->stream));
  map_delete_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->map, d_boxes, nb, d_n);
  c->launches++;
  uint32_t v = 0;
  LIO_CHECK(c, cudaMemcpyAsync(&v, d_n, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  cudaFreeAsync(d_boxes, c->stream);
  cudaFreeAsync(d_n, c->stream);
  if (n_deleted) *n_deleted = (int32_t)v;
  return LIO_OK;
}

}  // namespace lio

#include <algorithm>
#include <vector>

namespace lio {

// ≙ KD_TREE::acquire_removed_points: the points deleted since the last call that took them (fetch and clear).
int map_removed_points(lio_ctx* c, float* xyz, int64_t cap, int64_t* n) {
  uint32_t h[8];
  LIO_CHECK(c, cudaMemcpyAsync(h, c->map.counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  const int64_t logged = std::min<int64_t>(h[6], c->map.removed_cap);
  if (n) *n = logged;
  if (!xyz) return LIO_OK;  // a question about the size only
  const int64_t take = std::min<int64_t>(logged, cap);
  std::vector<float4> buf((size_t)take);
  if (take > 0)
    LIO_CHECK(c, cudaMemcpyAsync(buf.data(), c->map.removed, sizeof(float4) * (size_t)take, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaMemsetAsync(c->map.counters + 6, 0, 2 * sizeof(uint32_t), c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < take; ++i) {
    xyz[3 * i] = buf[(size_t)i].x;
    xyz[3 * i + 1] = buf[(size_t)i].y;
    xyz[3 * i + 2] = buf[(size_t)i].z;
  }
  if (h[7] != 0) {
    c->err = "removed-point log overflowed: the oldest entries were kept, later deletions of this period are missing";
    return LIO_E_CAPACITY;
  }
  return LIO_OK;
}

int map_dump(lio_ctx* c, float* xyz, int32_t* ids, int64_t cap, int64_t* n) {
  uint32_t h[8];
  LIO_CHECK(c, cudaMemcpyAsync(h, c->map.counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  const uint32_t live = h[2];
  if (n) *n = live;
  if ((!xyz && !ids) || live == 0) return LIO_OK;
  float4* d_out = nullptr;
  uint32_t* d_n = nullptr;
  LIO_CHECK(c, cudaMallocAsync(&d_out, sizeof(float4) * (size_t)live, c->stream));
  LIO_CHECK(c, cudaMallocAsync(&d_n, sizeof(uint32_t), c->stream));
  LIO_CHECK(c, cudaMemsetAsync(d_n, 0, sizeof(uint32_t), c->stream));
  map_dump_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->map, d_out, live, d_n);
  c->launches++;
  std::vector<float4> host((size_t)live);
  LIO_CHECK(c, cudaMemcpyAsync(host.data(), d_out, sizeof(float4) * (size_t)live, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  cudaFreeAsync(d_out, c->stream);
  cudaFreeAsync(d_n, c->stream);
  std::sort(host.begin(), host.end(), [](const float4& a, const float4& b) {
    int ia, ib;
    memcpy(&ia, &a.w, 4);
    memcpy(&ib, &b.w, 4);
    return ia < ib;
  });
  for (int64_t i = 0; i < (int64_t)live && i < cap; ++i) {
    if (xyz) {
      xyz[3 * i] = host[i].x;
      xyz[3 * i + 1] = host[i].y;
      xyz[3 * i + 2] = host[i].z;
    }
    if (ids) memcpy(&ids[i], &host[i].w, 4);
  }
  return LIO_OK;
}

int map_build_scan(lio_ctx* c, const lio_state* x) {
  int rc = map_reset(c);
  if (rc) return rc;
  if (c->scan_m <= 0) return LIO_OK;
  if (c->scan_m > c->caps.max_map_points) {
    c->err = "Build: more points than lio_caps.max_map_points";
    return LIO_E_CAPACITY;
  }
  LIO_CHECK(c, cudaMemcpyAsync(c->d_x, x, sizeof(lio_state), cudaMemcpyHostToDevice, c->stream));
  scan_to_world_kernel<<<(int)((c->scan_m + 255) / 256), 256, 0, c->stream>>>(c->d_body, c->d_scan_m, c->d_x,
                                                                               c->d_batch_pts);
  c->launches++;
  LIO_CHECK(c, cudaGetLastError());
  rc = map_append_batch(c, c->d_batch_pts, c->scan_m, 0, nullptr);
  if (rc) return rc;
  c->next_id = (int32_t)c->scan_m;
  c->map_built = true;
  return LIO_OK;
}

// map_incremental on the state at c->d_x, enqueue only: classification, ordered compaction, both Add_Points calls.  The
// class counts stay on the device (d_prep_counters[8..9]) and size the inserts there; bound = upper bound of M.
int map_incremental_enqueue(lio_ctx* c, float fsm, int ekf_inited, int min_m, int64_t bound) {
  if (bound <= 0) {  // empty scan: the counts a later report reads must still be this scan's
    LIO_CHECK(c, cudaMemsetAsync(c->d_prep_counters + 8, 0, 2 * sizeof(int), c->stream));
    LIO_CHECK(c, cudaMemsetAsync(c->map.counters + 4, 0, sizeof(uint32_t), c->stream));
    return LIO_OK;
  }
  const int grid = (int)((bound + 255) / 256);
  int* d_counts = c->d_prep_counters + 8;
  // rows the update's bounded search left empty get their nearest neighbour from the unbounded search -- all five
  // when filter_size_map is so large that the prefix argument of map_incr_classify_kernel does not hold
  const int need = (3.0f * fsm * fsm <= 0.95f * c->caps.knn_max_d2) ? 1 : LIO_K;
  int rc = launch_far_complete(c, c->d_near_q, -1, min_m, bound, need);
  if (rc) return rc;
  map_incr_classify_kernel<<<grid, 256, 0, c->stream>>>(c->d_body, c->d_scan_m, c->d_x, c->d_near, c->d_near_cnt,
                                                        c->map.counters, ekf_inited, fsm, min_m, c->d_world, c->d_cls);
  map_incr_compact_kernel<<<1, 1024, 0, c->stream>>>(c->d_world, c->d_cls, c->d_scan_m, min_m, c->d_add_a,
                                                     c->d_batch_flag, d_counts);
  c->launches += 2;
  // Add_Points(PointToAdd, true) decides on the first n_a points; one reserve / grow / fill appends what it leaves of
  // them together with Add_Points(PointNoNeedDownsample, false).  The downsample size is filter_size_map_min, as for
  // the classification (laserMapping.cpp:748: the same value goes into set_downsample_param).
  rc = vox_phase_enqueue(c, c->d_add_a, bound, d_counts, c->next_id, fsm);
  if (rc) return rc;
  return map_append_batch_enqueue(c, c->d_add_a, bound, d_counts + 3, c->next_id, nullptr, c->d_batch_flag);
}

int map_incremental(lio_ctx* c, const lio_state* x, float fsm, int ekf_inited, int32_t counts[3]) {
  if (!c->map_built) {
    c->err = "map_incremental needs a built map";
    return LIO_E_EMPTY_MAP;
  }
  counts[0] = counts[1] = counts[2] = 0;
  if (c->scan_m <= 0) return LIO_OK;
  if (const int ri = check_id_space(c, c->scan_m)) return ri;
  LIO_CHECK(c, cudaMemcpyAsync(c->d_x, x, sizeof(lio_state), cudaMemcpyHostToDevice, c->stream));
  int rc = map_incremental_enqueue(c, fsm, ekf_inited, 0, c->scan_m);
  if (rc) return rc;
  int hc[2] = {0, 0};
  uint32_t added = 0;
  LIO_CHECK(c, cudaMemcpyAsync(hc, c->d_prep_counters + 8, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaMemcpyAsync(&added, c->map.counters + 4, sizeof(added), cudaMemcpyDeviceToHost, c->stream));
  rc = check_map_error(c);  // synchronises
  if (rc) return rc;
  counts[0] = hc[0];
  counts[1] = hc[1];
  counts[2] = (int32_t)added;
  c->next_id += hc[0] + hc[1];
  return LIO_OK;
}

}  // namespace lio
