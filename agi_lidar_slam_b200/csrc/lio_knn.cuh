// Group-cooperative bounded 5-nearest-neighbour search on the voxel-hash map, and the FP32 plane fit.
// Replaces KD_TREE::Nearest_Search (ikd_Tree.cpp:370-402, 960-1101) and esti_plane (common_lib.h:102-134).
#pragma once
#include <float.h>

#include "lio_common.cuh"

namespace lio {

// Lane-local candidate list, ascending by key = (d2 bits << 32 | id).  d2 >= 0 so its IEEE bit pattern orders
// like the value; ids are unique, so the key order is the canonical (d2, id) order of the parity contract.
struct TopK {
  unsigned long long k0, k1, k2, k3, k4;
  uint32_t s0, s1, s2, s3, s4;  // pool slots
  __device__ __forceinline__ void init() {
    k0 = k1 = k2 = k3 = k4 = ~0ull;
    s0 = s1 = s2 = s3 = s4 = 0;
  }
  __device__ __forceinline__ void insert(unsigned long long k, uint32_t s) {
    if (k < k4) {
      k4 = k;
      s4 = s;
      if (k4 < k3) {
        unsigned long long tk = k3; k3 = k4; k4 = tk;
        uint32_t ts = s3; s3 = s4; s4 = ts;
        if (k3 < k2) {
          tk = k2; k2 = k3; k3 = tk;
          ts = s2; s2 = s3; s3 = ts;
          if (k2 < k1) {
            tk = k1; k1 = k2; k2 = tk;
            ts = s1; s1 = s2; s2 = ts;
            if (k1 < k0) {
              tk = k0; k0 = k1; k1 = tk;
              ts = s0; s0 = s1; s1 = ts;
            }
          }
        }
      }
    }
  }
  __device__ __forceinline__ void pop() {
    k0 = k1; k1 = k2; k2 = k3; k3 = k4; k4 = ~0ull;
    s0 = s1; s1 = s2; s2 = s3; s3 = s4;
  }
};

// ---------------------------------------------------------------------------------------------------------
// Shared-memory staging of the neighbour cells of a query TILE (north_star (1)).  The scan arrives sorted by voxel, so
// the queries of a tile are spatial neighbours and share most of the 27 cells each of them has to look at.  Per tile:
//   1. every (query, neighbour cell) pair is entered into a cell SET in shared memory (open addressing on the packed
//      cell key); the pair remembers its slot,
//   2. one thread per distinct cell probes the map's hash table (one dependent round trip for the whole tile),
//   3. the cell's bucket -- one contiguous run of float4 {x,y,z,id} in the point pool -- is brought into shared memory
//      by ONE bulk asynchronous copy (cp.async.bulk, completion counted in bytes on an mbarrier): no registers, no
//      per-point instructions, all copies of the tile in flight together,
//   4. the searches of the tile then stream their candidates from shared memory.
// Every distinct bucket crosses L2 -> SM once per tile instead of once per query that needs it.
// ---------------------------------------------------------------------------------------------------------
constexpr int KNN_CELLS = 32;   // cells of one round of a group's search (G lanes x ceil(27 / G) cells: 32 for G = 8, 16, 32)
constexpr int ST_TILE = 64;     // queries per tile at most
constexpr int ST_HASH = 2048;   // slots of the cell set (27 * ST_TILE = 1728 distinct cells at most)
constexpr int ST_PTS = 4096;    // staged map points per tile (64 KB); a tile that needs more searches global memory
constexpr uint32_t ST_FLAG = 0x80000000u;  // slot values with this bit index StageSmem::pts, the others the point pool
struct __align__(16) StageSmem {
  float4 pts[ST_PTS];
  unsigned long long key[ST_HASH];  // packed cell key or LIO_EMPTY_KEY
  uint32_t cell[ST_HASH];           // staged offset << 12 | count
  uint16_t list[ST_HASH];           // the slots in use
  uint16_t qslot[ST_TILE * 27];     // slot of every (query, neighbour cell) pair
  unsigned long long mbar;          // mbarrier of the bulk copies: THREADS arrivals + the bytes in flight per tile
  int n_list, n_pts, overflow, phase;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* mbar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* mbar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(mbar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(mbar)), "r"(parity)
      : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `mbar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

// Stages the neighbour cells of the tile's `cnt` queries (q[0 .. cnt): their FP32 p_world, filled and the block
// synchronised by the caller).  Whole block.  On return (block synchronised) st->overflow says whether the tile fits; if not, nothing of
// the staging may be used.
template <int NT>
__device__ __noinline__ void stage_cells(const MapView& map, StageSmem* st, const float4* q, int cnt,
                                            long long* dbg = nullptr) {
  const int tid = threadIdx.x;
  auto mark = [&](int tag) {  // LIO_TIMELINE instrumentation: last thread of block 0
    if (dbg != nullptr && blockIdx.x == 0 && tid == NT - 1) {
      const long long n = dbg[0];
      if (n < 62) {
        long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        dbg[1 + 2 * n] = tag;
        dbg[2 + 2 * n] = t;
        dbg[0] = n + 1;
      }
    }
  };
  mark(30);
  // 1. the cell set
#pragma unroll 1
  for (int p = tid; p < cnt * 27; p += NT) {
    const int qi = p / 27, c = p - qi * 27;
    const float4 qv = q[qi];
    const unsigned long long key = pack_cell(cell_coord(qv.x, map.inv_cell) + c % 3 - 1, cell_coord(qv.y, map.inv_cell) + (c / 3) % 3 - 1,
                                             cell_coord(qv.z, map.inv_cell) + c / 9 - 1);
    uint32_t h = hash64(key) & (ST_HASH - 1);
    for (;;) {
      const unsigned long long prev = atomicCAS(&st->key[h], LIO_EMPTY_KEY, key);
      if (prev == LIO_EMPTY_KEY) {
        st->list[atomicAdd(&st->n_list, 1)] = (uint16_t)h;
        break;
      }
      if (prev == key) break;
      h = (h + 1) & (ST_HASH - 1);
    }
    st->qslot[p] = (uint16_t)h;
  }
  __syncthreads();
  mark(31);
  // 2. + 3. one thread per distinct cell: probe the map, reserve room, start the copy
  uint32_t bytes = 0;
  const int n_list = st->n_list;
#pragma unroll 1
  for (int e = tid; e < n_list; e += NT) {
    const uint32_t h = st->list[e];
    uint32_t start, count;
    map_find(map, st->key[h], start, count);
    uint32_t info = 0;
    if (count > 0) {
      const uint32_t off = (uint32_t)atomicAdd(&st->n_pts, (int)count);
      if (off + count <= (uint32_t)ST_PTS && count < 4096u) {
        info = (off << 12) | count;
        bulk_g2s(&st->pts[off], map.pool + start, count * 16u, &st->mbar);
        bytes += count * 16u;
      } else {
        st->overflow = 1;
      }
    }
    st->cell[h] = info;
  }
  mark(32);
  if (bytes)
    mbar_arrive_expect_tx(&st->mbar, bytes);
  else
    mbar_arrive(&st->mbar);
  mbar_wait(&st->mbar, (uint32_t)st->phase & 1u);
  __syncthreads();
  mark(33);
  if (dbg != nullptr && blockIdx.x == 0 && tid == 0) {
    dbg[220] = st->n_list;
    dbg[221] = st->n_pts;
    dbg[222] = st->overflow;
  }
}

// A group of G lanes (G = 8, 16 or 32, aligned inside a warp) calls this with the SAME query; `gl` is the lane's index
// inside the group.  ALL 32 lanes of the warp must be here together (idle groups repeat a query): the shuffles below
// use the full mask.
// Round 0: the 27 cells of the 3x3x3 block around the query's cell are dealt round-robin to the lanes; a lane first
// issues the hash probes of all its cells (independent loads), then streams the buckets it found as ONE virtual list,
// four float4 loads in flight while the previous four are inserted into its private top-5.  The block holds every map
// point closer than `cell` to the query, so if the group has seen >= 5 points with d2 < cell^2 the result is already
// exact (the common case at one map point per 0.5 m voxel).  Otherwise further rounds visit, CPL cells per lane at a
// time through the same code, the remaining cells of the (2*rings+1)^3 block whose box distance is within the bound.
// The group then extracts the global top-5 with five rounds of redux.min on (d2 bits, id).
// Outputs are group-uniform: out_key[r] (d2 bits << 32 | id; ~0 when fewer than r+1 found), out_slot[r].
// With st != nullptr round 0 takes the candidates of the query's 27 cells from the tile's staging area (query qi of the
// tile) instead of global memory; the slots it reports then carry ST_FLAG and index st->pts.
template <int G>
__device__ __forceinline__ int group_knn5(const MapView& map, float qx, float qy, float qz, float max_d2, int rings,
                                          int gl, uint2* cells, unsigned long long out_key[LIO_K],
                                          uint32_t out_slot[LIO_K], long long* dbg = nullptr,
                                          StageSmem* st = nullptr, int qi = 0) {
  constexpr int CPL = (27 + G - 1) / G;  // cells per lane and round
  auto mark = [&](int slot) {  // LIO_TIMELINE instrumentation: block 0 / thread 0 only
    if (dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      dbg[slot] = t;
    }
  };
  mark(240);
  const int cx = cell_coord(qx, map.inv_cell), cy = cell_coord(qy, map.inv_cell), cz = cell_coord(qz, map.inv_cell);
  TopK top;
  top.init();
  const uint32_t max_bits = __float_as_uint(max_d2);
  const int side = 2 * rings + 1;
  const int ncell = side * side * side;
  int nrounds = 1;
  bool more = false;  // group-uniform: this query needs the rings beyond the 3x3x3 block
  uint32_t total = 0;
#pragma unroll 1
  for (int round = 0; round < nrounds; ++round) {
    if (round == 0 && st != nullptr) {
      // staged round 0: the lane's cells of the 3x3x3 block, candidates from shared memory
#pragma unroll 1
      for (int c = gl; c < 27; c += G) {
        const uint32_t info = st->cell[st->qslot[qi * 27 + c]];
        const uint32_t off = info >> 12, n = info & 0xfffu;
#pragma unroll 1
        for (uint32_t j = 0; j < n; ++j) {
          const float4 p = st->pts[off + j];
          const uint32_t d = __float_as_uint(dist2(qx, qy, qz, p.x, p.y, p.z));
          if (__float_as_int(p.w) >= 0 && d <= max_bits) top.insert(((unsigned long long)d << 32) | __float_as_uint(p.w), ST_FLAG | (off + j));
        }
      }
    } else {
    unsigned long long key[CPL];
    uint32_t h[CPL];
    uint4 e[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      key[u] = LIO_EMPTY_KEY;
      h[u] = 0;
      e[u] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
      int dx, dy, dz;
      bool take;
      if (round == 0) {
        const int c = gl + u * G;
        take = c < 27;
        dx = c % 3 - 1;
        dy = (c / 3) % 3 - 1;
        dz = c / 9 - 1;
      } else {
        const int c = (round - 1) * (G * CPL) + gl + u * G;
        dx = c % side - rings;
        dy = (c / side) % side - rings;
        dz = c / (side * side) - rings;
        take = more && c < ncell && !(dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1);
        if (take) {
          const float lx = (float)(cx + dx) * map.cell, ly = (float)(cy + dy) * map.cell,
                      lz = (float)(cz + dz) * map.cell;
          const float ex = fmaxf(fmaxf(lx - qx, qx - (lx + map.cell)), 0.f);
          const float ey = fmaxf(fmaxf(ly - qy, qy - (ly + map.cell)), 0.f);
          const float ez = fmaxf(fmaxf(lz - qz, qz - (lz + map.cell)), 0.f);
          // conservative (shrunk by 1e-3 relative) so rounding can never skip a cell that matters
          take = (ex * ex + ey * ey + ez * ez) * 0.999f <= max_d2;
        }
      }
      if (take) {
        key[u] = pack_cell(cx + dx, cy + dy, cz + dz);
        h[u] = hash64(key[u]) & map.hash_mask;
        e[u] = __ldg(reinterpret_cast<const uint4*>(map.table + h[u]));
      }
    }
    // resolve the probes
    uint32_t bst[CPL], bcn[CPL];
    uint32_t lsum = 0, lne = 0;  // the lane's candidates and non-empty cells
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
      uint32_t st = 0, cn = 0;
      if (key[u] != LIO_EMPTY_KEY) {
        for (;;) {
          const unsigned long long k = ((unsigned long long)e[u].y << 32) | e[u].x;
          if (k == key[u]) {
            st = e[u].z;
            cn = e[u].w;
            break;
          }
          if (k == LIO_EMPTY_KEY) break;
          h[u] = (h[u] + 1) & map.hash_mask;
          e[u] = __ldg(reinterpret_cast<const uint4*>(map.table + h[u]));
        }
      }
      bst[u] = st;
      bcn[u] = cn;
      lsum += cn;
      lne += cn > 0 ? 1u : 0u;
    }
    if (round == 0) mark(241);
    // The buckets of the GROUP form one virtual list [0, total) (lane-major, then the lane's cells): a scan over the
    // lanes gives every non-empty bucket its {end, base} (list index idx lives in pool slot base + idx of the first
    // bucket with idx < end), filed in shared memory.  Lane gl then takes the indices gl, gl + G, ...: every lane gets
    // the same number of candidates (+-1) whatever the cells hold -- dealt by CELLS, the lane that drew the full cells
    // of a wall set the pace of the whole tile with twice the mean (measured on the bench scans: 4.7 rounds of four
    // loads against 2.2) -- and the lanes of a group read consecutive points of a bucket in one request.
    {
      uint32_t isum = lsum, ine = lne;
#pragma unroll
      for (int off = 1; off < G; off <<= 1) {
        const uint32_t ts = __shfl_up_sync(0xffffffffu, isum, off, G);
        const uint32_t tn = __shfl_up_sync(0xffffffffu, ine, off, G);
        if (gl >= off) {
          isum += ts;
          ine += tn;
        }
      }
      total = __shfl_sync(0xffffffffu, isum, G - 1, G);
      uint32_t ex = isum - lsum, pos = ine - lne;
      if (round > 0) __syncwarp();  // the last round's list has been read
#pragma unroll
      for (int u = 0; u < CPL; ++u) {
        if (bcn[u] > 0) {
          cells[pos] = make_uint2(ex + bcn[u], bst[u] - ex);
          ++pos;
          ex += bcn[u];
        }
      }
      __syncwarp();
    }
    // stream the lane's share four points at a time, the next four already in flight while the current four are inserted
    const uint32_t nl = total > (uint32_t)gl ? (total - (uint32_t)gl + G - 1) / G : 0u;
    uint32_t cpos = 0;
    uint2 ce = cells[0];  // (unused when total == 0)
    auto slot_of = [&](uint32_t j) -> uint32_t {  // j-th candidate of the lane, j ascending over the calls
      const uint32_t idx = (uint32_t)gl + j * G;
      while (idx >= ce.x) ce = cells[++cpos];
      return ce.y + idx;
    };
    float4 cur[4], nxt[4];
    uint32_t cs[4], ns[4];
    if (nl > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cs[j] = slot_of(min((uint32_t)j, nl - 1));
        cur[j] = __ldg(map.pool + cs[j]);
      }
    }
#pragma unroll 1
    for (uint32_t t = 0; t < nl; t += 4) {
      if (t + 4 < nl) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ns[j] = slot_of(min(t + 4 + j, nl - 1));
          nxt[j] = __ldg(map.pool + ns[j]);
        }
      }
      // distance of the four points first (no selects), against the lane's current 5th best: most candidates are
      // farther and never reach the insertion code, of which there is ONE copy
      const uint32_t nj = min(4u, nl - t);
      const uint32_t thr = min(max_bits, (uint32_t)(top.k4 >> 32));
      uint32_t db[4];
      unsigned pass = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        db[j] = __float_as_uint(dist2(qx, qy, qz, cur[j].x, cur[j].y, cur[j].z));
        if ((uint32_t)j < nj && __float_as_int(cur[j].w) >= 0 && db[j] <= thr) pass |= 1u << j;
      }
#pragma unroll 1
      while (pass) {
        const int j = __ffs(pass) - 1;
        pass &= pass - 1;
        uint32_t d = db[0], id = __float_as_uint(cur[0].w), sl = cs[0];
        if (j == 1) { d = db[1]; id = __float_as_uint(cur[1].w); sl = cs[1]; }
        if (j == 2) { d = db[2]; id = __float_as_uint(cur[2].w); sl = cs[2]; }
        if (j == 3) { d = db[3]; id = __float_as_uint(cur[3].w); sl = cs[3]; }
        top.insert(((unsigned long long)d << 32) | id, sl);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cur[j] = nxt[j];
        cs[j] = ns[j];
      }
    }
    }
    if (round == 0) mark(242);
    if (round == 0 && rings > 1) {
      // exact already?  (0.999: any unseen point has true distance > cell, so its rounded d2 exceeds this bound)
      const uint32_t near_bits = __float_as_uint(fminf(map.cell * map.cell * 0.999f, max_d2));
      const int mine = (int)((uint32_t)(top.k0 >> 32) <= near_bits) + (int)((uint32_t)(top.k1 >> 32) <= near_bits) +
                       (int)((uint32_t)(top.k2 >> 32) <= near_bits) + (int)((uint32_t)(top.k3 >> 32) <= near_bits) +
                       (int)((uint32_t)(top.k4 >> 32) <= near_bits);
      int seen = mine;
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, off);
      // (the rounds hold warp-wide shuffles: the whole warp walks through them, the groups that are done with no cell)
      more = seen < LIO_K;
      if (__any_sync(0xffffffffu, more)) nrounds = 1 + (ncell + G * CPL - 1) / (G * CPL);
    }
  }
  mark(243);
  // Global top-5 of the group: five rounds of a butterfly minimum over (key, slot).  xor offsets below G never leave
  // the aligned group, so the shuffles use the FULL mask (the whole warp is here: no per-group masks, which cost a
  // warp-level rendezvous each).  Keys are unique (ids are), so exactly one lane pops per round.
  __syncwarp();
  int found = 0;
#pragma unroll
  for (int r = 0; r < LIO_K; ++r) {
    unsigned long long k = top.k0;
    uint32_t sl = top.s0;
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, off);
      const uint32_t os = __shfl_xor_sync(0xffffffffu, sl, off);
      if (ok < k) {
        k = ok;
        sl = os;
      }
    }
    out_key[r] = k;
    out_slot[r] = sl;
    if (k != ~0ull) {
      if (top.k0 == k) top.pop();
      ++found;
    }
  }
  mark(244);
  return found;
}

// ---------------------------------------------------------------------------------------------------------
// UNBOUNDED k-nearest search of one query by one warp: what KD_TREE::Nearest_Search returns with its default
// max_dist = INFINITY (ikd_Tree.h:285), the way esekfom.hpp:140-141 calls it.  The hot search above is bounded at
// d2 <= 5 (all the update's gate needs, esekfom.hpp:144-147); this one serves the rows it left incomplete -- map-frontier
// points whose neighbours lie farther away, which map_incremental still reads (laserMapping.cpp:391-423).
// Shells of cells at Chebyshev distance r = 0, 1, 2, ... around the query's cell, dealt to the 32 lanes and clipped to
// the bounding box of the cells the map has ever used (MapView::counters[8..13]).  After shell r every unseen point is
// farther than r * cell, so the search stops once the `need`-th best is closer than that, or the box is exhausted.
// `need` = 1 ... 5 neighbours wanted (the stop rule only; the lanes always keep five).  Outputs as group_knn5.
// To be called by blocks of ONE warp (it synchronises with __syncthreads).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, off);
    if (o < k) k = o;
  }
  return k;
}
// d2 bits of the need-th smallest key held by the warp's lanes (0xFFFFFFFF when there are fewer); `top` is not changed
__device__ __forceinline__ uint32_t warp_kth_bits(TopK t, int need) {
  unsigned long long k = ~0ull;
#pragma unroll 1
  for (int r = 0; r < need; ++r) {
    k = warp_min_u64(t.k0);
    if (k == ~0ull) break;
    if (t.k0 == k) t.pop();
  }
  return (uint32_t)(k >> 32);
}

__device__ __noinline__ int warp_knn_far(const MapView& map, float qx, float qy, float qz, int need,
                                         unsigned long long out_key[LIO_K], uint32_t out_slot[LIO_K]) {
  const int lane = threadIdx.x & 31;
  const int c[3] = {cell_coord(qx, map.inv_cell), cell_coord(qy, map.inv_cell), cell_coord(qz, map.inv_cell)};
  int lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = (int)__ldcg(map.counters + 8 + a);
    hi[a] = (int)__ldcg(map.counters + 11 + a);
  }
  TopK top;
  top.init();
  int r0 = 0, rmax = -1;  // first shell that touches the box, last shell that does
  if (lo[0] <= hi[0]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      r0 = max(r0, max(lo[a] - c[a], c[a] - hi[a]));
      rmax = max(rmax, max(abs(c[a] - lo[a]), abs(c[a] - hi[a])));
    }
  }
  uint32_t kth = 0xFFFFFFFFu;  // d2 bits of the need-th best after the last completed shell
  if (r0 > 256) {
    // the query is hundreds of cells away from everything: walking empty shells towards the box would cost more than
    // the box itself, so its cells are simply all visited (no LiDAR return is that far from its map; this keeps the
    // call exact and bounded whatever it is handed)
    const long long nx = (long long)hi[0] - lo[0] + 1, ny = (long long)hi[1] - lo[1] + 1, nz = (long long)hi[2] - lo[2] + 1;
#pragma unroll 1
    for (long long t = lane; t < nx * ny * nz; t += 32) {
      const int x = lo[0] + (int)(t % nx), y = lo[1] + (int)((t / nx) % ny), z = lo[2] + (int)(t / (nx * ny));
      uint32_t start, count;
      if (map_find(map, pack_cell(x, y, z), start, count) < 0) continue;
#pragma unroll 1
      for (uint32_t j = 0; j < count; ++j) {
        const float4 p = __ldg(map.pool + start + j);
        if (__float_as_int(p.w) < 0) continue;
        const uint32_t d = __float_as_uint(dist2(qx, qy, qz, p.x, p.y, p.z));
        top.insert(((unsigned long long)d << 32) | __float_as_uint(p.w), start + j);
      }
    }
    rmax = -1;  // skip the shells
  }
#pragma unroll 1
  for (int r = r0; r <= rmax; ++r) {
    if (r > r0 && kth != 0xFFFFFFFFu) {
      // every cell of shell r and beyond is farther than (r - 1) * cell from the query's cell, hence from the query.
      // 0.999: cell_coord rounds x * inv_cell, so a point can sit one ulp across the border of the cell it is filed in
      const float cover = (float)(r - 1) * map.cell * 0.999f;
      if (__uint_as_float(kth) < cover * cover) break;
    }
    const int side = 2 * r + 1;
    const int n_face = side * side;
    const int n_cells = r == 0 ? 1 : 2 * n_face + (side - 2) * 8 * r;
    // four cells per lane at a time: their first probes are issued together (the hash table is a dependent load each,
    // and a shell of radius r has ~24 r^2 cells)
#pragma unroll 1
    for (int t0 = lane; t0 < n_cells; t0 += 4 * 32) {
      unsigned long long ckey[4];
      uint32_t hh[4];
      uint4 ee[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + 32 * u;
        ckey[u] = LIO_EMPTY_KEY;
        hh[u] = 0;
        ee[u] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
        if (t >= n_cells) continue;
        int dx, dy, dz;
        if (t < 2 * n_face || r == 0) {  // the two full faces dz = -r, +r
          const int f = t >= n_face ? 1 : 0;
          const int e = t - f * n_face;
          dz = f ? r : -r;
          dy = e / side - r;
          dx = e % side - r;
        } else {  // the perimeter of the layers in between: 8r cells each
          const int e = t - 2 * n_face;
          const int layer = e / (8 * r), o = e % (8 * r);
          dz = -r + 1 + layer;
          const int sd = o / (2 * r), uu = o % (2 * r);
          dx = sd == 0 ? -r + uu : (sd == 1 ? r : (sd == 2 ? r - uu : -r));
          dy = sd == 0 ? -r : (sd == 1 ? -r + uu : (sd == 2 ? r : r - uu));
        }
        const int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
        if (x < lo[0] || x > hi[0] || y < lo[1] || y > hi[1] || z < lo[2] || z > hi[2]) continue;
        if (kth != 0xFFFFFFFFu) {  // cell farther than the need-th best already: nothing in it can matter
          const float lx = (float)x * map.cell, ly = (float)y * map.cell, lz = (float)z * map.cell;
          const float ex = fmaxf(fmaxf(lx - qx, qx - (lx + map.cell)), 0.f);
          const float ey = fmaxf(fmaxf(ly - qy, qy - (ly + map.cell)), 0.f);
          const float ez = fmaxf(fmaxf(lz - qz, qz - (lz + map.cell)), 0.f);
          if ((ex * ex + ey * ey + ez * ez) * 0.998f > __uint_as_float(kth)) continue;
        }
        ckey[u] = pack_cell(x, y, z);
        hh[u] = hash64(ckey[u]) & map.hash_mask;
        ee[u] = __ldg(reinterpret_cast<const uint4*>(map.table + hh[u]));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ckey[u] == LIO_EMPTY_KEY) continue;
        uint32_t start = 0, count = 0;
        for (;;) {
          const unsigned long long k = ((unsigned long long)ee[u].y << 32) | ee[u].x;
          if (k == ckey[u]) {
            start = ee[u].z;
            count = ee[u].w;
            break;
          }
          if (k == LIO_EMPTY_KEY) break;
          hh[u] = (hh[u] + 1) & map.hash_mask;
          ee[u] = __ldg(reinterpret_cast<const uint4*>(map.table + hh[u]));
        }
#pragma unroll 1
        for (uint32_t j = 0; j < count; ++j) {
          const float4 p = __ldg(map.pool + start + j);
          if (__float_as_int(p.w) < 0) continue;
          const uint32_t d = __float_as_uint(dist2(qx, qy, qz, p.x, p.y, p.z));
          top.insert(((unsigned long long)d << 32) | __float_as_uint(p.w), start + j);
        }
      }
    }
    __syncthreads();  // the block IS this warp: the barrier brings its lanes back together before the shuffles (a warp
                      // left split by the loops above pays ~1,000 cycles per shuffle; __syncwarp does not mend that)
    kth = warp_kth_bits(top, need);
  }
  __syncthreads();
  int found = 0;
#pragma unroll
  for (int r = 0; r < LIO_K; ++r) {
    unsigned long long k = top.k0;
    uint32_t sl = top.s0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, off);
      const uint32_t os = __shfl_xor_sync(0xffffffffu, sl, off);
      if (ok < k) {
        k = ok;
        sl = os;
      }
    }
    out_key[r] = k;
    out_slot[r] = sl;
    if (k != ~0ull) {
      if (top.k0 == k) top.pop();
      ++found;
    }
  }
  return found;
}

// ---------------------------------------------------------------------------------------------------------
// esti_plane<float> (common_lib.h:102-134): 5x3 FP32 column-pivoted Householder QR solve of A n = -1 in the
// operation order of SURVEY.md App. B.1 (Eigen ColPivHouseholderQR recipe), one IEEE rounding per operation.
// P[j] = neighbour j (x,y,z).  Returns true when all 5 neighbours lie within `thr` of the fitted plane.
// ---------------------------------------------------------------------------------------------------------
#define LIO_SWAPF(a, b) \
  {                     \
    float _t = (a);     \
    (a) = (b);          \
    (b) = _t;           \
  }

__device__ __noinline__ bool esti_plane(const float4 P[5], float thr, float pabcd[4]) {
  float A[5][3];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    A[i][0] = P[i].x;
    A[i][1] = P[i].y;
    A[i][2] = P[i].z;
  }
  float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
  float nd[3], nu[3], tau[3];
  int trans[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = s + A[i][j] * A[i][j];
    nd[j] = nu[j] = sqrtf(s);
  }
  const float maxn = fmaxf(nu[0], fmaxf(nu[1], nu[2]));
  const float eps = FLT_EPSILON;
  const float thr_helper = ((maxn * eps) * (maxn * eps)) / 5.0f;
  const float downdate_thr = sqrtf(eps);
  int nonzero = 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int jb = k;
    float best = nu[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j)
      if (nu[j] > best) {
        best = nu[j];
        jb = j;
      }
    if (nonzero == 3 && best * best < thr_helper * (float)(5 - k)) nonzero = k;
    trans[k] = jb;
#pragma unroll
    for (int j = k + 1; j < 3; ++j)
      if (jb == j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) LIO_SWAPF(A[i][k], A[i][j]);
        LIO_SWAPF(nu[k], nu[j]);
        LIO_SWAPF(nd[k], nd[j]);
      }
    float tailsq = 0.f;
#pragma unroll
    for (int i = k + 1; i < 5; ++i) tailsq = tailsq + A[i][k] * A[i][k];
    const float c0 = A[k][k];
    float beta, t;
    float ess[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (tailsq <= FLT_MIN) {
      t = 0.f;
      beta = c0;
    } else {
      beta = sqrtf(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) ess[i] = A[i][k] / den;
      t = (beta - c0) / beta;
    }
    tau[k] = t;
    A[k][k] = beta;
#pragma unroll
    for (int i = k + 1; i < 5; ++i) A[i][k] = ess[i];
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      float tmp = 0.f;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) tmp = tmp + ess[i] * A[i][j];
      tmp = tmp + A[k][j];
      A[k][j] = A[k][j] - t * tmp;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) A[i][j] = A[i][j] - (t * tmp) * ess[i];
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (nu[j] != 0.f) {
        float tt = fabsf(A[k][j]) / nu[j];
        tt = (1.f + tt) * (1.f - tt);
        if (tt < 0.f) tt = 0.f;
        const float ratio = nu[j] / nd[j];
        const float t2 = tt * (ratio * ratio);
        if (t2 <= downdate_thr) {
          float s = 0.f;
#pragma unroll
          for (int i = k + 1; i < 5; ++i) s = s + A[i][j] * A[i][j];
          nd[j] = sqrtf(s);
          nu[j] = nd[j];
        } else {
          nu[j] = nu[j] * sqrtf(tt);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < nonzero) {
      float tmp = 0.f;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) tmp = tmp + A[i][k] * c[i];
      tmp = tmp + c[k];
      c[k] = c[k] - tau[k] * tmp;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) c[i] = c[i] - (tau[k] * tmp) * A[i][k];
    }
  }
  float y[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < nonzero) {
      float s = c[i];
#pragma unroll
      for (int j = i + 1; j < 3; ++j)
        if (j < nonzero) s = s - A[i][j] * y[j];
      y[i] = s / A[i][i];
    }
  }
  // un-permute (column transpositions in reverse; trans[2] can only be 2)
  if (trans[1] == 2) LIO_SWAPF(y[1], y[2]);
  if (trans[0] == 1) LIO_SWAPF(y[0], y[1]);
  if (trans[0] == 2) LIO_SWAPF(y[0], y[2]);
  const float n = sqrtf((y[0] * y[0] + y[1] * y[1]) + y[2] * y[2]);
  pabcd[0] = y[0] / n;
  pabcd[1] = y[1] / n;
  pabcd[2] = y[2] / n;
  pabcd[3] = (float)(1.0 / (double)n);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float d = ((pabcd[0] * P[j].x + pabcd[1] * P[j].y) + pabcd[2] * P[j].z) + pabcd[3];
    if (fabsf(d) > thr) ok = false;
  }
  return ok;
}

}  // namespace lio
