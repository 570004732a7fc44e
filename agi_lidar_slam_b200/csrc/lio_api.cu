// extern "C" surface of liblio_b200.so (include/lio_b200.h): context, staging copies, graph capture of the whole
// update, and the small host-side sequential pieces (predict / boxplus / boxminus).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "lio_ctx.cuh"

namespace lio {
size_t preprocess_sort_bytes(int64_t n);
int preprocess_init_counters(lio_ctx* c);

__global__ void set_w_kernel(float4* dst, const float* w, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i].w = w[i];
}
__global__ void set_scan_m_kernel(int* d, int m) { *d = m; }
__global__ void xyz_to_float4_kernel(const float* xyz, float4* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
}

static uint32_t pow2_at_least(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return (uint32_t)p;
}

// Copy n point records from host memory into a device float4 array: x,y,z and w = the float at byte offset w_off
// of each record (stride 16: w_off is 12).  aux (optional) receives the float at aux_off.
static int stage_points(lio_ctx* c, const void* src, int64_t n, int stride, int w_off, float4* dst, int aux_off,
                        float* aux, float* tmp_w) {
  if (n <= 0) return LIO_OK;
  const char* s = static_cast<const char*>(src);
  if (stride == 16) {
    LIO_CHECK(c, cudaMemcpyAsync(dst, s, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return LIO_OK;
  }
  LIO_CHECK(c, cudaMemcpy2DAsync(dst, 16, s, stride, 16, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  LIO_CHECK(c, cudaMemcpy2DAsync(tmp_w, 4, s + w_off, stride, 4, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  set_w_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(dst, tmp_w, (int)n);
  c->launches++;
  if (aux) LIO_CHECK(c, cudaMemcpy2DAsync(aux, 4, s + aux_off, stride, 4, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  return LIO_OK;
}

static int refresh_scan_m(lio_ctx* c) {
  if (c->scan_m >= 0) return LIO_OK;
  int m = 0;
  LIO_CHECK(c, cudaMemcpyAsync(&m, c->d_scan_m, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  c->scan_m = m;
  return LIO_OK;
}

}  // namespace lio

using namespace lio;

// peer mailbox of the sharded-map exchange: blobs [8 ranks][2 slots][LIO_BLOB] doubles, stamps [8][2] u32, error flag
static const size_t MAILBOX_ERR = 8 * 2 * LIO_BLOB * 2 * sizeof(unsigned long long);  // behind the stamped words
static const size_t MAILBOX_BYTES = 24576;

extern "C" {

int lio_abi_version(void) { return LIO_ABI_VERSION; }

void lio_default_caps(lio_caps* caps) {
  caps->max_scan_points = 262144;
  caps->max_down_points = 100000;  // esekfom.hpp:23-29
  caps->max_map_points = 4194304;
  caps->map_cell = 1.5f;  // 3x3x3 block covers 1.5 m: the 5th neighbour is almost always inside
  caps->knn_max_d2 = 5.0f;      // esekfom.hpp:147
  caps->plane_thr = 0.1f;       // esekfom.hpp:157
  caps->map_downsample = 0.5f;  // filter_size_map (launch files)
}

const char* lio_last_error(lio_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t lio_launch_count(lio_ctx* ctx) { return ctx ? ctx->launches : 0; }

#define ALLOC(ptr, bytes) LIO_CHECK(c, cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes)))

static int create_impl(lio_ctx* c) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= c->device) {
    c->err = "no usable CUDA device (this library has no CPU fallback)";
    return LIO_E_NO_DEVICE;
  }
  LIO_CHECK(c, cudaSetDevice(c->device));
  cudaDeviceProp prop;
  LIO_CHECK(c, cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) {
    c->err = "device is not sm_100-class (kernels are built for sm_100a only)";
    return LIO_E_NO_DEVICE;
  }
  c->sm_count = prop.multiProcessorCount;
  LIO_CHECK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  const lio_caps& k = c->caps;
  // map
  c->hash_cap = pow2_at_least((uint64_t)k.max_map_points * 2);
  c->map.hash_mask = c->hash_cap - 1;
  c->map.pool_cap = (uint32_t)std::min<uint64_t>((uint64_t)k.max_map_points * 4, 0xFFFFFFF0ull);
  c->map.cell = k.map_cell;
  c->map.inv_cell = 1.0f / k.map_cell;
  c->knn_rings = (int)ceil(sqrt((double)k.knn_max_d2) / (double)k.map_cell);
  if (c->knn_rings < 1) c->knn_rings = 1;
  ALLOC(c->map.table, sizeof(CellEntry) * (size_t)c->hash_cap);
  ALLOC(c->map.cell_cap, 4 * (size_t)c->hash_cap);
  ALLOC(c->map.cell_pend, 4 * (size_t)c->hash_cap);
  ALLOC(c->map.cell_base, 4 * (size_t)c->hash_cap);
  ALLOC(c->map.pool, sizeof(float4) * (size_t)c->map.pool_cap);
  ALLOC(c->map.counters, 4 * 16);
  c->map.removed_cap = (uint32_t)std::min<uint64_t>((uint64_t)k.max_map_points, 1ull << 22);
  ALLOC(c->map.removed, sizeof(float4) * (size_t)c->map.removed_cap);
  c->map_downsample = k.map_downsample;
  c->batch_cap = std::max<int64_t>(k.max_map_points, k.max_down_points);
  ALLOC(c->d_batch_pts, sizeof(float4) * (size_t)c->batch_cap);
  ALLOC(c->d_batch_slot, 4 * (size_t)c->batch_cap);
  ALLOC(c->d_batch_rank, 4 * (size_t)c->batch_cap);
  ALLOC(c->d_batch_flag, (size_t)c->batch_cap);
  c->vox_cap = pow2_at_least((uint64_t)std::max<int64_t>(k.max_down_points, 131072) * 2);
  ALLOC(c->d_vox_best, 8 * (size_t)c->vox_cap);
  ALLOC(c->d_vox_key, 8 * (size_t)c->vox_cap);

  // scan
  const size_t M = (size_t)k.max_down_points;
  ALLOC(c->d_body, sizeof(float4) * M);
  ALLOC(c->d_world, sizeof(float4) * M);
  ALLOC(c->d_near, sizeof(float4) * M * LIO_K);
  ALLOC(c->d_near_d2, 4 * M * LIO_K);
  ALLOC(c->d_near_cnt, 4 * M);
  ALLOC(c->d_near_q, sizeof(float4) * M);
  ALLOC(c->d_selected, M);
  ALLOC(c->d_normvec, sizeof(float4) * M);
  ALLOC(c->d_plane, sizeof(float4) * M);
  ALLOC(c->d_partials, 8 * (size_t)(2 * LIO_BLOB) * pass_grid_blocks(c));
  LIO_CHECK(c, cudaMemset(c->d_partials, 0, 8 * (size_t)(2 * LIO_BLOB) * pass_grid_blocks(c)));
  ALLOC(c->d_blob_own, 8 * LIO_BLOB);
  c->d_blob = c->d_blob_own;
  ALLOC(c->d_prior, 8 * 288);
  ALLOC(c->d_pub, 8 * 1024);  // PUB_COPIES x PUB_STRIDE stamped words (lio_pass.cu)
  LIO_CHECK(c, cudaMemset(c->d_pub, 0, 8 * 1024));
  ALLOC(c->d_mailbox, MAILBOX_BYTES);
  LIO_CHECK(c, cudaMemset(c->d_mailbox, 0, MAILBOX_BYTES));
  {
    // LIO_ZERO_COPY=1: pass 0 reads a pinned scan in place over PCIe instead of a copy first (measured equal on B200)
    const char* env = getenv("LIO_ZERO_COPY");
    c->no_zero_copy = !(env && atoi(env));
  }
  {
    const char* env = getenv("LIO_STAGE_SEARCH");
    c->stage_search = env && atoi(env);
    const char* il = getenv("LIO_INTERLEAVE");
    c->interleave = il && atoi(il);
    const char* bf = getenv("LIO_FINISH_BATCHED");
    c->batched_finish = !(bf && !atoi(bf));
  }
  {
    const char* env = getenv("LIO_TIMELINE");
    if (env && atoi(env)) {
      ALLOC(c->d_dbg, 1024 * sizeof(long long));
      LIO_CHECK(c, cudaMemset(c->d_dbg, 0, 1024 * sizeof(long long)));
    }
  }
  ALLOC(c->d_cls, M);
  ALLOC(c->d_add_a, sizeof(float4) * M);
  LIO_CHECK(c, cudaMemset(c->d_selected, 0, M));
  LIO_CHECK(c, cudaMemset(c->d_near_cnt, 0, 4 * M));
  LIO_CHECK(c, cudaMemset(c->d_blob, 0, 8 * LIO_BLOB));

  // filter state: one block so that the prior goes up and the posterior comes down in one copy each
  {
    // {x 26, P 576, ctrl 4} {x0 26, P0 576, scan_m 1, sync 1} {xprop 26} {dx 24}: the second group is what one update
    // needs from the host, so lio_update_scan_host sends it in ONE copy (prior + scan size + zeroed barrier words)
    const size_t nd = 606 + 604 + 26 + 24 + 6;
    ALLOC(c->d_state_blk, 8 * nd);
    LIO_CHECK(c, cudaMemset(c->d_state_blk, 0, 8 * nd));
    double* b = c->d_state_blk;
    c->d_x = reinterpret_cast<StateD*>(b);
    c->d_P = b + 26;
    c->d_ctrl = reinterpret_cast<Ctrl*>(b + 602);
    c->d_x0 = reinterpret_cast<StateD*>(b + 606);
    c->d_P0 = b + 606 + 26;
    c->d_scan_m = reinterpret_cast<int*>(b + 606 + 602);
    c->d_sync = reinterpret_cast<unsigned*>(b + 606 + 603);
    c->d_xprop = reinterpret_cast<StateD*>(b + 1210);
    c->d_dx = b + 1236;
  }
  c->h_pinned_bytes = 8 * (1280 + LIO_BLOB + 4);
  LIO_CHECK(c, cudaMallocHost(&c->h_pinned, c->h_pinned_bytes));
  LIO_CHECK(c, cudaEventCreateWithFlags(&c->upload_done, cudaEventDisableTiming));
  LIO_CHECK(c, cudaEventCreateWithFlags(&c->multi_evt, cudaEventDisableTiming));
  LIO_CHECK(c, cudaEventCreateWithFlags(&c->ev_post, cudaEventDisableTiming));
  LIO_CHECK(c, cudaEventCreateWithFlags(&c->ev_growth, cudaEventDisableTiming));
  LIO_CHECK(c, cudaEventCreateWithFlags(&c->ev_prep, cudaEventDisableTiming));
  LIO_CHECK(c, cudaStreamCreateWithFlags(&c->prep_stream, cudaStreamNonBlocking));
  LIO_CHECK(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_out), 8 * 616, cudaHostAllocMapped));
  memset(c->h_out, 0, 8 * 616);
  LIO_CHECK(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->h_out_dev), c->h_out, 0));

  // preprocess
  const size_t N = (size_t)k.max_scan_points;
  ALLOC(c->d_raw, sizeof(float4) * N);
  ALLOC(c->d_raw_aux, 4 * N);
  ALLOC(c->d_undist, sizeof(float4) * N);
  ALLOC(c->d_vkeys, 12 * N);
  ALLOC(c->d_poses, sizeof(lio_pose6d) * 128);
  ALLOC(c->d_sort_keys_in, 4 * N);
  ALLOC(c->d_sort_keys_out, 4 * N);
  ALLOC(c->d_sort_vals_in, 4 * N);
  ALLOC(c->d_sort_vals_out, 4 * N);
  c->cub_tmp_bytes = preprocess_sort_bytes((int64_t)N);
  ALLOC(c->d_cub_tmp, c->cub_tmp_bytes + 16);
  ALLOC(c->d_prep_counters, 4 * 32);
  LIO_CHECK(c, cudaMemset(c->d_prep_counters, 0, 4 * 32));
  {
    // surf voxel filter: leaf hash (>= 2 slots per leaf the scan may have), point slots / segments, grid bitmap
    lio::VoxelFilter& v = c->vf;
    const uint32_t hs = pow2_at_least((uint64_t)k.max_down_points * 2);
    v.hash_mask = hs - 1;
    v.bitmap_bits = 1ull << 28;  // leaves of the scan's bounding grid (1024 x 1024 x 256 at most; PCL's own limit is 2^31)
    ALLOC(v.key, 8 * (size_t)hs);
    LIO_CHECK(c, cudaMemset(v.key, 0xFF, 8 * (size_t)hs));
    uint32_t** z4[] = {&v.cnt, &v.off, &v.rank, &v.lin, &v.list};
    for (uint32_t** pp : z4) {
      ALLOC(*pp, 4 * (size_t)hs);
      LIO_CHECK(c, cudaMemset(*pp, 0, 4 * (size_t)hs));
    }
    ALLOC(v.big, 16 * (N / 128 + 2));
    ALLOC(v.mid, 16 * (N / 16 + 2));
    ALLOC(v.slot, 4 * N);
    ALLOC(v.pos, 4 * N);
    ALLOC(v.seg, 4 * N);
    ALLOC(v.seg2, 4 * N);
    ALLOC(v.bitmap, (size_t)(v.bitmap_bits / 8));
    LIO_CHECK(c, cudaMemset(v.bitmap, 0, (size_t)(v.bitmap_bits / 8)));
    const size_t n_sb = (size_t)(v.bitmap_bits >> 10) + 2048;  // padded: the scan reads whole 16-byte groups per thread
    ALLOC(v.sbcount, 4 * n_sb);
    LIO_CHECK(c, cudaMemset(v.sbcount, 0, 4 * n_sb));
    ALLOC(v.sbprefix, 4 * n_sb);
    ALLOC(v.ctr, 4 * 8);
    LIO_CHECK(c, cudaMemset(v.ctr, 0, 4 * 8));
  }

  int rc = ensure_tables(c);
  if (rc) return rc;
  rc = map_reset(c);
  if (rc) return rc;
  rc = preprocess_init_counters(c);
  if (rc) return rc;
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

int lio_create(int device, const lio_caps* caps, lio_ctx** out) {
  if (!out) return LIO_E_INVALID;
  *out = nullptr;
  lio_ctx* c = new lio_ctx();
  c->device = device;
  if (caps)
    c->caps = *caps;
  else
    lio_default_caps(&c->caps);
  if (c->caps.max_scan_points <= 0 || c->caps.max_down_points <= 0 || c->caps.max_map_points <= 0 ||
      !(c->caps.map_cell > 0.f) || !(c->caps.knn_max_d2 > 0.f) || !(c->caps.map_downsample > 0.f)) {
    delete c;
    return LIO_E_INVALID;
  }
  const int rc = create_impl(c);
  if (rc != LIO_OK) {
    fprintf(stderr, "lio_create failed: %s\n", c->err.c_str());
    lio_destroy(c);
    return rc;
  }
  *out = c;
  return LIO_OK;
}

void lio_destroy(lio_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->prep_stream) cudaStreamSynchronize(c->prep_stream);  // a prefetched scan may still be on its way
  for (int r = 0; r < 8; ++r)
    if (c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  void* ptrs[] = {c->map.removed, c->map.table,   c->map.cell_cap,  c->map.cell_pend, c->map.cell_base, c->map.pool,
                  c->map.counters, c->d_batch_pts,  c->d_batch_slot,  c->d_batch_rank,  c->d_batch_flag,
                  c->d_vox_best,  c->d_vox_key,     c->d_body,        c->d_world,
                  c->d_near,      c->d_near_d2,     c->d_near_cnt,    c->d_selected,    c->d_normvec,     c->d_plane,
                  c->d_near_q,
                  c->d_partials,  c->d_blob_own,    c->d_cls,         c->d_add_a,
                  c->d_state_blk, c->d_prior,       c->d_dbg,         c->d_pub,         c->d_mailbox,
                  c->d_cloud,
                  c->d_raw,       c->d_raw_aux,     c->d_undist,
                  c->d_vkeys,     c->d_poses,
                  c->d_sort_keys_in, c->d_sort_keys_out, c->d_sort_vals_in, c->d_sort_vals_out, c->d_cub_tmp,
                  c->d_prep_counters, c->vf.key,    c->vf.cnt,        c->vf.mid,        c->vf.off,        c->vf.rank,
                  c->vf.lin,      c->vf.list,       c->vf.big,        c->vf.slot,       c->vf.seg,        c->vf.bitmap,
                  c->vf.sbcount,  c->vf.sbprefix,   c->vf.ctr,        c->vf.pos,        c->vf.seg2};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->upload_done) cudaEventDestroy(c->upload_done);
  if (c->multi_evt) cudaEventDestroy(c->multi_evt);
  if (c->ev_post) cudaEventDestroy(c->ev_post);
  if (c->ev_growth) cudaEventDestroy(c->ev_growth);
  if (c->ev_prep) cudaEventDestroy(c->ev_prep);
  if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
  delete c;
}

int lio_set_stream(lio_ctx* c, void* cuda_stream) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->prep_stream));
  const int rs = settle_growth(c);  // nothing of the old stream is left pending
  c->staged_ptr = nullptr;
  c->staged_on_prep = false;
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  c->own_stream = false;
  return rs;
}

int lio_synchronize(lio_ctx* c) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return settle_growth(c);
}

// ---------------------------------------------------------------- deferred map growth
// The reference publishes the odometry of a scan before map_incremental runs (laserMapping.cpp:776-785).  With
// lio_set_deferred_growth the step does the same: lio_scan_step_finish hands the posterior back as soon as it is on
// the host while the map growth of that scan is still running on the stream, under the host stage (IMU propagation)
// and the upload of the next scan.  The counts and errors of the growth are booked by the next call that needs the
// host's view of the map (settle_growth), at the latest by the growth of the next scan.
}  // extern "C"
namespace lio {
// Point ids are int32 and a negative id marks a dead slot: the id space must not wrap.  Every candidate of every insert
// takes an id (the rejected ones too), so a long run can use it up (~2^31 candidates: tens of hours at 10 Hz).
int check_id_space(lio_ctx* c, int64_t n_new) {
  if ((int64_t)c->next_id + n_new > (int64_t)INT32_MAX - 1) {
    c->err = "map point ids exhausted (2^31 insert candidates): rebuild the map from lio_map_dump to renumber";
    return LIO_E_CAPACITY;
  }
  return LIO_OK;
}
int settle_growth(lio_ctx* c) {
  if (!c->growth_pending) return LIO_OK;
  c->growth_pending = false;
  LIO_CHECK(c, cudaEventSynchronize(c->ev_growth));
  const double* hp = static_cast<const double*>(c->h_pinned);
  const int* cls = reinterpret_cast<const int*>(hp + 620);
  const uint32_t* mc = reinterpret_cast<const uint32_t*>(hp + 622);
  c->last_counts[0] = cls[0];
  c->last_counts[1] = cls[1];
  c->last_counts[2] = (int32_t)mc[4];
  c->next_id += cls[0] + cls[1];
  if (mc[3] != 0) {
    c->err = mc[3] == 1 ? "map hash table full (raise lio_caps.max_map_points)"
                        : "map point pool full (raise lio_caps.max_map_points)";
    return LIO_E_CAPACITY;
  }
  return LIO_OK;
}
}  // namespace lio
extern "C" {

int lio_set_deferred_growth(lio_ctx* c, int on) {
  if (!c) return LIO_E_INVALID;
  const int rc = settle_growth(c);
  c->deferred_growth = on != 0;
  return rc;
}

int lio_scan_step_settle(lio_ctx* c, int32_t counts[3]) {
  if (!c) return LIO_E_INVALID;
  const int rc = settle_growth(c);
  if (counts) memcpy(counts, c->last_counts, sizeof(c->last_counts));
  return rc;
}

// ---------------------------------------------------------------- map
int lio_map_build(lio_ctx* c, const void* pts, int64_t n, int stride) {
  if (!c || n < 0 || (stride != 16 && stride != 48) || (n > 0 && !pts)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  if (n > c->caps.max_map_points) {
    c->err = "Build: more points than lio_caps.max_map_points";
    return LIO_E_CAPACITY;
  }
  int rc = map_reset(c);
  if (rc) return rc;
  if (n == 0) return LIO_OK;  // KD_TREE::Build returns before creating a root (ikd_Tree.cpp:359)
  rc = stage_points(c, pts, n, stride, 32, c->d_batch_pts, 0, nullptr, reinterpret_cast<float*>(c->d_batch_slot));
  if (rc) return rc;
  rc = map_append_batch(c, c->d_batch_pts, n, 0, nullptr);
  if (rc) return rc;
  c->next_id = (int32_t)n;
  c->map_built = true;
  return LIO_OK;
}

int lio_map_set_downsample(lio_ctx* c, float downsample_size) {
  if (!c || !(downsample_size > 0.f)) return LIO_E_INVALID;
  c->map_downsample = downsample_size;
  return LIO_OK;
}

int lio_map_add(lio_ctx* c, const void* pts, int64_t n, int stride, int downsample_on, int32_t* n_added) {
  if (!c || n < 0 || (stride != 16 && stride != 48) || (n > 0 && !pts)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  if (n_added) *n_added = 0;
  if (!c->map_built) {
    c->err = "Add_Points on an empty map (the reference dereferences Root_Node here)";
    return LIO_E_EMPTY_MAP;
  }
  if (n == 0) return LIO_OK;
  if (n > c->batch_cap) {
    c->err = "Add_Points batch too large";
    return LIO_E_CAPACITY;
  }
  if (const int ri = check_id_space(c, n)) return ri;
  int rc = stage_points(c, pts, n, stride, 32, c->d_batch_pts, 0, nullptr, reinterpret_cast<float*>(c->d_batch_slot));
  if (rc) return rc;
  if (downsample_on) {
    rc = map_add_downsample(c, c->d_batch_pts, n, c->next_id, n_added);
  } else {
    rc = map_append_batch(c, c->d_batch_pts, n, c->next_id, nullptr);
    if (n_added) *n_added = (int32_t)n;
  }
  if (rc) return rc;
  c->next_id += (int32_t)n;
  return LIO_OK;
}

int lio_map_delete_boxes(lio_ctx* c, const float* boxes6, int nb, int32_t* n_deleted) {
  if (!c || nb < 0 || (nb > 0 && !boxes6)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  return map_delete_boxes(c, boxes6, nb, n_deleted);
}

int lio_map_size(lio_ctx* c, int64_t* total, int64_t* valid) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  uint32_t h[8];
  LIO_CHECK(c, cudaMemcpyAsync(h, c->map.counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  if (total) *total = c->next_id;
  if (valid) *valid = h[2];
  return LIO_OK;
}

int lio_map_removed_points(lio_ctx* c, float* xyz, int64_t cap, int64_t* n) {
  if (!c || cap < 0 || !n) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  return map_removed_points(c, xyz, cap, n);
}

int lio_map_dump(lio_ctx* c, float* xyz, int32_t* ids, int64_t cap, int64_t* n) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  return map_dump(c, xyz, ids, cap, n);
}

static int download_neighbors(lio_ctx* c, int64_t m, int32_t* idx5, float* d2_5, float* nbr_xyz) {
  if (m <= 0) return LIO_OK;
  if (idx5 || nbr_xyz) {
    std::vector<float4> h((size_t)m * LIO_K);
    LIO_CHECK(c, cudaMemcpyAsync(h.data(), c->d_near, sizeof(float4) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    LIO_CHECK(c, cudaStreamSynchronize(c->stream));
    for (size_t j = 0; j < h.size(); ++j) {
      if (idx5) memcpy(&idx5[j], &h[j].w, 4);
      if (nbr_xyz) {
        nbr_xyz[3 * j] = h[j].x;
        nbr_xyz[3 * j + 1] = h[j].y;
        nbr_xyz[3 * j + 2] = h[j].z;
      }
    }
  }
  if (d2_5) {
    LIO_CHECK(c, cudaMemcpyAsync(d2_5, c->d_near_d2, 4 * (size_t)m * LIO_K, cudaMemcpyDeviceToHost, c->stream));
    LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  return LIO_OK;
}

int lio_knn5(lio_ctx* c, const float* q_xyz, int64_t m, float max_d2, int32_t* idx5, float* d2_5, float* nbr_xyz) {
  if (!c || m < 0 || (m > 0 && !q_xyz) || !(max_d2 >= 0.f)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  // a bound beyond the hot search's (caps.knn_max_d2): the rows that search leaves short are completed by the
  // unbounded search, and whatever lies beyond max_d2 is cut off again below
  const bool beyond = max_d2 > c->caps.knn_max_d2;
  // processed in chunks of max_down_points through the scan-sized buffers (this invalidates cached neighbours)
  const int64_t chunk = c->caps.max_down_points;
  for (int64_t off = 0; off < m; off += chunk) {
    const int64_t cm = std::min(chunk, m - off);
    float* d_tmp = reinterpret_cast<float*>(c->d_add_a);  // 16 B/pt scratch >= 12 B/pt needed
    LIO_CHECK(c, cudaMemcpyAsync(d_tmp, q_xyz + 3 * off, 12 * (size_t)cm, cudaMemcpyHostToDevice, c->stream));
    xyz_to_float4_kernel<<<(int)((cm + 255) / 256), 256, 0, c->stream>>>(d_tmp, c->d_world, (int)cm);
    c->launches++;
    int rc = launch_knn_batch(c, c->d_world, cm);
    if (rc) return rc;
    if (beyond) {
      rc = launch_far_complete(c, c->d_world, cm, 0, cm, LIO_K);
      if (rc) return rc;
    }
    int32_t* ip = idx5 ? idx5 + off * LIO_K : nullptr;
    float* dp = d2_5 ? d2_5 + off * LIO_K : nullptr;
    float* xp = nbr_xyz ? nbr_xyz + off * LIO_K * 3 : nullptr;
    std::vector<float> dtmp;
    if (!dp && (beyond || max_d2 < c->caps.knn_max_d2)) {  // the cut needs the distances
      dtmp.resize((size_t)cm * LIO_K);
      dp = dtmp.data();
    }
    rc = download_neighbors(c, cm, ip, dp, xp);
    if (rc) return rc;
    if (dp && max_d2 != c->caps.knn_max_d2)
      for (int64_t j = 0; j < cm * LIO_K; ++j)
        if (!(dp[j] <= max_d2)) {  // `dist <= max_dist_sqr` (ikd_Tree.cpp:980)
          dp[j] = INFINITY;
          if (ip) ip[j] = -1;
          if (xp) xp[3 * j] = xp[3 * j + 1] = xp[3 * j + 2] = 0.f;
        }
  }
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

// Instrumentation: the batch kernel alone on the m queries the last lio_knn5 call left on the device (no copies, no
// sync), so that a tool can bracket exactly that kernel with CUDA events.
int lio_knn5_resident(lio_ctx* c, int64_t m) {
  if (!c || m < 0 || m > c->caps.max_down_points) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (!c->map_built) return LIO_E_EMPTY_MAP;
  return launch_knn_batch(c, c->d_world, m);
}

// ---------------------------------------------------------------- scan
// A lio_scan_step_prefetch copy may still be running on the second stream: order it before anything else touches d_raw
// on another stream (the growth was settled in between, or the caller came back with another buffer / entry point).
static int order_after_prefetch(lio_ctx* c) {
  if (c->staged_on_prep && c->stream != c->prep_stream) {
    LIO_CHECK(c, cudaEventRecord(c->ev_prep, c->prep_stream));
    LIO_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_prep, 0));
  }
  c->staged_on_prep = false;
  return LIO_OK;
}

static int preprocess_common(lio_ctx* c, const void* raw_pts, int64_t n, int stride, const lio_pose6d* poses,
                             int n_poses, const lio_state* end_state, float leaf, bool staged = false) {
  if (!c || n < 0 || (stride != 16 && stride != 48) || (n > 0 && !raw_pts) || !(leaf > 0.f)) return LIO_E_INVALID;
  if (n_poses >= 2 && (!poses || !end_state)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (n > c->caps.max_scan_points) {
    c->err = "scan larger than lio_caps.max_scan_points";
    return LIO_E_CAPACITY;
  }
  if (n_poses > 128) {
    c->err = "more than 128 IMU poses";
    return LIO_E_CAPACITY;
  }
  int rc = order_after_prefetch(c);
  if (rc) return rc;
  // stride 48: time = curvature (offset 36), intensity (offset 32) rides along as aux
  rc = staged ? LIO_OK
                  : stage_points(c, raw_pts, n, stride, 36, c->d_raw, 32, c->d_raw_aux, reinterpret_cast<float*>(c->d_vkeys));
  if (rc) return rc;
  if (n_poses >= 2)
    LIO_CHECK(c, cudaMemcpyAsync(c->d_poses, poses, sizeof(lio_pose6d) * n_poses, cudaMemcpyHostToDevice, c->stream));
  c->scan_m_bound = std::min<int64_t>(n, c->caps.max_down_points);
  return preprocess(c, n, n_poses >= 2 ? n_poses : 0, end_state, leaf, stride == 48);
}

static int preprocess_decode(lio_ctx* c, const int* h, int64_t* m);
static int preprocess_status(lio_ctx* c, int64_t* m) {
  int h[8];
  LIO_CHECK(c, cudaMemcpyAsync(h, c->d_prep_counters + 16, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return preprocess_decode(c, h, m);
}
// h = the first 8 preprocessing counters, already on the host
static int preprocess_decode(lio_ctx* c, const int* h, int64_t* m) {
  if (h[7] == 2) {
    c->err = "more occupied voxels than lio_caps.max_down_points";
    return LIO_E_CAPACITY;
  }
  if (h[7] == 3) {
    c->err = "VoxelGrid: leaf size too small for the scan extent (index would overflow)";
    c->scan_m = 0;
    if (m) *m = 0;
    return LIO_E_VOXEL_RANGE;
  }
  if (h[7] == 5) {
    c->err = "VoxelGrid: the scan's bounding grid has more than 2^28 leaves (leaf size too small for this extent)";
    return LIO_E_VOXEL_RANGE;
  }
  c->scan_m = std::min<int64_t>(h[0], c->caps.max_down_points);
  if (m) *m = c->scan_m;
  if (h[0] > 0) {
    const int64_t dx = (int64_t)h[4] - h[1] + 1, dy = (int64_t)h[5] - h[2] + 1, dz = (int64_t)h[6] - h[3] + 1;
    if (dx * dy * dz > (int64_t)INT32_MAX) {
      c->err = "VoxelGrid: leaf size too small for the scan extent (index would overflow)";
      return LIO_E_VOXEL_RANGE;
    }
  }
  return LIO_OK;
}

int lio_scan_preprocess(lio_ctx* c, const void* raw_pts, int64_t n, int stride, const lio_pose6d* poses, int n_poses,
                        const lio_state* end_state, float leaf, void* out_pts, int64_t* m, float* undistorted,
                        int32_t* voxel_key_xyz) {
  int rc = preprocess_common(c, raw_pts, n, stride, poses, n_poses, end_state, leaf);
  if (rc) return rc;
  int64_t mm = 0;
  rc = preprocess_status(c, &mm);
  if (m) *m = mm;
  if (rc) return rc;
  if (undistorted && n > 0)
    LIO_CHECK(c, cudaMemcpyAsync(undistorted, c->d_undist, 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (voxel_key_xyz && n > 0)
    LIO_CHECK(c, cudaMemcpyAsync(voxel_key_xyz, c->d_vkeys, 12 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (out_pts && mm > 0) {
    if (stride == 16) {
      LIO_CHECK(c, cudaMemcpyAsync(out_pts, c->d_body, 16 * (size_t)mm, cudaMemcpyDeviceToHost, c->stream));
    } else {
      std::vector<float4> hb((size_t)mm);
      std::vector<float> ht((size_t)mm);
      LIO_CHECK(c, cudaMemcpyAsync(hb.data(), c->d_body, 16 * (size_t)mm, cudaMemcpyDeviceToHost, c->stream));
      LIO_CHECK(c, cudaMemcpyAsync(ht.data(), c->d_normvec, 4 * (size_t)mm, cudaMemcpyDeviceToHost, c->stream));
      LIO_CHECK(c, cudaStreamSynchronize(c->stream));
      char* o = static_cast<char*>(out_pts);
      for (int64_t j = 0; j < mm; ++j) {
        float rec[12] = {hb[j].x, hb[j].y, hb[j].z, 1.0f, 0, 0, 0, 0, hb[j].w, ht[j], 0, 0};
        memcpy(o + 48 * j, rec, 48);
      }
    }
  }
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

int lio_scan_preprocess_resident(lio_ctx* c, const void* raw_pts, int64_t n, int stride, const lio_pose6d* poses,
                                 int n_poses, const lio_state* end_state, float leaf, int64_t* m) {
  int rc = preprocess_common(c, raw_pts, n, stride, poses, n_poses, end_state, leaf);
  if (rc) return rc;
  if (m) return preprocess_status(c, m);
  return LIO_OK;
}

int lio_scan_preprocess_cloud2(lio_ctx* c, const void* data, int64_t n, const lio_cloud_layout* L, const lio_pose6d* poses,
                               int n_poses, const lio_state* end_state, float leaf, int64_t* n_decoded, int64_t* m) {
  if (!c || n < 0 || (n > 0 && !data) || !L || !(leaf > 0.f)) return LIO_E_INVALID;
  const int time_bytes = L->time_type >= 2 ? 8 : 4;
  if (L->point_step < 12 || L->point_step > 256 || L->point_filter_num < 1 || L->rule < 1 || L->rule > 4 ||
      L->off_x < 0 || L->off_y < 0 || L->off_z < 0 || L->off_x + 4 > L->point_step || L->off_y + 4 > L->point_step ||
      L->off_z + 4 > L->point_step || L->intensity_type < 0 || L->intensity_type > 1 ||
      L->off_intensity + (L->intensity_type == 1 ? 1 : 4) > L->point_step ||
      (L->off_time >= 0 && L->off_time + time_bytes > L->point_step) || L->time_type < 0 || L->time_type > 3)
    return LIO_E_INVALID;
  const bool needs_ring = L->rule == 3 || ((L->rule == 2 || L->rule == 4) && L->yaw_time);
  if (needs_ring && (L->off_ring < 0 || L->ring_type < 0 || L->ring_type > 1 ||
                     L->off_ring + (L->ring_type == 1 ? 1 : 2) > L->point_step || L->n_scans < 1 || L->n_scans > 256))
    return LIO_E_INVALID;
  if (L->rule == 3 && (L->off_tag < 0 || L->off_tag + 1 > L->point_step || L->ring_type != 1 || L->off_time < 0 ||
                       L->time_type != 1))
    return LIO_E_INVALID;
  // given_offset_time (velodyne_handler :296-298, rs_handler :849-851): the LAST record's time field decides
  bool yaw_times = false;
  if ((L->rule == 2 || L->rule == 4) && L->yaw_time && n > 0) {
    if (L->scan_rate < 1) return LIO_E_INVALID;
    yaw_times = true;
    if (L->off_time >= 0) {
      const unsigned char* last = static_cast<const unsigned char*>(data) + (size_t)(n - 1) * L->point_step + L->off_time;
      if (L->time_type == 0) {
        float t;
        memcpy(&t, last, 4);
        yaw_times = !(t > 0);
      } else if (L->time_type == 1) {
        uint32_t t;
        memcpy(&t, last, 4);
        yaw_times = !(t > 0);
      } else {
        double t;
        memcpy(&t, last, 8);
        yaw_times = !(t > 0);
      }
    }
  }
  if (n_poses >= 2 && (!poses || !end_state)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (n > c->caps.max_scan_points) {
    c->err = "scan larger than lio_caps.max_scan_points";
    return LIO_E_CAPACITY;
  }
  if (n_poses > 128) {
    c->err = "more than 128 IMU poses";
    return LIO_E_CAPACITY;
  }
  const size_t need = (size_t)c->caps.max_scan_points * (size_t)L->point_step;
  if (need > c->cloud_bytes) {
    if (c->d_cloud) LIO_CHECK(c, cudaFree(c->d_cloud));
    c->d_cloud = nullptr;
    c->cloud_bytes = 0;
    LIO_CHECK(c, cudaMalloc(reinterpret_cast<void**>(&c->d_cloud), need));
    c->cloud_bytes = need;
  }
  if (const int ro = order_after_prefetch(c)) return ro;
  if (n > 0) LIO_CHECK(c, cudaMemcpyAsync(c->d_cloud, data, (size_t)n * L->point_step, cudaMemcpyHostToDevice, c->stream));
  int64_t nd = 0;
  int rc = decode_cloud2(c, n, *L, yaw_times, &nd);
  if (rc) return rc;
  c->n_decoded = nd;
  if (n_decoded) *n_decoded = nd;
  if (n_poses >= 2)
    LIO_CHECK(c, cudaMemcpyAsync(c->d_poses, poses, sizeof(lio_pose6d) * n_poses, cudaMemcpyHostToDevice, c->stream));
  rc = preprocess(c, nd, n_poses >= 2 ? n_poses : 0, end_state, leaf, true);
  if (rc) return rc;
  return preprocess_status(c, m);
}

int lio_scan_decoded(lio_ctx* c, float* xyzt, float* intensity, int64_t cap, int64_t* n) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (n) *n = c->n_decoded;
  const int64_t k = std::min<int64_t>(cap, c->n_decoded);
  // preprocess() reuses d_raw as the sorted-point buffer, so the decoded order is read back from the undistorted
  // cloud (input order; identical to the decoded cloud when no poses were given) and the intensity buffer
  if (xyzt && k > 0) LIO_CHECK(c, cudaMemcpyAsync(xyzt, c->d_undist, 16 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  if (intensity && k > 0)
    LIO_CHECK(c, cudaMemcpyAsync(intensity, c->d_raw_aux, 4 * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

int lio_scan_upload(lio_ctx* c, const void* down_pts, int64_t m, int stride) {
  if (!c || m < 0 || (stride != 16 && stride != 48) || (m > 0 && !down_pts)) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (m > c->caps.max_down_points) {
    c->err = "scan larger than lio_caps.max_down_points (the reference's arrays hold 100000, esekfom.hpp:23-29)";
    return LIO_E_CAPACITY;
  }
  int rc = stage_points(c, down_pts, m, stride, 32, c->d_body, 0, nullptr, reinterpret_cast<float*>(c->d_near_cnt));
  if (rc) return rc;
  set_scan_m_kernel<<<1, 1, 0, c->stream>>>(c->d_scan_m, (int)m);
  c->launches++;
  c->scan_m = m;
  return LIO_OK;
}

// ---------------------------------------------------------------- update
// upload staging area: wait until the previous asynchronous upload has left it
static int upload_area(lio_ctx* c, double** area) {
  LIO_CHECK(c, cudaEventSynchronize(c->upload_done));
  *area = static_cast<double*>(c->h_pinned) + 640;
  return LIO_OK;
}

int lio_update_pass(lio_ctx* c, const lio_state* x, int do_search, int extrinsic_est, double blob90[90],
                    int32_t* n_valid) {
  if (!c || !x) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (do_search && !c->map_built) {
    c->err = "update on an empty map";
    return LIO_E_EMPTY_MAP;
  }
  double* up = nullptr;
  int rc = upload_area(c, &up);
  if (rc) return rc;
  memcpy(up, x, sizeof(lio_state));
  LIO_CHECK(c, cudaMemcpyAsync(c->d_x, up, sizeof(lio_state), cudaMemcpyHostToDevice, c->stream));
  LIO_CHECK(c, cudaEventRecord(c->upload_done, c->stream));
  rc = launch_pass(c, do_search ? 1 : 0, extrinsic_est ? 1 : 0, -INFINITY, INFINITY);
  if (rc) return rc;
  double* hb = static_cast<double*>(c->h_pinned) + 1280;
  LIO_CHECK(c, cudaMemcpyAsync(hb, c->d_blob, 8 * LIO_BLOB, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  if (blob90) memcpy(blob90, hb, 8 * 90);
  if (n_valid) *n_valid = (int32_t)hb[90];
  return LIO_OK;
}

// One H2D copy into the snapshot {x0, P0}; to_current also makes it the current state {x, P}.
static int state_upload_impl(lio_ctx* c, const lio_state* x, const double P[576], bool to_current) {
  double* up = nullptr;
  int rc = upload_area(c, &up);
  if (rc) return rc;
  memcpy(up, x, sizeof(lio_state));
  memcpy(up + 26, P, 8 * 576);
  LIO_CHECK(c, cudaMemcpyAsync(c->d_x0, up, 8 * 602, cudaMemcpyHostToDevice, c->stream));
  LIO_CHECK(c, cudaEventRecord(c->upload_done, c->stream));
  if (to_current) LIO_CHECK(c, cudaMemcpyAsync(c->d_x, c->d_x0, 8 * 602, cudaMemcpyDeviceToDevice, c->stream));
  return LIO_OK;
}

int lio_state_upload(lio_ctx* c, const lio_state* x, const double P[576]) {
  if (!c || !x || !P) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  return state_upload_impl(c, x, P, true);
}

int lio_state_download(lio_ctx* c, lio_state* x, double P[576], int32_t* n_valid_last, int32_t* n_passes) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  double* hp = static_cast<double*>(c->h_pinned);
  LIO_CHECK(c, cudaMemcpyAsync(hp, c->d_x, 8 * 606, cudaMemcpyDeviceToHost, c->stream));  // {x, P, ctrl} in one copy
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  const Ctrl* hc = reinterpret_cast<const Ctrl*>(hp + 602);
  if (x) memcpy(x, hp, sizeof(lio_state));
  if (P) memcpy(P, hp + 26, 8 * 576);
  if (n_valid_last) *n_valid_last = hc->n_valid_last;
  if (n_passes) *n_passes = hc->n_passes;
  return LIO_OK;
}

int lio_update_enqueue(lio_ctx* c, double R, int max_iter, int extrinsic_est, int from_snapshot) {
  if (!c || max_iter < 0 || max_iter > 32) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (!c->map_built) {
    c->err = "update on an empty map";
    return LIO_E_EMPTY_MAP;
  }
  return launch_update(c, R, max_iter, extrinsic_est ? 1 : 0, from_snapshot ? 1 : 0);
}

int lio_update_scan(lio_ctx* c, lio_state* x_io, double P_io[576], double R, int max_iter, int extrinsic_est,
                    int32_t* n_valid_last, int32_t* n_passes) {
  if (!c || !x_io || !P_io || max_iter < 0 || max_iter > 32) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (!c->map_built) {
    c->err = "update on an empty map";
    return LIO_E_EMPTY_MAP;
  }
  // prior up in one copy, one kernel (which starts from the snapshot), posterior down in one copy
  int rc = state_upload_impl(c, x_io, P_io, false);
  if (rc) return rc;
  rc = launch_update(c, R, max_iter, extrinsic_est ? 1 : 0, 1);
  if (rc) return rc;
  return lio_state_download(c, x_io, P_io, n_valid_last, n_passes);
}

// The call a host makes per scan when the downsampled cloud is in host memory.  Host-direct: the prior rides in the
// kernel parameters, the posterior comes back through mapped pinned memory (the host spins on a sequence word instead
// of paying a copy + stream synchronisation).  With LIO_ZERO_COPY=1 a cloud in PINNED memory (cudaHostAlloc / torch
// pin_memory) is read in place over PCIe by pass 0 instead of being copied first (no gain measured: off by default).
int lio_update_scan_host(lio_ctx* c, const void* down_pts, int64_t m, int stride, lio_state* x_io, double P_io[576],
                         double R, int max_iter, int extrinsic_est, int32_t* n_valid_last, int32_t* n_passes) {
  if (!c || !x_io || !P_io || m < 0 || (m > 0 && !down_pts) || stride != 16 || max_iter < 0 || max_iter > 32)
    return LIO_E_INVALID;  // stride 48 clouds go through lio_scan_upload + lio_update_scan
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (!c->map_built) {
    c->err = "update on an empty map";
    return LIO_E_EMPTY_MAP;
  }
  if (m > c->caps.max_down_points) {
    c->err = "scan larger than lio_caps.max_down_points";
    return LIO_E_CAPACITY;
  }
  HostDirect hd;
  hd.x0 = reinterpret_cast<const double*>(x_io);
  hd.P0 = P_io;
  hd.m = (int)m;
  hd.body_src = nullptr;
  if (m > 0) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, down_pts) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
        attr.devicePointer != nullptr && !c->no_zero_copy) {
      hd.body_src = static_cast<const float4*>(attr.devicePointer);
    } else {
      cudaGetLastError();  // pageable memory: not an error, just a copy
      LIO_CHECK(c, cudaMemcpyAsync(c->d_body, down_pts, 16 * (size_t)m, cudaMemcpyHostToDevice, c->stream));
    }
  }
  unsigned long long* flag = reinterpret_cast<unsigned long long*>(c->h_out + 608);
  hd.out_dev = c->h_out_dev;
  hd.flag_dev = reinterpret_cast<unsigned long long*>(c->h_out_dev + 608);
  hd.seq = ++c->host_seq;
  c->scan_m = m;
  int rc = launch_update(c, R, max_iter, extrinsic_est ? 1 : 0, 1, -INFINITY, INFINITY, false, &hd);
  if (rc) return rc;
  // spin on the sequence word; look at the stream now and then so that a failed launch cannot hang the caller
  volatile unsigned long long* vf = flag;
  for (unsigned long long spins = 0; *vf != hd.seq; ++spins) {
    if ((spins & 0x3fff) == 0x3fff) {
      const cudaError_t q = cudaStreamQuery(c->stream);
      if (q != cudaErrorNotReady) {
        if (q != cudaSuccess) LIO_CHECK(c, q);
        if (*vf != hd.seq) {  // kernel finished without reporting: cannot happen with a healthy launch
          c->err = "update kernel finished without writing the posterior";
          return LIO_E_CUDA;
        }
      }
    }
  }
  const Ctrl* hc = reinterpret_cast<const Ctrl*>(c->h_out + 602);
  memcpy(x_io, c->h_out, sizeof(lio_state));
  memcpy(P_io, c->h_out + 26, 8 * 576);
  if (n_valid_last) *n_valid_last = hc->n_valid_last;
  if (n_passes) *n_passes = hc->n_passes;
  return LIO_OK;
}

// ---------------------------------------------------------------- one main-loop iteration per call
// laserMapping.cpp:737-785 for one scan -- VoxelGrid(UndistortPcl(scan)), the feats_down_size < 5 skip, the first-scan
// Build, update_iterated_dyn_share_modified, map_incremental -- enqueued back to back: M, the class counts and the insert
// sizes stay on the device, the host synchronises ONCE, at the end, for {posterior, M, counts, error flags}.
// The raw records of the next scan do not depend on the filter state: a host that still has the IMU propagation to do
// (lio_seq_process) sends them first, so the copy runs underneath that host work instead of after it.
int lio_scan_step_prefetch(lio_ctx* c, const void* raw_pts, int64_t n, int stride) {
  if (!c || n < 0 || (stride != 16 && stride != 48) || (n > 0 && !raw_pts)) return LIO_E_INVALID;
  c->staged_ptr = nullptr;
  if (n == 0 || !c->map_built || n > c->caps.max_scan_points) return LIO_OK;  // nothing to gain / lio_scan_step_begin reports
  LIO_CHECK(c, cudaSetDevice(c->device));
  const bool split = c->deferred_growth && c->growth_pending;
  cudaStream_t main_stream = c->stream;
  if (split) {
    LIO_CHECK(c, cudaStreamWaitEvent(c->prep_stream, c->ev_post, 0));
    c->stream = c->prep_stream;
  } else if (const int ro = order_after_prefetch(c)) {  // an earlier, dropped prefetch may still be copying there
    return ro;
  }
  const int rc = stage_points(c, raw_pts, n, stride, 36, c->d_raw, 32, c->d_raw_aux, reinterpret_cast<float*>(c->d_vkeys));
  c->stream = main_stream;
  if (rc) return rc;
  c->staged_ptr = raw_pts;
  c->staged_n = n;
  c->staged_stride = stride;
  c->staged_on_prep = split;
  return LIO_OK;
}

int lio_scan_step_begin(lio_ctx* c, const void* raw_pts, int64_t n, int stride, const lio_pose6d* poses, int n_poses,
                        const lio_state* x, const double P[576], float leaf_surf, int32_t* update_due) {
  if (!c || !x || !P || !update_due) return LIO_E_INVALID;
  *update_due = 0;
  c->step_phase = 0;
  c->min_m = 0;
  const bool staged = c->staged_ptr != nullptr && c->staged_ptr == raw_pts && c->staged_n == n && c->staged_stride == stride;
  c->staged_ptr = nullptr;
  // A deferred growth of the previous scan is still on the stream: upload, undistort and sort this scan on the second
  // stream next to it.  The growth reads d_body / d_scan_m (classification), so the one kernel that overwrites them
  // (centroids) waits for it; everything before touches preprocessing buffers only.  The host has the previous
  // posterior (ev_post), so the previous update is complete and nothing else reads those buffers.
  const bool split = c->deferred_growth && c->growth_pending && c->map_built;
  cudaStream_t main_stream = c->stream;
  if (split) {
    LIO_CHECK(c, cudaSetDevice(c->device));
    LIO_CHECK(c, cudaStreamWaitEvent(c->prep_stream, c->ev_post, 0));
    c->stream = c->prep_stream;
    c->centroid_wait = c->ev_growth;
  }

  int rc = preprocess_common(c, raw_pts, n, stride, poses, n_poses, x, leaf_surf, staged);
  if (split) {
    c->stream = main_stream;
    c->centroid_wait = nullptr;
    LIO_CHECK(c, cudaEventRecord(c->ev_prep, c->prep_stream));
    LIO_CHECK(c, cudaStreamWaitEvent(main_stream, c->ev_prep, 0));
  }
  if (rc) return rc;
  if (!c->map_built) {  // first scan with points: host-synchronous, happens once per sequence
    int64_t m = 0;
    rc = preprocess_status(c, &m);
    if (rc) return rc;
    c->step_m = m;
    c->step_status = LIO_SCAN_FEW_POINTS;
    if (m >= 5) {
      rc = map_build_scan(c, x);
      if (rc) return rc;
      c->step_status = LIO_SCAN_MAP_BUILT;
    }
    c->step_phase = 3;
    return LIO_OK;
  }
  rc = state_upload_impl(c, x, P, false);
  if (rc) return rc;
  c->min_m = 5;  // update and map growth see a scan of fewer than 5 points as empty
  c->step_phase = 1;
  *update_due = 1;
  return LIO_OK;
}

int lio_scan_step_end(lio_ctx* c, float leaf_map, int ekf_inited) {
  if (!c || !(leaf_map >= 0.f)) return LIO_E_INVALID;
  if (c->step_phase == 3) return LIO_OK;
  if (c->step_phase != 1) {
    c->err = "lio_scan_step_end without lio_scan_step_begin";
    return LIO_E_INVALID;
  }
  LIO_CHECK(c, cudaSetDevice(c->device));
  int rc = settle_growth(c);  // the previous scan's growth is long done by now: books next_id for this one
  if (rc) return rc;
  if (leaf_map > 0.f && (rc = check_id_space(c, c->scan_m_bound)) != LIO_OK) return rc;
  // posterior + preprocess counters first: with deferred growth the host resumes as soon as these have landed
  double* hp = static_cast<double*>(c->h_pinned);
  LIO_CHECK(c, cudaMemcpyAsync(hp, c->d_x, 8 * 606, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaMemcpyAsync(hp + 606, c->d_prep_counters + 16, 4 * 8, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaEventRecord(c->ev_post, c->stream));
  // leaf_map == 0: the map is static -- the relocalisation loop, where map_incremental() is commented out
  // (src/laserMapping_re.cpp:676); the counts of the report are then zero
  if (leaf_map > 0.f) {
    rc = map_incremental_enqueue(c, leaf_map, ekf_inited ? 1 : 0, 5, c->scan_m_bound);
    if (rc) return rc;
    LIO_CHECK(c, cudaMemcpyAsync(hp + 620, c->d_prep_counters + 8, 4 * 2, cudaMemcpyDeviceToHost, c->stream));
    LIO_CHECK(c, cudaMemcpyAsync(hp + 622, c->map.counters, 4 * 8, cudaMemcpyDeviceToHost, c->stream));
  } else {  // nothing is enqueued for a static map: the counts the report reads are zero by definition
    memset(hp + 620, 0, 8 * 6);
  }
  LIO_CHECK(c, cudaEventRecord(c->ev_growth, c->stream));
  c->growth_pending = true;
  c->step_phase = 2;
  return LIO_OK;
}

int lio_scan_step_finish(lio_ctx* c, lio_state* x_out, double P_out[576], lio_scan_report* rep) {
  if (!c || !rep) return LIO_E_INVALID;
  memset(rep, 0, sizeof(*rep));
  c->min_m = 0;
  if (c->step_phase == 3) {
    c->step_phase = 0;
    rep->m = c->step_m;
    rep->status = c->step_status;
    return LIO_OK;
  }
  if (c->step_phase != 2) {
    c->err = "lio_scan_step_finish without lio_scan_step_end";
    return LIO_E_INVALID;
  }
  c->step_phase = 0;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (c->deferred_growth) {
    cudaError_t q;
    while ((q = cudaEventQuery(c->ev_post)) == cudaErrorNotReady) {
    }
    LIO_CHECK(c, q);
  } else {
    LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  const double* hp = static_cast<const double*>(c->h_pinned);
  const int* prep = reinterpret_cast<const int*>(hp + 606);
  int64_t m = 0;
  int rc = preprocess_decode(c, prep, &m);
  rep->m = m;
  if (rc) return rc;
  if (!c->deferred_growth) {
    rc = settle_growth(c);
    if (rc) return rc;
  }
  if (m < 5) {
    rep->status = LIO_SCAN_FEW_POINTS;  // state untouched, as after the reference's `continue`
    return LIO_OK;
  }
  const Ctrl* hc = reinterpret_cast<const Ctrl*>(hp + 602);
  rep->status = LIO_SCAN_UPDATED;
  rep->n_valid = hc->n_valid_last;
  rep->n_passes = hc->n_passes;
  for (int k = 0; k < 3; ++k) rep->counts[k] = c->deferred_growth ? -1 : c->last_counts[k];
  if (x_out) memcpy(x_out, hp, sizeof(lio_state));
  if (P_out) memcpy(P_out, hp + 26, 8 * 576);
  return LIO_OK;
}

int lio_scan_step(lio_ctx* c, const void* raw_pts, int64_t n, int stride, const lio_pose6d* poses, int n_poses,
                  lio_state* x_io, double P_io[576], float leaf_surf, float leaf_map, double R, int max_iter,
                  int extrinsic_est, int ekf_inited, lio_scan_report* rep) {
  if (!c || !x_io || !P_io || !rep || max_iter < 0 || max_iter > 32 || !(leaf_map >= 0.f)) return LIO_E_INVALID;
  int32_t due = 0;
  int rc = lio_scan_step_begin(c, raw_pts, n, stride, poses, n_poses, x_io, P_io, leaf_surf, &due);
  if (rc) return rc;
  if (due) {
    rc = launch_update(c, R, max_iter, extrinsic_est ? 1 : 0, 1);
    if (rc) return rc;
    rc = lio_scan_step_end(c, leaf_map, ekf_inited);
    if (rc) return rc;
  }
  return lio_scan_step_finish(c, x_io, P_io, rep);
}

// ---------------------------------------------------------------- several independent sequences per launch
int lio_update_enqueue_multi(lio_ctx* const* ctxs, int n, double R, int max_iter, int extrinsic_est, int from_snapshot) {
  if (!ctxs || n < 1 || n > 8 || max_iter < 0 || max_iter > 32) return LIO_E_INVALID;
  lio_ctx* c = ctxs[0];
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  for (int q = 0; q < n; ++q) {
    if (!ctxs[q] || ctxs[q]->device != c->device) return LIO_E_INVALID;
    if (!ctxs[q]->map_built) {
      c->err = "update on an empty map";
      return LIO_E_EMPTY_MAP;
    }
    for (int r = 0; r < q; ++r)
      if (ctxs[r] == ctxs[q]) return LIO_E_INVALID;
  }
  // whatever the other contexts have enqueued on their own streams (scan / prior uploads) comes first ...
  for (int q = 1; q < n; ++q)
    if (ctxs[q]->stream != c->stream) {
      LIO_CHECK(c, cudaEventRecord(ctxs[q]->multi_evt, ctxs[q]->stream));
      LIO_CHECK(c, cudaStreamWaitEvent(c->stream, ctxs[q]->multi_evt, 0));
    }
  const int rc = launch_update_multi(ctxs, n, R, max_iter, extrinsic_est ? 1 : 0, from_snapshot ? 1 : 0);
  if (rc) return rc;
  // ... and what they enqueue next (downloads) comes after the launch
  LIO_CHECK(c, cudaEventRecord(c->multi_evt, c->stream));
  for (int q = 1; q < n; ++q)
    if (ctxs[q]->stream != c->stream) LIO_CHECK(c, cudaStreamWaitEvent(ctxs[q]->stream, c->multi_evt, 0));
  return LIO_OK;
}

// ---------------------------------------------------------------- sharded map over peer memory
int lio_peer_handle(lio_ctx* c, unsigned char handle[64]) {
  if (!c || !handle) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  LIO_CHECK(c, cudaIpcGetMemHandle(&h, c->d_mailbox));
  memcpy(handle, &h, 64);
  return LIO_OK;
}

int lio_peer_connect(lio_ctx* c, int rank, int world, const unsigned char* handles) {
  if (!c || !handles || world < 2 || world > 8 || rank < 0 || rank >= world) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  for (int r = 0; r < world; ++r) {
    void* base = c->d_mailbox;
    if (r != rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * r, 64);
      LIO_CHECK(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
      c->peer_base[r] = base;
    }
    c->peer_mbox[r] = static_cast<unsigned long long*>(base);
  }
  c->d_peer_err = reinterpret_cast<int*>(static_cast<char*>(c->d_mailbox) + MAILBOX_ERR);
  c->peer_rank = rank;
  c->peer_world = world;
  c->peer_epoch = 0;
  return LIO_OK;
}

int lio_set_shard_stripes(lio_ctx* c, float x_origin, float stripe_width, int world, int rank) {
  if (!c) return LIO_E_INVALID;
  if (world == 0) {  // back to the x windows of the calls
    c->stripe_world = 0;
    return LIO_OK;
  }
  if (world < 1 || world > 8 || rank < 0 || rank >= world || !(stripe_width > 0.f) || !(x_origin == x_origin))
    return LIO_E_INVALID;
  c->stripe_world = world;
  c->stripe_rank = rank;
  c->stripe_origin = x_origin;
  c->stripe_width = stripe_width;
  return LIO_OK;
}

int lio_update_enqueue_sharded(lio_ctx* c, double R, int max_iter, int extrinsic_est, int from_snapshot, float x_own_min,
                               float x_own_max) {
  if (!c || max_iter < 0 || max_iter > 32) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (!c->map_built) {
    c->err = "update on an empty map";
    return LIO_E_EMPTY_MAP;
  }
  return launch_update(c, R, max_iter, extrinsic_est ? 1 : 0, from_snapshot ? 1 : 0, x_own_min, x_own_max, true);
}

int lio_peer_status(lio_ctx* c, int32_t* timed_out) {
  if (!c || !timed_out) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  int v = 0;
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  LIO_CHECK(c, cudaMemcpy(&v, static_cast<char*>(c->d_mailbox) + MAILBOX_ERR, sizeof(int), cudaMemcpyDeviceToHost));
  *timed_out = v;
  return LIO_OK;
}

int lio_update_begin(lio_ctx* c, int max_iter, int extrinsic_est, int from_snapshot) {
  if (!c || max_iter < 0 || max_iter > 32) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  return launch_begin(c, max_iter, extrinsic_est ? 1 : 0, from_snapshot ? 1 : 0);
}

int lio_update_pass_enqueue(lio_ctx* c, int extrinsic_est, float x_own_min, float x_own_max) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  return launch_pass(c, -1, extrinsic_est ? 1 : 0, x_own_min, x_own_max, true);
}

int lio_update_step_enqueue(lio_ctx* c, double R, int extrinsic_est) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  return launch_solve(c, R, extrinsic_est ? 1 : 0);
}

int lio_pass_only_enqueue(lio_ctx* c, int do_search, int extrinsic_est) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (do_search && !c->map_built) return LIO_E_EMPTY_MAP;
  return launch_pass(c, do_search ? 1 : 0, extrinsic_est ? 1 : 0, -INFINITY, INFINITY);
}

int lio_debug_timeline(lio_ctx* c, int64_t out[256]) {
  if (!c || !out) return LIO_E_INVALID;
  if (!c->d_dbg) {
    c->err = "timeline not enabled (set LIO_TIMELINE=1 before lio_create)";
    return LIO_E_INVALID;
  }
  LIO_CHECK(c, cudaSetDevice(c->device));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  LIO_CHECK(c, cudaMemcpy(out, c->d_dbg, 256 * sizeof(long long), cudaMemcpyDeviceToHost));
  return LIO_OK;
}

int lio_debug_blocks(lio_ctx* c, int64_t out[768]) {
  if (!c || !out) return LIO_E_INVALID;
  if (!c->d_dbg) {
    c->err = "timeline not enabled (set LIO_TIMELINE=1 before lio_create)";
    return LIO_E_INVALID;
  }
  LIO_CHECK(c, cudaSetDevice(c->device));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  LIO_CHECK(c, cudaMemcpy(out, c->d_dbg + 256, 768 * sizeof(long long), cudaMemcpyDeviceToHost));
  return LIO_OK;
}

void* lio_blob_device_ptr(lio_ctx* c) { return c ? c->d_blob : nullptr; }

int lio_blob_upload(lio_ctx* c, const double blob92[92]) {
  if (!c || !blob92) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  double* up = nullptr;
  int rc = upload_area(c, &up);
  if (rc) return rc;
  memcpy(up, blob92, 8 * LIO_BLOB);
  LIO_CHECK(c, cudaMemcpyAsync(c->d_blob, up, 8 * LIO_BLOB, cudaMemcpyHostToDevice, c->stream));
  LIO_CHECK(c, cudaEventRecord(c->upload_done, c->stream));
  return LIO_OK;
}

int lio_blob_bind(lio_ctx* c, void* device_buffer) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  c->d_blob = device_buffer ? static_cast<double*>(device_buffer) : c->d_blob_own;
  return LIO_OK;
}

int lio_blob_download(lio_ctx* c, double blob92[92]) {
  if (!c || !blob92) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  LIO_CHECK(c, cudaMemcpyAsync(blob92, c->d_blob, 8 * LIO_BLOB, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

int lio_get_neighbors(lio_ctx* c, int32_t* idx5, float* d2_5, float* nbr_xyz, float* world_xyz, uint8_t* selected,
                      float* normvec) {
  if (!c) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  int rc = refresh_scan_m(c);
  if (rc) return rc;
  const int64_t m = c->scan_m;
  if (m <= 0) return LIO_OK;
  if (idx5 || d2_5 || nbr_xyz) {
    // the rows the update's bounded search left short become what the reference's unbounded Nearest_Search leaves in
    // Nearest_Points (esekfom.hpp:140-141): five neighbours wherever the map holds five points
    if (!c->map_built) return LIO_E_EMPTY_MAP;
    rc = launch_far_complete(c, c->d_near_q, m, 0, m, LIO_K);
    if (rc) return rc;
  }
  rc = download_neighbors(c, m, idx5, d2_5, nbr_xyz);
  if (rc) return rc;
  std::vector<float4> h;
  if (world_xyz) {
    h.resize((size_t)m);
    LIO_CHECK(c, cudaMemcpyAsync(h.data(), c->d_world, 16 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
    LIO_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < m; ++i) {
      world_xyz[3 * i] = h[i].x;
      world_xyz[3 * i + 1] = h[i].y;
      world_xyz[3 * i + 2] = h[i].z;
    }
  }
  if (selected) LIO_CHECK(c, cudaMemcpyAsync(selected, c->d_selected, (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  if (normvec) LIO_CHECK(c, cudaMemcpyAsync(normvec, c->d_normvec, 16 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  LIO_CHECK(c, cudaStreamSynchronize(c->stream));
  return LIO_OK;
}

int lio_map_incremental(lio_ctx* c, const lio_state* x, float filter_size_map, int ekf_inited, int32_t counts[3]) {
  if (!c || !x || !counts) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  int rc = refresh_scan_m(c);
  if (rc) return rc;
  return map_incremental(c, x, filter_size_map, ekf_inited, counts);
}

int lio_map_build_scan(lio_ctx* c, const lio_state* x) {
  if (!c || !x) return LIO_E_INVALID;
  LIO_CHECK(c, cudaSetDevice(c->device));
  if (const int rs = settle_growth(c)) return rs;
  int rc = refresh_scan_m(c);
  if (rc) return rc;
  return map_build_scan(c, x);
}

// ---------------------------------------------------------------- host-side sequential pieces
int lio_boxplus(const lio_state* x, const double f[24], lio_state* out) {
  if (!x || !f || !out) return LIO_E_INVALID;
  StateD r;
  boxplus(*reinterpret_cast<const StateD*>(x), f, r);
  memcpy(out, &r, sizeof(r));
  return LIO_OK;
}
int lio_boxminus(const lio_state* x1, const lio_state* x2, double out[24]) {
  if (!x1 || !x2 || !out) return LIO_E_INVALID;
  boxminus(*reinterpret_cast<const StateD*>(x1), *reinterpret_cast<const StateD*>(x2), out);
  return LIO_OK;
}

// esekf::predict (esekfom.hpp:82-95) with get_f / df_dx / df_dw of use-ikfom.hpp:57-123.  The Jacobians are sparse;
// the products are formed block-wise instead of as dense 24x24x24 loops.
int lio_predict(lio_state* xs, double P[576], double dt, const double Q[144], const double acc[3],
                const double gyro[3]) {
  if (!xs || !P || !Q || !acc || !gyro) return LIO_E_INVALID;
  StateD& x = *reinterpret_cast<StateD*>(xs);
  double R[9];
  quat_to_mat(x.rot, R);
  const double am[3] = {acc[0] - x.ba[0], acc[1] - x.ba[1], acc[2] - x.ba[2]};
  double a_in[3];
  mat3_vec(R, am, a_in);
  double f[24] = {0};
  for (int i = 0; i < 3; ++i) {
    f[i] = x.vel[i] * dt;
    f[3 + i] = (gyro[i] - x.bg[i]) * dt;
    f[12 + i] = (a_in[i] + x.grav[i]) * dt;
  }
  const double hat[9] = {0, -am[2], am[1], am[2], 0, -am[0], -am[1], am[0], 0};
  double Rh[9];
  mat3_mul(R, hat, Rh);
  StateD xn;
  boxplus(x, f, xn);
  x = xn;
  // P <- F P F^T + W Q W^T (esekfom.hpp:93-94) with F = I + dt df_dx, W = dt df_dw (use-ikfom.hpp:82-123).  The
  // Jacobians are sparse and their pattern is fixed, so the products are written out over the non-zero terms, every sum
  // in ascending k exactly as the dense triple loop adds them (the terms left out are exact zeros: same bits).  Rows of F:
  //   pos  i     : 1 at (i,i), dt at (i,12+i)           rot 3+i : 1 at (3+i,3+i), -dt at (3+i,15+i)
  //   vel  12+i  : A = -R hat(a) dt at cols 3..5, 1 at (12+i,12+i), Cm = -R dt at cols 18..20, dt at (12+i,21+i)
  //   all other rows: identity.       Rows of W: rot -dt at col i; vel -R dt at cols 3..5; bg dt at 6+i; ba dt at 9+i.
  // This runs ~20 times per scan on the host, on the critical path between two updates.
  double A[9], Cm[9];
  for (int i = 0; i < 9; ++i) {
    A[i] = -Rh[i] * dt;
    Cm[i] = -R[i] * dt;
  }
  const double ndt = -dt;
  double T[576], Pn[576], WQ[24 * 12];
  for (int i = 0; i < 24; ++i) {
    const double* Pi = P + i * 24;
    double* Ti = T + i * 24;
    if (i < 3) {
      const double* Pb = P + (12 + i) * 24;
      for (int j = 0; j < 24; ++j) Ti[j] = (0.0 + Pi[j]) + dt * Pb[j];
    } else if (i < 6) {
      const double* Pb = P + (12 + i) * 24;  // 15 + (i - 3)
      for (int j = 0; j < 24; ++j) Ti[j] = (0.0 + Pi[j]) + ndt * Pb[j];
    } else if (i >= 12 && i < 15) {
      const double* a = A + 3 * (i - 12);
      const double* cm = Cm + 3 * (i - 12);
      const double *P3 = P + 3 * 24, *P4 = P + 4 * 24, *P5 = P + 5 * 24, *P18 = P + 18 * 24, *P19 = P + 19 * 24,
                   *P20 = P + 20 * 24, *Pg = P + (9 + i) * 24;  // 21 + (i - 12)
      for (int j = 0; j < 24; ++j) {
        double v = 0.0 + a[0] * P3[j];
        v += a[1] * P4[j];
        v += a[2] * P5[j];
        v += Pi[j];
        v += cm[0] * P18[j];
        v += cm[1] * P19[j];
        v += cm[2] * P20[j];
        v += dt * Pg[j];
        Ti[j] = v;
      }
    } else {
      for (int j = 0; j < 24; ++j) Ti[j] = 0.0 + Pi[j];
    }
  }
  for (int i = 0; i < 24 * 12; ++i) WQ[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 12; ++j) {
      WQ[(3 + i) * 12 + j] = 0.0 + ndt * Q[i * 12 + j];
      double v = 0.0 + Cm[3 * i] * Q[3 * 12 + j];
      v += Cm[3 * i + 1] * Q[4 * 12 + j];
      v += Cm[3 * i + 2] * Q[5 * 12 + j];
      WQ[(12 + i) * 12 + j] = v;
      WQ[(15 + i) * 12 + j] = 0.0 + dt * Q[(6 + i) * 12 + j];
      WQ[(18 + i) * 12 + j] = 0.0 + dt * Q[(9 + i) * 12 + j];
    }
  for (int i = 0; i < 24; ++i) {
    const double* Ti = T + i * 24;
    const double* Wi = WQ + i * 12;
    double* Pi = Pn + i * 24;
    for (int j = 0; j < 3; ++j) {
      Pi[j] = ((0.0 + Ti[j]) + Ti[12 + j] * dt) + 0.0;
      Pi[3 + j] = ((0.0 + Ti[3 + j]) + Ti[15 + j] * ndt) + (0.0 + Wi[j] * ndt);
      Pi[6 + j] = 0.0 + Ti[6 + j];
      Pi[9 + j] = 0.0 + Ti[9 + j];
      double v = 0.0 + Ti[3] * A[3 * j];
      v += Ti[4] * A[3 * j + 1];
      v += Ti[5] * A[3 * j + 2];
      v += Ti[12 + j];
      v += Ti[18] * Cm[3 * j];
      v += Ti[19] * Cm[3 * j + 1];
      v += Ti[20] * Cm[3 * j + 2];
      v += Ti[21 + j] * dt;
      double w = 0.0 + Wi[3] * Cm[3 * j];
      w += Wi[4] * Cm[3 * j + 1];
      w += Wi[5] * Cm[3 * j + 2];
      Pi[12 + j] = v + w;
      Pi[15 + j] = (0.0 + Ti[15 + j]) + (0.0 + Wi[6 + j] * dt);
      Pi[18 + j] = (0.0 + Ti[18 + j]) + (0.0 + Wi[9 + j] * dt);
      Pi[21 + j] = 0.0 + Ti[21 + j];
    }
  }
  memcpy(P, Pn, 8 * 576);
  return LIO_OK;
}

}  // extern "C"
