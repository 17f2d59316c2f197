// tsc_api.cu — the C ABI of libtostore_cuda.so (include/tostore_cuda.h).
//
// Host-side mirror of the slice of VectorIndexManager that surrounds the engine
// call (core/vector_index_manager.dart:475-589) plus corpus ingestion. All
// compute runs in the sm_100a kernels of this library; there is no CPU fallback.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "tsc_index.h"
#include "tsc_ingest.cuh"

namespace tsc {

// ---- error string -----------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// ---- handle registry ----------------------------------------------------------
static std::mutex g_mu;
static std::unordered_map<uint64_t, Index *> g_index;
static uint64_t g_next_handle = 1;

static Index *lookup(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_index.find(h);
  if (it == g_index.end()) {
    set_error("unknown index handle %llu", (unsigned long long)h);
    return nullptr;
  }
  return it->second;
}

Index *lookup_index(uint64_t h) { return lookup(h); }

uint64_t register_index(Index *ix) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next_handle++;
  g_index[h] = ix;
  return h;
}

struct Ticket {
  Index *ix;
  uint64_t handle;
  cudaEvent_t done;
  int64_t *out_ids;
  double *out_dist;
  uint32_t *out_counts;
  uint32_t nq, k;
};
static std::unordered_map<uint64_t, Ticket *> g_tickets;
static uint64_t g_next_ticket = 1;

template <typename T>
static cudaError_t dev_alloc(Index *ix, T **p, size_t n) {
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e == cudaSuccess) ix->device_bytes += n * sizeof(T);
  return e;
}

static void free_index(Index *ix) {
  if (ix->host_only) {
    delete ix;
    return;
  }
  cudaSetDevice(ix->device);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  cudaFree(ix->d_rows);
  cudaFree(ix->d_deleted);
  cudaFree(ix->d_filter);
  cudaFree(ix->d_live);
  cudaFree(ix->d_delta);
  cudaFree(ix->d_live_count);
  cudaFree(ix->d_queries);
  cudaFree(ix->d_norm2);
  cudaFree(ix->d_q16);
  cudaFree(ix->d_progress);
  cudaFree(ix->d_cand);
  cudaFree(ix->d_out_ids);
  cudaFree(ix->d_out_dist);
  cudaFree(ix->d_out_counts);
  cudaFree(ix->d_stage);
  cudaFree(ix->d_page_status);
  cudaFree(ix->d_gather_send);
  cudaFree(ix->d_gather_recv);
  for (int r = 0; r < ix->n_ranks && r < 8; r++)
    if (ix->x_peer[r] && ix->x_peer[r] != ix->d_xbuf) cudaIpcCloseMemHandle(ix->x_peer[r]);
  cudaFree(ix->d_xbuf);
  if (ix->h_xstatus) cudaFreeHost(ix->h_xstatus);
  cudaFree(ix->d_where_args);
  for (auto &c : ix->columns) {
    cudaFree(c.d_values);
    cudaFree(c.d_null);
  }
  cudaFreeHost(ix->h_queries);
  cudaFreeHost(ix->h_out_ids);
  cudaFreeHost(ix->h_out_dist);
  cudaFreeHost(ix->h_out_counts);
  for (int i = 0; i < Index::kTimers; i++) {
    if (ix->t_beg[i]) cudaEventDestroy(ix->t_beg[i]);
    if (ix->t_end[i]) cudaEventDestroy(ix->t_end[i]);
  }
  if (ix->ev0) cudaEventDestroy(ix->ev0);
  if (ix->ev1) cudaEventDestroy(ix->ev1);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
}

int32_t hot_timer_resolve(Index *ix) {
  while (ix->t_pending > 0) {
    int i = (ix->t_head - ix->t_pending + 2 * Index::kTimers) % Index::kTimers;
    TSC_CUDA(cudaEventSynchronize(ix->t_end[i]));
    float ms = 0;
    TSC_CUDA(cudaEventElapsedTime(&ms, ix->t_beg[i], ix->t_end[i]));
    ix->hot_ms += ms;
    ix->hot_bytes += ix->t_bytes[i];
    ix->hot_flops += ix->t_flops[i];
    ix->hot_launches++;
    ix->t_pending--;
  }
  return TSC_OK;
}

int32_t hot_timer_begin(Index *ix, cudaStream_t st, int *slot) {
  if (ix->t_pending == Index::kTimers) {
    int32_t rc = hot_timer_resolve(ix);
    if (rc != TSC_OK) return rc;
  }
  *slot = ix->t_head;
  TSC_CUDA(cudaEventRecord(ix->t_beg[*slot], st));
  return TSC_OK;
}

int32_t hot_timer_end(Index *ix, cudaStream_t st, int slot, double bytes, double flops) {
  TSC_CUDA(cudaEventRecord(ix->t_end[slot], st));
  ix->t_bytes[slot] = bytes;
  ix->t_flops[slot] = flops;
  ix->t_head = (ix->t_head + 1) % Index::kTimers;
  ix->t_pending++;
  return TSC_OK;
}

static int32_t ensure_stage(Index *ix, size_t bytes) {
  if (ix->stage_bytes >= bytes) return TSC_OK;
  if (ix->d_stage) {
    cudaFree(ix->d_stage);
    ix->device_bytes -= ix->stage_bytes;
    ix->d_stage = nullptr;
    ix->stage_bytes = 0;
  }
  cudaError_t e = cudaMalloc((void **)&ix->d_stage, bytes);
  if (e != cudaSuccess) {
    set_error("staging buffer of %zu bytes: %s", bytes, cudaGetErrorString(e));
    return TSC_ERR_OOM;
  }
  ix->stage_bytes = bytes;
  ix->device_bytes += bytes;
  return TSC_OK;
}

int32_t ensure_stage_bytes(Index *ix, size_t bytes) { return ensure_stage(ix, bytes); }

static int32_t refresh_live(Index *ix, cudaStream_t st) {
  if (!ix->live_dirty && ix->live_rows_for == ix->rows) return TSC_OK;
  if (ix->has_deleted || ix->has_filter) {
    TSC_CUDA(cudaMemsetAsync(ix->d_live_count, 0, 8, st));
    combine_live_kernel<<<ix->sm_count * 4, 256, 0, st>>>(
        ix->has_deleted ? ix->d_deleted : nullptr, ix->has_filter ? ix->d_filter : nullptr,
        ix->d_live, ix->mask_words, ix->rows, ix->d_live_count);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    unsigned long long cnt = 0;
    TSC_CUDA(cudaMemcpyAsync(&cnt, ix->d_live_count, 8, cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaStreamSynchronize(st));
    ix->live_rows = cnt;
    ix->live_rows_for = ix->rows;
  }
  ix->live_dirty = false;
  return TSC_OK;
}

// queries already padded to [nq, qld] on the device
static int32_t search_padded(Index *ix, const float *d_q, uint32_t nq, uint32_t k,
                             double threshold, int64_t *d_ids, double *d_dist,
                             uint32_t *d_counts, cudaStream_t st) {
  if (ix->rows == 0) {  // meta.totalVectors == 0 -> const [] (ngh_graph_engine.dart:78)
    TSC_CUDA(cudaMemsetAsync(d_ids, 0xFF, (size_t)nq * k * 8, st));
    TSC_CUDA(cudaMemsetAsync(d_dist, 0xFF, (size_t)nq * k * 8, st));
    TSC_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)nq * 4, st));
    return TSC_OK;
  }
  int32_t rc = refresh_live(ix, st);
  if (rc != TSC_OK) return rc;
  const uint32_t kprime = kprime_for(k);
  TSC_CUDA(cudaEventRecord(ix->ev0, st));
  uint32_t lists = 0;
  const bool use_gemm = nq >= ix->gemm_min_nq && gemm_supported(ix, kprime);
  if (use_gemm)
    rc = launch_gemm(ix, d_q, nq, kprime, ix->d_cand, &lists, nullptr, st);
  else
    rc = launch_scan(ix, d_q, nq, kprime, ix->d_cand, &lists, st);
  if (rc != TSC_OK) return rc;
  rc = launch_select(ix, d_q, nq, k, kprime, ix->d_cand, lists * kprime, threshold, d_ids, d_dist,
                     d_counts, st);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaEventRecord(ix->ev1, st));
  ix->searches++;
  ix->last_path = use_gemm ? 2 : 1;
  ix->last_ms = -1.0;  // resolved lazily from the events
  uint32_t passes = (nq + 7) / 8;
  if (nq <= 4) passes = 1;
  ix->last_gbs = (double)passes * (double)ix->rows * ix->desc.dims * ix->elem_bytes;  // bytes
  return TSC_OK;
}

static int32_t check_search_args(Index *ix, const void *q, uint32_t nq, uint32_t k,
                                 const void *ids, const void *dist, const void *counts) {
  if (!q || !ids || !dist || !counts) {
    set_error("search: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  if (nq == 0 || nq > ix->nq_max) {
    set_error("search: nq=%u outside [1, nq_max=%u]", nq, ix->nq_max);
    return TSC_ERR_BAD_ARG;
  }
  if (k == 0 || k > ix->k_max) {
    set_error("search: k=%u outside [1, k_max=%u]", k, ix->k_max);
    return TSC_ERR_BAD_ARG;
  }
  return TSC_OK;
}

// ---- NCCL (dlopen'ed so that the library loads on hosts without it) ------------
struct Id128 {  // ncclUniqueId: 128 opaque bytes, passed by value
  char b[128];
};
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, Id128, int) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int32_t nccl_load() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_nccl.lib) return TSC_OK;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    set_error("NCCL not found: %s", dlerror());
    return TSC_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (int (*)(void *))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(
      lib, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    set_error("NCCL symbols missing");
    return TSC_ERR_NCCL;
  }
  g_nccl.lib = lib;
  return TSC_OK;
}
#define TSC_NCCL(expr)                                                                  \
  do {                                                                                  \
    int _r = (expr);                                                                    \
    if (_r != 0) {                                                                      \
      set_error("%s failed: %s", #expr,                                                 \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");      \
      return TSC_ERR_NCCL;                                                              \
    }                                                                                   \
  } while (0)

}  // namespace tsc

using namespace tsc;

extern "C" {

int32_t tsc_version(void) { return TSC_ABI_VERSION; }

int32_t tsc_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TSC_ERR_CUDA;
  }
  return n;
}

const char *tsc_last_error(void) { return g_err; }

const char *tsc_status_name(int32_t s) {
  switch (s) {
    case TSC_OK: return "TSC_OK";
    case TSC_ERR_BAD_HANDLE: return "TSC_ERR_BAD_HANDLE";
    case TSC_ERR_BAD_ARG: return "TSC_ERR_BAD_ARG";
    case TSC_ERR_BAD_DIMS: return "TSC_ERR_BAD_DIMS";
    case TSC_ERR_OOM: return "TSC_ERR_OOM";
    case TSC_ERR_CUDA: return "TSC_ERR_CUDA";
    case TSC_ERR_NCCL: return "TSC_ERR_NCCL";
    case TSC_ERR_PAGE: return "TSC_ERR_PAGE";
    case TSC_ERR_UNSUPPORTED: return "TSC_ERR_UNSUPPORTED";
    case TSC_ERR_NOT_READY: return "TSC_ERR_NOT_READY";
    default: return "TSC_ERR_UNKNOWN";
  }
}

int32_t tsc_index_create(const tsc_index_desc *d, uint64_t *out_handle) {
  if (!d || !out_handle) {
    set_error("index_create: NULL argument");
    return TSC_ERR_BAD_ARG;
  }
  if (d->struct_size != sizeof(tsc_index_desc)) {
    set_error("index_create: struct_size %u != %zu", d->struct_size, sizeof(tsc_index_desc));
    return TSC_ERR_BAD_ARG;
  }
  // The reference never validates dims; this boundary does.
  if (d->dims == 0 || d->dims > 65535) {
    set_error("index_create: dims=%u outside [1, 65535]", d->dims);
    return TSC_ERR_BAD_DIMS;
  }
  if (d->metric > TSC_METRIC_COSINE || d->src_precision > TSC_SRC_I8 ||
      d->dev_dtype > TSC_DEV_F16) {
    set_error("index_create: bad metric / precision / dtype code");
    return TSC_ERR_BAD_ARG;
  }
  if (d->capacity_rows == 0 || d->capacity_rows >= 0xFFFFFFFFull) {
    set_error("index_create: capacity_rows must be in [1, 2^32-2] per shard");
    return TSC_ERR_BAD_ARG;
  }
  if (d->k_max == 0 || d->k_max > 128 || d->nq_max == 0 || d->nq_max > 65535) {
    set_error("index_create: k_max must be in [1,128], nq_max in [1,65535]");
    return TSC_ERR_BAD_ARG;
  }
  int ndev = tsc_device_count();
  if (ndev < 0) return ndev;
  if (d->device_id < 0 || d->device_id >= ndev) {
    set_error("index_create: device %d not present (%d devices)", d->device_id, ndev);
    return TSC_ERR_CUDA;
  }
  TSC_CUDA(cudaSetDevice(d->device_id));
  cudaDeviceProp prop;
  TSC_CUDA(cudaGetDeviceProperties(&prop, d->device_id));
  if (prop.major != 10) {
    set_error("index_create: device %d is sm_%d%d; this library is built for sm_100a only",
              d->device_id, prop.major, prop.minor);
    return TSC_ERR_UNSUPPORTED;
  }
  Index *ix = new Index();
  ix->desc = *d;
  ix->device = d->device_id;
  ix->sm_count = prop.multiProcessorCount;
  ix->smem_optin = prop.sharedMemPerBlockOptin;
  ix->elem_bytes = d->dev_dtype == TSC_DEV_F32 ? 4 : 2;
  uint32_t epc = 16 / ix->elem_bytes;
  ix->ld = (d->dims + epc - 1) / epc * epc;
  ix->row_bytes = ix->ld * ix->elem_bytes;
  ix->qld = ix->ld;
  ix->capacity = d->capacity_rows;
  ix->k_max = d->k_max;
  ix->nq_max = d->nq_max;
  ix->kprime_max = kprime_for(d->k_max);
  int32_t rc = scan_configure(ix);
  if (rc != TSC_OK) {
    delete ix;
    return rc;
  }
  // scan: one list per CTA; tensor-core path: two per CTA (<= 2 x SM count)
  ix->cand_lists = 2ull * (uint64_t)(ix->scan.grid > ix->sm_count ? ix->scan.grid : ix->sm_count);
  ix->mask_words = (ix->capacity + 31) / 32 + 1;
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
  };
  ok(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
  ok(cudaEventCreate(&ix->ev0));
  ok(cudaEventCreate(&ix->ev1));
  for (int i = 0; i < Index::kTimers; i++) {
    ok(cudaEventCreate(&ix->t_beg[i]));
    ok(cudaEventCreate(&ix->t_end[i]));
  }
  ok(dev_alloc(ix, &ix->d_rows, (size_t)ix->capacity * ix->row_bytes));
  ok(dev_alloc(ix, &ix->d_deleted, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_filter, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_live, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_delta, 4));
  ok(dev_alloc(ix, &ix->d_live_count, 1));
  if (const char *ev = getenv("TSC_SCAN_SPARSE_FRAC")) ix->sparse_frac = atof(ev);
  ok(dev_alloc(ix, &ix->d_queries, (size_t)ix->nq_max * ix->qld));
  if (d->dev_dtype != TSC_DEV_F32) {
    ok(dev_alloc(ix, &ix->d_norm2, (size_t)ix->capacity));
    ok(dev_alloc(ix, &ix->d_q16, (size_t)ix->nq_max * ix->qld));
    ok(dev_alloc(ix, &ix->d_progress, (size_t)4096));
  } else if (const char *ev = getenv("TSC_GEMM_TF32")) {
    // opt-in (experimental, not yet measured): batches over an fp32 column on the tensor
    // cores as tf32 instead of looping the scan kernel 8 queries at a time
    if (atoi(ev) == 1) {
      ix->tf32 = true;
      ok(dev_alloc(ix, &ix->d_norm2, (size_t)ix->capacity));
    }
  }
  if (const char *ev = getenv("TSC_GEMM_MIN_NQ")) ix->gemm_min_nq = (uint32_t)atoi(ev);
  ok(dev_alloc(ix, &ix->d_cand, (size_t)ix->nq_max * ix->cand_lists * ix->kprime_max));
  ok(dev_alloc(ix, &ix->d_out_ids, (size_t)ix->nq_max * ix->k_max));
  ok(dev_alloc(ix, &ix->d_out_dist, (size_t)ix->nq_max * ix->k_max));
  ok(dev_alloc(ix, &ix->d_out_counts, (size_t)ix->nq_max));
  ok(cudaMallocHost((void **)&ix->h_queries, (size_t)ix->nq_max * ix->qld * 4));
  ok(cudaMallocHost((void **)&ix->h_out_ids, (size_t)ix->nq_max * ix->k_max * 8));
  ok(cudaMallocHost((void **)&ix->h_out_dist, (size_t)ix->nq_max * ix->k_max * 8));
  ok(cudaMallocHost((void **)&ix->h_out_counts, (size_t)ix->nq_max * 4));
  if (e == cudaSuccess) ok(cudaMemsetAsync(ix->d_deleted, 0, ix->mask_words * 4, ix->stream));
  if (e == cudaSuccess) ok(cudaMemsetAsync(ix->d_filter, 0xFF, ix->mask_words * 4, ix->stream));
  if (e == cudaSuccess) ok(cudaStreamSynchronize(ix->stream));
  if (e != cudaSuccess) {
    set_error("index_create: %s (rows need %.2f GB)", cudaGetErrorString(e),
              (double)ix->capacity * ix->row_bytes / 1e9);
    cudaGetLastError();
    free_index(ix);
    return e == cudaErrorMemoryAllocation ? TSC_ERR_OOM : TSC_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next_handle++;
  g_index[h] = ix;
  *out_handle = h;
  return TSC_OK;
}

int32_t tsc_index_destroy(uint64_t handle) {
  Index *ix;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_index.find(handle);
    if (it == g_index.end()) {
      set_error("unknown index handle %llu", (unsigned long long)handle);
      return TSC_ERR_BAD_HANDLE;
    }
    ix = it->second;
    g_index.erase(it);
  }
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ix->nccl_comm);
  }
  free_index(ix);
  return TSC_OK;
}

int32_t tsc_index_clear(uint64_t handle) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  TSC_CUDA(cudaMemsetAsync(ix->d_deleted, 0, ix->mask_words * 4, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_filter, 0xFF, ix->mask_words * 4, ix->stream));
  for (auto &c : ix->columns)
    TSC_CUDA(cudaMemsetAsync(c.d_null, 0xFF, ix->mask_words * 4, ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  ix->rows = 0;
  ix->deleted_rows = 0;
  ix->has_deleted = ix->has_filter = ix->live_dirty = false;
  for (auto &c : ix->columns) c.rows = 0;
  ix->pk_off.clear();
  ix->pk_len.clear();
  ix->pk_arena.clear();
  return TSC_OK;
}

static int32_t check_append_range(Index *ix, uint64_t first_node_id, uint64_t n_rows,
                                  uint64_t *row0) {
  uint64_t base = ix->desc.first_node_id;
  if (first_node_id < base || first_node_id - base > ix->rows) {
    set_error("append: first_node_id %llu is not contiguous with shard [%llu, %llu)",
              (unsigned long long)first_node_id, (unsigned long long)base,
              (unsigned long long)(base + ix->rows));
    return TSC_ERR_BAD_ARG;
  }
  *row0 = first_node_id - base;
  if (*row0 + n_rows > ix->capacity) {
    set_error("append: %llu rows at %llu exceed capacity %llu", (unsigned long long)n_rows,
              (unsigned long long)*row0, (unsigned long long)ix->capacity);
    return TSC_ERR_OOM;
  }
  return TSC_OK;
}

int32_t tsc_index_append_rows(uint64_t handle, uint64_t first_node_id, const void *rows,
                              uint64_t n_rows) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (n_rows == 0) return TSC_OK;
  if (!rows) {
    set_error("append_rows: NULL rows");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  uint64_t row0;
  int32_t rc = check_append_range(ix, first_node_id, n_rows, &row0);
  if (rc != TSC_OK) return rc;
  const uint32_t dims = ix->desc.dims;
  const int prec = ix->desc.src_precision;
  const uint32_t bpe = prec == TSC_SRC_F64 ? 8 : (prec == TSC_SRC_I8 ? 1 : 4);
  const size_t src_row = (size_t)dims * bpe;
  if (prec == TSC_SRC_F32 && ix->desc.dev_dtype == TSC_DEV_F32 && ix->ld == dims) {
    TSC_CUDA(cudaMemcpyAsync(ix->d_rows + row0 * ix->row_bytes, rows, n_rows * src_row,
                             cudaMemcpyHostToDevice, ix->stream));
  } else {
    const uint64_t chunk = (64ull << 20) / src_row ? (64ull << 20) / src_row : 1;
    rc = ensure_stage(ix, (size_t)(chunk < n_rows ? chunk : n_rows) * src_row);
    if (rc != TSC_OK) return rc;
    for (uint64_t r = 0; r < n_rows; r += chunk) {
      uint64_t n = n_rows - r < chunk ? n_rows - r : chunk;
      TSC_CUDA(cudaMemcpyAsync(ix->d_stage, (const uint8_t *)rows + r * src_row, n * src_row,
                               cudaMemcpyHostToDevice, ix->stream));
      convert_rows_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(
          ix->d_stage, n, dims, prec, bpe, ix->d_rows + (row0 + r) * ix->row_bytes, ix->row_bytes,
          ix->ld, ix->desc.dev_dtype);
      TSC_CUDA(cudaGetLastError());
      ix->launches++;
      TSC_CUDA(cudaStreamSynchronize(ix->stream));  // staging buffer is reused
    }
  }
  rc = gemm_update_norms(ix, row0, n_rows, ix->stream);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  if (row0 + n_rows > ix->rows) ix->rows = row0 + n_rows;
  ix->live_dirty = true;
  return TSC_OK;
}

int32_t tsc_index_append_synthetic(uint64_t handle, uint64_t seed, uint64_t first_node_id,
                                   uint64_t n_rows) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (n_rows == 0) return TSC_OK;
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  uint64_t row0;
  int32_t rc = check_append_range(ix, first_node_id, n_rows, &row0);
  if (rc != TSC_OK) return rc;
  synth_rows_kernel<<<ix->sm_count * 16, 256, 0, ix->stream>>>(
      seed, first_node_id, n_rows, ix->desc.dims, ix->d_rows + row0 * ix->row_bytes, ix->row_bytes,
      ix->ld, ix->desc.dev_dtype);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  rc = gemm_update_norms(ix, row0, n_rows, ix->stream);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  if (row0 + n_rows > ix->rows) ix->rows = row0 + n_rows;
  return TSC_OK;
}

static int32_t stage_and_check_pages(Index *ix, const uint8_t *pages, uint64_t n_pages,
                                     uint32_t page_size, uint32_t type, uint64_t page_base) {
  int32_t rc = ensure_stage(ix, (size_t)n_pages * page_size);
  if (rc != TSC_OK) return rc;
  if (ix->page_status_cap < n_pages + 1) {
    cudaFree(ix->d_page_status);
    ix->d_page_status = nullptr;
    TSC_CUDA(cudaMalloc((void **)&ix->d_page_status, (n_pages + 1) * 4));
    ix->page_status_cap = n_pages + 1;
  }
  TSC_CUDA(cudaMemcpyAsync(ix->d_stage, pages, (size_t)n_pages * page_size,
                           cudaMemcpyHostToDevice, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_page_status + n_pages, 0, 4, ix->stream));
  page_check_kernel<<<ix->sm_count * 4, 256, 0, ix->stream>>>(
      ix->d_stage, n_pages, page_size, type, ix->desc.dims, ix->d_page_status,
      ix->d_page_status + n_pages);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  uint32_t bad = 0;
  TSC_CUDA(cudaMemcpyAsync(&bad, ix->d_page_status + n_pages, 4, cudaMemcpyDeviceToHost,
                           ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  if (bad) {
    std::vector<uint32_t> st(n_pages);
    TSC_CUDA(cudaMemcpy(st.data(), ix->d_page_status, n_pages * 4, cudaMemcpyDeviceToHost));
    static const char *why[] = {"ok", "bad magic/header", "bad payload length", "CRC mismatch",
                                "wrong page type", "bad payload", "dims mismatch"};
    for (uint64_t i = 0; i < n_pages; i++)
      if (st[i]) {
        set_error("page %llu: %s (%u bad pages in this call)",
                  (unsigned long long)(page_base + i), why[st[i] < 7 ? st[i] : 0], bad);
        break;
      }
    return TSC_ERR_PAGE;
  }
  return TSC_OK;
}

int32_t tsc_index_append_pages(uint64_t handle, uint64_t first_logical_page, const uint8_t *pages,
                               uint64_t n_pages, uint32_t page_size, uint64_t live_rows) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (n_pages == 0) return TSC_OK;
  if (!pages || page_size < 128) {
    set_error("append_pages: NULL pages or page_size < 128");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  const int prec = ix->desc.src_precision;
  const uint32_t bpe = prec == TSC_SRC_F64 ? 8 : (prec == TSC_SRC_I8 ? 1 : 4);
  // NghPageSizer.vectorsPerRawPage, core/ngh_page.dart:575-579
  const int64_t usable = (int64_t)page_size - 20 - 8 - 64;
  const uint32_t rpp = usable > 0 ? (uint32_t)(usable / ((int64_t)ix->desc.dims * bpe)) : 0;
  if (rpp == 0) {
    set_error("append_pages: dims=%u does not fit a %u-byte page", ix->desc.dims, page_size);
    return TSC_ERR_BAD_DIMS;
  }
  const uint64_t chunk_pages = (64ull << 20) / page_size;
  for (uint64_t p0 = 0; p0 < n_pages; p0 += chunk_pages) {
    uint64_t np = n_pages - p0 < chunk_pages ? n_pages - p0 : chunk_pages;
    int32_t rc = stage_and_check_pages(ix, pages + p0 * page_size, np, page_size, kPtRawVec,
                                       first_logical_page + p0);
    if (rc != TSC_OK) return rc;
    uint64_t node0 = (first_logical_page + p0) * rpp;  // nodeId of slot 0 of this chunk
    uint64_t node1 = node0 + np * rpp;
    if (node1 > live_rows) node1 = live_rows;            // zero tail of the last page
    uint64_t lo = node0 > ix->desc.first_node_id ? node0 : ix->desc.first_node_id;
    uint64_t hi = node1 < ix->desc.first_node_id + ix->capacity
                      ? node1
                      : ix->desc.first_node_id + ix->capacity;
    if (lo >= hi) continue;
    uint64_t row0;
    rc = check_append_range(ix, lo, hi - lo, &row0);
    if (rc != TSC_OK) return rc;
    page_decode_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(
        ix->d_stage, page_size, rpp, lo - node0, hi - lo, ix->desc.dims,
        ix->d_rows + row0 * ix->row_bytes, ix->row_bytes, ix->ld, ix->desc.dev_dtype);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    rc = gemm_update_norms(ix, row0, hi - lo, ix->stream);
    if (rc != TSC_OK) return rc;
    TSC_CUDA(cudaStreamSynchronize(ix->stream));
    if (row0 + (hi - lo) > ix->rows) ix->rows = row0 + (hi - lo);
  }
  return TSC_OK;
}

int32_t tsc_index_set_deleted(uint64_t handle, const uint64_t *node_ids, uint64_t n,
                              uint8_t deleted) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (n == 0) return TSC_OK;
  if (!node_ids) {
    set_error("set_deleted: NULL node_ids");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  int32_t rc = ensure_stage(ix, n * 8);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaMemcpyAsync(ix->d_stage, node_ids, n * 8, cudaMemcpyHostToDevice, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_delta, 0, 4, ix->stream));
  set_bits_kernel<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0,
                    ix->stream>>>((const uint64_t *)ix->d_stage, n, ix->desc.first_node_id,
                                  ix->rows, ix->d_deleted, deleted ? 1 : 0, ix->d_delta);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  int delta = 0;
  TSC_CUDA(cudaMemcpyAsync(&delta, ix->d_delta, 4, cudaMemcpyDeviceToHost, ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  ix->deleted_rows = (uint64_t)((int64_t)ix->deleted_rows + delta);
  ix->has_deleted = ix->deleted_rows > 0;
  ix->live_dirty = true;
  return TSC_OK;
}

int32_t tsc_index_apply_graph_pages(uint64_t handle, uint64_t first_logical_page,
                                    const uint8_t *pages, uint64_t n_pages, uint32_t page_size) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (n_pages == 0) return TSC_OK;
  if (!pages || page_size < 128) {
    set_error("apply_graph_pages: NULL pages or page_size < 128");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint64_t chunk_pages = (64ull << 20) / page_size;
  for (uint64_t p0 = 0; p0 < n_pages; p0 += chunk_pages) {
    uint64_t np = n_pages - p0 < chunk_pages ? n_pages - p0 : chunk_pages;
    int32_t rc = stage_and_check_pages(ix, pages + p0 * page_size, np, page_size, kPtGraph,
                                       first_logical_page + p0);
    if (rc != TSC_OK) return rc;
    // slots per page = NghPageSizer.nodesPerGraphPage (ngh_page.dart:559-566), from the
    // maxDegree the page itself records; a page's own slotCount may be smaller (last page)
    uint16_t hdr[2] = {0, 0};   // [slotCount][maxDegree]
    TSC_CUDA(cudaMemcpy(hdr, ix->d_stage + kPageHeader, 4, cudaMemcpyDeviceToHost));
    const int64_t usable = (int64_t)page_size - 20 - 4 - 64;
    const uint32_t per_page = usable > 0 ? (uint32_t)(usable / (2 + (int64_t)hdr[1] * 4)) : 0;
    if (per_page == 0) {
      set_error("apply_graph_pages: page size %u holds no slot of degree %u", page_size, hdr[1]);
      return TSC_ERR_PAGE;
    }
    uint32_t set_before = 0;
    TSC_CUDA(cudaMemsetAsync(ix->d_delta, 0, 4, ix->stream));
    graph_flags_kernel<<<(unsigned)(np < 2048 ? np : 2048), 128, 0, ix->stream>>>(
        ix->d_stage, np, page_size, per_page, (first_logical_page + p0) * per_page,
        ix->desc.first_node_id, ix->rows, ix->d_deleted, (uint32_t *)ix->d_delta);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    TSC_CUDA(cudaMemcpyAsync(&set_before, ix->d_delta, 4, cudaMemcpyDeviceToHost, ix->stream));
    TSC_CUDA(cudaStreamSynchronize(ix->stream));
    ix->deleted_rows += set_before;
  }
  ix->has_deleted = ix->deleted_rows > 0;
  ix->live_dirty = true;
  return TSC_OK;
}

int32_t tsc_index_set_filter(uint64_t handle, const uint64_t *bitmap_words, uint64_t n_words) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  if (!bitmap_words) {
    ix->has_filter = false;
    ix->live_dirty = true;
    return TSC_OK;
  }
  uint64_t need = (ix->rows + 63) / 64;
  if (n_words < need) {
    set_error("set_filter: %llu words given, %llu needed for %llu rows",
              (unsigned long long)n_words, (unsigned long long)need,
              (unsigned long long)ix->rows);
    return TSC_ERR_BAD_ARG;
  }
  uint64_t words32 = need * 2 < ix->mask_words ? need * 2 : ix->mask_words;
  TSC_CUDA(cudaMemcpyAsync(ix->d_filter, bitmap_words, words32 * 4, cudaMemcpyHostToDevice,
                           ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  ix->has_filter = true;
  ix->live_dirty = true;
  return TSC_OK;
}

// ---- search -------------------------------------------------------------------
int32_t tsc_search_device(uint64_t handle, const float *d_queries, uint32_t nq, uint32_t k,
                          double threshold, int64_t *d_out_ids, double *d_out_dist,
                          uint32_t *d_out_counts, void *cuda_stream) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  int32_t rc = check_search_args(ix, d_queries, nq, k, d_out_ids, d_out_dist, d_out_counts);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  const float *q = d_queries;
  if (ix->qld != ix->desc.dims) {
    pad_queries_kernel<<<(nq * ix->qld + 255) / 256, 256, 0, st>>>(d_queries, nq, ix->desc.dims,
                                                                   ix->d_queries, ix->qld);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    q = ix->d_queries;
  }
  return search_padded(ix, q, nq, k, threshold, d_out_ids, d_out_dist, d_out_counts, st);
}

int32_t tsc_search_submit(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                          double threshold, int64_t *out_ids, double *out_dist,
                          uint32_t *out_counts, uint64_t *out_ticket) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!out_ticket) {
    set_error("search_submit: NULL out_ticket");
    return TSC_ERR_BAD_ARG;
  }
  std::unique_lock<std::mutex> lk(ix->mu);
  int32_t rc = check_search_args(ix, queries, nq, k, out_ids, out_dist, out_counts);
  if (rc != TSC_OK) return rc;
  {
    std::lock_guard<std::mutex> g(g_mu);
    for (auto &kv : g_tickets)
      if (kv.second->ix == ix) {
        set_error("search_submit: a search is already in flight on this index");
        return TSC_ERR_NOT_READY;
      }
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint32_t dims = ix->desc.dims, qld = ix->qld;
  for (uint32_t q = 0; q < nq; q++) {
    memcpy(ix->h_queries + (size_t)q * qld, queries + (size_t)q * dims, (size_t)dims * 4);
    for (uint32_t c = dims; c < qld; c++) ix->h_queries[(size_t)q * qld + c] = 0.0f;
  }
  cudaStream_t st = ix->stream;
  TSC_CUDA(cudaMemcpyAsync(ix->d_queries, ix->h_queries, (size_t)nq * qld * 4,
                           cudaMemcpyHostToDevice, st));
  rc = search_padded(ix, ix->d_queries, nq, k, threshold, ix->d_out_ids, ix->d_out_dist,
                     ix->d_out_counts, st);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaMemcpyAsync(ix->h_out_ids, ix->d_out_ids, (size_t)nq * k * 8,
                           cudaMemcpyDeviceToHost, st));
  TSC_CUDA(cudaMemcpyAsync(ix->h_out_dist, ix->d_out_dist, (size_t)nq * k * 8,
                           cudaMemcpyDeviceToHost, st));
  TSC_CUDA(cudaMemcpyAsync(ix->h_out_counts, ix->d_out_counts, (size_t)nq * 4,
                           cudaMemcpyDeviceToHost, st));
  Ticket *t = new Ticket{ix, handle, nullptr, out_ids, out_dist, out_counts, nq, k};
  cudaError_t e = cudaEventCreateWithFlags(&t->done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(t->done, st);
  if (e != cudaSuccess) {
    set_error("search_submit: %s", cudaGetErrorString(e));
    delete t;
    return TSC_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_mu);
  uint64_t id = g_next_ticket++;
  g_tickets[id] = t;
  *out_ticket = id;
  return TSC_OK;
}

static int32_t finish_ticket(uint64_t ticket, bool block, int32_t *out_done) {
  Ticket *t;
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_tickets.find(ticket);
    if (it == g_tickets.end()) {
      set_error("unknown ticket %llu", (unsigned long long)ticket);
      return TSC_ERR_BAD_HANDLE;
    }
    t = it->second;
  }
  cudaError_t e = block ? cudaEventSynchronize(t->done) : cudaEventQuery(t->done);
  if (e == cudaErrorNotReady) {
    if (out_done) *out_done = 0;
    return TSC_OK;
  }
  {
    std::lock_guard<std::mutex> g(g_mu);
    g_tickets.erase(ticket);
  }
  int32_t rc = TSC_OK;
  if (e != cudaSuccess) {
    set_error("search: %s", cudaGetErrorString(e));
    rc = TSC_ERR_CUDA;
  } else {
    std::lock_guard<std::mutex> lk(t->ix->mu);
    memcpy(t->out_ids, t->ix->h_out_ids, (size_t)t->nq * t->k * 8);
    memcpy(t->out_dist, t->ix->h_out_dist, (size_t)t->nq * t->k * 8);
    memcpy(t->out_counts, t->ix->h_out_counts, (size_t)t->nq * 4);
    if (out_done) *out_done = 1;
  }
  cudaEventDestroy(t->done);
  delete t;
  return rc;
}

int32_t tsc_search_poll(uint64_t ticket, int32_t *out_done) {
  if (!out_done) {
    set_error("search_poll: NULL out_done");
    return TSC_ERR_BAD_ARG;
  }
  return finish_ticket(ticket, false, out_done);
}

int32_t tsc_search_wait(uint64_t ticket) { return finish_ticket(ticket, true, nullptr); }

int32_t tsc_search(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                   double threshold, int64_t *out_ids, double *out_dist, uint32_t *out_counts) {
  uint64_t t = 0;
  int32_t rc = tsc_search_submit(handle, queries, nq, k, threshold, out_ids, out_dist, out_counts,
                                 &t);
  if (rc != TSC_OK) return rc;
  return tsc_search_wait(t);
}

// ---- VectorIndexManager.vectorSearch's arithmetic around the engine call --------------------
// _toFloat32 (vector_index_manager.dart:1385-1392): truncate / zero-pad to dims, fp64 -> fp32
// round to nearest even; cosine: _normalizeFloat32 (:1395-1408), magnitude in fp64 over the
// fp32 values, zero vector unchanged.
static void prep_query_f32(uint32_t dims, int metric, const double *values, uint64_t len,
                           float *q) {
  const uint64_t n = len < dims ? len : dims;
  for (uint64_t i = 0; i < n; i++) q[i] = (float)values[i];
  for (uint64_t i = n; i < dims; i++) q[i] = 0.0f;
  if (metric == TSC_METRIC_COSINE) {
    double mag = 0;
    for (uint32_t i = 0; i < dims; i++) mag += (double)q[i] * (double)q[i];
    mag = sqrt(mag);
    if (mag != 0) {
      double inv = 1.0 / mag;
      for (uint32_t i = 0; i < dims; i++) q[i] = (float)((double)q[i] * inv);
    }
  }
}

// _distanceToScore (:1411-1423)
static double distance_to_score(int metric, double d) {
  if (metric == TSC_METRIC_L2) return 1.0 / (1.0 + d);
  if (metric == TSC_METRIC_INNER_PRODUCT) return 1.0 / (1.0 + exp(d));   // d = -dot
  double s = 1.0 - d;
  if (s == s) s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  return s;
}

int32_t tsc_vector_search(uint64_t handle, const double *values, uint64_t len, uint32_t k,
                          double threshold, int64_t *out_ids, double *out_dist,
                          double *out_score, uint32_t *out_count) {
  return tsc_vector_search_batch(handle, values, len, 1, k, threshold, out_ids, out_dist, out_score,
                                 out_count);
}

// Batch form (additive: the reference's API is single-query): nq query vectors of `len`
// values each, prepared like single queries, searched in one call (the tcgen05 GEMM path for
// 16-bit columns and nq >= 9), scored. out_* are [nq][k], out_counts [nq].
int32_t tsc_vector_search_batch(uint64_t handle, const double *values, uint64_t len, uint32_t nq,
                                uint32_t k, double threshold, int64_t *out_ids, double *out_dist,
                                double *out_score, uint32_t *out_counts) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if ((!values && len) || !out_ids || !out_dist || !out_score || !out_counts || nq == 0) {
    set_error("vector_search: NULL buffer or nq == 0");
    return TSC_ERR_BAD_ARG;
  }
  const uint32_t dims = ix->desc.dims;
  const int metric = ix->desc.metric;
  std::vector<float> q((size_t)nq * dims);
  for (uint32_t i = 0; i < nq; i++)
    prep_query_f32(dims, metric, values ? values + (size_t)i * len : nullptr, len,
                   q.data() + (size_t)i * dims);
  int32_t rc = tsc_search(handle, q.data(), nq, k, threshold, out_ids, out_dist, out_counts);
  if (rc != TSC_OK) return rc;
  for (uint32_t i = 0; i < nq; i++)
    for (uint32_t j = 0; j < k; j++) {
      const size_t o = (size_t)i * k + j;
      out_score[o] = j < out_counts[i] ? distance_to_score(metric, out_dist[o]) : NAN;
    }
  return TSC_OK;
}

// Self-test hooks (no GPU, not fallbacks): the query preparation and the score mapping above,
// so the CPU tier can pin them bit for bit against the oracle's restatement.
int32_t tsc_selftest_query_prep(uint32_t dims, int32_t metric, const double *values, uint64_t len,
                                float *out_f32) {
  if (!out_f32 || dims == 0 || (!values && len)) {
    set_error("selftest_query_prep: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  prep_query_f32(dims, metric, values, len, out_f32);
  return TSC_OK;
}
double tsc_selftest_distance_to_score(int32_t metric, double distance) {
  return distance_to_score(metric, distance);
}

// ---- sharding -------------------------------------------------------------------
int32_t tsc_merge_shards(uint64_t handle, const int64_t *d_part_ids, const double *d_part_dist,
                         uint32_t n_parts, uint32_t nq, uint32_t k, int64_t *d_out_ids,
                         double *d_out_dist, uint32_t *d_out_counts, void *cuda_stream) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!d_part_ids || !d_part_dist || !d_out_ids || !d_out_dist || !d_out_counts || !n_parts ||
      !nq || !k) {
    set_error("merge_shards: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  return launch_merge(ix, d_part_ids, d_part_dist, (uint64_t)nq * k, n_parts, nq, k, d_out_ids,
                      d_out_dist, d_out_counts, st);
}

int32_t tsc_comm_unique_id(uint8_t *out_id128) {
  if (!out_id128) {
    set_error("comm_unique_id: NULL");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = nccl_load();
  if (rc != TSC_OK) return rc;
  TSC_NCCL(g_nccl.GetUniqueId(out_id128));
  return TSC_OK;
}

int32_t tsc_comm_init(uint64_t handle, const uint8_t *id128, int32_t n_ranks, int32_t rank) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) {
    set_error("comm_init: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = nccl_load();
  if (rc != TSC_OK) return rc;
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  Id128 id;
  memcpy(id.b, id128, 128);
  TSC_NCCL(g_nccl.CommInitRank(&ix->nccl_comm, n_ranks, id, rank));
  ix->n_ranks = n_ranks;
  ix->rank = rank;
  size_t part = (size_t)ix->nq_max * ix->k_max * 16;
  TSC_CUDA(dev_alloc(ix, &ix->d_gather_send, part));
  TSC_CUDA(dev_alloc(ix, &ix->d_gather_recv, part * n_ranks));
  return TSC_OK;
}

int32_t tsc_search_sharded(uint64_t handle, const float *d_queries, uint32_t nq, uint32_t k,
                           double threshold, int64_t *d_out_ids, double *d_out_dist,
                           uint32_t *d_out_counts, void *cuda_stream) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!ix->nccl_comm && !ix->p2p_ready) {
    set_error("search_sharded: neither tsc_comm_init nor tsc_comm_p2p_import has been called");
    return TSC_ERR_NCCL;
  }
  // per-shard exact top-k into the send block [ids | dist], then one exchange step
  const size_t nk = (size_t)nq * k;
  int64_t *s_ids = (int64_t *)ix->d_gather_send;
  double *s_dist = (double *)(ix->d_gather_send + nk * 8);
  int32_t rc = tsc_search_device(handle, d_queries, nq, k, threshold, s_ids, s_dist,
                                 d_out_counts, cuda_stream);
  if (rc != TSC_OK) return rc;
  std::lock_guard<std::mutex> lk(ix->mu);
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  if (ix->p2p_ready)   // opt-in: one kernel over NVLink peer memory (tsc_exchange.cuh)
    return launch_exchange(ix, s_ids, s_dist, nq, k, d_out_ids, d_out_dist, d_out_counts, st);
  TSC_NCCL(g_nccl.AllGather(ix->d_gather_send, ix->d_gather_recv, nk * 16, /*ncclUint8*/ 1,
                            ix->nccl_comm, st));
  return launch_merge(ix, (const int64_t *)ix->d_gather_recv,
                      (const double *)(ix->d_gather_recv + nk * 8), nk * 2, (uint32_t)ix->n_ranks,
                      nq, k, d_out_ids, d_out_dist, d_out_counts, st);
}

// ---- opt-in P2P exchange setup (experimental; see tsc_exchange.cuh) --------------------------
// export: allocate this rank's receive buffer and return its CUDA IPC handle (64 bytes);
// the caller all-gathers the handles of all ranks (any host transport) and passes them to
// import, which maps every peer's buffer. One process per GPU (IPC handles cannot be opened
// by the process that made them).
int32_t tsc_comm_p2p_export(uint64_t handle, int32_t n_ranks, int32_t rank, uint8_t *out_ipc64) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!out_ipc64 || n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks) {
    set_error("comm_p2p_export: bad argument (1 <= n_ranks <= 8)");
    return TSC_ERR_BAD_ARG;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lk(ix->mu);
  // one CTA per query spins on its peers' CTA of the same query: every CTA of a launch must
  // be resident at once or two ranks could wait on each other's unscheduled CTAs
  if (ix->nq_max > 512) {
    set_error("comm_p2p_export: nq_max=%u > 512 is not supported by the peer-memory exchange",
              ix->nq_max);
    return TSC_ERR_UNSUPPORTED;
  }
  if (ix->d_xbuf) {
    set_error("comm_p2p_export: already exported");
    return TSC_ERR_BAD_ARG;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint64_t slot = (uint64_t)ix->nq_max * ix->k_max * 16ull;
  const uint64_t bytes = ((2ull * n_ranks * slot + 127ull) & ~127ull) + 2ull * n_ranks * ix->nq_max * 4ull;
  if (bytes > (256ull << 20)) {
    set_error("comm_p2p_export: nq_max * k_max too large for the peer-memory exchange (%llu MB)",
              (unsigned long long)(bytes >> 20));
    return TSC_ERR_UNSUPPORTED;
  }
  TSC_CUDA(cudaMalloc((void **)&ix->d_xbuf, bytes));
  TSC_CUDA(cudaMemset(ix->d_xbuf, 0, bytes));          // flags = 0, epochs start at 1
  TSC_CUDA(cudaHostAlloc((void **)&ix->h_xstatus, 4, cudaHostAllocMapped));
  *ix->h_xstatus = 0;
  TSC_CUDA(cudaHostGetDevicePointer((void **)&ix->d_xstatus, ix->h_xstatus, 0));
  if (!ix->d_gather_send) {
    size_t part = (size_t)ix->nq_max * ix->k_max * 16;
    TSC_CUDA(dev_alloc(ix, &ix->d_gather_send, part));
  }
  ix->xbuf_bytes = bytes;
  ix->xslot_bytes = slot;
  ix->device_bytes += bytes;
  ix->n_ranks = n_ranks;
  ix->rank = rank;
  cudaIpcMemHandle_t hnd;
  TSC_CUDA(cudaIpcGetMemHandle(&hnd, ix->d_xbuf));
  memcpy(out_ipc64, &hnd, 64);
  return TSC_OK;
}

int32_t tsc_comm_p2p_import(uint64_t handle, const uint8_t *all_ipc) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!all_ipc || !ix->d_xbuf) {
    set_error("comm_p2p_import: NULL handles or tsc_comm_p2p_export not called");
    return TSC_ERR_BAD_ARG;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  for (int r = 0; r < ix->n_ranks; r++) {
    if (r == ix->rank) {
      ix->x_peer[r] = ix->d_xbuf;
      continue;
    }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, all_ipc + (size_t)r * 64, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("comm_p2p_import: cannot map rank %d's buffer: %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return TSC_ERR_CUDA;
    }
    ix->x_peer[r] = (uint8_t *)ptr;
  }
  ix->p2p_ready = true;
  return TSC_OK;
}

// Test hook: raw fp32 ranking keys of the tensor-core path for every (query, row),
// so tests can pin the UMMA / TMA / TMEM layouts against a plain matmul.
int32_t tsc_debug_gemm_keys(uint64_t handle, const float *queries, uint32_t nq, float *out_keys) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!queries || !out_keys || nq == 0 || nq > ix->nq_max || ix->rows == 0) {
    set_error("debug_gemm_keys: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  const uint32_t kprime = kprime_for(1);
  if (!gemm_supported(ix, kprime)) {
    set_error("debug_gemm_keys: index has no tensor-core path (fp32 storage)");
    return TSC_ERR_UNSUPPORTED;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint32_t dims = ix->desc.dims, qld = ix->qld;
  for (uint32_t q = 0; q < nq; q++) {
    memcpy(ix->h_queries + (size_t)q * qld, queries + (size_t)q * dims, (size_t)dims * 4);
    for (uint32_t c = dims; c < qld; c++) ix->h_queries[(size_t)q * qld + c] = 0.0f;
  }
  float *d_keys = nullptr;
  TSC_CUDA(cudaMalloc((void **)&d_keys, (size_t)nq * ix->rows * 4));
  cudaStream_t st = ix->stream;
  int32_t rc = refresh_live(ix, st);
  cudaError_t e = cudaMemcpyAsync(ix->d_queries, ix->h_queries, (size_t)nq * qld * 4,
                                  cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_keys, 0xFF, (size_t)nq * ix->rows * 4, st);
  uint32_t lists = 0;
  if (rc == TSC_OK && e == cudaSuccess)
    rc = launch_gemm(ix, ix->d_queries, nq, kprime, ix->d_cand, &lists, d_keys, st);
  if (rc == TSC_OK && e == cudaSuccess)
    e = cudaMemcpyAsync(out_keys, d_keys, (size_t)nq * ix->rows * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_keys);
  if (rc != TSC_OK) return rc;
  if (e != cudaSuccess) {
    set_error("debug_gemm_keys: %s", cudaGetErrorString(e));
    return TSC_ERR_CUDA;
  }
  return TSC_OK;
}

// Host re-enactment of page_check_kernel's warp-sliced CRC-32 (same helpers, 32
// simulated lanes) so the CPU test-suite can pin the slicing algebra without a GPU.
uint32_t tsc_selftest_crc32(const uint8_t *data, uint32_t len) {
  uint32_t tab[256];
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    tab[i] = c;
  }
  uint32_t per = (len + 31) / 32, acc = 0;
  for (uint32_t lane = 0; lane < 32; lane++) {
    uint32_t lo = per * lane < len ? per * lane : len;
    uint32_t hi = per * (lane + 1) < len ? per * (lane + 1) : len;
    uint32_t c = crc_bytes(tab, lane == 0 ? 0xFFFFFFFFu : 0u, data + lo, hi - lo);
    acc ^= crc_shift(c, len - hi);
  }
  return acc ^ 0xFFFFFFFFu;
}

// ---- observability ---------------------------------------------------------------
int32_t tsc_stats_get(uint64_t handle, tsc_stats *out) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!out || out->struct_size != sizeof(tsc_stats)) {
    set_error("stats_get: NULL or struct_size mismatch");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->searches && ix->last_ms < 0) {
    TSC_CUDA(cudaSetDevice(ix->device));
    TSC_CUDA(cudaEventSynchronize(ix->ev1));
    float ms = 0;
    TSC_CUDA(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
    ix->last_ms = ms;
    ix->last_gbs = ms > 0 ? ix->last_gbs / (ms * 1e6) : 0;
  }
  out->dims = ix->desc.dims;
  out->rows = ix->rows;
  out->deleted_rows = ix->deleted_rows;
  out->device_bytes = ix->device_bytes;
  out->row_stride_bytes = ix->row_bytes;
  out->searches = ix->searches;
  out->kernel_launches = ix->launches;
  out->last_search_ms = ix->last_ms < 0 ? 0 : ix->last_ms;
  out->last_scan_gbs = ix->last_ms < 0 ? 0 : ix->last_gbs;
  out->last_path = ix->last_path;
  out->reserved = 0;
  if (ix->t_pending) {
    TSC_CUDA(cudaSetDevice(ix->device));
    int32_t rc = hot_timer_resolve(ix);
    if (rc != TSC_OK) return rc;
  }
  out->hot_launches = ix->hot_launches;
  out->hot_ms_total = ix->hot_ms;
  out->hot_bytes_total = ix->hot_bytes;
  out->hot_flops_total = ix->hot_flops;
  return TSC_OK;
}

int32_t tsc_stats_reset(uint64_t handle) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  int32_t rc = hot_timer_resolve(ix);
  if (rc != TSC_OK) return rc;
  ix->hot_launches = 0;
  ix->hot_ms = ix->hot_bytes = ix->hot_flops = 0;
  return TSC_OK;
}

int32_t tsc_index_device_rows(uint64_t handle, void **out_ptr, uint64_t *out_rows,
                              uint64_t *out_row_stride_bytes) {
  Index *ix = lookup(handle);
  if (!ix) return TSC_ERR_BAD_HANDLE;
  if (!out_ptr || !out_rows || !out_row_stride_bytes) {
    set_error("device_rows: NULL");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  *out_ptr = ix->d_rows;
  *out_rows = ix->rows;
  *out_row_stride_bytes = ix->row_bytes;
  return TSC_OK;
}

}  // extern "C"
