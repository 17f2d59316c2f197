// tsc_api.cu — the C ABI of libtostore_cuda.so (include/tostore_cuda.h): handle registry,
// index lifetime, corpus ingestion, liveness, statistics. Search lives in tsc_search.cu,
// the multi-GPU group in tsc_group.cu.
//
// Host-side mirror of the slice of VectorIndexManager that surrounds the engine
// call (core/vector_index_manager.dart:475-589) plus corpus ingestion. All
// compute runs in the sm_100a kernels of this library; there is no CPU fallback.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include <cuda.h>

#include "tsc_index.h"
#include "tsc_ingest.cuh"
#include "tsc_tail.cuh"

namespace tsc {

// ---- error string -----------------------------------------------------------
static thread_local char g_err[512] = "";
const char *last_error_text() { return g_err; }
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// ---- handle registry ----------------------------------------------------------
// Handles map to shared references: a call that is running keeps its index alive even if
// another thread destroys the handle meanwhile; the device memory goes when the last
// reference does.
static std::mutex g_mu;
static std::unordered_map<uint64_t, IndexRef> g_index;
static std::unordered_map<uint64_t, GroupRef> g_group;
static uint64_t g_next_handle = 1;

IndexRef lookup_index(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_index.find(h);
  if (it == g_index.end()) {
    if (g_group.count(h))
      set_error("handle %llu is a multi-GPU group: this entry point needs a single shard",
                (unsigned long long)h);
    else
      set_error("unknown index handle %llu", (unsigned long long)h);
    return IndexRef();
  }
  return it->second;
}

GroupRef lookup_group(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_group.find(h);
  return it == g_group.end() ? GroupRef() : it->second;
}

uint64_t register_index(IndexRef ix) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next_handle++;
  g_index[h] = ix;
  return h;
}

uint64_t register_group(GroupRef g) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t h = g_next_handle++;
  g_group[h] = g;
  return h;
}

template <typename T>
static cudaError_t dev_alloc(Index *ix, T **p, size_t n) {
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e == cudaSuccess) ix->device_bytes += n * sizeof(T);
  return e;
}

// ---- the row block: virtual memory management ------------------------------------------
// (driver API through cudaGetDriverEntryPoint: the library links the runtime statically)
struct VmmApi {
  CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *,
                     unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*GetGranularity)(size_t *, const CUmemAllocationProp *,
                             CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};
static VmmApi g_vmm;
static std::mutex g_vmm_mu;
static int32_t vmm_load() {
  std::lock_guard<std::mutex> lk(g_vmm_mu);
  if (g_vmm.ok) return TSC_OK;
  auto get = [](const char *name, void **fn) {
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
      set_error("driver entry point %s unavailable: %s", name, cudaGetErrorString(e));
      cudaGetLastError();
      return false;
    }
    return true;
  };
  if (!get("cuMemAddressReserve", (void **)&g_vmm.AddressReserve) ||
      !get("cuMemAddressFree", (void **)&g_vmm.AddressFree) ||
      !get("cuMemCreate", (void **)&g_vmm.Create) || !get("cuMemRelease", (void **)&g_vmm.Release) ||
      !get("cuMemMap", (void **)&g_vmm.Map) || !get("cuMemUnmap", (void **)&g_vmm.Unmap) ||
      !get("cuMemSetAccess", (void **)&g_vmm.SetAccess) ||
      !get("cuMemGetAllocationGranularity", (void **)&g_vmm.GetGranularity))
    return TSC_ERR_CUDA;
  g_vmm.ok = true;
  return TSC_OK;
}
static CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}
// reserve the address range of the row block: as many rows as the device could ever hold
static int32_t rows_reserve(Index *ix) {
  int32_t rc = vmm_load();
  if (rc != TSC_OK) return rc;
  const CUmemAllocationProp prop = vmm_prop(ix->device);
  size_t gran = 0;
  if (g_vmm.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) {
    set_error("index_create: cuMemGetAllocationGranularity failed");
    return TSC_ERR_CUDA;
  }
  size_t free_b = 0, total_b = 0;
  TSC_CUDA(cudaMemGetInfo(&free_b, &total_b));
  size_t want = total_b;
  const size_t all_rows = (size_t)0xFFFFFFFEull * ix->row_bytes;
  if (want > all_rows) want = all_rows;
  const size_t first = (size_t)ix->capacity * ix->row_bytes;
  if (want < first) want = first;
  want = (want + gran - 1) / gran * gran;
  CUdeviceptr va = 0;
  if (g_vmm.AddressReserve(&va, want, 0, 0, 0) != CUDA_SUCCESS) {
    set_error("index_create: cannot reserve %.1f GB of device address space", want / 1e9);
    return TSC_ERR_OOM;
  }
  ix->rows_va = va;
  ix->rows_va_bytes = want;
  ix->vmm_gran = gran;
  ix->d_rows = reinterpret_cast<uint8_t *>(va);
  return TSC_OK;
}
// map physical memory so that bytes [0, need) of the row block are backed
static int32_t rows_map_to(Index *ix, size_t need) {
  if (need <= ix->rows_mapped) return TSC_OK;
  if (need > ix->rows_va_bytes) {
    set_error("the column would need %.1f GB: more than the device has", need / 1e9);
    return TSC_ERR_OOM;
  }
  const size_t add = (need - ix->rows_mapped + ix->vmm_gran - 1) / ix->vmm_gran * ix->vmm_gran;
  const CUmemAllocationProp prop = vmm_prop(ix->device);
  CUmemGenericAllocationHandle h = 0;
  if (g_vmm.Create(&h, add, &prop, 0) != CUDA_SUCCESS) {
    set_error("out of device memory: %.2f GB more for the embedding column", add / 1e9);
    return TSC_ERR_OOM;
  }
  if (g_vmm.Map(ix->rows_va + ix->rows_mapped, add, 0, h, 0) != CUDA_SUCCESS) {
    g_vmm.Release(h);
    set_error("cuMemMap failed for the embedding column");
    return TSC_ERR_CUDA;
  }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = ix->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (g_vmm.SetAccess(ix->rows_va + ix->rows_mapped, add, &acc, 1) != CUDA_SUCCESS) {
    g_vmm.Unmap(ix->rows_va + ix->rows_mapped, add);
    g_vmm.Release(h);
    set_error("cuMemSetAccess failed for the embedding column");
    return TSC_ERR_CUDA;
  }
  ix->rows_chunks.emplace_back((unsigned long long)h, add);
  ix->rows_mapped += add;
  ix->device_bytes += add;
  return TSC_OK;
}
static void rows_release(Index *ix) {
  if (!ix->rows_va || !g_vmm.ok) return;
  size_t off = 0;
  for (auto &c : ix->rows_chunks) {
    g_vmm.Unmap(ix->rows_va + off, c.second);
    g_vmm.Release((CUmemGenericAllocationHandle)c.first);
    off += c.second;
  }
  g_vmm.AddressFree(ix->rows_va, ix->rows_va_bytes);
  ix->rows_chunks.clear();
  ix->rows_va = 0;
  ix->d_rows = nullptr;
}

void free_index(Index *ix) {
  if (ix->host_only) {
    delete ix;
    return;
  }
  cudaSetDevice(ix->device);
  // searches may still be running on caller-owned streams; cudaFree would wait for them
  // implicitly, unmapping the row block (cuMemUnmap) does not
  cudaDeviceSynchronize();
  rows_release(ix);
  cudaFree(ix->d_deleted);
  cudaFree(ix->d_filter);
  cudaFree(ix->d_live);
  cudaFree(ix->d_delta);
  cudaFree(ix->d_live_count);
  cudaFree(ix->d_queries);
  cudaFree(ix->d_norm2);
  cudaFree(ix->d_maxnorm);
  cudaFree(ix->d_q16);
  cudaFree(ix->d_enorm);
  cudaFree(ix->d_progress);
  cudaFree(ix->d_cand);
  cudaFree(ix->d_out_block);
  cudaFree(ix->d_range_thr);
  cudaFree(ix->d_retry_list);
  cudaFree(ix->d_retry_n);
  cudaFree(ix->d_range_count);
  cudaFree(ix->d_range_buf);
  cudaFree(ix->d_done);
  cudaFree(ix->d_cert_stat);
  cudaFree(ix->d_loc_counts);
  cudaFree(ix->d_trace);
  if (ix->host_done) cudaEventDestroy(ix->host_done);
  cudaFree(ix->d_stage);
  cudaFree(ix->d_page_status);
  cudaFree(ix->d_gather_send);
  cudaFree(ix->d_gather_recv);
  if (ix->x_ipc)
    for (int r = 0; r < ix->n_ranks && r < 8; r++)
      if (ix->x_peer[r] && ix->x_peer[r] != ix->d_xbuf) cudaIpcCloseMemHandle(ix->x_peer[r]);
  cudaFree(ix->d_xbuf);
  if (ix->h_xstatus) cudaFreeHost(ix->h_xstatus);
  cudaFree(ix->d_where_args);
  cudaFree(ix->d_where_text);
  for (auto &c : ix->columns) {
    cudaFree(c.d_values);
    cudaFree(c.d_null);
    if (c.dict) {
      cudaFree(c.dict->d_units);
      cudaFree(c.dict->d_offs);
    }
  }
  cudaFreeHost(ix->h_queries);
  cudaFreeHost(ix->h_out_block);
  for (int i = 0; i < Index::kTimers; i++) {
    if (ix->t_beg[i]) cudaEventDestroy(ix->t_beg[i]);
    if (ix->t_end[i]) cudaEventDestroy(ix->t_end[i]);
  }
  if (ix->scratch_ev) cudaEventDestroy(ix->scratch_ev);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  cudaGetLastError();
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
  if (!ix->search_beg) ix->search_beg = ix->t_beg[*slot];
  return TSC_OK;
}

int32_t hot_timer_end(Index *ix, cudaStream_t st, int slot, double bytes, double flops) {
  TSC_CUDA(cudaEventRecord(ix->t_end[slot], st));
  ix->last_hot_end = ix->t_end[slot];
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
    cudaGetLastError();
    return TSC_ERR_OOM;
  }
  ix->stage_bytes = bytes;
  ix->device_bytes += bytes;
  return TSC_OK;
}

int32_t ensure_stage_bytes(Index *ix, size_t bytes) { return ensure_stage(ix, bytes); }

// The search scratch is one per index: a search on another stream waits for the previous one.
// Pipelined searches record no event of their own; the mark is then taken here, on the stream
// the previous search ran on (which must still exist).
int32_t order_after_last_search(Index *ix, cudaStream_t st) {
  if (!ix->scratch_used || ix->scratch_stream == st) return TSC_OK;
  if (!ix->scratch_mark) {
    TSC_CUDA(cudaEventRecord(ix->scratch_ev, ix->scratch_stream));
    ix->scratch_mark = ix->scratch_ev;
  }
  TSC_CUDA(cudaStreamWaitEvent(st, ix->scratch_mark, 0));
  return TSC_OK;
}
int32_t sync_last_search(Index *ix) {
  if (!ix->scratch_used) return TSC_OK;
  if (ix->scratch_mark) TSC_CUDA(cudaEventSynchronize(ix->scratch_mark));
  else TSC_CUDA(cudaStreamSynchronize(ix->scratch_stream));
  return TSC_OK;
}

int32_t refresh_live(Index *ix, cudaStream_t st) {
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

// [nq, dims] caller queries -> ix->d_queries [nq, qld], zero padded
int32_t launch_pad_queries(Index *ix, const float *d_queries, uint32_t nq, cudaStream_t st) {
  pad_queries_kernel<<<(nq * ix->qld + 255) / 256, 256, 0, st>>>(d_queries, nq, ix->desc.dims,
                                                                 ix->d_queries, ix->qld);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

// ---- index lifetime ---------------------------------------------------------------------
int32_t ix_create(const tsc_index_desc *d, IndexRef *out) {
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
  ix->desc.n_devices = 1;
  ix->device = d->device_id;
  ix->sm_count = prop.multiProcessorCount;
  ix->smem_optin = prop.sharedMemPerBlockOptin;
  ix->elem_bytes = d->dev_dtype == TSC_DEV_F32 ? 4 : 2;
  ix->gemm_min_nq = d->dev_dtype == TSC_DEV_F32 ? 9 : 5;
  uint32_t epc = 16 / ix->elem_bytes;
  ix->ld = (d->dims + epc - 1) / epc * epc;
  ix->row_bytes = ix->ld * ix->elem_bytes;
  ix->qld = ix->ld;
  ix->capacity = d->capacity_rows;
  ix->k_max = d->k_max;
  ix->nq_max = d->nq_max;
  ix->kprime_max = kprime_for(d->k_max) > 32 ? kprime_for(d->k_max) : 32;   // tensor path: K' >= 32
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
  ok(cudaEventCreate(&ix->scratch_ev));
  for (int i = 0; i < Index::kTimers; i++) {
    ok(cudaEventCreate(&ix->t_beg[i]));
    ok(cudaEventCreate(&ix->t_end[i]));
  }
  if (e == cudaSuccess) {
    int32_t vrc = rows_reserve(ix);
    if (vrc == TSC_OK) vrc = rows_map_to(ix, (size_t)ix->capacity * ix->row_bytes);
    if (vrc != TSC_OK) {
      free_index(ix);
      return vrc;
    }
  }
  ok(dev_alloc(ix, &ix->d_deleted, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_filter, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_live, ix->mask_words));
  ok(dev_alloc(ix, &ix->d_delta, 4));
  ok(dev_alloc(ix, &ix->d_live_count, 1));
  ok(dev_alloc(ix, &ix->d_queries, (size_t)ix->nq_max * ix->qld));
  // |row|^2 per row: tensor-path coefficients (16-bit: kind::f16, fp32: kind::tf32) and the
  // certificate's norm bound
  ok(dev_alloc(ix, &ix->d_norm2, (size_t)ix->capacity));
  ok(dev_alloc(ix, &ix->d_maxnorm, 1));
  ok(dev_alloc(ix, &ix->d_enorm, (size_t)ix->nq_max));
  ok(dev_alloc(ix, &ix->d_progress, (size_t)4096));
  if (d->dev_dtype != TSC_DEV_F32) ok(dev_alloc(ix, &ix->d_q16, (size_t)ix->nq_max * ix->qld));
  ok(dev_alloc(ix, &ix->d_cand, (size_t)ix->nq_max * ix->cand_lists * ix->kprime_max));
  {
    const size_t nk8 = (size_t)ix->nq_max * ix->k_max * 8, n4 = ((size_t)ix->nq_max * 4 + 7) & ~(size_t)7;
    ix->out_block_bytes = 2 * nk8 + 2 * n4;
    ok(dev_alloc(ix, &ix->d_out_block, ix->out_block_bytes));
    ok(cudaMallocHost((void **)&ix->h_out_block, ix->out_block_bytes));
    if (e == cudaSuccess) {
      ix->d_out_ids = reinterpret_cast<int64_t *>(ix->d_out_block);
      ix->d_out_dist = reinterpret_cast<double *>(ix->d_out_block + nk8);
      ix->d_out_counts = reinterpret_cast<uint32_t *>(ix->d_out_block + 2 * nk8);
      ix->d_flags = reinterpret_cast<uint32_t *>(ix->d_out_block + 2 * nk8 + n4);
      ix->h_out_ids = reinterpret_cast<int64_t *>(ix->h_out_block);
      ix->h_out_dist = reinterpret_cast<double *>(ix->h_out_block + nk8);
      ix->h_out_counts = reinterpret_cast<uint32_t *>(ix->h_out_block + 2 * nk8);
      ix->h_flags = reinterpret_cast<uint32_t *>(ix->h_out_block + 2 * nk8 + n4);
    }
  }
  ok(dev_alloc(ix, &ix->d_range_thr, (size_t)ix->nq_max));
  ok(dev_alloc(ix, &ix->d_retry_list, (size_t)ix->nq_max));
  ok(dev_alloc(ix, &ix->d_retry_n, 1));
  ok(dev_alloc(ix, &ix->d_range_count, (size_t)kRangeSlots));
  ok(dev_alloc(ix, &ix->d_range_buf, (size_t)kRangeSlots * kRangeCap));
  ok(dev_alloc(ix, &ix->d_done, 8));
  ok(dev_alloc(ix, &ix->d_cert_stat, (size_t)kStatSlots));
  ok(dev_alloc(ix, &ix->d_loc_counts, (size_t)ix->nq_max));
#ifdef TSC_DIAG
  ok(dev_alloc(ix, &ix->d_trace, (size_t)4096));
#endif
  ok(cudaEventCreateWithFlags(&ix->host_done, cudaEventDisableTiming));
  ok(cudaMallocHost((void **)&ix->h_queries, (size_t)ix->nq_max * ix->qld * 4));
  if (e == cudaSuccess) {
    ok(cudaMemsetAsync(ix->d_deleted, 0, ix->mask_words * 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_filter, 0xFF, ix->mask_words * 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_maxnorm, 0, 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_flags, 0, (size_t)ix->nq_max * 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_retry_n, 0, 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_range_count, 0, kRangeSlots * 4, ix->stream));
    ok(cudaMemsetAsync(ix->d_done, 0, 32, ix->stream));
    ok(cudaMemsetAsync(ix->d_cert_stat, 0, kStatSlots * 8, ix->stream));
    ok(cudaStreamSynchronize(ix->stream));
  }
  if (e != cudaSuccess) {
    set_error("index_create: %s (rows need %.2f GB)", cudaGetErrorString(e),
              (double)ix->capacity * ix->row_bytes / 1e9);
    cudaGetLastError();
    free_index(ix);
    return e == cudaErrorMemoryAllocation ? TSC_ERR_OOM : TSC_ERR_CUDA;
  }
  *out = IndexRef(ix, free_index);
  return TSC_OK;
}

int32_t ix_clear(Index *ix) {
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  TSC_CUDA(cudaMemsetAsync(ix->d_deleted, 0, ix->mask_words * 4, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_filter, 0xFF, ix->mask_words * 4, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_maxnorm, 0, 4, ix->stream));
  for (auto &c : ix->columns)
    TSC_CUDA(cudaMemsetAsync(c.d_null, 0xFF, ix->mask_words * 4, ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  ix->rows = 0;
  ix->max_norm2 = 0.0f;
  ix->deleted_rows = 0;
  ix->has_deleted = ix->has_filter = ix->live_dirty = false;
  for (auto &c : ix->columns) {
    c.rows = 0;
    if (c.dict) {   // device arrays are kept for the next strings
      c.dict->truncate(0);
      c.dict->n_codes = 0;
      c.dict->n_units = 0;
    }
  }
  ix->pk_off.clear();
  ix->pk_len.clear();
  ix->pk_arena.clear();
  ix->pk_rev.clear();
  ix->pk_rev_valid = false;
  return TSC_OK;
}

// Grow the column of an unsharded handle so that it holds rows [0, rows_needed): more
// physical memory behind the row block (no copy), the small per-row arrays (norms, bitmaps,
// attribute columns) re-allocated and copied. The reference grows by adding partition files
// (model/ngh_index_meta.dart:178-232); a shard's capacity is its node-id range and stays fixed.
template <typename T>
static int32_t regrow(Index *ix, T **p, size_t old_n, size_t new_n, int fill_byte) {
  T *q = nullptr;
  cudaError_t e = cudaMalloc((void **)&q, new_n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("out of device memory growing the column (%.2f GB)", new_n * sizeof(T) / 1e9);
    cudaGetLastError();
    return TSC_ERR_OOM;
  }
  TSC_CUDA(cudaMemcpyAsync(q, *p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, ix->stream));
  TSC_CUDA(cudaMemsetAsync(q + old_n, fill_byte, (new_n - old_n) * sizeof(T), ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  cudaFree(*p);
  *p = q;
  ix->device_bytes += (new_n - old_n) * sizeof(T);
  return TSC_OK;
}

int32_t ix_ensure_capacity(Index *ix, uint64_t rows_needed, const char *what) {
  if (rows_needed <= ix->capacity) return TSC_OK;
  if (ix->in_group || ix->p2p_ready || ix->nccl_comm || ix->host_only) {
    set_error("%s: rows up to %llu exceed the shard's capacity %llu (a shard's capacity is its "
              "node-id range)", what, (unsigned long long)rows_needed,
              (unsigned long long)ix->capacity);
    return TSC_ERR_OOM;
  }
  if (rows_needed >= 0xFFFFFFFFull) {
    set_error("%s: a shard holds at most 2^32-2 rows", what);
    return TSC_ERR_OOM;
  }
  uint64_t cap = ix->capacity + ix->capacity / 2;
  if (cap < rows_needed) cap = rows_needed;
  if (cap > 0xFFFFFFFEull) cap = 0xFFFFFFFEull;
  // if 1.5 x does not fit in memory any more, what is needed may still do
  for (int attempt = 0; attempt < 2; attempt++) {
    if ((size_t)cap * ix->row_bytes <= ix->rows_va_bytes) break;
    cap = rows_needed;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  int32_t rc = sync_last_search(ix);   // nothing may be reading the arrays that move
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  rc = rows_map_to(ix, (size_t)cap * ix->row_bytes);
  if (rc != TSC_OK && cap > rows_needed) {
    cap = rows_needed;
    rc = rows_map_to(ix, (size_t)cap * ix->row_bytes);
  }
  if (rc != TSC_OK) return rc;
  const uint64_t old_words = ix->mask_words, new_words = (cap + 31) / 32 + 1;
  if ((rc = regrow(ix, &ix->d_norm2, (size_t)ix->capacity, (size_t)cap, 0)) != TSC_OK) return rc;
  if ((rc = regrow(ix, &ix->d_deleted, (size_t)old_words, (size_t)new_words, 0)) != TSC_OK) return rc;
  if ((rc = regrow(ix, &ix->d_filter, (size_t)old_words, (size_t)new_words, 0xFF)) != TSC_OK) return rc;
  if ((rc = regrow(ix, &ix->d_live, (size_t)old_words, (size_t)new_words, 0xFF)) != TSC_OK) return rc;
  for (auto &c : ix->columns) {
    if ((rc = regrow(ix, &c.d_values, (size_t)ix->capacity, (size_t)cap, 0)) != TSC_OK) return rc;
    if ((rc = regrow(ix, &c.d_null, (size_t)old_words, (size_t)new_words, 0xFF)) != TSC_OK) return rc;
  }
  ix->capacity = cap;
  ix->desc.capacity_rows = cap;
  ix->mask_words = new_words;
  ix->live_dirty = true;
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
  return ix_ensure_capacity(ix, *row0 + n_rows, "append");
}

int32_t ix_append_rows(Index *ix, uint64_t first_node_id, const void *rows, uint64_t n_rows) {
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
    // two halves of the staging buffer alternate: the copy of chunk i+1 (a pageable-memory
    // cudaMemcpyAsync returns once the bytes are staged) overlaps the conversion of chunk i;
    // an event per half replaces the stream-wide sync per chunk
    uint64_t chunk = (32ull << 20) / src_row ? (32ull << 20) / src_row : 1;
    if (chunk > n_rows) chunk = n_rows;
    rc = ensure_stage(ix, (size_t)chunk * src_row * 2);
    if (rc != TSC_OK) return rc;
    cudaEvent_t done[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) TSC_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    cudaError_t e = cudaSuccess;
    int half = 0;
    for (uint64_t r = 0; r < n_rows && e == cudaSuccess; r += chunk, half ^= 1) {
      uint64_t n = n_rows - r < chunk ? n_rows - r : chunk;
      uint8_t *stage = ix->d_stage + (size_t)half * chunk * src_row;
      if (r >= 2 * chunk) e = cudaEventSynchronize(done[half]);   // the half is free again
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(stage, (const uint8_t *)rows + r * src_row, n * src_row,
                            cudaMemcpyHostToDevice, ix->stream);
      if (e != cudaSuccess) break;
      convert_rows_kernel<<<ix->sm_count * 8, 256, 0, ix->stream>>>(
          stage, n, dims, prec, bpe, ix->d_rows + (row0 + r) * ix->row_bytes, ix->row_bytes, ix->ld,
          ix->desc.dev_dtype);
      e = cudaGetLastError();
      ix->launches++;
      if (e == cudaSuccess) e = cudaEventRecord(done[half], ix->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    for (int i = 0; i < 2; i++) cudaEventDestroy(done[i]);
    if (e != cudaSuccess) {
      set_error("append_rows: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TSC_ERR_CUDA;
    }
  }
  rc = gemm_update_norms(ix, row0, n_rows, ix->stream);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  if (row0 + n_rows > ix->rows) ix->rows = row0 + n_rows;
  ix->live_dirty = true;
  return TSC_OK;
}

int32_t ix_append_synthetic(Index *ix, uint64_t seed, uint64_t first_node_id, uint64_t n_rows) {
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
  ix->live_dirty = true;
  return TSC_OK;
}

static int32_t stage_and_check_pages(Index *ix, const uint8_t *pages, uint64_t n_pages,
                                     uint32_t page_size, uint32_t type, uint64_t page_base) {
  int32_t rc = ensure_stage(ix, (size_t)n_pages * page_size);
  if (rc != TSC_OK) return rc;
  if (ix->page_status_cap < n_pages + 1) {
    cudaFree(ix->d_page_status);
    ix->d_page_status = nullptr;
    ix->page_status_cap = 0;
    TSC_CUDA(cudaMalloc((void **)&ix->d_page_status, (n_pages + 1) * 4));
    ix->page_status_cap = n_pages + 1;
  }
  TSC_CUDA(cudaMemcpyAsync(ix->d_stage, pages, (size_t)n_pages * page_size,
                           cudaMemcpyHostToDevice, ix->stream));
  TSC_CUDA(cudaMemsetAsync(ix->d_page_status + n_pages, 0, 4, ix->stream));
  page_check_kernel<<<ix->sm_count * 4, 256, 0, ix->stream>>>(
      ix->d_stage, n_pages, page_size, type, ix->desc.dims, ix->desc.src_precision,
      ix->d_page_status, ix->d_page_status + n_pages);
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
                                "wrong page type", "bad payload", "dims mismatch",
                                "precision differs from the index's src_precision"};
    for (uint64_t i = 0; i < n_pages; i++)
      if (st[i]) {
        set_error("page %llu: %s (%u bad pages in this call)",
                  (unsigned long long)(page_base + i), why[st[i] < 8 ? st[i] : 0], bad);
        break;
      }
    return TSC_ERR_PAGE;
  }
  return TSC_OK;
}

// pages staged per round trip (at least one, whatever the page size)
static uint64_t pages_per_chunk(uint32_t page_size) {
  uint64_t n = (64ull << 20) / page_size;
  return n ? n : 1;
}

int32_t ix_append_pages(Index *ix, uint64_t first_logical_page, const uint8_t *pages,
                        uint64_t n_pages, uint32_t page_size, uint64_t live_rows) {
  if (n_pages == 0) return TSC_OK;
  if (!pages || page_size < 128 || page_size > (1u << 30)) {
    set_error("append_pages: NULL pages or page_size outside [128, 2^30]");
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
  const uint64_t chunk_pages = pages_per_chunk(page_size);
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
    ix->live_dirty = true;
  }
  return TSC_OK;
}

int32_t ix_set_deleted(Index *ix, const uint64_t *node_ids, uint64_t n, uint8_t deleted) {
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

int32_t ix_apply_graph_pages(Index *ix, uint64_t first_logical_page, const uint8_t *pages,
                             uint64_t n_pages, uint32_t page_size) {
  if (n_pages == 0) return TSC_OK;
  if (!pages || page_size < 128 || page_size > (1u << 30)) {
    set_error("apply_graph_pages: NULL pages or page_size outside [128, 2^30]");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint64_t chunk_pages = pages_per_chunk(page_size);
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

int32_t ix_set_filter(Index *ix, const uint64_t *bitmap_words, uint64_t n_words) {
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

// ---- observability ---------------------------------------------------------------
int32_t ix_stats_get(Index *ix, tsc_stats *out) {
  std::lock_guard<std::mutex> lk(ix->mu);
  memset(reinterpret_cast<char *>(out) + 4, 0, sizeof(*out) - 4);
  out->dims = ix->desc.dims;
  out->rows = ix->rows;
  out->n_devices = 1;
  if (ix->host_only) return TSC_OK;
  if (ix->searches && ix->last_ms < 0) {
    TSC_CUDA(cudaSetDevice(ix->device));
    float ms = 0;
    if (ix->timed_beg && ix->timed_end) {
      TSC_CUDA(cudaEventSynchronize(ix->timed_end));
      TSC_CUDA(cudaEventElapsedTime(&ms, ix->timed_beg, ix->timed_end));
    }
    ix->last_ms = ms;
    ix->last_gbs = ms > 0 ? ix->last_gbs / (ms * 1e6) : 0;
  }
  out->deleted_rows = ix->deleted_rows;
  out->device_bytes = ix->device_bytes;
  out->row_stride_bytes = ix->row_bytes;
  out->searches = ix->searches;
  out->kernel_launches = ix->launches;
  out->last_search_ms = ix->last_ms < 0 ? 0 : ix->last_ms;
  out->last_scan_gbs = ix->last_ms < 0 ? 0 : ix->last_gbs;
  out->last_path = ix->last_path;
  if (ix->last_path == 2 && ix->last_ms > 0) {
    out->last_tflops = ix->last_flops / (ix->last_ms * 1e9);
    out->last_tensor_util = out->last_tflops / (ix->desc.dev_dtype == TSC_DEV_F32 ? 1125.0 : 2250.0);
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  if (ix->t_pending) {
    int32_t rc = hot_timer_resolve(ix);
    if (rc != TSC_OK) return rc;
  }
  out->hot_launches = ix->hot_launches;
  out->hot_ms_total = ix->hot_ms;
  out->hot_bytes_total = ix->hot_bytes;
  out->hot_flops_total = ix->hot_flops;
  unsigned long long cs[kStatSlots];
  TSC_CUDA(cudaMemcpyAsync(cs, ix->d_cert_stat, sizeof cs, cudaMemcpyDeviceToHost, ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  out->certified_queries = cs[kStatCertified];
  out->retried_queries = cs[kStatRetried];
  // asked for a range pass and never got one (device-buffer searches beyond the in-stream
  // range launches) counts as uncertified too
  const unsigned long long done = cs[kStatRetried] + cs[kStatUncertified];
  out->uncertified_queries =
      cs[kStatUncertified] + (cs[kStatRetryAsked] > done ? cs[kStatRetryAsked] - done : 0);
  out->range_rows = cs[kStatRangeRows];
  return TSC_OK;
}

int32_t ix_stats_reset(Index *ix) {
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->host_only) return TSC_OK;
  TSC_CUDA(cudaSetDevice(ix->device));
  int32_t rc = hot_timer_resolve(ix);
  if (rc != TSC_OK) return rc;
  ix->hot_launches = 0;
  ix->hot_ms = ix->hot_bytes = ix->hot_flops = 0;
  TSC_CUDA(cudaMemsetAsync(ix->d_cert_stat, 0, kStatSlots * 8, ix->stream));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  return TSC_OK;
}

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
  TSC_API_TRY
  if (!d || !out_handle) {
    set_error("index_create: NULL argument");
    return TSC_ERR_BAD_ARG;
  }
  if (d->struct_size != sizeof(tsc_index_desc)) {
    set_error("index_create: struct_size %u != %zu", d->struct_size, sizeof(tsc_index_desc));
    return TSC_ERR_BAD_ARG;
  }
  if (d->n_devices > 8) {
    set_error("index_create: n_devices=%u outside [0, 8]", d->n_devices);
    return TSC_ERR_BAD_ARG;
  }
  if (d->n_devices > 1) return grp_create(d, out_handle);
  tsc_index_desc one = *d;
  if (d->n_devices == 1) one.device_id = d->device_ids[0];
  IndexRef ix;
  int32_t rc = ix_create(&one, &ix);
  if (rc != TSC_OK) return rc;
  *out_handle = register_index(ix);
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_index_destroy(uint64_t handle) {
  TSC_API_TRY
  IndexRef ix;
  GroupRef g;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_index.find(handle);
    if (it != g_index.end()) {
      ix = it->second;
      g_index.erase(it);
    } else {
      auto ig = g_group.find(handle);
      if (ig == g_group.end()) {
        set_error("unknown index handle %llu", (unsigned long long)handle);
        return TSC_ERR_BAD_HANDLE;
      }
      g = ig->second;
      g_group.erase(ig);
    }
  }
  // tickets of this handle die with it (tsc_search.cu); calls still running on another
  // thread hold their own reference — the memory goes with the last one
  drop_tickets_of(handle);
  if (ix) {
    std::lock_guard<std::mutex> lk(ix->mu);   // wait for a call in progress
    ix->inflight.reset();                     // the ticket holds a reference to the index
    comm_release(ix.get());
  }
  if (g) {
    std::lock_guard<std::mutex> lk(g->mu);
    g->inflight.reset();
  }
  return TSC_OK;
  TSC_API_CATCH
}

#define TSC_DISPATCH(handle, grp_call, ix_call)         \
  TSC_API_TRY                                           \
  if (GroupRef g = lookup_group(handle)) {              \
    std::lock_guard<std::mutex> glk(g->mu);             \
    return grp_call;                                    \
  }                                                     \
  IndexRef ref = lookup_index(handle);                  \
  if (!ref) return TSC_ERR_BAD_HANDLE;                  \
  Index *ix = ref.get();                                \
  return ix_call;                                       \
  TSC_API_CATCH

static int32_t no_device(Index *ix) {
  if (!ix->host_only) return TSC_OK;
  set_error("this handle is a host-only self-test object: no device entry point works on it");
  return TSC_ERR_UNSUPPORTED;
}
#define TSC_DEV(call) (no_device(ix) != TSC_OK ? TSC_ERR_UNSUPPORTED : (call))

int32_t tsc_index_clear(uint64_t handle) {
  TSC_DISPATCH(handle, grp_clear(*g), TSC_DEV(ix_clear(ix)))
}

int32_t tsc_index_append_rows(uint64_t handle, uint64_t first_node_id, const void *rows,
                              uint64_t n_rows) {
  TSC_DISPATCH(handle, grp_append_rows(*g, first_node_id, rows, n_rows),
               TSC_DEV(ix_append_rows(ix, first_node_id, rows, n_rows)))
}

int32_t tsc_index_append_synthetic(uint64_t handle, uint64_t seed, uint64_t first_node_id,
                                   uint64_t n_rows) {
  TSC_DISPATCH(handle, grp_append_synthetic(*g, seed, first_node_id, n_rows),
               TSC_DEV(ix_append_synthetic(ix, seed, first_node_id, n_rows)))
}

int32_t tsc_index_append_pages(uint64_t handle, uint64_t first_logical_page, const uint8_t *pages,
                               uint64_t n_pages, uint32_t page_size, uint64_t live_rows) {
  TSC_DISPATCH(handle, grp_append_pages(*g, first_logical_page, pages, n_pages, page_size, live_rows),
               TSC_DEV(ix_append_pages(ix, first_logical_page, pages, n_pages, page_size, live_rows)))
}

int32_t tsc_index_set_deleted(uint64_t handle, const uint64_t *node_ids, uint64_t n,
                              uint8_t deleted) {
  TSC_DISPATCH(handle, grp_set_deleted(*g, node_ids, n, deleted),
               TSC_DEV(ix_set_deleted(ix, node_ids, n, deleted)))
}

int32_t tsc_index_apply_graph_pages(uint64_t handle, uint64_t first_logical_page,
                                    const uint8_t *pages, uint64_t n_pages, uint32_t page_size) {
  TSC_DISPATCH(handle, grp_apply_graph_pages(*g, first_logical_page, pages, n_pages, page_size),
               TSC_DEV(ix_apply_graph_pages(ix, first_logical_page, pages, n_pages, page_size)))
}

int32_t tsc_index_set_filter(uint64_t handle, const uint64_t *bitmap_words, uint64_t n_words) {
  TSC_DISPATCH(handle, grp_set_filter(*g, bitmap_words, n_words),
               TSC_DEV(ix_set_filter(ix, bitmap_words, n_words)))
}

int32_t tsc_stats_get(uint64_t handle, tsc_stats *out) {
  if (!out || out->struct_size != sizeof(tsc_stats)) {
    set_error("stats_get: NULL or struct_size mismatch");
    return TSC_ERR_BAD_ARG;
  }
  TSC_DISPATCH(handle, grp_stats_get(*g, out), ix_stats_get(ix, out))
}

// Pipelined device-buffer searches: see include/tostore_cuda.h
static int32_t ix_set_pipelining(Index *ix, int32_t on) {
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->host_only) {
    set_error("set_pipelining: host-only self-test handle");
    return TSC_ERR_UNSUPPORTED;
  }
  ix->pipeline = on != 0;
  ix->timer_every = on != 0 ? 16 : 1;
  ix->timer_tick = 0;
  return TSC_OK;
}
static int32_t grp_set_pipelining(Group &g, int32_t on) {
  for (auto &s : g.shards) {
    int32_t rc = ix_set_pipelining(s.get(), on);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}
int32_t tsc_index_set_pipelining(uint64_t handle, int32_t on) {
  TSC_DISPATCH(handle, grp_set_pipelining(*g, on), ix_set_pipelining(ix, on))
}

int32_t tsc_stats_reset(uint64_t handle) {
  TSC_DISPATCH(handle, grp_stats_reset(*g), ix_stats_reset(ix))
}

int32_t tsc_index_device_rows(uint64_t handle, void **out_ptr, uint64_t *out_rows,
                              uint64_t *out_row_stride_bytes) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  if (!out_ptr || !out_rows || !out_row_stride_bytes) {
    set_error("device_rows: NULL");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  *out_ptr = ix->d_rows;
  *out_rows = ix->rows;
  *out_row_stride_bytes = ix->row_bytes;
  return TSC_OK;
  TSC_API_CATCH
}

#ifdef TSC_DIAG
// diagnostics build only: phase timestamps of the last scan launch (tsc_tail.cuh trace slots);
// reset = 1 arms the next launch (slot 0 is taken with atomicMin)
int32_t tsc_diag_scan_trace(uint64_t handle, unsigned long long *out, uint32_t n, int32_t reset) {
  IndexRef ref = lookup_index(handle);
  if (!ref || !out || n > 4096) return TSC_ERR_BAD_ARG;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  TSC_CUDA(cudaStreamSynchronize(ix->stream));
  TSC_CUDA(cudaMemcpy(out, ix->d_trace, (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (reset) TSC_CUDA(cudaMemset(ix->d_trace, 0xFF, 4096 * 8));
  return TSC_OK;
}
#endif

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

}  // extern "C"
