// tsc_search.cu — the search entry points of the C ABI (include/tostore_cuda.h): launch
// sequencing of one search, host-buffer searches and tickets, the arithmetic around the
// engine call, and the communicator set-up for sharded indexes.
//
// One search = candidate stage (scan or tensor path) -> tail (select, exact fp64 re-rank,
// certificate) -> range pass for uncertified queries -> (sharded) exchange. For nq <= 8 on
// the scan path all of it is ONE kernel plus one range launch that exits at once when every
// query was certified.
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "tsc_index.h"
#include "tsc_exchange.cuh"

namespace tsc {

struct Ticket {
  uint64_t handle = 0;
  IndexRef ix;      // exactly one of ix / grp is set
  GroupRef grp;
  int64_t *out_ids = nullptr;
  double *out_dist = nullptr;
  uint32_t *out_counts = nullptr;
  uint32_t nq = 0, k = 0;
  bool retired = false;   // results delivered to out_*; rc is final
  int32_t rc = TSC_OK;
};
typedef std::shared_ptr<Ticket> TicketRef;
static std::mutex g_tmu;
static std::unordered_map<uint64_t, TicketRef> g_tickets;
static uint64_t g_next_ticket = 1;

void drop_tickets_of(uint64_t handle) {
  std::lock_guard<std::mutex> lk(g_tmu);
  for (auto it = g_tickets.begin(); it != g_tickets.end();)
    it = it->second->handle == handle ? g_tickets.erase(it) : std::next(it);
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
static std::mutex g_nccl_mu;
static int32_t nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
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

void comm_release(Index *ix) {
  if (ix->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ix->nccl_comm);
  ix->nccl_comm = nullptr;
}

bool ix_is_consumer(const Index *ix) {
  if (ix->n_ranks <= 1 || (!ix->p2p_ready && !ix->nccl_comm)) return true;
  if (!ix->p2p_ready) return true;   // NCCL all-gather: every rank holds the result
  return ix->xroot < 0 || ix->xroot == ix->rank;
}

// Receive buffer of the peer-memory exchange + the send block of the shard-local results.
int32_t ix_xchg_alloc(Index *ix, int n_ranks, int rank, int root) {
  if (n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || root >= n_ranks) {
    set_error("exchange: bad n_ranks / rank / root (1 <= n_ranks <= 8)");
    return TSC_ERR_BAD_ARG;
  }
  if (ix->d_xbuf) {
    set_error("exchange: already set up on this index");
    return TSC_ERR_BAD_ARG;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint64_t slot = (uint64_t)ix->nq_max * ix->k_max * 16ull;
  uint32_t depth = 4;
  while (depth > 2 && xchg_buf_bytes((uint32_t)n_ranks, depth, slot, ix->nq_max) > (256ull << 20))
    depth >>= 1;
  const uint64_t bytes = xchg_buf_bytes((uint32_t)n_ranks, depth, slot, ix->nq_max);
  if (bytes > (1ull << 30)) {
    set_error("exchange: nq_max * k_max too large for the peer-memory exchange (%llu MB)",
              (unsigned long long)(bytes >> 20));
    return TSC_ERR_UNSUPPORTED;
  }
  TSC_CUDA(cudaMalloc((void **)&ix->d_xbuf, bytes));
  TSC_CUDA(cudaMemset(ix->d_xbuf, 0, bytes));          // flags = acks = 0, epochs start at 1
  TSC_CUDA(cudaHostAlloc((void **)&ix->h_xstatus, 4, cudaHostAllocMapped));
  *ix->h_xstatus = 0;
  TSC_CUDA(cudaHostGetDevicePointer((void **)&ix->d_xstatus, ix->h_xstatus, 0));
  if (!ix->d_gather_send) {
    TSC_CUDA(cudaMalloc((void **)&ix->d_gather_send, (size_t)ix->nq_max * ix->k_max * 16));
    ix->device_bytes += (size_t)ix->nq_max * ix->k_max * 16;
  }
  ix->xbuf_bytes = bytes;
  ix->xslot_bytes = slot;
  ix->xdepth = depth;
  ix->device_bytes += bytes;
  ix->n_ranks = n_ranks;
  ix->rank = rank;
  ix->xroot = root;
  ix->xepoch = 0;
#ifdef TSC_DIAG
  if (const char *ev = getenv("TSC_P2P_TIMEOUT_MS"))
    if (atoll(ev) > 0) ix->x_timeout_cycles = atoll(ev) * 2000000ll;
#endif
  return TSC_OK;
}

// ---- one search on stream `st` ---------------------------------------------------------
// d_q: [nq, qld] fp32 on the device, padded. sharded: the result in d_ids / d_dist /
// d_counts is the merge over all shards (valid on the consumer ranks).
static int32_t run_search(Index *ix, const float *d_q, uint32_t nq, uint32_t k, double threshold,
                          int64_t *d_ids, double *d_dist, uint32_t *d_counts, cudaStream_t st,
                          bool sharded, bool device_api) {
  // the search scratch is one per index: order this search after the previous one
  {
    int32_t orc = order_after_last_search(ix, st);
    if (orc != TSC_OK) return orc;
  }
  SearchCtx c;
  c.d_q = d_q;
  c.nq = nq;
  c.k = k;
  c.kprime = kprime_for(k);
  c.threshold = threshold;
  c.st = st;
  c.sharded = sharded;
  c.pipelined = device_api && ix->pipeline;   // host-buffer searches always repair in-stream
  const size_t nk = (size_t)nq * k;
  if (sharded) {
    c.loc_ids = (int64_t *)ix->d_gather_send;
    c.loc_dist = (double *)(ix->d_gather_send + nk * 8);
    c.loc_counts = ix->d_loc_counts;
    c.x_ids = d_ids;
    c.x_dist = d_dist;
    c.x_counts = d_counts;
    ix->xepoch++;
    if (ix->h_xstatus && *reinterpret_cast<volatile uint32_t *>(ix->h_xstatus) != 0) {
      set_error("exchange: an earlier peer-memory exchange timed out (a rank fell behind or died)");
      return TSC_ERR_NCCL;
    }
  } else {
    c.loc_ids = d_ids;
    c.loc_dist = d_dist;
    c.loc_counts = d_counts;
  }
  const bool p2p = sharded && ix->p2p_ready;
  bool need_exchange = sharded;   // still owed after the kernels below
  bool ends_on_timer = false;     // the last thing enqueued is a timer event (fused scan path)
  ix->search_beg = nullptr;
  int32_t rc = TSC_OK;
  ix->search_timed = true;
  if (ix->rows == 0) {  // meta.totalVectors == 0 -> const [] (ngh_graph_engine.dart:78)
    TSC_CUDA(cudaMemsetAsync(c.loc_ids, 0xFF, nk * 8, st));
    TSC_CUDA(cudaMemsetAsync(c.loc_dist, 0xFF, nk * 8, st));
    TSC_CUDA(cudaMemsetAsync(c.loc_counts, 0, (size_t)nq * 4, st));
    TSC_CUDA(cudaMemsetAsync(ix->d_flags, 0, (size_t)nq * 4, st));
  } else {
    rc = refresh_live(ix, st);
    if (rc != TSC_OK) return rc;
    const bool use_gemm = nq >= ix->gemm_min_nq && gemm_supported(ix, c.kprime);
    // The tensor path's keys carry the rounding of the query to the storage type (bf16:
    // ~1e-3 |q||b|, tf32 likewise): with K' = 20 about 1 % of the queries of a Gaussian
    // corpus fail the certificate and each group of them costs a range pass. K' = 32 puts
    // the pivot ~7 sigma of the rank gap away: a range pass per ~10 batches instead.
    if (use_gemm && c.kprime < 32) c.kprime = 32;
    uint32_t lists = 0;
    const bool fused_path = !use_gemm && nq <= 8;
    // (the NCCL exchange enqueues work after the kernels: such searches keep their events)
    ix->search_timed = !(c.pipelined && fused_path && (!sharded || p2p)) ||
                       (ix->timer_tick++ % ix->timer_every) == 0;
    if (!use_gemm && nq <= 8) {
      // the whole search in one kernel (+ one range launch that normally exits at once)
      const int qb = nq == 1 ? 1 : (nq <= 4 ? 4 : 8);
      rc = launch_scan(ix, c, 0, 0, nq, qb, true, p2p, false, &lists);
      if (rc != TSC_OK) return rc;
      // Pipelined searches carry no range launch: a launch between two first passes keeps
      // the second from overlapping the first one's tail (a programmatic dependency reaches
      // one launch back; measured: 564 -> 557 us per query with it, 532 without, 1.25M rows).
      // A query whose certificate fails is flagged instead (tsc_search_flags = 1, counted as
      // uncertified in tsc_stats) and is re-issued by the caller.
      if (!c.pipelined) {
        rc = launch_scan(ix, c, 1, 0, 0, qb, true, p2p, true, nullptr);
        if (rc != TSC_OK) return rc;
      }
      if (p2p) need_exchange = false;
      ends_on_timer = ix->search_timed && !ix->hot_slot_open;
    } else {
      uint32_t list_kp = c.kprime;
      if (use_gemm) {
        if (list_kp > (uint32_t)kGemmMaxKp) list_kp = (uint32_t)kGemmMaxKp;
        c.gemm_list_kp = list_kp;
        rc = launch_gemm(ix, d_q, nq, list_kp, ix->d_cand, &lists, nullptr, st);
        if (rc != TSC_OK) return rc;
      } else {
        for (uint32_t q0 = 0; q0 < nq;) {
          const uint32_t left = nq - q0;
          const int qb = left >= 5 ? 8 : (left >= 2 ? 4 : 1);
          const uint32_t n = left < (uint32_t)qb ? left : (uint32_t)qb;
          rc = launch_scan(ix, c, 0, q0, n, qb, false, false, false, &lists);
          if (rc != TSC_OK) return rc;
          q0 += n;
        }
      }
      rc = launch_tail(ix, c, lists * list_kp, use_gemm);
      if (rc != TSC_OK) return rc;
      uint32_t n_retry = (nq + kRangeSlots - 1) / kRangeSlots;
      if (n_retry > kRetryLaunches) n_retry = kRetryLaunches;
      for (uint32_t g = 0; g < n_retry; g++) {
        rc = launch_scan(ix, c, 1, g * kRangeSlots, 0, (int)kRangeSlots, true, false,
                         g + 1 == n_retry, nullptr);
        if (rc != TSC_OK) return rc;
      }
    }
    ix->last_path = use_gemm ? 2 : 1;
    if (ix->search_timed) {   // an untimed (pipelined) search leaves the last timed one's figures
      uint32_t passes = (nq + 7) / 8;
      if (nq <= 4 || use_gemm) passes = 1;
      ix->last_gbs = (double)passes * (double)ix->rows * ix->desc.dims * ix->elem_bytes;  // bytes
      ix->last_flops = use_gemm ? 2.0 * nq * (double)ix->rows * ix->desc.dims : 0.0;
      ix->last_ms = -1.0;  // resolved lazily from the events
    }
  }
  if (need_exchange) {
    if (p2p) {
      rc = launch_exchange(ix, c);
      if (rc != TSC_OK) return rc;
    } else {
      TSC_NCCL(g_nccl.AllGather(ix->d_gather_send, ix->d_gather_recv, nk * 16, /*ncclUint8*/ 1,
                                ix->nccl_comm, st));
      rc = launch_merge(ix, (const int64_t *)ix->d_gather_recv,
                        (const double *)(ix->d_gather_recv + nk * 8), nk * 2,
                        (uint32_t)ix->n_ranks, nq, k, d_ids, d_dist, d_counts, st);
      if (rc != TSC_OK) return rc;
    }
  }
  ix->searches++;
  // Every event between two kernels costs stream time (a timestamp is a serialising
  // operation): the fused path runs with two per search, the timer pair around the scan
  // launch and its range launch; its end event also marks the end of the search.
  if (!ix->search_timed && !need_exchange) {
    ix->scratch_mark = nullptr;   // pipelined: no event between this search and the next
  } else if (ends_on_timer && !need_exchange) {
    ix->scratch_mark = ix->last_hot_end;
  } else {
    TSC_CUDA(cudaEventRecord(ix->scratch_ev, st));
    ix->scratch_mark = ix->scratch_ev;
  }
  if (ix->search_timed) {
    ix->timed_beg = ix->search_beg;
    ix->timed_end = ix->scratch_mark;
  }
  ix->scratch_stream = st;
  ix->scratch_used = true;
  return TSC_OK;
}

static int32_t check_search_args(Index *ix, const void *q, uint32_t nq, uint32_t k,
                                 const void *ids, const void *dist, const void *counts) {
  if (ix->host_only) {
    set_error("search: host-only self-test handle");
    return TSC_ERR_UNSUPPORTED;
  }
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

static bool ix_sharded(const Index *ix) { return ix->p2p_ready || ix->nccl_comm != nullptr; }

// ---- host-buffer search in two halves (caller holds ix->mu or the group's lock) --------
int32_t ix_search_begin(Index *ix, const float *queries, uint32_t nq, uint32_t k,
                        double threshold) {
  TSC_CUDA(cudaSetDevice(ix->device));
  const uint32_t dims = ix->desc.dims, qld = ix->qld;
  for (uint32_t q = 0; q < nq; q++) {
    memcpy(ix->h_queries + (size_t)q * qld, queries + (size_t)q * dims, (size_t)dims * 4);
    for (uint32_t c = dims; c < qld; c++) ix->h_queries[(size_t)q * qld + c] = 0.0f;
  }
  cudaStream_t st = ix->stream;
  TSC_CUDA(cudaMemcpyAsync(ix->d_queries, ix->h_queries, (size_t)nq * qld * 4,
                           cudaMemcpyHostToDevice, st));
  int32_t rc = run_search(ix, ix->d_queries, nq, k, threshold, ix->d_out_ids, ix->d_out_dist,
                          ix->d_out_counts, st, ix_sharded(ix), false);
  if (rc != TSC_OK) return rc;
  ix->host_consumer = ix_is_consumer(ix);
  if (ix->out_block_bytes <= 16384) {
    // small result block: everything (ids, dist, counts, flags) in one copy
    TSC_CUDA(cudaMemcpyAsync(ix->h_out_block, ix->d_out_block, ix->out_block_bytes,
                             cudaMemcpyDeviceToHost, st));
  } else {
    if (ix->host_consumer) {
      TSC_CUDA(cudaMemcpyAsync(ix->h_out_ids, ix->d_out_ids, (size_t)nq * k * 8,
                               cudaMemcpyDeviceToHost, st));
      TSC_CUDA(cudaMemcpyAsync(ix->h_out_dist, ix->d_out_dist, (size_t)nq * k * 8,
                               cudaMemcpyDeviceToHost, st));
      TSC_CUDA(cudaMemcpyAsync(ix->h_out_counts, ix->d_out_counts, (size_t)nq * 4,
                               cudaMemcpyDeviceToHost, st));
    }
    TSC_CUDA(cudaMemcpyAsync(ix->h_flags, ix->d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  }
  TSC_CUDA(cudaEventRecord(ix->host_done, st));
  return TSC_OK;
}

// Range passes the in-stream launches did not cover (a batch with more than
// kRetryLaunches * kRangeSlots uncertified queries): driven from the host until none is left.
static int32_t host_range_passes(Index *ix, uint32_t nq, uint32_t k, double threshold) {
  cudaStream_t st = ix->stream;
  for (;;) {
    std::vector<uint32_t> pend;
    for (uint32_t q = 0; q < nq; q++)
      if (ix->h_flags[q] == kFlagRetry) pend.push_back(q);
    if (pend.empty()) return TSC_OK;
    const uint32_t n = (uint32_t)pend.size();
    TSC_CUDA(cudaMemcpyAsync(ix->d_retry_list, pend.data(), (size_t)n * 4, cudaMemcpyHostToDevice,
                             st));
    TSC_CUDA(cudaMemcpyAsync(ix->d_retry_n, &n, 4, cudaMemcpyHostToDevice, st));
    TSC_CUDA(cudaStreamSynchronize(st));   // pend / n live on this stack frame
    SearchCtx c;
    c.d_q = ix->d_queries;
    c.nq = nq;
    c.k = k;
    c.kprime = kprime_for(k);
    c.threshold = threshold;
    c.st = st;
    c.loc_ids = ix->d_out_ids;
    c.loc_dist = ix->d_out_dist;
    c.loc_counts = ix->d_out_counts;
    const uint32_t groups = (n + kRangeSlots - 1) / kRangeSlots;
    for (uint32_t g = 0; g < groups; g++) {
      int32_t rc = launch_scan(ix, c, 1, g * kRangeSlots, 0, (int)kRangeSlots, true, false,
                               g + 1 == groups, nullptr);
      if (rc != TSC_OK) return rc;
    }
    TSC_CUDA(cudaMemcpyAsync(ix->h_out_ids, ix->d_out_ids, (size_t)nq * k * 8,
                             cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaMemcpyAsync(ix->h_out_dist, ix->d_out_dist, (size_t)nq * k * 8,
                             cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaMemcpyAsync(ix->h_out_counts, ix->d_out_counts, (size_t)nq * 4,
                             cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaMemcpyAsync(ix->h_flags, ix->d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaStreamSynchronize(st));
    for (uint32_t q : pend)
      if (ix->h_flags[q] == kFlagRetry) {   // cannot happen: a range pass always decides
        set_error("search: range pass left query %u undecided", q);
        return TSC_ERR_CUDA;
      }
  }
}

int32_t ix_search_end(Index *ix, uint32_t nq, uint32_t k, int64_t *out_ids, double *out_dist,
                      uint32_t *out_counts) {
  TSC_CUDA(cudaSetDevice(ix->device));
  TSC_CUDA(cudaEventSynchronize(ix->host_done));
  if (ix->h_xstatus && *reinterpret_cast<volatile uint32_t *>(ix->h_xstatus) != 0) {
    set_error("exchange: a peer did not deliver its results in time (a rank fell behind or died)");
    return TSC_ERR_NCCL;
  }
  if (!ix_sharded(ix)) {
    int32_t rc = host_range_passes(ix, nq, k, ix->last_threshold);
    if (rc != TSC_OK) return rc;
  }
  if (out_ids && ix->host_consumer) {
    memcpy(out_ids, ix->h_out_ids, (size_t)nq * k * 8);
    memcpy(out_dist, ix->h_out_dist, (size_t)nq * k * 8);
    memcpy(out_counts, ix->h_out_counts, (size_t)nq * 4);
  }
  return TSC_OK;
}

// ---- tickets -------------------------------------------------------------------------
// At most one host-buffer search is in flight per handle. Whoever needs the handle's pinned
// buffers next retires the ticket in flight first: its results are delivered to ITS buffers
// (caller-owned until the ticket is waited on), so nobody is rejected and nothing is lost.
static void retire_locked(TicketRef t) {   // by value: resets the holder it usually comes from; t->ix->mu or t->grp->mu is held
  if (t->retired) return;
  if (t->ix) {
    t->rc = ix_search_end(t->ix.get(), t->nq, t->k, t->out_ids, t->out_dist, t->out_counts);
    t->ix->inflight.reset();
  } else {
    t->rc = grp_search_end(*t->grp, t->out_ids, t->out_dist, t->out_counts);
    t->grp->inflight.reset();
  }
  t->retired = true;
}

static int32_t submit_ticket(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                             double threshold, int64_t *out_ids, double *out_dist,
                             uint32_t *out_counts, TicketRef *out) {
  TicketRef t = std::make_shared<Ticket>();
  t->handle = handle;
  t->out_ids = out_ids;
  t->out_dist = out_dist;
  t->out_counts = out_counts;
  t->nq = nq;
  t->k = k;
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (!queries || !out_ids || !out_dist || !out_counts) {
      set_error("search: NULL buffer");
      return TSC_ERR_BAD_ARG;
    }
    if (g->inflight) retire_locked(g->inflight);
    int32_t rc = grp_search_begin(*g, queries, nq, k, threshold);
    if (rc != TSC_OK) return rc;
    t->grp = g;
    g->inflight = t;
  } else {
    IndexRef ref = lookup_index(handle);
    if (!ref) return TSC_ERR_BAD_HANDLE;
    Index *ix = ref.get();
    std::lock_guard<std::mutex> lk(ix->mu);
    int32_t rc = check_search_args(ix, queries, nq, k, out_ids, out_dist, out_counts);
    if (rc != TSC_OK) return rc;
    if (ix->inflight) retire_locked(ix->inflight);
    ix->last_threshold = threshold;
    rc = ix_search_begin(ix, queries, nq, k, threshold);
    if (rc != TSC_OK) return rc;
    t->ix = ref;
    ix->inflight = t;
  }
  *out = t;
  return TSC_OK;
}

// done: 1 when the results are in the caller's buffers. block: wait for them.
static int32_t finish_ticket(uint64_t ticket, bool block, int32_t *out_done) {
  TicketRef t;
  {
    std::lock_guard<std::mutex> g(g_tmu);
    auto it = g_tickets.find(ticket);
    if (it == g_tickets.end()) {
      set_error("unknown ticket %llu", (unsigned long long)ticket);
      return TSC_ERR_BAD_HANDLE;
    }
    t = it->second;
  }
  {
    std::mutex &mu = t->ix ? t->ix->mu : t->grp->mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!t->retired) {
      if (!block) {
        Index *root = t->ix ? t->ix.get() : t->grp->shards[0].get();
        cudaSetDevice(root->device);
        cudaError_t e = cudaEventQuery(root->host_done);
        if (e == cudaErrorNotReady) {
          if (out_done) *out_done = 0;
          return TSC_OK;
        }
        cudaGetLastError();
      }
      retire_locked(t);
    }
  }
  {
    // exactly one caller removes the ticket; a second poller of the same id sees "unknown"
    std::lock_guard<std::mutex> g(g_tmu);
    g_tickets.erase(ticket);
  }
  if (out_done) *out_done = 1;
  return t->rc;
}

}  // namespace tsc

using namespace tsc;

extern "C" {

// ---- search -------------------------------------------------------------------
int32_t tsc_search_device(uint64_t handle, const float *d_queries, uint32_t nq, uint32_t k,
                          double threshold, int64_t *d_out_ids, double *d_out_dist,
                          uint32_t *d_out_counts, void *cuda_stream) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  int32_t rc = check_search_args(ix, d_queries, nq, k, d_out_ids, d_out_dist, d_out_counts);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  const float *q = d_queries;
  if (ix->qld != ix->desc.dims) {
    int32_t orc = order_after_last_search(ix, st);
    if (orc != TSC_OK) return orc;
    int32_t prc = launch_pad_queries(ix, d_queries, nq, st);
    if (prc != TSC_OK) return prc;
    q = ix->d_queries;
  }
  return run_search(ix, q, nq, k, threshold, d_out_ids, d_out_dist, d_out_counts, st, false, true);
  TSC_API_CATCH
}

int32_t tsc_search_sharded(uint64_t handle, const float *d_queries, uint32_t nq, uint32_t k,
                           double threshold, int64_t *d_out_ids, double *d_out_dist,
                           uint32_t *d_out_counts, void *cuda_stream) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix_sharded(ix)) {
    set_error("search_sharded: neither tsc_comm_init nor tsc_comm_p2p_import has been called");
    return TSC_ERR_NCCL;
  }
  int32_t rc = check_search_args(ix, d_queries, nq, k, d_out_ids, d_out_dist, d_out_counts);
  if (rc != TSC_OK) return rc;
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  const float *q = d_queries;
  if (ix->qld != ix->desc.dims) {
    int32_t orc = order_after_last_search(ix, st);
    if (orc != TSC_OK) return orc;
    int32_t prc = launch_pad_queries(ix, d_queries, nq, st);
    if (prc != TSC_OK) return prc;
    q = ix->d_queries;
  }
  return run_search(ix, q, nq, k, threshold, d_out_ids, d_out_dist, d_out_counts, st, true, true);
  TSC_API_CATCH
}

int32_t tsc_search_submit(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                          double threshold, int64_t *out_ids, double *out_dist,
                          uint32_t *out_counts, uint64_t *out_ticket) {
  TSC_API_TRY
  if (!out_ticket) {
    set_error("search_submit: NULL out_ticket");
    return TSC_ERR_BAD_ARG;
  }
  TicketRef t;
  int32_t rc = submit_ticket(handle, queries, nq, k, threshold, out_ids, out_dist, out_counts, &t);
  if (rc != TSC_OK) return rc;
  std::lock_guard<std::mutex> g(g_tmu);
  const uint64_t id = g_next_ticket++;
  g_tickets[id] = t;
  *out_ticket = id;
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_search_poll(uint64_t ticket, int32_t *out_done) {
  TSC_API_TRY
  if (!out_done) {
    set_error("search_poll: NULL out_done");
    return TSC_ERR_BAD_ARG;
  }
  return finish_ticket(ticket, false, out_done);
  TSC_API_CATCH
}

int32_t tsc_search_wait(uint64_t ticket) {
  TSC_API_TRY
  return finish_ticket(ticket, true, nullptr);
  TSC_API_CATCH
}

// Blocking search: the handle's lock is held from submission until the results are in the
// caller's buffers, so concurrent callers on one handle take turns.
int32_t tsc_search(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                   double threshold, int64_t *out_ids, double *out_dist, uint32_t *out_counts) {
  TSC_API_TRY
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (!queries || !out_ids || !out_dist || !out_counts) {
      set_error("search: NULL buffer");
      return TSC_ERR_BAD_ARG;
    }
    if (g->inflight) retire_locked(g->inflight);
    int32_t rc = grp_search_begin(*g, queries, nq, k, threshold);
    if (rc != TSC_OK) return rc;
    return grp_search_end(*g, out_ids, out_dist, out_counts);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  int32_t rc = check_search_args(ix, queries, nq, k, out_ids, out_dist, out_counts);
  if (rc != TSC_OK) return rc;
  if (ix->inflight) retire_locked(ix->inflight);
  ix->last_threshold = threshold;
  rc = ix_search_begin(ix, queries, nq, k, threshold);
  if (rc != TSC_OK) return rc;
  return ix_search_end(ix, nq, k, out_ids, out_dist, out_counts);
  TSC_API_CATCH
}

int32_t tsc_search_flags(uint64_t handle, uint32_t nq, uint32_t *out_flags) {
  TSC_API_TRY
  if (!out_flags || nq == 0) {
    set_error("search_flags: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> lk(g->mu);
    return grp_search_flags(*g, nq, out_flags);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->host_only || nq > ix->nq_max) {
    set_error("search_flags: nq=%u outside [1, nq_max]", nq);
    return TSC_ERR_BAD_ARG;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  if (ix->inflight) retire_locked(ix->inflight);
  {
    int32_t src = sync_last_search(ix);
    if (src != TSC_OK) return src;
  }
  TSC_CUDA(cudaMemcpy(out_flags, ix->d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost));
  return TSC_OK;
  TSC_API_CATCH
}

}  // extern "C"

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

extern "C" {

int32_t tsc_vector_search(uint64_t handle, const double *values, uint64_t len, uint32_t k,
                          double threshold, int64_t *out_ids, double *out_dist,
                          double *out_score, uint32_t *out_count) {
  return tsc_vector_search_batch(handle, values, len, 1, k, threshold, out_ids, out_dist, out_score,
                                 out_count);
}

// Batch form (additive: the reference's API is single-query): nq query vectors of `len`
// values each, prepared like single queries, searched in one call (the tcgen05 GEMM path for
// nq >= 9), scored. out_* are [nq][k], out_counts [nq].
int32_t tsc_vector_search_batch(uint64_t handle, const double *values, uint64_t len, uint32_t nq,
                                uint32_t k, double threshold, int64_t *out_ids, double *out_dist,
                                double *out_score, uint32_t *out_counts) {
  TSC_API_TRY
  uint32_t dims;
  int metric;
  if (GroupRef g = lookup_group(handle)) {
    dims = g->desc.dims;
    metric = g->desc.metric;
  } else {
    IndexRef ref = lookup_index(handle);
    if (!ref) return TSC_ERR_BAD_HANDLE;
    dims = ref->desc.dims;
    metric = ref->desc.metric;
  }
  if ((!values && len) || !out_ids || !out_dist || !out_score || !out_counts || nq == 0 ||
      nq > 65535) {
    set_error("vector_search: NULL buffer or nq outside [1, 65535]");
    return TSC_ERR_BAD_ARG;
  }
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
  TSC_API_CATCH
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

// ---- sharding, one process per GPU -----------------------------------------------------
int32_t tsc_merge_shards(uint64_t handle, const int64_t *d_part_ids, const double *d_part_dist,
                         uint32_t n_parts, uint32_t nq, uint32_t k, int64_t *d_out_ids,
                         double *d_out_dist, uint32_t *d_out_counts, void *cuda_stream) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  if (!d_part_ids || !d_part_dist || !d_out_ids || !d_out_dist || !d_out_counts || !n_parts ||
      !nq || !k || ix->host_only) {
    set_error("merge_shards: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->stream;
  return launch_merge(ix, d_part_ids, d_part_dist, (uint64_t)nq * k, n_parts, nq, k, d_out_ids,
                      d_out_dist, d_out_counts, st);
  TSC_API_CATCH
}

int32_t tsc_comm_unique_id(uint8_t *out_id128) {
  TSC_API_TRY
  if (!out_id128) {
    set_error("comm_unique_id: NULL");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = nccl_load();
  if (rc != TSC_OK) return rc;
  TSC_NCCL(g_nccl.GetUniqueId(out_id128));
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_comm_init(uint64_t handle, const uint8_t *id128, int32_t n_ranks, int32_t rank) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  if (!id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks || ix->host_only) {
    set_error("comm_init: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = nccl_load();
  if (rc != TSC_OK) return rc;
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->nccl_comm || ix->p2p_ready) {
    set_error("comm_init: this index already has a communicator");
    return TSC_ERR_BAD_ARG;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  Id128 id;
  memcpy(id.b, id128, 128);
  TSC_NCCL(g_nccl.CommInitRank(&ix->nccl_comm, n_ranks, id, rank));
  ix->n_ranks = n_ranks;
  ix->rank = rank;
  const size_t part = (size_t)ix->nq_max * ix->k_max * 16;
  if (!ix->d_gather_send) {
    TSC_CUDA(cudaMalloc((void **)&ix->d_gather_send, part));
    ix->device_bytes += part;
  }
  if (!ix->d_gather_recv) {
    TSC_CUDA(cudaMalloc((void **)&ix->d_gather_recv, part * n_ranks));
    ix->device_bytes += part * n_ranks;
  }
  return TSC_OK;
  TSC_API_CATCH
}

// export: allocate this rank's receive buffer and return its CUDA IPC handle (64 bytes);
// the caller all-gathers the handles of all ranks (any host transport) and passes them to
// import, which maps every peer's buffer. One process per GPU (IPC handles cannot be opened
// by the process that made them; inside one process use a group handle instead).
int32_t tsc_comm_p2p_export(uint64_t handle, int32_t n_ranks, int32_t rank, uint8_t *out_ipc64) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  if (!out_ipc64 || ix->host_only) {
    set_error("comm_p2p_export: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lk(ix->mu);
  int32_t rc = ix_xchg_alloc(ix, n_ranks, rank, -1);
  if (rc != TSC_OK) return rc;
  cudaIpcMemHandle_t hnd;
  TSC_CUDA(cudaIpcGetMemHandle(&hnd, ix->d_xbuf));
  memcpy(out_ipc64, &hnd, 64);
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_comm_p2p_import(uint64_t handle, const uint8_t *all_ipc, int32_t root) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!all_ipc || !ix->d_xbuf || ix->p2p_ready || root >= ix->n_ranks) {
    set_error("comm_p2p_import: NULL handles, bad root, or tsc_comm_p2p_export not called");
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
    ix->x_ipc = true;
  }
  ix->xroot = root < 0 ? -1 : root;
  ix->p2p_ready = true;
  return TSC_OK;
  TSC_API_CATCH
}

// Test hook: raw fp32 ranking keys of the tensor-core path for every (query, row),
// so tests can pin the UMMA / TMA / TMEM layouts against a plain matmul.
int32_t tsc_debug_gemm_keys(uint64_t handle, const float *queries, uint32_t nq, float *out_keys) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  Index *ix = ref.get();
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!queries || !out_keys || nq == 0 || ix->host_only || nq > ix->nq_max || ix->rows == 0) {
    set_error("debug_gemm_keys: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  const uint32_t kprime = kprime_for(1);
  if (!gemm_supported(ix, kprime)) {
    set_error("debug_gemm_keys: index has no tensor-core path");
    return TSC_ERR_UNSUPPORTED;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  if (ix->inflight) retire_locked(ix->inflight);
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
  TSC_API_CATCH
}

}  // extern "C"
