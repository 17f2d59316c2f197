// tsc_group.cu — one logical embedding column row-range sharded over the GPUs of ONE
// process (tsc_index_create with n_devices = 2..8; SURVEY.md §8b: "one process owns all
// GPUs"). This is the form the single-process Dart host reaches multi-GPU through: every
// host-buffer entry point of the C ABI works on the group handle, and one tsc_search call
// scans all shards and merges their exact top-k on shard 0 — each shard's scan kernel pushes
// its result into shard 0's memory over NVLink (tsc_exchange.cuh), shard 0's kernel merges.
// No row data ever crosses NVLink.
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <thread>

#include "tsc_index.h"

namespace tsc {

static uint32_t src_bpe(const tsc_index_desc &d) {
  return d.src_precision == TSC_SRC_F64 ? 8 : (d.src_precision == TSC_SRC_I8 ? 1 : 4);
}
static uint64_t shard_base(const Group &g, size_t s) { return g.desc.first_node_id + s * g.per_shard; }

// for every shard that intersects node ids [first, first + n): fn(shard, lo, hi)
template <typename Fn>
static int32_t for_shards(Group &g, uint64_t first, uint64_t n, Fn fn) {
  if (first < g.desc.first_node_id || first + n > g.desc.first_node_id + g.desc.capacity_rows) {
    set_error("node ids [%llu, %llu) outside the column [%llu, %llu)", (unsigned long long)first,
              (unsigned long long)(first + n), (unsigned long long)g.desc.first_node_id,
              (unsigned long long)(g.desc.first_node_id + g.desc.capacity_rows));
    return first + n > g.desc.first_node_id + g.desc.capacity_rows && first >= g.desc.first_node_id
               ? TSC_ERR_OOM
               : TSC_ERR_BAD_ARG;
  }
  for (size_t s = 0; s < g.shards.size(); s++) {
    const uint64_t b = shard_base(g, s);
    const uint64_t lo = first > b ? first : b;
    const uint64_t hi = first + n < b + g.per_shard ? first + n : b + g.per_shard;
    if (lo >= hi) continue;
    int32_t rc = fn(g.shards[s].get(), lo, hi);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

// ---- shard workers ---------------------------------------------------------------------
struct GroupWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  Index *ix = nullptr;
  enum Job { kNone, kBegin, kEnd, kQuit } job = kNone;
  bool done = true;
  // arguments / results of the job
  const float *queries = nullptr;
  uint32_t nq = 0, k = 0;
  double threshold = 0;
  int32_t rc = TSC_OK;
  std::string err;

  void run() {
    cudaSetDevice(ix->device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return job != kNone; });
      if (job == kQuit) return;
      const Job j = job;
      lk.unlock();
      int32_t r = TSC_OK;
      if (j == kBegin) {
        ix->last_threshold = threshold;
        r = ix_search_begin(ix, queries, nq, k, threshold);
      } else {
        r = ix_search_end(ix, nq, k, nullptr, nullptr, nullptr);
      }
      lk.lock();
      rc = r;
      if (r != TSC_OK) err = last_error_text();
      job = kNone;
      done = true;
      cv.notify_all();
    }
  }
  void post(Job j) {
    std::lock_guard<std::mutex> lk(mu);
    job = j;
    done = false;
    cv.notify_all();
  }
  int32_t wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    if (rc != TSC_OK) set_error("%s", err.c_str());
    return rc;
  }
};

static void start_workers(Group &g) {
  for (size_t s = 1; s < g.shards.size(); s++) {
    std::unique_ptr<GroupWorker> w(new GroupWorker);
    w->ix = g.shards[s].get();
    GroupWorker *raw = w.get();
    w->th = std::thread([raw] { raw->run(); });
    g.workers.push_back(std::move(w));
  }
}

Group::~Group() {
  for (auto &w : workers) {
    w->post(GroupWorker::kQuit);
    if (w->th.joinable()) w->th.join();
  }
}

int32_t grp_create(const tsc_index_desc *d, uint64_t *out_handle) {
  const uint32_t n = d->n_devices;
  if (d->capacity_rows < (uint64_t)n * 64) {
    set_error("index_create: capacity_rows=%llu too small to shard over %u devices",
              (unsigned long long)d->capacity_rows, n);
    return TSC_ERR_BAD_ARG;
  }
  for (uint32_t a = 0; a < n; a++)
    for (uint32_t b = a + 1; b < n; b++)
      if (d->device_ids[a] == d->device_ids[b]) {
        set_error("index_create: device %d listed twice", d->device_ids[a]);
        return TSC_ERR_BAD_ARG;
      }
  GroupRef g = std::make_shared<Group>();
  g->desc = *d;
  // shard boundaries on 64-row multiples: filter bitmaps split on whole 64-bit words
  g->per_shard = ((d->capacity_rows + n - 1) / n + 63) / 64 * 64;
  for (uint32_t s = 0; s < n; s++) {
    tsc_index_desc one = *d;
    one.n_devices = 1;
    one.device_id = d->device_ids[s];
    one.capacity_rows = g->per_shard;
    one.first_node_id = d->first_node_id + (uint64_t)s * g->per_shard;
    IndexRef ix;
    int32_t rc = ix_create(&one, &ix);
    if (rc != TSC_OK) return rc;
    ix->in_group = true;
    g->shards.push_back(ix);
  }
  // peer access in both directions, then every shard learns every receive buffer
  for (uint32_t a = 0; a < n; a++) {
    TSC_CUDA(cudaSetDevice(d->device_ids[a]));
    for (uint32_t b = 0; b < n; b++) {
      if (a == b) continue;
      int can = 0;
      TSC_CUDA(cudaDeviceCanAccessPeer(&can, d->device_ids[a], d->device_ids[b]));
      if (!can) {
        set_error("index_create: device %d cannot access device %d (no NVLink / P2P)",
                  d->device_ids[a], d->device_ids[b]);
        return TSC_ERR_UNSUPPORTED;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(d->device_ids[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("index_create: cudaDeviceEnablePeerAccess(%d -> %d): %s", d->device_ids[a],
                  d->device_ids[b], cudaGetErrorString(e));
        cudaGetLastError();
        return TSC_ERR_CUDA;
      }
      cudaGetLastError();
    }
  }
  for (uint32_t s = 0; s < n; s++) {
    int32_t rc = ix_xchg_alloc(g->shards[s].get(), (int)n, (int)s, /*root=*/0);
    if (rc != TSC_OK) return rc;
  }
  for (uint32_t s = 0; s < n; s++) {
    Index *ix = g->shards[s].get();
    for (uint32_t r = 0; r < n; r++) ix->x_peer[r] = g->shards[r]->d_xbuf;
    ix->p2p_ready = true;
  }
  start_workers(*g);
  *out_handle = register_group(g);
  return TSC_OK;
}

int32_t grp_clear(Group &g) {
  for (auto &s : g.shards) {
    int32_t rc = ix_clear(s.get());
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_append_rows(Group &g, uint64_t first_node_id, const void *rows, uint64_t n_rows) {
  if (n_rows == 0) return TSC_OK;
  if (!rows) {
    set_error("append_rows: NULL rows");
    return TSC_ERR_BAD_ARG;
  }
  const size_t src_row = (size_t)g.desc.dims * src_bpe(g.desc);
  return for_shards(g, first_node_id, n_rows, [&](Index *ix, uint64_t lo, uint64_t hi) {
    return ix_append_rows(ix, lo, (const uint8_t *)rows + (lo - first_node_id) * src_row, hi - lo);
  });
}

int32_t grp_append_synthetic(Group &g, uint64_t seed, uint64_t first_node_id, uint64_t n_rows) {
  if (n_rows == 0) return TSC_OK;
  return for_shards(g, first_node_id, n_rows, [&](Index *ix, uint64_t lo, uint64_t hi) {
    return ix_append_synthetic(ix, seed, lo, hi - lo);
  });
}

int32_t grp_append_pages(Group &g, uint64_t first_logical_page, const uint8_t *pages,
                         uint64_t n_pages, uint32_t page_size, uint64_t live_rows) {
  if (n_pages == 0) return TSC_OK;
  if (!pages || page_size < 128) {
    set_error("append_pages: NULL pages or page_size < 128");
    return TSC_ERR_BAD_ARG;
  }
  // NghPageSizer.vectorsPerRawPage, core/ngh_page.dart:575-579
  const int64_t usable = (int64_t)page_size - 20 - 8 - 64;
  const uint64_t rpp =
      usable > 0 ? (uint64_t)(usable / ((int64_t)g.desc.dims * src_bpe(g.desc))) : 0;
  if (rpp == 0) {
    set_error("append_pages: dims=%u does not fit a %u-byte page", g.desc.dims, page_size);
    return TSC_ERR_BAD_DIMS;
  }
  // every shard gets the pages that hold at least one of its node ids (a page may straddle
  // a shard boundary: both neighbours decode their own slots of it)
  for (size_t s = 0; s < g.shards.size(); s++) {
    const uint64_t b = shard_base(g, s);
    uint64_t p_lo = b / rpp, p_hi = (b + g.per_shard + rpp - 1) / rpp;
    if (p_lo < first_logical_page) p_lo = first_logical_page;
    if (p_hi > first_logical_page + n_pages) p_hi = first_logical_page + n_pages;
    if (p_lo >= p_hi) continue;
    int32_t rc = ix_append_pages(g.shards[s].get(), p_lo,
                                 pages + (p_lo - first_logical_page) * page_size, p_hi - p_lo,
                                 page_size, live_rows);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_set_deleted(Group &g, const uint64_t *node_ids, uint64_t n, uint8_t deleted) {
  for (auto &s : g.shards) {   // the kernel ignores ids outside the shard
    int32_t rc = ix_set_deleted(s.get(), node_ids, n, deleted);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_apply_graph_pages(Group &g, uint64_t first_logical_page, const uint8_t *pages,
                              uint64_t n_pages, uint32_t page_size) {
  for (auto &s : g.shards) {   // the kernel ignores slots outside the shard
    int32_t rc = ix_apply_graph_pages(s.get(), first_logical_page, pages, n_pages, page_size);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_set_filter(Group &g, const uint64_t *bitmap_words, uint64_t n_words) {
  for (size_t s = 0; s < g.shards.size(); s++) {
    Index *ix = g.shards[s].get();
    if (!bitmap_words) {
      int32_t rc = ix_set_filter(ix, nullptr, 0);
      if (rc != TSC_OK) return rc;
      continue;
    }
    const uint64_t w0 = s * (g.per_shard / 64);
    int32_t rc = ix_set_filter(ix, bitmap_words + (w0 < n_words ? w0 : n_words),
                               w0 < n_words ? n_words - w0 : 0);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_column_create(Group &g, uint32_t column_id, uint8_t col_type) {
  for (auto &s : g.shards) {
    int32_t rc = ix_column_create(s.get(), column_id, col_type);
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

int32_t grp_column_append(Group &g, uint32_t column_id, uint64_t first_node_id, const void *values,
                          const uint8_t *is_null, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!values) {
    set_error("column_append: NULL values");
    return TSC_ERR_BAD_ARG;
  }
  return for_shards(g, first_node_id, n, [&](Index *ix, uint64_t lo, uint64_t hi) {
    const uint64_t o = lo - first_node_id;
    return ix_column_append(ix, column_id, lo, (const uint8_t *)values + o * 8,
                            is_null ? is_null + o : nullptr, hi - lo);
  });
}

// every shard keeps its own dictionary of the strings of ITS rows
int32_t grp_column_append_text(Group &g, uint32_t column_id, uint64_t first_node_id,
                               const uint16_t *units, const uint64_t *offsets,
                               const uint8_t *is_null, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!offsets) {
    set_error("column_append_text: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  return for_shards(g, first_node_id, n, [&](Index *ix, uint64_t lo, uint64_t hi) {
    const uint64_t o = lo - first_node_id;
    return ix_column_append_text(ix, column_id, lo, units, offsets + o,
                                 is_null ? is_null + o : nullptr, hi - lo);
  });
}

int32_t grp_filter_where(Group &g, const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                         uint32_t n_in_args, const WhereTexts &texts, uint64_t *out_matched) {
  uint64_t total = 0;
  for (auto &s : g.shards) {
    uint64_t m = 0;
    int32_t rc = ix_filter_where(s.get(), ops, n_ops, in_args, n_in_args, texts, &m);
    if (rc != TSC_OK) return rc;
    total += m;
  }
  if (out_matched) *out_matched = total;
  return TSC_OK;
}

int32_t grp_set_primary_keys(Group &g, uint64_t first_node_id, const uint8_t *utf8,
                             const uint64_t *offsets, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!offsets) {
    set_error("set_primary_keys: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  return for_shards(g, first_node_id, n, [&](Index *ix, uint64_t lo, uint64_t hi) {
    return ix_set_primary_keys(ix, lo, utf8, offsets + (lo - first_node_id), hi - lo);
  });
}

// every shard looks the keys up in ITS table and installs its part of the filter
int32_t grp_filter_primary_keys(Group &g, const uint8_t *utf8, const uint64_t *offsets, uint64_t n,
                                uint64_t *out_matched) {
  uint64_t total = 0;
  for (auto &s : g.shards) {
    uint64_t m = 0;
    int32_t rc = ix_filter_primary_keys(s.get(), utf8, offsets, n, &m);
    if (rc != TSC_OK) return rc;
    total += m;
  }
  if (out_matched) *out_matched = total;
  return TSC_OK;
}

int32_t grp_get_primary_key(Group &g, uint64_t node_id, uint8_t *out_utf8, uint32_t capacity,
                            uint32_t *out_len) {
  if (!out_len) {
    set_error("get_primary_key: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  *out_len = 0;
  if (node_id < g.desc.first_node_id) return TSC_OK;
  const uint64_t s = (node_id - g.desc.first_node_id) / g.per_shard;
  if (s >= g.shards.size()) return TSC_OK;
  return ix_get_primary_key(g.shards[s].get(), node_id, out_utf8, capacity, out_len);
}

int32_t grp_load_ngh(Group &g, const char *index_dir, uint32_t flags, tsc_ngh_info *out) {
  if (out && out->struct_size != sizeof(tsc_ngh_info)) {
    set_error("index_load_ngh: struct_size mismatch");
    return TSC_ERR_BAD_ARG;
  }
  const auto t0 = std::chrono::steady_clock::now();
  tsc_ngh_info total;
  memset(&total, 0, sizeof total);
  for (auto &s : g.shards) {   // every shard reads only the partition files of its node ids
    tsc_ngh_info m;
    memset(&m, 0, sizeof m);
    m.struct_size = sizeof m;
    int32_t rc = ix_load_ngh(s.get(), index_dir, flags, &m);
    if (rc != TSC_OK) return rc;
    const uint64_t f = total.files_read + m.files_read, p = total.pages_read + m.pages_read,
                   b = total.bytes_read + m.bytes_read;
    total = m;
    total.files_read = f;
    total.pages_read = p;
    total.bytes_read = b;
  }
  total.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (out) *out = total;
  return TSC_OK;
}

int32_t grp_stats_get(Group &g, tsc_stats *out) {
  tsc_stats acc;
  memset(&acc, 0, sizeof acc);
  acc.struct_size = sizeof acc;
  for (size_t s = 0; s < g.shards.size(); s++) {
    tsc_stats st;
    memset(&st, 0, sizeof st);
    st.struct_size = sizeof st;
    int32_t rc = ix_stats_get(g.shards[s].get(), &st);
    if (rc != TSC_OK) return rc;
    acc.dims = st.dims;
    acc.rows += st.rows;
    acc.deleted_rows += st.deleted_rows;
    acc.device_bytes += st.device_bytes;
    acc.row_stride_bytes = st.row_stride_bytes;
    acc.kernel_launches += st.kernel_launches;
    acc.hot_launches += st.hot_launches;
    acc.hot_bytes_total += st.hot_bytes_total;
    acc.hot_flops_total += st.hot_flops_total;
    // shards run side by side: time is the slowest shard's, bytes add up
    if (st.hot_ms_total > acc.hot_ms_total) acc.hot_ms_total = st.hot_ms_total;
    if (st.last_search_ms > acc.last_search_ms) acc.last_search_ms = st.last_search_ms;
    acc.last_scan_gbs += st.last_scan_gbs;
    acc.last_tflops += st.last_tflops;                 // shards run side by side
    acc.last_tensor_util += st.last_tensor_util / (double)g.shards.size();
    acc.certified_queries += st.certified_queries;
    acc.retried_queries += st.retried_queries;
    acc.uncertified_queries += st.uncertified_queries;
    acc.range_rows += st.range_rows;
    if (s == 0) {
      acc.searches = st.searches;
      acc.last_path = st.last_path;
    }
  }
  acc.n_devices = (uint32_t)g.shards.size();
  *out = acc;
  return TSC_OK;
}

int32_t grp_stats_reset(Group &g) {
  for (auto &s : g.shards) {
    int32_t rc = ix_stats_reset(s.get());
    if (rc != TSC_OK) return rc;
  }
  return TSC_OK;
}

// One search over all shards: the workers enqueue shards 1..n-1 (they only push their top-k),
// the calling thread enqueues the root (shard 0: it waits for everybody's pairs and merges).
int32_t grp_search_begin(Group &g, const float *queries, uint32_t nq, uint32_t k,
                         double threshold) {
  Index *root = g.shards[0].get();
  if (nq == 0 || nq > root->nq_max) {
    set_error("search: nq=%u outside [1, nq_max=%u]", nq, root->nq_max);
    return TSC_ERR_BAD_ARG;
  }
  if (k == 0 || k > root->k_max) {
    set_error("search: k=%u outside [1, k_max=%u]", k, root->k_max);
    return TSC_ERR_BAD_ARG;
  }
  for (auto &w : g.workers) {
    w->queries = queries;
    w->nq = nq;
    w->k = k;
    w->threshold = threshold;
    w->post(GroupWorker::kBegin);
  }
  root->last_threshold = threshold;
  int32_t rc = ix_search_begin(root, queries, nq, k, threshold);
  for (auto &w : g.workers) {
    const int32_t r2 = w->wait();   // a failed shard: the exchange of this epoch times out on the others
    if (rc == TSC_OK) rc = r2;
  }
  if (rc != TSC_OK) return rc;
  g.search_pending = true;
  g.pend_nq = nq;
  g.pend_k = k;
  return TSC_OK;
}

int32_t grp_search_end(Group &g, int64_t *out_ids, double *out_dist, uint32_t *out_counts) {
  if (!g.search_pending) {
    set_error("search: no search in flight on this group");
    return TSC_ERR_BAD_ARG;
  }
  g.search_pending = false;
  for (auto &w : g.workers) {
    w->nq = g.pend_nq;
    w->k = g.pend_k;
    w->post(GroupWorker::kEnd);
  }
  int32_t rc = ix_search_end(g.shards[0].get(), g.pend_nq, g.pend_k, out_ids, out_dist, out_counts);
  for (auto &w : g.workers) {
    const int32_t r2 = w->wait();
    if (rc == TSC_OK) rc = r2;
  }
  return rc;
}

// a query is exact when every shard's local top-k was certified
int32_t grp_search_flags(Group &g, uint32_t nq, uint32_t *out_flags) {
  if (nq > g.shards[0]->nq_max) {
    set_error("search_flags: nq=%u outside [1, nq_max]", nq);
    return TSC_ERR_BAD_ARG;
  }
  if (g.inflight) {
    set_error("search_flags: a ticket is still in flight on this group");
    return TSC_ERR_NOT_READY;
  }
  memset(out_flags, 0, (size_t)nq * 4);
  std::vector<uint32_t> f(nq);
  for (auto &s : g.shards) {
    Index *ix = s.get();
    TSC_CUDA(cudaSetDevice(ix->device));
    {
      int32_t src = sync_last_search(ix);
      if (src != TSC_OK) return src;
    }
    TSC_CUDA(cudaMemcpy(f.data(), ix->d_flags, (size_t)nq * 4, cudaMemcpyDeviceToHost));
    for (uint32_t q = 0; q < nq; q++)
      if (f[q] > out_flags[q]) out_flags[q] = f[q];
  }
  return TSC_OK;
}

}  // namespace tsc
