// tsc_index.h — host-side index object behind the C ABI (include/tostore_cuda.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tostore_cuda.h"

namespace tsc {

void set_error(const char *fmt, ...);

#define TSC_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      tsc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                     __LINE__);                                                     \
      return TSC_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

struct ScanConfig {
  int grid = 0;        // CTAs (multiple of the SM count)
  int warps = 8;       // warps per CTA
  int warps16 = 16;    // ... for 16-bit columns: twice the elements per byte, latency-bound at 8 warps
                       // (bf16 cosine 10M x 768: 2.70 ms at 8 warps, 2.26 ms at 16; fp32 is faster at 8)
  int sparse_warps = 16;   // ... of the sparse scan (K6): latency-bound per warp, wants more of them
  int rows = 0;        // R rows per stage (0 = auto)
  int stages = 0;      // S (0 = auto)
  int stage_target = 6144;  // bytes per stage aimed for when R is auto
  int inflight_target = 96 * 1024;  // bytes of bulk copies in flight per CTA when S is auto
};

// numeric table field kept column-wise next to the embedding column (tsc_where.cuh)
// dictionary of a TSC_COL_TEXT column: every distinct string once, UTF-16 code units. The host
// keeps a mirror of the arena and an open-addressing table over it (no per-string allocation:
// interning a row is one hash, one probe run and at most one memcmp); the device copy is what
// dict_match_kernel reads.
struct TextDict {
  std::vector<uint16_t> h_units;    // all strings back to back (host mirror)
  std::vector<uint64_t> h_offs{0};  // string c = h_units [h_offs[c], h_offs[c + 1])
  std::vector<uint64_t> slots;      // (hash >> 32) << 32 | code + 1; 0 = empty; power-of-two size
  uint64_t n_units = 0;             // code units / strings present on the DEVICE (the host
  uint32_t n_codes = 0;             // vectors run ahead of them only inside an append)
  uint16_t *d_units = nullptr;      // [units_cap]
  uint64_t *d_offs = nullptr;       // [codes_cap + 1]
  uint64_t units_cap = 0, codes_cap = 0;

  static uint64_t hash(const uint16_t *p, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xFF51AFD7ED558CCDull);
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
      uint64_t w;
      memcpy(&w, p + i, 8);
      h = (h ^ w) * 0x9FB21C651E98DF25ull;
      h ^= h >> 32;
    }
    uint64_t w = 0;
    if (n > i) memcpy(&w, p + i, (n - i) * 2);
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
    return h ^ (h >> 29);
  }
  uint32_t host_codes() const { return (uint32_t)(h_offs.size() - 1); }
  void rebuild() {   // table over the strings of the host mirror
    size_t cap = 1024;
    while (cap < (size_t)host_codes() * 2 + 2) cap *= 2;
    slots.assign(cap, 0);
    for (uint32_t c = 0; c < host_codes(); c++)
      place(hash(h_units.data() + h_offs[c], (size_t)(h_offs[c + 1] - h_offs[c])), c);
  }
  void place(uint64_t h, uint32_t code) {
    const size_t mask = slots.size() - 1;
    size_t i = (size_t)h & mask;
    while (slots[i]) i = (i + 1) & mask;
    slots[i] = (h & 0xFFFFFFFF00000000ull) | ((uint64_t)code + 1);
  }
  // code of the string, interning it when it is new
  uint32_t intern(const uint16_t *p, size_t n) {
    if (slots.empty() || ((size_t)host_codes() + 1) * 2 > slots.size()) rebuild();
    const uint64_t h = hash(p, n);
    const size_t mask = slots.size() - 1;
    for (size_t i = (size_t)h & mask;; i = (i + 1) & mask) {
      const uint64_t v = slots[i];
      if (!v) break;
      if ((v ^ h) >> 32) continue;
      const uint32_t c = (uint32_t)v - 1;
      if (h_offs[c + 1] - h_offs[c] == n && (n == 0 || memcmp(h_units.data() + h_offs[c], p, n * 2) == 0))
        return c;
    }
    const uint32_t c = host_codes();
    h_units.insert(h_units.end(), p, p + n);
    h_offs.push_back(h_units.size());
    place(h, c);
    return c;
  }
  void truncate(uint32_t codes) {   // forget the strings interned after the first `codes`
    if (codes >= host_codes()) return;
    h_offs.resize((size_t)codes + 1);
    h_units.resize((size_t)h_offs.back());
    rebuild();
  }
};

struct AttrColumn {
  uint32_t id = 0;
  uint8_t type = 0;              // TSC_COL_I64 / TSC_COL_F64 / TSC_COL_TEXT
  uint64_t *d_values = nullptr;  // [capacity] raw 8-byte values (text: the dictionary code)
  uint32_t *d_null = nullptr;    // [mask_words] bit r = row r is NULL (allocated on first NULL)
  uint64_t rows = 0;             // rows appended so far
  std::shared_ptr<TextDict> dict;   // text columns only
};

struct Index {
  std::mutex mu;
  tsc_index_desc desc{};
  int device = 0;
  bool host_only = false;   // self-test object without device memory (tsc_selftest_host_index)
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;

  // geometry of the embedding block in HBM: dense row-major [capacity, ld]
  uint32_t elem_bytes = 4;
  uint32_t ld = 0;          // elements per row incl. zero padding to 16 bytes
  uint32_t row_bytes = 0;
  uint32_t qld = 0;         // padded query length (== ld)
  uint8_t *d_rows = nullptr;
  // The row block lives in a reserved virtual address range that physical memory is mapped
  // into chunk by chunk (CUDA virtual memory management): growing the column adds a chunk
  // behind the mapped part, the address of every existing row stays what it was.
  unsigned long long rows_va = 0;     // CUdeviceptr of the reservation (== d_rows)
  size_t rows_va_bytes = 0, rows_mapped = 0, vmm_gran = 0;
  std::vector<std::pair<unsigned long long, size_t>> rows_chunks;   // (allocation handle, bytes)
  bool in_group = false;              // shard of a group handle: its capacity is its node-id range
  uint64_t capacity = 0, rows = 0;

  // liveness: deleted (tombstones) and filter (WHERE) bitmaps -> live
  uint64_t mask_words = 0;  // 32-bit words
  uint32_t *d_deleted = nullptr, *d_filter = nullptr, *d_live = nullptr;
  bool has_deleted = false, has_filter = false, live_dirty = false;
  uint64_t deleted_rows = 0;
  uint64_t live_rows = 0;            // popcount of `live` over [0, rows) (valid when a mask is active)
  uint64_t live_rows_for = 0;        // value of `rows` the count was taken at
  unsigned long long *d_live_count = nullptr;
  double sparse_frac = 0.35;         // live fraction below which the per-live-row scan is used
  int *d_delta = nullptr;

  // search scratch
  uint32_t k_max = 0, nq_max = 0, kprime_max = 0;
  float *d_queries = nullptr;      // [nq_max, qld]
  float *h_queries = nullptr;      // pinned
  float *d_norm2 = nullptr;        // [capacity] sum of squares per stored row
  uint32_t *d_maxnorm = nullptr;   // bits of the largest norm2 seen (atomicMax; norms are >= 0)
  float max_norm2 = 0.0f;          // host copy: upper bound of |row|^2 (certificate, IP / L2)
  int *d_progress = nullptr;       // diagnostics scratch of the tensor-core path
  uint16_t *d_q16 = nullptr;       // [nq_max, qld] queries in the storage dtype (GEMM path)
  float *d_enorm = nullptr;        // [nq_max] |q16 - q| of the converted queries (certificate)
  // batches at least this large take the tensor-core path: 9 for fp32 columns (8 queries: scan
  // 10.9 ms, tf32 GEMM 12.4), 5 for 16-bit columns (8 queries: scan 15.0 ms, GEMM 8.2; 4 queries:
  // 6.9 vs 6.4) — 10M x 768, profiles/r02_small_batches.txt
  uint32_t gemm_min_nq = 9;
  // TMA descriptors of the tensor path, encoded once per (pointer, extent) instead of per launch
  struct TmapSlot {
    alignas(64) unsigned char bytes[128];   // a CUtensorMap
    const void *ptr = nullptr;
    uint64_t rows = 0;
    uint32_t box_rows = 0;
  };
  TmapSlot tmap_b[2], tmap_q;      // corpus (CTA / CTA pair boxes), queries
  // certificate / range retry state (tsc_tail.cuh)
  uint32_t *d_flags = nullptr;     // [nq_max] kFlag* of the last search
  uint32_t *h_flags = nullptr;     // pinned mirror
  uint32_t *d_range_thr = nullptr; // [nq_max]
  uint32_t *d_retry_list = nullptr;// [nq_max]
  uint32_t *d_retry_n = nullptr;   // zero between searches
  uint32_t *d_range_count = nullptr; // [kRangeSlots], zero between launches
  uint64_t *d_range_buf = nullptr; // [kRangeSlots][kRangeCap]
  uint32_t *d_done = nullptr;      // [8] zero between launches: [0] scan ticket, [1] exchange ticket, [2] scan work counter,
                                   // [3] / [4] work counter / ticket of the range launches (they may overlap the next scan)
  unsigned long long *d_cert_stat = nullptr;  // [kStatSlots]
  unsigned long long *d_trace = nullptr;      // diagnostics build: phase timestamps (tsc_tail.cuh)
  uint32_t *d_loc_counts = nullptr;  // [nq_max] shard-local result counts of a sharded search
  // host-buffer searches: at most one in flight; `host_done` follows its last D2H copy
  cudaEvent_t host_done = nullptr;
  std::shared_ptr<struct Ticket> inflight;
  bool host_consumer = true;         // the search in flight delivers its result on this shard
  double last_threshold = 0;         // of the host-buffer search in flight (owed range passes)
  // one search at a time uses the scratch above: searches on other streams wait for this
  double last_flops = 0;               // algorithmic flops of the last timed search (tensor path), else 0
  cudaEvent_t scratch_ev = nullptr;    // recorded at the end of a search unless a timer event already is
  // Pipelined device searches (tsc_index_set_pipelining): the scan launch of search i+1 is a
  // programmatic dependent of search i's range launch and overlaps search i's tail; stream
  // events would serialise the two, so only every timer_every-th search is timed and the
  // end-of-search mark is recorded lazily (scratch_mark == nullptr: not recorded yet).
  bool pipeline = false;
  uint32_t timer_every = 1, timer_tick = 0;
  bool search_timed = true;            // the search being enqueued carries the timer events
  cudaEvent_t scratch_mark = nullptr;  // the event that marks the end of the last search (not owned)
  cudaEvent_t search_beg = nullptr;    // first timer event of the search being enqueued (not owned)
  cudaEvent_t timed_beg = nullptr, timed_end = nullptr;   // event pair of the last TIMED search (not owned)
  cudaEvent_t last_hot_end = nullptr;  // most recent hot_timer_end event (not owned)
  cudaStream_t scratch_stream = nullptr;
  bool scratch_used = false;
  uint64_t *d_cand = nullptr;      // [nq_max][cand_lists][kprime_max]
  uint64_t cand_lists = 0;
  // results of host-buffer searches: ids | dist | counts | flags are ONE device block with
  // ONE pinned mirror, so a search brings them to the host with a single copy (four small
  // D2H copies cost ~25 us of a 0.6 ms query on an 8-GPU shard)
  uint8_t *d_out_block = nullptr, *h_out_block = nullptr;
  size_t out_block_bytes = 0;
  int64_t *d_out_ids = nullptr, *h_out_ids = nullptr;
  double *d_out_dist = nullptr, *h_out_dist = nullptr;
  uint32_t *d_out_counts = nullptr, *h_out_counts = nullptr;
  // ingest staging
  uint8_t *d_stage = nullptr;
  size_t stage_bytes = 0;
  uint32_t *d_page_status = nullptr;
  size_t page_status_cap = 0;

  ScanConfig scan;

  // nodeId -> primary key side table (host memory; role of the `__nid2pk` B+Tree)
  std::vector<uint64_t> pk_off;       // per shard row: start in pk_arena
  std::vector<uint32_t> pk_len;       // per shard row: byte length, 0 = unmapped / tombstone
  std::vector<char> pk_arena;         // append-only utf-8 bytes
  // key -> shard row (the role of `<index>__pk2nid`), rebuilt from the table above when a filter
  // by primary keys needs it after the table changed
  std::unordered_map<std::string, uint64_t> pk_rev;
  bool pk_rev_valid = false;

  // attribute columns for the WHERE prefilter
  std::vector<AttrColumn> columns;
  uint64_t *d_where_args = nullptr;   // IN-list keys
  size_t where_args_cap = 0;
  uint8_t *d_where_text = nullptr;    // text leaves: operand pool, IN-list pairs, bitmaps over codes
  size_t where_text_cap = 0;

  // sharding
  void *nccl_comm = nullptr;
  int n_ranks = 1, rank = 0;
  int xroot = -1;                   // consumer rank of the exchange, -1 = every rank
  uint8_t *d_gather_send = nullptr, *d_gather_recv = nullptr;
  // exchange over peer memory (tsc_exchange.cuh): receive buffer mapped by the peers (CUDA
  // IPC between processes, peer access inside one process)
  uint8_t *d_xbuf = nullptr;
  uint64_t xbuf_bytes = 0, xslot_bytes = 0;
  uint32_t xdepth = 0;
  uint8_t *x_peer[8] = {};          // receive buffers of all ranks (own entry = d_xbuf)
  bool x_ipc = false;               // x_peer entries were opened with cudaIpcOpenMemHandle
  bool p2p_ready = false;
  uint32_t xepoch = 0;
  uint32_t *h_xstatus = nullptr;    // mapped host word, set by the kernel on timeout
  uint32_t *d_xstatus = nullptr;    // device alias of h_xstatus
  long long x_timeout_cycles = 4000000000ll;

  // dominant-kernel timing: ring of event pairs, resolved lazily
  static constexpr int kTimers = 128;
  cudaEvent_t t_beg[kTimers] = {}, t_end[kTimers] = {};
  double t_bytes[kTimers] = {}, t_flops[kTimers] = {};
  int t_head = 0, t_pending = 0;  // pending pairs are [t_head - t_pending, t_head)
  uint64_t hot_launches = 0;
  double hot_ms = 0, hot_bytes = 0, hot_flops = 0;
  int hot_slot = 0;               // timer slot of the scan launch being bracketed
  double hot_slot_bytes = 0;
  bool hot_slot_open = false;     // its end event is still owed (after the range launch)

  // stats
  uint64_t searches = 0, launches = 0;
  double last_ms = 0, last_gbs = 0;
  uint32_t last_path = 0;
  uint64_t device_bytes = 0;
};

typedef std::shared_ptr<Index> IndexRef;
void free_index(Index *ix);               // deleter of IndexRef: releases the device memory
IndexRef lookup_index(uint64_t handle);   // empty + error string when unknown (or a group)
uint64_t register_index(IndexRef ix);
int32_t ensure_stage_bytes(Index *ix, size_t bytes);
// make room for rows [0, rows_needed) (unsharded handles grow; shards answer TSC_ERR_OOM). Caller holds ix->mu.
int32_t ix_ensure_capacity(Index *ix, uint64_t rows_needed, const char *what);
int32_t refresh_live(Index *ix, cudaStream_t st);
int32_t order_after_last_search(Index *ix, cudaStream_t st);   // st waits for the last search's scratch use
int32_t sync_last_search(Index *ix);                           // host waits for it
   // recombine deleted / filter -> live when stale
int32_t launch_pad_queries(Index *ix, const float *d_queries, uint32_t nq, cudaStream_t st);

// A group: one logical column row-range sharded over the GPUs of one process
// (tsc_index_create with n_devices > 1). Shard s owns node ids
// [first + s * per_shard, first + (s + 1) * per_shard) on device_ids[s]; shard 0 is the
// root of the exchange: it merges the shards' exact top-k and the host reads from it.
// One host thread per shard 1..n-1 of a group (shard 0 is driven by the calling thread): a
// search's launches go out to all GPUs side by side instead of one device after the other
// (8 shards: ~150 us of serial launch work per query otherwise).
struct GroupWorker;
const char *last_error_text();      // this thread's error string
struct Group {
  ~Group();
  std::vector<std::unique_ptr<GroupWorker>> workers;   // [n_shards - 1]
  std::mutex mu;                    // serialises calls on the group handle
  tsc_index_desc desc{};            // as given: capacity / first_node_id of the whole column
  std::vector<IndexRef> shards;
  uint64_t per_shard = 0;           // rows per shard (multiple of 64)
  bool search_pending = false;      // grp_search_begin without grp_search_end
  uint32_t pend_nq = 0, pend_k = 0;
  std::shared_ptr<struct Ticket> inflight;
};
typedef std::shared_ptr<Group> GroupRef;
GroupRef lookup_group(uint64_t handle);   // empty (no error set) when the handle is no group
uint64_t register_group(GroupRef g);

// ---- per-shard operations behind the C ABI (each locks ix->mu itself) ------------------
int32_t ix_create(const tsc_index_desc *d, IndexRef *out);
int32_t ix_clear(Index *ix);
int32_t ix_append_rows(Index *ix, uint64_t first_node_id, const void *rows, uint64_t n_rows);
int32_t ix_append_pages(Index *ix, uint64_t first_logical_page, const uint8_t *pages,
                        uint64_t n_pages, uint32_t page_size, uint64_t live_rows);
int32_t ix_append_synthetic(Index *ix, uint64_t seed, uint64_t first_node_id, uint64_t n_rows);
int32_t ix_set_deleted(Index *ix, const uint64_t *node_ids, uint64_t n, uint8_t deleted);
int32_t ix_apply_graph_pages(Index *ix, uint64_t first_logical_page, const uint8_t *pages,
                             uint64_t n_pages, uint32_t page_size);
int32_t ix_set_filter(Index *ix, const uint64_t *bitmap_words, uint64_t n_words);
int32_t ix_stats_get(Index *ix, tsc_stats *out);
int32_t ix_stats_reset(Index *ix);
int32_t ix_column_create(Index *ix, uint32_t column_id, uint8_t col_type);
int32_t ix_column_append(Index *ix, uint32_t column_id, uint64_t first_node_id, const void *values,
                         const uint8_t *is_null, uint64_t n);
int32_t ix_column_append_text(Index *ix, uint32_t column_id, uint64_t first_node_id,
                              const uint16_t *units, const uint64_t *offsets,
                              const uint8_t *is_null, uint64_t n);
// text operands of a program: string t = text_units [text_offsets[t], text_offsets[t + 1])
struct WhereTexts {
  const uint16_t *units = nullptr;
  const uint64_t *offsets = nullptr;
  uint32_t n = 0;
};
int32_t ix_filter_where(Index *ix, const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                        uint32_t n_in_args, const WhereTexts &texts, uint64_t *out_matched);
int32_t ix_set_primary_keys(Index *ix, uint64_t first_node_id, const uint8_t *utf8,
                            const uint64_t *offsets, uint64_t n);
int32_t ix_filter_primary_keys(Index *ix, const uint8_t *utf8, const uint64_t *offsets, uint64_t n,
                               uint64_t *out_matched);
int32_t ix_get_primary_key(Index *ix, uint64_t node_id, uint8_t *out_utf8, uint32_t capacity,
                           uint32_t *out_len);
int32_t ix_load_ngh(Index *ix, const char *index_dir, uint32_t flags, tsc_ngh_info *out);
// host-buffer search in two halves: begin enqueues H2D + kernels (+ D2H where this shard
// consumes the result) on the shard's stream, end waits, runs owed range passes and copies
// out. `queries` are [nq, dims] fp32 in host memory. ix->mu is NOT taken: the caller holds
// it (blocking search) or the group's lock.
int32_t ix_search_begin(Index *ix, const float *queries, uint32_t nq, uint32_t k, double threshold);
int32_t ix_search_end(Index *ix, uint32_t nq, uint32_t k, int64_t *out_ids, double *out_dist,
                      uint32_t *out_counts);
bool ix_is_consumer(const Index *ix);
void drop_tickets_of(uint64_t handle);    // tsc_index_destroy: tickets die with their handle
void comm_release(Index *ix);             // destroy the NCCL communicator, if any
// peer-memory exchange between the shards of one process (tsc_group.cu)
int32_t ix_xchg_alloc(Index *ix, int n_ranks, int rank, int root);

// ---- group operations (tsc_group.cu) -----------------------------------------------------
int32_t grp_create(const tsc_index_desc *d, uint64_t *out_handle);
int32_t grp_clear(Group &g);
int32_t grp_append_rows(Group &g, uint64_t first_node_id, const void *rows, uint64_t n_rows);
int32_t grp_append_pages(Group &g, uint64_t first_logical_page, const uint8_t *pages,
                         uint64_t n_pages, uint32_t page_size, uint64_t live_rows);
int32_t grp_append_synthetic(Group &g, uint64_t seed, uint64_t first_node_id, uint64_t n_rows);
int32_t grp_set_deleted(Group &g, const uint64_t *node_ids, uint64_t n, uint8_t deleted);
int32_t grp_apply_graph_pages(Group &g, uint64_t first_logical_page, const uint8_t *pages,
                              uint64_t n_pages, uint32_t page_size);
int32_t grp_set_filter(Group &g, const uint64_t *bitmap_words, uint64_t n_words);
int32_t grp_stats_get(Group &g, tsc_stats *out);
int32_t grp_stats_reset(Group &g);
int32_t grp_column_create(Group &g, uint32_t column_id, uint8_t col_type);
int32_t grp_column_append(Group &g, uint32_t column_id, uint64_t first_node_id, const void *values,
                          const uint8_t *is_null, uint64_t n);
int32_t grp_column_append_text(Group &g, uint32_t column_id, uint64_t first_node_id,
                               const uint16_t *units, const uint64_t *offsets,
                               const uint8_t *is_null, uint64_t n);
int32_t grp_filter_where(Group &g, const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                         uint32_t n_in_args, const WhereTexts &texts, uint64_t *out_matched);
int32_t grp_set_primary_keys(Group &g, uint64_t first_node_id, const uint8_t *utf8,
                             const uint64_t *offsets, uint64_t n);
int32_t grp_filter_primary_keys(Group &g, const uint8_t *utf8, const uint64_t *offsets, uint64_t n,
                                uint64_t *out_matched);
int32_t grp_get_primary_key(Group &g, uint64_t node_id, uint8_t *out_utf8, uint32_t capacity,
                            uint32_t *out_len);
int32_t grp_load_ngh(Group &g, const char *index_dir, uint32_t flags, tsc_ngh_info *out);
int32_t grp_search_begin(Group &g, const float *queries, uint32_t nq, uint32_t k, double threshold);
int32_t grp_search_end(Group &g, int64_t *out_ids, double *out_dist, uint32_t *out_counts);
int32_t grp_search_flags(Group &g, uint32_t nq, uint32_t *out_flags);

// every extern "C" body runs inside this: no exception crosses the C ABI
#define TSC_API_TRY try {
#define TSC_API_CATCH                                          \
  }                                                            \
  catch (const std::bad_alloc &) {                             \
    tsc::set_error("out of host memory");                      \
    return TSC_ERR_OOM;                                        \
  }                                                            \
  catch (...) {                                                \
    tsc::set_error("internal error (C++ exception)");          \
    return TSC_ERR_BAD_ARG;                                    \
  }

// One search as the launchers see it: the shard-local exact top-k goes to loc_*; for a
// sharded index the exchange merges the shards' loc_* into x_* (on the consumer ranks).
struct SearchCtx {
  const float *d_q = nullptr;    // [nq, qld] fp32, padded
  uint32_t nq = 0, k = 0, kprime = 0;
  uint32_t gemm_list_kp = 0;     // tensor path: entries per candidate list (<= kprime)
  bool pipelined = false;        // device-buffer search of a handle with pipelining on
  double threshold = 0;
  int64_t *loc_ids = nullptr;
  double *loc_dist = nullptr;
  uint32_t *loc_counts = nullptr;
  bool sharded = false;
  int64_t *x_ids = nullptr;
  double *x_dist = nullptr;
  uint32_t *x_counts = nullptr;
  cudaStream_t st = nullptr;
};

struct TailParams;
struct XchgParams;
// m candidates per query in ix->d_cand; gemm_keys: they come from the tensor path
void fill_tail(const Index *ix, const SearchCtx &c, uint32_t m, bool gemm_keys, TailParams *out);
void fill_xchg(const Index *ix, XchgParams *out);   // uses ix->xepoch as it stands

// launchers implemented per translation unit
// mode 0: first pass over queries [q_base, q_base + n) (n <= 8); mode 1: range pass over
// retry-list entries [q_base, q_base + qb). fused: the last CTA runs the tail (and, with
// xchg, the shard exchange of the whole search).
int32_t launch_scan(Index *ix, const SearchCtx &c, int mode, uint32_t q_base, uint32_t n, int qb,
                    bool fused, bool xchg, bool last_retry, uint32_t *out_lists);
int32_t launch_tail(Index *ix, const SearchCtx &c, uint32_t m, bool gemm_keys);
int32_t launch_merge(Index *ix, const int64_t *d_part_ids, const double *d_part_dist,
                     uint64_t part_stride, uint32_t n_parts, uint32_t nq, uint32_t k,
                     int64_t *d_ids, double *d_dist, uint32_t *d_counts, cudaStream_t st);
int32_t launch_exchange(Index *ix, const SearchCtx &c);
int32_t scan_configure(Index *ix);
bool gemm_supported(const Index *ix, uint32_t kprime);
int32_t gemm_update_norms(Index *ix, uint64_t first_row, uint64_t n, cudaStream_t st);
int32_t launch_gemm(Index *ix, const float *d_q, uint32_t nq, uint32_t kprime, uint64_t *d_cand,
                    uint32_t *out_lists, float *dbg_keys, cudaStream_t st);
// bracket one launch of the dominant kernel with events on `st`
int32_t hot_timer_begin(Index *ix, cudaStream_t st, int *slot);
int32_t hot_timer_end(Index *ix, cudaStream_t st, int slot, double bytes, double flops);
int32_t hot_timer_resolve(Index *ix);

inline uint32_t kprime_for(uint32_t k) {
  // rerankCount = max(2k, 20), core/ngh_graph_engine.dart:115; capped margin for large k
  uint32_t kp = k <= 64 ? (2 * k > 20 ? 2 * k : 20) : k + 64;
  return kp;
}

inline uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace tsc
