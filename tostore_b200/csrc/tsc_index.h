// tsc_index.h — host-side index object behind the C ABI (include/tostore_cuda.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
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
  int rows = 0;        // R rows per stage (0 = auto)
  int stages = 0;      // S (0 = auto)
  int stage_target = 6144;  // bytes per stage aimed for when R is auto
  int inflight_target = 96 * 1024;  // bytes of bulk copies in flight per CTA when S is auto
  bool sparse_pf = false;   // sparse scan with the prefetching block cursor (opt-in)
};

// numeric table field kept column-wise next to the embedding column (tsc_where.cuh)
struct AttrColumn {
  uint32_t id = 0;
  uint8_t type = 0;              // TSC_COL_I64 / TSC_COL_F64
  uint64_t *d_values = nullptr;  // [capacity] raw 8-byte values
  uint32_t *d_null = nullptr;    // [mask_words] bit r = row r is NULL (allocated on first NULL)
  uint64_t rows = 0;             // rows appended so far
};

struct Index {
  std::mutex mu;
  tsc_index_desc desc{};
  int device = 0;
  bool host_only = false;   // self-test object without device memory (tsc_selftest_host_index)
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // geometry of the embedding block in HBM: dense row-major [capacity, ld]
  uint32_t elem_bytes = 4;
  uint32_t ld = 0;          // elements per row incl. zero padding to 16 bytes
  uint32_t row_bytes = 0;
  uint32_t qld = 0;         // padded query length (== ld)
  uint8_t *d_rows = nullptr;
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
  float *d_norm2 = nullptr;        // [capacity] sum of squares per stored row (16-bit dtypes)
  int *d_progress = nullptr;       // lockstep counters of the tensor-core path
  uint16_t *d_q16 = nullptr;       // [nq_max, qld] queries in the storage dtype (GEMM path)
  uint32_t gemm_min_nq = 9;        // batches at least this large take the tensor-core path
  bool tf32 = false;               // fp32 column with the opt-in tf32 tensor path (TSC_GEMM_TF32=1)
  uint64_t *d_cand = nullptr;      // [nq_max][cand_lists][kprime_max]
  uint64_t cand_lists = 0;
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

  // attribute columns for the WHERE prefilter
  std::vector<AttrColumn> columns;
  uint64_t *d_where_args = nullptr;   // IN-list keys
  size_t where_args_cap = 0;

  // sharding
  void *nccl_comm = nullptr;
  int n_ranks = 1, rank = 0;
  uint8_t *d_gather_send = nullptr, *d_gather_recv = nullptr;
  // opt-in P2P exchange (tsc_exchange.cuh): receive buffer exported to the peers over CUDA IPC
  uint8_t *d_xbuf = nullptr;
  uint64_t xbuf_bytes = 0, xslot_bytes = 0;
  uint8_t *x_peer[8] = {};          // receive buffers of all ranks (own entry = d_xbuf)
  bool p2p_ready = false;
  uint32_t xepoch = 0;
  uint32_t *h_xstatus = nullptr;    // mapped host word, set by the kernel on timeout
  uint32_t *d_xstatus = nullptr;    // device alias of h_xstatus

  // dominant-kernel timing: ring of event pairs, resolved lazily
  static constexpr int kTimers = 128;
  cudaEvent_t t_beg[kTimers] = {}, t_end[kTimers] = {};
  double t_bytes[kTimers] = {}, t_flops[kTimers] = {};
  int t_head = 0, t_pending = 0;  // pending pairs are [t_head - t_pending, t_head)
  uint64_t hot_launches = 0;
  double hot_ms = 0, hot_bytes = 0, hot_flops = 0;

  // stats
  uint64_t searches = 0, launches = 0;
  double last_ms = 0, last_gbs = 0;
  uint32_t last_path = 0;
  uint64_t device_bytes = 0;
};

Index *lookup_index(uint64_t handle);   // NULL + error string when unknown
uint64_t register_index(Index *ix);
int32_t ensure_stage_bytes(Index *ix, size_t bytes);

// launchers implemented per translation unit
int32_t launch_scan(Index *ix, const float *d_q, uint32_t nq, uint32_t kprime, uint64_t *d_cand,
                    uint32_t *out_lists, cudaStream_t st);
int32_t launch_select(Index *ix, const float *d_q, uint32_t nq, uint32_t k, uint32_t kprime,
                      const uint64_t *d_cand, uint32_t m, double threshold, int64_t *d_ids,
                      double *d_dist, uint32_t *d_counts, cudaStream_t st);
int32_t launch_merge(Index *ix, const int64_t *d_part_ids, const double *d_part_dist,
                     uint64_t part_stride, uint32_t n_parts, uint32_t nq, uint32_t k,
                     int64_t *d_ids, double *d_dist, uint32_t *d_counts, cudaStream_t st);
int32_t launch_exchange(Index *ix, const int64_t *d_src_ids, const double *d_src_dist, uint32_t nq,
                        uint32_t k, int64_t *d_ids, double *d_dist, uint32_t *d_counts,
                        cudaStream_t st);
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
