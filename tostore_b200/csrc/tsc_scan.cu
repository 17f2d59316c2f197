// tsc_scan.cu — host launcher for K1 / K6 (tsc_scan.cuh): stage geometry + dispatch.
#include <stdlib.h>

#include "tsc_index.h"
#include "tsc_scan.cuh"

namespace tsc {

#ifdef TSC_DIAG
static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}
#else
// the shipping library reads no tuning switches from the environment
static int env_int(const char *, int dflt) { return dflt; }
#endif

int32_t scan_configure(Index *ix) {
  ix->scan.grid = env_int("TSC_SCAN_CTAS", 1) * ix->sm_count;
  ix->scan.warps = env_int("TSC_SCAN_WARPS", 8);
  ix->scan.rows = env_int("TSC_SCAN_ROWS", 0);
  ix->scan.stages = env_int("TSC_SCAN_STAGES", 0);
  ix->scan.stage_target = env_int("TSC_SCAN_STAGE_BYTES", 6144);
  ix->scan.inflight_target = env_int("TSC_SCAN_INFLIGHT_BYTES", 96 * 1024);
  if (ix->scan.warps < 1 || ix->scan.warps > 16 || ix->scan.grid < 1) {
    set_error("bad TSC_SCAN_* override");
    return TSC_ERR_BAD_ARG;
  }
  return TSC_OK;
}

struct ScanPlan {
  int warps, rows, stages, qb;
  uint32_t stage_bytes, sort_cap;
  size_t smem;
};

static bool plan_scan(const Index *ix, int qb, uint32_t kprime, ScanPlan *pl,
                      bool sparse = false) {
  const size_t budget = ix->smem_optin - 1024;
  for (int w = ix->scan.warps; w >= 1; w >>= 1) {
    int r = sparse ? 1 : ix->scan.rows;  // sparse: one live row per stage
    if (r <= 0) {
      r = 8;
      while (r > 1 && (size_t)r * ix->row_bytes > (size_t)ix->scan.stage_target) r >>= 1;
    }
    while (r > 1 && r * qb > 16) r >>= 1;
    uint32_t sort_cap = next_pow2((uint32_t)w * kprime);
    if (sort_cap < 2) sort_cap = 2;
    size_t fixed = scan_smem_query_bytes(qb, ix->qld) + scan_smem_sort_bytes(sort_cap);
    if (fixed >= budget) continue;
    size_t per_warp = ((budget - fixed) / w) & ~(size_t)127;
    for (; r >= 1; r >>= 1) {
      uint32_t stage_bytes = (uint32_t)r * ix->row_bytes;
      // Measured on B200 (profiles/r01_scan_sweep.txt): ~96 KB of bulk copies in
      // flight per SM is the sweet spot (7.36 TB/s); deeper rings lose 5-10 %.
      int smax = ix->scan.stages;
      if (smax <= 0) {
        smax = (int)((ix->scan.inflight_target + (size_t)w * stage_bytes / 2) /
                     ((size_t)w * stage_bytes));
        if (smax < 2) smax = 2;
        if (smax > (sparse ? 16 : 8)) smax = sparse ? 16 : 8;
      }
      int s = smax;
      while (s >= 2 && scan_smem_warp_bytes(qb, kprime, s, stage_bytes) > per_warp) s--;
      if (s >= 2) {
        pl->warps = w;
        pl->rows = r;
        pl->stages = s;
        pl->qb = qb;
        pl->stage_bytes = stage_bytes;
        pl->sort_cap = sort_cap;
        pl->smem = fixed + (size_t)w * scan_smem_warp_bytes(qb, kprime, s, stage_bytes);
        return true;
      }
    }
  }
  return false;
}

template <int METRIC, int DTYPE, int QB, int R>
static int32_t run_scan(Index *ix, const ScanParams &p, const ScanPlan &pl, cudaStream_t st) {
  static bool attr_done[64] = {false};
  auto kern = scan_topk_kernel<METRIC, DTYPE, QB, R>;
  if (!attr_done[ix->device & 63]) {
    // the fused tail has a few static __shared__ words: the dynamic part stops 1 KB short
    TSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ix->smem_optin - 1024));
    attr_done[ix->device & 63] = true;
  }
  int slot = 0;
  int32_t rc = TSC_OK;
  if (p.mode == 0) rc = hot_timer_begin(ix, st, &slot);
  if (rc != TSC_OK) return rc;
  kern<<<ix->scan.grid, pl.warps * 32, pl.smem, st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  if (p.mode != 0) return TSC_OK;   // range launches are not part of the roofline accounting
  // algorithmic bytes of one pass: live rows x dims x sizeof(elem)
  return hot_timer_end(ix, st, slot, (double)p.n_rows * ix->desc.dims * ix->elem_bytes, 0.0);
}

template <int METRIC, int DTYPE, int QB>
static int32_t run_sparse(Index *ix, const ScanParams &p, const ScanPlan &pl, cudaStream_t st) {
  static bool attr_done[64] = {false};
  auto kern = scan_topk_sparse_kernel<METRIC, DTYPE, QB>;
  if (!attr_done[ix->device & 63]) {
    // the fused tail has a few static __shared__ words: the dynamic part stops 1 KB short
    TSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ix->smem_optin - 1024));
    attr_done[ix->device & 63] = true;
  }
  int slot = 0;
  int32_t rc = TSC_OK;
  if (p.mode == 0) rc = hot_timer_begin(ix, st, &slot);
  if (rc != TSC_OK) return rc;
  kern<<<ix->scan.grid, pl.warps * 32, pl.smem, st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  if (p.mode != 0) return TSC_OK;
  // algorithmic bytes: only live rows are read (+ the bitmap)
  return hot_timer_end(ix, st, slot,
                       (double)ix->live_rows * ix->desc.dims * ix->elem_bytes + p.n_rows / 8.0, 0.0);
}

template <int METRIC, int DTYPE>
static int32_t dispatch_sparse(Index *ix, const ScanParams &p, const ScanPlan &pl,
                               cudaStream_t st) {
  if (pl.qb == 1) return run_sparse<METRIC, DTYPE, 1>(ix, p, pl, st);
  if (pl.qb == 4) return run_sparse<METRIC, DTYPE, 4>(ix, p, pl, st);
  return run_sparse<METRIC, DTYPE, 8>(ix, p, pl, st);
}

template <int METRIC>
static int32_t dispatch_sparse_dtype(Index *ix, const ScanParams &p, const ScanPlan &pl,
                                     cudaStream_t st) {
  switch (ix->desc.dev_dtype) {
    case TSC_DEV_F32: return dispatch_sparse<METRIC, kF32>(ix, p, pl, st);
    case TSC_DEV_BF16: return dispatch_sparse<METRIC, kBF16>(ix, p, pl, st);
    default: return dispatch_sparse<METRIC, kF16>(ix, p, pl, st);
  }
}

template <int METRIC, int DTYPE>
static int32_t dispatch_qr(Index *ix, const ScanParams &p, const ScanPlan &pl, cudaStream_t st) {
#define TSC_CASE(QB, R) \
  if (pl.qb == QB && pl.rows == R) return run_scan<METRIC, DTYPE, QB, R>(ix, p, pl, st);
  TSC_CASE(1, 1) TSC_CASE(1, 2) TSC_CASE(1, 4) TSC_CASE(1, 8)
  TSC_CASE(4, 1) TSC_CASE(4, 2) TSC_CASE(4, 4)
  TSC_CASE(8, 1) TSC_CASE(8, 2)
#undef TSC_CASE
  set_error("scan: no kernel for qb=%d rows=%d", pl.qb, pl.rows);
  return TSC_ERR_UNSUPPORTED;
}

template <int METRIC>
static int32_t dispatch_dtype(Index *ix, const ScanParams &p, const ScanPlan &pl,
                              cudaStream_t st) {
  switch (ix->desc.dev_dtype) {
    case TSC_DEV_F32: return dispatch_qr<METRIC, kF32>(ix, p, pl, st);
    case TSC_DEV_BF16: return dispatch_qr<METRIC, kBF16>(ix, p, pl, st);
    default: return dispatch_qr<METRIC, kF16>(ix, p, pl, st);
  }
}

// One scan launch over the shard. mode 0: first pass for queries [q_base, q_base + n), the
// kernel variant `qb` (1 / 4 / 8 queries per pass); d_cand receives [nq][grid][kprime]
// composites, *out_lists = grid. mode 1: range pass over retry-list entries
// [q_base, q_base + qb); exits at once on the device when there are none.
int32_t launch_scan(Index *ix, const SearchCtx &c, int mode, uint32_t q_base, uint32_t n, int qb,
                    bool fused, bool xchg, bool last_retry, uint32_t *out_lists) {
  if (out_lists) *out_lists = (uint32_t)ix->scan.grid;
  ScanPlan pl;
  // sparse liveness (WHERE prefilter): move only the live rows, one bulk copy each
  const bool masked = ix->has_deleted || ix->has_filter;
  const bool sparse = masked && ix->row_bytes >= 256 &&
                      (double)ix->live_rows < ix->sparse_frac * (double)ix->rows;
  if (!plan_scan(ix, qb, c.kprime, &pl, sparse)) {
    set_error("scan: dims=%u (row %u B) with k'=%u does not fit shared memory", ix->desc.dims,
              ix->row_bytes, c.kprime);
    return TSC_ERR_BAD_DIMS;
  }
  ScanParams p{};
  p.rows = ix->d_rows;
  p.n_rows = ix->rows;
  p.row_bytes = ix->row_bytes;
  p.chunks_per_row = ix->row_bytes / 16;
  p.queries = c.d_q;
  p.qld = ix->qld;
  p.q_base = q_base;
  p.nq = n;
  p.live_mask = masked ? ix->d_live : nullptr;
  p.kprime = c.kprime;
  p.stages = (uint32_t)pl.stages;
  p.stage_bytes = pl.stage_bytes;
  p.cand = ix->d_cand;
  p.sort_cap = pl.sort_cap;
  p.mode = mode;
  p.fused_tail = (fused || mode == 1) ? 1 : 0;
  p.xchg_in_tail = xchg ? 1 : 0;
  p.last_retry = last_retry ? 1 : 0;
  p.nq_total = c.nq;
  p.done_counter = ix->d_done;
  const uint32_t m = (uint32_t)ix->scan.grid * c.kprime;
  fill_tail(ix, c, m, false, &p.tail);
  p.tail_sort_cap = tail_sort_cap(m, c.kprime, mode == 1);
  if (xchg) {
    fill_xchg(ix, &p.xchg);
    p.x_out_ids = c.x_ids;
    p.x_out_dist = c.x_dist;
    p.x_out_counts = c.x_counts;
    uint32_t xcap = next_pow2((uint32_t)ix->n_ranks * c.k);
    if (xcap > p.tail_sort_cap) p.tail_sort_cap = xcap;
  }
  if (p.fused_tail) {
    // room to stage the candidates' rows for the re-rank: all of them when that fits
    const size_t lim = ix->smem_optin - 1024;
    uint32_t rows_staged = mode == 1 ? 64 : c.kprime;
    while (rows_staged > 1 &&
           tail_smem_bytes(p.tail_sort_cap, ix->qld, ix->row_bytes, rows_staged) > lim)
      rows_staged >>= 1;
    const size_t need = tail_smem_bytes(p.tail_sort_cap, ix->qld, ix->row_bytes, rows_staged);
    if (need > ix->smem_optin - 1024) {
      set_error("scan: the tail needs %zu bytes of shared memory", need);
      return TSC_ERR_BAD_DIMS;
    }
    if (need > pl.smem) pl.smem = need;
  }
  p.smem_bytes = (uint32_t)pl.smem;
  int32_t rc;
  if (sparse) {
    switch (ix->desc.metric) {
      case TSC_METRIC_L2: rc = dispatch_sparse_dtype<kL2>(ix, p, pl, c.st); break;
      case TSC_METRIC_INNER_PRODUCT: rc = dispatch_sparse_dtype<kIP>(ix, p, pl, c.st); break;
      default: rc = dispatch_sparse_dtype<kCos>(ix, p, pl, c.st); break;
    }
    return rc;
  }
  switch (ix->desc.metric) {
    case TSC_METRIC_L2: rc = dispatch_dtype<kL2>(ix, p, pl, c.st); break;
    case TSC_METRIC_INNER_PRODUCT: rc = dispatch_dtype<kIP>(ix, p, pl, c.st); break;
    default: rc = dispatch_dtype<kCos>(ix, p, pl, c.st); break;
  }
  return rc;
}

}  // namespace tsc
