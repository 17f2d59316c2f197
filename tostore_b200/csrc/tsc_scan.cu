// tsc_scan.cu — host launcher for K1 / K6 (tsc_scan.cuh): stage geometry + dispatch.
#include <stdlib.h>

#include "tsc_index.h"
#include "tsc_scan.cuh"
#include "tsc_scan_launch.h"

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
  // sparse scan: 16 warps per SM read 5.19 TB/s of live rows at 10 % density, 8 warps 3.89
  // (profiles/r02_sparse_warps.txt; the plain-load gather ceiling is 5.70)
  ix->scan.sparse_warps = env_int("TSC_SCAN_SPARSE_WARPS", 16);
  ix->scan.warps16 = env_int("TSC_SCAN_WARPS16", env_int("TSC_SCAN_WARPS", 16));
  ix->gemm_min_nq = (uint32_t)env_int("TSC_GEMM_MIN_NQ", (int)ix->gemm_min_nq);   // diagnostics build only
  ix->scan.rows = env_int("TSC_SCAN_ROWS", 0);
  ix->scan.stages = env_int("TSC_SCAN_STAGES", 0);
  ix->scan.stage_target = env_int("TSC_SCAN_STAGE_BYTES", 6144);
  ix->scan.inflight_target = env_int("TSC_SCAN_INFLIGHT_BYTES", 96 * 1024);
  if (ix->scan.warps < 1 || ix->scan.warps > 16 || ix->scan.grid < 1 || ix->scan.sparse_warps < 1 ||
      ix->scan.sparse_warps > 16 || ix->scan.warps16 < 1 || ix->scan.warps16 > 16) {
    set_error("bad TSC_SCAN_* override");
    return TSC_ERR_BAD_ARG;
  }
  return TSC_OK;
}

static bool plan_scan(const Index *ix, int qb, uint32_t kprime, ScanPlan *pl,
                      bool sparse = false) {
  const size_t budget = ix->smem_optin - 1024;
  const int w0 = sparse ? ix->scan.sparse_warps : (ix->elem_bytes == 2 ? ix->scan.warps16 : ix->scan.warps);
  for (int w = w0; w >= 1; w >>= 1) {
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
  p.no_range = (mode == 0 && p.fused_tail && c.pipelined) ? 1 : 0;
  p.nq_total = c.nq;
  // first passes and range launches keep separate tickets and work counters
  p.done_counter = ix->d_done + (mode == 1 ? 4 : 0);
  p.work_counter = ix->d_done + (mode == 1 ? 3 : 2);
  {
    // 7/8 of the stages round-robin (a whole number of rounds), the rest on demand
    const uint64_t total = (ix->rows + (uint64_t)pl.rows - 1) / (uint64_t)pl.rows;
    const uint64_t gw = (uint64_t)ix->scan.grid * pl.warps;
    p.static_stages = (total - total / 8) / gw * gw;
    // small shards: the demand-driven part would be a handful of atomics per warp, all latency
    if (total / gw < 16) p.static_stages = total;
  }
  const uint32_t m = (uint32_t)ix->scan.grid * c.kprime;
  fill_tail(ix, c, m, false, &p.tail);
  p.tail.defer_retry = p.no_range;
  p.tail_sort_cap = tail_sort_cap(m, c.kprime, p.tail.list_len, mode == 1);
  if (xchg) {
    fill_xchg(ix, &p.xchg);
    p.x_out_ids = c.x_ids;
    p.x_out_dist = c.x_dist;
    p.x_out_counts = c.x_counts;
    uint32_t xcap = next_pow2((uint32_t)ix->n_ranks * c.k);
    if (xcap > p.tail_sort_cap) p.tail_sort_cap = xcap;
  }
  if (p.fused_tail) {
    // the re-rank stages the candidates' rows in shared memory: whole rows for all K' + 1
    // chains when that fits (the scan's ring is usually larger already), else column chunks
    const size_t lim = ix->smem_optin - 1024;
    const uint32_t chains = mode == 1 ? (uint32_t)pl.warps * 32u : c.kprime + 1;
    const size_t need = tail_smem_bytes(p.tail_sort_cap, ix->qld, ix->row_bytes, chains, lim);
    if (need > lim) {
      set_error("scan: the tail needs %zu bytes of shared memory", need);
      return TSC_ERR_BAD_DIMS;
    }
    if (need > pl.smem) pl.smem = need;
    // the selection's scratch (list prefixes + the entries that pass the head pivot)
    const uint32_t lvl = (c.kprime + (uint32_t)ix->scan.grid - 1) / (uint32_t)ix->scan.grid;
    const size_t sel = tail_fixed_bytes(p.tail_sort_cap, ix->qld) +
                       ((size_t)ix->scan.grid * (4 * lvl < 8 ? 8 : 4 * lvl) + kHeadSelectCap) * 8;
    if (sel > pl.smem && sel <= lim) pl.smem = sel;
  }
  p.smem_bytes = (uint32_t)pl.smem;
  // the kernels live in one translation unit per metric (tsc_scan_l2 / _ip / _cos.cu)
  switch (ix->desc.metric) {
    case TSC_METRIC_L2: return scan_dispatch_l2(ix, p, pl, sparse, c.st);
    case TSC_METRIC_INNER_PRODUCT: return scan_dispatch_ip(ix, p, pl, sparse, c.st);
    default: return scan_dispatch_cos(ix, p, pl, sparse, c.st);
  }
}

}  // namespace tsc
