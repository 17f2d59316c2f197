// tsc_tail.cu — host side of K5 / K9 (tsc_tail.cuh, tsc_exchange.cuh): the error models of
// the certificate, the standalone tail / exchange / merge launchers.
#include <math.h>
#include <stdlib.h>

#include "tsc_exchange.cuh"
#include "tsc_index.h"

namespace tsc {

// ---- certificate: proven key-space error of each candidate stage ------------
// u = 2^-24 (fp32 unit roundoff). A lane of the scan kernel adds `terms` products with fma
// (one rounding each), five butterfly additions follow, L2 squares a rounded difference:
// |sum_computed - sum| <= gamma(terms + 8) * sum|term|; for L2 every term is >= 0 so the
// bound is relative, for dot products it is absolute through Cauchy-Schwarz.
static double scan_eps(const Index *ix) {
  const double u = 5.9604644775390625e-8;
  const uint32_t cpr = ix->row_bytes / 16, e = 16 / ix->elem_bytes;
  const double terms = (double)((cpr + 31) / 32) * e;
  return (terms + 12.0) * u * 1.01;
}
static CertModel cert_scan(const Index *ix) {
  CertModel m{0, 0, 0, 0, 0};
  const double eps = scan_eps(ix);
  const double bmax = sqrt((double)ix->max_norm2) * 1.001;
  switch (ix->desc.metric) {
    case TSC_METRIC_L2: m.c_rel = (float)eps; break;
    case TSC_METRIC_INNER_PRODUCT: m.a_q = (float)(eps * bmax); break;
    default: m.a_q = (float)(2.0 * eps + 9.5367431640625e-7); break;  // + rsqrtf, product rounding
  }
  return m;
}
// Tensor path: the query is rounded to the storage type (|q16 - q| enters through
// Cauchy-Schwarz, measured per query), products of 16-bit values are exact, the fp32
// accumulation of the tensor core is bounded by one ulp per addition (eps_mma). tf32 reads
// both fp32 operands truncated to 10 mantissa bits: every product is off by < 2^-9 relative.
static CertModel cert_gemm(const Index *ix) {
  CertModel m{0, 0, 0, 0, 0};
  const double eps_n = scan_eps(ix);   // row_norms_kernel accumulates like the scan
  const double eps_mma = ((double)ix->desc.dims + 16.0) * 1.1920928955078125e-7;
  const bool tf32 = ix->desc.dev_dtype == TSC_DEV_F32;
  const double e_in = tf32 ? 1.953125e-3 * 1.01 : 0.0;   // operand truncation (tf32 only)
  const double a_e = tf32 ? 0.0 : 1.001;                 // multiplies |q16 - q|
  const double bmax = sqrt((double)ix->max_norm2) * 1.001;
  const double tiny = 4.76837158203125e-7;               // coefficient / fma rounding
  switch (ix->desc.metric) {
    case TSC_METRIC_L2:   // key = |b|^2 - 2 q.b  (the constant |q|^2 is omitted)
      m.l2_shift = 1;
      m.a_q = (float)(2.0 * bmax * (eps_mma + e_in + tiny));
      m.a_e = (float)(2.0 * bmax * a_e);
      m.a_0 = (float)((eps_n + tiny) * bmax * bmax);
      break;
    case TSC_METRIC_INNER_PRODUCT:
      m.a_q = (float)(bmax * (eps_mma + e_in + tiny));
      m.a_e = (float)(bmax * a_e);
      break;
    default:              // key = -q.b / |b|
      m.a_q = (float)(eps_mma + e_in + eps_n + 2.0 * tiny);
      m.a_e = (float)a_e;
      break;
  }
  return m;
}

void fill_tail(const Index *ix, const SearchCtx &c, uint32_t m, bool gemm_keys, TailParams *t) {
  t->cand = ix->d_cand;
  t->m = m;
  t->list_len = gemm_keys ? 0 : c.kprime;   // scan lists are sorted ascending, K' entries each
  // tensor-path lists hold at most kGemmMaxKp entries: shorter than K' means rows were dropped
  // per list, which the certificate has to account for
  t->trunc_len = (gemm_keys && c.gemm_list_kp < c.kprime) ? c.gemm_list_kp : 0;
  t->kprime = c.kprime;
  t->k = c.k;
  t->rows = ix->d_rows;
  t->row_bytes = ix->row_bytes;
  t->dims = ix->desc.dims;
  t->queries = c.d_q;
  t->qld = ix->qld;
  t->threshold = c.threshold;
  t->first_node_id = (int64_t)ix->desc.first_node_id;
  t->out_ids = c.loc_ids;
  t->out_dist = c.loc_dist;
  t->out_counts = c.loc_counts;
  t->cert_range = cert_scan(ix);
  t->cert = gemm_keys ? cert_gemm(ix) : t->cert_range;
  t->enorm = (gemm_keys && ix->desc.dev_dtype != TSC_DEV_F32) ? ix->d_enorm : nullptr;
  t->flags = ix->d_flags;
  t->range_thr = ix->d_range_thr;
  t->retry_list = ix->d_retry_list;
  t->retry_n = ix->d_retry_n;
  t->range_count = ix->d_range_count;
  t->range_buf = ix->d_range_buf;
  t->stat = ix->d_cert_stat;
  t->diag = nullptr;
#ifdef TSC_DIAG
  t->diag = ix->d_trace;   // tools/scan_trace.py
#endif
}

void fill_xchg(const Index *ix, XchgParams *x) {
  for (int r = 0; r < kMaxRanks; r++) x->peer[r] = r < ix->n_ranks ? ix->x_peer[r] : nullptr;
  x->n_ranks = (uint32_t)ix->n_ranks;
  x->rank = (uint32_t)ix->rank;
  x->root = ix->xroot;
  x->k_stride = ix->k_max;
  x->nq_max = ix->nq_max;
  x->depth = ix->xdepth;
  x->slot_bytes = ix->xslot_bytes;
  x->flag_off = xchg_flag_off(x->n_ranks, x->depth, x->slot_bytes);
  x->ack_off = xchg_ack_off(x->n_ranks, x->depth, x->slot_bytes, x->nq_max);
  x->epoch = ix->xepoch;
  x->status = ix->d_xstatus;
  x->timeout_cycles = ix->x_timeout_cycles;
}

template <int METRIC, int DTYPE>
static int32_t run_tail(Index *ix, const TailParams &p, uint32_t nq, uint32_t sort_cap,
                        cudaStream_t st) {
  static bool attr_done[64] = {false};
  auto kern = tail_kernel<METRIC, DTYPE>;
  // whole candidate rows for all K' + 1 chains when that fits in ~100 KB (two CTAs per SM),
  // column chunks otherwise
  const size_t smem = tail_smem_bytes(sort_cap, p.qld, p.row_bytes, p.kprime + 1, 100 * 1024);
  if (!attr_done[ix->device & 63]) {
    TSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ix->smem_optin - 4 * 1024));
    attr_done[ix->device & 63] = true;
  }
  if (smem > ix->smem_optin - 4 * 1024) {
    set_error("tail: dims too large for the re-rank staging (%zu B)", smem);
    return TSC_ERR_BAD_DIMS;
  }
  kern<<<nq, kTailThreads, smem, st>>>(p, sort_cap, (uint32_t)smem);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

template <int METRIC>
static int32_t run_tail_dtype(Index *ix, const TailParams &p, uint32_t nq, uint32_t sort_cap,
                              cudaStream_t st) {
  switch (ix->desc.dev_dtype) {
    case TSC_DEV_F32: return run_tail<METRIC, kF32>(ix, p, nq, sort_cap, st);
    case TSC_DEV_BF16: return run_tail<METRIC, kBF16>(ix, p, nq, sort_cap, st);
    default: return run_tail<METRIC, kF16>(ix, p, nq, sort_cap, st);
  }
}

// standalone tail over all queries of the search: one CTA per query
int32_t launch_tail(Index *ix, const SearchCtx &c, uint32_t m, bool gemm_keys) {
  if (c.kprime > kMaxRerank) {
    set_error("tail: k'=%u exceeds %u", c.kprime, kMaxRerank);
    return TSC_ERR_BAD_ARG;
  }
  TailParams p{};
  fill_tail(ix, c, m, gemm_keys, &p);
  const uint32_t sort_cap = tail_sort_cap(m, c.kprime, p.list_len, false);
  switch (ix->desc.metric) {
    case TSC_METRIC_L2: return run_tail_dtype<kL2>(ix, p, c.nq, sort_cap, c.st);
    case TSC_METRIC_INNER_PRODUCT: return run_tail_dtype<kIP>(ix, p, c.nq, sort_cap, c.st);
    default: return run_tail_dtype<kCos>(ix, p, c.nq, sort_cap, c.st);
  }
}

// ---- shard merge: [n_parts][nq][k] -> [nq][k] (tsc_merge_shards, the NCCL all-gather path)
struct MergeParams {
  const int64_t *part_ids;   // part p, query q, rank j at [p * part_stride + q * k + j]
  const double *part_dist;
  uint64_t part_stride;      // elements between consecutive parts
  uint32_t n_parts, nq, k;
  int64_t *out_ids;
  double *out_dist;
  uint32_t *out_counts;
  uint32_t sort_cap;  // pow2 >= n_parts * k
};

__global__ void __launch_bounds__(256) merge_shards_kernel(const MergeParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Pair128 *buf = reinterpret_cast<Pair128 *>(smem);
  __shared__ uint32_t s_count;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t total = p.n_parts * p.k;
  if (tid == 0) s_count = 0;
  for (uint32_t i = tid; i < p.sort_cap; i += blockDim.x) {
    Pair128 e{~0ull, ~0ull};
    if (i < total) {
      uint32_t part = i / p.k, j = i % p.k;
      size_t o = (size_t)part * p.part_stride + (size_t)q * p.k + j;
      int64_t id = p.part_ids[o];
      if (id >= 0) {
        e.hi = ordered_key64(p.part_dist[o]);
        // node ids are non-negative; NaN distances (hi == ~0) still order by id
        e.lo = (uint64_t)id;
      }
    }
    buf[i] = e;
  }
  __syncthreads();
  bitonic_sort_pairs(buf, p.sort_cap);
  for (uint32_t j = tid; j < p.k; j += blockDim.x) {
    Pair128 e = buf[j];
    bool ok = e.lo != ~0ull;
    p.out_ids[(size_t)q * p.k + j] = ok ? (int64_t)e.lo : -1;
    p.out_dist[(size_t)q * p.k + j] =
        ok ? key64_to_double(e.hi) : __longlong_as_double(0x7FF8000000000000ll);
    if (ok) atomicAdd(&s_count, 1u);
  }
  __syncthreads();
  if (tid == 0) p.out_counts[q] = s_count;
}

int32_t launch_merge(Index *ix, const int64_t *d_part_ids, const double *d_part_dist,
                     uint64_t part_stride, uint32_t n_parts, uint32_t nq, uint32_t k,
                     int64_t *d_ids, double *d_dist, uint32_t *d_counts, cudaStream_t st) {
  MergeParams p{};
  p.part_ids = d_part_ids;
  p.part_dist = d_part_dist;
  p.part_stride = part_stride;
  p.n_parts = n_parts;
  p.nq = nq;
  p.k = k;
  p.out_ids = d_ids;
  p.out_dist = d_dist;
  p.out_counts = d_counts;
  p.sort_cap = next_pow2(n_parts * k);
  if (p.sort_cap < 2) p.sort_cap = 2;
  size_t smem = (size_t)p.sort_cap * sizeof(Pair128);
  if (smem > 48 * 1024) {
    set_error("merge: n_parts*k=%u too large", n_parts * k);
    return TSC_ERR_BAD_ARG;
  }
  merge_shards_kernel<<<nq, 256, smem, st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

// Batched paths: one CTA per query pushes, (consumers) waits + merges; the last CTA of a
// consumer to finish acknowledges the epoch. src_*: this shard's exact top-k [nq][k].
struct ExchangeKernelParams {
  XchgParams x;
  const int64_t *src_ids;
  const double *src_dist;
  uint32_t nq, k, sort_cap;
  int64_t *out_ids;
  double *out_dist;
  uint32_t *out_counts;
  uint32_t *done_counter;   // zero between launches
};

__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeKernelParams p) {
  extern __shared__ __align__(16) uint8_t xsmem[];
  __shared__ uint32_t s_last;
  const uint32_t q = blockIdx.x;
  xchg_push(p.x, q, p.k, p.src_ids + (size_t)q * p.k, p.src_dist + (size_t)q * p.k);
  if (!xchg_is_consumer(p.x, p.x.rank)) return;
  xchg_wait_merge(p.x, q, p.k, reinterpret_cast<Pair128 *>(xsmem), p.sort_cap,
                  p.out_ids + (size_t)q * p.k, p.out_dist + (size_t)q * p.k, p.out_counts + q);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(p.done_counter, 1u) == p.nq - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *p.done_counter = 0;
    xchg_ack(p.x);
  }
}

// K9 for the batched paths: push this shard's top-k to the consumers, wait, merge.
int32_t launch_exchange(Index *ix, const SearchCtx &c) {
  if (!ix->p2p_ready) {
    set_error("exchange: the peer-memory exchange has not been set up");
    return TSC_ERR_NCCL;
  }
  ExchangeKernelParams p{};
  fill_xchg(ix, &p.x);
  p.src_ids = c.loc_ids;
  p.src_dist = c.loc_dist;
  p.nq = c.nq;
  p.k = c.k;
  p.sort_cap = next_pow2((uint32_t)ix->n_ranks * c.k);
  if (p.sort_cap < 2) p.sort_cap = 2;
  p.out_ids = c.x_ids;
  p.out_dist = c.x_dist;
  p.out_counts = c.x_counts;
  p.done_counter = ix->d_done + 1;
  const size_t smem = (size_t)p.sort_cap * sizeof(Pair128);
  if (smem > 48 * 1024) {
    set_error("exchange: n_ranks*k=%u too large", ix->n_ranks * c.k);
    return TSC_ERR_BAD_ARG;
  }
  exchange_merge_kernel<<<c.nq, 256, smem, c.st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

}  // namespace tsc
