// tsc_select.cu — host launchers for K5 (tsc_select.cuh).
#include <stdlib.h>

#include "tsc_exchange.cuh"
#include "tsc_index.h"
#include "tsc_select.cuh"

namespace tsc {

template <int DTYPE>
static int32_t run_select(Index *ix, const SelectParams &p, uint32_t nq, cudaStream_t st) {
  static bool attr_done[64] = {false};
  auto kern = select_rerank_kernel<DTYPE>;
  size_t smem = select_smem_bytes(p.sort_cap, p.qld);
  if (!attr_done[ix->device & 63]) {
    TSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ix->smem_optin - 16 * 1024));
    attr_done[ix->device & 63] = true;
  }
  if (smem > ix->smem_optin - 16 * 1024) {
    set_error("select: dims too large for the re-rank staging (%zu B)", smem);
    return TSC_ERR_BAD_DIMS;
  }
  kern<<<nq, kSelectThreads, smem, st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

int32_t launch_select(Index *ix, const float *d_q, uint32_t nq, uint32_t k, uint32_t kprime,
                      const uint64_t *d_cand, uint32_t m, double threshold, int64_t *d_ids,
                      double *d_dist, uint32_t *d_counts, cudaStream_t st) {
  if (kprime > kMaxRerank) {
    set_error("select: k'=%u exceeds %u", kprime, kMaxRerank);
    return TSC_ERR_BAD_ARG;
  }
  SelectParams p{};
  p.cand = d_cand;
  p.m = m;
  p.kprime = kprime;
  p.k = k;
  p.rows = ix->d_rows;
  p.row_bytes = ix->row_bytes;
  p.dims = ix->desc.dims;
  p.queries = d_q;
  p.qld = ix->qld;
  p.metric = ix->desc.metric;
  p.threshold = threshold;
  p.first_node_id = (int64_t)ix->desc.first_node_id;
  p.out_ids = d_ids;
  p.out_dist = d_dist;
  p.out_counts = d_counts;
  p.sort_cap = m <= kSelectSortMax ? next_pow2(m) : next_pow2(kprime);
  if (p.sort_cap < 2) p.sort_cap = 2;
  switch (ix->desc.dev_dtype) {
    case TSC_DEV_F32: return run_select<kF32>(ix, p, nq, st);
    case TSC_DEV_BF16: return run_select<kBF16>(ix, p, nq, st);
    default: return run_select<kF16>(ix, p, nq, st);
  }
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

// K9 (opt-in): push this shard's top-k to every rank over peer memory, wait, merge.
int32_t launch_exchange(Index *ix, const int64_t *d_src_ids, const double *d_src_dist, uint32_t nq,
                        uint32_t k, int64_t *d_ids, double *d_dist, uint32_t *d_counts,
                        cudaStream_t st) {
  if (!ix->p2p_ready) {
    set_error("exchange: tsc_comm_p2p_import has not been called");
    return TSC_ERR_NCCL;
  }
  if (ix->h_xstatus && *reinterpret_cast<volatile uint32_t *>(ix->h_xstatus) != 0) {
    set_error("exchange: an earlier peer-memory exchange timed out (a rank fell behind or died)");
    return TSC_ERR_NCCL;
  }
  ExchangeParams p{};
  p.src_ids = d_src_ids;
  p.src_dist = d_src_dist;
  for (int r = 0; r < ix->n_ranks; r++) p.peer_base[r] = ix->x_peer[r];
  p.n_ranks = (uint32_t)ix->n_ranks;
  p.rank = (uint32_t)ix->rank;
  p.nq = nq;
  p.k = k;
  p.k_stride = ix->k_max;
  p.nq_max = ix->nq_max;
  p.slot_bytes = ix->xslot_bytes;
  p.flag_off = exchange_flag_off(p.n_ranks, p.slot_bytes);
  p.epoch = ++ix->xepoch;
  p.sort_cap = next_pow2(p.n_ranks * k);
  if (p.sort_cap < 2) p.sort_cap = 2;
  p.out_ids = d_ids;
  p.out_dist = d_dist;
  p.out_counts = d_counts;
  p.status = ix->d_xstatus;
  // ~2 s at 2 GHz unless overridden (TSC_P2P_TIMEOUT_MS)
  long long ms = 2000;
  if (const char *ev = getenv("TSC_P2P_TIMEOUT_MS")) ms = atoll(ev) > 0 ? atoll(ev) : ms;
  p.timeout_cycles = ms * 2000000ll;
  const size_t smem = (size_t)p.sort_cap * sizeof(Pair128);
  if (smem > 48 * 1024) {
    set_error("exchange: n_ranks*k=%u too large", p.n_ranks * k);
    return TSC_ERR_BAD_ARG;
  }
  exchange_merge_kernel<<<nq, 256, smem, st>>>(p);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  return TSC_OK;
}

}  // namespace tsc
