// tsc_scan.cuh — K1 / K6: HBM-bound exact scan with in-kernel top-K' selection and a fused
// tail (selection, exact re-rank, certificate, shard exchange) in the last CTA.
//
// Replaces the candidate-generation half of NghGraphEngine.search
// (core/ngh_graph_engine.dart:98-113: ADC beam search) with an exhaustive pass
// over the row-major embedding block. The fp32 ranking key computed here plays
// the role the PQ/ADC distance plays in the reference: it only has to put the
// true top-k inside the K' = max(2k, 20) candidates (ngh_graph_engine.dart:115)
// that the tail (tsc_tail.cuh) then re-ranks with the reference's exact fp64
// arithmetic — and the tail PROVES that it did (certificate), re-running the query
// in range mode when it cannot.
//
// Data movement: every warp owns a private ring of S stages in shared memory and
// keeps it full with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA
// unit) of R consecutive rows each, completion on a per-stage mbarrier. No
// CTA-wide barrier exists in the main loop. Lanes read 16-byte chunks of the
// staged rows (conflict-free LDS.128), accumulate in fp32, butterfly-reduce with
// warp shuffles, and insert into a per-warp sorted candidate list in shared memory.
//
// mode 0 (first pass): per-warp top-K' lists -> block merge -> one list per CTA.
// mode 1 (range pass): every row with key <= T[q] is appended to a global buffer; the
//   launch exits at once when no query of its group needs it (the common case).
// Tail: the last CTA to finish (atomic ticket) runs tsc_tail.cuh for the launch's queries
// and, for a sharded index, pushes the shard's exact top-k to the consumers over NVLink
// and merges (tsc_exchange.cuh) — scan, select, re-rank, exchange and merge are one kernel.
#pragma once

#include "tsc_exchange.cuh"

namespace tsc {

struct ScanParams {
  const uint8_t *rows;       // [n_rows, row_bytes] device storage dtype, 16B aligned
  uint64_t n_rows;
  uint32_t row_bytes;        // row stride in bytes (multiple of 16)
  uint32_t chunks_per_row;   // row_bytes / 16
  const float *queries;      // [nq_total, qld] fp32, zero padded to qld
  uint32_t qld;              // padded dims = chunks_per_row * elems_per_chunk
  uint32_t q_base;           // mode 0: first query of this launch; mode 1: first retry_list entry
  uint32_t nq;               // mode 0: valid queries in this launch (<= QB)
  const uint32_t *live_mask; // optional: bit r = row r may be returned; NULL = all live
  uint32_t kprime;           // candidates kept per list
  uint32_t stages;           // S
  uint32_t stage_bytes;      // R * row_bytes
  uint64_t *cand;            // [nq_total][gridDim.x][kprime] (ordered key << 32 | shard row)
  uint32_t sort_cap;         // pow2 >= warps * kprime (block-level merge buffer)
  int mode;                  // 0 = top-K' lists, 1 = range collect
  int fused_tail;            // the last CTA runs the tail (always in mode 1)
  int xchg_in_tail;          // ... and the shard exchange
  int last_retry;            // mode 1: the last range launch of the search resets retry_n
  int no_range;              // mode 0: no range launch follows this first pass (pipelined searches)
  uint32_t nq_total;         // queries of the whole search (exchange loop)
  uint32_t tail_sort_cap;    // Pair128 slots of the tail's sort buffer
  uint32_t smem_bytes;       // dynamic shared memory of this launch (the tail stages rows in it)
  uint32_t *done_counter;    // zero between launches
  uint32_t *work_counter;    // dynamic stage blocks handed out so far; zero between launches
  uint64_t static_stages;    // stages [0, static_stages) are dealt round-robin, the rest on demand
  TailParams tail;
  XchgParams xchg;
  int64_t *x_out_ids;        // exchange output (global top-k), [nq_total, k]
  double *x_out_dist;
  uint32_t *x_out_counts;
};

// Insert (key,id) into a sorted (ascending) list of kp entries held in shared
// memory by one warp; the last entry falls off. All lanes pass identical
// arguments. Rows reach a warp in increasing id order, so an equal key is placed
// after the existing ones and (key,id) stays lexicographically sorted.
// Returns the new threshold (largest kept key).
__device__ __forceinline__ uint32_t list_insert(uint32_t *keys, uint32_t *ids, int kp,
                                                uint32_t key, uint32_t id, int lane) {
  int pos = 0;
  for (int base = 0; base < kp; base += 32) {
    int j = base + lane;
    bool le = (j < kp) && (keys[j] <= key);
    unsigned m = __ballot_sync(0xFFFFFFFFu, le);
    pos += __popc(m);
    if (m != 0xFFFFFFFFu) break;
  }
  for (int base = ((kp - 1) >> 5) << 5; base >= 0 && base + 31 >= pos; base -= 32) {
    int j = base + lane;
    bool mv = (j > pos) && (j < kp);
    uint32_t k1 = 0, i1 = 0;
    if (mv) {
      k1 = keys[j - 1];
      i1 = ids[j - 1];
    }
    __syncwarp();
    if (mv) {
      keys[j] = k1;
      ids[j] = i1;
    } else if (j == pos) {
      keys[j] = key;
      ids[j] = id;
    }
  }
  __syncwarp();
  return keys[kp - 1];
}

// A row passed the threshold test: keep it. mode 0: sorted insertion into the warp's list
// (returns the new threshold); mode 1: append to the query slot's global range buffer
// (threshold unchanged). Warp-uniform.
__device__ __forceinline__ uint32_t scan_keep(const ScanParams &p, uint32_t *lkeys, uint32_t *lids,
                                              uint32_t slot, uint32_t thr, uint32_t uk,
                                              uint32_t row, int lane) {
  if (p.mode == 0) return list_insert(lkeys, lids, (int)p.kprime, uk, row, lane);
  if (lane == 0) {
    const uint32_t s = atomicAdd(p.tail.range_count + slot, 1u);
    if (s < kRangeCap) p.tail.range_buf[(size_t)slot * kRangeCap + s] = ((uint64_t)uk << 32) | row;
  }
  return thr;
}

template <int METRIC>
__device__ __forceinline__ uint32_t scan_key(float acc, float bb) {
  float key;
  if (METRIC == kL2) {
    key = acc;
  } else if (METRIC == kIP) {
    key = -acc;
  } else {
    key = (bb > 0.0f) ? -acc * rsqrtf(bb) : 0.0f;
  }
  key += 0.0f;  // -0.0 -> +0.0 so exact ties order by id only
  return ordered_key(key);
}

// Shared-memory footprint (must match the carve-up in the kernel).
// Query rows in shared memory. fp32 columns: as they are. 16-bit columns: a lane needs the 8
// query elements of its chunk = two LDS.128 at a 32-byte lane stride, which is a 2-way bank
// conflict (308 M conflicts in one C4 launch, ncu); so every block of 256 elements (one chunk
// per lane) is stored as its 32 first halves followed by its 32 second halves and both loads
// are conflict-free. Stride = qld rounded up to whole blocks.
__host__ __device__ inline uint32_t scan_q_stride(uint32_t qld) { return (qld + 255u) & ~255u; }
__host__ __device__ inline size_t scan_smem_query_bytes(int qb, uint32_t qld) {
  return ((size_t)qb * scan_q_stride(qld) * 4 + 127) & ~(size_t)127;
}
template <int E>
__device__ __forceinline__ uint32_t scan_q_slot(uint32_t e) {   // element index -> slot in its row
  if (E == 4) return e;
  const uint32_t r = e & 255u;
  return (e & ~255u) + ((r >> 2) & 1u) * 128u + (r >> 3) * 4u + (r & 3u);
}
template <int E>
__device__ __forceinline__ const float4 *scan_q_chunk(const float *row, uint32_t c, int h) {
  if (E == 4) return reinterpret_cast<const float4 *>(row + (size_t)c * 4);
  return reinterpret_cast<const float4 *>(row + ((size_t)(c >> 5) << 8) + (size_t)h * 128 + (size_t)(c & 31u) * 4);
}
__host__ __device__ inline size_t scan_smem_sort_bytes(uint32_t sort_cap) {
  return ((size_t)sort_cap * 8 + 127) & ~(size_t)127;
}
__host__ __device__ inline size_t scan_smem_warp_bytes(int qb, uint32_t kprime, uint32_t stages,
                                                      uint32_t stage_bytes) {
  size_t lists = ((size_t)qb * kprime * 8 + 15) & ~(size_t)15;
  size_t bars = ((size_t)stages * 8 + 15) & ~(size_t)15;
  size_t ids = ((size_t)stages * 4 + 15) & ~(size_t)15;   // which stage each ring slot holds
  size_t ring = (size_t)stages * stage_bytes;
  return (((lists + bars + ids + 127) & ~(size_t)127) + ring + 127) & ~(size_t)127;
}

// ---- block-level merge: W sorted lists -> one list of kp per query -------------
// (every bulk copy this CTA issued has been waited on, so no async write is outstanding)
__device__ __forceinline__ void scan_block_merge(const ScanParams &p, const uint32_t *qi,
                                                 uint32_t nq, uint64_t *sortbuf,
                                                 const uint32_t *lkeys, const uint32_t *lids,
                                                 uint32_t kp, int warp, int warps, int lane) {
  for (uint32_t q = 0; q < nq; q++) {
    __syncthreads();
    for (uint32_t i = lane; i < kp; i += 32)
      sortbuf[(size_t)warp * kp + i] =
          ((uint64_t)lkeys[(size_t)q * kp + i] << 32) | lids[(size_t)q * kp + i];
    for (uint32_t i = (uint32_t)warps * kp + threadIdx.x; i < p.sort_cap; i += blockDim.x)
      sortbuf[i] = ~0ull;
    __syncthreads();
    // Merge by ranking: every list is sorted, so the position of an element in the merged
    // order is its own index plus, for every other list, the number of entries below it (one
    // binary search each). Ties exist only between empty slots; they are broken by list
    // number (<= for lists before mine, < for lists after), which makes the ranks a
    // permutation. One pass, no barrier per stage (the bitonic network this replaces spent
    // ~6 us of every launch in its 36 barriers).
    uint64_t *out = p.cand + ((size_t)qi[q] * gridDim.x + blockIdx.x) * kp;
    for (uint32_t idx = threadIdx.x; idx < (uint32_t)warps * kp; idx += blockDim.x) {
      const uint32_t w = idx / kp, i = idx - w * kp;
      const uint64_t v = sortbuf[idx];
      uint32_t rank = i;
      for (uint32_t o = 0; o < (uint32_t)warps; o++) {
        if (o == w) continue;
        const uint64_t *l = sortbuf + (size_t)o * kp;
        uint32_t lo = 0, hi = kp;   // first index whose entry is > v (o < w) or >= v (o > w)
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          const uint64_t x = l[mid];
          const bool below = o < w ? x <= v : x < v;
          if (below) lo = mid + 1;
          else hi = mid;
        }
        rank += lo;
      }
      if (rank < kp) out[rank] = v;
    }
  }
}

// Which queries does this launch work on? mode 0: q_base .. q_base + nq. mode 1: the
// entries [q_base, q_base + QB) of the retry list; returns 0 when there are none (the
// launch exits at once). Uniform over the grid.
template <int QB>
__device__ __forceinline__ uint32_t scan_queries(const ScanParams &p, uint32_t (&qi)[QB]) {
  uint32_t nq = p.nq;
  if (p.mode == 1) {
    // launched with programmatic stream serialization: the first pass (whose tail writes
    // retry_n / range_thr) has to be complete and visible before anything is read. Whatever
    // follows this launch in the same way (the next search's first pass, pipelining) may be
    // scheduled right away: it waits for THIS grid before it publishes anything.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t n_retry = __ldcg(p.tail.retry_n);
    if (n_retry <= p.q_base) {
      if (p.last_retry && blockIdx.x == 0 && threadIdx.x == 0 && n_retry != 0) *p.tail.retry_n = 0;
      return 0;
    }
    nq = n_retry - p.q_base < (uint32_t)QB ? n_retry - p.q_base : (uint32_t)QB;
  }
#pragma unroll
  for (int q = 0; q < QB; q++) {
    const uint32_t j = (uint32_t)q < nq ? (uint32_t)q : 0u;
    qi[q] = p.mode == 1 ? __ldcg(p.tail.retry_list + p.q_base + j) : p.q_base + j;
  }
  return nq;
}

// Everything after the main loop: publish the CTA's candidates, then the LAST CTA of the
// grid runs the tail for the launch's queries and the exchange of the search.
template <int METRIC, int DTYPE, int QB>
__device__ __forceinline__ void scan_finish(const ScanParams &p, const uint32_t (&qi)[QB],
                                            uint32_t nq, uint8_t *smem, uint64_t *sortbuf,
                                            const uint32_t *lkeys, const uint32_t *lids, int warp,
                                            int warps, int lane) {
  __shared__ uint32_t s_ticket;
  TSC_TRACE(p.tail.diag, kTraceCta + blockIdx.x);
  // A pipelined first pass ran its main loop beside the previous search's tail / range launch;
  // from here on it touches what they use (candidate lists, tickets, flags, results): wait for
  // them to complete (returns at once for a launch without a programmatic dependency).
  if (p.mode == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.mode == 0) scan_block_merge(p, qi, nq, sortbuf, lkeys, lids, p.kprime, warp, warps, lane);
  TSC_TRACE(p.tail.diag, kTraceCta + gridDim.x + blockIdx.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(p.done_counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (threadIdx.x == 0) {   // stream order: the next launch sees zero
    *p.done_counter = 0;
    *p.work_counter = 0;
  }
  if (!p.fused_tail) return;
  __threadfence();
  TSC_TRACE(p.tail.diag, 1);
  // every other CTA has exited: let the range launch that follows be scheduled now
  if (p.mode == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (uint32_t q = 0; q < nq; q++)
    tail_query<METRIC, DTYPE>(p.tail, qi[q], p.mode, q, smem, p.smem_bytes, p.tail_sort_cap);
  // the last range launch of a search leaves the retry list empty for the next search
  // (queries beyond the in-stream range launches keep kFlagRetry: the host API re-runs them)
  if (p.mode == 1 && p.last_retry && threadIdx.x == 0) *p.tail.retry_n = 0;
  if (!p.xchg_in_tail) return;
  // a first pass with uncertified queries leaves the exchange to the range launch
  __syncthreads();
  if (p.mode == 0 && __ldcg(p.tail.retry_n) != 0) return;
  for (uint32_t q = 0; q < p.nq_total; q++) {
    __syncthreads();
    xchg_push(p.xchg, q, p.tail.k, p.tail.out_ids + (size_t)q * p.tail.k,
              p.tail.out_dist + (size_t)q * p.tail.k);
  }
  if (!xchg_is_consumer(p.xchg, p.xchg.rank)) return;
  uint32_t xcap = 2;
  while (xcap < p.xchg.n_ranks * p.tail.k) xcap <<= 1;
  for (uint32_t q = 0; q < p.nq_total; q++) {
    __syncthreads();
    xchg_wait_merge(p.xchg, q, p.tail.k, reinterpret_cast<Pair128 *>(smem), xcap,
                    p.x_out_ids + (size_t)q * p.tail.k, p.x_out_dist + (size_t)q * p.tail.k,
                    p.x_out_counts + q);
  }
  __syncthreads();
  xchg_ack(p.xchg);
  TSC_TRACE(p.tail.diag, 7);
}

template <int METRIC, int DTYPE, int QB, int R>
__global__ void __launch_bounds__(512, 1) scan_topk_kernel(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int E = Chunk<DTYPE>::kElems;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const uint32_t kp = p.kprime;
  const uint32_t S = p.stages;

  uint32_t qi[QB];
  const uint32_t nq = scan_queries<QB>(p, qi);
  if (nq == 0) return;
#ifdef TSC_DIAG
  if (p.tail.diag && threadIdx.x == 0) atomicMin(p.tail.diag, trace_now());
#endif

  // ---- carve shared memory -------------------------------------------------
  float *qs = reinterpret_cast<float *>(smem);
  uint64_t *sortbuf = reinterpret_cast<uint64_t *>(smem + scan_smem_query_bytes(QB, p.qld));
  uint8_t *wbase = smem + scan_smem_query_bytes(QB, p.qld) + scan_smem_sort_bytes(p.sort_cap) +
                   (size_t)warp * scan_smem_warp_bytes(QB, kp, S, p.stage_bytes);
  uint32_t *lkeys = reinterpret_cast<uint32_t *>(wbase);            // [QB][kp]
  uint32_t *lids = lkeys + (size_t)QB * kp;                         // [QB][kp]
  size_t off = ((size_t)QB * kp * 8 + 15) & ~(size_t)15;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + off);
  off += ((size_t)S * 8 + 15) & ~(size_t)15;
  uint32_t *slot_stage = reinterpret_cast<uint32_t *>(wbase + off);   // [S]
  off += ((size_t)S * 4 + 15) & ~(size_t)15;
  off = (off + 127) & ~(size_t)127;
  uint8_t *ring = wbase + off;

  // ---- queries -> smem (zero rows for q >= nq), lists -> empty --------------
  const uint32_t qstride = scan_q_stride(p.qld);
  for (uint32_t i = threadIdx.x; i < (uint32_t)QB * p.qld; i += blockDim.x) {
    const uint32_t q = i / p.qld, e = i - q * p.qld;
    qs[(size_t)q * qstride + scan_q_slot<E>(e)] =
        (q < nq) ? p.queries[(size_t)qi[q] * p.qld + e] : 0.0f;
  }
  for (uint32_t i = lane; i < (uint32_t)QB * kp; i += 32) {
    lkeys[i] = kEmptyKey;
    lids[i] = kInvalidRow;
  }
  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t policy = policy_evict_first();
  const uint64_t total_stages = (p.n_rows + R - 1) / R;
  const uint64_t gw = (uint64_t)blockIdx.x * warps + warp;
  const uint64_t GW = (uint64_t)gridDim.x * warps;

  // Work distribution. SMs do not get equal shares of HBM bandwidth (the last CTA of a
  // round-robin deal finished ~8 % after the median one: tools/scan_trace.py), so only the
  // first p.static_stages stages are dealt round-robin; the rest is handed out on demand in
  // blocks of kDynBlock consecutive stages from an atomic counter. Either way a warp sees its
  // rows in increasing order. Lane 0 runs the generator and records which stage each ring
  // slot holds (kNoStage = nothing more: the slots after it hold nothing either).
  constexpr uint32_t kNoStage = 0xFFFFFFFFu;
  constexpr uint32_t kDynBlock = 4;
  uint64_t next_static = gw;
  uint64_t dyn_next = 0;
  uint32_t dyn_left = 0;
  bool dyn_done = false;
  auto next_stage = [&]() -> uint32_t {
    if (next_static < p.static_stages) {
      const uint64_t st = next_static;
      next_static += GW;
      return (uint32_t)st;
    }
    if (dyn_left == 0) {
      if (dyn_done) return kNoStage;
      const uint64_t base = p.static_stages + (uint64_t)atomicAdd(p.work_counter, 1u) * kDynBlock;
      if (base >= total_stages) {
        dyn_done = true;
        return kNoStage;
      }
      dyn_next = base;
      dyn_left = total_stages - base < kDynBlock ? (uint32_t)(total_stages - base) : kDynBlock;
    }
    dyn_left--;
    return (uint32_t)dyn_next++;
  };

  auto issue = [&](uint32_t s, uint64_t st) {
    uint64_t row0 = st * R;
    uint64_t nrow = p.n_rows - row0;
    uint32_t bytes = (nrow >= (uint64_t)R) ? p.stage_bytes : (uint32_t)nrow * p.row_bytes;
    uint32_t bar = smem_u32(&bars[s]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(ring + (size_t)s * p.stage_bytes), p.rows + row0 * p.row_bytes, bytes, bar,
             policy);
  };

  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) {
      const uint32_t st = next_stage();
      slot_stage[s] = st;
      if (st != kNoStage) issue(s, st);
    }
  }
  __syncwarp();

  // mode 0: threshold = the list's largest key; mode 1: fixed T + 1 (key <= T passes)
  uint32_t thr[QB];
#pragma unroll
  for (int q = 0; q < QB; q++)
    thr[q] = (p.mode == 1 && (uint32_t)q < nq) ? __ldcg(p.tail.range_thr + qi[q]) + 1u : kEmptyKey;

  uint32_t s = 0, parity = 0;
  const uint32_t cpr = p.chunks_per_row;
  for (;;) {
    const uint64_t st = slot_stage[s];
    if (st == kNoStage) break;
    mbar_wait(smem_u32(&bars[s]), parity);

    float acc[QB][R];
    float bb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      bb[r] = 0.0f;
#pragma unroll
      for (int q = 0; q < QB; q++) acc[q][r] = 0.0f;
    }
    const uint4 *stage = reinterpret_cast<const uint4 *>(ring + (size_t)s * p.stage_bytes);
#pragma unroll 2
    for (uint32_t c = lane; c < cpr; c += 32) {
      float b[R][E];
#pragma unroll
      for (int r = 0; r < R; r++) {
        uint4 v = stage[(uint32_t)r * cpr + c];
        Chunk<DTYPE>::unpack(v, b[r]);
      }
      if (METRIC == kCos) {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int e = 0; e < E; e++) bb[r] = fmaf(b[r][e], b[r][e], bb[r]);
      }
#pragma unroll
      for (int q = 0; q < QB; q++) {
        float a[E];
#pragma unroll
        for (int h = 0; h < E / 4; h++) {
          const float4 t = *scan_q_chunk<E>(qs + (size_t)q * qstride, c, h);
          a[4 * h + 0] = t.x;
          a[4 * h + 1] = t.y;
          a[4 * h + 2] = t.z;
          a[4 * h + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int e = 0; e < E; e++) {
            if (METRIC == kL2) {
              float d = a[e] - b[r][e];
              acc[q][r] = fmaf(d, d, acc[q][r]);
            } else {
              acc[q][r] = fmaf(a[e], b[r][e], acc[q][r]);
            }
          }
      }
    }
    __syncwarp();  // every lane is done reading this stage (and its slot_stage entry)
    if (lane == 0) {
      const uint32_t nst = next_stage();
      slot_stage[s] = nst;
      if (nst != kNoStage) issue(s, nst);
    }

    // ---- butterfly reduction: every lane ends with the full sums -----------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (METRIC == kCos) bb[r] += __shfl_xor_sync(0xFFFFFFFFu, bb[r], o);
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q][r] += __shfl_xor_sync(0xFFFFFFFFu, acc[q][r], o);
      }
    }

    // ---- candidate insertion (warp-uniform control flow) --------------------
    const uint64_t row0 = st * R;
    uint32_t live = 0xFFFFFFFFu;
    if (p.live_mask != nullptr) {
      // R divides 32 and row0 is a multiple of R: the stage's bits sit in one word
      live = p.live_mask[row0 >> 5] >> (row0 & 31);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint64_t row = row0 + r;
      if (row >= p.n_rows || !((live >> r) & 1u)) continue;
#pragma unroll
      for (int q = 0; q < QB; q++) {
        const uint32_t uk = scan_key<METRIC>(acc[q][r], bb[r]);
        if (uk < thr[q] && (uint32_t)q < nq)
          thr[q] = scan_keep(p, lkeys + (size_t)q * kp, lids + (size_t)q * kp, (uint32_t)q, thr[q],
                             uk, (uint32_t)row, lane);
      }
    }

    if (++s == S) {
      s = 0;
      parity ^= 1u;
    }
  }

  // every copy has been waited on; the tail re-uses the ring (barriers included) as plain memory
  if (lane == 0)
    for (uint32_t i = 0; i < S; i++) mbar_inval(smem_u32(&bars[i]));
  scan_finish<METRIC, DTYPE, QB>(p, qi, nq, smem, sortbuf, lkeys, lids, warp, warps, lane);
}

// K6: scan of a sparsely live column (WHERE prefilter / heavy tombstoning).
// Same arithmetic, candidate lists, modes and tail as scan_topk_kernel, but the unit of
// data movement is ONE LIVE ROW: a warp walks its share of the liveness bitmap (one
// 32-bit word = 32 consecutive rows at a time) and issues a bulk copy only for rows
// whose bit is set, so dead rows cost no HBM bytes (rows are whole 16-byte-aligned
// byte ranges; at d=384 fp32 a row is exactly twelve 128-byte lines). The ring is a
// queue of S one-row stages; rows are consumed in issue (= increasing id) order, G at a
// time: one row per step left the warp latency-bound (ncu, profiles/r02_sparse_*: 168
// dependent instructions per row at 6.7 cycles each, 48 % of HBM) — G rows' LDS / FMA /
// shuffle chains are independent and interleave.
template <int QB>
struct SparseGroup {
  static constexpr int kRows = QB == 1 ? 4 : (QB == 4 ? 2 : 1);
};

template <int METRIC, int DTYPE, int QB>
__global__ void __launch_bounds__(512, 1) scan_topk_sparse_kernel(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int E = Chunk<DTYPE>::kElems;
  constexpr int G = SparseGroup<QB>::kRows;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const uint32_t kp = p.kprime;
  const uint32_t S = p.stages;

  uint32_t qi[QB];
  const uint32_t nq = scan_queries<QB>(p, qi);
  if (nq == 0) return;

  float *qs = reinterpret_cast<float *>(smem);
  uint64_t *sortbuf = reinterpret_cast<uint64_t *>(smem + scan_smem_query_bytes(QB, p.qld));
  uint8_t *wbase = smem + scan_smem_query_bytes(QB, p.qld) + scan_smem_sort_bytes(p.sort_cap) +
                   (size_t)warp * scan_smem_warp_bytes(QB, kp, S, p.stage_bytes);
  uint32_t *lkeys = reinterpret_cast<uint32_t *>(wbase);
  uint32_t *lids = lkeys + (size_t)QB * kp;
  size_t off = ((size_t)QB * kp * 8 + 15) & ~(size_t)15;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + off);
  off += ((size_t)S * 8 + 15) & ~(size_t)15;
  uint32_t *slot_stage = reinterpret_cast<uint32_t *>(wbase + off);   // [S]
  off += ((size_t)S * 4 + 15) & ~(size_t)15;
  off = (off + 127) & ~(size_t)127;
  uint8_t *ring = wbase + off;

  const uint32_t qstride = scan_q_stride(p.qld);
  for (uint32_t i = threadIdx.x; i < (uint32_t)QB * p.qld; i += blockDim.x) {
    const uint32_t q = i / p.qld, e = i - q * p.qld;
    qs[(size_t)q * qstride + scan_q_slot<E>(e)] =
        (q < nq) ? p.queries[(size_t)qi[q] * p.qld + e] : 0.0f;
  }
  for (uint32_t i = lane; i < (uint32_t)QB * kp; i += 32) {
    lkeys[i] = kEmptyKey;
    lids[i] = kInvalidRow;
  }
  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t policy = policy_evict_first();
  const uint64_t n_words = (p.n_rows + 31) / 32;
  const uint64_t gw = (uint64_t)blockIdx.x * warps + warp;
  const uint64_t GW = (uint64_t)gridDim.x * warps;
  const uint32_t cpr = p.chunks_per_row;

  uint32_t thr[QB];
#pragma unroll
  for (int q = 0; q < QB; q++)
    thr[q] = (p.mode == 1 && (uint32_t)q < nq) ? __ldcg(p.tail.range_thr + qi[q]) + 1u : kEmptyKey;

  // issue side: (word, remaining bits) cursor over this warp's bitmap words
  uint64_t word = gw;
  uint32_t bits = 0;
  auto load_word = [&]() {
    bits = 0;
    while (word < n_words) {
      bits = p.live_mask[word];
      uint64_t base_row = word * 32;
      if (base_row + 32 > p.n_rows) bits &= (uint32_t)((1ull << (p.n_rows - base_row)) - 1ull);
      if (bits) break;
      word += GW;
    }
  };
  load_word();
  // ring state: stages [head, head + inflight) mod S are in flight; the barrier of stage s
  // completes its phase number (uses of s so far) & 1
  uint32_t head = 0, tail = 0, inflight = 0, hpar = 0;
  // row ids of the stages in flight (uniform per warp): kept in registers of lane s
  uint32_t my_row = kInvalidRow;

  for (;;) {
    while (inflight < S && bits) {
      const uint32_t b = __ffs(bits) - 1;
      bits &= bits - 1;
      const uint32_t row = (uint32_t)(word * 32 + b);
      if (lane == (int)tail) my_row = row;
      if (lane == 0) {
        const uint32_t bar = smem_u32(&bars[tail]);
        mbar_expect_tx(bar, p.row_bytes);
        bulk_g2s(smem_u32(ring + (size_t)tail * p.stage_bytes), p.rows + (size_t)row * p.row_bytes,
                 p.row_bytes, bar, policy);
      }
      if (++tail == S) tail = 0;
      inflight++;
      if (!bits) {
        word += GW;
        load_word();
      }
    }
    if (inflight == 0) break;

    // ---- consume up to G rows at once (warp-uniform n) -------------------------------
    const uint32_t n = inflight < (uint32_t)G ? inflight : (uint32_t)G;
    uint32_t st_idx[G], rows_g[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      uint32_t h = head + ((uint32_t)g < n ? (uint32_t)g : 0u);
      uint32_t par = hpar;
      if (h >= S) {
        h -= S;
        par ^= 1u;
      }
      st_idx[g] = h;
      if ((uint32_t)g < n) mbar_wait(smem_u32(&bars[h]), par);
      rows_g[g] = __shfl_sync(0xFFFFFFFFu, my_row, (int)h);
    }
    float acc[QB][G];
    float bb[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      bb[g] = 0.0f;
#pragma unroll
      for (int q = 0; q < QB; q++) acc[q][g] = 0.0f;
    }
#pragma unroll 2
    for (uint32_t c = lane; c < cpr; c += 32) {
      float b[G][E];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const uint4 v =
            reinterpret_cast<const uint4 *>(ring + (size_t)st_idx[g] * p.stage_bytes)[c];
        Chunk<DTYPE>::unpack(v, b[g]);
      }
      if (METRIC == kCos) {
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int e = 0; e < E; e++) bb[g] = fmaf(b[g][e], b[g][e], bb[g]);
      }
#pragma unroll
      for (int q = 0; q < QB; q++) {
        float a[E];
#pragma unroll
        for (int h = 0; h < E / 4; h++) {
          const float4 t = *scan_q_chunk<E>(qs + (size_t)q * qstride, c, h);
          a[4 * h + 0] = t.x;
          a[4 * h + 1] = t.y;
          a[4 * h + 2] = t.z;
          a[4 * h + 3] = t.w;
        }
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
          for (int e = 0; e < E; e++) {
            if (METRIC == kL2) {
              float d = a[e] - b[g][e];
              acc[q][g] = fmaf(d, d, acc[q][g]);
            } else {
              acc[q][g] = fmaf(a[e], b[g][e], acc[q][g]);
            }
          }
      }
    }
    __syncwarp();  // the n stages are free again
    head += n;
    if (head >= S) {
      head -= S;
      hpar ^= 1u;
    }
    inflight -= n;

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int g = 0; g < G; g++) {
        if (METRIC == kCos) bb[g] += __shfl_xor_sync(0xFFFFFFFFu, bb[g], o);
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q][g] += __shfl_xor_sync(0xFFFFFFFFu, acc[q][g], o);
      }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      if ((uint32_t)g >= n) continue;
#pragma unroll
      for (int q = 0; q < QB; q++) {
        const uint32_t uk = scan_key<METRIC>(acc[q][g], bb[g]);
        if (uk < thr[q] && (uint32_t)q < nq)
          thr[q] = scan_keep(p, lkeys + (size_t)q * kp, lids + (size_t)q * kp, (uint32_t)q, thr[q],
                             uk, rows_g[g], lane);
      }
    }
  }
  __syncwarp();
  // every copy has been waited on; the tail re-uses the ring (barriers included) as plain memory
  if (lane == 0)
    for (uint32_t i = 0; i < S; i++) mbar_inval(smem_u32(&bars[i]));
  scan_finish<METRIC, DTYPE, QB>(p, qi, nq, smem, sortbuf, lkeys, lids, warp, warps, lane);
}

}  // namespace tsc
