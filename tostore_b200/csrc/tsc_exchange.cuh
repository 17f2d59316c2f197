// tsc_exchange.cuh — K9 (opt-in, experimental): the shard exchange as ONE kernel over
// NVLink peer memory instead of ncclAllGather + merge_shards_kernel.
//
// the only exchange of the path is k (distance, nodeId) pairs per query per
// shard (k x 16 B). It is latency-bound, not bandwidth-bound, so the collective is replaced
// by one-sided pushes: CTA q of every rank stores its shard's k pairs for query q straight
// into every peer's receive buffer (16-byte stores through NVLink / NVSwitch), publishes a
// release flag carrying the search epoch, waits for the n_ranks flags of query q in its OWN
// buffer, and merges the n_ranks x k pairs with the reference's final ordering rule
// (core/vector_index_manager.dart:587: ascending distance in double.compareTo order, ties
// by node id) — the same code as merge_shards_kernel.
//
// Buffers are double-buffered by epoch parity. A rank can be at most one epoch ahead of a
// peer (it cannot finish epoch e+1 without that peer's e+1 flags, which the peer only sends
// after it finished reading epoch e), so a slot is never overwritten while it is read.
// One CTA per query waits for the peers' CTA of the same query, so all CTAs of a launch must
// be co-resident (nq_max <= 512 is enforced at export; 148 SMs x >= 4 CTAs of 256 threads).
// The spin has a clock64 timeout: on expiry a status word in mapped host memory is set and
// the query returns empty instead of hanging the GPU.
//
// Written after round 1's GPU budget was spent: NOT yet run on hardware; enabled only by
// tsc_comm_p2p_export/import (TSC_EXCHANGE=p2p in bench.py / tests gated by TSC_TEST_P2P=1).
#pragma once

#include "tsc_select.cuh"

namespace tsc {

constexpr int kMaxRanks = 8;

struct ExchangeParams {
  const int64_t *src_ids;     // this shard's exact top-k, [nq][k]
  const double *src_dist;
  uint8_t *peer_base[kMaxRanks];   // receive buffer of every rank (own entry = local memory)
  uint32_t n_ranks, rank, nq, k;
  uint32_t k_stride;          // pairs reserved per query in a slot (k_max)
  uint32_t nq_max;            // queries reserved per slot / flags per (parity, source)
  uint64_t slot_bytes;        // nq_max * k_stride * 16
  uint64_t flag_off;          // byte offset of the flag area in a receive buffer
  uint32_t epoch;             // >= 1, increases by one per sharded search
  uint32_t sort_cap;          // pow2 >= n_ranks * k
  int64_t *out_ids;
  double *out_dist;
  uint32_t *out_counts;
  uint32_t *status;           // mapped host word: set to 1 on timeout
  long long timeout_cycles;
};

// receive buffer: [2 parities][n_ranks sources][slot_bytes] data, then
//                 [2][n_ranks][nq_max] uint32 flags (128-byte aligned)
__host__ __device__ inline uint64_t exchange_flag_off(uint32_t n_ranks, uint64_t slot_bytes) {
  return (2ull * n_ranks * slot_bytes + 127ull) & ~127ull;
}
__host__ __device__ inline uint64_t exchange_buf_bytes(uint32_t n_ranks, uint64_t slot_bytes,
                                                       uint32_t nq_max) {
  return exchange_flag_off(n_ranks, slot_bytes) + 2ull * n_ranks * nq_max * 4ull;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Pair128 *buf = reinterpret_cast<Pair128 *>(smem);
  __shared__ uint32_t s_count, s_timeout;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  const uint32_t par = p.epoch & 1u;
  if (tid == 0) {
    s_count = 0;
    s_timeout = 0;
  }

  // ---- 1. push this shard's k pairs of query q into every rank's buffer ----------------
  const uint64_t my_slot = ((uint64_t)par * p.n_ranks + p.rank) * p.slot_bytes +
                           (uint64_t)q * p.k_stride * 16ull;
  for (uint32_t i = tid; i < p.n_ranks * p.k; i += blockDim.x) {
    const uint32_t r = i / p.k, j = i % p.k;
    ulonglong2 v;
    v.x = (unsigned long long)p.src_ids[(size_t)q * p.k + j];
    v.y = (unsigned long long)__double_as_longlong(p.src_dist[(size_t)q * p.k + j]);
    // one 16-byte store (weak; made visible by the system fence + release flag below)
    *reinterpret_cast<ulonglong2 *>(p.peer_base[r] + my_slot + (uint64_t)j * 16ull) = v;
  }
  __threadfence_system();
  __syncthreads();

  // ---- 2. publish: flag (parity, source = me, query q) := epoch in every rank's buffer ----
  if (tid < p.n_ranks) {
    uint32_t *flag = reinterpret_cast<uint32_t *>(p.peer_base[tid] + p.flag_off) +
                     ((size_t)par * p.n_ranks + p.rank) * p.nq_max + q;
    st_release_sys(flag, p.epoch);
  }

  // ---- 3. wait for every source's flag of query q in MY buffer ---------------------------
  if (tid < p.n_ranks) {
    const uint32_t *flag = reinterpret_cast<const uint32_t *>(p.peer_base[p.rank] + p.flag_off) +
                           ((size_t)par * p.n_ranks + tid) * p.nq_max + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != p.epoch) {
      if (clock64() - t0 > p.timeout_cycles) {
        s_timeout = 1;
        *reinterpret_cast<volatile uint32_t *>(p.status) = 1u;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();

  // ---- 4. merge n_ranks x k pairs (same rule as merge_shards_kernel) ---------------------
  const uint32_t total = p.n_ranks * p.k;
  const bool dead = s_timeout != 0;
  for (uint32_t i = tid; i < p.sort_cap; i += blockDim.x) {
    Pair128 e{~0ull, ~0ull};
    if (i < total && !dead) {
      const uint32_t r = i / p.k, j = i % p.k;
      const uint8_t *src = p.peer_base[p.rank] + ((uint64_t)par * p.n_ranks + r) * p.slot_bytes +
                           (uint64_t)q * p.k_stride * 16ull + (uint64_t)j * 16ull;
      // written by a peer over NVLink: read around L1
      const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(src));
      const int64_t id = (int64_t)v.x;
      if (id >= 0) {
        e.hi = ordered_key64(__longlong_as_double((long long)v.y));
        e.lo = (uint64_t)id;
      }
    }
    buf[i] = e;
  }
  __syncthreads();
  bitonic_sort_pairs(buf, p.sort_cap);
  for (uint32_t j = tid; j < p.k; j += blockDim.x) {
    Pair128 e = buf[j];
    const bool ok = e.lo != ~0ull;
    int64_t id = -1;
    double d = __longlong_as_double(0x7FF8000000000000ll);
    if (ok) {
      id = (int64_t)e.lo;
      if (e.hi != ~0ull) {
        uint64_t b = (e.hi & 0x8000000000000000ull) ? (e.hi & 0x7FFFFFFFFFFFFFFFull) : ~e.hi;
        d = __longlong_as_double((long long)b);
      }
      atomicAdd(&s_count, 1u);
    }
    p.out_ids[(size_t)q * p.k + j] = id;
    p.out_dist[(size_t)q * p.k + j] = d;
  }
  __syncthreads();
  if (tid == 0) p.out_counts[q] = s_count;
}

}  // namespace tsc
