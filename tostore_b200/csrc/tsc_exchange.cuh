// tsc_exchange.cuh — K9: the shard exchange over NVLink peer memory.
//
// the only exchange of the path is k (distance, nodeId) pairs per query per
// shard (k x 16 B). It is latency-bound, not bandwidth-bound, so instead of a collective
// every shard PUSHES its exact top-k straight into the receive buffer of each consumer
// (16-byte stores through NVLink / NVSwitch) and publishes a release flag carrying the
// search epoch; a consumer waits for the n_ranks flags of a query in its OWN memory and
// merges the n_ranks x k pairs with the reference's final ordering rule
// (core/vector_index_manager.dart:587: ascending distance in double.compareTo order, ties
// by node id). Consumers are either every rank (all-gather semantics) or one root rank
// (the rank the host reads from; the others never wait and run ahead).
//
// These are device FUNCTIONS: the scan kernel's last CTA calls them right after the exact
// re-rank (tsc_scan.cuh), so scan + select + re-rank + exchange + merge is ONE kernel;
// exchange_merge_kernel wraps them for the batched paths (one CTA per query).
//
// Receive buffer of a rank (peer-mapped: CUDA IPC between processes, cudaDeviceEnablePeerAccess
// inside one process):
//   data  [depth][n_ranks sources][nq_max][k_stride] x 16 B
//   flags [depth][n_ranks][nq_max] u32      epoch of the data in the slot
//   acks  [n_ranks consumers] u32           last epoch consumer c has finished reading
// A search uses slot epoch % depth. Before a source overwrites a slot it checks that every
// consumer has acknowledged epoch - depth (flow control for root mode, where sources run
// ahead; with all-gather semantics the ranks are in lockstep and the check never waits).
// Spins carry a clock64 timeout: on expiry a status word in mapped host memory is set and
// the query returns empty instead of hanging the GPU.
#pragma once

#include "tsc_tail.cuh"

namespace tsc {

constexpr int kMaxRanks = 8;

struct XchgParams {
  uint8_t *peer[kMaxRanks];   // receive buffer of every rank (own entry = local memory)
  uint32_t n_ranks, rank;
  int32_t root;               // consumer rank, or -1: every rank consumes
  uint32_t k_stride;          // pairs reserved per query in a slot (k_max)
  uint32_t nq_max;            // queries reserved per slot
  uint32_t depth;             // slots
  uint64_t slot_bytes;        // nq_max * k_stride * 16
  uint64_t flag_off, ack_off; // byte offsets of the flag / ack areas
  uint32_t epoch;             // >= 1, increases by one per sharded search
  uint32_t *status;           // mapped host word: set to 1 on timeout
  long long timeout_cycles;
};

__host__ __device__ inline uint64_t xchg_flag_off(uint32_t n_ranks, uint32_t depth,
                                                  uint64_t slot_bytes) {
  return ((uint64_t)depth * n_ranks * slot_bytes + 127ull) & ~127ull;
}
__host__ __device__ inline uint64_t xchg_ack_off(uint32_t n_ranks, uint32_t depth,
                                                 uint64_t slot_bytes, uint32_t nq_max) {
  return (xchg_flag_off(n_ranks, depth, slot_bytes) + (uint64_t)depth * n_ranks * nq_max * 4ull +
          127ull) & ~127ull;
}
__host__ __device__ inline uint64_t xchg_buf_bytes(uint32_t n_ranks, uint32_t depth,
                                                   uint64_t slot_bytes, uint32_t nq_max) {
  return xchg_ack_off(n_ranks, depth, slot_bytes, nq_max) + 128ull;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ bool xchg_is_consumer(const XchgParams &x, uint32_t r) {
  return x.root < 0 || (uint32_t)x.root == r;
}

// spin until *flag (own memory, written by a peer) satisfies (int)(value - want) >= 0
__device__ __forceinline__ bool xchg_spin(const XchgParams &x, const uint32_t *flag, uint32_t want) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - want) < 0) {
    if (clock64() - t0 > x.timeout_cycles) {
      *reinterpret_cast<volatile uint32_t *>(x.status) = 1u;
      return false;
    }
    __nanosleep(32);
  }
  return true;
}

// Push this shard's k pairs of query q to every consumer and publish them. Block-wide;
// src_* are the shard's exact top-k of query q ([k], -1 padded), in global memory.
__device__ __forceinline__ void xchg_push(const XchgParams &x, uint32_t q, uint32_t k,
                                          const int64_t *src_ids, const double *src_dist) {
  const uint32_t tid = threadIdx.x;
  const uint32_t slot = x.epoch % x.depth;
  // flow control: every consumer is done with the search that used this slot last
  if (x.epoch > x.depth && tid < x.n_ranks && xchg_is_consumer(x, tid)) {
    const uint32_t *ack = reinterpret_cast<const uint32_t *>(x.peer[x.rank] + x.ack_off) + tid;
    xchg_spin(x, ack, x.epoch - x.depth);
  }
  __syncthreads();
  const uint64_t my_slot = ((uint64_t)slot * x.n_ranks + x.rank) * x.slot_bytes +
                           (uint64_t)q * x.k_stride * 16ull;
  for (uint32_t i = tid; i < x.n_ranks * k; i += blockDim.x) {
    const uint32_t r = i / k, j = i % k;
    if (!xchg_is_consumer(x, r)) continue;
    ulonglong2 v;
    v.x = (unsigned long long)src_ids[j];
    v.y = (unsigned long long)__double_as_longlong(src_dist[j]);
    // one 16-byte store (weak; made visible by the system fence + release flag below)
    *reinterpret_cast<ulonglong2 *>(x.peer[r] + my_slot + (uint64_t)j * 16ull) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < x.n_ranks && xchg_is_consumer(x, tid)) {
    uint32_t *flag = reinterpret_cast<uint32_t *>(x.peer[tid] + x.flag_off) +
                     ((size_t)slot * x.n_ranks + x.rank) * x.nq_max + q;
    st_release_sys(flag, x.epoch);
  }
}

// Consumer side: wait for every source's pairs of query q, merge, emit [k] results.
// Block-wide; buf = Pair128[sort_cap >= n_ranks * k] in shared memory.
__device__ __forceinline__ void xchg_wait_merge(const XchgParams &x, uint32_t q, uint32_t k,
                                                Pair128 *buf, uint32_t sort_cap, int64_t *out_ids,
                                                double *out_dist, uint32_t *out_count) {
  __shared__ uint32_t s_xcount, s_xdead;
  const uint32_t tid = threadIdx.x;
  const uint32_t slot = x.epoch % x.depth;
  if (tid == 0) {
    s_xcount = 0;
    s_xdead = 0;
  }
  __syncthreads();
  if (tid < x.n_ranks) {
    const uint32_t *flag = reinterpret_cast<const uint32_t *>(x.peer[x.rank] + x.flag_off) +
                           ((size_t)slot * x.n_ranks + tid) * x.nq_max + q;
    if (!xchg_spin(x, flag, x.epoch)) s_xdead = 1;
  }
  __syncthreads();
  const uint32_t total = x.n_ranks * k;
  const bool dead = s_xdead != 0;
  for (uint32_t i = tid; i < sort_cap; i += blockDim.x) {
    Pair128 e{~0ull, ~0ull};
    if (i < total && !dead) {
      const uint32_t r = i / k, j = i % k;
      const uint8_t *src = x.peer[x.rank] + ((uint64_t)slot * x.n_ranks + r) * x.slot_bytes +
                           (uint64_t)q * x.k_stride * 16ull + (uint64_t)j * 16ull;
      // written by a peer over NVLink: read around L1
      const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(src));
      const int64_t id = (int64_t)v.x;
      if (id >= 0) {
        e.hi = ordered_key64(__longlong_as_double((long long)v.y));
        e.lo = (uint64_t)id;   // node ids are non-negative; NaN distances still order by id
      }
    }
    buf[i] = e;
  }
  __syncthreads();
  bitonic_sort_pairs(buf, sort_cap);
  for (uint32_t j = tid; j < k; j += blockDim.x) {
    const Pair128 e = buf[j];
    const bool ok = e.lo != ~0ull;
    out_ids[j] = ok ? (int64_t)e.lo : -1;
    out_dist[j] = ok ? key64_to_double(e.hi) : __longlong_as_double(0x7FF8000000000000ll);
    if (ok) atomicAdd(&s_xcount, 1u);
  }
  __syncthreads();
  if (tid == 0) *out_count = s_xcount;
}

// Consumer: tell every source that this rank has finished reading epoch x.epoch.
__device__ __forceinline__ void xchg_ack(const XchgParams &x) {
  __threadfence_system();
  if (threadIdx.x < x.n_ranks) {
    uint32_t *ack = reinterpret_cast<uint32_t *>(x.peer[threadIdx.x] + x.ack_off) + x.rank;
    st_release_sys(ack, x.epoch);
  }
}

}  // namespace tsc
