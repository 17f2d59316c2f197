// tsc_select.cuh — K5: candidate selection, exact fp64 re-rank, final ordering.
//
// One CTA per query:
//   1. pick the K' best (fp32 key, row) composites out of the M the scan / GEMM
//      kernels published (shared-memory bitonic sort when M is small, 8-pass
//      radix select otherwise);
//   2. re-rank them with the reference's exact arithmetic —
//      `_exactDistance` core/ngh_graph_engine.dart:908-946: fp32 inputs widened
//      to fp64, sequential index order, every multiply and add a separate IEEE
//      double operation (no FMA: __dmul_rn/__dadd_rn), sqrt / divide correctly
//      rounded — so the distances returned are bit-identical to the Dart code;
//   3. drop `distance > threshold` (:127), sort ascending with Dart's
//      double.compareTo order (-0.0 < 0.0, NaN last; ties by node id), cut at k
//      (:133-134).
#pragma once

#include "tsc_common.cuh"

namespace tsc {

struct SelectParams {
  const uint64_t *cand;     // [nq][m] composites (ordered key << 32 | shard row)
  uint32_t m;               // candidates per query
  uint32_t kprime;          // candidates re-ranked (<= kMaxRerank)
  uint32_t k;               // results per query
  const uint8_t *rows;      // shard rows, device storage dtype
  uint32_t row_bytes;
  uint32_t dims;
  const float *queries;     // [nq, qld] fp32
  uint32_t qld;
  int metric;
  double threshold;         // NaN = none
  int64_t first_node_id;
  int64_t *out_ids;         // [nq, k]
  double *out_dist;         // [nq, k]
  uint32_t *out_counts;     // [nq]
  uint32_t sort_cap;        // pow2 >= m when m <= kSelectSortMax, else pow2 >= kprime
};

constexpr uint32_t kSelectSortMax = 4096;  // M above this goes through radix select
constexpr uint32_t kMaxRerank = 512;
constexpr int kSelectThreads = 512;

struct Pair128 {
  uint64_t hi, lo;
};
__device__ __forceinline__ bool pair_gt(const Pair128 &a, const Pair128 &b) {
  return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
}

// block-wide bitonic sort of n (power of two) pairs in shared memory, ascending
__device__ __forceinline__ void bitonic_sort_pairs(Pair128 *v, uint32_t n) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint32_t x = i ^ j;
        if (x > i) {
          Pair128 a = v[i], b = v[x];
          bool up = (i & k) == 0;
          if (pair_gt(a, b) == up) {
            v[i] = b;
            v[x] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <int DTYPE>
__device__ __forceinline__ float load_elem(const uint8_t *row, uint32_t i) {
  if (DTYPE == kF32) return reinterpret_cast<const float *>(row)[i];
  if (DTYPE == kBF16)
    return __uint_as_float((uint32_t)reinterpret_cast<const uint16_t *>(row)[i] << 16);
  return __half2float(reinterpret_cast<const __half *>(row)[i]);
}

// `_exactDistance` (ngh_graph_engine.dart:908-918) for one stored row.
template <int DTYPE>
__device__ double exact_distance(const float *q, const uint8_t *row, uint32_t d, int metric) {
  if (metric == kL2) {  // :920-927
    double sum = 0.0;
    for (uint32_t i = 0; i < d; i++) {
      double diff = __dsub_rn((double)q[i], (double)load_elem<DTYPE>(row, i));
      sum = __dadd_rn(sum, __dmul_rn(diff, diff));
    }
    return sqrt(sum);
  }
  if (metric == kIP) {  // :929-935, negated at :914
    double sum = 0.0;
    for (uint32_t i = 0; i < d; i++)
      sum = __dadd_rn(sum, __dmul_rn((double)q[i], (double)load_elem<DTYPE>(row, i)));
    return -sum;
  }
  double dot = 0.0, ma = 0.0, mb = 0.0;  // :937-946
  for (uint32_t i = 0; i < d; i++) {
    double a = (double)q[i], b = (double)load_elem<DTYPE>(row, i);
    dot = __dadd_rn(dot, __dmul_rn(a, b));
    ma = __dadd_rn(ma, __dmul_rn(a, a));
    mb = __dadd_rn(mb, __dmul_rn(b, b));
  }
  double denom = __dmul_rn(sqrt(ma), sqrt(mb));
  double sim = denom > 0.0 ? __ddiv_rn(dot, denom) : 0.0;
  return __dsub_rn(1.0, sim);
}

// dynamic smem: Pair128[sort_cap] | float q[qld] (+ small statics)
__host__ __device__ inline size_t select_smem_bytes(uint32_t sort_cap, uint32_t qld) {
  return (size_t)sort_cap * sizeof(Pair128) + (size_t)qld * 4 + 16;
}

template <int DTYPE>
__global__ void __launch_bounds__(kSelectThreads) select_rerank_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Pair128 *buf = reinterpret_cast<Pair128 *>(smem);
  float *qs = reinterpret_cast<float *>(smem + (size_t)p.sort_cap * sizeof(Pair128));
  __shared__ uint32_t hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_remaining, s_count;

  const uint32_t q = blockIdx.x;
  const uint64_t *cand = p.cand + (size_t)q * p.m;
  const uint32_t tid = threadIdx.x;

  for (uint32_t i = tid; i < p.qld; i += blockDim.x) qs[i] = p.queries[(size_t)q * p.qld + i];

  // ---- 1. K' best composites -> buf[0 .. ncand) ------------------------------
  uint32_t ncand;
  if (p.m <= kSelectSortMax) {
    for (uint32_t i = tid; i < p.sort_cap; i += blockDim.x) {
      buf[i].hi = (i < p.m) ? cand[i] : ~0ull;
      buf[i].lo = 0;
    }
    __syncthreads();
    bitonic_sort_pairs(buf, p.sort_cap);
    ncand = p.kprime < p.m ? p.kprime : p.m;
  } else {
    // radix select, 8 bits per pass from the top: the K'-th smallest composite
    if (tid == 0) {
      s_prefix = 0;
      s_remaining = p.kprime;
      s_count = 0;
    }
    for (int pass = 0; pass < 8; pass++) {
      const int shift = 56 - 8 * pass;
      for (uint32_t i = tid; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const uint64_t prefix = s_prefix;
      const uint64_t mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
      for (uint32_t i = tid; i < p.m; i += blockDim.x) {
        uint64_t v = cand[i];
        if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 0xFF], 1u);
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0: each lane owns 8 buckets; find the bucket where the running
        // count crosses the remaining rank (it exists: m >= kprime entries match)
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          loc[i] = hist[tid * 8 + i];
          sum += loc[i];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
          if ((int)tid >= o) inc += t;
        }
        const uint32_t exc = inc - sum, rem = s_remaining;
        if (exc < rem && rem <= inc) {
          uint32_t cum = exc, b = tid * 8;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (cum + loc[i] >= rem) {
              b = tid * 8 + i;
              break;
            }
            cum += loc[i];
          }
          s_remaining = rem - cum;
          s_prefix = prefix | ((uint64_t)b << shift);
        }
      }
      __syncthreads();
    }
    const uint64_t pivot = s_prefix;
    for (uint32_t i = tid; i < p.sort_cap; i += blockDim.x) {
      buf[i].hi = ~0ull;
      buf[i].lo = 0;
    }
    __syncthreads();
    for (uint32_t i = tid; i < p.m; i += blockDim.x) {
      uint64_t v = cand[i];
      if (v <= pivot && (uint32_t)v != kInvalidRow) {
        uint32_t slot = atomicAdd(&s_count, 1u);
        if (slot < p.sort_cap) buf[slot].hi = v;
      }
    }
    __syncthreads();
    ncand = s_count < p.kprime ? s_count : p.kprime;
  }
  __syncthreads();

  // ---- 2. exact fp64 re-rank, one thread per candidate ------------------------
  // (spread over warps so the sequential fp64 chains run on all four schedulers)
  {
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t c = (tid & 31) * nwarps + (tid >> 5);  // candidate index for this thread
    for (uint32_t i = c; i < p.sort_cap; i += blockDim.x) {
      uint64_t v = buf[i].hi;
      uint32_t row = (uint32_t)v;
      if (i < ncand && row != kInvalidRow) {
        double d = exact_distance<DTYPE>(qs, p.rows + (size_t)row * p.row_bytes, p.dims, p.metric);
        bool drop = (p.threshold == p.threshold) && (d > p.threshold);
        buf[i].hi = drop ? ~0ull : ordered_key64(d);
        buf[i].lo = drop ? ~0ull : (uint64_t)row;
      } else {
        buf[i].hi = ~0ull;
        buf[i].lo = ~0ull;
      }
    }
  }
  __syncthreads();

  // ---- 3. final order + emit ----------------------------------------------------
  uint32_t n2 = 1;
  while (n2 < ncand) n2 <<= 1;
  if (n2 < 2) n2 = 2;
  if (n2 > p.sort_cap) n2 = p.sort_cap;
  bitonic_sort_pairs(buf, n2);
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (uint32_t j = tid; j < p.k; j += blockDim.x) {
    bool ok = j < n2 && buf[j].lo != ~0ull;
    int64_t id = -1;
    double d = __longlong_as_double(0x7FF8000000000000ll);
    if (ok) {
      uint64_t kk = buf[j].hi;
      id = p.first_node_id + (int64_t)buf[j].lo;
      if (kk != ~0ull) {
        uint64_t b = (kk & 0x8000000000000000ull) ? (kk & 0x7FFFFFFFFFFFFFFFull) : ~kk;
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

// ---- shard merge: [n_parts][nq][k] -> [nq][k] ------------------------------------
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
