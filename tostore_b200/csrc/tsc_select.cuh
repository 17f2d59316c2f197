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

constexpr uint32_t kSelectSortMax = 1024;  // M above this goes through radix select
constexpr uint32_t kMaxRerank = 512;
constexpr int kSelectThreads = 512;
constexpr int kSelectWarps = kSelectThreads / 32;
constexpr int kRadixBins = 2048;   // 11-bit digits
constexpr int kProdChunk = 64;     // elements whose products one warp stages at a time

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

// `_exactDistance` (core/ngh_graph_engine.dart:908-946) for one stored row, by one
// warp. Every product is an individually rounded IEEE double multiply, so the
// lanes may compute them in parallel; the additions are then applied strictly
// left to right (i = 0..d-1, starting from +0.0) by a single lane per running
// sum, which is what makes the result bit-identical to the Dart loop. Products
// are staged through shared memory in chunks of kProdChunk, double buffered so
// the next chunk's global loads overlap the current chunk's add chain.
//   pbuf: [2 buffers][2 arrays][kProdChunk] doubles, private to the warp
//   mag_a: sum of q[i]^2 (cosine only; the same for every candidate of a query)
template <int DTYPE>
__device__ double warp_exact_distance(const float *q, const uint8_t *row, uint32_t d, int metric,
                                      double mag_a, double *pbuf, int lane) {
  constexpr int PER = kProdChunk / 32;
  double s0 = 0.0, s1 = 0.0;  // lane 0: dot / l2 sum, lane 1: magB
  const uint32_t nchunk = (d + kProdChunk - 1) / kProdChunk;
  float bv[PER];
  auto fetch = [&](uint32_t c) {
#pragma unroll
    for (int j = 0; j < PER; j++) {
      uint32_t i = c * kProdChunk + j * 32 + lane;
      bv[j] = i < d ? load_elem<DTYPE>(row, i) : 0.0f;
    }
  };
  fetch(0);
  for (uint32_t c = 0; c < nchunk; c++) {
    double *p0 = pbuf + (size_t)(c & 1) * 2 * kProdChunk;
    double *p1 = p0 + kProdChunk;
#pragma unroll
    for (int j = 0; j < PER; j++) {
      uint32_t e = j * 32 + lane, i = c * kProdChunk + e;
      double a = i < d ? (double)q[i] : 0.0, b = (double)bv[j];
      if (metric == kL2) {
        double diff = __dsub_rn(a, b);
        p0[e] = __dmul_rn(diff, diff);
      } else {
        p0[e] = __dmul_rn(a, b);
        if (metric == kCos) p1[e] = __dmul_rn(b, b);
      }
    }
    __syncwarp();
    if (c + 1 < nchunk) fetch(c + 1);
    const uint32_t n = min((uint32_t)kProdChunk, d - c * kProdChunk);
    if (lane < 2) {  // lane 0: dot / l2 chain, lane 1: magB chain, in lockstep
      const double *src = lane == 0 ? p0 : p1;
      double acc = lane == 0 ? s0 : s1;
      if (lane == 0 || metric == kCos) {
#pragma unroll 8
        for (uint32_t e = 0; e < n; e++) acc = __dadd_rn(acc, src[e]);
      }
      if (lane == 0) s0 = acc; else s1 = acc;
    }
    // the buffer written next iteration is the other one; the one after that is
    // this one again, and by then every lane has passed the __syncwarp above
  }
  s1 = __shfl_sync(0xFFFFFFFFu, s1, 1);
  s0 = __shfl_sync(0xFFFFFFFFu, s0, 0);
  if (metric == kL2) return sqrt(s0);                      // :920-927
  if (metric == kIP) return -s0;                           // :929-935, negated at :914
  double denom = __dmul_rn(sqrt(mag_a), sqrt(s1));         // :937-946
  double sim = denom > 0.0 ? __ddiv_rn(s0, denom) : 0.0;
  return __dsub_rn(1.0, sim);
}

// sum of q[i]^2, sequential (magA of _cosineSimlarity); q padded with zeros
__device__ double warp_mag_a(const float *q, uint32_t d, double *pbuf, int lane) {
  double s = 0.0;
  for (uint32_t c0 = 0; c0 < d; c0 += kProdChunk) {
    for (int e = lane; e < kProdChunk; e += 32) {
      double a = (c0 + e) < d ? (double)q[c0 + e] : 0.0;
      pbuf[e] = __dmul_rn(a, a);
    }
    __syncwarp();
    uint32_t n = min((uint32_t)kProdChunk, d - c0);
    if (lane == 0)
      for (uint32_t e = 0; e < n; e++) s = __dadd_rn(s, pbuf[e]);
    __syncwarp();
  }
  return __shfl_sync(0xFFFFFFFFu, s, 0);
}

// dynamic smem: Pair128[sort_cap] | double pbuf[warps][4*kProdChunk] | float q[qld]
__host__ __device__ inline size_t select_smem_bytes(uint32_t sort_cap, uint32_t qld) {
  return (size_t)sort_cap * sizeof(Pair128) + (size_t)kSelectWarps * 4 * kProdChunk * 8 +
         (size_t)qld * 4 + 16;
}

// One radix-select digit pass over the composites: histogram the `bits`-wide digit
// at `shift` of every entry matching (prefix, mask); warp 0 finds the bucket where
// the running count crosses s_remaining and narrows the prefix.
__device__ __forceinline__ void radix_pass(const uint64_t *cand, uint32_t m, int shift, int bits,
                                           uint64_t mask, uint32_t *hist, uint64_t *s_prefix,
                                           uint32_t *s_remaining, uint32_t *s_bucket_count) {
  const uint32_t tid = threadIdx.x;
  const uint32_t nb = 1u << bits;
  for (uint32_t i = tid; i < nb; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const uint64_t prefix = *s_prefix;
  for (uint32_t i = tid; i < m; i += blockDim.x) {
    uint64_t v = cand[i];
    if ((v & mask) == prefix) atomicAdd(&hist[(uint32_t)(v >> shift) & (nb - 1)], 1u);
  }
  __syncthreads();
  if (tid < 32) {
    const uint32_t per = nb / 32;  // buckets per lane (nb >= 32)
    uint32_t sum = 0;
    for (uint32_t i = 0; i < per; i++) sum += hist[tid * per + i];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if ((int)tid >= o) inc += t;
    }
    const uint32_t exc = inc - sum, rem = *s_remaining;
    if (exc < rem && rem <= inc) {  // exactly one lane: m >= remaining entries match
      uint32_t cum = exc, b = tid * per;
      for (uint32_t i = 0; i < per; i++) {
        uint32_t h = hist[tid * per + i];
        if (cum + h >= rem) {
          b = tid * per + i;
          *s_bucket_count = h;
          break;
        }
        cum += h;
      }
      *s_remaining = rem - cum;
      *s_prefix = prefix | ((uint64_t)b << shift);
    }
  }
  __syncthreads();
}

template <int DTYPE>
__global__ void __launch_bounds__(kSelectThreads) select_rerank_kernel(const SelectParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  Pair128 *buf = reinterpret_cast<Pair128 *>(smem);
  double *pbuf_all = reinterpret_cast<double *>(smem + (size_t)p.sort_cap * sizeof(Pair128));
  float *qs = reinterpret_cast<float *>(pbuf_all + (size_t)kSelectWarps * 4 * kProdChunk);
  __shared__ uint32_t hist[kRadixBins];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_remaining, s_count, s_bucket;
  __shared__ double s_mag_a;

  const uint32_t q = blockIdx.x;
  const uint64_t *cand = p.cand + (size_t)q * p.m;
  const uint32_t tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  double *pbuf = pbuf_all + (size_t)warp * 4 * kProdChunk;

  for (uint32_t i = tid; i < p.qld; i += blockDim.x) qs[i] = p.queries[(size_t)q * p.qld + i];
  if (tid == 0) {
    s_prefix = 0;
    s_remaining = p.kprime;
    s_count = 0;
    s_bucket = 0;
  }
  __syncthreads();

  // ---- 1. K' best composites -> buf[0 .. ncand).hi (unordered) ------------------
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
    // radix select of the K'-th smallest composite: three 11/11/10-bit passes over
    // the key half; the id half is only walked when the pivot key is shared by
    // more entries than are still needed (ties at the cut).
    radix_pass(cand, p.m, 53, 11, 0ull, hist, &s_prefix, &s_remaining, &s_bucket);
    radix_pass(cand, p.m, 42, 11, ~0ull << 53, hist, &s_prefix, &s_remaining, &s_bucket);
    radix_pass(cand, p.m, 32, 10, ~0ull << 42, hist, &s_prefix, &s_remaining, &s_bucket);
    uint64_t pivot;
    if (s_bucket == s_remaining) {
      pivot = s_prefix | 0xFFFFFFFFull;
    } else {
      radix_pass(cand, p.m, 21, 11, ~0ull << 32, hist, &s_prefix, &s_remaining, &s_bucket);
      radix_pass(cand, p.m, 10, 11, ~0ull << 21, hist, &s_prefix, &s_remaining, &s_bucket);
      radix_pass(cand, p.m, 0, 10, ~0ull << 10, hist, &s_prefix, &s_remaining, &s_bucket);
      pivot = s_prefix;
    }
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

  // ---- 2. exact fp64 re-rank, one warp per candidate ---------------------------------
  if (p.metric == kCos) {
    if (warp == 0) {
      double m = warp_mag_a(qs, p.dims, pbuf, lane);
      if (lane == 0) s_mag_a = m;
    }
    __syncthreads();
  }
  const double mag_a = p.metric == kCos ? s_mag_a : 0.0;
  for (uint32_t i = warp; i < p.sort_cap; i += kSelectWarps) {
    uint64_t v = buf[i].hi;
    uint32_t row = (uint32_t)v;
    uint64_t hi = ~0ull, lo = ~0ull;
    if (i < ncand && row != kInvalidRow) {
      double d = warp_exact_distance<DTYPE>(qs, p.rows + (size_t)row * p.row_bytes, p.dims,
                                            p.metric, mag_a, pbuf, lane);
      bool drop = (p.threshold == p.threshold) && (d > p.threshold);  // :127
      if (!drop) {
        hi = ordered_key64(d);
        lo = (uint64_t)row;
      }
    }
    __syncwarp();
    if (lane == 0) {
      buf[i].hi = hi;
      buf[i].lo = lo;
    }
  }
  __syncthreads();

  // ---- 3. final order + emit -------------------------------------------------------------
  uint32_t n2 = 2;
  while (n2 < ncand) n2 <<= 1;
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
