// tsc_tail.cuh — K5: the tail of every search — candidate selection, exact fp64 re-rank,
// exactness certificate, final ordering.
//
// One CTA works on one query (as a standalone kernel: one CTA per query; fused into the
// scan kernel: the last CTA to finish, tsc_scan.cuh):
//   1. pick the K' best (fp32 key, row) composites out of the M the scan / GEMM kernels
//      published. Scan lists are sorted, so the K'-th smallest of the list HEADS bounds the
//      K'-th smallest overall: one ranking pass over the heads, one filtering pass over the M
//      composites, one small sort. (General fallback: 11-bit radix select; M <= 1024: sort.)
//   2. re-rank them with the reference's exact arithmetic — `_exactDistance`
//      core/ngh_graph_engine.dart:908-946: fp32 inputs widened to fp64, sequential index
//      order, every multiply and add a separate IEEE double operation (no FMA:
//      __dmul_rn/__dadd_rn), sqrt / divide correctly rounded — so the distances returned
//      are bit-identical to the Dart code. The PRODUCTS are independent of each other, so
//      the whole CTA computes them tile by tile into shared memory (each one an individually
//      rounded multiply); the ADDITIONS are sequential by definition, so ONE LANE per
//      candidate applies them strictly left to right — a warp re-ranks 32 candidates in the
//      time of one, and a lane's critical path is one dependent DADD per element;
//   3. drop `distance > threshold` (:127), sort ascending with Dart's double.compareTo
//      order (-0.0 < 0.0, NaN last; ties by node id), cut at k (:133-134);
//   4. CERTIFY the candidate stage. The reference re-ranks everything it kept (:115-134);
//      this path keeps K' rows chosen by an approximate fp32 key, so it has to prove it
//      kept enough: with kappa_piv the K'-th smallest key (every row that is not a
//      candidate has key >= kappa_piv), D_adm the largest distance that could still enter
//      the result, K*(D) the key an error-free candidate stage would give a row at
//      distance D and |key - K*| <= c_rel |K*| + A the proven error of the stage
//      (CertModel, filled by the host per metric / dtype / path):
//          certified  <=>  K*(D_adm) < (kappa_piv - A) / (1 + c_rel)
//      Uncertified queries get a threshold T >= every key a row at distance <= D_adm can
//      have and are re-run by the RANGE pass (scan kernel, mode 1): it collects every row
//      with key <= T, this tail re-ranks all of them, and the result is exact regardless
//      of how many near-ties surround the k-th neighbour (up to kRangeCap rows).
#pragma once

#include "tsc_common.cuh"

namespace tsc {

constexpr uint32_t kSelectSortMax = 1024;  // M up to this is simply sorted
constexpr uint32_t kHeadSelectCap = 256;   // composites the head-pivot filter may let through
constexpr uint32_t kMaxRerank = 512;
constexpr int kRadixBins = 2048;           // 11-bit digits
constexpr uint32_t kRangeCap = 4096;       // rows the range pass can hold per query
constexpr uint32_t kRangeSlots = 8;        // queries per range pass (= the scan kernel's QB max)
constexpr uint32_t kRetryLaunches = 4;     // in-stream range passes per search
constexpr int kProdChunk = 64;             // elements per product tile

// per-query flag values (TailParams::flags)
enum : uint32_t { kFlagExact = 0, kFlagRetry = 1, kFlagUncertified = 2 };
// TailParams::stat slots
enum : int { kStatCertified = 0, kStatRetried = 1, kStatUncertified = 2, kStatRangeRows = 3,
             kStatRetryAsked = 4, kStatSlots = 8 };

// Diagnostics build only (-DTSC_DIAG, libtostore_cuda_diag.so): phase timestamps of the scan
// kernel and its tail (globaltimer ns) for tools/scan_trace.py. Compiles to nothing otherwise.
#ifdef TSC_DIAG
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TSC_TRACE(ptr, slot)                                            \
  do {                                                                  \
    if ((ptr) != nullptr && threadIdx.x == 0) (ptr)[slot] = trace_now(); \
  } while (0)
#else
#define TSC_TRACE(ptr, slot) \
  do {                       \
  } while (0)
#endif
// trace slots: 0 first CTA start (atomicMin), 1 tail begin, 2 candidates selected, 3 chains
// done, 4 sorted, 5 certificate done, 6 emitted, 7 exchange done;
// 16 + cta: main loop end, 16 + grid + cta: candidates published
constexpr int kTraceCta = 16;

struct Pair128 {
  uint64_t hi, lo;
};
__device__ __forceinline__ bool pair_gt(const Pair128 &a, const Pair128 &b) {
  return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo);
}

// Bitonic sort of n (power of two) pairs in shared memory, ascending, by the whole CTA.
// Up to 64 elements warp 0 does it alone (two per lane, __syncwarp between stages): a
// CTA-wide barrier per stage costs more than the stage. Ends with a CTA-wide barrier.
__device__ __forceinline__ void bitonic_sort_pairs(Pair128 *v, uint32_t n) {
  if (n <= 32) {
    // one element per lane of warp 0, compare-exchange through shuffles: no memory traffic
    if (threadIdx.x < 32) {
      const uint32_t i = threadIdx.x;
      Pair128 a{~0ull, ~0ull};
      if (i < n) a = v[i];
#pragma unroll
      for (uint32_t k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          Pair128 b;
          b.hi = __shfl_xor_sync(0xFFFFFFFFu, a.hi, j);
          b.lo = __shfl_xor_sync(0xFFFFFFFFu, a.lo, j);
          const bool take_min = ((i & j) == 0) == ((i & k) == 0);
          if (take_min ? pair_gt(a, b) : pair_gt(b, a)) a = b;
        }
      }
      if (i < n) v[i] = a;
    }
    __syncthreads();
    return;
  }
  if (n <= 64) {
    if (threadIdx.x < 32) {
      for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = threadIdx.x; i < n; i += 32) {
            const uint32_t x = i ^ j;
            if (x > i) {
              const Pair128 a = v[i], b = v[x];
              const bool up = (i & k) == 0;
              if (pair_gt(a, b) == up) {
                v[i] = b;
                v[x] = a;
              }
            }
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
    return;
  }
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

__device__ __forceinline__ double key64_to_double(uint64_t kk) {
  if (kk == ~0ull) return __longlong_as_double(0x7FF8000000000000ll);
  uint64_t b = (kk & 0x8000000000000000ull) ? (kk & 0x7FFFFFFFFFFFFFFFull) : ~kk;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ float key32_to_float(uint32_t uk) {
  uint32_t b = (uk & 0x80000000u) ? (uk & 0x7FFFFFFFu) : ~uk;
  return __uint_as_float(b);
}

template <int DTYPE>
__device__ __forceinline__ float load_elem(const uint8_t *row, uint32_t i) {
  if (DTYPE == kF32) return __ldg(reinterpret_cast<const float *>(row) + i);
  if (DTYPE == kBF16)
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const uint16_t *>(row) + i) << 16);
  return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const uint16_t *>(row) + i)));
}

// Error model of a candidate stage in key space: |key - K*| <= c_rel * |K*| + A(q),
// A(q) = a_q * ||q|| + a_e * ||q16 - q|| + a_0. K*(D) = D^2 (L2; minus ||q||^2 when
// l2_shift: the tensor path's keys omit that constant), D (inner product: D = -dot),
// (D - 1) * ||q|| (cosine: the stages never divide by ||q||).
struct CertModel {
  float c_rel, a_q, a_e, a_0;
  int l2_shift;
};

struct TailParams {
  const uint64_t *cand;     // [nq][m] composites (ordered key << 32 | shard row), first pass
  uint32_t m;               // candidates per query
  uint32_t list_len;        // > 0: cand[q] is m / list_len lists, each sorted ascending (scan)
  uint32_t kprime;          // candidates re-ranked by the first pass (<= kMaxRerank)
  uint32_t k;               // results per query
  const uint8_t *rows;      // shard rows, device storage dtype
  uint32_t row_bytes;
  uint32_t dims;
  const float *queries;     // [nq, qld] fp32
  uint32_t qld;
  double threshold;         // NaN = none
  int64_t first_node_id;
  int64_t *out_ids;         // [nq, k]
  double *out_dist;         // [nq, k]
  uint32_t *out_counts;     // [nq]
  CertModel cert;           // error model of the stage that produced `cand`
  CertModel cert_range;     // error model of the range pass (scan kernel, exact fp32 query)
  const float *enorm;       // [nq] ||q16 - q||_2 (tensor path) or NULL
  uint32_t *flags;          // [nq] kFlag*
  uint32_t *range_thr;      // [nq] ordered fp32 key T: the range pass collects key <= T
  uint32_t *retry_list;     // compacted indices of the queries that need the range pass
  uint32_t *retry_n;
  uint32_t *range_count;    // [kRangeSlots] rows collected per slot of the running range pass
  uint64_t *range_buf;      // [kRangeSlots][kRangeCap] composites
  unsigned long long *stat; // [kStatSlots] kStat*
  unsigned long long *diag; // phase timestamps (diagnostics build), else NULL
};

// shared memory of the tail (dynamic):
//   Pair128[sort_cap] | hist[kRadixBins] | q[qld] fp32 | product tiles
// A product tile is [2 buffers][1 or 2 arrays][kProdChunk][lanes | 1] doubles; `lanes`
// chains (candidates + the query's own |q|^2) run side by side per batch.
__host__ __device__ inline size_t tail_fixed_bytes(uint32_t sort_cap, uint32_t qld) {
  return (((size_t)sort_cap * sizeof(Pair128) + (size_t)kRadixBins * 4 + (size_t)qld * 4) + 15) &
         ~(size_t)15;
}
__host__ __device__ inline size_t tail_tile_bytes(uint32_t lanes, bool cosine) {
  return (size_t)2 * (cosine ? 2 : 1) * kProdChunk * (lanes | 1u) * 8;
}
__host__ __device__ inline size_t tail_smem_bytes(uint32_t sort_cap, uint32_t qld, uint32_t lanes,
                                                  bool cosine) {
  return tail_fixed_bytes(sort_cap, qld) + tail_tile_bytes(lanes, cosine) + 64;
}
__host__ __device__ inline uint32_t tail_sort_cap(uint32_t m, uint32_t kprime, uint32_t list_len,
                                                  bool range) {
  uint32_t need = range ? kRangeCap : (m <= kSelectSortMax ? m : kprime);
  if (!range && m > kSelectSortMax && list_len > 0 && need < kHeadSelectCap) need = kHeadSelectCap;
  uint32_t p = 2;
  while (p < need) p <<= 1;
  return p;
}

// 16 bytes of a stored row -> fp32 lanes (tsc_scan.cuh)
template <int DTYPE>
struct Chunk;
template <>
struct Chunk<kF32> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
};
template <>
struct Chunk<kBF16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[8]) {
    unpack_bf16x2(v.x, f[0], f[1]);
    unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]);
    unpack_bf16x2(v.w, f[6], f[7]);
  }
};
template <>
struct Chunk<kF16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[8]) {
    unpack_f16x2(v.x, f[0], f[1]);
    unpack_f16x2(v.y, f[2], f[3]);
    unpack_f16x2(v.z, f[4], f[5]);
    unpack_f16x2(v.w, f[6], f[7]);
  }
};

//   mag_a: sum of q[i]^2 (cosine only; the same for every candidate of a query)
template <int METRIC>
__device__ __forceinline__ double exact_finish(double s0, double s1, double mag_a) {
  if (METRIC == kL2) return sqrt(s0);                      // :920-927
  if (METRIC == kIP) return -s0;                           // :929-935, negated at :914
  const double denom = __dmul_rn(sqrt(mag_a), sqrt(s1));   // :937-946
  const double sim = denom > 0.0 ? __ddiv_rn(s0, denom) : 0.0;
  return __dsub_rn(1.0, sim);
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
    uint64_t v = __ldcg(cand + i);
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

// K*(D): the key an error-free candidate stage would give a row at exact distance D
__device__ __forceinline__ double key_star(int metric, double D, double qn, double q2,
                                           int l2_shift) {
  if (metric == kL2) return l2_shift ? D * D - q2 : D * D;
  if (metric == kIP) return D;
  return (D - 1.0) * qn;
}

// The tail for query `qi`, executed by every thread of the CTA (any blockDim that is a
// multiple of 32). mode 0: first pass over cand[qi][0..m); mode 1: range pass over
// range_buf[slot][0..range_count[slot]). `sm` = the CTA's dynamic shared memory, sm_bytes of
// it (>= tail_smem_bytes(sort_cap, qld, 1, cosine)).
// (__noinline__: one copy per (metric, dtype) and translation unit, shared by all the scan
// kernel variants that call it, instead of one inlined copy in each of them.)
template <int METRIC, int DTYPE>
__device__ __noinline__ void tail_query(const TailParams &p, uint32_t qi, int mode, uint32_t slot,
                                        uint8_t *sm, size_t sm_bytes, uint32_t sort_cap) {
  Pair128 *buf = reinterpret_cast<Pair128 *>(sm);
  uint32_t *hist = reinterpret_cast<uint32_t *>(sm + (size_t)sort_cap * sizeof(Pair128));
  float *qs = reinterpret_cast<float *>(hist + kRadixBins);
  __shared__ uint64_t s_prefix, s_pivot;
  __shared__ uint32_t s_remaining, s_count, s_bucket, s_valid;
  __shared__ double s_mag_a;

  const uint32_t tid = threadIdx.x;
  const uint64_t *cand = p.cand + (size_t)qi * p.m;
  // head-pivot selection applies: sorted lists, at least K' of them, heads fit the scratch
  const bool use_heads = mode == 0 && p.m > kSelectSortMax && p.list_len > 0 &&
                         p.m / p.list_len >= p.kprime && p.m / p.list_len <= kRadixBins / 2 &&
                         sort_cap >= kHeadSelectCap;
  // Everything the selection needs from global memory is requested up front, in ONE round
  // trip: the query, the list heads and this thread's share of the composites (registers).
  constexpr int kOwn = 16;
  uint64_t own[kOwn];
  const bool own_ok = use_heads && p.m <= (uint32_t)kOwn * blockDim.x;
  __syncthreads();   // the caller's use of `sm` is over
  if (own_ok) {
#pragma unroll
    for (int j = 0; j < kOwn; j++) {
      const uint32_t i = tid + (uint32_t)j * blockDim.x;
      own[j] = i < p.m ? __ldcg(cand + i) : ~0ull;
    }
  }
  if (use_heads) {
    uint64_t *heads = reinterpret_cast<uint64_t *>(hist);
    const uint32_t nl = p.m / p.list_len;
    for (uint32_t l = tid; l < nl; l += blockDim.x) heads[l] = __ldcg(cand + (size_t)l * p.list_len);
  }
  for (uint32_t i = tid; i < p.qld; i += blockDim.x) qs[i] = p.queries[(size_t)qi * p.qld + i];
  if (tid == 0) {
    s_prefix = 0;
    s_pivot = ~0ull;
    s_remaining = p.kprime;
    s_count = 0;
    s_bucket = 0;
    s_valid = 0;
  }
  __syncthreads();

  // ---- 1. candidates -> buf[0 .. ncand).hi (unordered), pivot, all_in ------------------
  uint32_t ncand = 0;
  uint64_t pivot = ~0ull;   // K'-th smallest composite of the first pass
  bool all_in = false;      // every live row of the shard is a candidate
  bool overflow = false;    // range pass: more rows than kRangeCap
  bool selected = false;
  if (mode == 1) {
    const uint32_t cnt = __ldcg(p.range_count + slot);
    overflow = cnt > kRangeCap;
    ncand = overflow ? 0u : cnt;
    const uint64_t *src = p.range_buf + (size_t)slot * kRangeCap;
    for (uint32_t i = tid; i < sort_cap; i += blockDim.x) {
      buf[i].hi = i < ncand ? __ldcg(src + i) : ~0ull;
      buf[i].lo = 0;
    }
    all_in = true;   // by construction: every row that can matter was collected
    selected = true;
  } else if (p.m <= kSelectSortMax) {
    uint32_t valid = 0;
    for (uint32_t i = tid; i < sort_cap; i += blockDim.x) {
      const uint64_t v = (i < p.m) ? __ldcg(cand + i) : ~0ull;
      buf[i].hi = v;
      buf[i].lo = 0;
      valid += (uint32_t)v != kInvalidRow;
    }
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    bitonic_sort_pairs(buf, sort_cap);
    const uint32_t nv = s_valid;
    ncand = p.kprime < nv ? p.kprime : nv;
    all_in = nv < p.kprime;
    if (!all_in) pivot = buf[p.kprime - 1].hi;
    selected = true;
  } else if (use_heads) {
    // Sorted lists: H = the K'-th smallest list head is >= the K'-th smallest composite
    // overall (the K' heads at or below it are K' composites <= H). Rank the heads against
    // each other, let every composite <= H through, sort the few that pass.
    const uint32_t nl = p.m / p.list_len;
    const uint64_t *heads = reinterpret_cast<const uint64_t *>(hist);   // loaded above
    for (uint32_t i = tid; i < sort_cap; i += blockDim.x) {
      buf[i].hi = ~0ull;
      buf[i].lo = 0;
    }
    for (uint32_t l = tid; l < nl; l += blockDim.x) {
      const uint64_t v = heads[l];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nl; j++) {
        const uint64_t w = heads[j];
        rank += (w < v) || (w == v && j < l);   // empty lists tie at ~0: order them by index
      }
      if (rank == p.kprime - 1) s_pivot = v;
    }
    __syncthreads();
    const uint64_t H = s_pivot;
    if ((uint32_t)(H >> 32) != kEmptyKey) {   // else: fewer than K' non-empty lists -> radix path
      if (own_ok) {
#pragma unroll
        for (int j = 0; j < kOwn; j++)
          if (own[j] <= H) {
            const uint32_t s = atomicAdd(&s_count, 1u);
            if (s < sort_cap) buf[s].hi = own[j];
          }
      } else {
        for (uint32_t i = tid; i < p.m; i += blockDim.x) {
          const uint64_t v = __ldcg(cand + i);
          if (v <= H) {
            const uint32_t s = atomicAdd(&s_count, 1u);
            if (s < sort_cap) buf[s].hi = v;
          }
        }
      }
      __syncthreads();
      const uint32_t cnt = s_count;   // >= K' by construction
      if (cnt <= sort_cap) {
        uint32_t n2 = 2;
        while (n2 < cnt) n2 <<= 1;
        bitonic_sort_pairs(buf, n2);
        ncand = p.kprime;
        pivot = buf[p.kprime - 1].hi;
        selected = true;
      }
      __syncthreads();
      if (tid == 0) s_count = 0;
      __syncthreads();
    }
  }
  if (!selected) {
    // radix select of the K'-th smallest composite: three 11/11/10-bit passes over
    // the key half; the id half is only walked when the pivot key is shared by
    // more entries than are still needed (ties at the cut).
    radix_pass(cand, p.m, 53, 11, 0ull, hist, &s_prefix, &s_remaining, &s_bucket);
    radix_pass(cand, p.m, 42, 11, ~0ull << 53, hist, &s_prefix, &s_remaining, &s_bucket);
    radix_pass(cand, p.m, 32, 10, ~0ull << 42, hist, &s_prefix, &s_remaining, &s_bucket);
    if (s_bucket == s_remaining) {
      pivot = s_prefix | 0xFFFFFFFFull;
    } else {
      radix_pass(cand, p.m, 21, 11, ~0ull << 32, hist, &s_prefix, &s_remaining, &s_bucket);
      radix_pass(cand, p.m, 10, 11, ~0ull << 21, hist, &s_prefix, &s_remaining, &s_bucket);
      radix_pass(cand, p.m, 0, 10, ~0ull << 10, hist, &s_prefix, &s_remaining, &s_bucket);
      pivot = s_prefix;
    }
    for (uint32_t i = tid; i < sort_cap; i += blockDim.x) {
      buf[i].hi = ~0ull;
      buf[i].lo = 0;
    }
    __syncthreads();
    for (uint32_t i = tid; i < p.m; i += blockDim.x) {
      const uint64_t v = __ldcg(cand + i);
      if (v <= pivot && (uint32_t)v != kInvalidRow) {
        const uint32_t s = atomicAdd(&s_count, 1u);
        if (s < sort_cap) buf[s].hi = v;
      }
    }
    __syncthreads();
    ncand = s_count < p.kprime ? s_count : p.kprime;
    // fewer than K' valid composites: the K'-th smallest is an empty slot (key 0xFFFFFFFF)
    all_in = (uint32_t)(pivot >> 32) == kEmptyKey;
  }
  __syncthreads();
  TSC_TRACE(p.diag, 2);

  // ---- 2. exact fp64 re-rank -------------------------------------------------------------
  // Chain 0 is the query's own |q|^2 (magA of _cosineSimlarity, also the certificate's
  // ||q||^2), chains 1..ncand the candidates. Per tile of kProdChunk elements every thread
  // computes products (each an individually rounded IEEE multiply of exactly widened fp32
  // values: order-free) into shared memory, transposed so that the chain lanes read
  // consecutive words; then lane c adds its chain's tile left to right. The products of the
  // next tile are prepared while the lanes add the current one (two buffers).
  constexpr int NA = METRIC == kCos ? 2 : 1;
  double *tiles = reinterpret_cast<double *>(sm + tail_fixed_bytes(sort_cap, p.qld));
  uint32_t lanes_max = 0;
  {
    const size_t fixed = tail_fixed_bytes(sort_cap, p.qld);
    const size_t avail = sm_bytes > fixed + 64 ? sm_bytes - fixed - 64 : 0;
    lanes_max = (uint32_t)(avail / ((size_t)2 * NA * kProdChunk * 8));
    if (lanes_max > 1 && !(lanes_max & 1u)) lanes_max--;   // the padded width (lanes | 1) must fit
    if (lanes_max > blockDim.x) lanes_max = blockDim.x;
  }
  // pull the candidates' rows towards L2 (they were streamed with evict-first)
  if (ncand <= kMaxRerank) {
    const uint32_t lines = (p.row_bytes + 127) / 128;
    for (uint32_t idx = tid; idx < ncand * lines; idx += blockDim.x) {
      const uint32_t row = (uint32_t)buf[idx / lines].hi;
      if (row != kInvalidRow)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.rows + (size_t)row * p.row_bytes +
                                                     (size_t)(idx % lines) * 128));
    }
  }
  const uint32_t n_chains = ncand + 1;
  const uint32_t n_tiles = (p.dims + kProdChunk - 1) / kProdChunk;
  for (uint32_t base = 0; base < n_chains && lanes_max > 0; base += lanes_max) {
    const uint32_t nb = n_chains - base < lanes_max ? n_chains - base : lanes_max;
    const uint32_t lp = nb | 1u;                        // padded tile width
    const size_t tile_doubles = (size_t)NA * kProdChunk * lp;
    // Products of tile t -> tiles[t & 1]. Item idx = (chain r, element e): consecutive threads
    // take consecutive elements of one row (coalesced), the tile is stored transposed.
    // Small batches (the usual K' + 1 chains) split the work in two so that the row loads of
    // tile t + 1 are in flight while the lanes add tile t: load_tile / store_tile.
    constexpr int kPf = 4;
    const bool pf = nb * kProdChunk <= kPf * blockDim.x;
    float bv[kPf];
    auto item = [&](uint32_t idx, uint32_t t, uint32_t &r, uint32_t &e, uint32_t &i, uint32_t &row) {
      r = idx / kProdChunk;
      e = idx - r * kProdChunk;
      i = t * kProdChunk + e;
      const uint32_t c = base + r;
      row = c == 0 ? kInvalidRow : (uint32_t)buf[c - 1].hi;
    };
    auto product = [&](double *dst, uint32_t r, uint32_t e, uint32_t i, bool is_q, bool has_row,
                       float bf) {
      double p0 = 0.0, p1 = 0.0;
      if (i < p.dims) {
        const double a = (double)qs[i];
        if (is_q) {
          p0 = __dmul_rn(a, a);
        } else if (has_row) {
          const double b = (double)bf;
          if (METRIC == kL2) {
            const double diff = __dsub_rn(a, b);
            p0 = __dmul_rn(diff, diff);
          } else {
            p0 = __dmul_rn(a, b);
            if (METRIC == kCos) p1 = __dmul_rn(b, b);
          }
        }
      }
      dst[(size_t)e * lp + r] = p0;
      if (METRIC == kCos) dst[(size_t)(kProdChunk + e) * lp + r] = p1;
    };
    auto load_tile = [&](uint32_t t) {
#pragma unroll
      for (int j = 0; j < kPf; j++) {
        const uint32_t idx = tid + (uint32_t)j * blockDim.x;
        bv[j] = 0.0f;
        if (idx < nb * kProdChunk) {
          uint32_t r, e, i, row;
          item(idx, t, r, e, i, row);
          if (row != kInvalidRow && i < p.dims)
            bv[j] = load_elem<DTYPE>(p.rows + (size_t)row * p.row_bytes, i);
        }
      }
    };
    auto store_tile = [&](uint32_t t) {
      double *dst = tiles + (size_t)(t & 1u) * tile_doubles;
#pragma unroll
      for (int j = 0; j < kPf; j++) {
        const uint32_t idx = tid + (uint32_t)j * blockDim.x;
        if (idx < nb * kProdChunk) {
          uint32_t r, e, i, row;
          item(idx, t, r, e, i, row);
          product(dst, r, e, i, base + r == 0, row != kInvalidRow, bv[j]);
        }
      }
    };
    auto make_tile = [&](uint32_t t) {
      double *dst = tiles + (size_t)(t & 1u) * tile_doubles;
      for (uint32_t idx = tid; idx < nb * kProdChunk; idx += blockDim.x) {
        uint32_t r, e, i, row;
        item(idx, t, r, e, i, row);
        float bf = 0.0f;
        if (row != kInvalidRow && i < p.dims)
          bf = load_elem<DTYPE>(p.rows + (size_t)row * p.row_bytes, i);
        product(dst, r, e, i, base + r == 0, row != kInvalidRow, bf);
      }
    };
    make_tile(0);
    __syncthreads();
    double s0 = 0.0, s1 = 0.0;   // +0.0: `double sum = 0.0` of the Dart loops
    for (uint32_t t = 0; t < n_tiles; t++) {
      if (t + 1 < n_tiles) {
        if (pf) load_tile(t + 1);
        else make_tile(t + 1);
      }
      if (tid < nb) {
        const double *src = tiles + (size_t)(t & 1u) * tile_doubles + tid;
        // elements past dims are +0.0 products: adding them changes nothing (a running
        // sum that started at +0.0 is never -0.0)
#pragma unroll 8
        for (int e = 0; e < kProdChunk; e++) {
          s0 = __dadd_rn(s0, src[(size_t)e * lp]);
          if (METRIC == kCos) s1 = __dadd_rn(s1, src[(size_t)(kProdChunk + e) * lp]);
        }
      }
      if (pf && t + 1 < n_tiles) store_tile(t + 1);
      __syncthreads();
    }
    if (base == 0 && tid == 0) s_mag_a = s0;
    __syncthreads();
    if (tid < nb && base + tid > 0) {
      const uint32_t ci = base + tid - 1;
      const uint32_t row = (uint32_t)buf[ci].hi;
      uint64_t hi = ~0ull, lo = ~0ull;
      if (row != kInvalidRow) {
        const double d = exact_finish<METRIC>(s0, s1, s_mag_a);
        const bool drop = (p.threshold == p.threshold) && (d > p.threshold);  // :127
        if (!drop) {
          hi = ordered_key64(d);
          lo = (uint64_t)row;
        }
      }
      buf[ci].hi = hi;
      buf[ci].lo = lo;
    }
    __syncthreads();
  }
  TSC_TRACE(p.diag, 3);

  // ---- 3. final order -------------------------------------------------------------------
  uint32_t n2 = 2;
  while (n2 < ncand) n2 <<= 1;
  if (n2 > sort_cap) n2 = sort_cap;
  for (uint32_t i = ncand + tid; i < n2; i += blockDim.x) {
    buf[i].hi = ~0ull;
    buf[i].lo = ~0ull;
  }
  __syncthreads();
  bitonic_sort_pairs(buf, n2);
  TSC_TRACE(p.diag, 4);

  // ---- 4. certificate (first pass) / verdict (range pass), by one thread ---------------------
  if (tid == 0) {
    uint32_t kept = 0;   // results that survive the threshold, up to k
    while (kept < p.k && kept < n2 && buf[kept].lo != ~0ull) kept++;
    s_count = kept;
    uint32_t flag = kFlagExact;
    if (mode == 1) {
      if (overflow) flag = kFlagUncertified;
      atomicAdd(p.stat + (overflow ? kStatUncertified : kStatRetried), 1ull);
      atomicAdd(p.stat + kStatRangeRows, (unsigned long long)__ldcg(p.range_count + slot));
      p.range_count[slot] = 0;
    } else if (!all_in) {
      const double q2 = s_mag_a, qn = sqrt(q2);
      const double en = p.enorm ? (double)p.enorm[qi] : 0.0;
      // D_adm: a row at a smaller distance would enter the result
      double d_adm;
      if (kept == p.k) {
        d_adm = key64_to_double(buf[p.k - 1].hi);
      } else {
        d_adm = (p.threshold == p.threshold) ? p.threshold
                                             : __longlong_as_double(0x7FF0000000000000ll);
      }
      const double inf = __longlong_as_double(0x7FF0000000000000ll);
      if (!(d_adm == d_adm)) d_adm = inf;   // NaN ranks last: anything would enter before it
      const bool zero_q = METRIC != kL2 && q2 == 0.0;   // every key and distance is the same
      bool ok = zero_q;
      uint32_t thr_key = kEmptyKey - 1u;
      if (!zero_q) {
        const double A = (double)p.cert.a_q * qn + (double)p.cert.a_e * en + (double)p.cert.a_0;
        double ks = key_star(METRIC, d_adm, qn, q2, p.cert.l2_shift);
        ks += fabs(ks) * 1e-12;
        const double kpiv = (double)key32_to_float((uint32_t)(pivot >> 32));
        const double lower = (kpiv - A) / (1.0 + (double)p.cert.c_rel);
        ok = ks < lower;   // false for NaN / inf on either side
        if (!ok) {
          // T: no row at distance <= D_adm can have a scan key above it
          const double Ar = (double)p.cert_range.a_q * qn + (double)p.cert_range.a_0;
          double ksr = key_star(METRIC, d_adm, qn, q2, p.cert_range.l2_shift);
          ksr += fabs(ksr) * 1e-12;
          const double t = ksr + fabs(ksr) * (double)p.cert_range.c_rel + Ar;
          float tf = (t == t) ? __double2float_ru(t) : __int_as_float(0x7F800000);
          tf += 0.0f;
          thr_key = ordered_key(tf);
        }
      }
      if (ok) {
        atomicAdd(p.stat + kStatCertified, 1ull);
      } else {
        flag = kFlagRetry;
        atomicAdd(p.stat + kStatRetryAsked, 1ull);
        p.range_thr[qi] = thr_key;
        const uint32_t s = atomicAdd(p.retry_n, 1u);
        p.retry_list[s] = qi;
      }
    } else {
      atomicAdd(p.stat + kStatCertified, 1ull);
    }
    p.flags[qi] = flag;
  }
  __syncthreads();
  TSC_TRACE(p.diag, 5);

  // ---- 5. emit (an overflowed range pass keeps the first pass's best-effort result) --------
  if (!(mode == 1 && overflow)) {
    const uint32_t kept = s_count;
    for (uint32_t j = tid; j < p.k; j += blockDim.x) {
      int64_t id = -1;
      double d = __longlong_as_double(0x7FF8000000000000ll);
      if (j < kept) {
        id = p.first_node_id + (int64_t)buf[j].lo;
        d = key64_to_double(buf[j].hi);
      }
      p.out_ids[(size_t)qi * p.k + j] = id;
      p.out_dist[(size_t)qi * p.k + j] = d;
    }
    if (tid == 0) p.out_counts[qi] = kept;
  }
  __syncthreads();
  TSC_TRACE(p.diag, 6);
}

// standalone form: one CTA per query (tensor path, multi-pass scans)
constexpr int kTailThreads = 256;
template <int METRIC, int DTYPE>
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const TailParams p, uint32_t sort_cap,
                                                            uint32_t smem_bytes) {
  extern __shared__ __align__(16) uint8_t tail_smem[];
  tail_query<METRIC, DTYPE>(p, blockIdx.x, 0, 0, tail_smem, smem_bytes, sort_cap);
}

}  // namespace tsc
