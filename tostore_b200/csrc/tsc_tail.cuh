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
//      are bit-identical to the Dart code. The ADDITIONS are sequential by definition, so
//      ONE LANE per candidate applies them strictly left to right: a warp re-ranks 32
//      candidates in the time of one and a lane's critical path is one dependent DADD
//      (~10 cycles) per element, with the element's widening, subtract and multiply issued
//      in its shadow. The candidates' rows are staged in shared memory by bulk async copies
//      (one per candidate, all in flight at once: ONE memory round trip), in column chunks
//      with two buffers when K' whole rows do not fit;
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
constexpr uint32_t kHeadSelectCap = 1024;  // composites the head-pivot filter may let through
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
  uint32_t trunc_len;       // > 0: cand[q] is m / trunc_len unsorted lists that each kept only
                            // their trunc_len < K' best rows (tensor path, large k)
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
  int defer_retry;          // no range launch follows (pipelined searches): an uncertified query is
                            // only flagged kFlagRetry; the caller re-issues it
  uint32_t *range_count;    // [kRangeSlots] rows collected per slot of the running range pass
  uint64_t *range_buf;      // [kRangeSlots][kRangeCap] composites
  unsigned long long *stat; // [kStatSlots] kStat*
  unsigned long long *diag; // phase timestamps (diagnostics build), else NULL
};

// shared memory of the tail (dynamic):
//   Pair128[sort_cap] | hist[kRadixBins] | q[qld] fp64 | row stage
// The row stage holds, per chain of a batch, `chunk` bytes of the candidate's row at a stride
// of an odd number of 16-byte units (a quarter warp's LDS.128 of eight consecutive chains hits
// eight different bank groups), once (whole rows fit) or twice (column chunks, double buffered).
__host__ __device__ inline size_t tail_fixed_bytes(uint32_t sort_cap, uint32_t qld) {
  return (((size_t)sort_cap * sizeof(Pair128) + (size_t)kRadixBins * 4 + (size_t)qld * 8) + 127) &
         ~(size_t)127;
}
__host__ __device__ inline uint32_t tail_stage_stride(uint32_t chunk_bytes) {
  return ((chunk_bytes >> 4) & 1u) ? chunk_bytes : chunk_bytes + 16u;
}
constexpr uint32_t kTailMinChunk = 128;   // bytes of a row per chain and buffer, at least
// what a launch should provide: whole rows for `chains` chains when that stays within `limit`,
// otherwise `limit` (the tail then works in column chunks and / or several batches)
__host__ __device__ inline size_t tail_smem_bytes(uint32_t sort_cap, uint32_t qld, uint32_t row_bytes,
                                                  uint32_t chains, size_t limit) {
  const size_t fixed = tail_fixed_bytes(sort_cap, qld);
  const size_t whole = fixed + (size_t)chains * tail_stage_stride(row_bytes) + 64;
  const size_t least = fixed + (size_t)2 * tail_stage_stride(kTailMinChunk) + 64;
  if (whole <= limit) return whole;
  return limit > least ? limit : least;
}
__host__ __device__ inline uint32_t tail_sort_cap(uint32_t m, uint32_t kprime, uint32_t list_len,
                                                  bool range) {
  uint32_t need = range ? kRangeCap : (m <= kSelectSortMax ? m : kprime);
  (void)list_len;
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
    __syncwarp();   // every lane has read s_remaining before the one that owns the bucket rewrites it
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
  double *qd = reinterpret_cast<double *>(hist + kRadixBins);   // the query, widened
  __shared__ uint64_t s_prefix, s_pivot;
  __shared__ uint32_t s_remaining, s_count, s_bucket, s_valid;
  __shared__ double s_mag_a;

  const uint32_t tid = threadIdx.x;
  const uint64_t *cand = p.cand + (size_t)qi * p.m;
  // Selection over sorted lists (the scan's per-CTA lists). With L lists, t = ceil(K' / L) and
  // r = ceil(K' / t): the r-th smallest of the lists' t-th entries, H, is >= the K'-th smallest
  // composite overall (the r lists at or below it hold t entries <= H each). Everything <= H
  // is collected (a few more than K' entries: each list is a random 1 / L sample of the rows)
  // and ranked by counting. Only a PREFIX of every list is read: P = max(8, 4t) entries, one
  // round trip together with the query; a list whose whole prefix is <= H is walked further.
  const uint32_t n_lists = p.list_len ? p.m / p.list_len : 0;
  const uint32_t lvl = n_lists ? (p.kprime + n_lists - 1) / n_lists : 0;          // t
  const uint32_t lvl_rank = lvl ? (p.kprime + lvl - 1) / lvl : 0;                  // r
  uint32_t pre = 4 * lvl < 8 ? 8 : 4 * lvl;                                        // P
  if (pre > p.list_len) pre = p.list_len;
  uint64_t *prefix = reinterpret_cast<uint64_t *>(sm + tail_fixed_bytes(sort_cap, p.qld));
  uint64_t *picked = prefix + (size_t)n_lists * pre;   // [kHeadSelectCap]
  const bool use_heads =
      mode == 0 && p.m > kSelectSortMax && p.list_len > 0 && lvl <= p.list_len &&
      n_lists <= kRadixBins / 2 &&
      tail_fixed_bytes(sort_cap, p.qld) + ((size_t)n_lists * pre + kHeadSelectCap) * 8 <= sm_bytes;
  __syncthreads();   // the caller's use of `sm` is over
  if (use_heads) {
    for (uint32_t idx = tid; idx < n_lists * pre; idx += blockDim.x) {
      const uint32_t l = idx / pre, j = idx - l * pre;
      prefix[idx] = __ldcg(cand + (size_t)l * p.list_len + j);
    }
  }
  for (uint32_t i = tid; i < p.qld; i += blockDim.x)
    qd[i] = (double)p.queries[(size_t)qi * p.qld + i];
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
    uint64_t *heads = reinterpret_cast<uint64_t *>(hist);
    for (uint32_t l = tid; l < n_lists; l += blockDim.x) heads[l] = prefix[(size_t)l * pre + lvl - 1];
    __syncthreads();
    for (uint32_t l = tid; l < n_lists; l += blockDim.x) {
      const uint64_t v = heads[l];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n_lists; j++) {
        const uint64_t w = heads[j];
        rank += (w < v) || (w == v && j < l);   // empty entries tie at ~0: order them by index
      }
      if (rank == lvl_rank - 1) s_pivot = v;
    }
    __syncthreads();
    const uint64_t H = s_pivot;
    if ((uint32_t)(H >> 32) != kEmptyKey) {   // else: fewer than K' entries in reach -> radix path
      for (uint32_t idx = tid; idx < n_lists * pre; idx += blockDim.x) {
        uint64_t v = prefix[idx];
        if (v > H) continue;
        uint32_t sl = atomicAdd(&s_count, 1u);
        if (sl < kHeadSelectCap) picked[sl] = v;
        const uint32_t l = idx / pre;
        if (idx - l * pre == pre - 1) {   // the whole prefix passed: walk on
          for (uint32_t j = pre; j < p.list_len; j++) {
            v = __ldcg(cand + (size_t)l * p.list_len + j);
            if (v > H) break;
            sl = atomicAdd(&s_count, 1u);
            if (sl < kHeadSelectCap) picked[sl] = v;
          }
        }
      }
      __syncthreads();
      const uint32_t cnt = s_count;   // >= K' by construction
      if (cnt <= kHeadSelectCap) {
        // composites are distinct (distinct rows): counting the smaller ones is the rank
        for (uint32_t i = tid; i < cnt; i += blockDim.x) {
          const uint64_t v = picked[i];
          uint32_t rank = 0;
          for (uint32_t j = 0; j < cnt; j++) rank += picked[j] < v;
          if (rank < p.kprime) {
            buf[rank].hi = v;
            buf[rank].lo = 0;
          }
          if (rank == p.kprime - 1) s_pivot = v;
        }
        __syncthreads();
        ncand = p.kprime;
        pivot = s_pivot;
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
  // Truncated lists (tensor path, K' above the kernel's list capacity): a row outside the
  // union of the lists was displaced by trunc_len better rows of ITS list, so its key is >= the
  // largest key of that (full) list. The certificate's "no candidate below this" bound is the
  // smaller of the K'-th smallest composite and every full list's largest key.
  if (mode == 0 && p.trunc_len != 0) {
    if (tid == 0) s_bucket = kEmptyKey;
    __syncthreads();
    const uint32_t nl = p.m / p.trunc_len;
    for (uint32_t l = tid; l < nl; l += blockDim.x) {
      uint32_t mx = 0;
      bool full = true;
      for (uint32_t j = 0; j < p.trunc_len; j++) {
        const uint64_t v = __ldcg(cand + (size_t)l * p.trunc_len + j);
        if ((uint32_t)v == kInvalidRow) full = false;
        else if ((uint32_t)(v >> 32) > mx) mx = (uint32_t)(v >> 32);
      }
      if (full) atomicMin(&s_bucket, mx);
    }
    __syncthreads();
    const uint32_t excl = s_bucket;
    if (excl != kEmptyKey) {
      all_in = false;
      if ((uint32_t)(pivot >> 32) > excl) pivot = (uint64_t)excl << 32;
    }
    __syncthreads();
  }
  TSC_TRACE(p.diag, 2);

  // ---- 2. exact fp64 re-rank -------------------------------------------------------------
  // Chain 0 is the query's own |q|^2 (magA of _cosineSimlarity, also the certificate's
  // ||q||^2), chains 1..ncand the candidates; thread c of a batch owns chain base + c. Every
  // thread with a live candidate issues ONE bulk copy of its row (or of the current column
  // chunk of it) into its stage slot and arrives on the buffer's mbarrier with the byte count;
  // everybody else just arrives. The chain lanes then walk their slot 16 bytes at a time:
  // widen, (subtract,) multiply — each an individually rounded IEEE operation on exactly
  // widened fp32 values — and add to the running sum, strictly in index order.
  constexpr int E = Chunk<DTYPE>::kElems;
  constexpr uint32_t kEsz = 16 / E;
  __shared__ __align__(8) uint64_t s_bar[2];
  uint8_t *stage = sm + tail_fixed_bytes(sort_cap, p.qld);
  const uint32_t need_bytes = ((p.dims + E - 1) / E) * 16u;   // <= row_bytes
  const uint32_t n_chains = ncand + 1;
  // batch geometry: `lanes` chains side by side, `chunk` bytes of each row per buffer
  uint32_t lanes = n_chains < blockDim.x ? n_chains : blockDim.x;
  uint32_t chunk = 0, nbuf = 1;
  {
    const size_t fixed = tail_fixed_bytes(sort_cap, p.qld);
    const size_t avail = sm_bytes > fixed + 64 ? sm_bytes - fixed - 64 : 0;
    for (;;) {
      if ((size_t)lanes * tail_stage_stride(need_bytes) <= avail) {
        chunk = need_bytes;
        nbuf = 1;
        break;
      }
      uint32_t c = (uint32_t)(avail / ((size_t)2 * lanes)) & ~15u;
      if (c >= 16 && tail_stage_stride(c) * (size_t)2 * lanes > avail) c -= 16;
      if (c >= kTailMinChunk || lanes == 1) {
        chunk = c;
        nbuf = 2;
        break;
      }
      lanes = (lanes + 1) >> 1;
    }
  }
  const uint32_t stride = tail_stage_stride(chunk);
  const uint32_t n_chunks = chunk ? (need_bytes + chunk - 1) / chunk : 0;
  const uint64_t policy = policy_evict_first();
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), blockDim.x);
    mbar_init(smem_u32(&s_bar[1]), blockDim.x);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t uses0 = 0, uses1 = 0;   // completed phases of the two barriers (uniform)
  for (uint32_t base = 0; base < n_chains && chunk != 0; base += lanes) {
    const uint32_t nb = n_chains - base < lanes ? n_chains - base : lanes;
    const uint32_t chain = base + tid;
    const bool mine = tid < nb;
    const bool is_q = mine && chain == 0;
    uint32_t row = kInvalidRow;
    if (mine && chain > 0) row = (uint32_t)buf[chain - 1].hi;
    const bool has_row = row != kInvalidRow;
    // chunk t of every chain of the batch -> buffer t & (nbuf - 1)
    auto issue = [&](uint32_t t) {
      const uint32_t b = t & (nbuf - 1u);
      const uint32_t bar = smem_u32(&s_bar[b]);
      const uint32_t off = t * chunk;
      const uint32_t bytes = need_bytes - off < chunk ? need_bytes - off : chunk;
      if (has_row) {
        mbar_expect_tx(bar, bytes);   // arrive + expect
        bulk_g2s(smem_u32(stage + ((size_t)b * lanes + tid) * stride),
                 p.rows + (size_t)row * p.row_bytes + off, bytes, bar, policy);
      } else {
        mbar_arrive(bar);
      }
    };
    issue(0);
    if (nbuf == 2 && n_chunks > 1) issue(1);
    double s0 = 0.0, s1 = 0.0;   // +0.0: `double sum = 0.0` of the Dart loops
    for (uint32_t t = 0; t < n_chunks; t++) {
      const uint32_t b = t & (nbuf - 1u);
      if (mine) {
        mbar_wait(smem_u32(&s_bar[b]), (b ? uses1 : uses0) & 1u);
        const uint32_t off = t * chunk;
        const uint32_t bytes = need_bytes - off < chunk ? need_bytes - off : chunk;
        const uint32_t e0 = off / kEsz;                       // first element of the chunk
        const uint32_t ne = p.dims - e0 < bytes / kEsz ? p.dims - e0 : bytes / kEsz;
        // a chain without a row (the query's, an empty candidate slot) reads slot 0's bytes
        // and discards them
        const uint4 *src = reinterpret_cast<const uint4 *>(
            stage + ((size_t)b * lanes + (has_row ? tid : 0u)) * stride);
        const double *qa = qd + e0;
        const uint32_t nv = ne / E;
#pragma unroll 2
        for (uint32_t v = 0; v < nv; v++) {
          float bf[E];
          Chunk<DTYPE>::unpack(src[v], bf);
#pragma unroll
          for (int e = 0; e < E; e += 2) {
            const double2 a2 = *reinterpret_cast<const double2 *>(qa + v * E + e);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const double a = h ? a2.y : a2.x;
              const double bw = (double)bf[e + h];
              if (METRIC == kL2) {
                const double x = is_q ? a : __dsub_rn(a, bw);
                s0 = __dadd_rn(s0, __dmul_rn(x, x));
              } else {
                const double x = is_q ? a : bw;
                s0 = __dadd_rn(s0, __dmul_rn(a, x));
                if (METRIC == kCos) s1 = __dadd_rn(s1, __dmul_rn(bw, bw));
              }
            }
          }
        }
        if (nv * E < ne) {   // dims is not a multiple of the vector width: the last few
          float bf[E];
          Chunk<DTYPE>::unpack(src[nv], bf);
#pragma unroll
          for (int e = 0; e < E; e++) {
            if (nv * E + e < ne) {
              const double a = qa[nv * E + e];
              const double bw = (double)bf[e];
              if (METRIC == kL2) {
                const double x = is_q ? a : __dsub_rn(a, bw);
                s0 = __dadd_rn(s0, __dmul_rn(x, x));
              } else {
                const double x = is_q ? a : bw;
                s0 = __dadd_rn(s0, __dmul_rn(a, x));
                if (METRIC == kCos) s1 = __dadd_rn(s1, __dmul_rn(bw, bw));
              }
            }
          }
        }
      }
      if (b) uses1++;
      else uses0++;
      __syncthreads();   // the buffer is free again
      if (t + nbuf < n_chunks) issue(t + nbuf);
    }
    if (is_q) s_mag_a = s0;
    __syncthreads();
    if (mine && chain > 0) {
      const uint32_t ci = chain - 1;
      uint64_t hi = ~0ull, lo = ~0ull;
      if (has_row) {
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
  if (tid == 0) {
    mbar_inval(smem_u32(&s_bar[0]));
    mbar_inval(smem_u32(&s_bar[1]));
  }
  TSC_TRACE(p.diag, 3);

  // ---- 3. final order -------------------------------------------------------------------
  // up to 32: bitonic network in registers of warp 0; up to 1024: rank by counting, out of
  // place through the (now free) row stage; beyond (range pass): bitonic sort in place
  uint32_t n2 = 2;
  while (n2 < ncand) n2 <<= 1;
  if (n2 > sort_cap) n2 = sort_cap;
  for (uint32_t i = ncand + tid; i < n2; i += blockDim.x) {
    buf[i].hi = ~0ull;
    buf[i].lo = ~0ull;
  }
  __syncthreads();
  if (n2 > 32 && ncand <= 1024 &&
      tail_fixed_bytes(sort_cap, p.qld) + (size_t)ncand * sizeof(Pair128) <= sm_bytes) {
    Pair128 *sorted = reinterpret_cast<Pair128 *>(stage);
    for (uint32_t i = tid; i < ncand; i += blockDim.x) {
      const Pair128 a = buf[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < ncand; j++) {
        const Pair128 o = buf[j];
        rank += pair_gt(a, o) || (a.hi == o.hi && a.lo == o.lo && j < i);   // dropped ones tie
      }
      sorted[rank] = a;
    }
    __syncthreads();
    for (uint32_t i = tid; i < ncand; i += blockDim.x) buf[i] = sorted[i];
    __syncthreads();
  } else {
    bitonic_sort_pairs(buf, n2);
  }
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
        if (!p.defer_retry) {
          p.range_thr[qi] = thr_key;
          const uint32_t s = atomicAdd(p.retry_n, 1u);
          p.retry_list[s] = qi;
        }
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
