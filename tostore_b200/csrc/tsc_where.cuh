// tsc_where.cuh — K8: structured WHERE prefilter evaluated on the GPU.
//
// SURVEY.md §8f row 3 / BASELINE config 5 ("structured WHERE prefilter + kNN"). The
// reference evaluates a condition tree per record on the CPU
// (ConditionRecordMatcher, handler/value_matcher.dart:337-625) and has no WHERE for
// vectors at all; here the numeric table fields a predicate needs are kept
// column-wise in HBM, aligned by node id, and ONE pass turns a condition tree into the
// liveness bitmap the scan kernels already honour — no PK list, no host round trip.
//
// Semantics restated from the reference (value_matcher.dart):
//   * tree: AND = every child, OR = any child, a childless AND / OR is true (:476-493)
//   * operators (:570-612): '=' cmp==0; '!=' cmp!=0 (so NULL != x is TRUE);
//     '>' '>=' '<' '<=' need a non-null value; IN false / NOT IN true on NULL;
//     BETWEEN = start <= v <= end, false on NULL; IS NULL / IS NOT NULL
//   * numeric order = Dart num.compareTo (:150-174): a total order in which
//     -0.0 < 0.0, NaN is above +inf and equal to itself
// Both column types are mapped to uint64 keys whose unsigned order is that order, so
// one kernel serves int and double fields; operands are converted on the host.
#pragma once

#include "tsc_common.cuh"

namespace tsc {

enum : uint8_t { kWLeaf = 0, kWAnd = 1, kWOr = 2 };
enum : uint8_t {
  kOpEq = 0, kOpNe, kOpGt, kOpGe, kOpLt, kOpLe, kOpBetween, kOpIn, kOpNotIn, kOpIsNull,
  kOpIsNotNull, kOpTrue, kOpFalse, kOpCount
};
constexpr int kWhereMaxOps = 64;     // program length / bit-stack depth
constexpr int kWhereMaxCols = 16;

struct WhereDevOp {        // 32 bytes, operands already in key space
  uint8_t kind, op;
  uint16_t n;              // children of AND / OR, or length of the IN list
  uint32_t col;            // slot in WhereCols
  uint64_t lo, hi;         // operand keys (BETWEEN: start, end)
  uint32_t args_off;       // IN list: first key in `args`
  uint32_t pad;
};

struct WhereCols {
  const uint64_t *values[kWhereMaxCols];  // raw 8-byte values (int64 or double bits)
  const uint32_t *nulls[kWhereMaxCols];   // optional: bit r = row r is NULL
  uint8_t is_f64[kWhereMaxCols];
};

struct WhereProgram {
  WhereDevOp ops[kWhereMaxOps];
  uint32_t n_ops;
};

__host__ __device__ __forceinline__ uint64_t where_key_i64(int64_t v) {
  return (uint64_t)v ^ 0x8000000000000000ull;
}
// Dart double.compareTo order; every NaN is the same (largest) key
__host__ __device__ __forceinline__ uint64_t where_key_f64_bits(uint64_t b) {
  const uint64_t k = (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  return (b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull ? 0xFFFFFFFFFFFFFFFFull : k;
}

// U rows of the program side by side. `load(slot, key[U], isnull[U])` fetches the U rows' values
// of a column as order-preserving keys. The program is walked ONCE for all U rows, so the U loads
// a leaf needs are independent of each other and in flight together (on the device: U
// coalesced 256-byte warp loads per leaf instead of one). Shared by the kernel (U = 4) and by
// the host self-test (tsc_selftest_where, U = 1), so the CPU test-suite exercises the same
// evaluation code.
// Stack = uint32_t for programs of up to 32 steps (the bit-stack cannot get deeper than the
// program is long; 64-bit shifts cost two instructions each), uint64_t otherwise.
template <int U, class Stack, class Load>
__host__ __device__ __forceinline__ void where_eval_rows_t(const WhereProgram &prog,
                                                           const uint64_t *args, Load load,
                                                           bool (&out)[U]) {
  Stack stack[U];
  uint64_t key[U];
  bool isnull[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    stack[u] = 0;
    key[u] = 0;
    isnull[u] = true;
  }
  uint32_t last_col = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < prog.n_ops; i++) {
    const WhereDevOp &op = prog.ops[i];
    if (op.kind == kWLeaf) {
      if (op.op < kOpTrue && op.col != last_col) {
        last_col = op.col;
        load(op.col, key, isnull);
      }
      // Every comparison is a range test in key space: r = NULL ? on_null : (lo <= key <= hi) ^ neg.
      // The operator is decoded ONCE per leaf (uniform over the warp) and the per-row work is two
      // 64-bit compares; IN / NOT IN walk their list. (A per-row switch over the operators was
      // if-converted by the compiler into ~80 instructions per row and leaf: the kernel was
      // issue-bound at 21 % of HBM, profiles/r02_where_*.)
      uint64_t lo = op.lo, hi = op.lo;
      bool neg = false, on_null = false, list = false;
      switch (op.op) {
        case kOpEq: break;
        case kOpNe: neg = true; on_null = true; break;
        case kOpGt: hi = ~0ull; if (lo == ~0ull) { lo = 1; hi = 0; } else lo = lo + 1; break;
        case kOpGe: hi = ~0ull; break;
        case kOpLt: if (hi == 0) { lo = 1; hi = 0; } else { hi = hi - 1; lo = 0; } break;
        case kOpLe: lo = 0; break;
        case kOpBetween: hi = op.hi; break;
        case kOpIn: list = true; break;
        case kOpNotIn: list = true; neg = true; on_null = true; break;
        case kOpIsNull: lo = 1; hi = 0; on_null = true; break;
        case kOpIsNotNull: lo = 0; hi = ~0ull; break;
        case kOpTrue: lo = 0; hi = ~0ull; on_null = true; break;
        default: lo = 1; hi = 0; break;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint64_t k = key[u];
        bool in;
        if (list) {
          in = false;
          for (uint32_t j = 0; j < op.n; j++) in |= (k == args[op.args_off + j]);
        } else {
          in = k >= lo && k <= hi;
        }
        const bool r = isnull[u] ? on_null : (in != neg);
        stack[u] = (Stack)(stack[u] << 1) | (Stack)(r ? 1 : 0);
      }
    } else {
      // n-ary AND / OR over the top n stack bits (n <= 63, checked on the host)
      const uint32_t n = op.n;
      const Stack m = (Stack)(((Stack)1 << (n & (sizeof(Stack) * 8 - 1))) - 1);   // n < width: host-checked
#pragma unroll
      for (int u = 0; u < U; u++) {
        const Stack top = stack[u] & m;
        const bool r = (n == 0) ? true : (op.kind == kWAnd ? top == m : top != 0);
        stack[u] = (Stack)(((stack[u] >> n) << 1) | (Stack)(r ? 1 : 0));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; u++) out[u] = prog.n_ops == 0 || (stack[u] & 1);
}
template <int U, class Load>
__host__ __device__ __forceinline__ void where_eval_rows(const WhereProgram &prog,
                                                         const uint64_t *args, Load load,
                                                         bool (&out)[U]) {
  if (prog.n_ops < 32) where_eval_rows_t<U, uint32_t>(prog, args, load, out);
  else where_eval_rows_t<U, uint64_t>(prog, args, load, out);
}

// one row: `load(slot, key, isnull)`
template <class Load>
__host__ __device__ __forceinline__ bool where_eval_row(const WhereProgram &prog,
                                                        const uint64_t *args, Load load) {
  bool out[1];
  where_eval_rows<1>(prog, args,
                     [&](uint32_t c, uint64_t (&key)[1], bool (&isnull)[1]) {
                       load(c, key[0], isnull[0]);
                     },
                     out);
  return out[0];
}

// One thread per row and bitmap word, kWhereWords consecutive words (128 rows) per warp and
// step: every leaf on a new column is kWhereWords independent coalesced 256-byte warp loads.
// (One word per step left a warp with 256 bytes in flight: 24 % of HBM, ncu in
// profiles/r02_first_call.log; L2 prefetching two steps ahead did not help.)
// HBM traffic: 8 bytes per row per distinct column + 4 bytes per 32 rows written.
constexpr int kWhereWords = 4;
__global__ void __launch_bounds__(256)
where_eval_kernel(const __grid_constant__ WhereProgram prog, const __grid_constant__ WhereCols cols,
                  const uint64_t *__restrict__ args, uint64_t n_rows, uint32_t n_slots,
                  uint32_t *__restrict__ out_bits, unsigned long long *__restrict__ matched) {
  (void)n_slots;
  const int lane = threadIdx.x & 31;
  const uint64_t n_words = (n_rows + 31) / 32;
  const uint64_t n_groups = (n_words + kWhereWords - 1) / kWhereWords;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long local = 0;
  for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups;
       g += warps) {
    const uint64_t row0 = g * kWhereWords * 32 + lane;
    bool res[kWhereWords];
    where_eval_rows<kWhereWords>(
        prog, args,
        [&](uint32_t c, uint64_t (&key)[kWhereWords], bool (&isnull)[kWhereWords]) {
          // all loads first, no branch between them (rows past the end re-read the last row:
          // their result is masked below), then the conversions
          uint64_t raw[kWhereWords];
          uint32_t nb[kWhereWords];
          const uint64_t *vals = cols.values[c];
          const uint32_t *nulls = cols.nulls[c];
#pragma unroll
          for (int u = 0; u < kWhereWords; u++) {
            uint64_t row = row0 + (uint64_t)u * 32;
            row = row < n_rows ? row : n_rows - 1;
            raw[u] = __ldg(vals + row);
            nb[u] = __ldg(nulls + (row >> 5)) >> (row & 31);
          }
          if (cols.is_f64[c] != 0) {   // uniform
#pragma unroll
            for (int u = 0; u < kWhereWords; u++) key[u] = where_key_f64_bits(raw[u]);
          } else {
#pragma unroll
            for (int u = 0; u < kWhereWords; u++) key[u] = raw[u] ^ 0x8000000000000000ull;
          }
#pragma unroll
          for (int u = 0; u < kWhereWords; u++) isnull[u] = nb[u] & 1u;
        },
        res);
#pragma unroll
    for (int u = 0; u < kWhereWords; u++) {
      const uint64_t row = row0 + (uint64_t)u * 32;
      const unsigned bits = __ballot_sync(0xFFFFFFFFu, res[u] && row < n_rows);
      const uint64_t w = g * kWhereWords + u;
      if (lane == 0 && w < n_words) {
        out_bits[w] = bits;
        local += __popc(bits);
      }
    }
  }
  if (lane == 0 && local) atomicAdd(matched, local);
}

// fixed-width columns arrive densely from the host; nulls as one byte per row
__global__ void pack_null_bits_kernel(const uint8_t *__restrict__ is_null, uint64_t n,
                                      uint64_t first_row, uint32_t *__restrict__ null_bits) {
  // rows [first_row, first_row + n): read-modify-write whole words with atomics because the
  // first and last word may be shared with rows appended earlier / later
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = first_row + i;
    const uint32_t bit = 1u << (row & 31);
    if (is_null[i]) atomicOr(null_bits + (row >> 5), bit);
    else atomicAnd(null_bits + (row >> 5), ~bit);
  }
}

}  // namespace tsc
