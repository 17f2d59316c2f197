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
  if ((b & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0xFFFFFFFFFFFFFFFFull;
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// One row of the program. `load(slot, key, isnull)` fetches the row's value of a column
// as an order-preserving key; shared by the kernel and by the host self-test
// (tsc_selftest_where), so the CPU test-suite exercises the same evaluation code.
template <class Load>
__host__ __device__ __forceinline__ bool where_eval_row(const WhereProgram &prog,
                                                        const uint64_t *args, Load load) {
  uint64_t stack = 0;
  uint32_t last_col = 0xFFFFFFFFu;
  uint64_t key = 0;
  bool isnull = true;
  for (uint32_t i = 0; i < prog.n_ops; i++) {
    const WhereDevOp &op = prog.ops[i];
    bool r;
    if (op.kind == kWLeaf) {
      if (op.op < kOpTrue && op.col != last_col) {
        last_col = op.col;
        load(op.col, key, isnull);
      }
      switch (op.op) {
        case kOpEq: r = !isnull && key == op.lo; break;
        case kOpNe: r = isnull || key != op.lo; break;
        case kOpGt: r = !isnull && key > op.lo; break;
        case kOpGe: r = !isnull && key >= op.lo; break;
        case kOpLt: r = !isnull && key < op.lo; break;
        case kOpLe: r = !isnull && key <= op.lo; break;
        case kOpBetween: r = !isnull && key >= op.lo && key <= op.hi; break;
        case kOpIn:
        case kOpNotIn: {
          bool any = false;
          for (uint32_t j = 0; j < op.n; j++) any |= (key == args[op.args_off + j]);
          r = op.op == kOpIn ? (!isnull && any) : (isnull || !any);
          break;
        }
        case kOpIsNull: r = isnull; break;
        case kOpIsNotNull: r = !isnull; break;
        case kOpTrue: r = true; break;
        default: r = false; break;
      }
      stack = (stack << 1) | (r ? 1ull : 0ull);
    } else {
      // n-ary AND / OR over the top n stack bits (n <= 63, checked on the host)
      const uint32_t n = op.n;
      const uint64_t m = (1ull << n) - 1ull;
      const uint64_t top = stack & m;
      r = (n == 0) ? true : (op.kind == kWAnd ? top == m : top != 0);
      stack = ((stack >> n) << 1) | (r ? 1ull : 0ull);
    }
  }
  return prog.n_ops == 0 || (stack & 1ull);
}

// One thread per row, one warp per 32-row bitmap word (ballot), grid-stride over words.
// HBM traffic: 8 bytes per row per leaf (coalesced 256-byte warp loads) + 4 bytes per 32
// rows written; a leaf on the column the previous leaf used re-uses the loaded value.
__global__ void __launch_bounds__(256)
where_eval_kernel(const __grid_constant__ WhereProgram prog, const __grid_constant__ WhereCols cols,
                  const uint64_t *__restrict__ args, uint64_t n_rows, uint32_t n_slots,
                  uint32_t *__restrict__ out_bits, unsigned long long *__restrict__ matched) {
  const int lane = threadIdx.x & 31;
  const uint64_t n_words = (n_rows + 31) / 32;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long local = 0;
  // The evaluator loads a column when the program first needs it, one dependent load after
  // the other: a warp had ~256 bytes in flight and the kernel ran at 24 % of HBM (ncu,
  // profiles/r02_first_call.log). Every column's lines of the words two iterations ahead are
  // therefore pulled into L2 first (two 128-byte lines per column and word).
  auto prefetch = [&](uint64_t w) {
    if (w < n_words && (lane & 15) == 0) {
      const uint64_t row = w * 32 + (uint64_t)lane;
      for (uint32_t c = 0; c < n_slots; c++)
        if (row < n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(cols.values[c] + row));
    }
  };
  const uint64_t w0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  prefetch(w0);
  prefetch(w0 + warps);
  for (uint64_t w = w0; w < n_words; w += warps) {
    prefetch(w + 2 * warps);
    const uint64_t row = w * 32 + lane;
    bool res = false;
    if (row < n_rows) {
      res = where_eval_row(prog, args, [&](uint32_t c, uint64_t &key, bool &isnull) {
        const uint64_t raw = __ldg(cols.values[c] + row);
        key = cols.is_f64[c] ? where_key_f64_bits(raw) : (raw ^ 0x8000000000000000ull);
        isnull = (__ldg(cols.nulls[c] + (row >> 5)) >> (row & 31)) & 1u;
      });
    }
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, res);
    if (lane == 0) {
      out_bits[w] = bits;
      local += __popc(bits);
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
