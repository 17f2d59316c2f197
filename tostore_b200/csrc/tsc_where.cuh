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
//
// TEXT fields (DataType.text) are dictionary-encoded: the distinct strings of a column live
// once in HBM as UTF-16 code units (Dart's own string form, so String.compareTo
// (:211-240) is a code-unit comparison and LIKE's `_` is one code unit), rows hold the
// 32-bit code. A text leaf is evaluated in two steps, both on the GPU: dict_match_kernel
// tests every DISTINCT string once (comparison / IN / BETWEEN / LIKE, :570-604 and
// matchesLike :318-331) into a bitmap over codes, and the row pass turns a row's code into
// that bit — a 12.5M-row column with 1000 distinct values costs 1000 string tests.
#pragma once

#include "tsc_common.cuh"

namespace tsc {

enum : uint8_t { kWLeaf = 0, kWAnd = 1, kWOr = 2 };
enum : uint8_t {
  kOpEq = 0, kOpNe, kOpGt, kOpGe, kOpLt, kOpLe, kOpBetween, kOpIn, kOpNotIn, kOpIsNull,
  kOpIsNotNull, kOpTrue, kOpFalse, kOpLike, kOpNotLike, kOpCount
};
constexpr int kWhereMaxOps = 64;     // program length / bit-stack depth
constexpr int kWhereMaxCols = 16;

// A leaf as the row pass sees it: already a range test in key space,
//   r = NULL ? on_null : ((lo <= key <= hi) != neg)
// (or a list / dictionary-bitmap membership instead of the range). The operator is decoded into
// this form ONCE, on the host (where_decode_leaf), so the row pass carries no operator switch.
enum : uint8_t { kLfNeg = 1, kLfOnNull = 2, kLfList = 4, kLfDict = 8, kLfConst = 16 };
struct WhereDevOp {        // 32 bytes, operands already in key space
  uint8_t kind, flags;     // kW*; kLf* (leaves)
  uint16_t n;              // children of AND / OR, or length of the IN list
  uint32_t col;            // slot in WhereCols
  uint64_t lo, hi;         // the range (lo > hi: empty)
  uint32_t args_off;       // kLfList: first key in `args`; kLfDict: first word in `dict_bits`
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

// Operator (numeric column, operand keys a / b) -> the row pass's range form
// (value_matcher.dart:570-612: '!=' and NOT IN are true on NULL, every ordering operator, IN and
// BETWEEN false; IN / NOT IN keep their list and use no range).
inline void where_decode_leaf(uint8_t op, uint64_t a, uint64_t b, WhereDevOp *d) {
  uint64_t lo = a, hi = a;
  uint8_t fl = 0;
  switch (op) {
    case kOpEq: break;
    case kOpNe: fl = kLfNeg | kLfOnNull; break;
    case kOpGt: if (a == ~0ull) { lo = 1; hi = 0; } else { lo = a + 1; hi = ~0ull; } break;
    case kOpGe: hi = ~0ull; break;
    case kOpLt: if (a == 0) { lo = 1; hi = 0; } else { lo = 0; hi = a - 1; } break;
    case kOpLe: lo = 0; break;
    case kOpBetween: hi = b; break;
    case kOpIn: fl = kLfList; break;
    case kOpNotIn: fl = kLfList | kLfNeg | kLfOnNull; break;
    case kOpIsNull: lo = 1; hi = 0; fl = kLfOnNull; break;
    case kOpIsNotNull: lo = 0; hi = ~0ull; break;
    case kOpTrue: lo = 0; hi = ~0ull; fl = kLfOnNull | kLfConst; break;
    default: lo = 1; hi = 0; fl = kLfConst; break;   // kOpFalse
  }
  d->lo = lo;
  d->hi = hi;
  d->flags = fl;
}

// U rows of the program side by side. `load(slot, key[U], isnull[U])` fetches the U rows' values
// of a column as order-preserving keys. The program is walked ONCE for all U rows, so the U loads
// a leaf needs are independent of each other and in flight together (on the device: U
// coalesced 256-byte warp loads per leaf instead of one). Shared by the kernel (U = 4) and by
// the host self-test (tsc_selftest_where, U = 1), so the CPU test-suite exercises the same
// evaluation code.
// Stack = uint32_t for programs of up to 32 steps (the bit-stack cannot get deeper than the
// program is long; 64-bit shifts cost two instructions each), uint64_t otherwise.
// TEXT: the program holds dictionary leaves (kLfDict).
template <int U, class Stack, bool TEXT, class Load>
__host__ __device__ __forceinline__ void where_eval_rows_t(const WhereProgram &prog,
                                                           const uint64_t *args,
                                                           const uint32_t *dict_bits, Load load,
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
      const uint32_t fl = op.flags;   // uniform over the warp
      if (!(fl & kLfConst) && op.col != last_col) {
        last_col = op.col;
        load(op.col, key, isnull);
      }
      // Per row and leaf: two 64-bit compares. (A per-row switch over the operators was
      // if-converted by the compiler into ~80 instructions per row and leaf — 21 % of HBM,
      // profiles/r02_where_*; decoding it per leaf in the kernel still cost ~1/3 of the
      // instructions of a two-leaf program, so the decode moved to the host.)
      const bool neg = (fl & kLfNeg) != 0, on_null = (fl & kLfOnNull) != 0;
      bool in[U];
      if (fl & kLfList) {
#pragma unroll
        for (int u = 0; u < U; u++) in[u] = false;
        for (uint32_t j = 0; j < op.n; j++) {
          const uint64_t a = args[op.args_off + j];
#pragma unroll
          for (int u = 0; u < U; u++) in[u] |= (key[u] == a);
        }
      } else if (TEXT && (fl & kLfDict)) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          // the key's low 32 bits are the row's dictionary code; a NULL row holds no code
          const uint32_t code = (uint32_t)key[u];
          in[u] = !isnull[u] && ((dict_bits[op.args_off + (code >> 5)] >> (code & 31)) & 1u);
        }
      } else {
        const uint64_t lo = op.lo, hi = op.hi;
#pragma unroll
        for (int u = 0; u < U; u++) in[u] = key[u] >= lo && key[u] <= hi;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const bool r = isnull[u] ? on_null : (in[u] != neg);
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
template <int U, bool TEXT, class Load>
__host__ __device__ __forceinline__ void where_eval_rows(const WhereProgram &prog,
                                                         const uint64_t *args,
                                                         const uint32_t *dict_bits, Load load,
                                                         bool (&out)[U]) {
  if (prog.n_ops < 32) where_eval_rows_t<U, uint32_t, TEXT>(prog, args, dict_bits, load, out);
  else where_eval_rows_t<U, uint64_t, TEXT>(prog, args, dict_bits, load, out);
}

// one row: `load(slot, key, isnull)`
template <class Load>
__host__ __device__ __forceinline__ bool where_eval_row(const WhereProgram &prog,
                                                        const uint64_t *args,
                                                        const uint32_t *dict_bits, Load load) {
  bool out[1];
  where_eval_rows<1, true>(prog, args, dict_bits,
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

// One step of a warp: rows [g * 128, g * 128 + 128). FULL: all of them exist — one base address
// per column with immediate offsets, the four NULL words as one 16-byte load, the four result
// words as one 16-byte store. Otherwise (the last step of the column) rows past the end re-read
// the last row and are masked out of the result.
template <bool TEXT, bool FULL>
__device__ __forceinline__ unsigned where_step(const WhereProgram &prog, const WhereCols &cols,
                                               const uint64_t *__restrict__ args,
                                               const uint32_t *__restrict__ dict_bits,
                                               uint64_t n_rows, uint64_t n_words, uint64_t g, int lane,
                                               uint32_t *__restrict__ out_bits) {
  const uint64_t row0 = g * kWhereWords * 32 + lane;
  uint64_t rowc[kWhereWords];
#pragma unroll
  for (int u = 0; u < kWhereWords; u++) {
    const uint64_t row = row0 + (uint64_t)u * 32;
    rowc[u] = FULL || row < n_rows ? row : n_rows - 1;
  }
  bool res[kWhereWords];
  where_eval_rows<kWhereWords, TEXT>(
      prog, args, dict_bits,
      [&](uint32_t c, uint64_t (&key)[kWhereWords], bool (&isnull)[kWhereWords]) {
        // all loads first, no branch between them, then the conversions
        uint64_t raw[kWhereWords];
        uint32_t nb[kWhereWords];
        const uint64_t *vals = cols.values[c];
        const uint32_t *nulls = cols.nulls[c];
        if (FULL) {
          const uint64_t *vp = vals + row0;
#pragma unroll
          for (int u = 0; u < kWhereWords; u++) raw[u] = __ldg(vp + u * 32);
          static_assert(kWhereWords == 4, "the NULL words of a step are one uint4");
          const uint4 nw = __ldg(reinterpret_cast<const uint4 *>(nulls) + g);
          nb[0] = nw.x >> lane;
          nb[1] = nw.y >> lane;
          nb[2] = nw.z >> lane;
          nb[3] = nw.w >> lane;
        } else {
#pragma unroll
          for (int u = 0; u < kWhereWords; u++) {
            raw[u] = __ldg(vals + rowc[u]);
            nb[u] = __ldg(nulls + (rowc[u] >> 5)) >> (rowc[u] & 31);
          }
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
  unsigned bits[kWhereWords];
#pragma unroll
  for (int u = 0; u < kWhereWords; u++)
    bits[u] = __ballot_sync(0xFFFFFFFFu, res[u] && (FULL || row0 + (uint64_t)u * 32 < n_rows));
  unsigned count = 0;
  if (lane == 0) {
    if (FULL) {
      reinterpret_cast<uint4 *>(out_bits)[g] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
#pragma unroll
      for (int u = 0; u < kWhereWords; u++) count += __popc(bits[u]);
    } else {
#pragma unroll
      for (int u = 0; u < kWhereWords; u++) {
        const uint64_t w = g * kWhereWords + u;
        if (w < n_words) {
          out_bits[w] = bits[u];
          count += __popc(bits[u]);
        }
      }
    }
  }
  return count;
}

template <bool TEXT>
__global__ void __launch_bounds__(256)
where_eval_kernel(const __grid_constant__ WhereProgram prog, const __grid_constant__ WhereCols cols,
                  const uint64_t *__restrict__ args, const uint32_t *__restrict__ dict_bits,
                  uint64_t n_rows, uint32_t n_slots, uint32_t *__restrict__ out_bits,
                  unsigned long long *__restrict__ matched) {
  (void)n_slots;
  const int lane = threadIdx.x & 31;
  const uint64_t n_words = (n_rows + 31) / 32;
  const uint64_t n_groups = (n_words + kWhereWords - 1) / kWhereWords;
  const uint64_t full_groups = n_rows / (kWhereWords * 32);
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long local = 0;
  for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups;
       g += warps) {
    if (g < full_groups)
      local += where_step<TEXT, true>(prog, cols, args, dict_bits, n_rows, n_words, g, lane, out_bits);
    else
      local += where_step<TEXT, false>(prog, cols, args, dict_bits, n_rows, n_words, g, lane, out_bits);
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

// ---- text leaves ---------------------------------------------------------------------------
// One text leaf as the dictionary pass sees it. Operands live in the program's operand pool
// (UTF-16 code units); `list` holds (offset, length) pairs of an IN list.
struct TextLeaf {
  uint8_t op;              // kOpEq / kOpGt / kOpGe / kOpLt / kOpLe / kOpBetween / kOpIn / kOpLike
  uint8_t pad[3];
  uint32_t a_off, a_len;   // operand (BETWEEN: start; LIKE: pattern)
  uint32_t b_off, b_len;   // BETWEEN: end
  uint32_t list_off, n;    // IN: first pair in `list`, pairs
  uint32_t bits_off;       // first word of this leaf's bitmap over codes
};

// Dart String.compareTo: lexicographic over UTF-16 code units, shorter prefix first
__host__ __device__ __forceinline__ int text_compare(const uint16_t *a, uint32_t na,
                                                     const uint16_t *b, uint32_t nb) {
  const uint32_t n = na < nb ? na : nb;
  for (uint32_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na == nb ? 0 : (na < nb ? -1 : 1);
}

// what `.` of a Dart RegExp (no dotAll, no unicode flag) does NOT match
__host__ __device__ __forceinline__ bool text_line_terminator(uint16_t c) {
  return c == 0x000A || c == 0x000D || c == 0x2028 || c == 0x2029;
}

// ValueMatcher.matchesLike (value_matcher.dart:318-331): the pattern becomes the regular
// expression ^...$ with `%` -> `.*`, `_` -> `.` and every other character literal (there is no
// escape: a backslash is a literal backslash), matched against the whole string, case-sensitive.
// `.` is one UTF-16 code unit other than a line terminator, so neither wildcard crosses a line
// break. Iterative matcher with one backtrack point (the latest `%`): a `%` that would have to
// absorb a line terminator fails the match — no earlier `%` could absorb it either.
__host__ __device__ __forceinline__ bool text_like(const uint16_t *s, uint32_t n,
                                                   const uint16_t *pat, uint32_t m) {
  uint32_t si = 0, pi = 0, star_p = 0xFFFFFFFFu, star_s = 0;
  while (si < n) {
    if (pi < m && pat[pi] == (uint16_t)'%') {
      star_p = pi++;
      star_s = si;
    } else if (pi < m && (pat[pi] == (uint16_t)'_' ? !text_line_terminator(s[si])
                                                    : pat[pi] == s[si])) {
      si++;
      pi++;
    } else if (star_p != 0xFFFFFFFFu) {
      if (text_line_terminator(s[star_s])) return false;
      si = ++star_s;
      pi = star_p + 1;
    } else {
      return false;
    }
  }
  while (pi < m && pat[pi] == (uint16_t)'%') pi++;
  return pi == m;
}

// the positive predicate of a text leaf for one string ('!=', NOT IN and NOT LIKE negate it in
// the row pass, where NULL handling lives)
__host__ __device__ __forceinline__ bool text_leaf_match(const TextLeaf &lf, const uint16_t *s,
                                                         uint32_t n, const uint16_t *pool,
                                                         const uint2 *list) {
  const uint16_t *a = pool + lf.a_off;
  switch (lf.op) {
    case kOpEq: return text_compare(s, n, a, lf.a_len) == 0;
    case kOpGt: return text_compare(s, n, a, lf.a_len) > 0;
    case kOpGe: return text_compare(s, n, a, lf.a_len) >= 0;
    case kOpLt: return text_compare(s, n, a, lf.a_len) < 0;
    case kOpLe: return text_compare(s, n, a, lf.a_len) <= 0;
    case kOpBetween:
      return text_compare(s, n, a, lf.a_len) >= 0 &&
             text_compare(s, n, pool + lf.b_off, lf.b_len) <= 0;
    case kOpIn:
      for (uint32_t j = 0; j < lf.n; j++) {
        const uint2 e = list[lf.list_off + j];
        if (text_compare(s, n, pool + e.x, e.y) == 0) return true;
      }
      return false;
    case kOpLike: return text_like(s, n, a, lf.a_len);
    default: return false;
  }
}

// One thread per distinct string of the column; a warp writes one word of the leaf's bitmap.
// HBM traffic: the dictionary's bytes once per text leaf (strings are read front to back by
// their own thread: neighbouring threads read neighbouring strings of the arena).
__global__ void __launch_bounds__(256)
dict_match_kernel(const __grid_constant__ TextLeaf leaf, const uint16_t *__restrict__ units,
                  const uint64_t *__restrict__ offs, uint32_t n_codes,
                  const uint16_t *__restrict__ pool, const uint2 *__restrict__ list,
                  uint32_t *__restrict__ dict_bits) {
  const uint32_t n_words = (n_codes + 31) / 32;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const uint32_t code = w * 32 + lane;
    bool r = false;
    if (code < n_codes) {
      const uint64_t o = offs[code];
      r = text_leaf_match(leaf, units + o, (uint32_t)(offs[code + 1] - o), pool, list);
    }
    const unsigned bits = __ballot_sync(0xFFFFFFFFu, r);
    if (lane == 0) dict_bits[leaf.bits_off + w] = bits;
  }
}

}  // namespace tsc
