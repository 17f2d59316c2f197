// tsc_where.cu — attribute columns + the structured WHERE prefilter entry points
// (include/tostore_cuda.h: tsc_index_column_create / _append, tsc_index_filter_where).
// Kernel and semantics: tsc_where.cuh.
#include <string.h>

#include "tsc_index.h"
#include "tsc_where.cuh"

using namespace tsc;

static AttrColumn *find_column(Index *ix, uint32_t id) {
  for (auto &c : ix->columns)
    if (c.id == id) return &c;
  return nullptr;
}

struct WhereColInfo {
  uint32_t id;
  uint8_t type;
};

// Validate a caller's postfix program and translate it into the device form: operands
// become order-preserving keys of the leaf's column type, columns become slots.
// Shared by tsc_index_filter_where and the host self-test.
static int32_t where_build(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                           uint32_t n_in_args, const WhereColInfo *cinfo, uint32_t n_cinfo,
                           WhereProgram *prog, uint32_t *slot_ids, uint32_t *n_slots_out,
                           std::vector<uint64_t> *keys_out) {
  memset(prog, 0, sizeof *prog);
  uint32_t n_slots = 0;
  std::vector<uint64_t> &keys = *keys_out;
  keys.assign(n_in_args, 0);
  std::vector<uint8_t> keyed(n_in_args, 0);
  int depth = 0;   // stack depth check: the program must leave exactly one value
  for (uint32_t i = 0; i < n_ops; i++) {
    const tsc_where_op &o = ops[i];
    WhereDevOp &d = prog->ops[i];
    d.kind = o.kind;
    d.op = o.op;
    d.n = o.n;
    if (o.kind == TSC_W_AND || o.kind == TSC_W_OR) {
      if ((int)o.n > depth || o.n > 63) {
        set_error("filter_where: step %u pops %u values, stack holds %d", i, o.n, depth);
        return TSC_ERR_BAD_ARG;
      }
      depth -= (int)o.n - 1;
      continue;
    }
    if (o.kind != TSC_W_LEAF || o.op >= kOpCount) {
      set_error("filter_where: step %u has unknown kind %u / op %u", i, o.kind, o.op);
      return TSC_ERR_BAD_ARG;
    }
    if (++depth > kWhereMaxOps) {
      set_error("filter_where: stack deeper than %d", kWhereMaxOps);
      return TSC_ERR_BAD_ARG;
    }
    if (o.op == TSC_OP_TRUE || o.op == TSC_OP_FALSE) continue;
    const WhereColInfo *c = nullptr;
    for (uint32_t j = 0; j < n_cinfo; j++)
      if (cinfo[j].id == o.column_id) c = &cinfo[j];
    if (!c) {
      set_error("filter_where: step %u names unknown column %u", i, o.column_id);
      return TSC_ERR_BAD_ARG;
    }
    uint32_t s = 0;
    while (s < n_slots && slot_ids[s] != o.column_id) s++;
    if (s == n_slots) slot_ids[n_slots++] = o.column_id;
    d.col = s;
    const bool f64 = c->type == TSC_COL_F64;
    auto fkey = [](double v) {
      uint64_t b;
      memcpy(&b, &v, 8);
      return where_key_f64_bits(b);
    };
    d.lo = f64 ? fkey(o.f_lo) : where_key_i64(o.i_lo);
    d.hi = f64 ? fkey(o.f_hi) : where_key_i64(o.i_hi);
    if (o.op == TSC_OP_IN || o.op == TSC_OP_NOT_IN) {
      if ((uint64_t)o.args_offset + o.n > n_in_args) {
        set_error("filter_where: step %u IN list [%u, %u) outside in_args (%u)", i, o.args_offset,
                  o.args_offset + o.n, n_in_args);
        return TSC_ERR_BAD_ARG;
      }
      d.args_off = o.args_offset;
      for (uint32_t j = 0; j < o.n; j++) {
        const uint32_t a = o.args_offset + j;
        uint64_t raw;
        memcpy(&raw, (const uint8_t *)in_args + (size_t)a * 8, 8);
        const uint64_t k = f64 ? where_key_f64_bits(raw) : (raw ^ 0x8000000000000000ull);
        if (keyed[a] && keys[a] != k) {
          set_error("filter_where: IN value %u is shared by columns of different types", a);
          return TSC_ERR_BAD_ARG;
        }
        keys[a] = k;
        keyed[a] = 1;
      }
    }
  }
  if (n_ops && depth != 1) {
    set_error("filter_where: program leaves %d values on the stack (must be 1)", depth);
    return TSC_ERR_BAD_ARG;
  }
  prog->n_ops = n_ops;
  *n_slots_out = n_slots;
  return TSC_OK;
}

namespace tsc {

int32_t ix_column_create(Index *ix, uint32_t column_id, uint8_t col_type) {
  if (col_type > TSC_COL_F64) {
    set_error("column_create: unknown column type %u", col_type);
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  if (find_column(ix, column_id)) {
    set_error("column_create: column %u already exists", column_id);
    return TSC_ERR_BAD_ARG;
  }
  if (ix->columns.size() >= (size_t)kWhereMaxCols) {
    set_error("column_create: at most %d columns per index", kWhereMaxCols);
    return TSC_ERR_UNSUPPORTED;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  AttrColumn c;
  c.id = column_id;
  c.type = col_type;
  cudaError_t e = cudaMalloc((void **)&c.d_values, ix->capacity * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&c.d_null, ix->mask_words * 4);
  // every row is NULL until a value is appended for it
  if (e == cudaSuccess) e = cudaMemsetAsync(c.d_null, 0xFF, ix->mask_words * 4, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  if (e != cudaSuccess) {
    set_error("column_create: %s (%.2f GB)", cudaGetErrorString(e), ix->capacity * 8 / 1e9);
    cudaGetLastError();
    cudaFree(c.d_values);
    cudaFree(c.d_null);
    return e == cudaErrorMemoryAllocation ? TSC_ERR_OOM : TSC_ERR_CUDA;
  }
  ix->device_bytes += ix->capacity * 8 + ix->mask_words * 4;
  ix->columns.push_back(c);
  return TSC_OK;
}

int32_t ix_column_append(Index *ix, uint32_t column_id, uint64_t first_node_id, const void *values,
                         const uint8_t *is_null, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!values) {
    set_error("column_append: NULL values");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  AttrColumn *c = find_column(ix, column_id);
  if (!c) {
    set_error("column_append: unknown column %u", column_id);
    return TSC_ERR_BAD_ARG;
  }
  const uint64_t base = ix->desc.first_node_id;
  if (first_node_id < base || first_node_id - base > c->rows) {
    set_error("column_append: first_node_id %llu is not contiguous with column rows [%llu, %llu)",
              (unsigned long long)first_node_id, (unsigned long long)base,
              (unsigned long long)(base + c->rows));
    return TSC_ERR_BAD_ARG;
  }
  const uint64_t row0 = first_node_id - base;
  {
    int32_t grc = ix_ensure_capacity(ix, row0 + n, "column_append");
    if (grc != TSC_OK) return grc;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  TSC_CUDA(cudaMemcpyAsync(c->d_values + row0, values, n * 8, cudaMemcpyHostToDevice, ix->stream));
  // null bits: staged as bytes, packed on the device (or cleared when there are none)
  const uint64_t chunk = 64ull << 20;
  for (uint64_t r = 0; r < n; r += chunk) {
    const uint64_t m = n - r < chunk ? n - r : chunk;
    int32_t rc = ensure_stage_bytes(ix, (size_t)m);
    if (rc != TSC_OK) return rc;
    if (is_null)
      TSC_CUDA(cudaMemcpyAsync(ix->d_stage, is_null + r, m, cudaMemcpyHostToDevice, ix->stream));
    else
      TSC_CUDA(cudaMemsetAsync(ix->d_stage, 0, m, ix->stream));
    const unsigned blocks = (unsigned)((m + 255) / 256 < (uint64_t)ix->sm_count * 8
                                           ? (m + 255) / 256 : ix->sm_count * 8);
    pack_null_bits_kernel<<<blocks, 256, 0, ix->stream>>>(ix->d_stage, m, row0 + r, c->d_null);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    TSC_CUDA(cudaStreamSynchronize(ix->stream));   // staging buffer is reused
  }
  if (row0 + n > c->rows) c->rows = row0 + n;
  return TSC_OK;
}

int32_t ix_filter_where(Index *ix, const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                        uint32_t n_in_args, uint64_t *out_matched) {
  if ((n_ops && !ops) || n_ops > (uint32_t)kWhereMaxOps || (n_in_args && !in_args) ||
      n_in_args > 4096) {
    set_error("filter_where: bad program (n_ops=%u <= %d, n_in_args=%u <= 4096)", n_ops,
              kWhereMaxOps, n_in_args);
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  WhereProgram prog;
  WhereCols cols;
  memset(&cols, 0, sizeof cols);
  std::vector<WhereColInfo> info;
  for (auto &c : ix->columns) info.push_back({c.id, c.type});
  uint32_t slot_ids[kWhereMaxCols];
  uint32_t n_slots = 0;
  std::vector<uint64_t> keys;
  int32_t brc = where_build(ops, n_ops, in_args, n_in_args, info.data(), (uint32_t)info.size(), &prog,
                            slot_ids, &n_slots, &keys);
  if (brc != TSC_OK) return brc;
  for (uint32_t s = 0; s < n_slots; s++) {
    AttrColumn *c = find_column(ix, slot_ids[s]);
    cols.values[s] = c->d_values;
    cols.nulls[s] = c->d_null;
    cols.is_f64[s] = c->type == TSC_COL_F64;
  }
  TSC_CUDA(cudaSetDevice(ix->device));
  cudaStream_t st = ix->stream;
  if (n_in_args > ix->where_args_cap) {
    cudaFree(ix->d_where_args);
    ix->d_where_args = nullptr;
    ix->where_args_cap = 0;
    TSC_CUDA(cudaMalloc((void **)&ix->d_where_args, (size_t)4096 * 8));
    ix->where_args_cap = 4096;
  }
  if (n_in_args)
    TSC_CUDA(cudaMemcpyAsync(ix->d_where_args, keys.data(), (size_t)n_in_args * 8,
                             cudaMemcpyHostToDevice, st));
  unsigned long long matched = 0;
  if (ix->rows) {
    TSC_CUDA(cudaMemsetAsync(ix->d_live_count, 0, 8, st));
    const uint64_t words = (ix->rows + 31) / 32;
    const uint64_t want = (words + 8 * kWhereWords - 1) / (8 * kWhereWords);   // 8 warps per CTA
    const unsigned blocks = (unsigned)(want < (uint64_t)ix->sm_count * 8 ? want : ix->sm_count * 8);
    where_eval_kernel<<<blocks, 256, 0, st>>>(prog, cols, ix->d_where_args, ix->rows, n_slots,
                                              ix->d_filter, ix->d_live_count);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
    TSC_CUDA(cudaMemcpyAsync(&matched, ix->d_live_count, 8, cudaMemcpyDeviceToHost, st));
  }
  TSC_CUDA(cudaStreamSynchronize(st));
  ix->has_filter = true;
  ix->live_dirty = true;
  if (out_matched) *out_matched = matched;
  return TSC_OK;
}

}  // namespace tsc

extern "C" {

#define TSC_WHERE_DISPATCH(handle, grp_call, ix_call)                                   \
  TSC_API_TRY                                                                           \
  if (GroupRef g = lookup_group(handle)) {                                              \
    std::lock_guard<std::mutex> glk(g->mu);                                             \
    return grp_call;                                                                    \
  }                                                                                     \
  IndexRef ref = lookup_index(handle);                                                  \
  if (!ref) return TSC_ERR_BAD_HANDLE;                                                  \
  Index *ix = ref.get();                                                                \
  if (ix->host_only) {                                                                  \
    set_error("host-only self-test handle: no device entry point works on it");         \
    return TSC_ERR_UNSUPPORTED;                                                         \
  }                                                                                     \
  return ix_call;                                                                       \
  TSC_API_CATCH

int32_t tsc_index_column_create(uint64_t handle, uint32_t column_id, uint8_t col_type) {
  TSC_WHERE_DISPATCH(handle, grp_column_create(*g, column_id, col_type),
                     ix_column_create(ix, column_id, col_type))
}

int32_t tsc_index_column_append(uint64_t handle, uint32_t column_id, uint64_t first_node_id,
                                const void *values, const uint8_t *is_null, uint64_t n) {
  TSC_WHERE_DISPATCH(handle, grp_column_append(*g, column_id, first_node_id, values, is_null, n),
                     ix_column_append(ix, column_id, first_node_id, values, is_null, n))
}

int32_t tsc_index_filter_where(uint64_t handle, const tsc_where_op *ops, uint32_t n_ops,
                               const void *in_args, uint32_t n_in_args, uint64_t *out_matched) {
  TSC_WHERE_DISPATCH(handle, grp_filter_where(*g, ops, n_ops, in_args, n_in_args, out_matched),
                     ix_filter_where(ix, ops, n_ops, in_args, n_in_args, out_matched))
}

// Self-test hook (no GPU): the same program translation (where_build) and the same
// per-row evaluation (where_eval_row) as tsc_index_filter_where, over host arrays.
// col_values [n_cols][n_rows] raw 8-byte values, col_is_null [n_cols][n_rows] bytes.
int32_t tsc_selftest_where(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                           uint32_t n_in_args, uint32_t n_cols, const uint32_t *col_ids,
                           const uint8_t *col_types, const uint64_t *col_values,
                           const uint8_t *col_is_null, uint64_t n_rows, uint8_t *out_match) {
  if ((n_ops && !ops) || n_ops > (uint32_t)kWhereMaxOps || (n_in_args && !in_args) ||
      n_in_args > 4096 || n_cols > (uint32_t)kWhereMaxCols || (n_rows && !out_match)) {
    set_error("selftest_where: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  std::vector<WhereColInfo> info;
  for (uint32_t i = 0; i < n_cols; i++) info.push_back({col_ids[i], col_types[i]});
  static WhereProgram prog;   // 2 KB: keep it off small thread stacks
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  uint32_t slot_ids[kWhereMaxCols];
  uint32_t n_slots = 0;
  std::vector<uint64_t> keys;
  int32_t rc = where_build(ops, n_ops, in_args, n_in_args, info.data(), n_cols, &prog, slot_ids,
                           &n_slots, &keys);
  if (rc != TSC_OK) return rc;
  uint32_t src[kWhereMaxCols];   // slot -> index into the caller's column arrays
  for (uint32_t s = 0; s < n_slots; s++)
    for (uint32_t i = 0; i < n_cols; i++)
      if (col_ids[i] == slot_ids[s]) src[s] = i;
  for (uint64_t row = 0; row < n_rows; row++)
    out_match[row] = where_eval_row(prog, keys.data(), [&](uint32_t c, uint64_t &key, bool &isnull) {
      const uint64_t raw = col_values[(size_t)src[c] * n_rows + row];
      key = col_types[src[c]] == TSC_COL_F64 ? where_key_f64_bits(raw)
                                             : (raw ^ 0x8000000000000000ull);
      isnull = col_is_null[(size_t)src[c] * n_rows + row] != 0;
    }) ? 1 : 0;
  return TSC_OK;
}

}  // extern "C"
