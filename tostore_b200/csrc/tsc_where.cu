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
  uint32_t n_codes;   // text columns: distinct strings in the dictionary
};

// what the dictionary pass needs for the text leaves of one program
struct TextPlan {
  std::vector<TextLeaf> leaves;
  std::vector<uint32_t> leaf_slot;   // column slot of each leaf
  std::vector<uint2> list;           // IN lists: (offset, length) into the operand pool
  uint32_t bits_words = 0;           // words of all leaves' bitmaps
};

// Validate a caller's postfix program and translate it into the device form: operands
// become order-preserving keys of the leaf's column type, columns become slots.
// Shared by tsc_index_filter_where and the host self-test.
static int32_t where_build(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                           uint32_t n_in_args, const WhereTexts &texts, const WhereColInfo *cinfo,
                           uint32_t n_cinfo, WhereProgram *prog, uint32_t *slot_ids,
                           uint32_t *n_slots_out, std::vector<uint64_t> *keys_out,
                           TextPlan *plan) {
  memset(prog, 0, sizeof *prog);
  // operand t of the text pool as (offset, length), checked
  auto text_ref = [&](int64_t t, uint32_t step, uint2 *out) -> bool {
    if (t < 0 || (uint64_t)t >= texts.n) {
      set_error("filter_where: step %u names text operand %lld, the program has %u", step,
                (long long)t, texts.n);
      return false;
    }
    const uint64_t a = texts.offsets[t], b = texts.offsets[t + 1];
    if (b < a || b - a > 0xFFFFFFFFull || a > 0xFFFFFFFFull) {
      set_error("filter_where: text operand %lld has a bad range", (long long)t);
      return false;
    }
    *out = make_uint2((uint32_t)a, (uint32_t)(b - a));
    return true;
  };
  uint32_t n_slots = 0;
  std::vector<uint64_t> &keys = *keys_out;
  keys.assign(n_in_args, 0);
  std::vector<uint8_t> keyed(n_in_args, 0);
  int depth = 0;   // stack depth check: the program must leave exactly one value
  for (uint32_t i = 0; i < n_ops; i++) {
    const tsc_where_op &o = ops[i];
    WhereDevOp &d = prog->ops[i];
    d.kind = o.kind;
    d.n = o.n;
    if (o.kind == TSC_W_AND || o.kind == TSC_W_OR) {
      if ((int)o.n > depth || o.n > 63) {
        set_error("filter_where: step %u pops %u values, stack holds %d", i, o.n, depth);
        return TSC_ERR_BAD_ARG;
      }
      depth -= (int)o.n - 1;
      continue;
    }
    if (o.kind != TSC_W_LEAF || o.op >= (uint8_t)kOpCount) {
      set_error("filter_where: step %u has unknown kind %u / op %u", i, o.kind, o.op);
      return TSC_ERR_BAD_ARG;
    }
    if (++depth > kWhereMaxOps) {
      set_error("filter_where: stack deeper than %d", kWhereMaxOps);
      return TSC_ERR_BAD_ARG;
    }
    if (o.op == TSC_OP_TRUE || o.op == TSC_OP_FALSE) {
      where_decode_leaf(o.op, 0, 0, &d);
      continue;
    }
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
    if (c->type == TSC_COL_TEXT) {
      if (o.op == TSC_OP_IS_NULL || o.op == TSC_OP_IS_NOT_NULL) {   // the NULL bitmap answers
        where_decode_leaf(o.op, 0, 0, &d);
        continue;
      }
      // positive predicate over the dictionary + how the row pass reads it
      TextLeaf lf;
      memset(&lf, 0, sizeof lf);
      uint8_t flags = kLfDict;
      switch (o.op) {
        case TSC_OP_NE: lf.op = kOpEq; flags |= kLfNeg | kLfOnNull; break;
        case TSC_OP_NOT_IN: lf.op = kOpIn; flags |= kLfNeg | kLfOnNull; break;
        case TSC_OP_NOT_LIKE: lf.op = kOpLike; flags |= kLfNeg; break;   // false on NULL (:602-604)
        default: lf.op = o.op; break;
      }
      uint2 r;
      if (lf.op == kOpIn) {
        if ((uint64_t)o.args_offset + o.n > n_in_args) {
          set_error("filter_where: step %u IN list [%u, %u) outside in_args (%u)", i,
                    o.args_offset, o.args_offset + o.n, n_in_args);
          return TSC_ERR_BAD_ARG;
        }
        lf.list_off = (uint32_t)plan->list.size();
        lf.n = o.n;
        for (uint32_t j = 0; j < o.n; j++) {
          uint64_t t;
          memcpy(&t, (const uint8_t *)in_args + (size_t)(o.args_offset + j) * 8, 8);
          if (!text_ref(t > 0x7FFFFFFFull ? -1 : (int64_t)t, i, &r)) return TSC_ERR_BAD_ARG;
          plan->list.push_back(r);
        }
      } else {
        if (!text_ref(o.i_lo, i, &r)) return TSC_ERR_BAD_ARG;
        lf.a_off = r.x;
        lf.a_len = r.y;
        if (lf.op == kOpBetween) {
          if (!text_ref(o.i_hi, i, &r)) return TSC_ERR_BAD_ARG;
          lf.b_off = r.x;
          lf.b_len = r.y;
        }
      }
      lf.bits_off = plan->bits_words;
      plan->bits_words += (c->n_codes + 31) / 32 + 1;   // never empty: a NULL row may read word 0
      plan->leaves.push_back(lf);
      plan->leaf_slot.push_back(s);
      d.flags = flags;
      d.args_off = lf.bits_off;
      continue;
    }
    if (o.op == TSC_OP_LIKE || o.op == TSC_OP_NOT_LIKE) {
      set_error("filter_where: step %u: LIKE needs a text column (column %u is numeric)", i,
                o.column_id);
      return TSC_ERR_UNSUPPORTED;
    }
    const bool f64 = c->type == TSC_COL_F64;
    auto fkey = [](double v) {
      uint64_t b;
      memcpy(&b, &v, 8);
      return where_key_f64_bits(b);
    };
    where_decode_leaf(o.op, f64 ? fkey(o.f_lo) : where_key_i64(o.i_lo),
                      f64 ? fkey(o.f_hi) : where_key_i64(o.i_hi), &d);
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
  if (col_type > TSC_COL_TEXT) {
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
  if (e == cudaSuccess && col_type == TSC_COL_TEXT)   // codes of unwritten rows stay in range
    e = cudaMemsetAsync(c.d_values, 0, ix->capacity * 8, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  if (e != cudaSuccess) {
    set_error("column_create: %s (%.2f GB)", cudaGetErrorString(e), ix->capacity * 8 / 1e9);
    cudaGetLastError();
    cudaFree(c.d_values);
    cudaFree(c.d_null);
    return e == cudaErrorMemoryAllocation ? TSC_ERR_OOM : TSC_ERR_CUDA;
  }
  ix->device_bytes += ix->capacity * 8 + ix->mask_words * 4;
  if (col_type == TSC_COL_TEXT) c.dict = std::make_shared<TextDict>();
  ix->columns.push_back(c);
  return TSC_OK;
}

// rows [first_node_id, first_node_id + n) of a column <- n 8-byte values (+ NULL bytes); ix->mu held
static int32_t column_write_locked(Index *ix, AttrColumn *c, uint64_t first_node_id,
                                   const void *values, const uint8_t *is_null, uint64_t n) {
  const uint64_t base = ix->desc.first_node_id;
  if (first_node_id < base || first_node_id - base > c->rows) {
    set_error("column_append: first_node_id %llu is not contiguous with column rows [%llu, %llu)",
              (unsigned long long)first_node_id, (unsigned long long)base,
              (unsigned long long)(base + c->rows));
    return TSC_ERR_BAD_ARG;
  }
  const uint64_t row0 = first_node_id - base;
  {
    const uint32_t cid = c->id;
    int32_t grc = ix_ensure_capacity(ix, row0 + n, "column_append");
    if (grc != TSC_OK) return grc;
    c = find_column(ix, cid);
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
  if (c->type == TSC_COL_TEXT) {
    set_error("column_append: column %u is a text column (tsc_index_column_append_text)", column_id);
    return TSC_ERR_BAD_ARG;
  }
  return column_write_locked(ix, c, first_node_id, values, is_null, n);
}

// device arrays of a dictionary hold at least `units` code units and `codes` strings
static int32_t dict_reserve(Index *ix, TextDict *d, uint64_t units, uint64_t codes) {
  if (units > d->units_cap) {
    uint64_t cap = d->units_cap ? d->units_cap * 2 : 4096;
    while (cap < units) cap *= 2;
    uint16_t *q = nullptr;
    if (cudaMalloc((void **)&q, cap * 2) != cudaSuccess) {
      cudaGetLastError();
      set_error("column_append_text: out of device memory for the dictionary (%.2f GB)", cap * 2 / 1e9);
      return TSC_ERR_OOM;
    }
    if (d->n_units)
      TSC_CUDA(cudaMemcpyAsync(q, d->d_units, d->n_units * 2, cudaMemcpyDeviceToDevice, ix->stream));
    TSC_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(d->d_units);
    ix->device_bytes += (cap - d->units_cap) * 2;
    d->d_units = q;
    d->units_cap = cap;
  }
  if (codes > d->codes_cap) {
    uint64_t cap = d->codes_cap ? d->codes_cap * 2 : 1024;
    while (cap < codes) cap *= 2;
    uint64_t *q = nullptr;
    if (cudaMalloc((void **)&q, (cap + 1) * 8) != cudaSuccess) {
      cudaGetLastError();
      set_error("column_append_text: out of device memory for the dictionary (%.2f GB)", cap * 8 / 1e9);
      return TSC_ERR_OOM;
    }
    if (d->d_offs)
      TSC_CUDA(cudaMemcpyAsync(q, d->d_offs, ((uint64_t)d->n_codes + 1) * 8, cudaMemcpyDeviceToDevice,
                               ix->stream));
    else
      TSC_CUDA(cudaMemsetAsync(q, 0, 8, ix->stream));   // offs[0] = 0
    TSC_CUDA(cudaStreamSynchronize(ix->stream));
    cudaFree(d->d_offs);
    ix->device_bytes += (cap - d->codes_cap) * 8;
    d->d_offs = q;
    d->codes_cap = cap;
  }
  return TSC_OK;
}

// Intern the strings of rows [0, n): codes[i] <- the row's code; strings seen for the first
// time extend the dictionary's host mirror. Shared by the device path and the host self-test.
// A NULL row gets code 0 and stores nothing.
static int32_t dict_intern(TextDict *d, const uint16_t *units, const uint64_t *offsets,
                           const uint8_t *is_null, uint64_t n, uint64_t *codes) {
  for (uint64_t i = 0; i < n; i++) {
    codes[i] = 0;
    if (is_null && is_null[i]) continue;
    const uint64_t a = offsets[i], b = offsets[i + 1];
    if (b < a || b - a > 0x7FFFFFFFull) {
      set_error("column_append_text: row %llu has a bad string range", (unsigned long long)i);
      return TSC_ERR_BAD_ARG;
    }
    if (d->host_codes() == 0xFFFFFFFEu) {
      set_error("column_append_text: more than 2^32 - 2 distinct strings");
      return TSC_ERR_UNSUPPORTED;
    }
    codes[i] = d->intern(units + a, (size_t)(b - a));
  }
  return TSC_OK;
}

int32_t ix_column_append_text(Index *ix, uint32_t column_id, uint64_t first_node_id,
                              const uint16_t *units, const uint64_t *offsets,
                              const uint8_t *is_null, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!offsets || (!units && offsets[n] != offsets[0])) {
    set_error("column_append_text: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  AttrColumn *c = find_column(ix, column_id);
  if (!c || c->type != TSC_COL_TEXT) {
    set_error("column_append_text: column %u is not a text column of this index", column_id);
    return TSC_ERR_BAD_ARG;
  }
  TextDict *d = c->dict.get();
  std::vector<uint64_t> codes(n);
  int32_t rc = dict_intern(d, units, offsets, is_null, n, codes.data());
  if (rc != TSC_OK) {
    d->truncate(d->n_codes);   // forget the strings interned by this call
    return rc;
  }
  if (d->host_codes() > d->n_codes) {   // upload the new tail of the dictionary
    TSC_CUDA(cudaSetDevice(ix->device));
    rc = dict_reserve(ix, d, d->h_units.size(), d->host_codes());
    if (rc != TSC_OK) {
      d->truncate(d->n_codes);
      return rc;
    }
    cudaError_t e = cudaSuccess;
    if (d->h_units.size() > d->n_units)
      e = cudaMemcpyAsync(d->d_units + d->n_units, d->h_units.data() + d->n_units,
                          (d->h_units.size() - d->n_units) * 2, cudaMemcpyHostToDevice, ix->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d->d_offs + d->n_codes + 1, d->h_offs.data() + d->n_codes + 1,
                          (size_t)(d->host_codes() - d->n_codes) * 8, cudaMemcpyHostToDevice,
                          ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) {
      d->truncate(d->n_codes);
      set_error("column_append_text: dictionary upload failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TSC_ERR_CUDA;
    }
    d->n_units = d->h_units.size();
    d->n_codes = d->host_codes();
  }
  return column_write_locked(ix, c, first_node_id, codes.data(), is_null, n);
}

int32_t ix_filter_where(Index *ix, const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                        uint32_t n_in_args, const WhereTexts &texts, uint64_t *out_matched) {
  if ((n_ops && !ops) || n_ops > (uint32_t)kWhereMaxOps || (n_in_args && !in_args) ||
      n_in_args > 4096 || texts.n > 4096 || (texts.n && (!texts.offsets || !texts.units))) {
    set_error("filter_where: bad program (n_ops=%u <= %d, n_in_args=%u <= 4096, n_texts=%u <= 4096)",
              n_ops, kWhereMaxOps, n_in_args, texts.n);
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  WhereProgram prog;
  WhereCols cols;
  memset(&cols, 0, sizeof cols);
  std::vector<WhereColInfo> info;
  for (auto &c : ix->columns) info.push_back({c.id, c.type, c.dict ? c.dict->n_codes : 0u});
  uint32_t slot_ids[kWhereMaxCols];
  uint32_t n_slots = 0;
  std::vector<uint64_t> keys;
  TextPlan plan;
  int32_t brc = where_build(ops, n_ops, in_args, n_in_args, texts, info.data(), (uint32_t)info.size(),
                            &prog, slot_ids, &n_slots, &keys, &plan);
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
  // text leaves: every distinct string of the leaf's column is tested once, into a bitmap over codes
  const uint32_t *d_dict_bits = nullptr;
  if (!plan.leaves.empty()) {
    const uint64_t pool_units = texts.n ? texts.offsets[texts.n] : 0;
    const size_t pool_bytes = ((size_t)pool_units * 2 + 15) & ~(size_t)15;
    const size_t list_bytes = (plan.list.size() * sizeof(uint2) + 15) & ~(size_t)15;
    const size_t need = pool_bytes + list_bytes + (size_t)plan.bits_words * 4;
    if (need > ix->where_text_cap) {
      cudaFree(ix->d_where_text);
      ix->d_where_text = nullptr;
      ix->where_text_cap = 0;
      const size_t cap = need + (need >> 2) + 4096;
      if (cudaMalloc((void **)&ix->d_where_text, cap) != cudaSuccess) {
        cudaGetLastError();
        set_error("filter_where: out of device memory for the text leaves (%.2f GB)", cap / 1e9);
        return TSC_ERR_OOM;
      }
      ix->where_text_cap = cap;
    }
    uint16_t *d_pool = reinterpret_cast<uint16_t *>(ix->d_where_text);
    uint2 *d_list = reinterpret_cast<uint2 *>(ix->d_where_text + pool_bytes);
    uint32_t *d_bits = reinterpret_cast<uint32_t *>(ix->d_where_text + pool_bytes + list_bytes);
    if (pool_units)
      TSC_CUDA(cudaMemcpyAsync(d_pool, texts.units, (size_t)pool_units * 2, cudaMemcpyHostToDevice, st));
    if (!plan.list.empty())
      TSC_CUDA(cudaMemcpyAsync(d_list, plan.list.data(), plan.list.size() * sizeof(uint2),
                               cudaMemcpyHostToDevice, st));
    TSC_CUDA(cudaMemsetAsync(d_bits, 0, (size_t)plan.bits_words * 4, st));
    for (size_t l = 0; l < plan.leaves.size(); l++) {
      const TextDict *d = find_column(ix, slot_ids[plan.leaf_slot[l]])->dict.get();
      if (d->n_codes == 0) continue;
      const uint32_t words = (d->n_codes + 31) / 32;
      const unsigned want = (words + 7) / 8;   // 8 warps per CTA, one word per warp and step
      const unsigned blocks = want < (unsigned)ix->sm_count * 8 ? want : (unsigned)ix->sm_count * 8;
      dict_match_kernel<<<blocks, 256, 0, st>>>(plan.leaves[l], d->d_units, d->d_offs, d->n_codes,
                                                d_pool, d_list, d_bits);
      TSC_CUDA(cudaGetLastError());
      ix->launches++;
    }
    d_dict_bits = d_bits;
  }
  unsigned long long matched = 0;
  if (ix->rows) {
    TSC_CUDA(cudaMemsetAsync(ix->d_live_count, 0, 8, st));
    const uint64_t words = (ix->rows + 31) / 32;
    const uint64_t want = (words + 8 * kWhereWords - 1) / (8 * kWhereWords);   // 8 warps per CTA
    const unsigned blocks = (unsigned)(want < (uint64_t)ix->sm_count * 8 ? want : ix->sm_count * 8);
    if (d_dict_bits)
      where_eval_kernel<true><<<blocks, 256, 0, st>>>(prog, cols, ix->d_where_args, d_dict_bits,
                                                      ix->rows, n_slots, ix->d_filter,
                                                      ix->d_live_count);
    else
      where_eval_kernel<false><<<blocks, 256, 0, st>>>(prog, cols, ix->d_where_args, nullptr,
                                                       ix->rows, n_slots, ix->d_filter,
                                                       ix->d_live_count);
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

int32_t tsc_index_column_append_text(uint64_t handle, uint32_t column_id, uint64_t first_node_id,
                                     const uint16_t *units, const uint64_t *offsets,
                                     const uint8_t *is_null, uint64_t n) {
  TSC_WHERE_DISPATCH(handle,
                     grp_column_append_text(*g, column_id, first_node_id, units, offsets, is_null, n),
                     ix_column_append_text(ix, column_id, first_node_id, units, offsets, is_null, n))
}

int32_t tsc_index_filter_where(uint64_t handle, const tsc_where_op *ops, uint32_t n_ops,
                               const void *in_args, uint32_t n_in_args, uint64_t *out_matched) {
  const WhereTexts none;
  TSC_WHERE_DISPATCH(handle, grp_filter_where(*g, ops, n_ops, in_args, n_in_args, none, out_matched),
                     ix_filter_where(ix, ops, n_ops, in_args, n_in_args, none, out_matched))
}

int32_t tsc_index_filter_where_text(uint64_t handle, const tsc_where_op *ops, uint32_t n_ops,
                                    const void *in_args, uint32_t n_in_args,
                                    const uint16_t *text_units, const uint64_t *text_offsets,
                                    uint32_t n_texts, uint64_t *out_matched) {
  WhereTexts texts;
  texts.units = text_units;
  texts.offsets = text_offsets;
  texts.n = n_texts;
  TSC_WHERE_DISPATCH(handle, grp_filter_where(*g, ops, n_ops, in_args, n_in_args, texts, out_matched),
                     ix_filter_where(ix, ops, n_ops, in_args, n_in_args, texts, out_matched))
}

// Self-test hooks (no GPU): the same program translation (where_build), the same dictionary
// builder (dict_intern), the same per-string test (text_leaf_match) and the same per-row
// evaluation (where_eval_row) as tsc_index_filter_where, over host arrays.
// col_values [n_cols][n_rows] raw 8-byte values, col_is_null [n_cols][n_rows] bytes.
int32_t tsc_selftest_where_text(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                                uint32_t n_in_args, const uint16_t *text_units,
                                const uint64_t *text_offsets, uint32_t n_texts, uint32_t n_cols,
                                const uint32_t *col_ids, const uint8_t *col_types,
                                const uint64_t *col_values, const uint8_t *col_is_null,
                                const uint16_t *row_units, const uint64_t *row_offsets,
                                uint64_t n_rows, uint8_t *out_match) {
  TSC_API_TRY
  if ((n_ops && !ops) || n_ops > (uint32_t)kWhereMaxOps || (n_in_args && !in_args) ||
      n_in_args > 4096 || n_texts > 4096 || (n_texts && (!text_units || !text_offsets)) ||
      n_cols > (uint32_t)kWhereMaxCols || (n_rows && !out_match)) {
    set_error("selftest_where: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  // text columns: intern the rows like tsc_index_column_append_text does
  struct HostDict {
    TextDict d;
    std::vector<uint64_t> codes;
  };
  std::vector<HostDict> dicts(n_cols);
  std::vector<WhereColInfo> info;
  for (uint32_t i = 0; i < n_cols; i++) {
    if (col_types[i] == TSC_COL_TEXT) {
      if (!row_offsets) {
        set_error("selftest_where: text column without row strings");
        return TSC_ERR_BAD_ARG;
      }
      HostDict &h = dicts[i];
      h.codes.resize(n_rows);
      int32_t rc = dict_intern(&h.d, row_units, row_offsets + (size_t)i * (n_rows + 1),
                               col_is_null + (size_t)i * n_rows, n_rows, h.codes.data());
      if (rc != TSC_OK) return rc;
      h.d.n_codes = h.d.host_codes();
    }
    info.push_back({col_ids[i], col_types[i], dicts[i].d.n_codes});
  }
  WhereTexts texts;
  texts.units = text_units;
  texts.offsets = text_offsets;
  texts.n = n_texts;
  static WhereProgram prog;   // 2 KB: keep it off small thread stacks
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  uint32_t slot_ids[kWhereMaxCols];
  uint32_t n_slots = 0;
  std::vector<uint64_t> keys;
  TextPlan plan;
  int32_t rc = where_build(ops, n_ops, in_args, n_in_args, texts, info.data(), n_cols, &prog,
                           slot_ids, &n_slots, &keys, &plan);
  if (rc != TSC_OK) return rc;
  uint32_t src[kWhereMaxCols];   // slot -> index into the caller's column arrays
  for (uint32_t s = 0; s < n_slots; s++)
    for (uint32_t i = 0; i < n_cols; i++)
      if (col_ids[i] == slot_ids[s]) src[s] = i;
  std::vector<uint32_t> bits(plan.bits_words + 1, 0);
  for (size_t l = 0; l < plan.leaves.size(); l++) {
    const HostDict &h = dicts[src[plan.leaf_slot[l]]];
    for (uint32_t code = 0; code < h.d.n_codes; code++)
      if (text_leaf_match(plan.leaves[l], h.d.h_units.data() + h.d.h_offs[code],
                          (uint32_t)(h.d.h_offs[code + 1] - h.d.h_offs[code]), text_units,
                          plan.list.data()))
        bits[plan.leaves[l].bits_off + (code >> 5)] |= 1u << (code & 31);
  }
  for (uint64_t row = 0; row < n_rows; row++)
    out_match[row] = where_eval_row(prog, keys.data(), bits.data(),
                                    [&](uint32_t c, uint64_t &key, bool &isnull) {
      const uint32_t i = src[c];
      const uint64_t raw = col_types[i] == TSC_COL_TEXT ? dicts[i].codes[row]
                                                        : col_values[(size_t)i * n_rows + row];
      key = col_types[i] == TSC_COL_F64 ? where_key_f64_bits(raw) : (raw ^ 0x8000000000000000ull);
      isnull = col_is_null[(size_t)i * n_rows + row] != 0;
    }) ? 1 : 0;
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_selftest_where(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                           uint32_t n_in_args, uint32_t n_cols, const uint32_t *col_ids,
                           const uint8_t *col_types, const uint64_t *col_values,
                           const uint8_t *col_is_null, uint64_t n_rows, uint8_t *out_match) {
  return tsc_selftest_where_text(ops, n_ops, in_args, n_in_args, nullptr, nullptr, 0, n_cols,
                                 col_ids, col_types, col_values, col_is_null, nullptr, nullptr,
                                 n_rows, out_match);
}

}  // extern "C"
