// tsc_pk.cu — nodeId -> primary key side table and PK-returning search
// (include/tostore_cuda.h: tsc_index_set_primary_keys / _get_primary_key,
// tsc_vector_search_pk), and the WHERE prefilter from a set of primary keys
// (tsc_index_filter_primary_keys: the role of the `<index>__pk2nid` lookup,
// core/vector_index_manager.dart:1350-1363). Host memory only; the reference keeps this mapping in the
// `<index>__nid2pk` B+Tree (core/vector_index_manager.dart:553-588, :1276-1293).
#include <string.h>

#include "tsc_index.h"

namespace tsc {

int32_t ix_set_primary_keys(Index *ix, uint64_t first_node_id, const uint8_t *utf8,
                            const uint64_t *offsets, uint64_t n) {
  if (n == 0) return TSC_OK;
  if (!offsets || (!utf8 && offsets[n] != offsets[0])) {
    set_error("set_primary_keys: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  const uint64_t base = ix->desc.first_node_id;
  // (mappings follow their rows: an append that grew the column has raised the capacity already)
  if (first_node_id < base || first_node_id - base + n > ix->capacity) {
    set_error("set_primary_keys: node ids [%llu, %llu) outside shard [%llu, %llu)",
              (unsigned long long)first_node_id, (unsigned long long)(first_node_id + n),
              (unsigned long long)base, (unsigned long long)(base + ix->capacity));
    return TSC_ERR_BAD_ARG;
  }
  for (uint64_t i = 0; i < n; i++)
    if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 0xFFFFFFFFull) {
      set_error("set_primary_keys: offsets must be non-decreasing (entry %llu)",
                (unsigned long long)i);
      return TSC_ERR_BAD_ARG;
    }
  const uint64_t row0 = first_node_id - base;
  if (ix->pk_off.size() < row0 + n) {
    ix->pk_off.resize(row0 + n, 0);
    ix->pk_len.resize(row0 + n, 0);
  }
  const uint64_t arena0 = ix->pk_arena.size();
  ix->pk_arena.insert(ix->pk_arena.end(), (const char *)utf8 + offsets[0],
                      (const char *)utf8 + offsets[n]);
  for (uint64_t i = 0; i < n; i++) {
    ix->pk_off[row0 + i] = arena0 + (offsets[i] - offsets[0]);
    ix->pk_len[row0 + i] = (uint32_t)(offsets[i + 1] - offsets[i]);
  }
  ix->pk_rev_valid = false;
  return TSC_OK;
}

// Bitmap over the shard's rows: bit r is set iff the key of row r is one of the n given keys
// (64-bit words, LSB first: the form tsc_index_set_filter takes). The reverse map is rebuilt from
// the nodeId -> key table when that table changed; when a key is mapped by several rows the
// highest node id wins (the latest insert, like an upsert of `__pk2nid`), tombstoned (empty)
// mappings are not searchable. ix->mu held.
static int32_t pk_filter_bitmap_locked(Index *ix, const uint8_t *utf8, const uint64_t *offsets,
                                       uint64_t n, std::vector<uint64_t> *words,
                                       uint64_t *out_matched) {
  if (n && (!offsets || (!utf8 && offsets[n] != offsets[0]))) {
    set_error("filter_primary_keys: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  for (uint64_t i = 0; i < n; i++)
    if (offsets[i + 1] < offsets[i]) {
      set_error("filter_primary_keys: offsets must be non-decreasing (entry %llu)",
                (unsigned long long)i);
      return TSC_ERR_BAD_ARG;
    }
  if (!ix->pk_rev_valid) {
    ix->pk_rev.clear();
    ix->pk_rev.reserve(ix->pk_len.size());
    for (uint64_t r = 0; r < ix->pk_len.size(); r++)
      if (ix->pk_len[r])
        ix->pk_rev[std::string(ix->pk_arena.data() + ix->pk_off[r], ix->pk_len[r])] = r;
    ix->pk_rev_valid = true;
  }
  const uint64_t rows = ix->rows > ix->pk_len.size() ? ix->rows : ix->pk_len.size();
  words->assign((rows + 63) / 64 + 1, 0);
  uint64_t matched = 0;
  for (uint64_t i = 0; i < n; i++) {
    if (offsets[i + 1] == offsets[i]) continue;   // the empty key is the tombstone mapping
    auto it = ix->pk_rev.find(std::string((const char *)utf8 + offsets[i], offsets[i + 1] - offsets[i]));
    if (it == ix->pk_rev.end()) continue;
    uint64_t &w = (*words)[it->second >> 6];
    const uint64_t bit = 1ull << (it->second & 63);
    matched += !(w & bit);
    w |= bit;
  }
  if (out_matched) *out_matched = matched;
  return TSC_OK;
}

int32_t ix_filter_primary_keys(Index *ix, const uint8_t *utf8, const uint64_t *offsets, uint64_t n,
                               uint64_t *out_matched) {
  std::vector<uint64_t> words;
  {
    std::lock_guard<std::mutex> lk(ix->mu);
    int32_t rc = pk_filter_bitmap_locked(ix, utf8, offsets, n, &words, out_matched);
    if (rc != TSC_OK) return rc;
  }
  return ix_set_filter(ix, words.data(), words.size());   // takes the lock itself
}

int32_t ix_get_primary_key(Index *ix, uint64_t node_id, uint8_t *out_utf8, uint32_t capacity,
                           uint32_t *out_len) {
  if (!out_len || (!out_utf8 && capacity)) {
    set_error("get_primary_key: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  *out_len = 0;
  const uint64_t base = ix->desc.first_node_id;
  if (node_id < base || node_id - base >= ix->pk_len.size()) return TSC_OK;
  const uint64_t r = node_id - base;
  *out_len = ix->pk_len[r];
  if (ix->pk_len[r] > capacity) {
    set_error("get_primary_key: key of %u bytes does not fit %u", ix->pk_len[r], capacity);
    return TSC_ERR_BAD_ARG;
  }
  memcpy(out_utf8, ix->pk_arena.data() + ix->pk_off[r], ix->pk_len[r]);
  return TSC_OK;
}

// Result assembly after the engine call (vector_index_manager.dart:576-587): drop hits
// whose node has no primary-key mapping, keep ascending distance order, compact in place.
// `get` reads the key of one node id (one shard's table, or the group's routing).
template <typename GetKey>
static int32_t pk_assemble(GetKey get, uint32_t k, int64_t *out_ids, double *out_dist,
                           double *out_score, uint8_t *out_pk_utf8, uint64_t pk_capacity,
                           uint64_t *out_pk_offsets, uint32_t *out_count) {
  uint32_t kept = 0;
  uint64_t used = 0;
  out_pk_offsets[0] = 0;
  for (uint32_t j = 0; j < *out_count && j < k; j++) {
    if (out_ids[j] < 0) continue;
    uint32_t len = 0;
    const uint64_t room = pk_capacity - used;
    int32_t rc = get((uint64_t)out_ids[j], out_pk_utf8 ? out_pk_utf8 + used : nullptr,
                     room > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)room, &len);
    if (rc != TSC_OK) {
      set_error("vector_search_pk: keys need more than %llu bytes",
                (unsigned long long)pk_capacity);
      return TSC_ERR_BAD_ARG;
    }
    if (len == 0) continue;   // `if (pk == null) continue;` vector_index_manager.dart:578-579
    used += len;
    out_ids[kept] = out_ids[j];
    out_dist[kept] = out_dist[j];
    out_score[kept] = out_score[j];
    out_pk_offsets[++kept] = used;
  }
  for (uint32_t j = kept; j < k; j++) {
    out_ids[j] = -1;
    out_dist[j] = out_score[j] = __builtin_nan("");
    out_pk_offsets[j + 1] = used;
  }
  *out_count = kept;
  return TSC_OK;
}

}  // namespace tsc

using namespace tsc;

extern "C" {

int32_t tsc_index_set_primary_keys(uint64_t handle, uint64_t first_node_id, const uint8_t *utf8,
                                   const uint64_t *offsets, uint64_t n) {
  TSC_API_TRY
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> glk(g->mu);
    return grp_set_primary_keys(*g, first_node_id, utf8, offsets, n);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  return ix_set_primary_keys(ref.get(), first_node_id, utf8, offsets, n);
  TSC_API_CATCH
}

int32_t tsc_index_get_primary_key(uint64_t handle, uint64_t node_id, uint8_t *out_utf8,
                                  uint32_t capacity, uint32_t *out_len) {
  TSC_API_TRY
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> glk(g->mu);
    return grp_get_primary_key(*g, node_id, out_utf8, capacity, out_len);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  return ix_get_primary_key(ref.get(), node_id, out_utf8, capacity, out_len);
  TSC_API_CATCH
}

int32_t tsc_index_filter_primary_keys(uint64_t handle, const uint8_t *utf8, const uint64_t *offsets,
                                      uint64_t n, uint64_t *out_matched) {
  TSC_API_TRY
  if (GroupRef g = lookup_group(handle)) {
    std::lock_guard<std::mutex> glk(g->mu);
    return grp_filter_primary_keys(*g, utf8, offsets, n, out_matched);
  }
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  if (ref->host_only) {
    set_error("host-only self-test handle: no device entry point works on it");
    return TSC_ERR_UNSUPPORTED;
  }
  return ix_filter_primary_keys(ref.get(), utf8, offsets, n, out_matched);
  TSC_API_CATCH
}

// Self-test hook (no GPU): the bitmap tsc_index_filter_primary_keys would install, computed by
// the same code over the handle's primary-key table (works on a host-only index).
int32_t tsc_selftest_pk_filter_bitmap(uint64_t handle, const uint8_t *utf8, const uint64_t *offsets,
                                      uint64_t n, uint64_t *out_words, uint64_t n_words,
                                      uint64_t *out_matched) {
  TSC_API_TRY
  IndexRef ref = lookup_index(handle);
  if (!ref) return TSC_ERR_BAD_HANDLE;
  if (!out_words && n_words) {
    set_error("selftest_pk_filter_bitmap: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  std::vector<uint64_t> words;
  std::lock_guard<std::mutex> lk(ref->mu);
  int32_t rc = pk_filter_bitmap_locked(ref.get(), utf8, offsets, n, &words, out_matched);
  if (rc != TSC_OK) return rc;
  for (uint64_t i = 0; i < n_words; i++) out_words[i] = i < words.size() ? words[i] : 0;
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_vector_search_pk(uint64_t handle, const double *values, uint64_t len, uint32_t k,
                             double threshold, int64_t *out_ids, double *out_dist,
                             double *out_score, uint8_t *out_pk_utf8, uint64_t pk_capacity,
                             uint64_t *out_pk_offsets, uint32_t *out_count) {
  TSC_API_TRY
  if (!lookup_group(handle) && !lookup_index(handle)) return TSC_ERR_BAD_HANDLE;
  if (!out_pk_offsets || (!out_pk_utf8 && pk_capacity)) {
    set_error("vector_search_pk: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  int32_t rc = tsc_vector_search(handle, values, len, k, threshold, out_ids, out_dist, out_score,
                                 out_count);
  if (rc != TSC_OK) return rc;
  auto get = [&](uint64_t node, uint8_t *dst, uint32_t cap, uint32_t *out_len) {
    return tsc_index_get_primary_key(handle, node, dst, cap, out_len);
  };
  return pk_assemble(get, k, out_ids, out_dist, out_score, out_pk_utf8, pk_capacity, out_pk_offsets,
                     out_count);
  TSC_API_CATCH
}

// Self-test hooks (no GPU): a host-only index object that carries nothing but the
// primary-key table, and the result-assembly step of tsc_vector_search_pk applied to
// caller-supplied hits. Every compute entry point fails on such a handle (no device
// memory behind it); release it with tsc_index_destroy.
int32_t tsc_selftest_host_index(uint64_t capacity_rows, uint64_t first_node_id,
                                uint64_t *out_handle) {
  TSC_API_TRY
  if (!out_handle || capacity_rows == 0) {
    set_error("selftest_host_index: bad argument");
    return TSC_ERR_BAD_ARG;
  }
  Index *ix = new Index();
  ix->desc.struct_size = sizeof(tsc_index_desc);
  ix->desc.first_node_id = first_node_id;
  ix->desc.capacity_rows = capacity_rows;
  ix->capacity = capacity_rows;
  ix->host_only = true;
  *out_handle = register_index(IndexRef(ix, free_index));
  return TSC_OK;
  TSC_API_CATCH
}

int32_t tsc_selftest_pk_assemble(uint64_t handle, uint32_t k, int64_t *ids, double *dist,
                                 double *score, uint8_t *out_pk_utf8, uint64_t pk_capacity,
                                 uint64_t *out_pk_offsets, uint32_t *inout_count) {
  TSC_API_TRY
  if (!lookup_group(handle) && !lookup_index(handle)) return TSC_ERR_BAD_HANDLE;
  if (!ids || !dist || !score || !out_pk_offsets || !inout_count) {
    set_error("selftest_pk_assemble: NULL buffer");
    return TSC_ERR_BAD_ARG;
  }
  auto get = [&](uint64_t node, uint8_t *dst, uint32_t cap, uint32_t *out_len) {
    return tsc_index_get_primary_key(handle, node, dst, cap, out_len);
  };
  return pk_assemble(get, k, ids, dist, score, out_pk_utf8, pk_capacity, out_pk_offsets,
                     inout_count);
  TSC_API_CATCH
}

}  // extern "C"
