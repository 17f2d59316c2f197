// tostore_vector.hpp — C++ host layer over the C ABI (tostore_cuda.h), mirroring the slice
// of ToStore's Dart API that serves vector search: same type names, argument meaning and
// error behaviour as the reference (paths relative to /root/reference/lib):
//
//   VectorData, VectorFieldConfig, VectorPrecision, VectorDistanceMetric, VectorIndexConfig
//                                                   src/model/table_schema.dart:2109-2675
//   VectorSearchResult                              src/model/query_result.dart:207-228
//   ToStore.vectorSearch(tableName, fieldName:, queryVector:, topK: 10, efSearch:,
//                        distanceThreshold:)        tostore.dart:493-511
//   VectorIndexManager.vectorSearch / writeChanges  src/core/vector_index_manager.dart:297-589
//   QueryCondition (where / or / whereIn / whereBetween ...), operator semantics
//                                                   src/query/query_condition.dart:117-660,
//                                                   src/handler/value_matcher.dart:570-612
//
// The reference's host language is Dart (binding source: dart/tostore_cuda_bindings.dart);
// this header is the same host layer in the compiled language this build image has, and
// tostore_b200/vector_store.py is its Python twin (the one the pytest suite drives).
// Header-only; link with -ltostore_cuda. Like `vectorSearch` in the reference, lookups that
// miss (unknown table / field, empty index) return an empty list instead of throwing;
// library failures throw TscError. There is no CPU fallback.
#ifndef TOSTORE_VECTOR_HPP
#define TOSTORE_VECTOR_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "tostore_cuda.h"

namespace tostore {

enum class VectorPrecision : uint8_t { float64 = 0, float32 = 1, int8 = 2 };        // enum order :2481-2498
enum class VectorDistanceMetric : uint8_t { l2 = 0, innerProduct = 1, cosine = 2 }; // enum order :2511-2531
enum class DeviceDType : uint8_t { float32 = 0, bfloat16 = 1, float16 = 2 };         // new: storage in HBM
enum class DataType : uint8_t { integer = 0, doubleType = 1, text = 2 };             // attribute fields

struct TscError : std::runtime_error {
  int32_t status;
  TscError(int32_t s, const std::string &where)
      : std::runtime_error(where + ": " + tsc_status_name(s) + ": " + tsc_last_error()), status(s) {}
};
inline void check(int32_t rc, const char *where) {
  if (rc != TSC_OK) throw TscError(rc, where);
}

struct VectorData {
  std::vector<double> values;
  static VectorData fromList(std::vector<double> v) { return VectorData{std::move(v)}; }
  int dimensions() const { return (int)values.size(); }
};

struct VectorFieldConfig {
  int dimensions = 0;
  VectorPrecision precision = VectorPrecision::float64;
};

struct VectorIndexConfig {
  VectorDistanceMetric distanceMetric = VectorDistanceMetric::cosine;
  std::optional<int> maxDegree, efSearch, constructionEf, pqSubspaces;   // accepted, unused: the
  std::optional<double> pruneAlpha;                                       // exact scan has no graph
};

struct VectorSearchResult {
  std::string primaryKey;
  double distance = 0.0;
  double score = 0.0;
};

// ---- text values: UTF-8 at this layer, UTF-16 code units (a Dart String's own form) below ----
inline std::u16string utf8ToUtf16(const std::string &s) {
  std::u16string out;
  for (size_t i = 0; i < s.size();) {
    const unsigned char c = (unsigned char)s[i];
    uint32_t cp = 0xFFFD;
    size_t n = 1;
    if (c < 0x80) cp = c;
    else if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; n = 2; }
    else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; n = 3; }
    else if ((c & 0xF8) == 0xF0) { cp = c & 0x07; n = 4; }
    if (i + n > s.size()) { cp = 0xFFFD; n = 1; }
    for (size_t j = 1; j < n; j++) {
      const unsigned char d = (unsigned char)s[i + j];
      if ((d & 0xC0) != 0x80) { cp = 0xFFFD; n = 1; break; }
      cp = (cp << 6) | (d & 0x3F);
    }
    i += n;
    if (cp >= 0x10000) {
      cp -= 0x10000;
      out.push_back((char16_t)(0xD800 + (cp >> 10)));
      out.push_back((char16_t)(0xDC00 + (cp & 0x3FF)));
    } else {
      out.push_back((char16_t)cp);
    }
  }
  return out;
}
// String.trim(): Unicode White_Space plus the byte-order mark
inline std::u16string dartTrim(const std::u16string &s) {
  auto ws = [](char16_t c) {
    return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 ||
           (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F ||
           c == 0x3000 || c == 0xFEFF;
  };
  size_t a = 0, b = s.size();
  while (a < b && ws(s[a])) a++;
  while (b > a && ws(s[b - 1])) b--;
  return s.substr(a, b - a);
}

// ---- QueryCondition: builds the postfix program tsc_index_filter_where[_text] evaluates ----
using Value = std::variant<std::monostate, int64_t, double, std::string>;   // monostate = null; string = UTF-8

// convertValue for DataType.text (table_schema.dart:1421-1442): toString().trim()
inline std::u16string convertText(const Value &v) {
  if (std::holds_alternative<std::string>(v)) return dartTrim(utf8ToUtf16(std::get<std::string>(v)));
  if (std::holds_alternative<int64_t>(v)) return utf8ToUtf16(std::to_string(std::get<int64_t>(v)));
  throw std::invalid_argument("text field operand must be a string or an integer");
}

class QueryCondition {
 public:
  // `where` ANDs onto the current group, `orWhere` starts an alternative
  // (query_condition.dart:117-260); operators as in `_buildCondition` (:478-520).
  QueryCondition &where(const std::string &field, const std::string &op, Value v = {}) {
    groups_.back().push_back(Leaf{field, upper(op), {std::move(v)}});
    return *this;
  }
  QueryCondition &orWhere(const std::string &field, const std::string &op, Value v = {}) {
    groups_.emplace_back();
    return where(field, op, std::move(v));
  }
  QueryCondition &whereIn(const std::string &field, std::vector<Value> vs) {
    groups_.back().push_back(Leaf{field, "IN", std::move(vs)});
    return *this;
  }
  QueryCondition &whereNotIn(const std::string &field, std::vector<Value> vs) {
    groups_.back().push_back(Leaf{field, "NOT IN", std::move(vs)});
    return *this;
  }
  QueryCondition &whereBetween(const std::string &field, Value start, Value end) {
    groups_.back().push_back(Leaf{field, "BETWEEN", {std::move(start), std::move(end)}});
    return *this;
  }
  // convenience forms, one `where` each (query_condition.dart:574-656); text fields only
  QueryCondition &whereLike(const std::string &field, const std::string &pattern) { return where(field, "LIKE", pattern); }
  QueryCondition &whereNotLike(const std::string &field, const std::string &pattern) {
    return where(field, "NOT LIKE", pattern);
  }
  QueryCondition &whereContains(const std::string &field, const std::string &v) { return where(field, "LIKE", "%" + v + "%"); }
  QueryCondition &whereNotContains(const std::string &field, const std::string &v) {
    return where(field, "NOT LIKE", "%" + v + "%");
  }
  QueryCondition &whereStartsWith(const std::string &field, const std::string &v) { return where(field, "LIKE", v + "%"); }
  QueryCondition &whereEndsWith(const std::string &field, const std::string &v) { return where(field, "LIKE", "%" + v); }
  QueryCondition &whereNull(const std::string &field) { return where(field, "IS"); }
  QueryCondition &whereNotNull(const std::string &field) { return where(field, "IS NOT"); }
  bool isEmpty() const {
    for (auto &g : groups_)
      if (!g.empty()) return false;
    return true;
  }

  struct Program {
    std::vector<tsc_where_op> ops;
    std::vector<uint64_t> in_args;   // raw 8-byte values typed like the leaf's column
    std::u16string text_units;       // operand pool of the text leaves ...
    std::vector<uint64_t> text_offsets{0};   // ... operand t = units [offsets[t], offsets[t + 1])
    uint32_t n_texts() const { return (uint32_t)text_offsets.size() - 1; }
    int64_t addText(const std::u16string &s) {
      text_units += s;
      text_offsets.push_back(text_units.size());
      return (int64_t)text_offsets.size() - 2;
    }
  };
  // columns: field name -> (column id, type). Operands are converted to the field's type
  // exactly like FieldSchema.convertValue (table_schema.dart:1371-1421): integer fields
  // round doubles half away from zero, double fields widen integers.
  Program compile(const std::map<std::string, std::pair<uint32_t, DataType>> &columns) const {
    Program p;
    size_t n_groups = 0;
    for (auto &g : groups_) {
      if (g.empty()) continue;
      for (auto &leaf : g) emit(leaf, columns, &p);
      if (g.size() != 1) p.ops.push_back(node(TSC_W_AND, (uint16_t)g.size()));
      n_groups++;
    }
    if (n_groups > 1) p.ops.push_back(node(TSC_W_OR, (uint16_t)n_groups));
    if (p.ops.size() > 64) throw std::invalid_argument("condition compiles to more than 64 steps");
    return p;
  }

 private:
  struct Leaf {
    std::string field, op;
    std::vector<Value> args;
  };
  std::vector<std::vector<Leaf>> groups_{1};

  static std::string upper(std::string s) {
    for (auto &c : s) c = (char)std::toupper((unsigned char)c);
    return s;
  }
  static tsc_where_op node(uint8_t kind, uint16_t n) {
    tsc_where_op o;
    std::memset(&o, 0, sizeof o);
    o.kind = kind;
    o.n = n;
    return o;
  }
  static int64_t dartRound(double x) {   // double.round(): half away from zero, clamped to int64
    if (std::isnan(x) || std::isinf(x))
      throw std::invalid_argument("cannot convert NaN / infinity to an integer field operand");
    const double a = std::fabs(x);
    double r = a;
    if (a < 4503599627370496.0) {
      r = std::floor(a);
      if (a - r >= 0.5) r += 1.0;
    }
    if (r >= 9223372036854775808.0) return x >= 0 ? INT64_MAX : INT64_MIN;
    return x >= 0 ? (int64_t)r : -(int64_t)r;
  }
  static void setOperand(tsc_where_op *o, DataType t, const Value &v, bool hi, Program *p) {
    if (t == DataType::text) {
      (hi ? o->i_hi : o->i_lo) = p->addText(convertText(v));
    } else if (t == DataType::integer) {
      const int64_t x = std::holds_alternative<double>(v) ? dartRound(std::get<double>(v)) : std::get<int64_t>(v);
      (hi ? o->i_hi : o->i_lo) = x;
    } else {
      const double x = std::holds_alternative<int64_t>(v) ? (double)std::get<int64_t>(v) : std::get<double>(v);
      (hi ? o->f_hi : o->f_lo) = x;
    }
  }
  static void emit(const Leaf &leaf, const std::map<std::string, std::pair<uint32_t, DataType>> &columns,
                   Program *p) {
    auto it = columns.find(leaf.field);
    if (it == columns.end()) throw std::invalid_argument("WHERE names field '" + leaf.field + "' which has no attribute column");
    const DataType t = it->second.second;
    tsc_where_op o = node(TSC_W_LEAF, 0);
    o.column_id = it->second.first;
    const bool null0 = leaf.args.empty() || std::holds_alternative<std::monostate>(leaf.args[0]);
    static const std::map<std::string, uint8_t> simple = {{"=", TSC_OP_EQ}, {"!=", TSC_OP_NE}, {"<>", TSC_OP_NE},
                                                          {">", TSC_OP_GT}, {">=", TSC_OP_GE}, {"<", TSC_OP_LT},
                                                          {"<=", TSC_OP_LE}};
    auto s = simple.find(leaf.op);
    if (s != simple.end()) {
      if (null0) {   // matcher(value, null): 0 iff value is null, else +1 (value_matcher.dart:160-163)
        switch (s->second) {
          case TSC_OP_EQ: o.op = TSC_OP_IS_NULL; break;
          case TSC_OP_NE: case TSC_OP_GT: case TSC_OP_GE: o.op = TSC_OP_IS_NOT_NULL; break;
          default: o.op = TSC_OP_FALSE; break;
        }
      } else {
        o.op = s->second;
        setOperand(&o, t, leaf.args[0], false, p);
      }
    } else if (leaf.op == "BETWEEN") {
      if (leaf.args.size() != 2 || null0 || std::holds_alternative<std::monostate>(leaf.args[1])) {
        o.op = TSC_OP_FALSE;
      } else {
        o.op = TSC_OP_BETWEEN;
        setOperand(&o, t, leaf.args[0], false, p);
        setOperand(&o, t, leaf.args[1], true, p);
      }
    } else if (leaf.op == "IN" || leaf.op == "NOT IN") {
      o.op = leaf.op == "IN" ? TSC_OP_IN : TSC_OP_NOT_IN;
      o.args_offset = (uint32_t)p->in_args.size();
      for (auto &v : leaf.args) {
        if (std::holds_alternative<std::monostate>(v)) continue;   // never equal to a non-null value
        tsc_where_op tmp = node(TSC_W_LEAF, 0);
        setOperand(&tmp, t, v, false, p);
        uint64_t raw;
        if (t != DataType::doubleType) std::memcpy(&raw, &tmp.i_lo, 8);   // text: the operand's pool index
        else std::memcpy(&raw, &tmp.f_lo, 8);
        p->in_args.push_back(raw);
        o.n++;
      }
    } else if (leaf.op == "IS") {
      o.op = null0 ? TSC_OP_IS_NULL : TSC_OP_FALSE;
    } else if (leaf.op == "IS NOT") {
      o.op = null0 ? TSC_OP_IS_NOT_NULL : TSC_OP_FALSE;
    } else if (leaf.op == "LIKE" || leaf.op == "NOT LIKE") {   // value_matcher.dart:599-604
      if (t != DataType::text) throw std::invalid_argument(leaf.op + " on a numeric field has no columnar GPU form");
      if (null0) {
        o.op = TSC_OP_FALSE;
      } else {
        o.op = leaf.op == "LIKE" ? TSC_OP_LIKE : TSC_OP_NOT_LIKE;
        setOperand(&o, t, leaf.args[0], false, p);
      }
    } else {
      throw std::invalid_argument("operator '" + leaf.op + "' has no columnar GPU form");
    }
    p->ops.push_back(o);
  }
};

// ---- one record of a batch insert ---------------------------------------------------------
struct Record {
  std::string id;                          // primary key ("" -> skipped, like prepareVectorBatchChunk)
  std::optional<VectorData> embedding;     // absent -> skipped
  std::map<std::string, Value> fields;     // attribute fields; absent / monostate -> NULL
};

// ---- the slice of ToStore / VectorIndexManager that serves vectorSearch --------------------
class GpuVectorStore {
 public:
  explicit GpuVectorStore(int deviceId = 0, uint64_t capacityRows = 1u << 20,
                          DeviceDType deviceDType = DeviceDType::float32, uint32_t kMax = 128)
      : device_(deviceId), capacity_(capacityRows), dtype_(deviceDType), k_max_(kMax) {}
  // One process, several GPUs: every index of this store is a GROUP handle — the column is
  // row-range sharded over `deviceIds` inside the library (tsc_index_create with n_devices > 1)
  // and vectorSearch merges the shards over NVLink. Nothing else in this class changes.
  GpuVectorStore(const std::vector<int> &deviceIds, uint64_t capacityRows,
                 DeviceDType deviceDType = DeviceDType::float32, uint32_t kMax = 128)
      : device_(deviceIds.empty() ? 0 : deviceIds[0]), devices_(deviceIds), capacity_(capacityRows),
        dtype_(deviceDType), k_max_(kMax) {}
  ~GpuVectorStore() { close(); }
  GpuVectorStore(const GpuVectorStore &) = delete;
  GpuVectorStore &operator=(const GpuVectorStore &) = delete;

  // TableSchema vector field + IndexSchema(type: IndexType.vector). attributeFields (new,
  // additive): integer / double / text table fields mirrored column-wise on the GPU (text:
  // dictionary-encoded) for vectorSearch(where:).
  void createVectorIndex(const std::string &tableName, const std::string &fieldName,
                         const VectorFieldConfig &fieldConfig, const VectorIndexConfig &indexConfig = {},
                         const std::map<std::string, DataType> &attributeFields = {}) {
    tsc_index_desc d;
    std::memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    d.dims = (uint32_t)fieldConfig.dimensions;
    d.metric = (uint8_t)indexConfig.distanceMetric;
    d.src_precision = TSC_SRC_F32;       // rows cross the boundary as fp32 (`_toFloat32`)
    d.dev_dtype = (uint8_t)dtype_;
    d.device_id = device_;
    if (devices_.size() > 1) {
      d.n_devices = (uint32_t)(devices_.size() < 8 ? devices_.size() : 8);
      for (uint32_t i = 0; i < d.n_devices; i++) d.device_ids[i] = devices_[i];
    }
    d.capacity_rows = capacity_;
    d.k_max = k_max_;
    d.nq_max = 64;
    auto ix = std::make_unique<Index>();
    check(tsc_index_create(&d, &ix->handle), "tsc_index_create");
    ix->fieldName = fieldName;
    ix->field = fieldConfig;
    ix->config = indexConfig;
    uint32_t cid = 0;
    for (auto &kv : attributeFields) {
      check(tsc_index_column_create(ix->handle, cid, (uint8_t)kv.second), "tsc_index_column_create");
      ix->attributes[kv.first] = {cid++, kv.second};
    }
    tables_[tableName].push_back(std::move(ix));
  }

  // VectorIndexManager.writeChanges (:297-466): rows become node ids in insertion order
  // (meta.nextNodeId++, ngh_graph_engine.dart:321-322); `__nid2pk` deltas go to the library.
  size_t batchInsert(const std::string &tableName, const std::vector<Record> &records) {
    size_t done = 0;
    for (auto &ixp : tables_[tableName]) {
      Index &ix = *ixp;
      const uint32_t dims = (uint32_t)ix.field.dimensions;
      std::vector<float> rows;
      std::vector<const Record *> kept;
      for (auto &r : records) {
        if (!r.embedding || r.id.empty()) continue;
        const size_t base = rows.size();
        rows.resize(base + dims, 0.0f);                       // _toFloat32: truncate / zero-pad
        const auto &v = r.embedding->values;
        for (size_t i = 0; i < v.size() && i < dims; i++) rows[base + i] = (float)v[i];
        if (ix.field.precision == VectorPrecision::int8)      // what the reference keeps on disk
          for (size_t i = 0; i < dims; i++) rows[base + i] = int8RoundTrip(rows[base + i]);
        kept.push_back(&r);
      }
      if (kept.empty()) continue;
      const uint64_t start = ix.nextNodeId;
      check(tsc_index_append_rows(ix.handle, start, rows.data(), kept.size()), "tsc_index_append_rows");
      std::string bytes;
      std::vector<uint64_t> offs{0};
      for (auto *r : kept) {
        bytes += r->id;
        offs.push_back(bytes.size());
        ix.pk2nid[r->id] = ix.nextNodeId++;
      }
      check(tsc_index_set_primary_keys(ix.handle, start, (const uint8_t *)bytes.data(), offs.data(), kept.size()),
            "tsc_index_set_primary_keys");
      for (auto &a : ix.attributes) {
        std::vector<uint64_t> vals(kept.size(), 0);
        std::vector<uint8_t> nulls(kept.size(), 1);
        if (a.second.second == DataType::text) {   // stored like convertValue stores it: trimmed
          std::u16string units;
          std::vector<uint64_t> toffs{0};
          for (size_t i = 0; i < kept.size(); i++) {
            auto f = kept[i]->fields.find(a.first);
            if (f != kept[i]->fields.end() && !std::holds_alternative<std::monostate>(f->second)) {
              nulls[i] = 0;
              units += convertText(f->second);
            }
            toffs.push_back(units.size());
          }
          check(tsc_index_column_append_text(ix.handle, a.second.first, start, (const uint16_t *)units.data(),
                                             toffs.data(), nulls.data(), kept.size()),
                "tsc_index_column_append_text");
          continue;
        }
        for (size_t i = 0; i < kept.size(); i++) {
          auto f = kept[i]->fields.find(a.first);
          if (f == kept[i]->fields.end() || std::holds_alternative<std::monostate>(f->second)) continue;
          nulls[i] = 0;
          vals[i] = rawValue(f->second, a.second.second);
        }
        check(tsc_index_column_append(ix.handle, a.second.first, start, vals.data(), nulls.data(), kept.size()),
              "tsc_index_column_append");
      }
      done = std::max(done, kept.size());
    }
    return done;
  }
  size_t insert(const std::string &tableName, const Record &r) { return batchInsert(tableName, {r}); }

  // In-place update of an existing row (same nodeId). Additive: the reference does not
  // forward embedding updates to the vector index (core/index_manager.dart:3125-3133).
  size_t update(const std::string &tableName, const std::string &primaryKey, const Record &r) {
    size_t n = 0;
    for (auto &ixp : tables_[tableName]) {
      Index &ix = *ixp;
      auto it = ix.pk2nid.find(primaryKey);
      if (it == ix.pk2nid.end()) continue;
      const uint64_t nid = it->second;
      if (r.embedding) {
        const uint32_t dims = (uint32_t)ix.field.dimensions;
        std::vector<float> row(dims, 0.0f);
        const auto &v = r.embedding->values;
        for (size_t i = 0; i < v.size() && i < dims; i++) row[i] = (float)v[i];
        if (ix.field.precision == VectorPrecision::int8)
          for (auto &x : row) x = int8RoundTrip(x);
        check(tsc_index_append_rows(ix.handle, nid, row.data(), 1), "tsc_index_append_rows");
      }
      for (auto &a : ix.attributes) {
        auto f = r.fields.find(a.first);
        if (f == r.fields.end()) continue;
        const uint8_t isnull = std::holds_alternative<std::monostate>(f->second) ? 1 : 0;
        if (a.second.second == DataType::text) {
          const std::u16string units = isnull ? std::u16string() : convertText(f->second);
          const uint64_t toffs[2] = {0, units.size()};
          check(tsc_index_column_append_text(ix.handle, a.second.first, nid, (const uint16_t *)units.data(), toffs,
                                             &isnull, 1), "tsc_index_column_append_text");
          continue;
        }
        const uint64_t raw = isnull ? 0 : rawValue(f->second, a.second.second);
        check(tsc_index_column_append(ix.handle, a.second.first, nid, &raw, &isnull, 1), "tsc_index_column_append");
      }
      n++;
    }
    return n;
  }

  // Cold start: stream an existing on-disk NGH index (`<indexDir>/ngh/...`) into the GPU
  // index of (tableName, fieldName); primary keys come from the caller (`__nid2pk`).
  uint64_t loadIndex(const std::string &tableName, const std::string &fieldName, const std::string &indexDir,
                     bool tombstones = true) {
    Index *ix = find(tableName, fieldName);
    if (!ix) return 0;
    tsc_ngh_info info;
    std::memset(&info, 0, sizeof info);
    info.struct_size = sizeof info;
    check(tsc_index_load_ngh(ix->handle, indexDir.c_str(), tombstones ? TSC_LOAD_TOMBSTONES : 0, &info),
          "tsc_index_load_ngh");
    ix->nextNodeId = std::max<uint64_t>(ix->nextNodeId, info.next_node_id);
    return info.next_node_id;
  }
  void setPrimaryKeys(const std::string &tableName, const std::string &fieldName, uint64_t firstNodeId,
                      const std::vector<std::string> &pks) {
    Index *ix = find(tableName, fieldName);
    if (!ix || pks.empty()) return;
    std::string bytes;
    std::vector<uint64_t> offs{0};
    for (size_t i = 0; i < pks.size(); i++) {
      bytes += pks[i];
      offs.push_back(bytes.size());
      if (!pks[i].empty()) ix->pk2nid[pks[i]] = firstNodeId + i;
    }
    check(tsc_index_set_primary_keys(ix->handle, firstNodeId, (const uint8_t *)bytes.data(), offs.data(), pks.size()),
          "tsc_index_set_primary_keys");
  }

  // deleteBatch (ngh_graph_engine.dart:411-445) + tombstone mapping (vector_index_manager.dart:416-434)
  size_t deleteKeys(const std::string &tableName, const std::vector<std::string> &primaryKeys) {
    size_t n = 0;
    for (auto &ixp : tables_[tableName]) {
      std::vector<uint64_t> nids;
      for (auto &pk : primaryKeys) {
        auto it = ixp->pk2nid.find(pk);
        if (it == ixp->pk2nid.end()) continue;
        nids.push_back(it->second);
        ixp->pk2nid.erase(it);
      }
      if (nids.empty()) continue;
      check(tsc_index_set_deleted(ixp->handle, nids.data(), nids.size(), 1), "tsc_index_set_deleted");
      const uint64_t zero[2] = {0, 0};
      for (uint64_t nid : nids)
        check(tsc_index_set_primary_keys(ixp->handle, nid, (const uint8_t *)"", zero, 1), "tsc_index_set_primary_keys");
      n = std::max(n, nids.size());
    }
    return n;
  }

  // WHERE prefilter from a set of primary keys (new, additive) — e.g. the result of any query the
  // reference's executor ran: restrict the next searches of (tableName, fieldName) to these rows;
  // std::nullopt clears it. Returns the number of rows selected.
  uint64_t setWhereFilter(const std::string &tableName, const std::string &fieldName,
                          const std::optional<std::vector<std::string>> &primaryKeys) {
    Index *ix = find(tableName, fieldName);
    if (!ix) return 0;
    if (!primaryKeys) {
      check(tsc_index_set_filter(ix->handle, nullptr, 0), "tsc_index_set_filter");
      return 0;
    }
    std::string bytes;
    std::vector<uint64_t> offs{0};
    for (auto &pk : *primaryKeys) {
      bytes += pk;
      offs.push_back(bytes.size());
    }
    uint64_t matched = 0;
    check(tsc_index_filter_primary_keys(ix->handle, (const uint8_t *)bytes.data(), offs.data(), primaryKeys->size(),
                                        &matched), "tsc_index_filter_primary_keys");
    return matched;
  }

  // ToStore.vectorSearch (tostore.dart:493-511). `where` (new, additive) is evaluated on the
  // GPU into the prefilter bitmap; nullptr searches every live row.
  std::vector<VectorSearchResult> vectorSearch(const std::string &tableName, const std::string &fieldName,
                                               const VectorData &queryVector, int topK = 10,
                                               std::optional<int> efSearch = std::nullopt,
                                               std::optional<double> distanceThreshold = std::nullopt,
                                               const QueryCondition *where = nullptr) {
    (void)efSearch;                                   // exact scan: no expansion factor
    Index *ix = find(tableName, fieldName);
    if (!ix || ix->nextNodeId == 0 || topK <= 0) return {};   // :485-504 -> const []
    if (where) {
      std::map<std::string, std::pair<uint32_t, DataType>> cols(ix->attributes.begin(), ix->attributes.end());
      auto prog = where->compile(cols);
      if (prog.n_texts())
        check(tsc_index_filter_where_text(ix->handle, prog.ops.data(), (uint32_t)prog.ops.size(),
                                          prog.in_args.data(), (uint32_t)prog.in_args.size(),
                                          (const uint16_t *)prog.text_units.data(), prog.text_offsets.data(),
                                          prog.n_texts(), nullptr), "tsc_index_filter_where_text");
      else
        check(tsc_index_filter_where(ix->handle, prog.ops.data(), (uint32_t)prog.ops.size(), prog.in_args.data(),
                                     (uint32_t)prog.in_args.size(), nullptr), "tsc_index_filter_where");
      ix->whereActive = true;
    } else if (ix->whereActive) {
      check(tsc_index_set_filter(ix->handle, nullptr, 0), "tsc_index_set_filter");
      ix->whereActive = false;
    }
    const uint32_t k = (uint32_t)topK;
    std::vector<int64_t> ids(k);
    std::vector<double> dist(k), score(k);
    std::vector<uint8_t> pks(64 * 1024);
    std::vector<uint64_t> offs(k + 1);
    uint32_t count = 0;
    check(tsc_vector_search_pk(ix->handle, queryVector.values.data(), queryVector.values.size(), k,
                               distanceThreshold ? *distanceThreshold : std::nan(""), ids.data(), dist.data(),
                               score.data(), pks.data(), pks.size(), offs.data(), &count),
          "tsc_vector_search_pk");
    std::vector<VectorSearchResult> out;
    for (uint32_t j = 0; j < count; j++)
      out.push_back({std::string((const char *)pks.data() + offs[j], (size_t)(offs[j + 1] - offs[j])), dist[j], score[j]});
    return out;                                        // ascending distance (:587)
  }

  // Batch form (additive; the reference's API is single-query): one result list per query,
  // each what vectorSearch would return; one library call (tensor-core path for 16-bit
  // columns and >= 9 queries).
  std::vector<std::vector<VectorSearchResult>> vectorSearchBatch(
      const std::string &tableName, const std::string &fieldName, const std::vector<VectorData> &queryVectors,
      int topK = 10, std::optional<double> distanceThreshold = std::nullopt) {
    std::vector<std::vector<VectorSearchResult>> out(queryVectors.size());
    Index *ix = find(tableName, fieldName);
    if (!ix || ix->nextNodeId == 0 || topK <= 0 || queryVectors.empty()) return out;
    const size_t dims = (size_t)ix->field.dimensions, k = (size_t)topK;
    for (size_t b = 0; b < queryVectors.size(); b += 64) {          // nq_max of the index
      const size_t nq = std::min<size_t>(64, queryVectors.size() - b);
      std::vector<double> vals(nq * dims, 0.0);
      for (size_t i = 0; i < nq; i++) {
        const auto &v = queryVectors[b + i].values;
        for (size_t c = 0; c < v.size() && c < dims; c++) vals[i * dims + c] = v[c];
      }
      std::vector<int64_t> ids(nq * k);
      std::vector<double> dist(nq * k), score(nq * k);
      std::vector<uint32_t> counts(nq);
      check(tsc_vector_search_batch(ix->handle, vals.data(), dims, (uint32_t)nq, (uint32_t)k,
                                    distanceThreshold ? *distanceThreshold : std::nan(""), ids.data(),
                                    dist.data(), score.data(), counts.data()),
            "tsc_vector_search_batch");
      for (size_t i = 0; i < nq; i++)
        for (uint32_t j = 0; j < counts[i]; j++) {
          uint8_t pk[4096];
          uint32_t len = 0;
          check(tsc_index_get_primary_key(ix->handle, (uint64_t)ids[i * k + j], pk, sizeof pk, &len),
                "tsc_index_get_primary_key");
          if (len == 0) continue;                                    // unmapped / tombstoned (:578-579)
          out[b + i].push_back({std::string((const char *)pk, len), dist[i * k + j], score[i * k + j]});
        }
    }
    return out;
  }

  void close() {
    for (auto &t : tables_)
      for (auto &ix : t.second)
        if (ix->handle) {
          tsc_index_destroy(ix->handle);
          ix->handle = 0;
        }
    tables_.clear();
  }

 private:
  struct Index {
    uint64_t handle = 0;
    std::string fieldName;
    VectorFieldConfig field;
    VectorIndexConfig config;
    uint64_t nextNodeId = 0;
    std::map<std::string, uint64_t> pk2nid;                                  // role of `__pk2nid`
    std::map<std::string, std::pair<uint32_t, DataType>> attributes;
    bool whereActive = false;
  };
  Index *find(const std::string &table, const std::string &field) {
    auto t = tables_.find(table);
    if (t == tables_.end()) return nullptr;
    for (auto &ix : t->second)
      if (ix->fieldName == field) return ix.get();
    return nullptr;
  }
  static float int8RoundTrip(float v) {   // setVectorFromFloat32 / getVectorAsFloat32, ngh_page.dart:368-412
    double c = (double)v < -1.0 ? -1.0 : ((double)v > 1.0 ? 1.0 : (double)v);
    c *= 127.0;
    const double q = c >= 0 ? std::floor(c + 0.5) : std::ceil(c - 0.5);
    return (float)(q / 127.0);
  }
  static uint64_t rawValue(const Value &v, DataType t) {
    uint64_t raw;
    if (t == DataType::integer) {
      const int64_t x = std::holds_alternative<double>(v) ? (int64_t)std::llround(std::get<double>(v)) : std::get<int64_t>(v);
      std::memcpy(&raw, &x, 8);
    } else {
      const double x = std::holds_alternative<int64_t>(v) ? (double)std::get<int64_t>(v) : std::get<double>(v);
      std::memcpy(&raw, &x, 8);
    }
    return raw;
  }

  int device_;
  std::vector<int> devices_;   // > 1 entry: group handle over these GPUs
  uint64_t capacity_;
  DeviceDType dtype_;
  uint32_t k_max_;
  std::map<std::string, std::vector<std::unique_ptr<Index>>> tables_;
};

}  // namespace tostore

#endif  // TOSTORE_VECTOR_HPP
