// vector_store_demo.cc — the C++ host layer (include/tostore_vector.hpp) in use.
//
//   g++ -std=c++17 -I include examples/vector_store_demo.cc -L tostore_b200 -ltostore_cuda
//       -Wl,-rpath,$PWD/tostore_b200 -o /tmp/vsdemo && /tmp/vsdemo
//
// Part 1 runs anywhere: QueryCondition programs built with the C++ builder are evaluated by
// the library's host self-test (the same translation + per-row evaluator the GPU kernel
// uses) over small fixed columns, one bitmap line per condition — tests/test_cpp_host.py
// compares them with the oracle. Part 2 needs a B200: the ToStore-shaped flow
// (createVectorIndex / batchInsert / vectorSearch with and without WHERE / delete).
#include <cstdio>
#include <limits>

#include "tostore_vector.hpp"

using namespace tostore;

static const int kRows = 12;
// price: integer field with NULLs; rating: double field with -0.0 / NaN / inf / NULL
static const std::optional<int64_t> kPrice[kRows] = {5, std::nullopt, 30, 31, -1, 7, 7, 100, std::nullopt, 0, 19, 20};
static std::optional<double> rating(int r) {
  const double nan = std::numeric_limits<double>::quiet_NaN(), inf = std::numeric_limits<double>::infinity();
  const std::optional<double> v[kRows] = {0.0, -0.0, nan, std::nullopt, 2.5, 4.5, -4.5, inf, -inf, 1.0, std::nullopt, 3.0};
  return v[r];
}

// name: text field (UTF-8 here, UTF-16 code units below), one NULL, one value with a line break
static const char *kName[kRows] = {"alice", "Alice", nullptr, "bob", "al", "alice\nsmith", "a%b", "zo\xc3\xab",
                                   "\xf0\x9f\x98\x80 grin", "", "bobby", "a_b"};

static void evaluate(const char *name, const QueryCondition &qc) {
  const std::map<std::string, std::pair<uint32_t, DataType>> cols = {{"price", {0, DataType::integer}},
                                                                      {"rating", {1, DataType::doubleType}},
                                                                      {"name", {2, DataType::text}}};
  auto prog = qc.compile(cols);
  const uint32_t ids[3] = {0, 1, 2};
  const uint8_t types[3] = {TSC_COL_I64, TSC_COL_F64, TSC_COL_TEXT};
  uint64_t vals[3][kRows] = {};
  uint8_t nulls[3][kRows];
  std::u16string row_units;
  uint64_t row_offs[3][kRows + 1] = {};
  for (int r = 0; r < kRows; r++) {
    nulls[2][r] = !kName[r];
    if (kName[r]) row_units += utf8ToUtf16(kName[r]);
    row_offs[2][r + 1] = row_units.size();
  }
  for (int r = 0; r < kRows; r++) {
    nulls[0][r] = !kPrice[r];
    int64_t p = kPrice[r].value_or(0);
    std::memcpy(&vals[0][r], &p, 8);
    auto rt = rating(r);
    nulls[1][r] = !rt;
    double d = rt.value_or(0.0);
    std::memcpy(&vals[1][r], &d, 8);
  }
  uint8_t match[kRows];
  check(tsc_selftest_where_text(prog.ops.data(), (uint32_t)prog.ops.size(), prog.in_args.data(),
                                (uint32_t)prog.in_args.size(), (const uint16_t *)prog.text_units.data(),
                                prog.text_offsets.data(), prog.n_texts(), 3, ids, types, &vals[0][0], &nulls[0][0],
                                (const uint16_t *)row_units.data(), &row_offs[0][0], kRows, match),
        "tsc_selftest_where_text");
  std::printf("%s ", name);
  for (int r = 0; r < kRows; r++) std::putchar(match[r] ? '1' : '0');
  std::putchar('\n');
}

int main() {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  // ---- part 1: conditions (names are the keys tests/test_cpp_host.py looks up) ----
  evaluate("empty", QueryCondition());
  evaluate("price_lt_20", QueryCondition().where("price", "<", int64_t{20}));
  evaluate("price_ge_19p5", QueryCondition().where("price", ">=", 19.5));            // -> round() = 20
  evaluate("price_ne_7", QueryCondition().where("price", "!=", int64_t{7}));          // NULL != 7 is true
  evaluate("price_not_in", QueryCondition().whereNotIn("price", {int64_t{5}, int64_t{30}, Value{}}));
  evaluate("price_between_and_rating", QueryCondition().whereBetween("price", int64_t{5}, int64_t{30})
                                           .where("rating", ">", 2.0));
  evaluate("rating_eq_neg_zero", QueryCondition().where("rating", "=", -0.0));
  evaluate("rating_ge_nan", QueryCondition().where("rating", ">=", nan));
  evaluate("rating_lt_zero", QueryCondition().where("rating", "<", 0.0));
  evaluate("rating_in", QueryCondition().whereIn("rating", {4.5, int64_t{1}}));       // 1 -> 1.0
  evaluate("rating_null", QueryCondition().whereNull("rating"));
  evaluate("price_gt_null", QueryCondition().where("price", ">"));                    // any non-null value
  evaluate("or_groups", QueryCondition().where("price", "<", int64_t{0}).orWhere("rating", ">=", int64_t{4})
                            .where("price", "IS NOT").orWhere("price", "=", int64_t{0}));

  // text field: String.compareTo order, operands trimmed, LIKE as ValueMatcher.matchesLike
  evaluate("name_eq", QueryCondition().where("name", "=", std::string("  alice ")));
  evaluate("name_ne", QueryCondition().where("name", "!=", std::string("bob")));        // NULL != x is true
  evaluate("name_gt", QueryCondition().where("name", ">", std::string("b")));
  evaluate("name_in", QueryCondition().whereIn("name", {std::string("bob"), std::string("zo\xc3\xab"), std::string("")}));
  evaluate("name_like_prefix", QueryCondition().whereStartsWith("name", "al"));
  evaluate("name_like_any", QueryCondition().where("name", "LIKE", std::string("%")));  // not across a line break
  evaluate("name_like_one", QueryCondition().where("name", "LIKE", std::string("a_b")));
  evaluate("name_like_astral", QueryCondition().where("name", "LIKE", std::string("__ grin")));
  evaluate("name_not_like_and_price", QueryCondition().where("name", "NOT LIKE", std::string("%b%"))
                                          .where("price", ">=", int64_t{7}));

  // ---- part 2: the ToStore-shaped flow on a GPU ----
  if (tsc_device_count() <= 0) {
    std::printf("no CUDA device: skipping the GPU part (%s)\n", tsc_last_error());
    return 0;
  }
  try {
    GpuVectorStore db(0, 4096);
    VectorIndexConfig cfg;
    cfg.distanceMetric = VectorDistanceMetric::cosine;
    db.createVectorIndex("items", "embedding", VectorFieldConfig{64, VectorPrecision::float32}, cfg,
                         {{"price", DataType::integer}, {"rating", DataType::doubleType}, {"name", DataType::text}});
    std::vector<Record> recs;
    for (int r = 0; r < 1000; r++) {
      Record rec;
      rec.id = "item-" + std::to_string(r);
      std::vector<double> v(64);
      for (int i = 0; i < 64; i++) v[i] = std::sin(0.37 * r + 0.11 * i) + 0.001 * r;
      rec.embedding = VectorData::fromList(v);
      rec.fields["price"] = int64_t{r % 50};
      if (r % 9) rec.fields["rating"] = (r % 11) * 0.5;
      rec.fields["name"] = std::string(r % 3 ? "widget " : "gadget ") + std::to_string(r % 7);
      recs.push_back(std::move(rec));
    }
    std::printf("inserted %zu\n", db.batchInsert("items", recs));
    std::vector<double> q(64);
    for (int i = 0; i < 64; i++) q[i] = std::sin(0.37 * 123 + 0.11 * i);
    for (auto &r : db.vectorSearch("items", "embedding", VectorData::fromList(q), 5))
      std::printf("  %s distance=%.12g score=%.6f\n", r.primaryKey.c_str(), r.distance, r.score);
    QueryCondition qc;
    qc.where("price", "<", int64_t{10}).where("rating", ">=", 2.0);
    std::printf("WHERE price < 10 AND rating >= 2:\n");
    for (auto &r : db.vectorSearch("items", "embedding", VectorData::fromList(q), 5, std::nullopt, std::nullopt, &qc))
      std::printf("  %s distance=%.12g score=%.6f\n", r.primaryKey.c_str(), r.distance, r.score);
    QueryCondition qt;
    qt.where("name", "LIKE", std::string("gadget%")).where("price", "<", int64_t{25});
    std::printf("WHERE name LIKE 'gadget%%' AND price < 25:\n");
    for (auto &r : db.vectorSearch("items", "embedding", VectorData::fromList(q), 5, std::nullopt, std::nullopt, &qt))
      std::printf("  %s distance=%.12g score=%.6f\n", r.primaryKey.c_str(), r.distance, r.score);
    std::printf("deleted %zu\n", db.deleteKeys("items", {"item-123"}));
    for (auto &r : db.vectorSearch("items", "embedding", VectorData::fromList(q), 3))
      std::printf("  %s distance=%.12g\n", r.primaryKey.c_str(), r.distance);
    std::printf("unknown field -> %zu results\n", db.vectorSearch("items", "nope", VectorData::fromList(q)).size());
  } catch (const TscError &e) {
    std::fprintf(stderr, "TscError: %s\n", e.what());
    return 1;
  }
  return 0;
}
