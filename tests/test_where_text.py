"""Text fields in the structured WHERE prefilter: String.compareTo order, IN / BETWEEN and
LIKE / NOT LIKE (handler/value_matcher.dart:211-240, :318-331, :570-612; operand trim()
model/table_schema.dart:1421-1442), dictionary-encoded on the GPU (tsc_where.cuh).

CPU part: hand-written known answers pin the oracle's LIKE; the library's own dictionary
builder + per-string test + row evaluator (the code the kernels share) run on the host
through tsc_selftest_where_text against the oracle, including random patterns heavy in
wildcards and line terminators. GPU part: the same conditions through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle import where_oracle as wo
from tostore_b200 import where as W

I64, F64, TEXT = W.COL_I64, W.COL_F64, W.COL_TEXT
COLS = {"age": (1, I64), "name": (3, TEXT), "tag": (4, TEXT)}
TYPES = {"age": "i64", "name": "text", "tag": "text"}

NAMES = ["alice", "Alice", "bob", "bobby", "", "al", "alice\nsmith", "a%b", "a_b", "a.b", "a\\b",
         "zoë", "zoe", "\U0001F600 grin", "\ufb01 fi", "x\r\ny", "x y", "[ab]", "a*b", "(a|b)",
         "alice smith", "ALICE", "b", "bo", "élan", "a b"]
TAGS = ["red", "green", "blue", "dark red", "red-ish", "RED"]


def _columns(n=300, seed=5):
    rng = np.random.default_rng(seed)
    age = [None if rng.random() < 0.1 else int(rng.integers(0, 80)) for _ in range(n)]
    name = [None if rng.random() < 0.1 else NAMES[int(rng.integers(0, len(NAMES)))] for _ in range(n)]
    tag = [None if rng.random() < 0.2 else TAGS[int(rng.integers(0, len(TAGS)))] for _ in range(n - 30)]
    return {"age": age, "name": name, "tag": tag}          # tag: shorter column, NULL tail


CONDITIONS = [
    {"name": "alice"},
    {"name": {"=": "Alice"}},
    {"name": {"=": "  alice \t"}},                      # operands are trim()med
    {"name": {"!=": "alice"}},                          # NULL != x is true
    {"name": None},
    {"name": {"IS NOT": None}},
    {"name": {">": "b"}},                               # code-unit order: 'bob' > 'b', 'Alice' < 'b'
    {"name": {">=": "bob", "<": "al"}},                 # operator map = OR
    {"name": {"<=": ""}},
    {"name": {"<": "\ufb01"}},                          # code units: an astral char (D83D ..) sorts BELOW U+FB01
    {"name": {">": "\uffff"}},                          # nothing: surrogates D83D < FFFF
    {"name": {"BETWEEN": {"start": "a", "end": "b"}}},
    {"name": {"IN": ["bob", "zoë", "nobody", ""]}},
    {"name": {"NOT IN": ["bob", "zoë", "nobody"]}},     # NULL NOT IN is true
    {"name": {"IN": []}},
    {"name": {"LIKE": "al%"}},
    {"name": {"LIKE": "%b"}},
    {"name": {"LIKE": "%"}},                            # not 'alice\nsmith': .* stops at a line break
    {"name": {"LIKE": "a_b"}},                          # 'a%b', 'a_b', 'a.b', 'a\\b', 'a*b', 'a b'
    {"name": {"LIKE": "a.b"}},                          # '.' is literal
    {"name": {"LIKE": "a\\b"}},                         # no escape character: literal backslash
    {"name": {"LIKE": "a\\%b"}},
    {"name": {"LIKE": "[ab]"}},
    {"name": {"LIKE": "(a|b)"}},
    {"name": {"LIKE": "a*b"}},
    {"name": {"LIKE": "%\n%"}},                         # a literal line break in the pattern
    {"name": {"LIKE": "x%y"}},                          # neither 'x\r\ny' nor 'x y'
    {"name": {"LIKE": "x__y"}},
    {"name": {"LIKE": "x\r\ny"}},
    {"name": {"LIKE": "__ grin"}},                      # an astral character is two code units
    {"name": {"LIKE": "_ grin"}},
    {"name": {"LIKE": ""}},
    {"name": {"LIKE": "%li%e%"}},
    {"name": {"LIKE": "ALICE"}},                        # case-sensitive
    {"name": {"NOT LIKE": "al%"}},                      # NOT LIKE is FALSE on NULL
    {"name": {"LIKE": None}},
    {"name": {"LIKE": " al% "}},                        # the pattern is trimmed like any operand
    {"tag": {"LIKE": "%red%"}},
    {"tag": {"!=": "red"}},                             # short column: NULL tail counts
    {"tag": "RED", "age": {">": 30}},                   # two fields in one leaf = AND
    {"AND": [{"name": {"LIKE": "%o%"}}, {"OR": [{"tag": {"IN": ["red", "blue"]}}, {"age": {"<": 20}}]}]},
    {"OR": [{"name": {"LIKE": "z%"}}, {"AND": [{"tag": {"NOT LIKE": "%red%"}}, {"age": {">=": 50}}]}]},
    {"name": {"=": 5}},                                 # operand -> toString()
]


def test_oracle_like_known_answers():
    like = wo.matches_like
    assert like("alice", "al%") and like("alice", "%") and like("", "%") and like("", "")
    assert not like("alice", "al") and not like("alice", "AL%") and not like("", "_")
    assert like("a%b", "a_b") and like("a.b", "a.b") and not like("axb", "a.b")
    assert like("a\\b", "a\\b") and not like("a%b", "a\\%b") and like("a\\xyzb", "a\\%b")
    assert not like("alice\nsmith", "%") and like("alice\nsmith", "%\n%")
    assert not like("x y", "x_y") and like("x y", "x y") and not like("x\ry", "x%")
    assert like("\U0001F600", "__") and not like("\U0001F600", "_")
    assert like("[ab]", "[ab]") and not like("a", "[ab]") and like("a*b", "a*b") and not like("aab", "a*b")
    assert wo.dart_string_compare("\ufb01", "\U0001F600") > 0      # FB01 > D83D: code units, not code points
    assert wo.dart_string_compare("b", "Alice") > 0 and wo.dart_string_compare("al", "alice") < 0
    assert wo.dart_trim("\ufeff a b\u3000\n") == "a b" and wo.dart_trim("\x1fa\x1f") == "\x1fa\x1f"
    assert W.dart_trim("\ufeff a b\u3000\n") == "a b" and W.dart_trim("\x1fa\x1f") == "\x1fa\x1f"


def _selftest(prog: W.WhereProgram, columns, n_rows, col_map=COLS):
    """tsc_selftest_where_text: the library's translation, dictionary builder, per-string test
    and row evaluator (shared with the kernels) on the host."""
    from tostore_b200 import _native as N
    names = list(col_map)
    ids = np.array([col_map[n][0] for n in names], dtype=np.uint32)
    types = np.array([col_map[n][1] for n in names], dtype=np.uint8)
    vals = np.zeros((len(names), n_rows), dtype=np.uint64)
    nulls = np.ones((len(names), n_rows), dtype=np.uint8)
    row_strings = []
    for ci, name in enumerate(names):
        col = list(columns[name][:n_rows]) + [None] * max(0, n_rows - len(columns[name]))
        for r, v in enumerate(col):
            if v is None:
                continue
            nulls[ci, r] = 0
            if types[ci] == I64:
                vals[ci, r] = np.array([v], dtype=np.int64).view(np.uint64)[0]
            elif types[ci] == F64:
                vals[ci, r] = np.array([v], dtype=np.float64).view(np.uint64)[0]
        row_strings += [v if (types[ci] == TEXT and v is not None) else "" for v in col]
    units, flat = W.utf16_pool(row_strings)
    # [n_cols][n_rows + 1] absolute offsets
    offs = np.zeros((len(names), n_rows + 1), dtype=np.uint64)
    for ci in range(len(names)):
        offs[ci] = flat[ci * n_rows: ci * n_rows + n_rows + 1]
    ops, n_ops, raw, n_args = prog.buffers()
    t_units, t_offs = prog.text_buffers()
    out = np.zeros(n_rows, dtype=np.uint8)
    N.check(N.lib().tsc_selftest_where_text(
        C.cast(ops, C.c_void_p), n_ops, raw.ctypes.data, n_args, t_units.ctypes.data,
        t_offs.ctypes.data, len(prog.texts), len(names), ids.ctypes.data, types.ctypes.data,
        vals.ctypes.data, nulls.ctypes.data, units.ctypes.data, offs.ctypes.data, n_rows,
        out.ctypes.data), "tsc_selftest_where_text")
    return out.astype(bool).tolist()


@pytest.mark.parametrize("ci", range(len(CONDITIONS)))
def test_library_text_evaluation_equals_oracle_on_host(ci):
    cols = _columns()
    n = len(cols["age"])
    cond = CONDITIONS[ci]
    want = wo.evaluate_columns(cond, cols, TYPES, n_rows=n)
    got = _selftest(W.compile_condition(cond, COLS), cols, n)
    assert got == want, cond


def test_like_known_rows():
    """A few answers spelled out (independent of both implementations' code)."""
    cols = {"age": [1] * 6, "name": ["alice", "alice\nsmith", None, "a%b", "x\r\ny", "bob"], "tag": []}
    def run(cond):
        return _selftest(W.compile_condition(cond, COLS), cols, 6)
    assert run({"name": {"LIKE": "%"}}) == [True, False, False, True, False, True]
    assert run({"name": {"NOT LIKE": "%"}}) == [False, True, False, False, True, False]
    assert run({"name": {"LIKE": "a_b"}}) == [False, False, False, True, False, False]
    assert run({"name": {"!=": "bob"}}) == [True, True, True, True, True, False]
    assert run({"name": {">": "alice"}}) == [False, True, False, False, True, True]
    assert run({"tag": None}) == [True] * 6 and run({"tag": {"LIKE": "%"}}) == [False] * 6


def test_random_like_patterns_equal_regex_oracle():
    """text_like (two-pointer matcher with one backtrack point) against the regex restatement,
    on an alphabet where wildcards, repeats and line terminators are common."""
    rng = np.random.default_rng(2024)
    alpha = ["a", "b", "a", "b", "%", "_", "\n", " ", "\r", "c"]
    strs = ["".join(alpha[int(i)] for i in rng.integers(0, len(alpha), int(rng.integers(0, 9))))
            for _ in range(400)]
    strs = [s for s in strs if wo.dart_trim(s) == s]        # stored values are trimmed by the host
    cols = {"age": [1] * len(strs), "name": strs, "tag": []}
    n = len(strs)
    for _ in range(150):
        pat = "".join(alpha[int(i)] for i in rng.integers(0, len(alpha), int(rng.integers(0, 7))))
        for op in ("LIKE", "NOT LIKE"):
            cond = {"name": {op: pat}}
            want = wo.evaluate_columns(cond, cols, TYPES, n_rows=n)
            assert _selftest(W.compile_condition(cond, COLS), cols, n) == want, (op, pat)


def test_random_text_comparisons_equal_oracle():
    rng = np.random.default_rng(7)
    alpha = ["a", "b", "B", "é", "\ufb01", "\U0001F600", "\uffff", "0"]
    strs = ["".join(alpha[int(i)] for i in rng.integers(0, len(alpha), int(rng.integers(0, 5))))
            for _ in range(300)]
    cols = {"age": [1] * len(strs), "name": [None if i % 17 == 0 else s for i, s in enumerate(strs)],
            "tag": []}
    n = len(strs)
    for _ in range(60):
        a, b = strs[int(rng.integers(0, n))], strs[int(rng.integers(0, n))]
        for cond in ({"name": {">": a}}, {"name": {"<=": a}}, {"name": {"BETWEEN": {"start": a, "end": b}}},
                     {"name": {"IN": [a, b]}}, {"name": {"NOT IN": [a, b]}}, {"name": {"!=": a}}):
            want = wo.evaluate_columns(cond, cols, TYPES, n_rows=n)
            assert _selftest(W.compile_condition(cond, COLS), cols, n) == want, cond


def test_library_rejects_malformed_text_programs():
    from tostore_b200 import _native as N
    cols = _columns(n=40)
    prog = W.compile_condition({"name": {"LIKE": "a%"}}, COLS)
    prog.ops[0].i_lo = 7                                     # operand index outside the pool
    with pytest.raises(N.TscError):
        _selftest(prog, cols, 40)
    prog = W.compile_condition({"name": {"IN": ["a", "b"]}}, COLS)
    prog.args[1] = (TEXT, 99)
    with pytest.raises(N.TscError):
        _selftest(prog, cols, 40)
    prog = W.compile_condition({"name": {"LIKE": "a%"}}, COLS)
    prog.ops[0].column_id = 1                                # LIKE on an integer column
    with pytest.raises(N.TscError):
        _selftest(prog, cols, 40)
    with pytest.raises(NotImplementedError):
        W.compile_condition({"age": {"LIKE": "1%"}}, COLS)
    with pytest.raises(TypeError):
        W.compile_condition({"name": {"=": 1.5}}, COLS)      # no exact Dart double.toString() here


# ---- GPU ------------------------------------------------------------------------------------
def _gpu_index(cols, n, d=32, **kw):
    import oracle
    from tostore_b200 import GpuVectorIndex
    ix = GpuVectorIndex(d, 0, capacity_rows=n + 64, k_max=16, nq_max=8, **kw)
    rows = oracle.synth_rows(43, 0, n, d)
    ix.append_rows(rows)
    for name, (cid, t) in COLS.items():
        ix.column_create(cid, t)
        ix.column_append(cid, cols[name])
    return ix, rows


def _check_all(ix, rows, cols, n, conditions):
    import oracle
    q = oracle.synth_rows(44, 0, 1, rows.shape[1])[0]
    for cond in conditions:
        want = np.array(wo.evaluate_columns(cond, cols, TYPES, n_rows=n), dtype=bool)
        matched = ix.filter_where(W.compile_condition(cond, COLS))
        assert matched == int(want.sum()), (cond, matched, int(want.sum()))
        ids, dist, cnt = ix.search(q, 10)
        oi, od = oracle.search(rows, q, 0, 10, filter=want)
        assert cnt[0] == len(oi) and (ids[0, : len(oi)] == oi).all(), cond
        assert (dist[0, : len(od)].view(np.int64) == od.view(np.int64)).all(), cond


@pytest.mark.gpu
def test_gpu_text_where_equals_oracle():
    cols = _columns(n=1500, seed=21)
    n = len(cols["age"])
    ix, rows = _gpu_index(cols, n)
    with ix:
        _check_all(ix, rows, cols, n, CONDITIONS)


@pytest.mark.gpu
def test_gpu_text_column_many_distinct_strings_and_updates():
    """A dictionary of thousands of strings (several warps and words of the code bitmap), then
    in-place updates that add new strings."""
    from tostore_b200 import TscError
    rng = np.random.default_rng(3)
    n = 5000
    name = [None if i % 50 == 0 else f"user{int(rng.integers(0, 3000)):04d}{'x' * int(rng.integers(0, 40))}"
            for i in range(n)]
    cols = {"age": [int(i % 90) for i in range(n)], "name": name, "tag": [TAGS[i % len(TAGS)] for i in range(n)]}
    ix, rows = _gpu_index(cols, n)
    conds = [{"name": {"LIKE": "user1%"}}, {"name": {"LIKE": "%7xx%"}}, {"name": {">=": "user2000"}},
             {"name": {"NOT LIKE": "user0%"}, "tag": {"IN": ["red", "RED"]}},
             {"name": {"IN": [name[1], name[2], "nobody"]}}, {"name": {"LIKE": "user____"}}]
    with ix:
        _check_all(ix, rows, cols, n, conds)
        # overwrite 100 rows with strings the dictionary has not seen, and one with NULL
        new = [f"fresh-{i}" for i in range(100)]
        ix.column_append(3, new, first_node_id=200)
        ix.column_append(3, [None], first_node_id=10)
        cols["name"][200:300] = new
        cols["name"][10] = None
        _check_all(ix, rows, cols, n, conds + [{"name": {"LIKE": "fresh-%"}}, {"name": None}])
        from tostore_b200 import _native as N
        v = np.zeros(1, dtype=np.int64)
        with pytest.raises(TscError):                       # numeric append on a text column
            N.check(ix._lib.tsc_index_column_append(ix.handle, 3, 0, v.ctypes.data, None, 1), "append")


@pytest.mark.gpu
def test_gpu_text_where_growth_and_clear():
    """The code column regrows with the index; clear() forgets the dictionary."""
    cols = _columns(n=200, seed=9)
    ix, rows = _gpu_index(cols, 200)
    import oracle
    with ix:
        more = oracle.synth_rows(45, 200, 600, 32)
        ix.append_rows(more)                                # beyond capacity 264: grows
        extra = [NAMES[i % len(NAMES)] for i in range(600)]
        ix.column_append(3, extra)
        ix.column_append(1, [5] * 600)
        cols2 = {"age": cols["age"] + [5] * 600, "name": cols["name"] + extra, "tag": cols["tag"]}
        allrows = np.concatenate([rows, more])
        _check_all(ix, allrows, cols2, 800, [{"name": {"LIKE": "al%"}}, {"name": {">": "b"}, "age": 5},
                                            {"tag": {"LIKE": "%red%"}}])
        ix.clear()
        ix.append_rows(rows[:100])
        ix.column_append(3, ["only"] * 100)
        assert ix.filter_where(W.compile_condition({"name": "only"}, COLS)) == 100
        assert ix.filter_where(W.compile_condition({"name": {"LIKE": "al%"}}, COLS)) == 0


@pytest.mark.gpu
def test_gpu_text_where_on_group_handle():
    """Row-range shards keep their own dictionaries; one program serves all of them."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cols = _columns(n=4000, seed=33)
    n = len(cols["age"])
    ix, rows = _gpu_index(cols, n, device_ids=[0, 1])
    with ix:
        _check_all(ix, rows, cols, n, CONDITIONS[:20])


@pytest.mark.gpu
def test_store_vector_search_with_text_where():
    """`GpuVectorStore.vectorSearch(where=...)` over a text attribute: stored values go through
    convertValue (trim) like the reference's insert path, operands too."""
    import oracle
    from tostore_b200 import (GpuVectorStore, QueryCondition, VectorData, VectorDistanceMetric,
                              VectorFieldConfig, VectorIndexConfig, VectorPrecision)
    n, d = 400, 16
    rows = oracle.synth_rows(93, 0, n, d)
    title = [None if i % 19 == 0 else f"  {NAMES[i % len(NAMES)]} " for i in range(n)]   # padded on purpose
    stored = [None if t is None else wo.dart_trim(t) for t in title]
    price = [i % 50 for i in range(n)]
    st = GpuVectorStore(capacity_rows=1024)
    try:
        st.createVectorIndex("docs", "emb", VectorFieldConfig(d, VectorPrecision.float32),
                             VectorIndexConfig(VectorDistanceMetric.l2),
                             attributeFields={"title": "text", "price": "integer"})
        st.batchInsert("docs", [{"id": f"pk{i}", "emb": VectorData.fromList(rows[i]), "title": title[i],
                                 "price": price[i]} for i in range(n)])
        q = oracle.synth_rows(94, 0, 1, d)[0]
        qc = QueryCondition().where("title", "LIKE", "al%").where("price", "<", 30).orWhere("title", "=", " bob")
        want = np.array(wo.evaluate_columns(qc.build(), {"title": stored, "price": price},
                                            {"title": "text", "price": "i64"}), dtype=bool)
        assert want.any() and not want.all()
        res = st.vectorSearch("docs", fieldName="emb", queryVector=VectorData.fromList(q), topK=9, where=qc)
        oi, od = oracle.search(rows, q, 0, 9, filter=want)
        assert [r.primaryKey for r in res] == [f"pk{i}" for i in oi]
        assert [r.distance for r in res] == od.tolist()
        st.update("docs", f"pk{int(oi[0])}", {"title": "zzz"})              # leaves the filter's set
        stored[int(oi[0])] = "zzz"
        want = np.array(wo.evaluate_columns(qc.build(), {"title": stored, "price": price},
                                            {"title": "text", "price": "i64"}), dtype=bool)
        res = st.vectorSearch("docs", fieldName="emb", queryVector=VectorData.fromList(q), topK=9, where=qc)
        oi2, _ = oracle.search(rows, q, 0, 9, filter=want)
        assert [r.primaryKey for r in res] == [f"pk{i}" for i in oi2] and oi2[0] != oi[0]
    finally:
        st.close()


# ---- committed golden fixtures (tests/golden/golden_where_text.json) --------------------------
def _dec(v):
    if isinstance(v, dict) and set(v) == {"u"}:
        return "".join(chr(c) for c in v["u"])           # one str character per UTF-16 code unit
    if isinstance(v, list):
        return [_dec(x) for x in v]
    if isinstance(v, dict):
        return {k: _dec(x) for k, x in v.items()}
    return v


def _join_surrogates(v):
    """The fixture holds code units; Python strings want surrogate PAIRS joined (a lone
    surrogate stays as it is)."""
    if isinstance(v, str):
        return v.encode("utf-16-le", "surrogatepass").decode("utf-16-le", "surrogatepass")
    if isinstance(v, list):
        return [_join_surrogates(x) for x in v]
    if isinstance(v, dict):
        return {k: _join_surrogates(x) for k, x in v.items()}
    return v


def test_golden_text_where_fixtures_pin_oracle_and_library():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_where_text.json")) as f:
        g = json.load(f)
    cols = _join_surrogates(_dec(g["columns"]))
    cmap = {"title": (0, TEXT), "lang": (1, TEXT), "year": (2, I64)}
    assert len(g["cases"]) >= 40
    for case in g["cases"]:
        cond = _join_surrogates(_dec(case["cond"]))
        want = [c == "1" for c in case["match"]]
        assert wo.evaluate_columns(cond, cols, g["types"], n_rows=g["rows"]) == want, cond
        assert _selftest(W.compile_condition(cond, cmap), cols, g["rows"], col_map=cmap) == want, cond


# ---- property test: random trees over text + integer fields (host) -----------------------------
from hypothesis import HealthCheck, given, settings, strategies as st   # noqa: E402

_ALPHA = ["a", "b", "B", "%", "_", "\n", " ", "é", "\U0001F600", "0"]
_text = st.lists(st.sampled_from(_ALPHA), max_size=5).map("".join)
_ints = st.integers(-3, 80)


def _text_leaf():
    simple = st.tuples(st.sampled_from(["=", "!=", "<>", ">", ">=", "<", "<="]), _text | st.none())
    between = st.tuples(st.just("BETWEEN"), st.fixed_dictionaries({"start": _text, "end": _text}))
    inlist = st.tuples(st.sampled_from(["IN", "NOT IN"]), st.lists(_text | st.none(), max_size=4))
    like = st.tuples(st.sampled_from(["LIKE", "NOT LIKE"]), _text | st.none())
    isnull = st.tuples(st.sampled_from(["IS", "IS NOT"]), st.none())
    opmap = st.lists(simple | between | inlist | like | isnull, min_size=0, max_size=3).map(dict)
    return st.dictionaries(st.sampled_from(["name", "tag"]), opmap | _text | st.none(), min_size=1, max_size=2)


def _int_leaf():
    simple = st.tuples(st.sampled_from(["=", "!=", ">", "<="]), _ints | st.none())
    return st.fixed_dictionaries({"age": st.lists(simple, min_size=1, max_size=2).map(dict) | _ints})


_tree = st.recursive(_text_leaf() | _int_leaf(),
                     lambda kids: st.fixed_dictionaries({"AND": st.lists(kids, max_size=3)})
                     | st.fixed_dictionaries({"OR": st.lists(kids, max_size=3)}), max_leaves=6)


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(_tree, st.lists(_text | st.none(), min_size=1, max_size=40), st.integers(0, 1000))
def test_property_random_text_trees_library_equals_oracle(cond, names, seed):
    rng = np.random.default_rng(seed)
    n = len(names)
    names = [None if v is None else wo.dart_trim(v) for v in names]      # stored values are trimmed
    cols = {"age": [None if rng.random() < 0.1 else int(rng.integers(0, 80)) for _ in range(n)],
            "name": names,
            "tag": [names[int(rng.integers(0, n))] for _ in range(max(n - 3, 0))]}
    prog = W.compile_condition(cond, COLS)
    if len(prog.ops) > W.MAX_OPS:
        return
    want = wo.evaluate_columns(cond, cols, TYPES, n_rows=n)
    assert _selftest(prog, cols, n) == want


# ---- boolean fields: 0 / 1 int64 column on the device, convertValue on the host ----------------
def test_boolean_field_conditions_equal_oracle():
    """DataType.boolean (convertValue table_schema.dart:1450-1459, matcher value_matcher.dart:242-253:
    false < true): the host maps values and operands to 0 / 1, the device sees an int64 column."""
    rng = np.random.default_rng(12)
    n = 200
    active = [None if rng.random() < 0.15 else bool(rng.integers(0, 2)) for _ in range(n)]
    name = [NAMES[int(rng.integers(0, len(NAMES)))] for _ in range(n)]
    cols = {"active": active, "name": name}
    types = {"active": "bool", "name": "text"}
    compile_map = {"active": (5, W.COL_BOOL), "name": (3, TEXT)}
    device_map = {"active": (5, I64), "name": (3, TEXT)}
    device_cols = {"active": [None if v is None else int(v) for v in active], "name": name}
    conds = [{"active": True}, {"active": False}, {"active": {"=": 1}}, {"active": {"=": 0.0}},
             {"active": {"=": "yes"}}, {"active": {"=": "TRUE"}}, {"active": {"=": "no"}},
             {"active": {"!=": True}}, {"active": {">": False}}, {"active": {"<=": False}},
             {"active": {"IN": [True, None]}}, {"active": {"NOT IN": [False]}}, {"active": None},
             {"active": {"IS NOT": None}}, {"active": {"BETWEEN": {"start": False, "end": True}}},
             {"active": {">": None}}, {"active": 7},
             {"AND": [{"active": True}, {"name": {"LIKE": "a%"}}]},
             {"OR": [{"active": {"IS": None}}, {"AND": [{"active": False}, {"name": {">": "b"}}]}]}]
    for cond in conds:
        want = wo.evaluate_columns(cond, cols, types, n_rows=n)
        got = _selftest(W.compile_condition(cond, compile_map), device_cols, n, col_map=device_map)
        assert got == want, cond
    assert 0 < sum(wo.evaluate_columns({"active": True}, cols, types, n_rows=n)) < n
    with pytest.raises(NotImplementedError):
        W.compile_condition({"active": {"LIKE": "t%"}}, compile_map)
    with pytest.raises(TypeError):
        W.compile_condition({"active": {"=": [1]}}, compile_map)


def test_datetime_field_is_a_text_column_of_iso_strings():
    """DataType.datetime: stored as DateTime.toIso8601String(), compared as a string (the datetime
    matcher is the text matcher, value_matcher.dart:211-240)."""
    import datetime as dt
    cv = W.convert_datetime
    assert cv(dt.datetime(2024, 1, 2, 3, 4, 5, 6000)) == "2024-01-02T03:04:05.006"
    assert cv(dt.datetime(2024, 1, 2, 3, 4, 5, 6007)) == "2024-01-02T03:04:05.006007"
    assert cv(dt.datetime(2024, 1, 2, 3, 4, 5)) == "2024-01-02T03:04:05.000"
    assert cv(dt.datetime(987, 12, 31, 23, 59, 59, 999000, tzinfo=dt.timezone.utc)) == "0987-12-31T23:59:59.999Z"
    assert cv("2024-01-02T03:04:05.000") == "2024-01-02T03:04:05.000"
    with pytest.raises(ValueError):
        cv(dt.datetime(2024, 1, 2, tzinfo=dt.timezone(dt.timedelta(hours=2))))
    days = [None if i % 9 == 0 else cv(dt.datetime(2024, 1 + i % 12, 1 + i % 28, i % 24)) for i in range(150)]
    cols = {"created": days}
    compile_map, device_map = {"created": (6, W.COL_DATETIME)}, {"created": (6, TEXT)}
    for cond in ({"created": {">=": dt.datetime(2024, 6, 1)}},
                 {"created": {"BETWEEN": {"start": dt.datetime(2024, 3, 1), "end": "2024-04-30T23:59:59.999"}}},
                 {"created": {"LIKE": "2024-02-%"}}, {"created": {"!=": days[1]}}, {"created": None},
                 {"created": {"IN": [days[1], days[2], dt.datetime(1999, 1, 1)]}}):
        as_text = {"created": W._map_operands(cond["created"], cv, "datetime")}
        want = wo.evaluate_columns(as_text, cols, {"created": "text"}, n_rows=150)
        got = _selftest(W.compile_condition(cond, compile_map), cols, 150, col_map=device_map)
        assert got == want, cond
    assert sum(wo.evaluate_columns({"created": {"LIKE": "2024-02-%"}}, cols, {"created": "text"}, n_rows=150)) > 0


def test_query_condition_convenience_builders():
    """whereLike / whereContains / whereStartsWith / whereEndsWith / whereEmpty / whereContainsAny ...
    (query/query_condition.dart:574-678) produce the same maps as the spelled-out `where` calls."""
    Q = W.QueryCondition
    assert Q().whereContains("name", "li").build() == {"name": {"LIKE": "%li%"}}
    assert Q().whereNotContains("name", "li").build() == {"name": {"NOT LIKE": "%li%"}}
    assert Q().whereStartsWith("name", "al").build() == {"name": {"LIKE": "al%"}}
    assert Q().whereEndsWith("name", "b").build() == {"name": {"LIKE": "%b"}}
    assert Q().whereLike("name", "a_b").whereGreaterThan("age", 3).build() == \
        {"AND": [{"name": {"LIKE": "a_b"}}, {"age": {">": 3}}]}
    assert Q().whereEmpty("name").build() == {"OR": [{"name": {"IS": None}}, {"name": {"=": ""}}]}
    assert Q().whereNotEmpty("name").build() == {"AND": [{"name": {"IS NOT": None}}, {"name": {"!=": ""}}]}
    cols = _columns()
    n = len(cols["age"])
    for qc in (Q().whereContains("name", "li").whereLessThanOrEqualTo("age", 40),
               Q().whereEmpty("name").whereNotEqual("tag", "red"),
               Q().where("age", ">", 70).orWhere("name", "LIKE", "z%").whereContainsAny("tag", ["ee", "lu"]),
               Q().whereContainsAny("name", ["ob", "\n", "%"])):
        cond = qc.build()
        want = wo.evaluate_columns(cond, cols, TYPES, n_rows=n)
        assert _selftest(W.compile_condition(cond, COLS), cols, n) == want, cond
        assert any(want) and not all(want), cond


def test_c_oracle_like_and_compare_equal_the_python_restatement():
    """Three independent LIKE implementations pin each other: the C oracle's dynamic programme,
    where_oracle's regex, and (through the other tests) the library's two-pointer matcher."""
    import oracle
    lib = oracle.c_oracle()
    rng = np.random.default_rng(99)
    alpha = ["a", "b", "%", "_", "\n", "\r", " ", " ", "é", "\U0001F600"]

    def rnd(maxlen):
        return "".join(alpha[int(i)] for i in rng.integers(0, len(alpha), int(rng.integers(0, maxlen))))

    def units(s):
        u = np.array(wo.code_units(s), dtype=np.uint16)
        return (u if u.size else np.zeros(1, dtype=np.uint16)), len(wo.code_units(s))

    for _ in range(3000):
        s, p = rnd(9), rnd(7)
        (us, ns), (up, npat) = units(s), units(p)
        assert lib.tso_like_match(us, ns, up, npat) == int(wo.matches_like(s, p)), (s, p)
        assert lib.tso_string_compare(us, ns, up, npat) == wo.dart_string_compare(s, p), (s, p)
    for s, p, want in (("alice", "al%", 1), ("alice\nsmith", "%", 0), ("alice\nsmith", "%\n%", 1), ("", "", 1),
                       ("", "%%", 1), ("a", "", 0), ("\U0001F600", "__", 1), ("a%b", "a\\%b", 0)):
        (us, ns), (up, npat) = units(s), units(p)
        assert lib.tso_like_match(us, ns, up, npat) == want, (s, p)
