"""Generate tests/golden/golden_where_text.json with oracle/where_oracle.py: text fields in
the WHERE prefilter (String.compareTo order, IN / BETWEEN, LIKE / NOT LIKE).

The reference has no test of ConditionRecordMatcher or matchesLike, so these fixtures come
from the Python restatement (regex-based LIKE; hand-checked known answers pin it in
tests/test_where_text.py) and then pin the library's dictionary builder, per-string test and
row evaluator (tsc_selftest_where_text — the code dict_match_kernel / where_eval_kernel share).
Strings are stored as lists of UTF-16 code units (lone surrogates and control characters
survive any JSON tooling that way).
Run from the repo root:  python tests/golden/make_golden_where_text.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import where_oracle as wo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_where_text.json")
TYPES = {"title": "text", "lang": "text", "year": "i64"}


def enc(v):
    if isinstance(v, str):
        return {"u": wo.code_units(v)}
    if isinstance(v, list):
        return [enc(x) for x in v]
    if isinstance(v, dict):
        return {k: enc(x) for k, x in v.items()}
    return v


def main():
    rng = np.random.default_rng(20261018)
    n = 120
    words = ["alpha", "Alpha", "beta", "alphabet", "", "a", "al pha", "alpha\nbeta", "a%", "a_", "50%", "x.y",
             "x\\y", "café", "cafe", "\U0001F680 launch", "ﬁn", "tab\there", "line\r\nbreak", "\ud800",
             "zeta", "ALPHA", "be", "b", "[set]", "a|b", "(x)", "^start", "end$", "q?", "plus+", "{1}"]
    cols = {
        "title": [None if rng.random() < 0.1 else words[int(rng.integers(0, len(words)))] for _ in range(n)],
        "lang": [None if rng.random() < 0.15 else ["en", "de", "en-GB", "EN", "fr"][int(rng.integers(0, 5))]
                 for _ in range(n - 20)],
        "year": [None if rng.random() < 0.1 else int(rng.integers(1990, 2030)) for _ in range(n)],
    }
    conds = [
        {"title": "alpha"},
        {"title": {"=": " alpha\t"}},
        {"title": {"!=": "alpha"}},
        {"title": {">": "b"}},
        {"title": {"<": "ﬁ"}},
        {"title": {">=": "\ud800"}},
        {"title": {"<=": ""}},
        {"title": {"BETWEEN": {"start": "a", "end": "b"}}},
        {"title": {"IN": ["beta", "café", "", "nobody"]}},
        {"title": {"NOT IN": ["beta", "café"]}},
        {"title": {"LIKE": "al%"}},
        {"title": {"LIKE": "%a"}},
        {"title": {"LIKE": "%"}},
        {"title": {"LIKE": "_"}},
        {"title": {"LIKE": "a_"}},
        {"title": {"LIKE": "a%"}},
        {"title": {"LIKE": "%\n%"}},
        {"title": {"LIKE": "line%break"}},
        {"title": {"LIKE": "line_\nbreak"}},
        {"title": {"LIKE": "__ launch"}},
        {"title": {"LIKE": "x.y"}},
        {"title": {"LIKE": "x\\y"}},
        {"title": {"LIKE": "[set]"}},
        {"title": {"LIKE": "a|b"}},
        {"title": {"LIKE": "(x)"}},
        {"title": {"LIKE": "^start"}},
        {"title": {"LIKE": "end$"}},
        {"title": {"LIKE": "q?"}},
        {"title": {"LIKE": "plus+"}},
        {"title": {"LIKE": "{1}"}},
        {"title": {"LIKE": "%l%h%"}},
        {"title": {"NOT LIKE": "al%"}},
        {"title": {"LIKE": ""}},
        {"title": {"LIKE": None}},
        {"title": None},
        {"lang": {"LIKE": "en%"}},
        {"lang": {"!=": "en"}},
        {"lang": {"IS": None}},
        {"lang": "EN", "year": {">=": 2000}},
        {"AND": [{"title": {"LIKE": "%a%"}}, {"OR": [{"lang": {"IN": ["de", "fr"]}}, {"year": {"<": 2000}}]}]},
        {"OR": [{"title": {"LIKE": "z%"}}, {"AND": [{"lang": {"NOT LIKE": "en%"}}, {"year": {"IS NOT": None}}]}]},
        {"title": {"=": 50}},
    ]
    cases = []
    for c in conds:
        m = wo.evaluate_columns(c, cols, TYPES, n_rows=n)
        cases.append({"cond": enc(c), "match": "".join("1" if x else "0" for x in m)})
    with open(OUT, "w") as f:
        json.dump({"rows": n, "types": TYPES, "columns": enc(cols), "cases": cases}, f, indent=0)
    print(f"wrote {OUT}: {len(cases)} cases over {n} rows")


if __name__ == "__main__":
    main()
