"""Generate tests/golden/golden_where.json with oracle/where_oracle.py.

The reference has no WHERE for vectors and no test of ConditionRecordMatcher, so these
fixtures come from the Python restatement of its rules (hand-checked known answers pin
that restatement in tests/test_where.py) and then pin the library's evaluator on the CPU
(tsc_selftest_where) and on the GPU (tsc_index_filter_where).
Run from the repo root:  python tests/golden/make_golden_where.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import where_oracle as wo  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_where.json")
TYPES = {"price": "i64", "rating": "f64", "stock": "i64"}


def enc(v):
    """JSON has no NaN / inf / -0.0: encode doubles as hex strings."""
    if isinstance(v, float):
        return {"f": v.hex()}
    if isinstance(v, list):
        return [enc(x) for x in v]
    if isinstance(v, dict):
        return {k: enc(x) for k, x in v.items()}
    return v


def main():
    rng = np.random.default_rng(20261017)
    n = 96
    special = [0.0, -0.0, math.nan, math.inf, -math.inf, 4.5, -4.5]
    cols = {
        "price": [None if rng.random() < 0.12 else int(rng.integers(-3, 40)) for _ in range(n)],
        "rating": [None if rng.random() < 0.12 else
                   (special[int(rng.integers(0, len(special)))] if rng.random() < 0.25
                    else float(np.round(rng.random() * 5, 1))) for _ in range(n)],
        "stock": [int(rng.integers(0, 3)) for _ in range(n - 16)],
    }
    conds = [
        {"price": {"<": 20}},
        {"price": {">=": 19.5}},
        {"price": {"!=": 7}},
        {"price": {"NOT IN": [1, 2, 3]}},
        {"price": {"BETWEEN": {"start": 5, "end": 15}}, "rating": {">": 2.0}},
        {"rating": {"=": -0.0}},
        {"rating": {">=": math.nan}},
        {"rating": {"<": 0.0}},
        {"rating": {"IN": [4.5, math.inf]}},
        {"stock": {"=": 0}},
        {"stock": {"!=": 0}},
        {"stock": None},
        {"OR": [{"price": {"<": 0}}, {"AND": [{"rating": {">=": 4}}, {"stock": {">": 0}}]}]},
        {"AND": [{"price": {">": 5, "<": 0}}, {"rating": {"IS NOT": None}}]},
        {"OR": []},
    ]
    cases = []
    for c in conds:
        match = wo.evaluate_columns(c, cols, TYPES, n_rows=n)
        cases.append({"cond": enc(c), "match": "".join("1" if m else "0" for m in match)})
    with open(OUT, "w") as f:
        json.dump({"types": TYPES, "rows": n, "columns": enc(cols), "cases": cases}, f, indent=1)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
