"""C++ host layer (include/tostore_vector.hpp): compiles as C++17 against the C ABI, and its
QueryCondition builder produces programs whose evaluation (library host self-test, the code
the GPU kernel shares) equals the oracle's restatement of the reference's matcher
(handler/value_matcher.dart:476-612). The GPU half of examples/vector_store_demo.cc runs on
a B200 only (tools/history/gpu_r2_first.sh)."""
import math
import os
import shutil
import subprocess

import pytest

from oracle import where_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAN, INF = math.nan, math.inf
COLS = {"price": [5, None, 30, 31, -1, 7, 7, 100, None, 0, 19, 20],
        "rating": [0.0, -0.0, NAN, None, 2.5, 4.5, -4.5, INF, -INF, 1.0, None, 3.0],
        "name": ["alice", "Alice", None, "bob", "al", "alice\nsmith", "a%b", "zo\u00eb", "\U0001F600 grin", "",
                 "bobby", "a_b"]}
TYPES = {"price": "i64", "rating": "f64", "name": "text"}
# the same conditions examples/vector_store_demo.cc builds with the C++ builder, in map form
CONDITIONS = {
    "empty": {},
    "price_lt_20": {"price": {"<": 20}},
    "price_ge_19p5": {"price": {">=": 19.5}},
    "price_ne_7": {"price": {"!=": 7}},
    "price_not_in": {"price": {"NOT IN": [5, 30, None]}},
    "price_between_and_rating": {"AND": [{"price": {"BETWEEN": {"start": 5, "end": 30}}}, {"rating": {">": 2.0}}]},
    "rating_eq_neg_zero": {"rating": {"=": -0.0}},
    "rating_ge_nan": {"rating": {">=": NAN}},
    "rating_lt_zero": {"rating": {"<": 0.0}},
    "rating_in": {"rating": {"IN": [4.5, 1]}},
    "rating_null": {"rating": {"IS": None}},
    "price_gt_null": {"price": {">": None}},
    "or_groups": {"OR": [{"price": {"<": 0}},
                         {"AND": [{"rating": {">=": 4}}, {"price": {"IS NOT": None}}]},
                         {"price": {"=": 0}}]},
    "name_eq": {"name": {"=": "  alice "}},
    "name_ne": {"name": {"!=": "bob"}},
    "name_gt": {"name": {">": "b"}},
    "name_in": {"name": {"IN": ["bob", "zo\u00eb", ""]}},
    "name_like_prefix": {"name": {"LIKE": "al%"}},
    "name_like_any": {"name": {"LIKE": "%"}},
    "name_like_one": {"name": {"LIKE": "a_b"}},
    "name_like_astral": {"name": {"LIKE": "__ grin"}},
    "name_not_like_and_price": {"AND": [{"name": {"NOT LIKE": "%b%"}}, {"price": {">=": 7}}]},
}


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    from tostore_b200 import _native as N
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = tmp_path_factory.mktemp("cpp") / "vsdemo"
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "vector_store_demo.cc"), "-o", str(exe),
                           "-L", libdir, "-ltostore_cuda", f"-Wl,-rpath,{libdir}"])
    return str(exe)


def test_cpp_query_condition_builder_equals_oracle(demo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the demo's second half would run for real (tools/history/gpu_r2_first.sh)")
    out = subprocess.run([demo], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    got = dict(line.split() for line in out.stdout.splitlines() if " " in line and set(line.split()[-1]) <= {"0", "1"})
    assert set(CONDITIONS) <= set(got)
    for name, cond in CONDITIONS.items():
        want = "".join("1" if m else "0" for m in wo.evaluate_columns(cond, COLS, TYPES, n_rows=12))
        assert got[name] == want, (name, got[name], want)
    assert "no CUDA device" in out.stdout          # the product half refuses to run without a GPU
