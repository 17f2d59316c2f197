"""Host-side arithmetic of tsc_vector_search — `_toFloat32`, `_normalizeFloat32`
(core/vector_index_manager.dart:1385-1408) and `_distanceToScore` (:1411-1423) — pinned bit
for bit against the oracle's restatements through the library's self-test hooks (no GPU)."""
import ctypes as C
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import oracle_np as onp
from tostore_b200 import _native as N


def lib_prep(values, dims, metric):
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.full(dims, 7.0, dtype=np.float32)
    N.check(N.lib().tsc_selftest_query_prep(dims, metric, v.ctypes.data if v.size else None, v.size,
                                            out.ctypes.data), "prep")
    return out


def oracle_prep(values, dims, metric):
    q = onp.to_float32(np.asarray(values, dtype=np.float64), dims)
    return onp.normalize_f32(q) if metric == 2 else q


@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("length", [0, 1, 31, 48, 49, 200])
def test_query_prep_truncates_pads_and_normalises_like_the_oracle(metric, length):
    dims = 48
    rng = np.random.default_rng(length * 3 + metric)
    v = rng.standard_normal(length) * 1.7
    got, want = lib_prep(v, dims, metric), oracle_prep(v, dims, metric)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


def test_query_prep_special_values():
    dims = 8
    for metric in (0, 2):
        zero = lib_prep(np.zeros(dims), dims, metric)
        assert not zero.any()                                    # zero vector returned unchanged
        big = np.array([1e300, -1e300, 1e-320, 3.0])             # overflow to inf / underflow on the fp32 store
        got, want = lib_prep(big, dims, metric), oracle_prep(big, dims, metric)
        assert (got.view(np.uint32) == want.view(np.uint32)).all() or (np.isnan(got) == np.isnan(want)).all()
    demo = [i * 0.015 for i in range(128)]                       # example/lib/tostore_example.dart:388-406
    got, want = lib_prep(demo, 128, 2), oracle_prep(demo, 128, 2)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert abs(float(np.sqrt((got.astype(np.float64) ** 2).sum())) - 1.0) < 1e-6


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(-1e6, 1e6, allow_nan=False), min_size=0, max_size=40), st.integers(1, 33), st.integers(0, 2))
def test_property_query_prep_equals_oracle(values, dims, metric):
    got, want = lib_prep(values, dims, metric), oracle_prep(values, dims, metric)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_distance_to_score_equals_oracle(metric):
    clib = oracle.c_oracle()
    f = N.lib().tsc_selftest_distance_to_score
    for d in [0.0, -0.0, 1e-12, 0.5, 1.0, 1.5, 2.0, 2.5, -3.0, 40.0, -40.0, 800.0, -800.0, math.inf, -math.inf]:
        got, want = f(metric, d), clib.tso_distance_to_score(d, metric)
        assert np.float64(got).view(np.int64) == np.float64(want).view(np.int64), (metric, d, got, want)
    assert math.isnan(f(metric, math.nan)) == math.isnan(clib.tso_distance_to_score(math.nan, metric))
