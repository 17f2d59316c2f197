"""CPU tests: the C oracle against the numpy oracle, the golden fixtures, the
reference's demo vectors (closed form) and the page codec. No GPU, no product code."""
import math
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import oracle_np as onp
from conftest import golden_cases


def bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


def test_crc32_check_value():
    lib = oracle.c_oracle()
    assert lib.tso_crc32(b"123456789", 9) == 0xCBF43926       # CRC-32/IEEE check value
    data = np.random.default_rng(0).integers(0, 256, 70000, dtype=np.uint8).tobytes()
    assert lib.tso_crc32(data, len(data)) == zlib.crc32(data)


def test_demo_vectors_closed_form():
    """example/lib/tostore_example.dart:388-406: ramps i*0.01, i*0.02+0.5, query i*0.015."""
    i = np.arange(128, dtype=np.float64)
    v1 = onp.to_float32(i * 0.01, 128)
    v2 = onp.to_float32(i * 0.02 + 0.5, 128)
    q = onp.to_float32(i * 0.015, 128)
    s2 = float((i * i).sum())                                  # 690880
    lib = oracle.c_oracle()
    assert lib.tso_l2_distance(q, v1, 128) == pytest.approx(0.005 * math.sqrt(s2), rel=1e-6)
    assert lib.tso_inner_product(q, v1, 128) == pytest.approx(0.015 * 0.01 * s2, rel=1e-6)
    qn = onp.normalize_f32(q)
    d1 = lib.tso_exact_distance(qn, v1, 128, onp.COSINE)       # parallel ramps
    d2 = lib.tso_exact_distance(qn, v2, 128, onp.COSINE)
    assert abs(d1) < 1e-7 and d1 < d2
    cos2 = float((i * 0.015) @ (i * 0.02 + 0.5)) / math.sqrt(
        float((i * 0.015) @ (i * 0.015)) * float((i * 0.02 + 0.5) @ (i * 0.02 + 0.5)))
    assert d2 == pytest.approx(1.0 - cos2, abs=1e-6)
    ids, dist = oracle.search(np.stack([v1, v2]), qn, onp.COSINE, 5)
    assert ids.tolist() == [0, 1]
    for d in dist:                                              # demo's own check :1133-1141
        assert -1.0001 <= lib.tso_distance_to_score(d, onp.COSINE) <= 1.0001


def test_to_float32_pad_truncate_and_round():
    lib = oracle.c_oracle()
    v = np.array([1.0 + 2.0 ** -24, 1.0 + 3 * 2.0 ** -24, 1e39, -0.0, 5.5], dtype=np.float64)
    for dims in (3, 5, 8):
        out = np.full(dims, 7.0, dtype=np.float32)
        lib.tso_to_float32(v, v.size, dims, out)
        ref = onp.to_float32(v, dims)
        assert (out.view(np.uint32) == ref.view(np.uint32)).all()
    out = np.zeros(5, dtype=np.float32)
    lib.tso_to_float32(v, 5, 5, out)
    assert out[0] == 1.0 and out[1] == np.float32(1.0 + 2.0 ** -22) and np.isinf(out[2])


def test_normalize_zero_vector_unchanged():
    lib = oracle.c_oracle()
    z = np.zeros(16, dtype=np.float32)
    out = np.ones(16, dtype=np.float32)
    assert lib.tso_normalize_f32(z, 16, out) == 0 and not out.any()
    v = onp.synth_rows(5, 0, 1, 64)[0]
    out = np.empty(64, dtype=np.float32)
    assert lib.tso_normalize_f32(v, 64, out) == 1
    assert (out.view(np.uint32) == onp.normalize_f32(v).view(np.uint32)).all()


def test_score_mapping():
    lib = oracle.c_oracle()
    for m in (onp.L2, onp.INNER_PRODUCT, onp.COSINE):
        for d in (0.0, 0.25, 1.0, 1.5, 2.5, -0.5, -30.0, 800.0):
            assert lib.tso_distance_to_score(d, m) == onp.distance_to_score(d, m)
    assert lib.tso_distance_to_score(1.5, onp.COSINE) == 0.0
    assert lib.tso_distance_to_score(-0.5, onp.COSINE) == 1.0


def test_compare_total_order():
    lib = oracle.c_oracle()
    nan = float("nan")
    assert lib.tso_compare(-0.0, 5, 0.0, 1) < 0            # -0.0 before 0.0 (double.compareTo)
    assert lib.tso_compare(1.0, 1, nan, 0) < 0 and lib.tso_compare(nan, 0, 1e308, 9) > 0
    assert lib.tso_compare(nan, 1, nan, 2) < 0             # equal NaNs -> node id
    assert lib.tso_compare(2.0, 3, 2.0, 3) == 0


@pytest.mark.parametrize("metric", [onp.L2, onp.INNER_PRODUCT, onp.COSINE])
def test_c_oracle_equals_numpy_oracle(metric):
    rng = np.random.default_rng(metric)
    rows = rng.standard_normal((3000, 96)).astype(np.float32)
    rows[17] = 0.0                                         # zero-norm row -> cosine distance 1.0
    rows[40] = rows[41]                                    # exact tie -> node-id order
    q = rng.standard_normal(96).astype(np.float32)
    if metric == onp.COSINE:
        q = onp.normalize_f32(q)
    deleted = rng.random(3000) < 0.1
    filt = rng.random(3000) < 0.5
    for kw in ({}, {"deleted": deleted}, {"deleted": deleted, "filter": filt}):
        a = oracle.search(rows, q, metric, 25, **kw)
        b = onp.search(rows, q, metric, 25, **kw)
        c = oracle.search(rows, q, metric, 25, threads=4, **kw)
        assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()
        assert (a[0] == c[0]).all() and (bits(a[1]) == bits(c[1])).all()
    if metric == onp.COSINE:
        d = onp.exact_distances(q, rows[17:18], metric)
        assert d[0] == 1.0


def test_golden_fixtures_pin_c_oracle(golden):
    for name, seed, n, dims, dt, k in golden_cases(golden):
        rows = oracle.synth_rows(seed, 0, n, dims)
        oracle.c_oracle().tso_round_rows(rows.ctypes.data, rows.size, dt)
        qs = oracle.synth_rows(seed + 1000, 0, 3, dims)
        deleted, filt = golden[name + "/deleted"], golden[name + "/filter"]
        for metric in (0, 1, 2):
            for qi in range(3):
                q = onp.normalize_f32(qs[qi]) if metric == 2 else qs[qi]
                key = f"{name}/m{metric}/q{qi}"
                for tag, kw in (("", {}), ("del_", {"deleted": deleted}),
                                ("delfil_", {"deleted": deleted, "filter": filt})):
                    ids, dist = oracle.search(rows, q, metric, k, **kw)
                    assert (ids == golden[key + f"/{tag}ids"]).all(), key + tag
                    assert (bits(dist) == bits(golden[key + f"/{tag}dist"])).all(), key + tag
                if key + "/thr" in golden:
                    ids, dist = oracle.search(rows, q, metric, k, threshold=float(golden[key + "/thr"]))
                    assert (ids == golden[key + "/thr_ids"]).all()
                    assert len(ids) >= 4 or n < 4


def test_search_synth_equals_array_search():
    q = oracle.synth_rows(9, 0, 1, 200)[0]
    for dt in (0, 1, 2):
        rows = oracle.synth_rows(77, 0, 5000, 200)
        oracle.c_oracle().tso_round_rows(rows.ctypes.data, rows.size, dt)
        a = oracle.search(rows, q, onp.L2, 10)
        b = oracle.search_synth(77, 5000, 200, dt, q, onp.L2, 10, threads=3)
        assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all()


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 400), st.integers(1, 70), st.integers(1, 30), st.integers(0, 2),
       st.integers(0, 2 ** 31))
def test_property_k_threshold_permutation(n, dims, k, metric, seed):
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal(dims).astype(np.float32)
    ids, dist = oracle.search(rows, q, metric, k)
    assert len(ids) == min(k, n)                            # k > live rows -> short result
    assert all(oracle.c_oracle().tso_compare(dist[i], ids[i], dist[i + 1], ids[i + 1]) < 0
               for i in range(len(ids) - 1))
    perm = rng.permutation(n)                               # permutation invariance of distances
    ids2, dist2 = oracle.search(rows[perm], q, metric, k)
    assert (bits(dist) == bits(dist2)).all()
    if len(dist) > 1:                                       # strict '>' threshold
        ids3, dist3 = oracle.search(rows, q, metric, k, threshold=float(dist[0]))
        assert len(ids3) >= 1 and (dist3 <= dist[0]).all()
    dead = np.zeros(n, dtype=bool)
    dead[ids[: max(1, len(ids) // 2)]] = True               # tombstoned rows never returned
    ids4, _ = oracle.search(rows, q, metric, k, deleted=dead)
    assert not set(ids4.tolist()) & set(np.nonzero(dead)[0].tolist())


# ---------------------------------------------------------------- page codec
@pytest.mark.parametrize("dims,expect", [(128, 31), (384, 10), (512, 7), (768, 5), (1536, 2)])
def test_rows_per_page_f32(dims, expect):
    lib = oracle.c_oracle()
    assert lib.tso_vectors_per_raw_page(16384, dims, 4) == expect
    assert onp.vectors_per_raw_page(16384, dims, 4) == expect
    assert lib.tso_nodes_per_graph_page(16384, 64) == 63 == onp.nodes_per_graph_page(16384, 64)


@pytest.mark.parametrize("prec", [onp.F64, onp.F32, onp.I8])
def test_rawvec_page_roundtrip_c_vs_numpy(prec):
    lib = oracle.c_oracle()
    dims, ps = 96, 16384
    cap = onp.vectors_per_raw_page(ps, dims, onp.bytes_per_element(prec))
    rows = (np.random.default_rng(prec).standard_normal((cap - 3, dims)) * 0.6).astype(np.float32)
    page_c = np.zeros(ps, dtype=np.uint8)
    assert lib.tso_build_rawvec_page(rows, rows.shape[0], dims, prec, ps, page_c) == 0
    page_np = onp.build_rawvec_page(rows, dims, prec, ps)
    assert page_c.tobytes() == page_np                       # byte-identical writers
    out = np.zeros((cap, dims), dtype=np.float32)
    assert lib.tso_parse_rawvec_page(page_c, ps, dims, out, cap) == cap
    ref = onp.parse_rawvec_page(page_np, dims)
    assert (out.view(np.uint32) == ref.view(np.uint32)).all()
    assert not out[cap - 3:].any()                           # zero-filled tail slots
    if prec != onp.I8:
        assert (out[: cap - 3] == rows).all()
    else:
        assert np.abs(out[: cap - 3] - np.clip(rows, -1, 1)).max() <= 0.5 / 127 + 1e-7
    bad = page_c.copy()
    bad[100] ^= 1
    assert lib.tso_parse_rawvec_page(bad, ps, dims, out, cap) == -3      # CRC mismatch
    bad = page_c.copy()
    bad[0] ^= 1
    assert lib.tso_parse_rawvec_page(bad, ps, dims, out, cap) == -1      # magic
    assert lib.tso_parse_rawvec_page(page_c, ps, dims + 1, out, cap) == -6
    with pytest.raises(ValueError):
        onp.parse_page(bytes(bad))


def test_graph_page_flags_and_addressing():
    lib = oracle.c_oracle()
    flags = np.array([0, 1, 0, 3, 2, 1], dtype=np.uint8)
    page = np.zeros(16384, dtype=np.uint8)
    assert lib.tso_build_graph_page(flags, flags.size, 64, 16384, page) == 0
    assert page.tobytes() == onp.build_graph_page(flags.tolist(), 64, 16384)
    out = np.zeros(63, dtype=np.uint8)
    assert lib.tso_parse_graph_page_flags(page, 16384, out, 63) == 63
    assert out[:6].tolist() == flags.tolist() and not out[6:].any()
    import ctypes as C
    part, lp, slot = C.c_uint64(), C.c_uint32(), C.c_uint32()
    for nid in (0, 4, 5, 5119, 5120, 5 * 1024 * 3 + 7):
        lib.tso_node_location(nid, 5, 1024, C.byref(part), C.byref(lp), C.byref(slot))
        assert (part.value, lp.value, slot.value) == onp.node_location(nid, 5, 1024)
    assert onp.node_location(5120, 5, 1024) == (1, 1, 0)     # first page of partition 1
