"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the
C ABI, against the CPU oracle on the same seeded inputs. Bar: ids identical,
fp64 distances bit-identical (the re-rank kernel restates the reference's
arithmetic with separate IEEE multiply/add, sequential order)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
from conftest import golden_cases

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def t():
    import tostore_b200
    return tostore_b200


def prep(q, metric):
    return onp.normalize_f32(q) if metric == onp.COSINE else q


def assert_same(ids, dist, counts, q, oi, od, k, tag=""):
    assert counts[q] == len(oi), (tag, counts[q], len(oi))
    assert (ids[q, : len(oi)] == oi).all(), (tag, ids[q], oi)
    assert (bits(dist[q, : len(oi)]) == bits(od)).all(), (tag, dist[q], od)
    assert (ids[q, len(oi):] == -1).all() and np.isnan(dist[q, len(oi):]).all(), tag


def test_native_library_is_the_path():
    T = t()
    assert os.path.exists(T.LIB_PATH)
    from tostore_b200 import _native
    assert _native.lib().tsc_device_count() >= 1


def test_golden_fixtures(golden):
    T = t()
    for name, seed, n, dims, dt, k in golden_cases(golden):
        qs = oracle.synth_rows(seed + 1000, 0, 3, dims)
        deleted, filt = golden[name + "/deleted"], golden[name + "/filter"]
        for metric in (0, 1, 2):
            with T.GpuVectorIndex(dims, metric, capacity_rows=n + 8, dev_dtype=dt, k_max=128,
                                  nq_max=8) as ix:
                ix.append_synthetic(seed, n)
                Q = np.stack([prep(qs[i], metric) for i in range(3)])
                ids, dist, cnt = ix.search(Q, k)
                for qi in range(3):
                    key = f"{name}/m{metric}/q{qi}"
                    assert_same(ids, dist, cnt, qi, golden[key + "/ids"], golden[key + "/dist"], k, key)
                    if key + "/thr" in golden:
                        i2, d2, c2 = ix.search(Q[qi], k, threshold=float(golden[key + "/thr"]))
                        assert_same(i2, d2, c2, 0, golden[key + "/thr_ids"], golden[key + "/thr_dist"], k, key + "/thr")
                ix.set_deleted(np.nonzero(deleted)[0])
                assert ix.stats().deleted_rows == int(deleted.sum())
                ids, dist, cnt = ix.search(Q, k)
                for qi in range(3):
                    key = f"{name}/m{metric}/q{qi}"
                    assert_same(ids, dist, cnt, qi, golden[key + "/del_ids"], golden[key + "/del_dist"], k, key + "/del")
                ix.set_filter(filt)
                ids, dist, cnt = ix.search(Q, k)
                for qi in range(3):
                    key = f"{name}/m{metric}/q{qi}"
                    assert_same(ids, dist, cnt, qi, golden[key + "/delfil_ids"], golden[key + "/delfil_dist"], k, key + "/delfil")
                ix.set_filter(None)
                ix.set_deleted(np.nonzero(deleted)[0], deleted=False)
                assert ix.stats().deleted_rows == 0
                ids, dist, cnt = ix.search(Q[:1], k)
                assert_same(ids, dist, cnt, 0, golden[f"{name}/m{metric}/q0/ids"],
                            golden[f"{name}/m{metric}/q0/dist"], k, "undelete")


@pytest.mark.parametrize("dims", [1, 3, 4, 33, 128, 200, 768, 1536, 2048])
@pytest.mark.parametrize("metric", [0, 1, 2])
def test_random_rows_host_append(dims, metric):
    T = t()
    rng = np.random.default_rng(dims * 10 + metric)
    n = 3001
    rows = rng.standard_normal((n, dims)).astype(np.float32)
    rows[5] = 0.0                                   # zero-norm row
    rows[100] = rows[7]                             # exact duplicate -> tie broken by node id
    Q = rng.standard_normal((9, dims)).astype(np.float32)
    Q[8] = rows[7]                                  # query equal to a stored row
    Qp = np.stack([prep(q, metric) for q in Q])
    if dims == 1 and metric == 2:
        pytest.skip("1-d cosine distances are all exactly 0 or 2: ties only, order undefined upstream")
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, k_max=64, nq_max=16) as ix:
        ix.append_rows(rows[:1000])
        ix.append_rows(rows[1000:])                 # incremental flush-time appends
        assert ix.stats().rows == n
        for k, nq in ((10, 9), (1, 1), (64, 4), (17, 5)):
            ids, dist, cnt = ix.search(Qp[:nq], k)
            for q in range(nq):
                oi, od = oracle.search(rows, Qp[q], metric, k)
                assert_same(ids, dist, cnt, q, oi, od, k, f"d{dims} m{metric} k{k} q{q}")


def test_k_larger_than_live_rows_and_empty_index():
    T = t()
    rows = oracle.synth_rows(3, 0, 6, 32)
    q = oracle.synth_rows(4, 0, 1, 32)
    with T.GpuVectorIndex(32, 0, capacity_rows=64, k_max=32, nq_max=4) as ix:
        ids, dist, cnt = ix.search(q, 10)           # totalVectors == 0 -> []
        assert cnt[0] == 0 and (ids == -1).all()
        ix.append_rows(rows)
        ix.set_deleted([1, 4])
        ids, dist, cnt = ix.search(q, 10)
        oi, od = oracle.search(rows, q[0], 0, 10, deleted=np.isin(np.arange(6), [1, 4]))
        assert len(oi) == 4
        assert_same(ids, dist, cnt, 0, oi, od, 10)
        ix.set_deleted([0, 2, 3, 5])                # everything dead
        ids, dist, cnt = ix.search(q, 10)
        assert cnt[0] == 0
        ix.clear()
        assert ix.stats().rows == 0


@pytest.mark.parametrize("dt", [1, 2])
def test_sixteen_bit_storage_uses_rounded_rows(dt):
    T = t()
    n, dims = 5000, 320
    rows = onp.round_dev((np.random.default_rng(dt).standard_normal((n, dims)) * 3).astype(np.float32), dt)
    raw = (np.random.default_rng(dt).standard_normal((n, dims)) * 3).astype(np.float32)
    Q = np.random.default_rng(9).standard_normal((5, dims)).astype(np.float32)
    for metric in (0, 1, 2):
        Qp = np.stack([prep(q, metric) for q in Q])
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=32, nq_max=8) as ix:
            ix.append_rows(raw)                     # library rounds fp32 -> bf16 / f16 (RNE)
            ids, dist, cnt = ix.search(Qp, 10)
            for q in range(5):
                oi, od = oracle.search(rows, Qp[q], metric, 10)
                assert_same(ids, dist, cnt, q, oi, od, 10, f"dt{dt} m{metric}")


@pytest.mark.parametrize("prec", [0, 1, 2])
def test_reference_pages_equal_dense_append(prec):
    """oracle page writer -> tsc_index_append_pages == append_rows of the decoded rows."""
    T = t()
    dims, ps = 96, 16384
    bpe = onp.bytes_per_element(prec)
    rpp = onp.vectors_per_raw_page(ps, dims, bpe)
    n = rpp * 7 + 3                                  # last page partially filled
    rng = np.random.default_rng(prec)
    rows = (rng.standard_normal((n, dims)) * 0.5).astype(np.float32)
    pages = b"".join(onp.build_rawvec_page(rows[i: i + rpp], dims, prec, ps) for i in range(0, n, rpp))
    decoded = np.concatenate([onp.parse_rawvec_page(pages[i: i + ps], dims)
                              for i in range(0, len(pages), ps)])[:n]
    q = rng.standard_normal((2, dims)).astype(np.float32)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n + rpp, src_precision=prec, k_max=16, nq_max=4) as ix:
        ix.append_pages(pages[: 3 * ps], 0, ps, live_rows=n)
        ix.append_pages(pages[3 * ps:], 3, ps, live_rows=n)       # second batch of pages
        assert ix.stats().rows == n                                # zero tail slots excluded
        ids, dist, cnt = ix.search(q, 10)
        for qi in range(2):
            oi, od = oracle.search(decoded, q[qi], 0, 10)
            assert_same(ids, dist, cnt, qi, oi, od, 10, f"prec{prec}")
        bad = bytearray(pages[:ps])
        bad[200] ^= 0x40
        with pytest.raises(T.TscError) as e:
            ix.append_pages(bytes(bad), 0, ps, live_rows=n)
        assert e.value.status == -7 and "CRC" in str(e.value)
        with pytest.raises(T.TscError):
            ix.append_pages(onp.build_graph_page([0, 1], 64, ps), 0, ps, live_rows=n)


def test_graph_page_tombstones():
    T = t()
    dims, ps, n = 64, 16384, 200
    rows = oracle.synth_rows(11, 0, n, dims)
    npg = onp.nodes_per_graph_page(ps, 64)
    flags = np.zeros(n, dtype=np.uint8)
    dead = np.array([0, 5, 62, 63, 64, 130, 199])
    flags[dead] = 1
    flags[7] = 2                                     # 'updated' flag is not a tombstone
    pages = b"".join(onp.build_graph_page(flags[i: i + npg].tolist(), 64, ps) for i in range(0, n, npg))
    q = oracle.synth_rows(12, 0, 1, dims)
    with T.GpuVectorIndex(dims, 2, capacity_rows=n, k_max=128, nq_max=2) as ix:
        ix.append_rows(rows)
        ix.apply_graph_pages(pages, 0, ps)
        assert ix.stats().deleted_rows == len(dead)
        qp = prep(q[0], 2)
        ids, dist, cnt = ix.search(qp, 128)
        oi, od = oracle.search(rows, qp, 2, 128, deleted=np.isin(np.arange(n), dead))
        assert_same(ids, dist, cnt, 0, oi, od, 128)


def test_vector_search_prep_and_score():
    """tsc_vector_search == _toFloat32 + _normalizeFloat32 + search + _distanceToScore."""
    T = t()
    lib = oracle.c_oracle()
    dims, n = 48, 800
    rows = oracle.synth_rows(21, 0, n, dims)
    rng = np.random.default_rng(5)
    for metric in (0, 1, 2):
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, k_max=16, nq_max=2) as ix:
            ix.append_rows(rows)
            for length in (dims, dims - 9, dims + 20, 0):      # truncate / zero-pad, no error
                v = rng.standard_normal(length) * 1.7
                ids, dist, score = ix.vector_search(v, 5)
                q = prep(onp.to_float32(v, dims), metric)
                oi, od = oracle.search(rows, q, metric, 5)
                assert (ids == oi).all() and (bits(dist) == bits(od)).all()
                for d, s in zip(od, score):
                    assert s == lib.tso_distance_to_score(d, metric)


def test_store_api_mirrors_reference_behaviour():
    T = t()
    st = T.GpuVectorStore(capacity_rows=4096)
    assert st.vectorSearch("nope", fieldName="e", queryVector=T.VectorData.fromList([1.0])) == []
    st.createVectorIndex("embeddings", "embedding",
                         T.VectorFieldConfig(dimensions=128, precision=T.VectorPrecision.float32),
                         T.VectorIndexConfig(distanceMetric=T.VectorDistanceMetric.cosine, maxDegree=32,
                                             efSearch=64, constructionEf=128))
    q = T.VectorData.fromList([i * 0.015 for i in range(128)])
    assert st.vectorSearch("embeddings", fieldName="embedding", queryVector=q) == []   # empty index
    st.insert("embeddings", {"id": "a", "embedding": T.VectorData.fromList([i * 0.01 for i in range(128)])})
    st.insert("embeddings", {"id": "b", "embedding": T.VectorData.fromList([i * 0.02 + 0.5 for i in range(128)])})
    st.batchInsert("embeddings", [{"id": "c", "embedding": [1.0] * 64},       # short -> zero padded
                                  {"id": "", "embedding": [1.0] * 128},       # empty key skipped
                                  {"id": "d"}])                               # no vector skipped
    res = st.vectorSearch("embeddings", fieldName="embedding", queryVector=q, topK=5, efSearch=64)
    assert [r.primaryKey for r in res] == ["a", "b", "c"]
    assert abs(res[0].distance) < 1e-7 and 0.0 <= res[2].score <= res[1].score <= res[0].score <= 1.0
    assert st.vectorSearch("embeddings", fieldName="other", queryVector=q) == []
    st.delete("embeddings", ["a"])
    res = st.vectorSearch("embeddings", fieldName="embedding", queryVector=q, topK=5)
    assert [r.primaryKey for r in res] == ["b", "c"]
    res = st.vectorSearch("embeddings", fieldName="embedding", queryVector=q, topK=5,
                          distanceThreshold=res[0].distance)
    assert [r.primaryKey for r in res] == ["b"]
    st.setWhereFilter("embeddings", "embedding", ["c"])
    res = st.vectorSearch("embeddings", fieldName="embedding", queryVector=q, topK=5)
    assert [r.primaryKey for r in res] == ["c"] and set(res[0].toJson()) == {"primaryKey", "distance", "score"}
    st.close()


def test_submit_poll_and_device_buffers():
    import torch
    T = t()
    n, dims, k = 50000, 256, 10
    q = oracle.synth_rows(31, 0, 6, dims)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=8) as ix:
        ix.append_synthetic(30, n)
        ref = ix.search(q, k)
        poll = ix.search_async(q, k)
        out = None
        for _ in range(100000):
            out = poll()
            if out is not None:
                break
        assert out is not None and (out[0] == ref[0]).all() and (bits(out[1]) == bits(ref[1])).all()
        dq = torch.from_numpy(q).cuda()
        d_ids = torch.empty((6, k), dtype=torch.int64, device="cuda")
        d_dist = torch.empty((6, k), dtype=torch.float64, device="cuda")
        d_cnt = torch.empty(6, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        ix.search_device(dq.data_ptr(), 6, k, d_ids.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr(),
                         stream=s.cuda_stream)
        s.synchronize()
        assert (d_ids.cpu().numpy() == ref[0]).all()
        assert (bits(d_dist.cpu().numpy()) == bits(ref[1])).all()
        rows = oracle.synth_rows(30, 0, n, dims)
        oi, od = oracle.search(rows, q[0], 0, k, threads=4)
        assert (ref[0][0] == oi).all() and (bits(ref[1][0]) == bits(od)).all()


def test_row_range_shards_merge_to_unsharded_result():
    """Two shards of one corpus on one GPU: merge of per-shard top-k == global top-k."""
    import torch
    T = t()
    n, dims, k, nq = 40000, 128, 10, 5
    q = oracle.synth_rows(41, 0, nq, dims)
    cut = 17000
    with T.GpuVectorIndex(dims, 1, capacity_rows=n, k_max=16, nq_max=8) as whole, \
            T.GpuVectorIndex(dims, 1, capacity_rows=cut, k_max=16, nq_max=8) as a, \
            T.GpuVectorIndex(dims, 1, capacity_rows=n - cut, first_node_id=cut, k_max=16, nq_max=8) as b:
        whole.append_synthetic(40, n)
        a.append_synthetic(40, cut)
        b.append_synthetic(40, n - cut, first_node_id=cut)
        ref = whole.search(q, k)
        pa, pb = a.search(q, k), b.search(q, k)
        assert pb[0].min() >= cut
        part_ids = torch.from_numpy(np.stack([pa[0], pb[0]])).cuda()
        part_dist = torch.from_numpy(np.stack([pa[1], pb[1]])).cuda()
        o_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        o_dist = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        o_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        a.merge_shards(part_ids.data_ptr(), part_dist.data_ptr(), 2, nq, k, o_ids.data_ptr(),
                       o_dist.data_ptr(), o_cnt.data_ptr(), stream=torch.cuda.Stream().cuda_stream)
        torch.cuda.synchronize()
        assert (o_ids.cpu().numpy() == ref[0]).all()
        assert (bits(o_dist.cpu().numpy()) == bits(ref[1])).all()
        assert (o_cnt.cpu().numpy() == k).all()


def test_full_size_c2_properties():
    """BASELINE config 2 (N=10M, d=768 fp32, L2, k=10) through size-independent
    properties: planted neighbours come back in the planted order, returned
    distances are bit-identical to the oracle on those rows, no sampled row beats
    the k-th result, and the scan is idempotent."""
    T = t()
    n, dims, k, seed = 10_000_000, 768, 10, 0x705702E2
    q = oracle.synth_rows(seed + 1, 0, 1, dims)[0]
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=8) as ix:
        ix.append_synthetic(seed, n)
        assert ix.stats().rows == n
        planted = np.array([9_999_999, 0, 4_321_987, 77, 5_000_000], dtype=np.int64)
        noise = oracle.synth_rows(seed + 2, 0, len(planted), dims)
        needles = np.stack([(q + noise[j] * np.float32(0.01 * (j + 1))).astype(np.float32)
                            for j in range(len(planted))])
        for j, r in enumerate(planted):              # overwrite rows in place
            ix.append_rows(needles[j: j + 1], first_node_id=int(r))
        ids, dist, cnt = ix.search(q, k)
        ids2, dist2, _ = ix.search(q, k)
        assert (ids == ids2).all() and (bits(dist) == bits(dist2)).all()       # idempotent
        assert cnt[0] == k and ids[0, : len(planted)].tolist() == planted.tolist()
        lib = oracle.c_oracle()
        for j in range(k):                            # distances bit-exact on the returned rows
            r = int(ids[0, j])
            hit = np.nonzero(planted == r)[0]
            row = needles[hit[0]] if len(hit) else oracle.synth_rows(seed, r, 1, dims)[0]
            assert bits(dist[0, j])[()] == bits(lib.tso_l2_distance(q, row, dims))[()]
        assert all(lib.tso_compare(dist[0, j], ids[0, j], dist[0, j + 1], ids[0, j + 1]) < 0
                   for j in range(k - 1))
        rng = np.random.default_rng(1)                # sampled completeness check
        for r0 in rng.integers(0, n - 20000, 12):
            blk = oracle.synth_rows(seed, int(r0), 20000, dims)
            d = onp.exact_distances(q, blk, 0)
            better = np.nonzero(d < dist[0, k - 1])[0] + int(r0)
            assert set(better.tolist()) <= set(ids[0].tolist()) | set(planted.tolist())
        st = ix.stats()
        assert st.last_path == 1 and st.last_search_ms > 0


@pytest.mark.parametrize("dt,dims", [(0, 384), (1, 256), (0, 100)])
def test_sparse_where_filter_uses_per_row_scan(dt, dims):
    """Low-selectivity WHERE bitmap -> per-live-row bulk copies (K6); same results as the oracle."""
    T = t()
    n, k = 60000, 10
    rows = onp.round_dev(oracle.synth_rows(71, 0, n, dims), dt)
    Q = oracle.synth_rows(72, 0, 5, dims)
    rng = np.random.default_rng(3)
    for metric in (0, 2):
        Qp = np.stack([prep(q, metric) for q in Q])
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=8) as ix:
            ix.append_synthetic(71, n)
            for frac in (0.10, 0.001, 0.0):
                mask = rng.random(n) < frac
                if frac == 0.001:
                    mask[n - 1] = mask[0] = True          # edges of the bitmap
                ix.set_filter(mask)
                ix.set_deleted(np.nonzero(mask)[0][:3])    # tombstones on top of the filter
                dead = np.zeros(n, dtype=bool)
                dead[np.nonzero(mask)[0][:3]] = True
                for nq in (1, 5):
                    ids, dist, cnt = ix.search(Qp[:nq], k)
                    for q in range(nq):
                        oi, od = oracle.search(rows, Qp[q], metric, k, deleted=dead, filter=mask)
                        assert_same(ids, dist, cnt, q, oi, od, k, f"dt{dt} d{dims} m{metric} f{frac}")
                ix.set_deleted(np.nonzero(dead)[0], deleted=False)
