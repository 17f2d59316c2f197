"""Exactness certificate of the candidate stage (tsc_tail.cuh) and the range pass behind it.

The reference re-ranks everything it kept (core/ngh_graph_engine.dart:115-134); this path
keeps K' = max(2k, 20) rows chosen by an approximate key, so it must prove it kept enough.
These tests build corpora where it did NOT — hundreds of rows within fp32 / bf16 key noise
of the k-th neighbour — and require the ids to equal the oracle's anyway (range pass), plus
the bookkeeping in tsc_stats / tsc_search_flags. i.i.d. Gaussian data never gets here: there
the rank-k to rank-K' gap dwarfs the key error and every query certifies on the first pass.
"""
import threading

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def t():
    import tostore_b200
    return tostore_b200


def near_tie_corpus(n, dims, n_cluster, rel, dt, seed):
    """Gaussian corpus + a cluster of `n_cluster` near-copies of one row: copy j differs from
    the centre in coordinate j % dims by a relative step ~rel (at least one ulp of the storage
    type), so their exact distances to a query next to the centre are all different but sit
    within ~rel of each other. The query is the centre plus a small offset: its nearest
    neighbours are the cluster."""
    rng = np.random.default_rng(seed)
    rows = oracle.synth_rows(seed, 0, n, dims).copy()
    centre = rows[0].copy()
    pos = rng.choice(np.arange(1, n), size=n_cluster, replace=False)
    for j, p in enumerate(pos):
        c = centre.copy()
        i = j % dims
        step = abs(c[i]) * rel * (1 + j // dims)
        ulp = np.spacing(np.float32(abs(c[i]))) * (1 << (16 if dt == 1 else (13 if dt == 2 else 0)))
        c[i] += max(step, float(ulp) * (1 + j // dims)) * (1 if j % 2 else -1)
        rows[p] = c
    rows = onp.round_dev(rows, dt)
    q = (centre + 0.05 * rng.standard_normal(dims)).astype(np.float32)
    return rows, q


@pytest.mark.parametrize("rel", [1e-6, 1e-4])
@pytest.mark.parametrize("metric", [0, 1, 2])
@pytest.mark.parametrize("k,n_cluster", [(10, 200), (100, 300)])
def test_scan_path_near_ties_need_the_range_pass(rel, metric, k, n_cluster):
    T = t()
    n, dims = 40000, 128
    rows, q = near_tie_corpus(n, dims, n_cluster, rel, 0, 900 + k)
    if metric == 2:
        q = onp.normalize_f32(q)
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, k_max=128, nq_max=8) as ix:
        ix.append_rows(rows)
        for nq in (1, 3):
            Q = np.stack([q] * nq)
            ix.stats_reset()
            ids, dist, cnt = ix.search(Q, k)
            oi, od = oracle.search(rows, q, metric, k)
            for qi in range(nq):
                assert cnt[qi] == k
                assert (ids[qi] == oi).all(), (metric, k, rel, ids[qi][:12], oi[:12])
                assert (bits(dist[qi]) == bits(od)).all()
            st = ix.stats()
            assert st.certified_queries + st.retried_queries == nq and st.uncertified_queries == 0
            assert (ix.search_flags(nq) == 0).all()
        if metric == 0 and rel < 1e-5:
            # the cluster is far denser than the fp32 key can resolve: the first pass alone
            # cannot have been certified
            assert st.retried_queries > 0 and st.range_rows >= k


@pytest.mark.parametrize("dt", [1, 2])
@pytest.mark.parametrize("metric", [0, 1, 2])
def test_tensor_path_near_ties_need_the_range_pass(dt, metric):
    """16-bit column, batch of 24 queries -> tcgen05 path with the query rounded to the storage
    type: key noise ~2^-9 |q||b|. A 200-row cluster one storage-ulp apart is far inside it."""
    T = t()
    n, dims, k, nq = 30000, 128, 10, 24
    rows, q = near_tie_corpus(n, dims, 200, 1e-4, dt, 77 + dt)
    others = oracle.synth_rows(5, 0, nq - 1, dims)
    Q = np.vstack([q[None, :], others]).astype(np.float32)
    if metric == 2:
        Q = np.stack([onp.normalize_f32(x) for x in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=32) as ix:
        ix.append_rows(rows)
        ids, dist, cnt = ix.search(Q, k)
        st = ix.stats()
        assert st.last_path == 2
        for qi in range(nq):
            oi, od = oracle.search(rows, Q[qi], metric, k)
            assert cnt[qi] == k and (ids[qi] == oi).all(), (dt, metric, qi, ids[qi], oi)
            assert (bits(dist[qi]) == bits(od)).all()
        assert st.certified_queries + st.retried_queries == nq and st.uncertified_queries == 0
        assert st.retried_queries >= 1          # query 0 sits in the cluster
        assert (ix.search_flags(nq) == 0).all()


def test_tensor_path_more_uncertified_queries_than_in_stream_range_launches():
    """40 of 48 queries sit in a near-tie cluster: more than the 4 x 8 the in-stream range
    launches cover; the host-buffer API runs the remaining range passes itself."""
    T = t()
    n, dims, k, nq = 20000, 64, 10, 48
    rows, q = near_tie_corpus(n, dims, 150, 1e-4, 1, 31)
    rng = np.random.default_rng(8)
    Q = np.vstack([q[None, :] + 1e-3 * rng.standard_normal((40, dims)).astype(np.float32),
                   oracle.synth_rows(6, 0, 8, dims)]).astype(np.float32)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, dev_dtype=1, k_max=16, nq_max=64) as ix:
        ix.append_rows(rows)
        ids, dist, cnt = ix.search(Q, k)
        for qi in range(nq):
            oi, od = oracle.search(rows, Q[qi], 0, k)
            assert (ids[qi] == oi).all() and (bits(dist[qi]) == bits(od)).all(), qi
        st = ix.stats()
        assert st.uncertified_queries == 0 and st.certified_queries + st.retried_queries == nq


def test_gaussian_data_certifies_on_the_first_pass():
    T = t()
    n, dims, k = 100000, 256, 10
    Q = oracle.synth_rows(12, 0, 6, dims)
    for metric, dt in ((0, 0), (2, 1), (1, 2)):
        Qp = np.stack([onp.normalize_f32(x) if metric == 2 else x for x in Q])
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=8) as ix:
            ix.append_synthetic(11, n)
            ix.search(Qp, k)
            ix.search(Qp[:1], k)
            st = ix.stats()
            assert st.certified_queries == 7 and st.retried_queries == 0 and st.range_rows == 0


def test_more_exact_duplicates_than_the_range_pass_holds_is_reported():
    """6000 identical rows nearest to the query: every key and every distance ties, the proof
    cannot be given and the range pass overflows (4096 rows) -> flag 2 / uncertified_queries.
    The result is still the reference's (ties order by node id), it is just not provable."""
    T = t()
    n, dims, k = 20000, 32, 10
    rows = oracle.synth_rows(3, 0, n, dims).copy()
    rows[1000:7000] = rows[0]
    q = (rows[0] + 0.01).astype(np.float32)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=4) as ix:
        ix.append_rows(rows)
        ids, dist, cnt = ix.search(q, k)
        assert (ix.search_flags(1) == 2).all()
        st = ix.stats()
        assert st.uncertified_queries == 1 and st.retried_queries == 0
        oi, od = oracle.search(rows, q, 0, k)
        assert (ids[0] == oi).all() and (bits(dist[0]) == bits(od)).all()


def test_zero_query_and_tiny_index_are_certified_trivially():
    T = t()
    dims = 16
    rows = oracle.synth_rows(2, 0, 15, dims)                 # fewer rows than K'
    with T.GpuVectorIndex(dims, 2, capacity_rows=64, k_max=16, nq_max=4) as ix:
        ix.append_rows(rows)
        for q in (np.zeros(dims, np.float32), onp.normalize_f32(rows[3])):
            ids, dist, cnt = ix.search(q, 10)
            oi, od = oracle.search(rows, q, 2, 10)
            assert (ids[0] == oi).all() and (bits(dist[0]) == bits(od)).all()
        assert ix.stats().uncertified_queries == 0 and ix.stats().retried_queries == 0


def test_concurrent_blocking_searches_on_one_handle_take_turns():
    """include/tostore_cuda.h: thread-safe per handle. Four threads, one handle, different
    queries and k: nobody is rejected, nobody sees another thread's result."""
    T = t()
    n, dims = 30000, 96
    rows = oracle.synth_rows(41, 0, n, dims)
    Q = oracle.synth_rows(42, 0, 8, dims)
    want = {(i, k): oracle.search(rows, Q[i], 0, k) for i in range(8) for k in (5, 10)}
    errors = []
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=4) as ix:
        ix.append_synthetic(41, n)

        def work(tid):
            try:
                for rep in range(15):
                    i, k = (tid * 2 + rep) % 8, (5, 10)[(tid + rep) % 2]
                    ids, dist, cnt = ix.search(Q[i], k)
                    oi, od = want[(i, k)]
                    assert (ids[0] == oi).all() and (bits(dist[0]) == bits(od)).all(), (tid, rep)
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=work, args=(j,)) for j in range(4)]
        [x.start() for x in th]
        [x.join() for x in th]
    assert not errors, errors


def test_second_submit_retires_the_ticket_in_flight():
    T = t()
    n, dims, k = 20000, 64, 10
    rows = oracle.synth_rows(51, 0, n, dims)
    Q = oracle.synth_rows(52, 0, 3, dims)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=4) as ix:
        ix.append_synthetic(51, n)
        polls = [ix.search_async(Q[i], k) for i in range(3)]     # three tickets, none waited on
        blocking = ix.search(Q[0], k)                             # retires the last one first
        for i in (2, 0, 1):                                       # any order
            ids, dist, cnt = polls[i](block=True)
            oi, od = oracle.search(rows, Q[i], 0, k)
            assert (ids[0] == oi).all() and (bits(dist[0]) == bits(od)).all(), i
        oi, od = oracle.search(rows, Q[0], 0, k)
        assert (blocking[0][0] == oi).all()
