"""Pipelined device-buffer searches (tsc_index_set_pipelining): the scan launch of search i+1
is a programmatic dependent of search i's kernels and overlaps its tail. Results must be the
ones the unpipelined path returns — including tiny corpora (the next scan's main loop ends
BEFORE the previous tail does) and corpora whose near-ties fail the certificate and must
be flagged instead of repaired in-stream."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def run_stream(ix, Q, k, nq_per_call, pipelined):
    """len(Q) / nq_per_call back-to-back tsc_search_device calls on one stream."""
    import torch
    n_calls = Q.shape[0] // nq_per_call
    dq = torch.from_numpy(Q).cuda()
    o_ids = torch.full((Q.shape[0], k), -7, dtype=torch.int64, device="cuda")
    o_dist = torch.zeros((Q.shape[0], k), dtype=torch.float64, device="cuda")
    o_cnt = torch.zeros(Q.shape[0], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    ix.set_pipelining(pipelined)
    d = Q.shape[1]
    for c in range(n_calls):
        o = c * nq_per_call
        ix.search_device(dq.data_ptr() + o * d * 4, nq_per_call, k, o_ids.data_ptr() + o * k * 8,
                         o_dist.data_ptr() + o * k * 8, o_cnt.data_ptr() + o * 4, stream=s.cuda_stream)
    s.synchronize()
    ix.set_pipelining(False)
    return o_ids.cpu().numpy(), o_dist.cpu().numpy(), o_cnt.cpu().numpy()


@pytest.mark.parametrize("n,dims,metric,dt,nq_per_call", [
    (200_000, 256, 0, 0, 1), (3_000, 64, 2, 0, 1), (500, 32, 1, 0, 2), (120_000, 128, 0, 1, 4),
    (60_000, 768, 2, 2, 1)])
def test_pipelined_equals_unpipelined_and_oracle(n, dims, metric, dt, nq_per_call):
    import tostore_b200 as T
    k, n_q = 10, 48
    rows = onp.round_dev(oracle.synth_rows(31, 0, n, dims), dt)
    Q = oracle.synth_rows(32, 0, n_q, dims)
    if metric == 2:
        Q = np.stack([onp.normalize_f32(q) for q in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=8) as ix:
        ix.append_synthetic(31, n)
        a = run_stream(ix, Q, k, nq_per_call, False)
        for _ in range(3):                                   # timing-dependent: a few rounds
            b = run_stream(ix, Q, k, nq_per_call, True)
            assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all() and (a[2] == b[2]).all()
        for q in range(0, n_q, 9):
            oi, od = oracle.search(rows, Q[q], metric, k)
            assert b[2][q] == k and (b[0][q] == oi).all()
            assert (bits(b[1][q]) == bits(od)).all()
        st = ix.stats()
        assert st.uncertified_queries == 0 and st.hot_launches > 0   # sampled timer still reports
        if n >= 100_000:
            # the sparse (per-live-row) scan kernel shares the pipelined chain
            rng = np.random.default_rng(7)
            keep = rng.random(n) < 0.1
            ix.set_filter(keep)
            a = run_stream(ix, Q, k, nq_per_call, False)
            b = run_stream(ix, Q, k, nq_per_call, True)
            assert (a[0] == b[0]).all() and (bits(a[1]) == bits(b[1])).all() and (a[2] == b[2]).all()
            oi, od = oracle.search(rows, Q[5], metric, k, deleted=~keep)
            assert (b[0][5] == oi).all() and (bits(b[1][5]) == bits(od)).all()


def test_pipelined_search_flags_what_it_cannot_certify():
    """Pipelined searches carry no range launch. Every second query sits next to a 200-row
    near-tie cluster: its first pass cannot be certified — it must be FLAGGED (stats count it as
    uncertified, tsc_search_flags says a range pass is owed), never silently wrong; the same
    query re-issued with pipelining off is exact, and the easy queries in between are exact."""
    import tostore_b200 as T
    from test_gpu_certificate import near_tie_corpus
    n, dims, k = 40000, 128, 10
    rows, q_hard = near_tie_corpus(n, dims, 200, 1e-6, 0, 910)
    easy = oracle.synth_rows(33, 0, 16, dims)
    Q = np.empty((32, dims), dtype=np.float32)
    Q[0::2] = easy
    Q[1::2] = q_hard                                         # the last search is a hard one
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, k_max=16, nq_max=8) as ix:
        ix.append_rows(rows)
        ix.stats_reset()
        ids, dist, cnt = run_stream(ix, Q, k, 1, True)
        st = ix.stats()
        assert st.uncertified_queries == 16 and st.certified_queries == 16 and st.retried_queries == 0
        assert ix.search_flags(1)[0] == 1                    # range pass owed for the last search
        for q in (0, 14, 30):
            oi2, od2 = oracle.search(rows, Q[q], 0, k)
            assert (ids[q] == oi2).all() and (bits(dist[q]) == bits(od2)).all()
        ix.stats_reset()
        ids1, dist1, cnt1 = ix.search(q_hard, k)              # pipelining is off again
        oi, od = oracle.search(rows, q_hard, 0, k)
        assert cnt1[0] == k and (ids1[0] == oi).all() and (bits(dist1[0]) == bits(od)).all()
        st = ix.stats()
        assert st.retried_queries == 1 and st.uncertified_queries == 0
