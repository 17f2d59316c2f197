"""The column of an unsharded handle grows on demand (the reference grows by partition files,
model/ngh_index_meta.dart:178-232): appends beyond `capacity_rows` map more physical memory
behind the row block and re-allocate the per-row side arrays; everything stored before keeps
its value, searches stay identical to the oracle. A shard of a group keeps a fixed capacity."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


@pytest.mark.parametrize("dt,metric", [(0, 0), (1, 2), (2, 1)])
def test_append_beyond_capacity_grows_the_column(dt, metric):
    import tostore_b200 as T
    from tostore_b200 import where as W
    dims, k, total = 96, 10, 9000
    rows = oracle.synth_rows(51, 0, total, dims)
    rows_dev = onp.round_dev(rows, dt)
    Q = oracle.synth_rows(52, 0, 12, dims)
    if metric == 2:
        Q = np.stack([onp.normalize_f32(q) for q in Q])
    price = np.arange(total, dtype=np.int64) % 100
    with T.GpuVectorIndex(dims, metric, capacity_rows=1000, dev_dtype=dt, k_max=16, nq_max=16) as ix:
        ix.column_create(0, W.COL_I64)
        bytes0 = ix.stats().device_bytes
        done = 0
        for step in (700, 1300, 2500, 4500):                 # crosses the capacity three times
            ix.append_rows(rows[done:done + step], first_node_id=done)
            ix.column_append(0, price[done:done + step], first_node_id=done)
            done += step
            if done == 2000:
                ix.set_deleted(np.array([3, 1500, 1999]))     # tombstones survive the next growth
            ids, dist, cnt = ix.search(Q[:3], k)               # scan path after every step
            mask = np.zeros(done, dtype=bool)
            if done >= 2000:
                mask[[3, 1500, 1999]] = True
            for q in range(3):
                oi, od = oracle.search(rows_dev[:done], Q[q], metric, k, deleted=mask)
                assert cnt[q] == k and (ids[q] == oi).all(), (done, q, ids[q], oi)
                assert (bits(dist[q]) == bits(od)).all()
        st = ix.stats()
        assert st.rows == total and st.device_bytes > bytes0
        ids, dist, cnt = ix.search(Q, k)                       # tensor path (12 queries) on the grown column
        assert ix.stats().last_path == 2
        mask = np.zeros(total, dtype=bool)
        mask[[3, 1500, 1999]] = True
        for q in (0, 5, 11):
            oi, od = oracle.search(rows_dev, Q[q], metric, k, deleted=mask)
            assert (ids[q] == oi).all() and (bits(dist[q]) == bits(od)).all()
        # the attribute column grew with it: WHERE price < 10 over all rows
        matched = ix.filter_where(W.compile_condition({"price": {"<": 10}}, {"price": (0, W.COL_I64)}))
        assert matched == int((price < 10).sum())
        ids, dist, cnt = ix.search(Q[:2], k)
        keep = (price < 10) & ~mask
        for q in range(2):
            oi, od = oracle.search(rows_dev, Q[q], metric, k, deleted=~keep)
            assert (ids[q] == oi).all() and (bits(dist[q]) == bits(od)).all()


def test_synthetic_append_grows_too_and_rows_keep_their_address():
    import tostore_b200 as T
    dims = 64
    with T.GpuVectorIndex(dims, 0, capacity_rows=4096, k_max=16, nq_max=8) as ix:
        ix.append_synthetic(9, 4096)
        p0 = ix.device_rows()[0]
        ix.append_synthetic(9, 60000, first_node_id=4096)
        assert ix.device_rows()[0] == p0 and ix.stats().rows == 64096
        rows = oracle.synth_rows(9, 0, 64096, dims)
        q = oracle.synth_rows(10, 0, 1, dims)[0]
        ids, dist, cnt = ix.search(q, 10)
        oi, od = oracle.search(rows, q, 0, 10)
        assert (ids[0] == oi).all() and (bits(dist[0]) == bits(od)).all()
