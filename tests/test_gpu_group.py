"""One process, several GPUs, ONE handle (include/tostore_cuda.h: tsc_index_create with
n_devices > 1; SURVEY.md §8b "one process owns all GPUs"): the form the single-process Dart
host reaches multi-GPU through (call site core/vector_index_manager.dart:538-548). Needs
>= 2 B200s; skipped otherwise."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def devices():
    from tostore_b200 import _native
    n = _native.lib().tsc_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    return list(range(min(n, 8)))


@pytest.mark.parametrize("metric,dt", [(0, 0), (2, 1), (1, 2)])
def test_group_search_equals_oracle(metric, dt):
    import tostore_b200 as T
    devs = devices()
    n, dims, k = 60_000, 192, 10
    rows = onp.round_dev(oracle.synth_rows(31, 0, n, dims), dt)
    Q = oracle.synth_rows(32, 0, 40, dims)
    Qp = np.stack([onp.normalize_f32(q) if metric == 2 else q for q in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n + 100, dev_dtype=dt, k_max=16, nq_max=64,
                          device_ids=devs) as ix:
        # three appends that straddle shard boundaries
        ix.append_rows(oracle.synth_rows(31, 0, 1000, dims), first_node_id=0)
        ix.append_rows(oracle.synth_rows(31, 1000, n - 5000, dims), first_node_id=1000)
        ix.append_synthetic(31, 4000, first_node_id=n - 4000)
        st = ix.stats()
        assert st.rows == n and st.n_devices == len(devs)
        dead = np.zeros(n, dtype=bool)
        dead[[5, n // 2, n - 1]] = True
        ix.set_deleted(np.nonzero(dead)[0])
        for nq in (1, 5, 40):                       # fused single kernel / scan batch / tensor path
            ids, dist, cnt = ix.search(Qp[:nq], k)
            for q in range(nq):
                oi, od = oracle.search(rows, Qp[q], metric, k, deleted=dead)
                assert cnt[q] == k and (ids[q] == oi).all(), (metric, nq, q, ids[q], oi)
                assert (bits(dist[q]) == bits(od)).all()
            assert (ix.search_flags(nq) == 0).all()
        # WHERE bitmap over the whole column, split on the shard boundaries by the library
        mask = np.random.default_rng(4).random(n) < 0.05
        ix.set_filter(mask)
        ids, dist, cnt = ix.search(Qp[:3], k)
        for q in range(3):
            oi, od = oracle.search(rows, Qp[q], metric, k, deleted=dead, filter=mask)
            assert (ids[q, : len(oi)] == oi).all() and (bits(dist[q, : len(oi)]) == bits(od)).all()
        ix.set_filter(None)
        # threshold + async ticket
        oi, od = oracle.search(rows, Qp[0], metric, k, deleted=dead)
        poll = ix.search_async(Qp[0], k, threshold=float(od[4]))
        ids, dist, cnt = poll(block=True)
        assert cnt[0] == 5 and (ids[0, :5] == oi[:5]).all()


def test_group_primary_keys_where_and_vector_search():
    import tostore_b200 as T
    from tostore_b200 import where as W
    devs = devices()
    n, dims, k = 20_000, 64, 8
    rows = oracle.synth_rows(61, 0, n, dims)
    with T.GpuVectorIndex(dims, 2, capacity_rows=n, k_max=16, nq_max=8, device_ids=devs) as ix:
        ix.append_rows(rows)
        ix.set_primary_keys([f"pk-{i}" if i % 7 else None for i in range(n)])
        assert ix.get_primary_key(8) == "pk-8" and ix.get_primary_key(7) is None
        assert ix.get_primary_key(n - 1) == (f"pk-{n - 1}" if (n - 1) % 7 else None)
        year = (np.arange(n) % 30 + 1995).astype(np.int64)
        ix.column_create(1, W.COL_I64)
        ix.column_append(1, year)
        prog = W.compile_condition({"year": {">=": 2020}}, {"year": (1, W.COL_I64)})
        matched = ix.filter_where(prog)
        mask = year >= 2020
        assert matched == int(mask.sum())
        q = rows[123].astype(np.float64) + 0.01
        pks, ids, dist, score = ix.vector_search_pk(q, k)
        qp = onp.normalize_f32(onp.to_float32(q, dims))
        oi, od = oracle.search(rows, qp, 2, k, filter=mask)
        keep = [i for i in oi if i % 7]
        assert list(ids) == keep and pks == [f"pk-{i}" for i in keep]


def test_group_rejects_single_shard_entry_points():
    import tostore_b200 as T
    from tostore_b200 import _native as N
    devs = devices()
    with T.GpuVectorIndex(16, 0, capacity_rows=4096, k_max=16, nq_max=4, device_ids=devs) as ix:
        with pytest.raises(N.TscError) as e:
            ix.device_rows()
        assert e.value.status == N.TSC_ERR_BAD_HANDLE
