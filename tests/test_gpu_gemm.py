"""GPU tests of the tensor-core path (tcgen05 GEMM + fused top-k) through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def gemm_keys(ix, q):
    from tostore_b200 import _native as N
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty((q.shape[0], ix.stats().rows), dtype=np.float32)
    N.check(N.lib().tsc_debug_gemm_keys(ix.handle, q.ctypes.data, q.shape[0], out.ctypes.data),
            "tsc_debug_gemm_keys")
    return out


@pytest.mark.parametrize("dt", [1, 2])
@pytest.mark.parametrize("dims,n,nq", [(64, 300, 5), (128, 1000, 130), (768, 777, 64), (100, 513, 257),
                                         (1024, 600, 140), (776, 300, 9)])
def test_gemm_keys_match_matmul(dt, dims, n, nq):
    """UMMA descriptors / TMA swizzle / TMEM epilogue: raw keys vs a float64 matmul
    over the same rounded operands (tolerance: fp32 accumulation order)."""
    import tostore_b200 as T
    rng = np.random.default_rng(dims + n)
    rows = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    rr, qr = onp.round_dev(rows, dt).astype(np.float64), onp.round_dev(q, dt).astype(np.float64)
    s = qr @ rr.T
    n2 = (rr * rr).sum(axis=1)
    for metric in (0, 1, 2):
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=512) as ix:
            ix.append_rows(rows)
            keys = gemm_keys(ix, q).astype(np.float64)
            ref = {0: n2[None, :] - 2 * s, 1: -s, 2: -s / np.sqrt(n2)[None, :]}[metric]
            err = np.abs(keys - ref)
            tol = 1e-3 * (1.0 + np.abs(ref))
            bad = np.argwhere(err > tol)
            assert bad.size == 0, (metric, dims, n, nq, bad[:8], keys[tuple(bad[0])], ref[tuple(bad[0])])


@pytest.mark.parametrize("dt", [1, 2])
@pytest.mark.parametrize("metric", [0, 1, 2])
def test_gemm_path_parity_with_oracle(dt, metric):
    import tostore_b200 as T
    n, dims, nq, k = 20000, 256, 200, 10
    rows = onp.round_dev(oracle.synth_rows(61, 0, n, dims), dt)
    Q = oracle.synth_rows(62, 0, nq, dims)
    Qp = np.stack([onp.normalize_f32(q) if metric == 2 else q for q in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=256) as ix:
        ix.append_synthetic(61, n)
        ids, dist, cnt = ix.search(Qp, k)
        assert ix.stats().last_path == 2                    # tensor-core path taken
        ids8, dist8, _ = ix.search(Qp[:4], k)               # scan path on the same index
        assert ix.stats().last_path == 1
        assert (ids[:4] == ids8).all() and (bits(dist[:4]) == bits(dist8)).all()
        ids6, dist6, _ = ix.search(Qp[:6], k)               # 16-bit columns: tensor path from 5 queries
        assert ix.stats().last_path == 2
        assert (ids[:6] == ids6).all() and (bits(dist[:6]) == bits(dist6)).all()
        for q in range(0, nq, 7):
            oi, od = oracle.search(rows, Qp[q], metric, k)
            assert cnt[q] == k and (ids[q] == oi).all(), (q, ids[q], oi)
            assert (bits(dist[q]) == bits(od)).all()
        dead = np.unique(ids[:, :3].ravel())                # tombstones + filter on the GEMM path
        ix.set_deleted(dead)
        ids2, dist2, cnt2 = ix.search(Qp, k)
        assert not np.isin(ids2, dead).any()
        mask = np.zeros(n, dtype=bool)
        mask[dead] = True
        oi, od = oracle.search(rows, Qp[3], metric, k, deleted=mask)
        assert (ids2[3] == oi).all() and (bits(dist2[3]) == bits(od)).all()


def test_full_size_c3_properties():
    """BASELINE config 3 (batch-1024 cosine, N=10M d=768 bf16, k=10): planted
    neighbours, bit-exact distances on returned rows, sampled completeness."""
    import tostore_b200 as T
    n, dims, k, nq, seed = 10_000_000, 768, 10, 1024, 0x705702E3
    Q = oracle.synth_rows(seed + 1, 0, nq, dims)
    Qp = np.stack([onp.normalize_f32(q) for q in Q])
    with T.GpuVectorIndex(dims, 2, capacity_rows=n, dev_dtype=1, k_max=16, nq_max=1024) as ix:
        ix.append_synthetic(seed, n)
        planted = {0: 9_999_999, 500: 123, 1023: 5_000_001}
        for qi, r in planted.items():                       # row parallel to the query -> distance ~ 0
            ix.append_rows((Q[qi] * np.float32(3.0))[None, :], first_node_id=r)
        ids, dist, cnt = ix.search(Qp, k)
        st = ix.stats()
        assert st.last_path == 2 and (cnt == k).all()
        for qi, r in planted.items():
            assert ids[qi, 0] == r and dist[qi, 0] < 1e-4
        lib = oracle.c_oracle()
        rng = np.random.default_rng(3)
        for qi in rng.integers(0, nq, 6):
            for j in range(k):                              # bit-exact fp64 distances
                r = int(ids[qi, j])
                if r in planted.values():
                    continue
                row = onp.round_dev(oracle.synth_rows(seed, r, 1, dims), 1)[0]
                assert bits(dist[qi, j])[()] == bits(lib.tso_exact_distance(Qp[qi], row, dims, 2))[()]
            r0 = int(rng.integers(0, n - 50000))            # sampled completeness
            blk = onp.round_dev(oracle.synth_rows(seed, r0, 50000, dims), 1)
            d = onp.exact_distances(Qp[qi], blk, 2)
            better = set((np.nonzero(d < dist[qi, k - 1])[0] + r0).tolist())
            assert better <= set(ids[qi].tolist()) | set(planted.values())


# ---- fp32 column multiplied as tf32 on the tensor cores (batches of >= 9 queries) ------------

def _tf32_trunc(a):
    """What kind::tf32 reads from an fp32 operand: the low 13 mantissa bits are ignored."""
    return (np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("dims,n,nq", [(32, 300, 5), (128, 1000, 130), (768, 777, 64), (100, 513, 257)])
def test_tf32_keys_match_matmul(dims, n, nq):
    import tostore_b200 as T
    rng = np.random.default_rng(dims + n)
    rows = rng.standard_normal((n, dims)).astype(np.float32)
    q = rng.standard_normal((nq, dims)).astype(np.float32)
    s = _tf32_trunc(q).astype(np.float64) @ _tf32_trunc(rows).astype(np.float64).T
    n2 = (rows.astype(np.float64) ** 2).sum(axis=1)          # norms use the full fp32 values
    for metric in (0, 1, 2):
        with T.GpuVectorIndex(dims, metric, capacity_rows=n, k_max=16, nq_max=512) as ix:
            ix.append_rows(rows)
            keys = gemm_keys(ix, q).astype(np.float64)
            ref = {0: n2[None, :] - 2 * s, 1: -s, 2: -s / np.sqrt(n2)[None, :]}[metric]
            # the hardware may round instead of truncate: allow one tf32 ulp per product
            tol = 2e-3 * (1.0 + np.abs(ref)) + 2.0 ** -9 * np.sqrt(dims)
            assert (np.abs(keys - ref) <= tol).all(), (metric, dims, np.abs(keys - ref).max())


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_tf32_path_parity_with_oracle(metric):
    import tostore_b200 as T
    n, dims, nq, k = 20000, 256, 200, 10
    rows = oracle.synth_rows(61, 0, n, dims)
    Q = oracle.synth_rows(62, 0, nq, dims)
    Qp = np.stack([onp.normalize_f32(q) if metric == 2 else q for q in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, k_max=16, nq_max=256) as ix:
        ix.append_synthetic(61, n)
        ids, dist, cnt = ix.search(Qp, k)
        assert ix.stats().last_path == 2
        for q in range(0, nq, 7):
            oi, od = oracle.search(rows, Qp[q], metric, k)
            assert cnt[q] == k and (ids[q] == oi).all(), (q, ids[q], oi)
            assert (bits(dist[q]) == bits(od)).all()


# ---- k above the kernel's list capacity (K' > 32): truncated lists + certificate ---------------

@pytest.mark.parametrize("dt,metric", [(1, 2), (2, 1), (0, 0)])
def test_gemm_path_large_k(dt, metric):
    """A k=100 batch stays on the tensor cores: every (CTA, column half) list keeps its 32 best
    rows, the tail re-ranks the K' = k + 64 best of their union and the certificate bounds what
    the truncated lists may have dropped. ids / distances equal the oracle."""
    import tostore_b200 as T
    n, dims, nq, k = 60000, 128, 40, 100
    rows = onp.round_dev(oracle.synth_rows(71, 0, n, dims), dt)
    Q = oracle.synth_rows(72, 0, nq, dims)
    Qp = np.stack([onp.normalize_f32(q) if metric == 2 else q for q in Q])
    with T.GpuVectorIndex(dims, metric, capacity_rows=n, dev_dtype=dt, k_max=128, nq_max=64) as ix:
        ix.append_synthetic(71, n)
        ix.stats_reset()
        ids, dist, cnt = ix.search(Qp, k)
        st = ix.stats()
        assert st.last_path == 2 and st.uncertified_queries == 0
        for q in range(0, nq, 5):
            oi, od = oracle.search(rows, Qp[q], metric, k)
            assert cnt[q] == k and (ids[q] == oi).all(), (q, ids[q][:8], oi[:8])
            assert (bits(dist[q]) == bits(od)).all()


def test_gemm_path_large_k_clustered_rows_take_the_range_pass():
    """Adversarial layout for truncated lists: the 100 nearest rows of query 0 are CONSECUTIVE
    rows, so one list would have to hold all of them and drops most. The certificate must notice
    (the full list's largest key is below the K'-th candidate) and the range pass must repair
    the result."""
    import tostore_b200 as T
    n, dims, nq, k = 40000, 128, 16, 100
    rng = np.random.default_rng(5)
    rows = oracle.synth_rows(73, 0, n, dims).copy()
    Q = oracle.synth_rows(74, 0, nq, dims).copy()
    start = 12345
    rows[start:start + 120] = Q[0][None, :] + 0.01 * rng.standard_normal((120, dims)).astype(np.float32)
    rows16 = onp.round_dev(rows, 1)
    with T.GpuVectorIndex(dims, 0, capacity_rows=n, dev_dtype=1, k_max=128, nq_max=16) as ix:
        ix.append_rows(rows)
        ix.stats_reset()
        ids, dist, cnt = ix.search(Q, k)
        st = ix.stats()
        assert st.last_path == 2
        assert st.retried_queries >= 1 and st.uncertified_queries == 0
        for q in (0, 1, 7):
            oi, od = oracle.search(rows16, Q[q], 0, k)
            assert cnt[q] == k and (ids[q] == oi).all(), (q, ids[q][:8], oi[:8])
            assert (bits(dist[q]) == bits(od)).all()
