"""CPU tests of the N>1 host logic: row-range plan, top-k merge, and the
all-gather exchange over torch.distributed with the gloo backend (world_size 2).
The per-shard search is stood in by the oracle here (no GPU in this container);
on B200s the same exchange runs in-library over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

import oracle
from tostore_b200.sharding import merge_topk, shard_rows


def test_shard_rows_cover_and_align():
    for n in (0, 1, 31, 32, 33, 10_000_000, 12_345_679):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 32 == 0 for lo, hi in spans if hi > lo)
    assert shard_rows(10_000_000, 8, 3) == (3 * 1_250_016, 4 * 1_250_016)
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_merge_topk_matches_global_oracle_and_order_rules():
    rows = oracle.synth_rows(5, 0, 5000, 48)
    q = oracle.synth_rows(6, 0, 3, 48)
    k = 10
    for metric in (0, 1, 2):
        parts_i, parts_d = [], []
        for lo, hi in (shard_rows(5000, 3, r) for r in range(3)):
            pi, pd = [], []
            for qi in range(3):
                i, d = oracle.search(rows[lo:hi], q[qi], metric, k, first_node_id=lo)
                pi.append(np.pad(i, (0, k - len(i)), constant_values=-1))
                pd.append(np.pad(d, (0, k - len(d)), constant_values=np.nan))
            parts_i.append(pi), parts_d.append(pd)
        ids, dist, cnt = merge_topk(np.array(parts_i), np.array(parts_d), k)
        for qi in range(3):
            oi, od = oracle.search(rows, q[qi], metric, k)
            assert cnt[qi] == k and (ids[qi] == oi).all()
            assert (dist[qi].view(np.int64) == od.view(np.int64)).all()
    # ties -> node id; -0.0 before 0.0; NaN last; empty slots skipped
    pi = np.array([[[7, 3, -1]], [[5, 9, 2]]])
    pd = np.array([[[0.0, 1.0, np.nan]], [[-0.0, 1.0, np.nan]]])
    ids, dist, cnt = merge_topk(pi, pd, 5)
    assert ids[0].tolist() == [5, 7, 3, 9, 2] and cnt[0] == 5 and np.isnan(dist[0, 4])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tostore_b200.sharding import ShardedSearcher
        n, d, k = 6000, 64, 10
        lo, hi = shard_rows(n, world, rank)
        rows = oracle.synth_rows(11, lo, hi - lo, d)            # this rank's shard only
        dead = np.zeros(hi - lo, dtype=bool)
        dead[::7] = True

        def local_search(queries, kk, thr):
            out_i = np.full((len(queries), kk), -1, dtype=np.int64)
            out_d = np.full((len(queries), kk), np.nan)
            for qi, q in enumerate(queries):
                i, dd = oracle.search(rows, q, 0, kk, threshold=thr, deleted=dead, first_node_id=lo)
                out_i[qi, : len(i)], out_d[qi, : len(i)] = i, dd
            return out_i, out_d, None

        q = oracle.synth_rows(12, 0, 4, d)
        ids, dist_, cnt = ShardedSearcher(local_search).search(q, k)
        ret[rank] = (ids, dist_, cnt)
    finally:
        dist.destroy_process_group()


def test_all_gather_merge_world_size_2_gloo():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        n, d, k = 6000, 64, 10
        rows = oracle.synth_rows(11, 0, n, d)
        dead = np.zeros(n, dtype=bool)
        for r in range(world):
            lo, hi = shard_rows(n, world, r)
            dead[lo:hi][::7] = True
        q = oracle.synth_rows(12, 0, 4, d)
        for r in range(world):                                   # every rank holds the global result
            ids, dist_, cnt = ret[r]
            for qi in range(4):
                oi, od = oracle.search(rows, q[qi], 0, k, deleted=dead)
                assert cnt[qi] == k and (ids[qi] == oi).all()
                assert (dist_[qi].view(np.int64) == od.view(np.int64)).all()
