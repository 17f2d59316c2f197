"""Multi-GPU test (needs >= 2 B200s; skipped otherwise): in-library sharded search
(local scan -> ncclAllGather of per-shard top-k -> merge kernel) equals the oracle
on the whole corpus, on every rank."""
import os
import socket

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret, exchange="nccl"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from tostore_b200 import GpuVectorIndex
        from tostore_b200.sharding import shard_rows
        n, d, k, nq = 200_000, 256, 10, 6
        lo, hi = shard_rows(n, world, rank)
        ix = GpuVectorIndex(d, 2, capacity_rows=hi - lo, device_id=rank, first_node_id=lo, k_max=16, nq_max=8)
        ix.append_synthetic(21, hi - lo, first_node_id=lo)
        if exchange == "p2p":
            ix.comm_init_p2p(dist, world, rank)
        else:
            uid = [GpuVectorIndex.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            ix.comm_init(uid[0], world, rank)
        from oracle import oracle_np as onp
        Q = np.stack([onp.normalize_f32(q) for q in oracle.synth_rows(22, 0, nq, d)])
        dq = torch.from_numpy(Q).cuda()
        o_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        o_dist = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        o_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        for _ in range(3):
            ix.search_device(dq.data_ptr(), nq, k, o_ids.data_ptr(), o_dist.data_ptr(), o_cnt.data_ptr(),
                             stream=s.cuda_stream, sharded=True)
        s.synchronize()
        ret[rank] = (o_ids.cpu().numpy(), o_dist.cpu().numpy(), o_cnt.cpu().numpy())
        ix.close()
    finally:
        dist.destroy_process_group()


# exchange="p2p": push over NVLink peer memory fused into the scan kernel's tail
# (tsc_exchange.cuh); "nccl": ncclAllGather + merge kernel
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_search_nccl_all_gather(exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle_np as onp
    world, port = min(torch.cuda.device_count(), 8), _free_port()   # 2, 4 or 8 shards
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret, exchange), nprocs=world, join=True)
        n, d, k, nq = 200_000, 256, 10, 6
        rows = oracle.synth_rows(21, 0, n, d)
        Q = np.stack([onp.normalize_f32(q) for q in oracle.synth_rows(22, 0, nq, d)])
        for r in range(world):
            ids, dist_, cnt = ret[r]
            for qi in range(nq):
                oi, od = oracle.search(rows, Q[qi], 2, k, threads=4)
                assert cnt[qi] == k and (ids[qi] == oi).all(), (r, qi, ids[qi], oi)
                assert (dist_[qi].view(np.int64) == od.view(np.int64)).all()
