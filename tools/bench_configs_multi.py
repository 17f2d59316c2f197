"""Row-sharded BASELINE configs on the GPUs of one box (torchrun, one rank per GPU):

  c4  inner-product, N=100M d=1536 fp16, k=100, 8 shards of 12.5M rows, top-k all-gather
  c5  WHERE prefilter + kNN, N=50M d=384 fp32 L2 k=10, selectivity 10%, 4 shards
  c2  the headline corpus (10M x 768 fp32 L2 k=10) row-sharded over the ranks

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29541 tools/bench_configs_multi.py c4 [--rows-per-gpu N] [--check]

Every step = local scan -> select/re-rank -> ncclAllGather of the per-shard exact
top-k (inside the library) -> merge kernel, timed with CUDA events on the launching
stream, max over ranks. Weak-scaling geometry: rows per GPU fixed by the config, so a
run on fewer GPUs than the config names measures a proportionally smaller corpus (said
in the output). --check runs the C oracle over the WHOLE sharded corpus for one query
on rank 0 (ids identical, fp64 distances bit-identical)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402  (checker only)
from tostore_b200 import GpuVectorIndex  # noqa: E402

CONFIGS = {
    # name: (rows_per_gpu, dims, metric, dev_dtype, k, mask_frac, gpus named by BASELINE)
    "c4": (12_500_000, 1536, 1, 2, 100, None, 8),
    "c5": (12_500_000, 384, 0, 0, 10, 0.10, 4),
    "c2": (None, 768, 0, 0, 10, None, 1),
}
SEED = 0x705702E4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--rows-per-gpu", type=int, default=0)
    ap.add_argument("--total-rows", type=int, default=0,
                    help="strong scaling: split this many rows over the ranks (C4: 100000000 at 2/4/8 GPUs)")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    per, d, metric, dt, k, mask_frac, named = CONFIGS[args.config]
    if per is None:
        per = (10_000_000 + world - 1) // world
    if args.total_rows:
        per = (args.total_rows + world - 1) // world
    if args.rows_per_gpu:
        per = args.rows_per_gpu
    per = (per + 31) // 32 * 32
    n_total = per * world
    lo = rank * per

    ix = GpuVectorIndex(d, metric, capacity_rows=per, dev_dtype=dt, device_id=local, first_node_id=lo,
                        k_max=max(k, 16), nq_max=8)
    ix.append_synthetic(SEED, per, first_node_id=lo)
    mask_all = None
    if mask_frac is not None:
        # stands in for a WHERE result: Bernoulli(mask_frac) per node id, same stream on every rank
        rng = np.random.default_rng(SEED + 2)
        mask_all = rng.random(n_total) < mask_frac
        ix.set_filter(mask_all[lo:lo + per])
    if world > 1 and args.exchange == "p2p":
        ix.comm_init_p2p(dist, world, rank)
    elif world > 1:
        uid = [GpuVectorIndex.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ix.comm_init(uid[0], world, rank)

    nqs = args.steps + args.warmup
    Q = oracle.synth_rows(SEED + 1, 0, nqs, d)
    q_dev = torch.from_numpy(Q).cuda()
    o_ids = torch.empty((nqs, k), dtype=torch.int64, device="cuda")
    o_dist = torch.empty((nqs, k), dtype=torch.float64, device="cuda")
    o_cnt = torch.empty((nqs,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def step(i):
        ix.search_device(q_dev.data_ptr() + i * d * 4, 1, k, o_ids.data_ptr() + i * k * 8,
                         o_dist.data_ptr() + i * k * 8, o_cnt.data_ptr() + i * 4,
                         stream=stream.cuda_stream, sharded=world > 1)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync_all()
    ix.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.warmup, nqs):
        step(i)
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    st = ix.stats()
    hot = torch.tensor([st.hot_ms_total / max(st.hot_launches, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(hot, op=dist.ReduceOp.MAX)
    step_ms = float(ms.item()) / args.steps
    hot_ms = float(hot.item())
    live_frac = 1.0 if mask_frac is None else mask_frac
    esz = 4 if dt == 0 else 2
    bytes_per_gpu = per * d * esz * live_frac

    # results of the last query to the host, then leave the process group BEFORE the oracle
    # runs: the CPU check takes seconds to minutes and must not sit inside a collective
    # (a rank waiting at a barrier for 10 min trips the NCCL watchdog)
    i = nqs - 1
    ids, dd, cnt_i = o_ids[i].cpu().numpy(), o_dist[i].cpu().numpy(), int(o_cnt[i])
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    check = None
    if args.check and rank == 0:
        t0 = time.perf_counter()
        # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
        threads = len(os.sched_getaffinity(0))
        oi, od = oracle.search_synth(SEED, n_total, d, dt, Q[i], metric, k, filter=mask_all,
                                     threads=threads)
        check = {"oracle_rows": n_total, "oracle_s": time.perf_counter() - t0,
                 "oracle_threads": threads,
                 "ids_identical": bool((ids[: len(oi)] == oi).all() and cnt_i == len(oi)),
                 "dist_bit_identical": bool((dd[: len(od)].view(np.int64) == od.view(np.int64)).all())}

    if rank == 0:
        peak = 6551.0
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        out = {"config": args.config, "gpus": world, "exchange": args.exchange, "gpus_named_by_baseline": named,
               "rows_total": n_total, "rows_per_gpu": per, "dims": d, "metric": metric, "dev_dtype": dt,
               "k": k, "mask_frac": mask_frac, "steps": args.steps, "ms_per_query": step_ms,
               "qps": 1e3 / step_ms, "scan_kernel_ms_max_over_ranks": hot_ms,
               "algorithmic_bytes_per_gpu": bytes_per_gpu,
               "scan_gbs_per_gpu": bytes_per_gpu / hot_ms / 1e6,
               "scan_frac_of_measured_hbm": bytes_per_gpu / hot_ms / 1e6 / peak,
               "exchange_and_select_ms": step_ms - hot_ms, "check": check}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
