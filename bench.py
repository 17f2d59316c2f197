#!/usr/bin/env python
"""bench.py — headline benchmark of the vector-search hot path (BASELINE.json).

Metric: vector QPS @ recall=1.0, d=768 N=10M k=10 (single-query L2 over fp32
rows, BASELINE config 2), with the dominant kernel's achieved HBM GB/s against
the measured roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W]         this repo (CUDA path)
  python bench.py --impl reference [...]                      CPU arm (oracle port)

A "step" is one single-query search over the whole corpus (one scan pass +
select/re-rank). N>1 (torchrun, one rank per GPU): the same 10M-row corpus is
row-range sharded across ranks (strong scaling); every step each rank scans its
shard, the per-shard exact top-k are exchanged with one ncclAllGather inside the
library and merged on every rank. `--mode replicas` (not the default) instead gives every
GPU a full copy of the corpus and its own stream of queries: no exchange, weak scaling.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC_NAME = "vector QPS @ recall=1.0, d=768 N=10M k=10; achieved GB/s vs HBM roofline"
SEED = 0x705702E2
N_ROWS, DIMS, K = 10_000_000, 768, 10
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="corpus rows (default = the metric's 10M)")
    ap.add_argument("--dims", type=int, default=DIMS)
    ap.add_argument("--k", type=int, default=K)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the `configs` block (BASELINE configs 1/3/4/5 measured after the headline)")
    ap.add_argument("--recall-queries", type=int, default=3,
                    help="full oracle searches (all N rows, every host core) run after the timed regions")
    ap.add_argument("--mode", default="shard", choices=["shard", "replicas"],
                    help="N>1: row-range shards of ONE corpus with a top-k exchange per query (default, "
                         "what BASELINE's north_star asks for) or N independent full replicas, each "
                         "serving its own queries (no exchange; the corpus fits one GPU)")
    ap.add_argument("--diag-lib", action="store_true", help="tools only: load libtostore_cuda_diag.so")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="device loop without tsc_index_set_pipelining (every search then waits for the "
                         "previous one's tail and carries timer events)")
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "p2p"],
                    help="N>1: push over NVLink peer memory fused into the scan kernel (default) or "
                         "ncclAllGather + merge kernel")
    return ap.parse_args()


def ncu_traffic(n, d, world):
    """dram__bytes_read+write of the dominant kernel from the committed `ncu --set full`
    capture (profiles/), valid only for the exact workload it was taken on."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_scan_traffic.json")) as f:
            t = json.load(f)
        if world == 1 and n == 10_000_000 and d == 768:
            return float(t["traffic"]), t["source"]
    except Exception:
        pass
    return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# ---------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm),
                "power_w_max": max(power), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------
# CPU arm: the oracle port (restated reference arithmetic), all host threads
# ---------------------------------------------------------------------------------
def cpu_arm(rows_total: int, dims: int, k: int, sample_rows: int, steps: int, warmup: int):
    """Exhaustive fp64 scan with the reference's own arithmetic over a bounded
    sample of the corpus; QPS is scaled to the full corpus (linear in rows)."""
    import numpy as np
    import oracle

    oracle.c_oracle()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    sample_rows = min(sample_rows, rows_total)
    rows = oracle.synth_rows(SEED, 0, sample_rows, dims)
    queries = oracle.synth_rows(SEED + 1, 0, steps + warmup, dims)
    for i in range(warmup):
        oracle.search(rows, queries[i], 0, k, threads=threads)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        oracle.search(rows, queries[i], 0, k, threads=threads)
    dt = time.perf_counter() - t0
    scale = sample_rows / rows_total
    qps = steps / dt * scale
    import shutil
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
            "dart_sdk_on_this_host": bool(shutil.which("dart") or shutil.which("flutter")),
            "sample": f"{steps} queries x {sample_rows} of {rows_total} rows (d={dims}, fp64 scalar "
                      f"_exactDistance + heap top-k, OpenMP row split), QPS scaled x{scale:g} to the full corpus",
            "ms_per_step_sample": dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 30)
    base = cpu_arm(args.rows, args.dims, args.k, args.cpu_sample_rows, steps, min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": base["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
        "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"single-query L2, N={args.rows} d={args.dims} fp32, k={args.k} "
                               "(BASELINE config 2)",
                   "implementation": "reference arithmetic (_exactDistance, fp64, sequential) restated in C "
                                     "(oracle port, exhaustive scan, OpenMP over rows); the Dart reference "
                                     "itself cannot run here (no Dart SDK)",
                   "l2_flush": "inputs larger than L2"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch

    import oracle  # only for the cpu_baseline leg and the recall check (after the timed regions)
    if args.diag_lib:   # tools only: experiments with the diagnostics build's switches
        from tostore_b200 import _native
        _native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "libtostore_cuda_diag.so")
    from tostore_b200 import METRIC_L2, GpuVectorIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, d, k = args.rows, args.dims, args.k
    steps, warmup = args.steps, max(args.warmup, 3)
    # row-range sharding: contiguous node-id ranges, aligned to 32 rows
    replicas = world > 1 and args.mode == "replicas"
    sharded = world > 1 and not replicas
    per = ((n + world - 1) // world + 31) // 32 * 32
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    if replicas:
        lo, hi = 0, n
    ix = GpuVectorIndex(d, METRIC_L2, capacity_rows=max(hi - lo, 1), device_id=local,
                        first_node_id=lo, k_max=16, nq_max=8)
    ix.append_synthetic(SEED, hi - lo, first_node_id=lo)
    if sharded and args.exchange == "p2p":
        # push exchange over NVLink peer memory, fused into the scan kernel's last CTA; the
        # merged result is needed where the host reads it: rank 0 (the other ranks only push)
        ix.comm_init_p2p(dist, world, rank, root=0)
    elif sharded:
        uid = [GpuVectorIndex.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ix.comm_init(uid[0], world, rank)

    nqs = steps + warmup
    # replicas serve different queries; shards of one corpus all see the same query
    q_np = oracle.synth_rows(SEED + 1, (rank * nqs) if replicas else 0, nqs, d)
    q_host = torch.from_numpy(q_np).pin_memory()
    q_dev = q_host.cuda()
    o_ids = torch.full((nqs, k), -1, dtype=torch.int64, device="cuda")
    o_dist = torch.empty((nqs, k), dtype=torch.float64, device="cuda")
    o_cnt = torch.zeros((nqs,), dtype=torch.int32, device="cuda")
    # a real (non-default) stream: every kernel of the timed region is launched on
    # it and the CUDA events that time the region are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    esz = q_dev.element_size() * d

    def step_device(i):
        ix.search_device(q_dev.data_ptr() + i * esz, 1, k, o_ids.data_ptr() + i * k * 8,
                         o_dist.data_ptr() + i * k * 8, o_cnt.data_ptr() + i * 4,
                         stream=sptr, sharded=sharded)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (`value`) -----------------------------------------
    # Throughput mode of the public device API: the HBM pass of query i+1 overlaps the tail
    # (selection, fp64 re-rank, certificate, shard exchange) of query i; all work of every
    # query is still done and results are complete in stream order.
    pipelined = not args.no_pipeline
    ix.set_pipelining(pipelined)
    for i in range(warmup):
        step_device(i)
    sync_all()
    ix.stats_reset()
    launches0 = ix.stats().kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for i in range(warmup, warmup + steps):
        step_device(i)
    e1.record(stream)
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    st = ix.stats()
    launches = st.kernel_launches - launches0
    hot_ms = st.hot_ms_total / max(st.hot_launches, 1)
    hot_bytes = st.hot_bytes_total / max(st.hot_launches, 1)
    dev_ids = o_ids.cpu().numpy()
    dev_dist = o_dist.cpu().numpy()
    ix.set_pipelining(False)

    # ---- end to end through the host-buffer plugin call (`e2e`) ----------------------
    # tsc_search with HOST buffers on every rank: H2D of the query, the kernels (for a
    # sharded index: + the exchange over NVLink), D2H of the result, all inside the call.
    e2e_out = [None]

    def step_e2e(i):
        e2e_out[0] = ix.search(q_np[i], k)

    for i in range(warmup):
        step_e2e(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        step_e2e(i)
    wall_ms = (time.perf_counter() - t0) * 1e3
    sync_all()
    ms2 = torch.tensor([wall_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())
    e2e_last = e2e_out[0]
    cert = ix.stats()
    flags_bad = int((ix.search_flags(1) != 0).sum())

    # every rank's certificate counters (a query is exact when every shard certified it)
    cert_t = torch.tensor([cert.certified_queries, cert.retried_queries, cert.uncertified_queries],
                          dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(cert_t, op=dist.ReduceOp.SUM)
    cert_sum = [int(x) for x in cert_t.cpu()]
    ix.close()
    del q_dev, o_ids, o_dist, o_cnt
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()   # nothing CPU-heavy happens inside a live process group
    if rank != 0:
        return

    # ---- recall check on rank 0, after the timed regions, with every host core ------------
    # FULL oracle searches (the C restatement of the reference arithmetic streams all N
    # synthetic rows, ~5 s per query on 16 cores) for the last timed queries of both loops:
    # ids identical, fp64 distances bit-identical = recall 1.0 for those queries, at any N.
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    checked, ok_ids, ok_bits = [], True, True
    last = warmup + steps - 1
    for which, qi, got_ids, got_dist in (("device loop", last, dev_ids[last], dev_dist[last]),
                                         ("host-buffer loop", last, e2e_last[0][0], e2e_last[1][0]),
                                         ("device loop", last - 1, dev_ids[last - 1],
                                          dev_dist[last - 1]))[:max(args.recall_queries, 0)]:
        oi, od = oracle.search_synth(SEED, n, d, 0, q_np[qi], 0, k, threads=threads)
        ok_ids &= bool((np.asarray(got_ids) == oi).all())
        ok_bits &= bool((np.asarray(got_dist, dtype=np.float64).view(np.int64) == od.view(np.int64)).all())
        checked.append(f"{which} query {qi}")
    recall = {"checked": f"full oracle search over all {n} rows for: " + ", ".join(checked),
              "ids_identical": ok_ids, "bit_exact": ok_bits, "recall": 1.0 if ok_ids else None,
              "certificate": {"certified": cert_sum[0], "range_pass": cert_sum[1],
                              "uncertified": cert_sum[2],
                              "note": "per shard and query over both timed loops + warm-up of the "
                                      "second; uncertified must be 0 for recall=1.0 to be proven"},
              "flags_nonzero_last_search": flags_bad}

    peak, which = peaks()
    achieved = hot_bytes / (hot_ms * 1e-3) / 1e9 if hot_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic(n, d, world)
    xname = ("none" if not sharded else
             "push over NVLink peer memory to rank 0, fused into the scan kernel's last CTA"
             if args.exchange == "p2p" else "ncclAllGather of per-shard top-k + merge kernel (in-library)")
    line = {
        # replicas: every rank answers one query per step
        "metric": METRIC_NAME, "value": steps * (world if replicas else 1) / (total_ms * 1e-3),
        "unit": "queries/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "weak" if replicas else "strong", "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"single-query L2, N={n} d={d} fp32, k={k} (BASELINE config 2)",
                   "implementation": "one kernel per query: HBM-bound scan + in-kernel top-k + exact fp64 "
                                     f"re-rank + exactness certificate (+ shard exchange) on {world}xB200",
                   "rows_per_gpu": hi - lo,
                   "sharding": "replicas (each GPU holds the whole corpus and serves its own queries)"
                   if replicas else ("row-range" if world > 1 else "none"),
                   "exchange": xname,
                   "l2_flush": "inputs larger than L2 (each pass streams the whole shard; "
                               f"{(hi - lo) * d * 4 / 1e9:.2f} GB per GPU vs 126 MB L2)",
                   "queries": "distinct synthetic query per step",
                   "device_loop": ("pipelined (tsc_index_set_pipelining): the scan of query i+1 overlaps the "
                                   "tail of query i; the hot-kernel timer samples every 16th search"
                                   if pipelined else "one search after the other, every search timed")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)" if which == "measured"
                     else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "peak_note": "MEASURED_PEAKS hbm_gbs is a torch copy (read + write) figure; a "
                                  "read-only stream pays no write turnarounds, so frac can exceed 1 "
                                  "(DESIGN.md K1: 90 % of the 8.18 TB/s ncu reports as DRAM peak)",
                     "kernel": "scan_topk_kernel<L2,f32,QB=1> (scan + fused tail)",
                     "kernel_ms": hot_ms, "algorithmic_bytes_per_launch": hot_bytes,
                     "kernel_share_of_step": hot_ms / (total_ms / steps) if total_ms else None,
                     "kernel_share_note": ("kernel_ms is the event-timed duration of the sampled, "
                                           "non-overlapped launches; in the pipelined loop consecutive "
                                           "launches overlap by the tail, so the share can exceed 1")
                     if pipelined else None},
        "e2e": {"value": steps * (world if replicas else 1) / (e2e_ms * 1e-3), "unit": "queries/s",
                "h2d_bytes_per_step": d * 4 * world,            # every rank copies its query in
                # the library brings ids | dist | counts | flags back as ONE block sized for
                # (nq_max = 8, k_max = 16): 2 * 8 * 16 * 8 + 2 * 32 bytes, on every rank
                "d2h_bytes_per_step": (2 * 8 * 16 * 8 + 2 * 32) * world,
                "ms_per_step": e2e_ms / steps,
                "api": "tsc_search (host buffers)" + (", called on every rank; rank 0 receives the "
                                                      "merged result" if sharded else "")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "recall_check": recall,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_arm(n, d, k, args.cpu_sample_rows, 8, 1)
    if world == 1 and not args.no_configs:
        # the other BASELINE configs (one GPU's share of the sharded ones), measured after the
        # headline region: device ms, achieved GB/s or TFLOP/s and fraction of the measured peak
        try:
            from tools.bench_configs import measure_all
            line["configs"] = measure_all(("c1", "c3", "c4", "c5", "c5w", "c5t"))
        except Exception as e:  # noqa: BLE001
            line["configs"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
