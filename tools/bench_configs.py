"""Per-config measurements (BASELINE.json configs 1-5, one GPU's share of the
sharded ones). Prints one JSON line per config. Usage:
    python tools/bench_configs.py [c1 c2 c3 c4 c5]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from tostore_b200 import GpuVectorIndex  # noqa: E402

PEAKS = {"hbm_gbs": 6545.6, "bf16_tflops": 1622.2, "bf16_tflops_sustained": 1365.6}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def run(name, n, d, metric, dt, nq, k, reps, mask_frac=None, warm=3):
    Q = oracle.synth_rows(99, 0, nq * (reps + warm), d).reshape(reps + warm, nq, d)
    if metric == 2:
        Q = np.stack([np.stack([onp.normalize_f32(q) for q in b]) for b in Q])
    with GpuVectorIndex(d, metric, capacity_rows=n, dev_dtype=dt, k_max=max(k, 16), nq_max=max(nq, 8)) as ix:
        ix.append_synthetic(7, n)
        if mask_frac is not None:
            ix.set_filter(np.random.default_rng(9).random(n) < mask_frac)
        for i in range(warm):
            ix.search(Q[i], k)
        ix.stats_reset()
        tot = 0.0
        t0 = time.perf_counter()
        for i in range(warm, warm + reps):
            ix.search(Q[i], k)
            tot += ix.stats().last_search_ms
        wall = (time.perf_counter() - t0) / reps * 1e3
        st = ix.stats()
        hot_ms = st.hot_ms_total / st.hot_launches
        per_search_launches = st.hot_launches / reps
        out = {"config": name, "n": n, "d": d, "metric": metric, "dev_dtype": dt, "nq": nq, "k": k,
               "path": st.last_path, "device_ms_per_search": tot / reps, "e2e_ms_per_search": wall,
               "qps_device": nq / (tot / reps) * 1e3, "qps_e2e": nq / wall * 1e3,
               "hot_kernel_ms": hot_ms, "hot_launches_per_search": per_search_launches,
               "hbm_gbs": st.hot_bytes_total / st.hot_launches / hot_ms / 1e6,
               "hbm_frac_of_measured": st.hot_bytes_total / st.hot_launches / hot_ms / 1e6 / PEAKS["hbm_gbs"]}
        if st.hot_flops_total > 0:
            tf = st.hot_flops_total / st.hot_launches / hot_ms / 1e9
            out.update(tflops=tf, tensor_frac_of_measured_burst=tf / PEAKS["bf16_tflops"],
                       tensor_frac_of_measured_sustained=tf / PEAKS["bf16_tflops_sustained"])
        if mask_frac is not None:
            out["filtered_bytes_gbs"] = out["hbm_gbs"] * mask_frac
        print(json.dumps(out), flush=True)


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which:
    run("c1 brute-force L2 10k x 128 fp32 k=10", 10_000, 128, 0, 0, 1, 10, 200)
if "c2" in which:
    run("c2 single-query L2 10M x 768 fp32 k=10", 10_000_000, 768, 0, 0, 1, 10, 30)
if "c2b" in which:
    run("c2b 8-query L2 10M x 768 fp32 k=10", 10_000_000, 768, 0, 0, 8, 10, 10)
if "c3" in which:
    run("c3 batch-1024 cosine 10M x 768 bf16 k=10", 10_000_000, 768, 2, 1, 1024, 10, 10)
if "c3s" in which:
    run("c3s single-query cosine 10M x 768 bf16 k=10 (scan)", 10_000_000, 768, 2, 1, 1, 10, 20)
if "c4" in which:
    run("c4 shard: IP 12.5M x 1536 fp16 k=100 (1 of 8 GPUs)", 12_500_000, 1536, 1, 2, 1, 100, 20)
if "c5" in which:
    run("c5 shard: L2 12.5M x 384 fp32 k=10, 10% WHERE mask (1 of 4 GPUs)", 12_500_000, 384, 0, 0, 1, 10, 20,
        mask_frac=0.10)
    run("c5u shard unfiltered: L2 12.5M x 384 fp32 k=10", 12_500_000, 384, 0, 0, 1, 10, 20)
