"""Single-process multi-GPU: ONE handle over all GPUs of the box (tsc_index_create with
n_devices > 1), searched through tsc_search with host buffers — the call the Dart host makes
(core/vector_index_manager.dart:538-548). Prints one JSON line: end-to-end QPS of BASELINE
config 2 sharded over N GPUs, with a full-corpus oracle check of the last query.
    python tools/bench_group.py [n_gpus] [rows] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tostore_b200 import METRIC_L2, GpuVectorIndex, _native  # noqa: E402

SEED = 0x705702E2
ndev = _native.lib().tsc_device_count()
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else ndev
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
d, k, warm = 768, 10, 10
Q = oracle.synth_rows(SEED + 1, 0, steps + warm, d)
kw = {"device_ids": list(range(n_gpus))} if n_gpus > 1 else {}
with GpuVectorIndex(d, METRIC_L2, capacity_rows=rows, k_max=16, nq_max=8, **kw) as ix:
    ix.append_synthetic(SEED, rows)
    for i in range(warm):
        ix.search(Q[i], k)
    ix.stats_reset()
    lat = []
    t0 = time.perf_counter()
    for i in range(warm, warm + steps):
        t1 = time.perf_counter()
        ids, dist, cnt = ix.search(Q[i], k)
        lat.append(time.perf_counter() - t1)
    wall = time.perf_counter() - t0
    st = ix.stats()
    flags = ix.search_flags(1)
oi, od = oracle.search_synth(SEED, rows, d, 0, Q[warm + steps - 1], 0, k, threads=len(os.sched_getaffinity(0)))
lat = np.sort(np.array(lat)) * 1e3
print(json.dumps({
    "what": "single-process group handle, tsc_search with host buffers", "n_gpus": n_gpus, "rows": rows,
    "dims": d, "k": k, "steps": steps, "qps_e2e": steps / wall, "ms_per_query": wall / steps * 1e3,
    "latency_ms_p50": float(lat[len(lat) // 2]), "latency_ms_p99": float(lat[int(len(lat) * 0.99)]),
    "scan_kernel_ms_slowest_shard": st.hot_ms_total / max(st.hot_launches / n_gpus, 1),
    "certified": int(st.certified_queries), "range_pass": int(st.retried_queries),
    "uncertified": int(st.uncertified_queries), "flags_last": int(flags[0]),
    "ids_identical_to_full_oracle": bool((ids[0] == oi).all()),
    "bit_exact": bool((dist[0].view(np.int64) == od.view(np.int64)).all())}))
