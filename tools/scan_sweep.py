"""Sweep the scan kernel's stage geometry on the C2 workload (one process, one
index per config; geometry is read from TSC_SCAN_* when the index is created)."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from tostore_b200 import GpuVectorIndex  # noqa: E402

N, D, K = int(os.environ.get("SWEEP_ROWS", 10_000_000)), int(os.environ.get("SWEEP_DIMS", 768)), 10
DT = int(os.environ.get("SWEEP_DTYPE", 0))
METRIC = int(os.environ.get("SWEEP_METRIC", 0))
q = oracle.synth_rows(5, 0, 32, D)
configs = []
for ctas, warps, stages, rows in itertools.product((1, 2), (4, 8, 12, 16), (2, 3, 4, 6), (1, 2, 4)):
    configs.append((ctas, warps, stages, rows))
only = os.environ.get("SWEEP_ONLY")
for ctas, warps, stages, rows in configs:
    os.environ.update(TSC_SCAN_CTAS=str(ctas), TSC_SCAN_WARPS=str(warps),
                      TSC_SCAN_STAGES=str(stages), TSC_SCAN_ROWS=str(rows))
    try:
        with GpuVectorIndex(D, METRIC, capacity_rows=N, dev_dtype=DT, k_max=16, nq_max=8) as ix:
            ix.append_synthetic(1, N)
            for i in range(3):
                ix.search(q[i], K)
            ix.stats_reset()
            for i in range(3, 19):
                ix.search(q[i], K)
            st = ix.stats()
            ms = st.hot_ms_total / st.hot_launches
            gbs = st.hot_bytes_total / st.hot_launches / ms / 1e6
            inflight = warps * stages * rows * st.row_stride_bytes * ctas / 1024
            print(f"ctas={ctas} warps={warps:2d} stages={stages} rows={rows} inflight={inflight:6.0f}KB "
                  f"scan_ms={ms:.3f} GB/s={gbs:.0f} total_ms={st.last_search_ms:.3f}", flush=True)
    except Exception as e:  # config does not fit shared memory
        print(f"ctas={ctas} warps={warps} stages={stages} rows={rows} -> {str(e)[:90]}", flush=True)
