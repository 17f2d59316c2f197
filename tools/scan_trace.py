"""Where does a scan launch spend its time? Needs the diagnostics library
(`make -C tostore_b200/csrc diag`): per-CTA main-loop end times and the phases of the fused
tail, from globaltimer stamps (tsc_tail.cuh TSC_TRACE).
    python tools/scan_trace.py [rows] [dims] [k] [reps] [metric] [dev_dtype]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tostore_b200 import _native  # noqa: E402

_native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "libtostore_cuda_diag.so")
import oracle  # noqa: E402
from tostore_b200 import GpuVectorIndex  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
metric = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dt = int(sys.argv[6]) if len(sys.argv) > 6 else 0
L = _native.lib()
L.tsc_diag_scan_trace.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_int32]
Q = oracle.synth_rows(5, 0, reps + 3, d)
with GpuVectorIndex(d, metric, capacity_rows=n, dev_dtype=dt, k_max=max(16, k), nq_max=8) as ix:
    ix.append_synthetic(7, n)
    grid = 148
    buf = np.zeros(16 + 2 * grid, dtype=np.uint64)
    for i in range(reps + 3):
        L.tsc_diag_scan_trace(ix.handle, buf.ctypes.data, buf.size, 1)
        ix.search(Q[i], k)
        L.tsc_diag_scan_trace(ix.handle, buf.ctypes.data, buf.size, 0)
        if i < 3:
            continue
        t0 = int(buf[0])
        loop = (buf[16:16 + grid].astype(np.int64) - t0) / 1e3
        pub = (buf[16 + grid:16 + 2 * grid].astype(np.int64) - t0) / 1e3
        ph = [(int(x) - t0) / 1e3 for x in buf[1:7]]
        print(f"rep {i}: main loop ends us min/med/max {loop.min():.1f}/{np.median(loop):.1f}/{loop.max():.1f}"
              f" | published max {pub.max():.1f} | tail begin {ph[0]:.1f} selected {ph[1]:.1f} chains {ph[2]:.1f}"
              f" sorted {ph[3]:.1f} cert {ph[4]:.1f} emitted {ph[5]:.1f} | kernel(event) {ix.stats().last_search_ms * 1e3:.1f}")
