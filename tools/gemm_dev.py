"""Diagnostic for the tcgen05 path: raw keys vs numpy on a small problem."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import oracle_np as onp  # noqa: E402
import tostore_b200 as T  # noqa: E402
from tostore_b200 import _native as N  # noqa: E402

dims, n, nq, dt = int(os.environ.get("D", 64)), int(os.environ.get("NR", 300)), int(os.environ.get("NQ", 5)), 1
rng = np.random.default_rng(0)
rows = rng.standard_normal((n, dims)).astype(np.float32)
q = rng.standard_normal((nq, dims)).astype(np.float32)
rr, qr = onp.round_dev(rows, dt).astype(np.float64), onp.round_dev(q, dt).astype(np.float64)
ref = -(qr @ rr.T)
with T.GpuVectorIndex(dims, 1, capacity_rows=n, dev_dtype=dt, k_max=16, nq_max=512) as ix:
    ix.append_rows(rows)
    out = np.empty((nq, n), dtype=np.float32)
    N.check(N.lib().tsc_debug_gemm_keys(ix.handle, q.ctypes.data, nq, out.ctypes.data), "dbg")
    err = np.abs(out - ref)
    print("max err", err.max(), "nan", np.isnan(out).sum(), "of", out.size)
    print("out[0,:8]", out[0, :8])
    print("ref[0,:8]", ref[0, :8])
    bad = np.argwhere(~(err < 1e-2 * (1 + np.abs(ref))))
    print("bad count", len(bad), "first", bad[:10].tolist())
    if len(bad):
        print("bad rows(q) uniq", np.unique(bad[:, 0])[:20], "bad cols(n) uniq", np.unique(bad[:, 1])[:40])
