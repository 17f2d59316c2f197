"""A/B of the WHERE row pass's L2 prefetch distance (diagnostics build: TSC_WHERE_PF) on one
GPU, plus the text-column append rate. One index, 12.5M rows; numeric program = config c5w's,
text program = config c5t's. Prints one JSON line.
  python tools/where_ab.py            (needs `make diag`)

HISTORY: ran once (profiles/r02f_where_prefetch_ab.json). The prefetch lost at every distance
(numeric call 88 us without, 95-101 us with), so the kernel code and the TSC_WHERE_PF switch it
drove were deleted again; kept as the record of how the A/B was made (it ran from tools/)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tostore_b200 import _native  # noqa: E402

_native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "libtostore_cuda_diag.so")
from tostore_b200 import GpuVectorIndex, where as W  # noqa: E402

N_ROWS = 12_500_000


def text_units(codes, distinct):
    num = (codes * (10_000_000 // distinct)).astype(np.int64)
    units = np.empty((codes.size, 11), dtype=np.uint16)
    units[:, :4] = np.frombuffer("cat-".encode("utf-16-le"), dtype=np.uint16)
    for j in range(7):
        units[:, 10 - j] = 48 + (num // 10 ** j) % 10
    return units, np.arange(codes.size + 1, dtype=np.uint64) * 11


def main():
    n = N_ROWS
    rng = np.random.default_rng(9)
    out = {"rows": n}
    with GpuVectorIndex(16, 0, capacity_rows=n, k_max=16, nq_max=8) as ix:
        ix.append_synthetic(7, n)
        ix.column_create(0, W.COL_I64)
        ix.column_create(1, W.COL_F64)
        ix.column_append(0, rng.integers(0, 1000, n))
        ix.column_append(1, rng.random(n))
        for cid, distinct in ((2, 1000), (3, 1_000_000)):
            ix.column_create(cid, W.COL_TEXT)
            units, offs = text_units(rng.integers(0, distinct, n), distinct)
            t0 = time.perf_counter()
            _native.check(ix._lib.tsc_index_column_append_text(ix.handle, cid, 0, units.ctypes.data,
                                                               offs.ctypes.data, None, n), "append_text")
            dt = time.perf_counter() - t0
            out[f"append_text_{distinct}_distinct_s"] = dt
            out[f"append_text_{distinct}_distinct_mrows_s"] = n / dt / 1e6
        cols = {"price": (0, W.COL_I64), "rating": (1, W.COL_F64), "cat": (2, W.COL_TEXT), "cat1m": (3, W.COL_TEXT)}
        progs = {"numeric": W.compile_condition({"AND": [{"price": {"<": 316}}, {"rating": {"<": 0.316}}]}, cols),
                 "text": W.compile_condition({"AND": [{"cat": {"LIKE": "cat-1%"}}, {"price": {"<": 500}}]}, cols),
                 "text1m": W.compile_condition({"AND": [{"cat1m": {"LIKE": "cat-1%"}}, {"price": {"<": 500}}]}, cols)}
        res = {}
        for rnd in range(2):
            for pf in (0, 1, 2, 4):
                os.environ["TSC_WHERE_PF"] = str(pf)
                for name, prog in progs.items():
                    matched = ix.filter_where(prog)
                    t0 = time.perf_counter()
                    for _ in range(40):
                        ix.filter_where(prog)
                    ms = (time.perf_counter() - t0) / 40 * 1e3
                    key = f"{name}_pf{pf}"
                    res[key] = min(res.get(key, 1e9), ms)
                    res[f"{name}_matched"] = int(matched)
        out["where_ms_incl_sync"] = res
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
