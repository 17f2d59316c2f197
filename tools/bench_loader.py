"""Cold-start throughput of the native NGH loader (tsc_index_load_ngh).

    python tools/bench_loader.py [--rows 1000000] [--dims 768] [--dir /tmp/ngh_bench] [--dry]

Writes a synthetic on-disk index in the reference's format (oracle_np.write_ngh_index: real
page envelopes, CRCs, partition files), then loads it into a GPU index with the native
loader and, for comparison, with the Python walk. --dry (no GPU) only times the library's
directory walk + double-buffered reader with a CRC sink (tsc_selftest_ngh_walk)."""
import argparse
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import oracle_np as onp  # noqa: E402  (test infrastructure: writes the fixture)
from tostore_b200 import ngh_loader as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dims", type=int, default=768)
    ap.add_argument("--dir", default="/tmp/ngh_bench")
    ap.add_argument("--dry", action="store_true")
    args = ap.parse_args()

    if not os.path.exists(os.path.join(args.dir, "ngh", "meta.json")):
        shutil.rmtree(args.dir, ignore_errors=True)
        os.makedirs(args.dir)
        rng = np.random.default_rng(1)
        rows = rng.standard_normal((args.rows, args.dims)).astype(np.float32)
        t0 = time.perf_counter()
        onp.write_ngh_index(args.dir, rows, "l2", onp.F32)
        print(f"# wrote {args.rows} x {args.dims} in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    meta = L.read_meta_native(args.dir)
    out = {"rows": meta.next_node_id, "dims": meta.dimensions}

    if args.dry:
        import ctypes as C
        from tostore_b200 import _native as N
        cap = 1 << 16
        first = np.zeros(cap, dtype=np.uint64)
        cnt = np.zeros(cap, dtype=np.uint64)
        crc = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        t0 = time.perf_counter()
        N.check(N.lib().tsc_selftest_ngh_walk(args.dir.encode(), 0, 0, meta.next_node_id, 4096,
                                              first.ctypes.data, cnt.ctypes.data, crc.ctypes.data, cap,
                                              C.byref(n)), "walk")
        dt = time.perf_counter() - t0
        nbytes = int(cnt[: n.value].sum()) * meta.page_size
        out.update(mode="dry (host CRC sink)", chunks=n.value, bytes=nbytes, seconds=dt, gbs=nbytes / dt / 1e9)
        print(json.dumps(out))
        return

    import tostore_b200 as T

    def make(m):
        return T.GpuVectorIndex(m.dimensions, m.metric, capacity_rows=m.next_node_id, src_precision=m.precision,
                                k_max=16, nq_max=4)

    for native in (True, False, True):
        t0 = time.perf_counter()
        ix, _ = L.load_ngh_index(args.dir, make, native=native)
        dt = time.perf_counter() - t0
        st = ix.stats()
        nbytes = st.rows * meta.dimensions * 4
        print(json.dumps(dict(out, loader="native" if native else "python", seconds=dt, rows_loaded=st.rows,
                              row_gbs=nbytes / dt / 1e9)))
        ix.close()


if __name__ == "__main__":
    main()
