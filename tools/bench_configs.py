"""Per-config measurements (BASELINE.json configs 1-5, one GPU's share of the sharded ones).
Importable (bench.py appends `measure_all()` to its JSON line as the `configs` block) and a
command line:  python tools/bench_configs.py [c1 c2 c3 c4 c5 ...]  -> one JSON line per config."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

if "--diag" in sys.argv:   # tools only: the diagnostics build honours the TSC_SCAN_* / TSC_GEMM_* switches
    sys.argv.remove("--diag")
    from tostore_b200 import _native  # noqa: E402
    _native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "libtostore_cuda_diag.so")

PEAKS = {"hbm_gbs": 6545.6, "bf16_tflops": 1622.2, "bf16_tflops_sustained": 1365.6}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def measure(name, n, d, metric, dt, nq, k, reps, mask_frac=None, warm=3, where=False, device_id=0):
    """One config on one GPU: `reps` searches through tsc_search (host buffers), device time
    from the library's CUDA events, dominant-kernel time from its per-launch event pairs.
    where: produce the mask on the GPU from two attribute columns (tsc_index_filter_where)
    instead of uploading a bitmap, and time that call."""
    import oracle
    from oracle import oracle_np as onp
    from tostore_b200 import GpuVectorIndex
    Q = oracle.synth_rows(99, 0, nq * (reps + warm), d).reshape(reps + warm, nq, d)
    if metric == 2:
        Q = np.stack([np.stack([onp.normalize_f32(q) for q in b]) for b in Q])
    with GpuVectorIndex(d, metric, capacity_rows=n, dev_dtype=dt, k_max=max(k, 16), nq_max=max(nq, 8),
                        device_id=device_id) as ix:
        ix.append_synthetic(7, n)
        where_ms = None
        if mask_frac is not None and where:
            from tostore_b200 import where as W
            rng = np.random.default_rng(9)
            ix.column_create(0, W.COL_I64)
            ix.column_create(1, W.COL_F64)
            ix.column_append(0, rng.integers(0, 1000, n))
            ix.column_append(1, rng.random(n))
            # price < 1000 * sqrt(frac) AND rating < sqrt(frac)  ->  selectivity ~ frac
            cut = mask_frac ** 0.5
            prog = W.compile_condition({"AND": [{"price": {"<": int(1000 * cut)}}, {"rating": {"<": cut}}]},
                                       {"price": (0, W.COL_I64), "rating": (1, W.COL_F64)})
            ix.filter_where(prog)
            t0 = time.perf_counter()
            for _ in range(20):
                matched = ix.filter_where(prog)
            where_ms = (time.perf_counter() - t0) / 20 * 1e3
        elif mask_frac is not None:
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
        out = {"config": name, "n": n, "d": d, "metric": metric, "dev_dtype": dt, "nq": nq, "k": k,
               "path": "tensor" if st.last_path == 2 else "scan",
               "device_ms_per_search": tot / reps, "e2e_ms_per_search": wall,
               "qps_device": nq / (tot / reps) * 1e3, "qps_e2e": nq / wall * 1e3,
               "hot_kernel_ms": hot_ms, "hot_launches_per_search": st.hot_launches / reps,
               "hbm_gbs": st.hot_bytes_total / st.hot_launches / hot_ms / 1e6,
               "hbm_frac": st.hot_bytes_total / st.hot_launches / hot_ms / 1e6 / PEAKS["hbm_gbs"],
               "certified": int(st.certified_queries), "retried": int(st.retried_queries),
               "uncertified": int(st.uncertified_queries)}
        if st.hot_flops_total > 0:
            tf = st.hot_flops_total / st.hot_launches / hot_ms / 1e9
            out.update(tflops=tf, tensor_frac_burst=tf / PEAKS["bf16_tflops"],
                       tensor_frac_sustained=tf / PEAKS["bf16_tflops_sustained"])
        if mask_frac is not None:
            out["full_scan_equivalent_gbs"] = out["hbm_gbs"] / mask_frac
        if where_ms is not None:
            # two 8-byte columns read + one bit per row written (tsc_where.cuh)
            out.update(where_ms_incl_sync=where_ms, where_matched=int(matched),
                       where_gbs=n * 16.125 / where_ms / 1e6)
        return out


def measure_where_text(name, n, distinct, reps=20, device_id=0):
    """WHERE over a TEXT column on one GPU: `category LIKE 'cat-1%' AND price < 500` over n rows
    whose category is one of `distinct` strings ("cat-" + 7 digits). Times tsc_index_filter_where_text
    (dictionary pass + row pass + count read-back, wall clock incl. the sync) and the column append
    (host interning + upload)."""
    import ctypes as C
    from tostore_b200 import GpuVectorIndex, _native as N, where as W
    rng = np.random.default_rng(11)
    with GpuVectorIndex(16, 0, capacity_rows=n, k_max=16, nq_max=8, device_id=device_id) as ix:
        ix.append_synthetic(7, n)
        ix.column_create(0, W.COL_I64)
        ix.column_create(1, W.COL_TEXT)
        ix.column_append(0, rng.integers(0, 1000, n))
        codes = rng.integers(0, distinct, n)
        # "cat-" + zero-padded number spread over [0, 10^7): fixed 11 code units per row, built with numpy
        num = (codes * (10_000_000 // distinct)).astype(np.int64)
        units = np.empty((n, 11), dtype=np.uint16)
        units[:, :4] = np.frombuffer("cat-".encode("utf-16-le"), dtype=np.uint16)
        for j in range(7):
            units[:, 10 - j] = 48 + (num // 10 ** j) % 10
        offsets = np.arange(n + 1, dtype=np.uint64) * 11
        t0 = time.perf_counter()
        N.check(ix._lib.tsc_index_column_append_text(ix.handle, 1, 0, units.ctypes.data, offsets.ctypes.data,
                                                     None, n), "tsc_index_column_append_text")
        append_s = time.perf_counter() - t0
        prog = W.compile_condition({"AND": [{"category": {"LIKE": "cat-1%"}}, {"price": {"<": 500}}]},
                                   {"price": (0, W.COL_I64), "category": (1, W.COL_TEXT)})
        matched = ix.filter_where(prog)
        t0 = time.perf_counter()
        for _ in range(reps):
            matched = ix.filter_where(prog)
        ms = (time.perf_counter() - t0) / reps * 1e3
        want = int(((num // 1_000_000 == 1) & (np.random.default_rng(11).integers(0, 1000, n) < 500)).sum())
        n_distinct = int(np.unique(codes).size)
        # row pass: two 8-byte columns + one bit per row; dictionary pass: the distinct strings once
        bytes_moved = n * 16.125 + n_distinct * (11 * 2 + 8)
        return {"config": name, "n": n, "distinct_strings": n_distinct, "where_ms_incl_sync": ms,
                "where_matched": int(matched), "where_matched_expected": want, "where_gbs": bytes_moved / ms / 1e6,
                "column_append_s": append_s, "column_append_mrows_s": n / append_s / 1e6}


CONFIGS = {
    "c1": lambda: [measure("c1 brute-force L2 10k x 128 fp32 k=10", 10_000, 128, 0, 0, 1, 10, 200)],
    "c2": lambda: [measure("c2 single-query L2 10M x 768 fp32 k=10", 10_000_000, 768, 0, 0, 1, 10, 30)],
    "c2b": lambda: [measure("c2b 8-query L2 10M x 768 fp32 k=10 (scan, one pass)", 10_000_000, 768, 0, 0, 8, 10, 10)],
    "c2q": lambda: [measure(f"c2q {q}-query L2 10M x 768 fp32 k=10", 10_000_000, 768, 0, 0, q, 10, 8) for q in (2, 4, 8)],
    "c3q": lambda: [measure(f"c3q {q}-query cosine 10M x 768 bf16 k=10", 10_000_000, 768, 2, 1, q, 10, 8) for q in (2, 4, 8)],
    "c2t": lambda: [measure("c2t batch-1024 L2 10M x 768 fp32 k=10 (tf32 tensor path)", 10_000_000, 768, 0, 0,
                            1024, 10, 5)],
    "c3": lambda: [measure("c3 batch-1024 cosine 10M x 768 bf16 k=10", 10_000_000, 768, 2, 1, 1024, 10, 10)],
    "c3k": lambda: [measure("c3k batch-1024 cosine 10M x 768 bf16 k=100 (tensor path, truncated lists)", 10_000_000,
                            768, 2, 1, 1024, 100, 5)],
    "c3s": lambda: [measure("c3s single-query cosine 10M x 768 bf16 k=10 (scan)", 10_000_000, 768, 2, 1, 1, 10, 20)],
    "c4": lambda: [measure("c4 shard: IP 12.5M x 1536 fp16 k=100 (1 of 8 GPUs)", 12_500_000, 1536, 1, 2, 1, 100, 20)],
    "c5": lambda: [measure("c5 shard: L2 12.5M x 384 fp32 k=10, 10% WHERE mask (1 of 4 GPUs)", 12_500_000, 384,
                           0, 0, 1, 10, 20, mask_frac=0.10),
                   measure("c5u shard unfiltered: L2 12.5M x 384 fp32 k=10", 12_500_000, 384, 0, 0, 1, 10, 20)],
    "c5s1": lambda: [measure("c5s1 shard: 1% mask", 12_500_000, 384, 0, 0, 1, 10, 20, mask_frac=0.01)],
    "c5w": lambda: [measure("c5w shard: L2 12.5M x 384 fp32 k=10, WHERE price<316 AND rating<0.316 (~10%) "
                            "evaluated on the GPU", 12_500_000, 384, 0, 0, 1, 10, 20, mask_frac=0.10, where=True)],
    "c5t": lambda: [measure_where_text("c5t WHERE category LIKE 'cat-1%' AND price < 500 over 12.5M rows, "
                                       "1000 distinct strings (text column, dictionary-encoded)", 12_500_000, 1000),
                    measure_where_text("c5t same, 1M distinct strings", 12_500_000, 1_000_000, reps=10)],
}


def measure_all(which=("c1", "c3", "c4", "c5")):
    """One failing config does not take the others' numbers with it."""
    out = []
    for w in which:
        try:
            out.extend(CONFIGS[w]())
        except Exception as e:     # noqa: BLE001
            out.append({"config": w, "error": repr(e)})
    return out


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]):
        for line in CONFIGS[w]():
            print(json.dumps(line), flush=True)
