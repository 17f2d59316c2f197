"""C3 (batch-1024 cosine, 10M x 768 bf16, k=10) under a matrix of tensor-core-path
settings, one index build. Every line: settings, dominant-kernel ms (CUDA events inside
the library), TFLOP/s, fraction of the measured cuBLAS peaks, SM clock / power sampled
by NVML during the timed loop. Settings are TSC_GEMM_* environment variables read by
the launcher at every search; TSC_GEMM_EXP != 0 runs an isolation experiment whose
results are INVALID (timing only).
    python tools/gemm_matrix.py "K=V,K=V" "K=V" ...       ('' = defaults)
"""
import json
import os
import sys
import threading
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

N = int(os.environ.get("C3_N", 10_000_000))
D = int(os.environ.get("C3_D", 768))
NQ = int(os.environ.get("C3_NQ", 1024))
REPS = int(os.environ.get("C3_REPS", 8))
WARM = 3


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.on = False
        self.stop = False
        self.sm, self.pw = [], []

    def run(self):
        while not self.stop:
            if self.on:
                try:
                    self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
            time.sleep(0.01)


def main():
    settings = sys.argv[1:] or [""]
    Q = oracle.synth_rows(99, 0, NQ * (REPS + WARM), D).reshape(REPS + WARM, NQ, D)
    Q = np.stack([np.stack([onp.normalize_f32(q) for q in b]) for b in Q])
    try:
        smp = Sampler()
        smp.start()
    except Exception as e:  # no NVML: still time
        print("nvml unavailable:", e, file=sys.stderr)
        smp = None
    keys = set()
    with GpuVectorIndex(D, 2, capacity_rows=N, dev_dtype=1, k_max=16, nq_max=NQ) as ix:
        ix.append_synthetic(7, N)
        ref_ids = None
        for spec in settings:
            for k in keys:
                os.environ.pop(k, None)
            kv = dict(x.split("=") for x in spec.split(",") if x)
            prof = kv.pop("PROF", None)
            for k, v in kv.items():
                os.environ["TSC_GEMM_" + k] = v
                keys.add("TSC_GEMM_" + k)
            for i in range(WARM):
                ids, _, _ = ix.search(Q[i], 10)
            valid = int(kv.get("EXP", 0)) == 0
            if valid:
                if ref_ids is None:
                    ref_ids = ids.copy()
                same = bool((ids == ref_ids).all())
            ix.stats_reset()
            if smp:
                smp.sm, smp.pw, smp.on = [], [], True
            tot = 0.0
            for i in range(WARM, WARM + REPS):
                ix.search(Q[i], 10)
                tot += ix.stats().last_search_ms
            if smp:
                smp.on = False
            st = ix.stats()
            hot = st.hot_ms_total / st.hot_launches
            tf = st.hot_flops_total / st.hot_launches / hot / 1e9
            out = {"set": spec or "default", "hot_ms": round(hot, 3), "total_ms": round(tot / REPS, 3),
                   "tflops": round(tf, 1), "frac_burst": round(tf / PEAKS["bf16_tflops"], 4),
                   "frac_sustained": round(tf / PEAKS["bf16_tflops_sustained"], 4)}
            if smp and smp.sm:
                out.update(sm_mhz=float(np.median(smp.sm)), power_w=round(float(np.mean(smp.pw)), 1),
                           samples=len(smp.sm))
            if valid:
                out["ids_same_as_first_valid"] = same
            print(json.dumps(out), flush=True)
            if prof:   # one extra launch through the profiling kernel (prints to stderr)
                os.environ["TSC_GEMM_PROF"] = "1"
                ix.search(Q[0], 10)
                os.environ.pop("TSC_GEMM_PROF", None)
    if smp:
        smp.stop = True


if __name__ == "__main__":
    main()
