"""Generate tests/golden/golden_small.npz with the numpy-float64 oracle.

The reference (tocreator/tostore, Dart) ships no golden vectors for vectorSearch
and cannot run here, so these fixtures are produced by oracle/oracle_np.py — the
independent restatement — and then pin BOTH the C oracle (CPU tests) and the
CUDA path (GPU tests). Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_np as onp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")


def main():
    out = {}
    cases = []
    # (name, seed, n, dims, dev_dtype, k)
    specs = [
        ("c1_like", 101, 4000, 128, onp.DEV_F32, 10),
        ("d768", 102, 600, 768, onp.DEV_F32, 10),
        ("ragged_d100", 103, 1500, 100, onp.DEV_F32, 7),
        ("bf16_d256", 104, 2000, 256, onp.DEV_BF16, 10),
        ("f16_d384_k100", 105, 3000, 384, onp.DEV_F16, 100),
        ("tiny_n5", 106, 5, 64, onp.DEV_F32, 10),
    ]
    for name, seed, n, dims, dt, k in specs:
        rows = onp.round_dev(onp.synth_rows(seed, 0, n, dims), dt)
        qs = onp.synth_rows(seed + 1000, 0, 3, dims)
        rng = np.random.default_rng(seed)
        deleted = rng.random(n) < 0.2
        filt = rng.random(n) < 0.3
        for metric in (onp.L2, onp.INNER_PRODUCT, onp.COSINE):
            for qi in range(qs.shape[0]):
                q = onp.normalize_f32(qs[qi]) if metric == onp.COSINE else qs[qi]
                ids, dist = onp.search(rows, q, metric, k)
                key = f"{name}/m{metric}/q{qi}"
                out[key + "/ids"], out[key + "/dist"] = ids, dist
                ids, dist = onp.search(rows, q, metric, k, deleted=deleted)
                out[key + "/del_ids"], out[key + "/del_dist"] = ids, dist
                ids, dist = onp.search(rows, q, metric, k, deleted=deleted, filter=filt)
                out[key + "/delfil_ids"], out[key + "/delfil_dist"] = ids, dist
                # threshold = distance of the 4th result: strict '>' keeps exactly 4
                full_ids, full = onp.search(rows, q, metric, k)
                if len(full) >= 4:
                    ids, dist = onp.search(rows, q, metric, k, threshold=float(full[3]))
                    out[key + "/thr"] = np.float64(full[3])
                    out[key + "/thr_ids"], out[key + "/thr_dist"] = ids, dist
        out[name + "/deleted"], out[name + "/filter"] = deleted, filt
        cases.append((name, seed, n, dims, dt, k))
    out["cases"] = np.array([c[0] for c in cases])
    out["case_params"] = np.array([[c[1], c[2], c[3], c[4], c[5]] for c in cases], dtype=np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
