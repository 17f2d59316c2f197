"""bench.py's output contract, checked on the CPU: the CUDA arm is run in a subprocess with the
GPU index and the torch.cuda calls replaced by inert stand-ins (fixed fake timings), purely to
exercise bench.py's own control flow and the JSON line it prints — keys, types, which rank
prints, how `value` is aggregated in each multi-GPU mode. No number produced here means
anything; the real arm needs a B200. The reference arm (`--impl reference`) runs for real."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MOCK = textwrap.dedent('''
    import contextlib, io, json, os, runpy, sys
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    import tostore_b200

    class FakeStats:
        kernel_launches = 0; hot_launches = 100; hot_ms_total = 418.0; hot_bytes_total = 100 * 30.72e9
        certified_queries = 1; retried_queries = 0; uncertified_queries = 0
    class FakeIndex:
        def __init__(self, *a, **k): pass
        def append_synthetic(self, *a, **k): pass
        def comm_init(self, *a): pass
        def comm_init_p2p(self, d, n, r, root=-1): d.barrier()
        def search_flags(self, nq): return np.zeros(nq, dtype=np.uint32)
        @staticmethod
        def comm_unique_id(): return b"x" * 128
        def search_device(self, *a, **k): FakeStats.kernel_launches += 2
        def set_pipelining(self, on): pass
        def search(self, q, k): return (np.zeros((1, k), dtype=np.int64), np.zeros((1, k)), np.ones(1, dtype=np.uint32))
        def stats(self): return FakeStats
        def stats_reset(self): pass
        def close(self): pass
    class FakeStream:
        cuda_stream = 0
        def synchronize(self): pass
    class FakeEvent:
        def __init__(self, **k): pass
        def record(self, s=None): pass
        def elapsed_time(self, other): return 422.0
    tostore_b200.GpuVectorIndex = FakeIndex
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.Stream, torch.cuda.Event, torch.cuda.set_stream = FakeStream, FakeEvent, (lambda s: None)
    torch.Tensor.pin_memory = lambda self: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    _empty, _tensor, _init = torch.empty, torch.tensor, dist.init_process_group
    _full, _zeros = torch.full, torch.zeros
    torch.cuda.empty_cache = lambda: None
    torch.full = lambda *a, **k: _full(*a, **{x: v for x, v in k.items() if x != "device"})
    torch.zeros = lambda *a, **k: _zeros(*a, **{x: v for x, v in k.items() if x != "device"})
    torch.empty = lambda *a, **k: _empty(*a, **{x: v for x, v in k.items() if x != "device"})
    torch.tensor = lambda *a, **k: _tensor(*a, **{x: v for x, v in k.items() if x != "device"})
    dist.init_process_group = lambda backend, **k: _init("gloo")
    sys.argv = ["bench.py"] + sys.argv[1:]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
    out = buf.getvalue().strip()
    if int(os.environ.get("RANK", "0")) != 0:
        assert out == "", out
    else:
        print(out.splitlines()[-1])
''')

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"}


def _run(tmp_path, args, nproc=1):
    script = tmp_path / "mock_bench.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + MOCK)
    args = args + ["--no-configs", "--recall-queries", "1"]   # one full oracle search (~10 s here)
    cmd = [sys.executable, str(script)] + args
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)] + args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout                       # ONE JSON line, rank 0 only
    return json.loads(lines[0])


def _check_common(line, n_gpus):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["n_gpus"] == n_gpus and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "l2_flush" in line["config"] and "model" not in line["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["unit"] == "GB/s"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == 768 * 4 * n_gpus and line["e2e"]["d2h_bytes_per_step"] > 0
    rc = line["recall_check"]
    assert {"checked", "ids_identical", "bit_exact", "certificate"} <= set(rc)
    assert rc["certificate"]["uncertified"] == 0
    assert line["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert line["warmup"] >= 3


def test_cuda_arm_json_contract_single_gpu(tmp_path):
    line = _run(tmp_path, ["--steps", "20", "--warmup", "1", "--cpu-sample-rows", "20000"])
    _check_common(line, 1)
    assert line["steps"] == 20 and line["scaling"] == "strong"
    assert abs(line["value"] - 20 / 0.422) < 1e-6 and abs(line["ms_per_step"] - 21.1) < 1e-9
    assert line["roofline"]["traffic"] and line["roofline"]["traffic"] > 3.0e10        # committed ncu capture
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] == "port" and cb["cores"] >= 1


@pytest.mark.parametrize("extra,scaling,mult", [([], "strong", 1), (["--mode", "replicas"], "weak", 2),
                                                (["--exchange", "p2p"], "strong", 1)])
def test_cuda_arm_json_contract_two_ranks(tmp_path, extra, scaling, mult):
    line = _run(tmp_path, ["--gpus", "2", "--steps", "10", "--warmup", "3"] + extra, nproc=2)
    _check_common(line, 2)
    assert line["scaling"] == scaling and abs(line["value"] - mult * 10 / 0.422) < 1e-6
    assert "cpu_baseline" not in line and line["roofline"]["traffic"] is None


def test_reference_arm_runs_for_real():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--cpu-sample-rows", "20000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "queries/s"
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["workload"].startswith("single-query L2, N=10000000 d=768 fp32, k=10")
