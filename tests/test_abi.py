"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every
symbol include/tostore_cuda.h declares, fails loudly (no CPU fallback), and the
product package never touches the oracle."""
import ctypes as C
import os
import re
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "tostore_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tsc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from tostore_b200 import _native as N
    lib = N.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in tostore_cuda.h but not exported"
    assert sorted(N.EXPORTS) == syms, "python binding list out of sync with the header"
    assert lib.tsc_version() == 2


def test_struct_layout_matches_header():
    from tostore_b200 import _native as N
    from tostore_b200 import where as W
    assert C.sizeof(N.IndexDesc) == 80          # 40 (v1) + n_devices 4 + device_ids 32 + reserved 4
    assert C.sizeof(N.Stats) == 168 and N.Stats.certified_queries.offset == 112   # + last_tflops, last_tensor_util
    assert N.IndexDesc.capacity_rows.offset == 16 and N.IndexDesc.k_max.offset == 32
    assert C.sizeof(W.WhereOp) == 48 and W.WhereOp.i_lo.offset == 8 and W.WhereOp.args_offset.offset == 40
    assert C.sizeof(N.NghInfo) == 72 and N.NghInfo.next_node_id.offset == 24


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (what dart:ffi / cgo / any
    FFI generator consumes), struct sizes must equal the ctypes mirrors, and a C program
    must link against the .so and get the no-GPU error path (no C++ runtime needed by the
    caller)."""
    import shutil
    import subprocess
    from tostore_b200 import _native as N
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "tostore_cuda.h"
_Static_assert(sizeof(tsc_index_desc) == 80, "tsc_index_desc");
_Static_assert(sizeof(tsc_stats) == 168, "tsc_stats");
_Static_assert(sizeof(tsc_where_op) == 48, "tsc_where_op");
_Static_assert(sizeof(tsc_ngh_info) == 72, "tsc_ngh_info");
int main(void) {
  if (tsc_version() != TSC_ABI_VERSION) return 2;
  tsc_index_desc d;
  memset(&d, 0, sizeof d);
  d.struct_size = sizeof d;
  d.dims = 0;                              /* rejected before any CUDA call */
  uint64_t h = 0;
  int32_t rc = tsc_index_create(&d, &h);
  printf("%d %s|%s\n", rc, tsc_status_name(rc), tsc_last_error());
  uint8_t check[9] = {'1','2','3','4','5','6','7','8','9'};
  printf("%08x\n", tsc_selftest_crc32(check, 9));
  return rc == TSC_ERR_BAD_DIMS ? 0 : 3;
}
""")
    exe = tmp_path / "caller"
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-ltostore_cuda",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("-3 TSC_ERR_BAD_DIMS|") and "dims" in lines[0]
    assert lines[1] == "cbf43926"                 # CRC-32/IEEE check value


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tostore_b200 as t
    with pytest.raises(t.TscError) as e:
        t.GpuVectorIndex(128, capacity_rows=16)
    assert e.value.status == -5                  # TSC_ERR_CUDA, loudly
    store = t.GpuVectorStore()
    with pytest.raises(t.TscError):
        store.createVectorIndex("t", "e", t.VectorFieldConfig(dimensions=8))


def test_bad_arguments_are_rejected_before_cuda():
    from tostore_b200 import _native as N
    lib = N.lib()
    h = C.c_uint64(0)
    assert lib.tsc_index_create(None, C.byref(h)) == N.TSC_ERR_BAD_ARG
    d = N.IndexDesc(struct_size=C.sizeof(N.IndexDesc), dims=0, metric=0, src_precision=1,
                    dev_dtype=0, device_id=0, capacity_rows=10, k_max=10, nq_max=1)
    assert lib.tsc_index_create(C.byref(d), C.byref(h)) == N.TSC_ERR_BAD_DIMS
    assert b"dims" in lib.tsc_last_error()
    d.dims, d.metric = 8, 9
    assert lib.tsc_index_create(C.byref(d), C.byref(h)) == N.TSC_ERR_BAD_ARG
    d.metric, d.k_max = 0, 1000
    assert lib.tsc_index_create(C.byref(d), C.byref(h)) == N.TSC_ERR_BAD_ARG
    assert lib.tsc_index_destroy(12345) == N.TSC_ERR_BAD_HANDLE
    assert lib.tsc_search(777, None, 1, 1, 0.0, None, None, None) == N.TSC_ERR_BAD_HANDLE
    assert lib.tsc_status_name(-7) == b"TSC_ERR_PAGE"


def test_warp_sliced_crc_matches_crc32_ieee():
    from tostore_b200 import _native as N
    lib = N.lib()
    rng = np.random.default_rng(1)
    for n in (0, 1, 9, 31, 32, 33, 1000, 16364, 16384, 70001):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        buf = data.tobytes()
        assert lib.tsc_selftest_crc32(buf, n) == zlib.crc32(buf), n
    assert lib.tsc_selftest_crc32(b"123456789", 9) == 0xCBF43926


def test_product_path_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "tostore_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dp, f)).read()
                code = "\n".join(l for l in src.splitlines()
                                 if not l.lstrip().startswith(("#", "//", "*", "/*", '"""')))
                assert not re.search(r"^\s*(import|from)\s+oracle", code, flags=re.M), f
                assert "liboracle" not in code and "tso_" not in code.replace("tso_synth_value", ""), f


def test_c_example_compiles_and_fails_loudly_without_gpu(tmp_path):
    """examples/minimal.c (the whole C ABI end to end) builds as C99 against the header and
    the .so; on a host without a GPU it reports TSC_ERR_CUDA instead of computing anything."""
    import shutil
    import subprocess
    import torch
    from tostore_b200 import _native as N
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "minimal"
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "minimal.c"), "-o", str(exe), "-L", libdir,
                           "-ltostore_cuda", f"-Wl,-rpath,{libdir}", "-lm"])
    if torch.cuda.is_available():
        pytest.skip("GPU present: the example would run for real (see tools/history/gpu_r2_first.sh)")
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 1 and "TSC_ERR_CUDA" in out.stderr
