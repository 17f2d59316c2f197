"""CPU oracle package — TEST INFRASTRUCTURE ONLY.

`oracle.c_oracle()` loads `oracle/liboracle.so` (built by `oracle/Makefile`,
called from `__graft_entry__.build()`), the plain-C restatement of the
reference's vector-search arithmetic; `oracle.oracle_np` is the independent
numpy-float64 restatement that pins it. Only tests/, `__graft_entry__.smoke()`
and bench.py's cpu_baseline / `--impl reference` legs may import this package.
PARITY UNPINNED BY THE REFERENCE (no golden vectors exist upstream; no Dart SDK
here) — see tostore_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "tostore_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def c_oracle():
    """Load (building if needed) the C oracle and declare its prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    lib = C.CDLL(build())
    vp = C.c_void_p
    lib.tso_to_float32.argtypes = [_f64p, C.c_uint64, C.c_uint32, _f32p]
    lib.tso_to_float32.restype = None
    lib.tso_normalize_f32.argtypes = [_f32p, C.c_uint32, _f32p]
    lib.tso_normalize_f32.restype = C.c_int
    lib.tso_distance_to_score.argtypes = [C.c_double, C.c_int]
    lib.tso_distance_to_score.restype = C.c_double
    for name in ("tso_l2_distance", "tso_inner_product", "tso_cosine_similarity"):
        f = getattr(lib, name)
        f.argtypes = [_f32p, _f32p, C.c_uint32]
        f.restype = C.c_double
    lib.tso_exact_distance.argtypes = [_f32p, _f32p, C.c_uint32, C.c_int]
    lib.tso_exact_distance.restype = C.c_double
    lib.tso_compare.argtypes = [C.c_double, C.c_int64, C.c_double, C.c_int64]
    lib.tso_compare.restype = C.c_int
    lib.tso_search.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int64, vp, vp,
                               _f32p, C.c_int, C.c_uint32, C.c_double, C.c_int, _i64p, _f64p]
    lib.tso_search.restype = C.c_uint32
    lib.tso_search_synth.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int64,
                                     vp, vp, _f32p, C.c_int, C.c_uint32, C.c_double, C.c_int,
                                     _i64p, _f64p]
    lib.tso_search_synth.restype = C.c_uint32
    lib.tso_synth_value.argtypes = [C.c_uint64, C.c_uint64]
    lib.tso_synth_value.restype = C.c_float
    lib.tso_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, vp]
    lib.tso_synth_rows.restype = None
    lib.tso_round_bf16.argtypes = [C.c_float]
    lib.tso_round_bf16.restype = C.c_float
    lib.tso_round_f16.argtypes = [C.c_float]
    lib.tso_round_f16.restype = C.c_float
    lib.tso_round_rows.argtypes = [vp, C.c_uint64, C.c_int]
    lib.tso_round_rows.restype = None
    lib.tso_crc32.argtypes = [vp, C.c_size_t]
    lib.tso_crc32.restype = C.c_uint32
    lib.tso_build_page.argtypes = [C.c_uint8, vp, C.c_uint32, C.c_uint32, _u8p]
    lib.tso_build_page.restype = C.c_int
    lib.tso_bytes_per_element.argtypes = [C.c_int]
    lib.tso_bytes_per_element.restype = C.c_uint32
    lib.tso_vectors_per_raw_page.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    lib.tso_vectors_per_raw_page.restype = C.c_uint32
    lib.tso_nodes_per_graph_page.argtypes = [C.c_uint32, C.c_uint32]
    lib.tso_nodes_per_graph_page.restype = C.c_uint32
    lib.tso_build_rawvec_page.argtypes = [_f32p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _u8p]
    lib.tso_build_rawvec_page.restype = C.c_int
    lib.tso_parse_rawvec_page.argtypes = [_u8p, C.c_uint32, C.c_uint32, _f32p, C.c_uint32]
    lib.tso_parse_rawvec_page.restype = C.c_int32
    lib.tso_build_graph_page.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, _u8p]
    lib.tso_build_graph_page.restype = C.c_int
    lib.tso_parse_graph_page_flags.argtypes = [_u8p, C.c_uint32, _u8p, C.c_uint32]
    lib.tso_parse_graph_page_flags.restype = C.c_int32
    lib.tso_node_location.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint32)]
    lib.tso_node_location.restype = None
    lib.tso_max_threads.restype = C.c_int
    _u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
    lib.tso_string_compare.argtypes = [_u16p, C.c_uint32, _u16p, C.c_uint32]
    lib.tso_string_compare.restype = C.c_int
    lib.tso_like_match.argtypes = [_u16p, C.c_uint32, _u16p, C.c_uint32]
    lib.tso_like_match.restype = C.c_int
    _LIB = lib
    return lib


def _bitmap(mask, n):
    """bool[n] -> uint64 words, bit i of word i//64 (LSB first)."""
    if mask is None:
        return None
    m = np.zeros(((n + 63) // 64) * 64, dtype=bool)
    m[:n] = np.asarray(mask, dtype=bool)[:n]
    return np.packbits(m.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint64).copy()


def search(rows, query, metric, k, threshold=None, deleted=None, filter=None,
           first_node_id=0, threads=1):
    """C-oracle exhaustive search over an fp32 [n, ld] array (ld >= dims)."""
    lib = c_oracle()
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    n, ld = rows.shape
    dims = query.shape[0]
    ids = np.full(k, -1, dtype=np.int64)
    dist = np.full(k, np.nan, dtype=np.float64)
    db, fb = _bitmap(deleted, n), _bitmap(filter, n)
    thr = float("nan") if threshold is None else float(threshold)
    m = lib.tso_search(rows.ctypes.data, n, dims, ld, first_node_id,
                       None if db is None else db.ctypes.data,
                       None if fb is None else fb.ctypes.data,
                       query, metric, k, thr, threads, ids, dist)
    return ids[:m], dist[:m]


def search_synth(seed, n, dims, dev_dtype, query, metric, k, threshold=None,
                 deleted=None, filter=None, first_node_id=0, threads=0):
    """C-oracle exhaustive search over on-the-fly synthetic rows."""
    lib = c_oracle()
    query = np.ascontiguousarray(query, dtype=np.float32)
    ids = np.full(k, -1, dtype=np.int64)
    dist = np.full(k, np.nan, dtype=np.float64)
    db, fb = _bitmap(deleted, n), _bitmap(filter, n)
    thr = float("nan") if threshold is None else float(threshold)
    if threads <= 0:
        threads = lib.tso_max_threads()
    m = lib.tso_search_synth(seed, n, dims, dev_dtype, first_node_id,
                             None if db is None else db.ctypes.data,
                             None if fb is None else fb.ctypes.data,
                             query, metric, k, thr, threads, ids, dist)
    return ids[:m], dist[:m]


def synth_rows(seed, first_row, n, dims, ld=None):
    lib = c_oracle()
    ld = dims if ld is None else ld
    out = np.empty((n, ld), dtype=np.float32)
    lib.tso_synth_rows(seed, first_row, n, dims, ld, out.ctypes.data)
    return out
