"""ctypes binding of libtostore_cuda.so (include/tostore_cuda.h).

This is the Python twin of dart/tostore_cuda_bindings.dart: same entry points,
same ownership rules (caller-allocated buffers, int32 status, opaque uint64
handles). There is no fallback of any kind: if the shared library is missing or
a call fails, a `TscError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtostore_cuda.so")

# every symbol include/tostore_cuda.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "tsc_version", "tsc_device_count", "tsc_last_error", "tsc_status_name",
    "tsc_index_create", "tsc_index_destroy", "tsc_index_clear",
    "tsc_index_append_rows", "tsc_index_append_pages", "tsc_index_append_synthetic",
    "tsc_index_set_deleted", "tsc_index_apply_graph_pages", "tsc_index_set_filter",
    "tsc_search", "tsc_search_submit", "tsc_search_poll", "tsc_search_wait", "tsc_search_flags",
    "tsc_search_device", "tsc_index_set_pipelining", "tsc_vector_search", "tsc_vector_search_batch", "tsc_merge_shards",
    "tsc_selftest_query_prep", "tsc_selftest_distance_to_score",
    "tsc_comm_unique_id", "tsc_comm_init", "tsc_search_sharded",
    "tsc_comm_p2p_export", "tsc_comm_p2p_import",
    "tsc_stats_get", "tsc_stats_reset", "tsc_index_device_rows", "tsc_selftest_crc32",
    "tsc_debug_gemm_keys",
    "tsc_index_column_create", "tsc_index_column_append", "tsc_index_filter_where",
    "tsc_index_column_append_text", "tsc_index_filter_where_text", "tsc_selftest_where_text",
    "tsc_ngh_read_meta", "tsc_index_load_ngh", "tsc_selftest_ngh_walk",
    "tsc_selftest_where", "tsc_selftest_host_index", "tsc_selftest_pk_assemble",
    "tsc_index_set_primary_keys", "tsc_index_get_primary_key", "tsc_vector_search_pk",
    "tsc_index_filter_primary_keys", "tsc_selftest_pk_filter_bitmap",
)

TSC_OK = 0
TSC_ERR_BAD_HANDLE, TSC_ERR_BAD_ARG, TSC_ERR_BAD_DIMS, TSC_ERR_OOM = -1, -2, -3, -4
TSC_ERR_CUDA, TSC_ERR_NCCL, TSC_ERR_PAGE, TSC_ERR_UNSUPPORTED, TSC_ERR_NOT_READY = -5, -6, -7, -8, -9


class TscError(RuntimeError):
    def __init__(self, status: int, where: str, message: str):
        super().__init__(f"{where}: status {status} ({message})")
        self.status = status


class IndexDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dims", C.c_uint32),
        ("metric", C.c_uint8), ("src_precision", C.c_uint8),
        ("dev_dtype", C.c_uint8), ("reserved0", C.c_uint8),
        ("device_id", C.c_int32),
        ("capacity_rows", C.c_uint64), ("first_node_id", C.c_uint64),
        ("k_max", C.c_uint32), ("nq_max", C.c_uint32),
        ("n_devices", C.c_uint32), ("device_ids", C.c_int32 * 8), ("reserved1", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dims", C.c_uint32),
        ("rows", C.c_uint64), ("deleted_rows", C.c_uint64),
        ("device_bytes", C.c_uint64), ("row_stride_bytes", C.c_uint64),
        ("searches", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("last_search_ms", C.c_double), ("last_scan_gbs", C.c_double),
        ("last_path", C.c_uint32), ("reserved", C.c_uint32),
        ("hot_launches", C.c_uint64), ("hot_ms_total", C.c_double),
        ("hot_bytes_total", C.c_double), ("hot_flops_total", C.c_double),
        ("certified_queries", C.c_uint64), ("retried_queries", C.c_uint64),
        ("uncertified_queries", C.c_uint64), ("range_rows", C.c_uint64),
        ("n_devices", C.c_uint32), ("reserved2", C.c_uint32),
        ("last_tflops", C.c_double), ("last_tensor_util", C.c_double),
    ]


class NghInfo(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dims", C.c_uint32),
        ("metric", C.c_uint8), ("precision", C.c_uint8), ("reserved", C.c_uint16),
        ("page_size", C.c_uint32), ("max_degree", C.c_uint32), ("reserved2", C.c_uint32),
        ("next_node_id", C.c_uint64), ("max_partition_file_size", C.c_uint64),
        ("files_read", C.c_uint64), ("pages_read", C.c_uint64), ("bytes_read", C.c_uint64),
        ("seconds", C.c_double),
    ]


_lib = None


def lib():
    """Load libtostore_cuda.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TscError(TSC_ERR_UNSUPPORTED, "load",
                       f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; "
                       f"g.build()'` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    L.tsc_version.restype = i32
    L.tsc_device_count.restype = i32
    L.tsc_last_error.restype = C.c_char_p
    L.tsc_status_name.restype = C.c_char_p
    L.tsc_status_name.argtypes = [i32]
    L.tsc_index_create.argtypes = [C.POINTER(IndexDesc), C.POINTER(u64)]
    L.tsc_index_destroy.argtypes = [u64]
    L.tsc_index_clear.argtypes = [u64]
    L.tsc_index_append_rows.argtypes = [u64, u64, vp, u64]
    L.tsc_index_append_pages.argtypes = [u64, u64, vp, u64, u32, u64]
    L.tsc_index_append_synthetic.argtypes = [u64, u64, u64, u64]
    L.tsc_index_set_deleted.argtypes = [u64, vp, u64, C.c_uint8]
    L.tsc_index_apply_graph_pages.argtypes = [u64, u64, vp, u64, u32]
    L.tsc_index_set_filter.argtypes = [u64, vp, u64]
    L.tsc_search.argtypes = [u64, vp, u32, u32, C.c_double, vp, vp, vp]
    L.tsc_search_submit.argtypes = [u64, vp, u32, u32, C.c_double, vp, vp, vp, C.POINTER(u64)]
    L.tsc_search_poll.argtypes = [u64, C.POINTER(i32)]
    L.tsc_search_wait.argtypes = [u64]
    L.tsc_search_device.argtypes = [u64, vp, u32, u32, C.c_double, vp, vp, vp, vp]
    L.tsc_index_set_pipelining.argtypes = [u64, i32]
    L.tsc_vector_search.argtypes = [u64, vp, u64, u32, C.c_double, vp, vp, vp, C.POINTER(u32)]
    L.tsc_vector_search_batch.argtypes = [u64, vp, u64, u32, u32, C.c_double, vp, vp, vp, vp]
    L.tsc_selftest_query_prep.argtypes = [u32, i32, vp, u64, vp]
    L.tsc_selftest_distance_to_score.argtypes = [i32, C.c_double]
    L.tsc_selftest_distance_to_score.restype = C.c_double
    L.tsc_merge_shards.argtypes = [u64, vp, vp, u32, u32, u32, vp, vp, vp, vp]
    L.tsc_comm_unique_id.argtypes = [vp]
    L.tsc_comm_init.argtypes = [u64, vp, i32, i32]
    L.tsc_comm_p2p_export.argtypes = [u64, i32, i32, vp]
    L.tsc_comm_p2p_import.argtypes = [u64, vp, i32]
    L.tsc_search_flags.argtypes = [u64, u32, vp]
    L.tsc_search_sharded.argtypes = [u64, vp, u32, u32, C.c_double, vp, vp, vp, vp]
    L.tsc_stats_get.argtypes = [u64, C.POINTER(Stats)]
    L.tsc_stats_reset.argtypes = [u64]
    L.tsc_index_device_rows.argtypes = [u64, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
    L.tsc_debug_gemm_keys.argtypes = [u64, vp, u32, vp]
    L.tsc_index_column_create.argtypes = [u64, u32, C.c_uint8]
    L.tsc_index_column_append.argtypes = [u64, u32, u64, vp, vp, u64]
    L.tsc_index_filter_where.argtypes = [u64, vp, u32, vp, u32, C.POINTER(u64)]
    L.tsc_index_column_append_text.argtypes = [u64, u32, u64, vp, vp, vp, u64]
    L.tsc_index_filter_where_text.argtypes = [u64, vp, u32, vp, u32, vp, vp, u32, C.POINTER(u64)]
    L.tsc_selftest_where_text.argtypes = [vp, u32, vp, u32, vp, vp, u32, u32, vp, vp, vp, vp, vp,
                                          vp, u64, vp]
    L.tsc_ngh_read_meta.argtypes = [C.c_char_p, C.POINTER(NghInfo)]
    L.tsc_index_load_ngh.argtypes = [u64, C.c_char_p, u32, C.POINTER(NghInfo)]
    L.tsc_selftest_ngh_walk.argtypes = [C.c_char_p, u32, u64, u64, u32, vp, vp, vp, u32,
                                        C.POINTER(u32)]
    L.tsc_selftest_where.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, vp, u64, vp]
    L.tsc_selftest_host_index.argtypes = [u64, u64, C.POINTER(u64)]
    L.tsc_selftest_pk_assemble.argtypes = [u64, u32, vp, vp, vp, vp, u64, vp, C.POINTER(u32)]
    L.tsc_index_set_primary_keys.argtypes = [u64, u64, vp, vp, u64]
    L.tsc_index_get_primary_key.argtypes = [u64, u64, vp, u32, C.POINTER(u32)]
    L.tsc_index_filter_primary_keys.argtypes = [u64, vp, vp, u64, C.POINTER(u64)]
    L.tsc_selftest_pk_filter_bitmap.argtypes = [u64, vp, vp, u64, vp, u64, C.POINTER(u64)]
    L.tsc_vector_search_pk.argtypes = [u64, vp, u64, u32, C.c_double, vp, vp, vp, vp, u64, vp,
                                       C.POINTER(u32)]
    L.tsc_selftest_crc32.argtypes = [vp, u32]
    L.tsc_selftest_crc32.restype = u32
    for name in EXPORTS:
        f = getattr(L, name)
        if f.restype is C.c_int:  # default restype: all remaining entry points return int32
            f.restype = i32
    _lib = L
    return L


def check(status: int, where: str) -> None:
    if status != TSC_OK:
        L = lib()
        msg = L.tsc_last_error().decode("utf-8", "replace")
        name = L.tsc_status_name(status).decode()
        raise TscError(status, where, f"{name}: {msg}")
