"""GpuVectorIndex — thin object wrapper over the C ABI for one shard on one GPU.

Stands where `NghGraphEngine` stands in the reference
(/root/reference/lib/src/core/ngh_graph_engine.dart:67-135): given an already
prepared fp32 query it returns (nodeId, fp64 distance) pairs in ascending
distance order. All compute happens in libtostore_cuda.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _native as N

METRIC_L2, METRIC_INNER_PRODUCT, METRIC_COSINE = 0, 1, 2
SRC_F64, SRC_F32, SRC_I8 = 0, 1, 2
DEV_F32, DEV_BF16, DEV_F16 = 0, 1, 2

_SRC_NP = {SRC_F64: np.float64, SRC_F32: np.float32, SRC_I8: np.int8}


class GpuVectorIndex:
    """One shard on one GPU, or — with `device_ids=[...]` (2..8 devices) — a GROUP: the column
    row-range sharded over the GPUs of this process behind one handle; every method that takes
    host buffers works on it unchanged and `search` merges the shards over NVLink."""

    def __init__(self, dims: int, metric: int = METRIC_COSINE, *, capacity_rows: int,
                 src_precision: int = SRC_F32, dev_dtype: int = DEV_F32, device_id: int = 0,
                 first_node_id: int = 0, k_max: int = 32, nq_max: int = 64, device_ids=None):
        self._lib = N.lib()
        self.dims, self.metric = int(dims), int(metric)
        self.src_precision, self.dev_dtype = int(src_precision), int(dev_dtype)
        self.device_id, self.first_node_id = int(device_id), int(first_node_id)
        self.k_max, self.nq_max = int(k_max), int(nq_max)
        self.device_ids = [int(d) for d in device_ids] if device_ids else None
        desc = N.IndexDesc(struct_size=C.sizeof(N.IndexDesc), dims=self.dims, metric=self.metric,
                           src_precision=self.src_precision, dev_dtype=self.dev_dtype,
                           device_id=self.device_id, capacity_rows=int(capacity_rows),
                           first_node_id=self.first_node_id, k_max=self.k_max, nq_max=self.nq_max)
        if self.device_ids:
            if len(self.device_ids) > 8:
                raise ValueError("at most 8 devices per group")
            desc.n_devices = len(self.device_ids)
            for i, dv in enumerate(self.device_ids):
                desc.device_ids[i] = dv
        h = C.c_uint64(0)
        N.check(self._lib.tsc_index_create(C.byref(desc), C.byref(h)), "tsc_index_create")
        self.handle = h.value

    # -- lifetime ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "handle", 0):
            self._lib.tsc_index_destroy(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def clear(self) -> None:
        N.check(self._lib.tsc_index_clear(self.handle), "tsc_index_clear")
        self._col_rows = {}            # attribute columns restart at the first node id too

    # -- ingestion -------------------------------------------------------------
    def append_rows(self, rows, first_node_id: Optional[int] = None) -> None:
        rows = np.ascontiguousarray(rows, dtype=_SRC_NP[self.src_precision])
        if rows.ndim != 2 or rows.shape[1] != self.dims:
            raise ValueError(f"rows must be [n, {self.dims}]")
        if first_node_id is None:
            first_node_id = self.first_node_id + self.stats().rows
        N.check(self._lib.tsc_index_append_rows(self.handle, int(first_node_id),
                                                rows.ctypes.data, rows.shape[0]),
                "tsc_index_append_rows")

    def append_pages(self, pages: bytes, first_logical_page: int, page_size: int,
                     live_rows: int) -> None:
        buf = np.frombuffer(pages, dtype=np.uint8)
        if buf.size % page_size:
            raise ValueError("pages is not a whole number of pages")
        N.check(self._lib.tsc_index_append_pages(self.handle, int(first_logical_page),
                                                 buf.ctypes.data, buf.size // page_size,
                                                 int(page_size), int(live_rows)),
                "tsc_index_append_pages")

    def append_synthetic(self, seed: int, n_rows: int, first_node_id: Optional[int] = None) -> None:
        if first_node_id is None:
            first_node_id = self.first_node_id + self.stats().rows
        N.check(self._lib.tsc_index_append_synthetic(self.handle, int(seed), int(first_node_id),
                                                     int(n_rows)),
                "tsc_index_append_synthetic")

    def load_ngh(self, index_dir: str, tombstones: bool = True) -> N.NghInfo:
        """Cold start from an on-disk ToStore NGH index directory (the one that holds
        `ngh/meta.json`): the library's reader thread streams this shard's partition
        files through pinned double buffers into append_pages / apply_graph_pages."""
        info = N.NghInfo(struct_size=C.sizeof(N.NghInfo))
        N.check(self._lib.tsc_index_load_ngh(self.handle, str(index_dir).encode("utf-8"),
                                             1 if tombstones else 0, C.byref(info)),
                "tsc_index_load_ngh")
        return info

    # -- liveness ----------------------------------------------------------------
    def set_deleted(self, node_ids, deleted: bool = True) -> None:
        ids = np.ascontiguousarray(node_ids, dtype=np.uint64)
        N.check(self._lib.tsc_index_set_deleted(self.handle, ids.ctypes.data, ids.size,
                                                1 if deleted else 0), "tsc_index_set_deleted")

    def apply_graph_pages(self, pages: bytes, first_logical_page: int, page_size: int) -> None:
        buf = np.frombuffer(pages, dtype=np.uint8)
        N.check(self._lib.tsc_index_apply_graph_pages(self.handle, int(first_logical_page),
                                                      buf.ctypes.data, buf.size // page_size,
                                                      int(page_size)),
                "tsc_index_apply_graph_pages")

    def set_filter(self, mask) -> None:
        """mask: bool[rows] (True = row may be returned) or None to clear."""
        if mask is None:
            N.check(self._lib.tsc_index_set_filter(self.handle, None, 0), "tsc_index_set_filter")
            return
        m = np.asarray(mask, dtype=bool)
        padded = np.zeros(((m.size + 63) // 64) * 64, dtype=bool)
        padded[: m.size] = m
        words = np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
        words = words.view(np.uint64).copy()
        N.check(self._lib.tsc_index_set_filter(self.handle, words.ctypes.data, words.size),
                "tsc_index_set_filter")

    # -- structured WHERE prefilter (attribute columns, evaluated on the GPU) --------
    def column_create(self, column_id: int, col_type: int) -> None:
        """col_type: where.COL_I64 / where.COL_F64 / where.COL_TEXT."""
        N.check(self._lib.tsc_index_column_create(self.handle, int(column_id), int(col_type)),
                "tsc_index_column_create")
        self._col_types = getattr(self, "_col_types", {})
        self._col_types[int(column_id)] = int(col_type)

    def column_append(self, column_id: int, values, is_null=None,
                      first_node_id: Optional[int] = None) -> None:
        """values: int64 / float64 array per the column type; is_null: bool array or None.
        A list may hold None for NULL."""
        t = getattr(self, "_col_types", {}).get(int(column_id))
        if t is None:
            raise KeyError(f"column {column_id} was not created on this index")
        if t == 2:
            return self._column_append_text(column_id, values, is_null, first_node_id)
        if is_null is None and isinstance(values, (list, tuple)) and any(v is None for v in values):
            is_null = np.array([v is None for v in values], dtype=bool)
            values = [0 if v is None else v for v in values]
        vals = np.ascontiguousarray(values, dtype=np.int64 if t == 0 else np.float64)
        nulls = None if is_null is None else np.ascontiguousarray(is_null, dtype=np.uint8)
        if nulls is not None and nulls.size != vals.size:
            raise ValueError("is_null must have one entry per value")
        if first_node_id is None:
            self._col_rows = getattr(self, "_col_rows", {})
            first_node_id = self.first_node_id + self._col_rows.get(int(column_id), 0)
        N.check(self._lib.tsc_index_column_append(self.handle, int(column_id), int(first_node_id),
                                                  vals.ctypes.data,
                                                  None if nulls is None else nulls.ctypes.data,
                                                  vals.size), "tsc_index_column_append")
        self._col_rows = getattr(self, "_col_rows", {})
        end = int(first_node_id) - self.first_node_id + vals.size
        self._col_rows[int(column_id)] = max(self._col_rows.get(int(column_id), 0), end)

    def _column_append_text(self, column_id: int, values, is_null, first_node_id) -> None:
        """values: strings (None = NULL), stored as their UTF-16 code units — a Dart String's
        own form — through tsc_index_column_append_text."""
        from .where import utf16_pool
        values = list(values)
        nulls = np.array([v is None for v in values], dtype=np.uint8)
        if is_null is not None:
            nulls |= np.ascontiguousarray(is_null, dtype=np.uint8) != 0
        units, offsets = utf16_pool(["" if n else v for v, n in zip(values, nulls)])
        if first_node_id is None:
            self._col_rows = getattr(self, "_col_rows", {})
            first_node_id = self.first_node_id + self._col_rows.get(int(column_id), 0)
        N.check(self._lib.tsc_index_column_append_text(
            self.handle, int(column_id), int(first_node_id), units.ctypes.data, offsets.ctypes.data,
            nulls.ctypes.data if nulls.any() else None, len(values)), "tsc_index_column_append_text")
        self._col_rows = getattr(self, "_col_rows", {})
        end = int(first_node_id) - self.first_node_id + len(values)
        self._col_rows[int(column_id)] = max(self._col_rows.get(int(column_id), 0), end)

    def filter_where(self, program) -> int:
        """Install the rows matching a compiled `where.WhereProgram` as the filter;
        returns how many rows passed."""
        ops, n_ops, raw, n_args = program.buffers()
        matched = C.c_uint64(0)
        if program.texts:
            units, offsets = program.text_buffers()
            N.check(self._lib.tsc_index_filter_where_text(
                self.handle, C.cast(ops, C.c_void_p), n_ops, raw.ctypes.data, n_args,
                units.ctypes.data, offsets.ctypes.data, len(program.texts), C.byref(matched)),
                "tsc_index_filter_where_text")
        else:
            N.check(self._lib.tsc_index_filter_where(self.handle, C.cast(ops, C.c_void_p), n_ops,
                                                     raw.ctypes.data, n_args, C.byref(matched)),
                    "tsc_index_filter_where")
        return matched.value

    # -- search --------------------------------------------------------------------
    def search(self, queries, k: int, threshold: Optional[float] = None):
        """queries: fp32 [nq, dims] (prepared as the reference prepares them).
        Returns (ids int64 [nq,k] -1 padded, dist float64 [nq,k], counts uint32 [nq])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dims:
            raise ValueError(f"queries must be [nq, {self.dims}]")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        counts = np.empty(nq, dtype=np.uint32)
        thr = math.nan if threshold is None else float(threshold)
        N.check(self._lib.tsc_search(self.handle, q.ctypes.data, nq, k, thr, ids.ctypes.data,
                                     dist.ctypes.data, counts.ctypes.data), "tsc_search")
        return ids, dist, counts

    def search_async(self, queries, k: int, threshold: Optional[float] = None):
        """submit / poll pair; returns a callable `poll()` -> None | (ids, dist, counts)."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dims)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        counts = np.empty(nq, dtype=np.uint32)
        t = C.c_uint64(0)
        thr = math.nan if threshold is None else float(threshold)
        N.check(self._lib.tsc_search_submit(self.handle, q.ctypes.data, nq, k, thr,
                                            ids.ctypes.data, dist.ctypes.data, counts.ctypes.data,
                                            C.byref(t)), "tsc_search_submit")
        keep = (q,)

        def poll(block: bool = False):
            _ = keep
            if block:
                N.check(self._lib.tsc_search_wait(t.value), "tsc_search_wait")
                return ids, dist, counts
            done = C.c_int32(0)
            N.check(self._lib.tsc_search_poll(t.value, C.byref(done)), "tsc_search_poll")
            return (ids, dist, counts) if done.value else None

        return poll

    def search_device(self, d_queries: int, nq: int, k: int, d_ids: int, d_dist: int,
                      d_counts: int, threshold: Optional[float] = None, stream: int = 0,
                      sharded: bool = False) -> None:
        """Raw device pointers (ints); asynchronous on `stream` (0 = index stream)."""
        thr = math.nan if threshold is None else float(threshold)
        fn = self._lib.tsc_search_sharded if sharded else self._lib.tsc_search_device
        N.check(fn(self.handle, d_queries, nq, k, thr, d_ids, d_dist, d_counts, stream or None),
                "tsc_search_sharded" if sharded else "tsc_search_device")

    def set_pipelining(self, on: bool) -> None:
        """Throughput mode for back-to-back `search_device` calls on one stream: the HBM pass of
        search i+1 overlaps the tail / exchange of search i (tsc_index_set_pipelining; see the
        header for the contract on query buffers)."""
        N.check(self._lib.tsc_index_set_pipelining(self.handle, 1 if on else 0),
                "tsc_index_set_pipelining")

    def vector_search(self, values, k: int, threshold: Optional[float] = None):
        """fp64 query of any length -> (ids, dist, score) with the reference's
        query preparation and score mapping done inside the library."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        ids = np.empty(k, dtype=np.int64)
        dist = np.empty(k, dtype=np.float64)
        score = np.empty(k, dtype=np.float64)
        cnt = C.c_uint32(0)
        thr = math.nan if threshold is None else float(threshold)
        N.check(self._lib.tsc_vector_search(self.handle, v.ctypes.data, v.size, k, thr,
                                            ids.ctypes.data, dist.ctypes.data, score.ctypes.data,
                                            C.byref(cnt)), "tsc_vector_search")
        n = cnt.value
        return ids[:n], dist[:n], score[:n]

    def vector_search_batch(self, values, k: int, threshold: Optional[float] = None):
        """Batch form of `vector_search` (additive): values [nq, len] fp64, every query
        prepared like a single one, one search call (tensor-core path for 16-bit columns and
        nq >= 5 on 16-bit columns, nq >= 9 on fp32 columns). Returns (ids [nq,k], dist [nq,k], score [nq,k], counts [nq])."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        if v.ndim != 2:
            raise ValueError("values must be [nq, len]")
        nq = v.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        score = np.empty((nq, k), dtype=np.float64)
        counts = np.empty(nq, dtype=np.uint32)
        thr = math.nan if threshold is None else float(threshold)
        N.check(self._lib.tsc_vector_search_batch(self.handle, v.ctypes.data, v.shape[1], nq, k, thr,
                                                  ids.ctypes.data, dist.ctypes.data,
                                                  score.ctypes.data, counts.ctypes.data),
                "tsc_vector_search_batch")
        return ids, dist, score, counts

    # -- nodeId -> primary key side table (role of `__nid2pk`) ------------------------
    def set_primary_keys(self, pks, first_node_id: Optional[int] = None) -> None:
        """pks: sequence of str (None / '' = tombstone mapping) for consecutive node ids."""
        enc = [b"" if p is None else str(p).encode("utf-8") for p in pks]
        offs = np.zeros(len(enc) + 1, dtype=np.uint64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
        if first_node_id is None:
            first_node_id = self.first_node_id
        N.check(self._lib.tsc_index_set_primary_keys(self.handle, int(first_node_id),
                                                     blob.ctypes.data, offs.ctypes.data, len(enc)),
                "tsc_index_set_primary_keys")

    @staticmethod
    def _key_blob(pks):
        enc = [b"" if p is None else str(p).encode("utf-8") for p in pks]
        offs = np.zeros(len(enc) + 1, dtype=np.uint64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        return np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8), offs, len(enc)

    def filter_primary_keys(self, pks) -> int:
        """WHERE prefilter from a set of primary keys (what any ToStore query returns): the
        rows whose key is in the set stay searchable; returns how many rows that is."""
        blob, offs, n = self._key_blob(pks)
        matched = C.c_uint64(0)
        N.check(self._lib.tsc_index_filter_primary_keys(self.handle, blob.ctypes.data, offs.ctypes.data,
                                                        n, C.byref(matched)),
                "tsc_index_filter_primary_keys")
        return matched.value

    def get_primary_key(self, node_id: int) -> Optional[str]:
        buf = (C.c_uint8 * 4096)()
        n = C.c_uint32(0)
        N.check(self._lib.tsc_index_get_primary_key(self.handle, int(node_id), buf, 4096,
                                                    C.byref(n)), "tsc_index_get_primary_key")
        return bytes(buf[: n.value]).decode("utf-8") if n.value else None

    def vector_search_pk(self, values, k: int, threshold: Optional[float] = None,
                         pk_capacity: int = 1 << 16):
        """`vector_search` + result assembly inside the library: (pks, ids, dist, score);
        results whose node has no primary-key mapping are dropped."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        ids = np.empty(k, dtype=np.int64)
        dist = np.empty(k, dtype=np.float64)
        score = np.empty(k, dtype=np.float64)
        pkb = np.empty(pk_capacity, dtype=np.uint8)
        offs = np.zeros(k + 1, dtype=np.uint64)
        cnt = C.c_uint32(0)
        thr = math.nan if threshold is None else float(threshold)
        N.check(self._lib.tsc_vector_search_pk(self.handle, v.ctypes.data, v.size, k, thr,
                                               ids.ctypes.data, dist.ctypes.data,
                                               score.ctypes.data, pkb.ctypes.data, pk_capacity,
                                               offs.ctypes.data, C.byref(cnt)),
                "tsc_vector_search_pk")
        n = cnt.value
        raw = pkb.tobytes()
        pks = [raw[int(offs[i]): int(offs[i + 1])].decode("utf-8") for i in range(n)]
        return pks, ids[:n], dist[:n], score[:n]

    # -- sharding --------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        N.check(N.lib().tsc_comm_unique_id(buf), "tsc_comm_unique_id")
        return bytes(buf)

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        N.check(self._lib.tsc_comm_init(self.handle, buf, n_ranks, rank), "tsc_comm_init")

    def comm_init_p2p(self, dist, n_ranks: int, rank: int, root: int = -1) -> None:
        """Peer-memory exchange (tsc_exchange.cuh): every shard pushes its top-k over NVLink;
        root = -1: every rank receives the global result, root = r: only rank r does. `dist` is
        an initialised torch.distributed module (any backend) used only to all-gather the
        64-byte CUDA IPC handles of the receive buffers."""
        buf = (C.c_uint8 * 64)()
        N.check(self._lib.tsc_comm_p2p_export(self.handle, n_ranks, rank, buf), "tsc_comm_p2p_export")
        handles = [None] * n_ranks
        dist.all_gather_object(handles, bytes(buf))
        blob = (C.c_uint8 * (64 * n_ranks)).from_buffer_copy(b"".join(handles))
        N.check(self._lib.tsc_comm_p2p_import(self.handle, blob, int(root)), "tsc_comm_p2p_import")
        dist.barrier()          # every rank has mapped every buffer before the first search

    def merge_shards(self, d_part_ids: int, d_part_dist: int, n_parts: int, nq: int, k: int,
                     d_ids: int, d_dist: int, d_counts: int, stream: int = 0) -> None:
        N.check(self._lib.tsc_merge_shards(self.handle, d_part_ids, d_part_dist, n_parts, nq, k,
                                           d_ids, d_dist, d_counts, stream or None),
                "tsc_merge_shards")

    def search_flags(self, nq: int):
        """Per-query verdict of the exactness certificate for the last search:
        0 exact, 1 range pass owed (device-buffer searches only), 2 uncertified."""
        f = np.zeros(nq, dtype=np.uint32)
        N.check(self._lib.tsc_search_flags(self.handle, int(nq), f.ctypes.data), "tsc_search_flags")
        return f

    # -- observability -----------------------------------------------------------------
    def stats(self) -> N.Stats:
        s = N.Stats(struct_size=C.sizeof(N.Stats))
        N.check(self._lib.tsc_stats_get(self.handle, C.byref(s)), "tsc_stats_get")
        return s

    def stats_reset(self) -> None:
        N.check(self._lib.tsc_stats_reset(self.handle), "tsc_stats_reset")

    def device_rows(self):
        p, n, stride = C.c_void_p(0), C.c_uint64(0), C.c_uint64(0)
        N.check(self._lib.tsc_index_device_rows(self.handle, C.byref(p), C.byref(n),
                                                C.byref(stride)), "tsc_index_device_rows")
        return p.value, n.value, stride.value
