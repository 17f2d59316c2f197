"""numpy-float64 ORACLE — test infrastructure only, never on the product path.

An independent second restatement of the reference's vector-search arithmetic
(tocreator/tostore @ 130da06, Dart). It exists to pin the C oracle
(`oracle/tostore_oracle.c`): both must agree bit-for-bit on distances and on
result order. PARITY UNPINNED BY THE REFERENCE — the reference has no test,
golden vector or fixture for `vectorSearch` and cannot be executed here (no Dart
SDK), see SURVEY.md §8c.

Citations are relative to /root/reference/lib/src.
"""
from __future__ import annotations

import math
import struct
import zlib

import numpy as np

L2, INNER_PRODUCT, COSINE = 0, 1, 2          # model/table_schema.dart:2511-2531
F64, F32, I8 = 0, 1, 2                       # model/table_schema.dart:2481-2498
DEV_F32, DEV_BF16, DEV_F16 = 0, 1, 2

PAGE_MAGIC = 0x32475054                      # core/btree_page.dart:134
PAGE_HEADER = 20                             # core/btree_page.dart:133
PT_NGH_GRAPH, PT_NGH_RAWVEC = 6, 8           # core/btree_page.dart:14-55


# --------------------------------------------------------------------------
# query preparation — core/vector_index_manager.dart
# --------------------------------------------------------------------------
def to_float32(values, dims: int) -> np.ndarray:
    """_toFloat32 :1385-1392 — zero-filled, min(len, dims) copied, RNE to fp32."""
    out = np.zeros(dims, dtype=np.float32)
    v = np.asarray(values, dtype=np.float64)[:dims]
    with np.errstate(over="ignore"):
        out[: len(v)] = v.astype(np.float32)
    return out


def _seq_sum(terms: np.ndarray) -> np.ndarray:
    """Left-to-right fp64 sum starting from +0.0 (cumsum is sequential)."""
    z = np.zeros(terms.shape[:-1] + (1,), dtype=np.float64)
    return np.cumsum(np.concatenate([z, terms], axis=-1), axis=-1)[..., -1]


def normalize_f32(v: np.ndarray) -> np.ndarray:
    """_normalizeFloat32 :1395-1408."""
    v = np.asarray(v, dtype=np.float32)
    v64 = v.astype(np.float64)
    mag = math.sqrt(float(_seq_sum(v64 * v64)))
    if mag == 0:
        return v
    inv = 1.0 / mag
    return (v64 * inv).astype(np.float32)


def distance_to_score(distance: float, metric: int) -> float:
    """_distanceToScore :1411-1423."""
    if metric == L2:
        return 1.0 / (1.0 + distance)
    if metric == INNER_PRODUCT:
        try:
            return 1.0 / (1.0 + math.exp(-(-distance)))
        except OverflowError:
            return 0.0
    s = 1.0 - distance
    if s != s:
        return s
    return min(max(s, 0.0), 1.0)


# --------------------------------------------------------------------------
# exact distances — core/ngh_graph_engine.dart:908-946
# --------------------------------------------------------------------------
def exact_distances(query: np.ndarray, rows: np.ndarray, metric: int) -> np.ndarray:
    """`_exactDistance(query, row)` for every row of `rows` ([n, d] fp32)."""
    a = np.asarray(query, dtype=np.float32).astype(np.float64)[None, :]
    b = np.asarray(rows, dtype=np.float32).astype(np.float64)
    b = b[:, : a.shape[1]]
    with np.errstate(all="ignore"):
        if metric == L2:                                   # :920-927
            diff = a - b
            return np.sqrt(_seq_sum(diff * diff))
        if metric == INNER_PRODUCT:                        # :929-935, negated :914
            return -_seq_sum(a * b)
        dot = _seq_sum(a * b)                              # :937-946
        mag_a = _seq_sum(a * a)
        mag_b = _seq_sum(b * b)
        denom = np.sqrt(mag_a) * np.sqrt(mag_b)
        sim = np.where(denom > 0, dot / np.where(denom > 0, denom, 1.0), 0.0)
        return 1.0 - sim


def _order_key(d: np.ndarray) -> np.ndarray:
    """Map doubles to int64 keys ordered like Dart's double.compareTo
    (-0.0 < 0.0, NaN last)."""
    d = np.where(np.isnan(d), np.float64("nan"), d).astype(np.float64)
    bits = d.view(np.int64).copy()
    nan = np.isnan(d)
    bits[nan] = np.int64(0x7FF8000000000000)
    neg = bits < 0
    # negatives: larger magnitude -> smaller key; -0.0 lands at -1, just below +0.0
    bits[neg] = np.int64(-(2**63)) - bits[neg] - np.int64(1)
    return bits


def search(rows, query, metric: int, k: int, threshold=None, deleted=None,
           filter=None, first_node_id: int = 0):
    """Exhaustive search with the reference's re-rank semantics
    (ngh_graph_engine.dart:122-134): drop `d > threshold`, sort ascending by
    (d, nodeId), first k. `deleted` / `filter` are boolean arrays of len n."""
    rows = np.asarray(rows, dtype=np.float32)
    n = rows.shape[0]
    d = exact_distances(query, rows, metric)
    live = np.ones(n, dtype=bool)
    if deleted is not None:
        live &= ~np.asarray(deleted, dtype=bool)
    if filter is not None:
        live &= np.asarray(filter, dtype=bool)
    if threshold is not None and not (threshold != threshold):
        with np.errstate(invalid="ignore"):
            live &= ~(d > threshold)
    ids = np.nonzero(live)[0].astype(np.int64)
    dd = d[ids]
    order = np.lexsort((ids, _order_key(dd)))[:k]
    return ids[order] + first_node_id, dd[order]


# --------------------------------------------------------------------------
# synthetic rows (bit-identical to tso_synth_value and csrc/synth.cuh)
# --------------------------------------------------------------------------
def synth_rows(seed: int, first_row: int, n: int, dims: int) -> np.ndarray:
    idx = (np.arange(n, dtype=np.uint64)[:, None] + np.uint64(first_row)) * np.uint64(dims) \
        + np.arange(dims, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    lanes = z.view(np.int16).reshape(n, dims, 4).astype(np.int32).sum(axis=2)
    return (lanes.astype(np.float32) * np.float32(2.0 ** -15)).astype(np.float32)


def round_dev(rows: np.ndarray, dev_dtype: int) -> np.ndarray:
    """Round fp32 through the device storage dtype (RNE) and back."""
    rows = np.asarray(rows, dtype=np.float32)
    if dev_dtype == DEV_F16:
        with np.errstate(over="ignore"):
            return rows.astype(np.float16).astype(np.float32)
    if dev_dtype == DEV_BF16:
        u = rows.view(np.uint32).astype(np.uint64)
        special = (u & 0x7F800000) == 0x7F800000
        r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        s = np.where((u & 0x007FFFFF) != 0, u | 0x00400000, u) & 0xFFFF0000
        return np.where(special, s, r).astype(np.uint32).view(np.float32)
    return rows


# --------------------------------------------------------------------------
# page envelope — core/btree_page.dart; NGH pages — core/ngh_page.dart
# --------------------------------------------------------------------------
def crc32(data: bytes) -> int:
    """Crc32.of :81-88 is plain CRC-32/IEEE == zlib.crc32."""
    return zlib.crc32(data) & 0xFFFFFFFF


def build_page(page_type: int, payload: bytes, page_size: int) -> bytes:
    """BTreePageHeader.encode :148-160 + BTreePageIO.buildPageBytes :173-203."""
    if PAGE_HEADER + len(payload) > page_size:
        raise ValueError("page overflow")
    hdr = struct.pack("<IHBBIII", PAGE_MAGIC, PAGE_HEADER, page_type, 0,
                      len(payload), crc32(payload), 0)
    return hdr + payload + b"\0" * (page_size - PAGE_HEADER - len(payload))


def parse_page(page: bytes):
    """BTreePageIO.parsePageBytes :206-226 → (type, payload)."""
    magic, hs, pt, _flags, ln, crc, _r = struct.unpack_from("<IHBBIII", page, 0)
    if magic != PAGE_MAGIC or hs != PAGE_HEADER or pt >= 10:
        raise ValueError("Invalid page header/magic")
    if PAGE_HEADER + ln > len(page):
        raise ValueError("Invalid payload length")
    payload = page[PAGE_HEADER:PAGE_HEADER + ln]
    if crc32(payload) != crc:
        raise ValueError("Page CRC mismatch")
    return pt, payload


def bytes_per_element(precision: int) -> int:          # ngh_page.dart:331-340
    return 8 if precision == F64 else (1 if precision == I8 else 4)


def vectors_per_raw_page(page_size: int, dims: int, bpe: int) -> int:  # :575-579
    usable = page_size - 20 - 8 - 64
    vec = dims * bpe
    return usable // vec if usable > 0 and vec > 0 else 0


def nodes_per_graph_page(page_size: int, max_degree: int) -> int:      # :559-566
    usable = page_size - 20 - 4 - 64
    return usable // (2 + max_degree * 4) if usable > 0 else 0


def encode_rows(rows: np.ndarray, precision: int) -> bytes:
    """setVectorFromFloat32 :394-412."""
    rows = np.asarray(rows, dtype=np.float32)
    if precision == F32:
        return rows.astype("<f4").tobytes()
    if precision == F64:
        return rows.astype("<f8").tobytes()
    c = np.clip(rows.astype(np.float64), -1.0, 1.0) * 127.0
    q = np.where(c >= 0, np.floor(c + 0.5), np.ceil(c - 0.5))   # Dart round(): half away
    return q.astype(np.int8).tobytes()


def decode_rows(data: bytes, count: int, dims: int, precision: int) -> np.ndarray:
    """getVectorAsFloat32 :368-389."""
    if precision == F32:
        a = np.frombuffer(data, dtype="<f4", count=count * dims)
        return a.reshape(count, dims).astype(np.float32)
    if precision == F64:
        a = np.frombuffer(data, dtype="<f8", count=count * dims)
        with np.errstate(over="ignore"):
            return a.reshape(count, dims).astype(np.float32)
    a = np.frombuffer(data, dtype=np.int8, count=count * dims).astype(np.float64)
    return (a / 127.0).astype(np.float32).reshape(count, dims)


def build_rawvec_page(rows: np.ndarray, dims: int, precision: int, page_size: int) -> bytes:
    """Full-capacity zero-padded page: NghRawVectorPage.empty :346-362 +
    encodePayload :414-425."""
    bpe = bytes_per_element(precision)
    cap = vectors_per_raw_page(page_size, dims, bpe)
    rows = np.asarray(rows, dtype=np.float32).reshape(-1, dims)
    if cap == 0 or rows.shape[0] > cap:
        raise ValueError("rows do not fit the page")
    data = encode_rows(rows, precision)
    data += b"\0" * (cap * dims * bpe - len(data))
    payload = struct.pack("<HHB3x", cap, dims, precision) + data
    return build_page(PT_NGH_RAWVEC, payload, page_size)


def parse_rawvec_page(page: bytes, expect_dims: int) -> np.ndarray:
    """tryDecodePayload :427-447 → all vectorCount rows as fp32."""
    pt, payload = parse_page(page)
    if pt != PT_NGH_RAWVEC or len(payload) < 8:
        raise ValueError("not a raw-vector page")
    vcount, dims, prec = struct.unpack_from("<HHB", payload, 0)
    bpe = bytes_per_element(prec)
    if dims == 0 or len(payload) < 8 + vcount * dims * bpe or dims != expect_dims:
        raise ValueError("bad raw-vector payload")
    return decode_rows(payload[8:], vcount, dims, prec)


def build_graph_page(flags, max_degree: int, page_size: int) -> bytes:
    """NghGraphPage.encodePayload :166-187 with empty neighbour lists."""
    cap = nodes_per_graph_page(page_size, max_degree)
    slot = 2 + max_degree * 4
    flags = list(flags)
    if len(flags) > cap:
        raise ValueError("too many slots")
    body = bytearray(cap * slot)
    for i, f in enumerate(flags):
        body[i * slot] = f
    return build_page(PT_NGH_GRAPH, struct.pack("<HH", cap, max_degree) + bytes(body), page_size)


def parse_graph_page_flags(page: bytes):
    pt, payload = parse_page(page)
    if pt != PT_NGH_GRAPH or len(payload) < 4:
        raise ValueError("not a graph page")
    count, deg = struct.unpack_from("<HH", payload, 0)
    slot = 2 + deg * 4
    if deg == 0 or len(payload) < 4 + count * slot:
        raise ValueError("bad graph payload")
    return [payload[4 + i * slot] for i in range(count)]


def node_location(node_id: int, per_page: int, pages_per_partition: int):
    """model/ngh_index_meta.dart:451-490."""
    logical = node_id // per_page
    return (logical // pages_per_partition, 1 + logical % pages_per_partition,
            node_id % per_page)


# --------------------------------------------------------------------------
# synthetic on-disk NGH index (test fixture writer): the files a flushed ToStore
# database holds for one vector index — meta.json, rawvec/ and graph/ partition
# files with a per-file meta page at page 0 (core/ngh_page.dart:29-98).
# --------------------------------------------------------------------------
PT_NGH_META = 5


def _partition_meta_page(partition: int, category: int, total: int, file_size: int,
                         page_size: int) -> bytes:
    payload = struct.pack("<IHHiiqqii", 0x3148474E, 1, category, partition, 0, total, file_size,
                          -1, 0)
    payload += b"\0" * (128 - len(payload))
    return build_page(PT_NGH_META, payload, page_size)


def write_ngh_index(index_dir: str, rows: np.ndarray, metric: str, precision: int,
                    deleted=None, page_size: int = 16384, max_partition_file_size: int = 16 * 1024 * 1024,
                    max_degree: int = 64, max_entries_per_dir: int = 500):
    import json
    import os
    rows = np.asarray(rows, dtype=np.float32)
    n, dims = rows.shape
    bpe = bytes_per_element(precision)
    rpp = vectors_per_raw_page(page_size, dims, bpe)
    npg = nodes_per_graph_page(page_size, max_degree)
    ppp = max_partition_file_size // page_size
    flags = np.zeros(n, dtype=np.uint8)
    if deleted is not None:
        flags[np.asarray(deleted, dtype=bool)] = 1

    def write_category(cat_name, cat_code, per_page, make_page):
        n_logical = -(-n // per_page) if n else 0
        for part in range(-(-n_logical // ppp) if n_logical else 0):
            d = os.path.join(index_dir, "ngh", cat_name, f"dir_{part // max_entries_per_dir}")
            os.makedirs(d, exist_ok=True)
            pages = []
            for lp in range(part * ppp, min((part + 1) * ppp, n_logical)):
                pages.append(make_page(lp * per_page, min(n, (lp + 1) * per_page)))
            size = (1 + len(pages)) * page_size
            with open(os.path.join(d, f"p{part}.ngh"), "wb") as f:
                f.write(_partition_meta_page(part, cat_code, sum(1 for _ in pages), size, page_size))
                f.write(b"".join(pages))

    os.makedirs(os.path.join(index_dir, "ngh"), exist_ok=True)
    write_category("rawvec", 2, rpp, lambda a, b: build_rawvec_page(rows[a:b], dims, precision, page_size))
    write_category("graph", 0, npg, lambda a, b: build_graph_page(flags[a:b].tolist(), max_degree, page_size))
    meta = {"version": 1, "name": "idx_embedding", "tableName": "t", "fieldName": "embedding",
            "dimensions": dims, "distanceMetric": metric,
            "precision": {F64: "float64", F32: "float32", I8: "int8"}[precision],
            "maxDegree": max_degree, "totalVectors": n, "deletedCount": int(flags.sum()),
            "nextNodeId": n, "nghPageSize": page_size, "rawVectorPartitionCount": 1,
            "graphPartitionCount": 1, "maxPartitionFileSize": max_partition_file_size}
    with open(os.path.join(index_dir, "ngh", "meta.json"), "w") as f:
        json.dump(meta, f)
    return meta
