"""Loader of on-disk NGH indexes: CPU part checks the directory walk / addressing
against the oracle's page codec; the GPU part loads the files and searches."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
from tostore_b200 import ngh_loader as L


def _make(tmp_path, n, dims, prec, metric="l2", mpfs=16 * 1024 * 1024, dead=None):
    rows = (np.random.default_rng(n + dims).standard_normal((n, dims)) * 0.5).astype(np.float32)
    onp.write_ngh_index(str(tmp_path), rows, metric, prec, deleted=dead, max_partition_file_size=mpfs)
    return rows


def test_directory_walk_matches_oracle_codec(tmp_path):
    n, dims = 1000, 96
    rows = _make(tmp_path, n, dims, onp.F32, mpfs=4 * 16384)       # 4 pages per file -> many files
    meta = L.read_meta(str(tmp_path))
    assert (meta.dimensions, meta.metric, meta.precision, meta.next_node_id) == (dims, 0, 1, n)
    assert meta.pages_per_partition == 4
    assert meta.vectors_per_raw_page == onp.vectors_per_raw_page(16384, dims, 4)
    assert meta.nodes_per_graph_page == 63
    got = np.zeros((0, dims), dtype=np.float32)
    expect_first = 0
    for first, data in L.iter_partition_pages(str(tmp_path), "rawvec", meta, meta.vectors_per_raw_page, n):
        assert first == expect_first                                 # contiguous logical pages
        for off in range(0, len(data), meta.page_size):
            got = np.concatenate([got, onp.parse_rawvec_page(data[off: off + meta.page_size], dims)])
        expect_first = first + len(data) // meta.page_size
    assert (got[:n] == rows).all() and not got[n:].any()
    part, local, slot = onp.node_location(n - 1, meta.vectors_per_raw_page, meta.pages_per_partition)
    assert L.partition_path(str(tmp_path), "rawvec", part).endswith(f"dir_0/p{part}.ngh")
    assert local >= 1 and slot == (n - 1) % meta.vectors_per_raw_page


def _native_walk(index_dir, category, lo, hi, chunk_pages, max_chunks=4096):
    import ctypes as C
    from tostore_b200 import _native as N
    first = np.zeros(max_chunks, dtype=np.uint64)
    npages = np.zeros(max_chunks, dtype=np.uint64)
    crc = np.zeros(max_chunks, dtype=np.uint32)
    n = C.c_uint32(0)
    N.check(N.lib().tsc_selftest_ngh_walk(str(index_dir).encode(), category, lo, hi, chunk_pages,
                                          first.ctypes.data, npages.ctypes.data, crc.ctypes.data,
                                          max_chunks, C.byref(n)), "tsc_selftest_ngh_walk")
    return [(int(first[i]), int(npages[i]), int(crc[i])) for i in range(n.value)]


def _python_pages(index_dir, category, meta, per_page, n):
    """{logical page: bytes} from the Python walk (the earlier, GPU-tested loader)."""
    out = {}
    for first, data in L.iter_partition_pages(str(index_dir), category, meta, per_page, n):
        for j in range(len(data) // meta.page_size):
            out[first + j] = data[j * meta.page_size: (j + 1) * meta.page_size]
    return out


def test_native_meta_parser_matches_python(tmp_path):
    import json
    _make(tmp_path, 300, 48, onp.I8, metric="innerProduct", mpfs=4 * 16384)
    assert L.read_meta_native(str(tmp_path)) == L.read_meta(str(tmp_path))
    # nested objects that repeat top-level key names, floats, escapes, missing optional keys
    meta_path = tmp_path / "ngh" / "meta.json"
    j = json.loads(meta_path.read_text())
    j["nodeIdToPkMeta"] = {"dimensions": 7, "name": "x\"}{", "nextNodeId": 1, "list": [1, {"nghPageSize": 3}]}
    j["timestamps"] = {"created": "2026-01-01T00:00:00", "modified": None}
    j["nextNodeId"] = 300.0
    del j["maxDegree"], j["distanceMetric"]
    meta_path.write_text(json.dumps(j, indent=2))
    m = L.read_meta_native(str(tmp_path))
    assert (m.dimensions, m.next_node_id, m.max_degree, m.metric, m.precision) == (48, 300, 64, 2, 2)
    assert m == L.read_meta(str(tmp_path))
    meta_path.write_text("[1, 2")
    from tostore_b200 import TscError
    with pytest.raises(TscError):
        L.read_meta_native(str(tmp_path))
    with pytest.raises(TscError):
        L.read_meta_native(str(tmp_path / "nowhere"))


@pytest.mark.parametrize("chunk_pages", [1, 3, 4, 64])
def test_native_walk_delivers_the_same_pages_as_the_python_walk(tmp_path, chunk_pages):
    import zlib
    n, dims = 1000, 96
    dead = np.zeros(n, dtype=bool)
    dead[::7] = True
    _make(tmp_path, n, dims, onp.F32, mpfs=4 * 16384, dead=dead)     # 4 data pages per file
    meta = L.read_meta(str(tmp_path))
    for cat, name, per in ((0, "rawvec", meta.vectors_per_raw_page), (1, "graph", meta.nodes_per_graph_page)):
        pages = _python_pages(tmp_path, name, meta, per, n)
        chunks = _native_walk(tmp_path, cat, 0, n, chunk_pages)
        seen = []
        for first, cnt, crc in chunks:
            assert 1 <= cnt <= chunk_pages
            assert first // 4 == (first + cnt - 1) // 4              # never across a partition file
            assert crc == zlib.crc32(b"".join(pages[p] for p in range(first, first + cnt)))
            seen.extend(range(first, first + cnt))
        assert seen == sorted(pages)                                 # every page once, in order


def test_native_walk_reads_only_the_shards_pages_and_survives_missing_files(tmp_path):
    import os
    n, dims = 2000, 96
    _make(tmp_path, n, dims, onp.F32, mpfs=4 * 16384)
    meta = L.read_meta(str(tmp_path))
    per = meta.vectors_per_raw_page
    lo, hi = 700, 1300                                               # a shard's node-id range
    chunks = _native_walk(tmp_path, 0, lo, hi, 64)
    got = [p for first, cnt, _ in chunks for p in range(first, first + cnt)]
    assert got == list(range(lo // per, -(-hi // per)))
    assert _native_walk(tmp_path, 0, 5, 5, 64) == []                 # empty range
    assert _native_walk(tmp_path, 0, 0, 10 ** 9, 64)[-1][0] + _native_walk(tmp_path, 0, 0, 10 ** 9, 64)[-1][1] \
        == -(-n // per)                                              # clipped at nextNodeId
    # a missing partition file leaves a hole; a truncated one delivers its whole pages only
    os.remove(L.partition_path(str(tmp_path), "rawvec", 1))
    with open(L.partition_path(str(tmp_path), "rawvec", 2), "r+b") as f:
        f.truncate(16384 * 2 + 100)                                  # meta page + 1 data page + junk
    got = [p for first, cnt, _ in _native_walk(tmp_path, 0, 0, n, 64) for p in range(first, first + cnt)]
    want = [p for p in range(-(-n // per)) if p // 4 != 1 and not (p // 4 == 2 and p % 4 >= 1)]
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("prec", [onp.F64, onp.F32, onp.I8])
def test_load_and_search(tmp_path, prec):
    import tostore_b200 as T
    n, dims, k = 3000, 128, 10
    dead = np.zeros(n, dtype=bool)
    dead[[0, 5, 1234, n - 1]] = True
    rows = _make(tmp_path, n, dims, prec, metric="cosine", mpfs=8 * 16384, dead=dead)
    decoded = onp.decode_rows(onp.encode_rows(rows, prec), n, dims, prec)

    def make(meta):
        return T.GpuVectorIndex(meta.dimensions, meta.metric, capacity_rows=meta.next_node_id,
                                src_precision=meta.precision, k_max=16, nq_max=4)

    ix, meta = L.load_ngh_index(str(tmp_path), make, native=False)     # the Python walk
    with ix:
        st = ix.stats()
        assert st.rows == n and st.deleted_rows == int(dead.sum())
        q = onp.normalize_f32(np.random.default_rng(1).standard_normal(dims).astype(np.float32))
        ids, dist, cnt = ix.search(q, k)
        oi, od = oracle.search(decoded, q, 2, k, deleted=dead)
        assert cnt[0] == k and (ids[0] == oi).all()
        assert (dist[0].view(np.int64) == od.view(np.int64)).all()


# ---- property test: the library's meta.json parser against Python's json ------------------
from hypothesis import given, settings, strategies as st   # noqa: E402

_junk = st.recursive(st.none() | st.booleans() | st.integers(-10, 10 ** 12) | st.floats(allow_nan=False, allow_infinity=False)
                     | st.text(max_size=12),
                     lambda kids: st.lists(kids, max_size=3) | st.dictionaries(
                         st.sampled_from(["dimensions", "nextNodeId", "nghPageSize", "name", "x", "maxDegree"]),
                         kids, max_size=3), max_leaves=8)


@settings(max_examples=120, deadline=None)
@given(st.integers(1, 4096), st.sampled_from(["l2", "innerProduct", "cosine", "weird", None]),
       st.sampled_from(["float64", "float32", "int8", None]), st.integers(0, 10 ** 10),
       st.sampled_from([4096, 16384, 65536]), st.integers(1, 64),
       st.dictionaries(st.text(min_size=1, max_size=10).filter(
           lambda k: k not in {"dimensions", "distanceMetric", "precision", "nextNodeId", "nghPageSize",
                               "maxPartitionFileSize", "maxDegree"}), _junk, max_size=5),
       st.booleans())
def test_property_native_meta_parser_equals_python_json(tmp_path_factory, dims, metric, prec, nxt, ps, files,
                                                        extra, pretty):
    import json
    d = tmp_path_factory.mktemp("meta")
    (d / "ngh").mkdir()
    j = dict(extra)
    j.update({"dimensions": dims, "nextNodeId": nxt, "nghPageSize": ps, "maxPartitionFileSize": ps * files * 4})
    if metric is not None:
        j["distanceMetric"] = metric
    if prec is not None:
        j["precision"] = prec
    (d / "ngh" / "meta.json").write_text(json.dumps(j, indent=2 if pretty else None, ensure_ascii=not pretty))
    assert L.read_meta_native(str(d)) == L.read_meta(str(d))
