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

    ix, meta = L.load_ngh_index(str(tmp_path), make)
    with ix:
        st = ix.stats()
        assert st.rows == n and st.deleted_rows == int(dead.sum())
        q = onp.normalize_f32(np.random.default_rng(1).standard_normal(dims).astype(np.float32))
        ids, dist, cnt = ix.search(q, k)
        oi, od = oracle.search(decoded, q, 2, k, deleted=dead)
        assert cnt[0] == k and (ids[0] == oi).all()
        assert (dist[0].view(np.int64) == od.view(np.int64)).all()
