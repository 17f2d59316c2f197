"""nodeId -> primary key side table (role of the reference's `__nid2pk` B+Tree,
core/vector_index_manager.dart:553-588, :1276-1293): host-memory logic of the library,
exercised WITHOUT a GPU through the self-test hooks (a host-only index object and the
result-assembly step of tsc_vector_search_pk)."""
import ctypes as C

import numpy as np
import pytest

from tostore_b200 import GpuVectorIndex, TscError
from tostore_b200 import _native as N


class HostIndex(GpuVectorIndex):
    """GpuVectorIndex bound to a host-only self-test handle (primary-key table only)."""

    def __init__(self, capacity, first_node_id=0):
        self._lib = N.lib()
        self.first_node_id = first_node_id
        h = C.c_uint64(0)
        N.check(self._lib.tsc_selftest_host_index(capacity, first_node_id, C.byref(h)), "host_index")
        self.handle = h.value

    def assemble(self, ids, dist, score, k, cap=4096):
        ids = np.array(list(ids) + [-1] * (k - len(ids)), dtype=np.int64)
        n = C.c_uint32(len(dist))
        dist = np.array(list(dist) + [np.nan] * (k - len(dist)), dtype=np.float64)
        score = np.array(list(score) + [np.nan] * (k - len(score)), dtype=np.float64)
        pkb = np.zeros(cap, dtype=np.uint8)
        offs = np.zeros(k + 1, dtype=np.uint64)
        N.check(self._lib.tsc_selftest_pk_assemble(self.handle, k, ids.ctypes.data, dist.ctypes.data,
                                                   score.ctypes.data, pkb.ctypes.data, cap,
                                                   offs.ctypes.data, C.byref(n)), "pk_assemble")
        raw = pkb.tobytes()
        pks = [raw[int(offs[i]): int(offs[i + 1])].decode() for i in range(n.value)]
        return pks, ids, dist, score, n.value, offs


def test_set_get_overwrite_and_bounds():
    with HostIndex(100, first_node_id=1000) as ix:
        pks = [f"user_{i}" if i % 4 else f"用户-{i}" for i in range(60)]
        ix.set_primary_keys(pks, first_node_id=1000)
        assert [ix.get_primary_key(1000 + i) for i in range(60)] == pks
        assert ix.get_primary_key(999) is None and ix.get_primary_key(1060) is None
        assert ix.get_primary_key(5_000_000) is None
        ix.set_primary_keys(["replaced", None, ""], first_node_id=1010)        # update + tombstones
        assert ix.get_primary_key(1010) == "replaced"
        assert ix.get_primary_key(1011) is None and ix.get_primary_key(1012) is None
        assert ix.get_primary_key(1013) == pks[13]
        ix.set_primary_keys(["tail"], first_node_id=1099)                       # sparse: gap stays unmapped
        assert ix.get_primary_key(1099) == "tail" and ix.get_primary_key(1080) is None
        with pytest.raises(TscError):
            ix.set_primary_keys(["x"], first_node_id=1100)                      # beyond capacity
        with pytest.raises(TscError):
            ix.set_primary_keys(["x"], first_node_id=10)                        # below the shard
        ix.set_primary_keys([], first_node_id=1000)                             # no-op


def test_result_assembly_drops_unmapped_and_keeps_order():
    with HostIndex(50) as ix:
        ix.set_primary_keys([f"pk{i}" for i in range(40)])
        ix.set_primary_keys([None], first_node_id=7)                            # tombstone mapping
        hits = [5, 7, 39, 45, 0]                                                # 7 tombstoned, 45 unmapped
        dist = [0.1, 0.2, 0.3, 0.4, 0.5]
        score = [0.9, 0.8, 0.7, 0.6, 0.5]
        pks, ids, d, s, n, offs = ix.assemble(hits, dist, score, k=8)
        assert n == 3 and pks == ["pk5", "pk39", "pk0"]
        assert ids[:3].tolist() == [5, 39, 0] and d[:3].tolist() == [0.1, 0.3, 0.5]
        assert s[:3].tolist() == [0.9, 0.7, 0.5]
        assert (ids[3:] == -1).all() and np.isnan(d[3:]).all() and np.isnan(s[3:]).all()
        assert offs.tolist() == [0, 3, 7, 10] + [10] * 5
        with pytest.raises(TscError):
            ix.assemble(hits, dist, score, k=8, cap=5)                          # keys do not fit
        pks, *_ , n, _ = ix.assemble([], [], [], k=4)
        assert n == 0 and pks == []


def test_host_only_handle_refuses_compute():
    with HostIndex(10) as ix:
        with pytest.raises(TscError):
            ix.search(np.zeros((1, 4), dtype=np.float32), 1) if hasattr(ix, "dims") else \
                N.check(ix._lib.tsc_search(ix.handle, None, 1, 1, float("nan"), None, None, None), "tsc_search")


def test_concurrent_callers_share_one_handle_safely():
    """Thread-safe per handle (internal mutex): many threads hammer set / get / assemble on
    one host-only handle while others translate + evaluate WHERE programs; every reader must
    see either the old or the new key of a row, never a torn one."""
    import threading
    from tostore_b200 import where as W
    n, rounds = 64, 300
    errors = []
    with HostIndex(n) as ix:
        ix.set_primary_keys([f"a{i:04d}" for i in range(n)])

        def writer(tag):
            try:
                for r in range(rounds):
                    ix.set_primary_keys([f"{tag}{(r + i) % 10000:04d}" for i in range(8)], first_node_id=(r * 8) % (n - 8))
            except Exception as e:          # noqa: BLE001
                errors.append(e)

        def reader():
            try:
                for r in range(rounds):
                    pk = ix.get_primary_key(r % n)
                    assert pk is not None and len(pk) == 5 and pk[0] in "abc" and pk[1:].isdigit(), pk
                    pks, *_ = ix.assemble([r % n, (r + 1) % n], [0.1, 0.2], [0.9, 0.8], k=4)
                    assert len(pks) == 2 and all(len(p) == 5 for p in pks)
            except Exception as e:          # noqa: BLE001
                errors.append(e)

        def evaluator():
            try:
                import ctypes as C
                from tostore_b200 import _native as N
                prog = W.compile_condition({"x": {"IN": [1, 2, 3]}}, {"x": (0, W.COL_I64)})
                ops, n_ops, raw, n_args = prog.buffers()
                ids = np.array([0], dtype=np.uint32)
                types = np.array([0], dtype=np.uint8)
                vals = np.arange(16, dtype=np.uint64)
                nulls = np.zeros(16, dtype=np.uint8)
                out = np.zeros(16, dtype=np.uint8)
                for _ in range(rounds):
                    N.check(N.lib().tsc_selftest_where(C.cast(ops, C.c_void_p), n_ops, raw.ctypes.data, n_args, 1,
                                                       ids.ctypes.data, types.ctypes.data, vals.ctypes.data,
                                                       nulls.ctypes.data, 16, out.ctypes.data), "where")
                    assert out.tolist() == [0, 1, 1, 1] + [0] * 12
            except Exception as e:          # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=writer, args=("b",)), threading.Thread(target=writer, args=("c",)),
                   threading.Thread(target=reader), threading.Thread(target=reader),
                   threading.Thread(target=evaluator), threading.Thread(target=evaluator)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    assert not errors, errors[:3]


def _pk_bitmap(ix, pks, n_rows):
    """tsc_selftest_pk_filter_bitmap: the bitmap tsc_index_filter_primary_keys would install."""
    blob, offs, n = ix._key_blob(pks)
    words = np.zeros((n_rows + 63) // 64 + 2, dtype=np.uint64)
    matched = C.c_uint64(0)
    N.check(ix._lib.tsc_selftest_pk_filter_bitmap(ix.handle, blob.ctypes.data, offs.ctypes.data, n,
                                                  words.ctypes.data, words.size, C.byref(matched)),
            "pk_filter_bitmap")
    bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n_rows].astype(bool)
    assert not np.unpackbits(words.view(np.uint8), bitorder="little")[n_rows:].any()
    return bits, matched.value


def test_filter_by_primary_keys_bitmap():
    """The `__pk2nid` role (core/vector_index_manager.dart:1350-1363): a set of primary keys -> the
    rows that stay searchable. Unknown / empty keys are ignored, duplicates count once, tombstoned
    mappings are not selectable, a re-mapped key selects the highest node id, and the reverse map
    follows later changes of the table."""
    n = 300
    with HostIndex(400, first_node_id=50) as ix:
        pks = [f"k{i}" if i % 7 else f"键-{i}" for i in range(n)]
        ix.set_primary_keys(pks, first_node_id=50)
        rng = np.random.default_rng(4)
        pick = sorted(set(int(i) for i in rng.integers(0, n, 90)))
        bits, matched = _pk_bitmap(ix, [pks[i] for i in pick] + ["nobody", "", None, pks[pick[0]]], n)
        assert matched == len(pick) and np.nonzero(bits)[0].tolist() == pick
        bits, matched = _pk_bitmap(ix, [], n)
        assert matched == 0 and not bits.any()
        # tombstone a selected row, re-map another key to a later node id
        ix.set_primary_keys([""], first_node_id=50 + pick[1])
        ix.set_primary_keys([pks[pick[2]]], first_node_id=50 + n)          # same key, node id 350
        bits, matched = _pk_bitmap(ix, [pks[i] for i in pick], n + 1)
        want = [i for i in pick if i not in (pick[1], pick[2])] + [n]
        assert np.nonzero(bits)[0].tolist() == sorted(want) and matched == len(want)
        with pytest.raises(TscError):                                       # device entry point: no GPU behind it
            ix.filter_primary_keys(["k1"])
