"""Cold start: stream an existing on-disk ToStore NGH vector index into a GPU index.

Host orchestration only — the bytes of the reference's partition files go to the
library untouched (`tsc_index_append_pages` validates magic / CRC / dims and decodes
on the GPU; `tsc_index_apply_graph_pages` reads the tombstone flags). Layout facts
restated from the reference (paths relative to /root/reference/lib/src):

  <index>/ngh/meta.json                      model/ngh_index_meta.dart:410-446
  <index>/ngh/rawvec/dir_<p // 500>/p<p>.ngh core/path_manager.dart:317-324,
  <index>/ngh/graph/dir_<p // 500>/p<p>.ngh  handler/common.dart:43 (maxEntriesPerDir)
  page 0 of every file = per-file meta page  core/ngh_page.dart:29-98
  pages per file  P = maxPartitionFileSize // nghPageSize   model/ngh_index_meta.dart:178
  nodeId -> logical page = nodeId // perPage; file = logical // P;
            local page = 1 + logical % P     model/ngh_index_meta.dart:451-490
  rows per raw page  = (pageSize-20-8-64) // (dims*bpe)     core/ngh_page.dart:575-579
  slots per graph page = (pageSize-20-4-64) // (2+4*maxDegree)   :559-566
Extents are derived from nextNodeId, never from the *PartitionCount fields (the
reference does not maintain them).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Callable, Iterator, Tuple

MAX_ENTRIES_PER_DIR = 500
_PRECISION = {"float64": 0, "float32": 1, "int8": 2}
_METRIC = {"l2": 0, "innerProduct": 1, "cosine": 2}
_BPE = {0: 8, 1: 4, 2: 1}


@dataclass(frozen=True)
class NghMeta:
    dimensions: int
    metric: int
    precision: int
    next_node_id: int
    page_size: int
    max_partition_file_size: int
    max_degree: int

    @property
    def pages_per_partition(self) -> int:
        return self.max_partition_file_size // self.page_size

    @property
    def vectors_per_raw_page(self) -> int:
        usable = self.page_size - 20 - 8 - 64
        vec = self.dimensions * _BPE[self.precision]
        return usable // vec if usable > 0 and vec > 0 else 0

    @property
    def nodes_per_graph_page(self) -> int:
        usable = self.page_size - 20 - 4 - 64
        return usable // (2 + self.max_degree * 4) if usable > 0 else 0


def read_meta(index_dir: str) -> NghMeta:
    with open(os.path.join(index_dir, "ngh", "meta.json"), "r", encoding="utf-8") as f:
        j = json.load(f)
    return NghMeta(
        dimensions=int(j["dimensions"]),
        metric=_METRIC.get(j.get("distanceMetric"), 2),          # default cosine (:503)
        precision=_PRECISION.get(j.get("precision"), 1),
        next_node_id=int(j.get("nextNodeId", 0)),
        page_size=int(j.get("nghPageSize", 16 * 1024)),
        max_partition_file_size=int(j.get("maxPartitionFileSize", 16 * 1024 * 1024)),
        max_degree=int(j.get("maxDegree", 64)),
    )


def partition_path(index_dir: str, category: str, partition: int) -> str:
    return os.path.join(index_dir, "ngh", category, f"dir_{partition // MAX_ENTRIES_PER_DIR}",
                        f"p{partition}.ngh")


def iter_partition_pages(index_dir: str, category: str, meta: NghMeta, per_page: int,
                         n_nodes: int) -> Iterator[Tuple[int, bytes]]:
    """Yield (first_logical_page, bytes of consecutive data pages) per partition file,
    covering node ids [0, n_nodes). Missing / short files yield what exists (the
    reference treats unreadable pages as empty, ngh_partition_manager.dart:262-287)."""
    if per_page <= 0 or n_nodes <= 0:
        return
    n_logical = -(-n_nodes // per_page)
    ppp, ps = meta.pages_per_partition, meta.page_size
    for part in range(-(-n_logical // ppp)):
        path = partition_path(index_dir, category, part)
        if not os.path.exists(path):
            continue
        want = min(ppp, n_logical - part * ppp)
        with open(path, "rb") as f:
            f.seek(ps)                                         # page 0 = per-file meta page
            data = f.read(want * ps)
        data = data[: len(data) // ps * ps]
        if data:
            yield part * ppp, data


def read_meta_native(index_dir: str) -> NghMeta:
    """meta.json through the library's own parser (`tsc_ngh_read_meta`)."""
    import ctypes as C

    from . import _native as N
    info = N.NghInfo(struct_size=C.sizeof(N.NghInfo))
    N.check(N.lib().tsc_ngh_read_meta(str(index_dir).encode("utf-8"), C.byref(info)),
            "tsc_ngh_read_meta")
    return NghMeta(dimensions=info.dims, metric=info.metric, precision=info.precision,
                   next_node_id=info.next_node_id, page_size=info.page_size,
                   max_partition_file_size=info.max_partition_file_size,
                   max_degree=info.max_degree)


def load_ngh_index(index_dir: str, make_index: Callable, load_tombstones: bool = True,
                   native: bool = True):
    """`make_index(meta)` must return a `GpuVectorIndex` whose dims / metric /
    src_precision match `meta` and whose capacity covers `meta.next_node_id` (or the
    shard's share of it). Returns (index, meta). native=True (default) uses the
    library's pipelined loader (`tsc_index_load_ngh`: reader thread + pinned double
    buffer); native=False walks the files from Python, one blocking call per file."""
    if native:
        meta = read_meta_native(index_dir)
        ix = make_index(meta)
        ix.load_ngh(index_dir, tombstones=load_tombstones)
        return ix, meta
    meta = read_meta(index_dir)
    ix = make_index(meta)
    for first_page, data in iter_partition_pages(index_dir, "rawvec", meta,
                                                 meta.vectors_per_raw_page, meta.next_node_id):
        ix.append_pages(data, first_page, meta.page_size, live_rows=meta.next_node_id)
    if load_tombstones:
        for first_page, data in iter_partition_pages(index_dir, "graph", meta,
                                                     meta.nodes_per_graph_page, meta.next_node_id):
            ix.apply_graph_pages(data, first_page, meta.page_size)
    return ix, meta
