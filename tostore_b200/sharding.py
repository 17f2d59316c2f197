"""Row-range sharding of one embedding column across the GPUs of a box.

Host-side logic only: which node-id range a rank owns, and how the
per-shard exact top-k lists combine into the global top-k. The reference has no
distributed code at all; the exchange here is the single step the path needs —
an all-gather of k (distance, nodeId) pairs per query per shard, followed by the
same ordering rule as the reference's final sort
(/root/reference/lib/src/core/vector_index_manager.dart:587: ascending distance,
Dart `double.compareTo` order, ties by node id).

Two ways to run the exchange:
  * in the library: `tsc_comm_init` + `tsc_search_sharded` (ncclAllGather on the
    search stream + the merge kernel) — what bench.py uses on B200s;
  * on the host through any `torch.distributed` backend (`ShardedSearcher`), which
    is also what the gloo world_size=2 CPU tests drive.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_rows(n_total: int, world: int, rank: int, align: int = 32) -> Tuple[int, int]:
    """Contiguous node-id range [lo, hi) of `rank`; boundaries are multiples of
    `align` so that bitmap words (and, with align = lcm(32, rows per page), whole
    reference pages) never straddle two GPUs."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    per = -(-n_total // world)
    per = -(-per // align) * align
    lo = min(n_total, rank * per)
    hi = min(n_total, (rank + 1) * per)
    return lo, hi


def _order_key(d: np.ndarray) -> np.ndarray:
    """float64 -> int64 keys ordered like Dart's double.compareTo (-0.0 < 0.0, NaN last)."""
    d = np.ascontiguousarray(d, dtype=np.float64)
    bits = d.view(np.int64).copy()
    bits[np.isnan(d)] = np.int64(0x7FF8000000000000)
    neg = bits < 0
    bits[neg] = np.int64(-(2 ** 63)) - bits[neg] - np.int64(1)
    return bits


def merge_topk(part_ids: np.ndarray, part_dist: np.ndarray, k: int):
    """[P, nq, k] per-shard results (-1 = empty slot) -> global (ids, dist, counts)."""
    part_ids = np.asarray(part_ids, dtype=np.int64)
    part_dist = np.asarray(part_dist, dtype=np.float64)
    p, nq, kk = part_ids.shape
    ids = np.full((nq, k), -1, dtype=np.int64)
    dist = np.full((nq, k), np.nan, dtype=np.float64)
    counts = np.zeros(nq, dtype=np.uint32)
    for q in range(nq):
        ci = part_ids[:, q, :].reshape(-1)
        cd = part_dist[:, q, :].reshape(-1)
        keep = ci >= 0
        ci, cd = ci[keep], cd[keep]
        order = np.lexsort((ci, _order_key(cd)))[:k]
        n = len(order)
        ids[q, :n], dist[q, :n], counts[q] = ci[order], cd[order], n
    return ids, dist, counts


class ShardedSearcher:
    """One rank's view of a sharded index: local search + all-gather + merge.

    `local_search(queries, k, threshold) -> (ids[nq,k], dist[nq,k], counts[nq])`
    is `GpuVectorIndex.search` of the rank's shard in production."""

    def __init__(self, local_search: Callable, group=None):
        import torch.distributed as dist
        self._dist = dist
        self._search = local_search
        self._group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def search(self, queries, k: int, threshold: Optional[float] = None):
        import torch
        ids, dist_, _ = self._search(queries, k, threshold)
        ids = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
        dd = torch.from_numpy(np.ascontiguousarray(dist_, dtype=np.float64))
        g_ids = [torch.empty_like(ids) for _ in range(self.world)]
        g_dd = [torch.empty_like(dd) for _ in range(self.world)]
        self._dist.all_gather(g_ids, ids, group=self._group)      # the one exchange step
        self._dist.all_gather(g_dd, dd, group=self._group)
        return merge_topk(torch.stack(g_ids).numpy(), torch.stack(g_dd).numpy(), k)
