"""Host mirror of ToStore's vector API for the GPU path.

Keeps the reference's names, argument meaning and error behaviour for the hot
path only (citations relative to /root/reference/lib):

  * `VectorData`, `VectorFieldConfig`, `VectorPrecision`, `VectorDistanceMetric`,
    `VectorIndexConfig`                     src/model/table_schema.dart:2109-2675
  * `VectorSearchResult`                    src/model/query_result.dart:207-228
  * `ToStore.vectorSearch(tableName, fieldName:, queryVector:, topK: 10,
    efSearch:, distanceThreshold:)`         tostore.dart:493-511
  * `VectorIndexManager.vectorSearch / writeChanges`
                                            src/core/vector_index_manager.dart:297-589

What is *not* here: tables, B+Trees, WAL, transactions, schema migration. Rows
become visible to search as soon as they are appended (the reference makes them
visible at flush); nodeId = insertion order, exactly as
`meta.nextNodeId++` (src/core/ngh_graph_engine.dart:321-322). The nodeId -> primary
key B+Tree of the reference (`__nid2pk`) is a dense host-memory table inside the
library (`tsc_index_set_primary_keys`, `tsc_vector_search_pk`).

The Dart production binding is dart/tostore_cuda_bindings.dart; this module is
the equivalent host layer in the one host language this image can run.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import where as _where
from .engine import GpuVectorIndex


class VectorPrecision(enum.IntEnum):       # table_schema.dart:2481-2498 (enum order)
    float64 = 0
    float32 = 1
    int8 = 2


class VectorDistanceMetric(enum.IntEnum):  # table_schema.dart:2511-2531 (enum order)
    l2 = 0
    innerProduct = 1
    cosine = 2


class DeviceDType(enum.IntEnum):           # new: storage type of the column in HBM
    float32 = 0
    bfloat16 = 1
    float16 = 2


@dataclass(frozen=True)
class VectorData:                          # table_schema.dart:2109-2401
    values: Sequence[float]

    @staticmethod
    def fromList(values: Iterable[float]) -> "VectorData":
        return VectorData(tuple(float(v) for v in values))

    @property
    def dimensions(self) -> int:
        return len(self.values)


@dataclass(frozen=True)
class VectorFieldConfig:                   # table_schema.dart:2406-2475
    dimensions: int
    precision: VectorPrecision = VectorPrecision.float64


@dataclass(frozen=True)
class VectorIndexConfig:                   # table_schema.dart:2547-2597
    distanceMetric: VectorDistanceMetric = VectorDistanceMetric.cosine
    maxDegree: Optional[int] = None        # accepted for API compatibility; the
    efSearch: Optional[int] = None         # exact scan has no graph and ignores
    constructionEf: Optional[int] = None   # these four
    pruneAlpha: Optional[float] = None
    pqSubspaces: Optional[int] = None


@dataclass(frozen=True)
class VectorSearchResult:                  # query_result.dart:207-228
    primaryKey: str
    distance: float
    score: float

    def toJson(self) -> Dict[str, object]:
        return {"primaryKey": self.primaryKey, "distance": self.distance, "score": self.score}


@dataclass
class _VectorIndex:
    table: str
    fieldName: str
    field: VectorFieldConfig
    config: VectorIndexConfig
    engine: GpuVectorIndex
    nid2pk: List[Optional[str]] = field(default_factory=list)   # role of `__nid2pk`
    pk2nid: Dict[str, int] = field(default_factory=dict)        # role of `__pk2nid`
    # table fields mirrored column-wise on the GPU for WHERE: name -> (column id, type)
    attributes: Dict[str, tuple] = field(default_factory=dict)


class GpuVectorStore:
    """The slice of `ToStore` / `VectorIndexManager` that serves `vectorSearch`."""

    def __init__(self, device_id: int = 0, capacity_rows: int = 1 << 20,
                 device_dtype: DeviceDType = DeviceDType.float32, k_max: int = 128):
        self._device_id = device_id
        self._capacity = capacity_rows
        self._dtype = device_dtype
        self._k_max = k_max
        self._indexes: Dict[str, List[_VectorIndex]] = {}

    # -- schema -------------------------------------------------------------------------
    def createVectorIndex(self, tableName: str, fieldName: str, fieldConfig: VectorFieldConfig,
                          indexConfig: Optional[VectorIndexConfig] = None,
                          attributeFields: Optional[Dict[str, str]] = None) -> None:
        """`TableSchema` vector field + `IndexSchema(type: IndexType.vector)`.
        Unlike the reference (no validation) bad dims raise.
        attributeFields (new, additive): fields of the table — name -> 'integer' | 'double' |
        'text' | 'boolean' | 'datetime' (DataType names, model/table_schema.dart) — kept
        column-wise on the GPU (text: dictionary-encoded; boolean: 0 / 1; datetime: its stored
        ISO-8601 string, compared as text like the reference does) so
        `vectorSearch(where=...)` can prefilter on them."""
        cfg = indexConfig or VectorIndexConfig()
        eng = GpuVectorIndex(fieldConfig.dimensions, int(cfg.distanceMetric),
                             capacity_rows=self._capacity, src_precision=1,
                             dev_dtype=int(self._dtype), device_id=self._device_id,
                             k_max=self._k_max, nq_max=64)
        ix = _VectorIndex(tableName, fieldName, fieldConfig, cfg, eng)
        for cid, (name, dtype) in enumerate((attributeFields or {}).items()):
            kinds = {"integer": _where.COL_I64, "double": _where.COL_F64, "text": _where.COL_TEXT,
                     "boolean": _where.COL_BOOL, "datetime": _where.COL_DATETIME}
            if dtype not in kinds:
                raise ValueError(f"attribute field {name!r}: only {' / '.join(kinds)} fields have a GPU column")
            t = kinds[dtype]
            eng.column_create(cid, {_where.COL_BOOL: _where.COL_I64,
                                    _where.COL_DATETIME: _where.COL_TEXT}.get(t, t))
            ix.attributes[name] = (cid, t)
        self._indexes.setdefault(tableName, []).append(ix)

    def _find(self, tableName: str, fieldName: str) -> Optional[_VectorIndex]:
        for ix in self._indexes.get(tableName, ()):      # vector_index_manager.dart:484-499
            if ix.fieldName == fieldName:
                return ix
        return None

    # -- writes (VectorIndexManager.writeChanges, :297-466) -----------------------------
    @staticmethod
    def _to_float32(values, dims: int) -> np.ndarray:
        """`_toFloat32` (compute/vector_batch_prepare_compute.dart:79-86): truncate or
        zero-pad, fp64 -> fp32 round-to-nearest-even."""
        out = np.zeros(dims, dtype=np.float32)
        v = np.asarray(values.values if isinstance(values, VectorData) else values,
                       dtype=np.float64)[:dims]
        with np.errstate(over="ignore"):
            out[: v.size] = v.astype(np.float32)
        return out

    @staticmethod
    def _store_round_trip(rows: np.ndarray, precision: VectorPrecision) -> np.ndarray:
        """What the reference keeps on disk for `precision` and reads back as fp32
        (`setVectorFromFloat32` / `getVectorAsFloat32`, core/ngh_page.dart:368-412)."""
        if precision == VectorPrecision.int8:
            c = np.clip(rows.astype(np.float64), -1.0, 1.0) * 127.0
            q = np.where(c >= 0, np.floor(c + 0.5), np.ceil(c - 0.5))
            return (q / 127.0).astype(np.float32)
        return rows                                     # f32 exact; f64 holds the fp32 value

    def batchInsert(self, tableName: str, records: Sequence[Dict[str, object]],
                    primaryKey: str = "id") -> int:
        """Insert records; every vector index of the table receives the rows.
        Records without the vector field or with an empty key are skipped
        (`prepareVectorBatchChunk`, compute/vector_batch_prepare_compute.dart:34-67)."""
        n_done = 0
        for ix in self._indexes.get(tableName, ()):
            vecs, pks, kept = [], [], []
            for rec in records:
                val = rec.get(ix.fieldName)
                if val is None:
                    continue
                pk = rec.get(primaryKey)
                if pk is None or str(pk) == "":
                    continue
                vecs.append(self._to_float32(val, ix.field.dimensions))
                pks.append(str(pk))
                kept.append(rec)
            if not vecs:
                continue
            rows = self._store_round_trip(np.stack(vecs), ix.field.precision)
            start = len(ix.nid2pk)                      # startNodeId = meta.nextNodeId
            ix.engine.append_rows(rows, first_node_id=start)
            for name, (cid, t) in ix.attributes.items():
                ix.engine.column_append(cid, [_where._convert(r[name], t) if r.get(name) is not None
                                              else None for r in kept], first_node_id=start)
            ix.engine.set_primary_keys(pks, first_node_id=start)     # `__nid2pk` deltas (:1276-1293)
            for j, pk in enumerate(pks):
                ix.nid2pk.append(pk)
                ix.pk2nid[pk] = start + j
            n_done = max(n_done, len(vecs))
        return n_done

    def insert(self, tableName: str, record: Dict[str, object], primaryKey: str = "id") -> int:
        return self.batchInsert(tableName, [record], primaryKey)

    def update(self, tableName: str, primaryKey: object, record: Dict[str, object]) -> int:
        """Overwrite the embedding and / or attribute fields of an existing row in place
        (same nodeId). Additive: the reference does not forward embedding updates to the
        vector index at all (core/index_manager.dart:3125-3133); here
        the row in HBM, its norm and its attribute columns are rewritten, so the next
        search sees the new values. Fields absent from `record` keep their value."""
        n = 0
        for ix in self._indexes.get(tableName, ()):
            nid = ix.pk2nid.get(str(primaryKey))
            if nid is None:
                continue
            val = record.get(ix.fieldName)
            if val is not None:
                row = self._store_round_trip(self._to_float32(val, ix.field.dimensions)[None, :],
                                             ix.field.precision)
                ix.engine.append_rows(row, first_node_id=nid)
            for name, (cid, t) in ix.attributes.items():
                if name in record:
                    v = record[name]
                    ix.engine.column_append(cid, [None if v is None else _where._convert(v, t)],
                                            first_node_id=nid)
            n += 1
        return n

    def delete(self, tableName: str, primaryKeys: Iterable[object]) -> int:
        """Tombstone rows (deleteBatch, ngh_graph_engine.dart:411-445; mapping
        tombstone, vector_index_manager.dart:416-434)."""
        n = 0
        for ix in self._indexes.get(tableName, ()):
            nids = []
            for pk in primaryKeys:
                nid = ix.pk2nid.pop(str(pk), None)
                if nid is not None:
                    ix.nid2pk[nid] = None
                    nids.append(nid)
            if nids:
                ix.engine.set_deleted(nids, True)
                for nid in nids:                                 # tombstone mapping, value [1]
                    ix.engine.set_primary_keys([None], first_node_id=nid)
                n = max(n, len(nids))
        return n

    def setWhereFilter(self, tableName: str, fieldName: str,
                       primaryKeys: Optional[Iterable[object]]) -> None:
        """WHERE prefilter (new, additive): restrict the next searches to these keys."""
        ix = self._find(tableName, fieldName)
        if ix is None:
            return
        if primaryKeys is None:
            ix.engine.set_filter(None)
            return
        mask = np.zeros(len(ix.nid2pk), dtype=bool)
        for pk in primaryKeys:
            nid = ix.pk2nid.get(str(pk))
            if nid is not None:
                mask[nid] = True
        ix.engine.set_filter(mask)

    # -- ToStore.vectorSearch (tostore.dart:493-511) --------------------------------------
    def vectorSearch(self, tableName: str, *, fieldName: str, queryVector: VectorData,
                     topK: int = 10, efSearch: Optional[int] = None,
                     distanceThreshold: Optional[float] = None,
                     where=None) -> List[VectorSearchResult]:
        """`where` (new, additive): a `QueryCondition` / its map form over the index's
        attribute fields; evaluated on the GPU into the prefilter bitmap before the scan.
        `where=None` searches every live row (and clears a previous condition)."""
        ix = self._find(tableName, fieldName)
        if ix is None or not ix.nid2pk:                  # :485-504 -> const []
            return []
        _ = efSearch                                     # exact scan: no expansion factor
        if where is not None:
            cond = where.build() if hasattr(where, "build") else where
            ix.engine.filter_where(_where.compile_condition(cond, ix.attributes))
            ix.where_active = True
        elif getattr(ix, "where_active", False):
            ix.engine.set_filter(None)
            ix.where_active = False
        values = queryVector.values if isinstance(queryVector, VectorData) else queryVector
        # query prep, search, nodeId -> PK (:553-588) and score all happen inside the library
        pks, _ids, dist, score = ix.engine.vector_search_pk(values, topK, distanceThreshold)
        return [VectorSearchResult(primaryKey=pk, distance=d, score=s)
                for pk, d, s in zip(pks, dist.tolist(), score.tolist())]   # ascending (:587)

    def vectorSearchBatch(self, tableName: str, *, fieldName: str, queryVectors: Sequence[VectorData],
                          topK: int = 10, distanceThreshold: Optional[float] = None
                          ) -> List[List[VectorSearchResult]]:
        """Batch form of `vectorSearch` (additive; the reference's API is single-query): one
        result list per query vector, each exactly what `vectorSearch` would return. Queries
        are truncated / zero-padded to the field's dimensions like single ones."""
        ix = self._find(tableName, fieldName)
        if ix is None or not ix.nid2pk or not queryVectors:
            return [[] for _ in queryVectors]
        dims = ix.field.dimensions
        vals = np.zeros((len(queryVectors), dims), dtype=np.float64)
        for i, qv in enumerate(queryVectors):
            v = np.asarray(qv.values if isinstance(qv, VectorData) else qv, dtype=np.float64)[:dims]
            vals[i, : v.size] = v
        out: List[List[VectorSearchResult]] = []
        step = ix.engine.nq_max
        for b in range(0, len(queryVectors), step):
            ids, dist, score, counts = ix.engine.vector_search_batch(vals[b: b + step], topK, distanceThreshold)
            for i in range(ids.shape[0]):
                res = []
                for j in range(int(counts[i])):
                    nid = int(ids[i, j])
                    pk = ix.nid2pk[nid] if 0 <= nid < len(ix.nid2pk) else None
                    if pk is None:                               # :578-579 (deleted / unmapped)
                        continue
                    res.append(VectorSearchResult(primaryKey=pk, distance=float(dist[i, j]),
                                                  score=float(score[i, j])))
                out.append(res)
        return out

    def close(self) -> None:
        for lst in self._indexes.values():
            for ix in lst:
                ix.engine.close()
        self._indexes.clear()
