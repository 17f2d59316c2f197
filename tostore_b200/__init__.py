"""tostore_b200 — B200-native exact vector search behind ToStore's `vectorSearch`.

Only what the hot path needs: `csrc/` (sm_100a kernels + the C ABI,
built into `libtostore_cuda.so`), `_native` (ctypes binding), `engine`
(`GpuVectorIndex`, the stand-in for `NghGraphEngine`) and `vector_store`
(host mirror of the reference's vector API). No CPU fallback exists.
"""
from .engine import (DEV_BF16, DEV_F16, DEV_F32, METRIC_COSINE, METRIC_INNER_PRODUCT,
                     METRIC_L2, SRC_F32, SRC_F64, SRC_I8, GpuVectorIndex)
from .vector_store import (DeviceDType, GpuVectorStore, VectorData, VectorDistanceMetric,
                           VectorFieldConfig, VectorIndexConfig, VectorPrecision,
                           VectorSearchResult)
from .where import QueryCondition, compile_condition
from ._native import LIB_PATH, TscError

__all__ = [
    "GpuVectorIndex", "GpuVectorStore", "VectorData", "VectorDistanceMetric",
    "VectorFieldConfig", "VectorIndexConfig", "VectorPrecision", "VectorSearchResult",
    "DeviceDType", "TscError", "LIB_PATH", "QueryCondition", "compile_condition",
    "METRIC_L2", "METRIC_INNER_PRODUCT", "METRIC_COSINE", "SRC_F64", "SRC_F32", "SRC_I8",
    "DEV_F32", "DEV_BF16", "DEV_F16",
]
