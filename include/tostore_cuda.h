/*
 * tostore_cuda.h — C ABI of libtostore_cuda.so, the B200 (sm_100a) exact
 * vector-search executor that drops in underneath ToStore's `vectorSearch`.
 *
 * This is the boundary the reference's Dart host binds with dart:ffi
 * (binding source: dart/tostore_cuda_bindings.dart, wiring: INTEGRATION.md) and
 * that this repo's Python host mirror binds with ctypes. Plain pointers and
 * sizes only: no C++ types, no exceptions, no torch types, no stdout.
 *
 * Reference interfaces replaced / consumed (paths relative to
 * /root/reference/lib/src, tocreator/tostore @ 130da06):
 *   - NghGraphEngine.search                 core/ngh_graph_engine.dart:67-135
 *     (call site core/vector_index_manager.dart:538-548)       -> tsc_search*
 *   - _exactDistance/_l2Distance/_innerProduct/_cosineSimlarity :908-946
 *   - result ordering :133-134, vector_index_manager.dart:587
 *   - VectorIndexManager._toFloat32/_normalizeFloat32/_distanceToScore
 *     core/vector_index_manager.dart:1385-1423                  -> tsc_vector_search
 *   - NghRawVectorPage / BTreePageIO page format
 *     core/ngh_page.dart:310-450, core/btree_page.dart:132-234  -> tsc_index_append_pages
 *   - tombstones NghNodeFlags.deleted core/ngh_page.dart:104-108,
 *     deleteBatch core/ngh_graph_engine.dart:411-445            -> tsc_index_set_deleted,
 *                                                                  tsc_index_apply_graph_pages
 *   - nodeId -> (partition,page,slot) model/ngh_index_meta.dart:451-490
 *   - ConditionRecordMatcher (condition tree + operators over numeric fields)
 *     handler/value_matcher.dart:337-625, query/query_condition.dart:486-520
 *                                                              -> tsc_index_filter_where
 *
 * Conventions (mirroring lib/src/handler/system_ffi_helper.dart): every buffer
 * is caller-allocated and caller-freed; int32 status, 0 = success, negative =
 * tsc_status; handles are opaque uint64 (safe to pass between isolates).
 * Thread-safe per handle (internal mutex): searches on one handle are serialised —
 * a second blocking search waits for the first, a second tsc_search_submit first
 * retires the ticket in flight (its results are delivered to its buffers) — and
 * device-buffer searches on different streams are ordered by an event, because a
 * handle owns one set of search scratch. No exception crosses the boundary.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point
 * returns TSC_ERR_CUDA.
 *
 * One process can own all GPUs of a box: tsc_index_create with n_devices = 2..8
 * returns a GROUP handle — the embedding column is row-range sharded over the
 * devices, every entry point that takes HOST buffers (append, load, filter, WHERE,
 * tsc_search, tsc_vector_search*, primary keys, stats) works on it unchanged, and
 * one tsc_search call scans all shards and merges their exact top-k on the first
 * device through NVLink peer memory. That is the form the single-process Dart host
 * binds (call site core/vector_index_manager.dart:538-548). The one-process-per-GPU
 * form (tsc_comm_*, tsc_search_sharded) exists for torchrun-style launches.
 *
 * Exactness: candidates are chosen by an fp32 key and re-ranked in exact fp64; the
 * library PROVES per query that no row outside the candidates can belong to the
 * result (certificate) and otherwise re-runs the query in range mode,
 * which collects every row the key error bound cannot exclude. tsc_stats reports
 * certified / retried / uncertified query counts.
 */
#ifndef TOSTORE_CUDA_H
#define TOSTORE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSC_ABI_VERSION 2

typedef enum tsc_status {
  TSC_OK = 0,
  TSC_ERR_BAD_HANDLE = -1,
  TSC_ERR_BAD_ARG = -2,      /* NULL pointer, k == 0, k > k_max, nq > nq_max ... */
  TSC_ERR_BAD_DIMS = -3,     /* dims == 0, dims too large, page dims mismatch     */
  TSC_ERR_OOM = -4,          /* host or device allocation failed / capacity full  */
  TSC_ERR_CUDA = -5,         /* CUDA runtime / driver error, no device            */
  TSC_ERR_NCCL = -6,         /* NCCL missing or failed                            */
  TSC_ERR_PAGE = -7,         /* bad magic / header / length / CRC / type          */
  TSC_ERR_UNSUPPORTED = -8,
  TSC_ERR_NOT_READY = -9     /* poll: ticket still running                        */
} tsc_status;

/* VectorDistanceMetric enum order, model/table_schema.dart:2511-2531 */
enum { TSC_METRIC_L2 = 0, TSC_METRIC_INNER_PRODUCT = 1, TSC_METRIC_COSINE = 2 };
/* VectorPrecision enum order, model/table_schema.dart:2481-2498 */
enum { TSC_SRC_F64 = 0, TSC_SRC_F32 = 1, TSC_SRC_I8 = 2 };
/* storage type of the embedding column in HBM (new; reference has fp32 only) */
enum { TSC_DEV_F32 = 0, TSC_DEV_BF16 = 1, TSC_DEV_F16 = 2 };

typedef struct tsc_index_desc {
  uint32_t struct_size;    /* sizeof(tsc_index_desc), for forward compatibility */
  uint32_t dims;           /* NghIndexMeta.dimensions                            */
  uint8_t metric;          /* TSC_METRIC_*                                       */
  uint8_t src_precision;   /* TSC_SRC_*: element type of append_rows / pages     */
  uint8_t dev_dtype;       /* TSC_DEV_*                                          */
  uint8_t reserved0;
  int32_t device_id;       /* CUDA device ordinal that holds this shard          */
  uint64_t capacity_rows;  /* rows reserved in HBM for this shard; an UNSHARDED   */
                           /* column grows beyond it on demand (append_rows, ...) */
  uint64_t first_node_id;  /* nodeId of shard row 0 (row-range sharding)         */
  uint32_t k_max;          /* largest topK that will be requested (<= 128)       */
  uint32_t nq_max;         /* largest query batch per call                       */
  uint32_t n_devices;      /* 0 / 1: one shard on device_id. 2..8: a GROUP — the  */
  int32_t device_ids[8];   /* column is row-range sharded over device_ids[0..n),  */
  uint32_t reserved1;      /* capacity_rows / first_node_id describe the WHOLE    */
} tsc_index_desc;          /* column; device_id is ignored                        */

typedef struct tsc_stats {
  uint32_t struct_size;
  uint32_t dims;
  uint64_t rows;             /* rows appended ([0, nextNodeId) of this shard)  */
  uint64_t deleted_rows;
  uint64_t device_bytes;     /* HBM held by this index                         */
  uint64_t row_stride_bytes; /* bytes per row in HBM                           */
  uint64_t searches;         /* completed search calls                         */
  uint64_t kernel_launches;  /* kernels launched by this index since creation  */
  double last_search_ms;     /* device time of the last search (CUDA events)   */
  double last_scan_gbs;      /* algorithmic row bytes / last_search_ms         */
  uint32_t last_path;        /* 1 = HBM scan, 2 = tcgen05 GEMM                 */
  uint32_t reserved;
  /* dominant-kernel accounting since creation / tsc_stats_reset: every scan (or
   * GEMM) launch is bracketed by CUDA events on its own stream */
  uint64_t hot_launches;     /* timed launches of the dominant kernel          */
  double hot_ms_total;       /* sum of their device durations                  */
  double hot_bytes_total;    /* algorithmic bytes they covered (rows*dims*elem)*/
  double hot_flops_total;    /* algorithmic flops (GEMM path), else 0          */
  /* exactness certificate (per query, since creation / tsc_stats_reset) */
  uint64_t certified_queries;   /* first pass proven complete                  */
  uint64_t retried_queries;     /* re-run in range mode, exact afterwards      */
  uint64_t uncertified_queries; /* more near-ties than the range pass holds
                                   (4096 rows): best-effort result            */
  uint64_t range_rows;          /* rows the range passes re-ranked             */
  uint32_t n_devices;           /* 1, or the shard count of a group            */
  uint32_t reserved2;
  /* tensor-core path (last_path == 2): algorithmic flops (2 * nq * rows * dims) of the last
   * search over its device time, and that as a fraction of B200's nominal dense peak for the
   * MMA kind used (kind::f16 2250 TFLOP/s, kind::tf32 1125); 0 on the scan path */
  double last_tflops;
  double last_tensor_util;
} tsc_stats;

/* ---- library ---- */
int32_t tsc_version(void);                       /* TSC_ABI_VERSION             */
int32_t tsc_device_count(void);                  /* >= 0, or negative status    */
const char *tsc_last_error(void);                /* thread-local static string  */
const char *tsc_status_name(int32_t status);

/* ---- index lifetime ---- */
int32_t tsc_index_create(const tsc_index_desc *desc, uint64_t *out_handle);
int32_t tsc_index_destroy(uint64_t handle);
int32_t tsc_index_clear(uint64_t handle);        /* drop all rows, keep capacity;
                                                    clearCacheForIndex hook,
                                                    vector_index_manager.dart:1192-1202 */

/* ---- corpus ingestion (flush-time hooks, vector_index_manager.dart:378-387) ---- */
/* rows: dense row-major [n_rows, dims] of src_precision, HOST memory. node ids
 * [first_node_id, first_node_id+n_rows) must lie inside the shard and be
 * appended densely (first_node_id <= shard_first + rows). */
int32_t tsc_index_append_rows(uint64_t handle, uint64_t first_node_id,
                              const void *rows, uint64_t n_rows);
/* pages: n_pages consecutive reference raw-vector pages (page_size bytes each,
 * as stored in rawvec/dir_k/p<n>.ngh after the per-file meta page), HOST
 * memory. first_logical_page = nodeId / vectorsPerRawPage of the first page.
 * The library validates magic / header / CRC-32 / page type / dims, strips the
 * 28-byte headers and decodes f64 / f32 / i8 elements exactly like
 * NghRawVectorPage.getVectorAsFloat32. live_rows = nextNodeId of the index:
 * slots at or beyond it (zero-filled tail of the last page) are ignored. */
int32_t tsc_index_append_pages(uint64_t handle, uint64_t first_logical_page,
                               const uint8_t *pages, uint64_t n_pages,
                               uint32_t page_size, uint64_t live_rows);
/* test / benchmark helper: append n_rows of the deterministic synthetic corpus
 * (seed, global row index) generated on the device; see DESIGN.md. */
int32_t tsc_index_append_synthetic(uint64_t handle, uint64_t seed,
                                   uint64_t first_node_id, uint64_t n_rows);

/* ---- cold start from an existing on-disk index ----
 * index_dir is the directory that holds `ngh/meta.json`, `ngh/rawvec/dir_k/p<n>.ngh` and
 * `ngh/graph/dir_k/p<n>.ngh` (core/path_manager.dart:317-324). The loader reads only the
 * partition files that hold this shard's node ids [first_node_id, first_node_id +
 * capacity_rows) ∩ [0, nextNodeId), on a reader thread that runs one 64 MiB chunk ahead of
 * the GPU (pinned double buffer), and feeds them to tsc_index_append_pages /
 * tsc_index_apply_graph_pages. dims and precision of the index on disk must equal the
 * GPU index's dims / src_precision. */
enum { TSC_LOAD_TOMBSTONES = 1 };     /* also read the graph pages' deleted flags */
typedef struct tsc_ngh_info {
  uint32_t struct_size;
  uint32_t dims;                     /* meta.json "dimensions"                     */
  uint8_t metric;                    /* TSC_METRIC_* ("distanceMetric", default cosine) */
  uint8_t precision;                 /* TSC_SRC_*    ("precision", default float32)      */
  uint16_t reserved;
  uint32_t page_size;                /* "nghPageSize", default 16384               */
  uint32_t max_degree;               /* "maxDegree", default 64                    */
  uint32_t reserved2;
  uint64_t next_node_id;             /* "nextNodeId": rows are node ids [0, next)  */
  uint64_t max_partition_file_size;  /* "maxPartitionFileSize", default 16 MiB     */
  uint64_t files_read, pages_read, bytes_read;   /* filled by tsc_index_load_ngh   */
  double seconds;
} tsc_ngh_info;
int32_t tsc_ngh_read_meta(const char *index_dir, tsc_ngh_info *out);   /* host only */
int32_t tsc_index_load_ngh(uint64_t handle, const char *index_dir, uint32_t flags,
                           tsc_ngh_info *out /* optional */);

/* ---- liveness ---- */
int32_t tsc_index_set_deleted(uint64_t handle, const uint64_t *node_ids,
                              uint64_t n, uint8_t deleted);
/* graph pages (graph/dir_k/p<n>.ngh data pages): only the per-slot flags byte is
 * read; bit 0x01 marks a tombstone. first_logical_page = nodeId / nodesPerGraphPage. */
int32_t tsc_index_apply_graph_pages(uint64_t handle, uint64_t first_logical_page,
                                    const uint8_t *pages, uint64_t n_pages,
                                    uint32_t page_size);
/* WHERE prefilter: bit (nodeId - first_node_id) of word (..)/64, LSB first,
 * 1 = row may be returned. NULL clears the filter. */
int32_t tsc_index_set_filter(uint64_t handle, const uint64_t *bitmap_words,
                             uint64_t n_words);

/* ---- structured WHERE prefilter on the GPU (BASELINE config 5; additive: the
 * reference has no WHERE for vectors). Numeric table fields are kept column-wise in
 * HBM, aligned by node id; a condition tree in postfix order is evaluated in one pass
 * into the filter bitmap. Operator semantics restate ConditionRecordMatcher
 * (handler/value_matcher.dart:570-612; numeric order = Dart num.compareTo :150-174:
 * -0.0 < 0.0, NaN above +inf and equal to itself; NULL != x is true, NULL NOT IN is
 * true, every ordering operator / IN / BETWEEN is false on NULL; a childless AND or OR
 * is true :476-493). ---- */
enum { TSC_COL_I64 = 0, TSC_COL_F64 = 1, TSC_COL_TEXT = 2 };
enum { TSC_W_LEAF = 0, TSC_W_AND = 1, TSC_W_OR = 2 };
enum {
  TSC_OP_EQ = 0, TSC_OP_NE = 1, TSC_OP_GT = 2, TSC_OP_GE = 3, TSC_OP_LT = 4, TSC_OP_LE = 5,
  TSC_OP_BETWEEN = 6, TSC_OP_IN = 7, TSC_OP_NOT_IN = 8, TSC_OP_IS_NULL = 9,
  TSC_OP_IS_NOT_NULL = 10, TSC_OP_TRUE = 11, TSC_OP_FALSE = 12,
  TSC_OP_LIKE = 13, TSC_OP_NOT_LIKE = 14   /* text columns only */
};
typedef struct tsc_where_op {
  uint8_t kind;          /* TSC_W_*                                               */
  uint8_t op;            /* TSC_OP_* (leaves)                                     */
  uint16_t n;            /* AND / OR: children popped; IN / NOT IN: list length   */
  uint32_t column_id;    /* leaves: the column the operator reads                 */
  int64_t i_lo, i_hi;    /* operand(s) when the column is TSC_COL_I64; TSC_COL_TEXT:
                          * index of the operand string in the program's text pool   */
  double f_lo, f_hi;     /* operand(s) when the column is TSC_COL_F64             */
  uint32_t args_offset;  /* IN / NOT IN: first element in `in_args`               */
  uint32_t reserved;
} tsc_where_op;
int32_t tsc_index_column_create(uint64_t handle, uint32_t column_id, uint8_t col_type);
/* values: n x 8 bytes (int64 or double per the column type), HOST. is_null: n bytes
 * (non-zero = NULL) or NULL pointer for "no NULLs". Rows are node ids
 * [first_node_id, first_node_id + n), appended densely like the embedding rows;
 * overwriting already-appended rows is allowed (updates). */
int32_t tsc_index_column_append(uint64_t handle, uint32_t column_id, uint64_t first_node_id,
                                const void *values, const uint8_t *is_null, uint64_t n);
/* Evaluate the postfix program (<= 64 steps, <= 4096 IN-list values, in_args are 8-byte
 * values typed like the column of the leaf that uses them) over rows [0, rows) of every
 * column it names and install the result as the index's filter (as tsc_index_set_filter
 * does). An empty program matches every row. out_matched (optional): rows that passed.
 * Columns shorter than the embedding column read as NULL beyond their end. */
int32_t tsc_index_filter_where(uint64_t handle, const tsc_where_op *ops, uint32_t n_ops,
                               const void *in_args, uint32_t n_in_args, uint64_t *out_matched);

/* ---- text fields (DataType.text) in the WHERE prefilter. A TSC_COL_TEXT column is
 * dictionary-encoded in HBM: distinct strings once, as UTF-16 code units (a Dart String's
 * own code units, `String.codeUnits`), rows hold the code. Order is Dart's
 * String.compareTo — lexicographic over code units (text matcher,
 * handler/value_matcher.dart:211-240); LIKE / NOT LIKE follow ValueMatcher.matchesLike
 * (:318-331, :599-604): `%` = any run of code units, `_` = exactly one, neither matches a
 * line terminator (\n \r U+2028 U+2029: the pattern becomes a RegExp without dotAll),
 * everything else is literal (no escape character), whole-string, case-sensitive, false on
 * NULL for both LIKE and NOT LIKE. The other operators behave as on numeric columns.
 * Every distinct string is tested once per text leaf on the GPU (dict_match_kernel), the
 * row pass looks the row's code up in the resulting bitmap. ---- */
/* strings: row i = units [offsets[i], offsets[i + 1]) — n + 1 offsets, in code units,
 * absolute into `units`. is_null as for tsc_index_column_append (a NULL row's range is
 * ignored). HOST buffers. Values are stored as given (the host applies the field's
 * convertValue, i.e. trim(), model/table_schema.dart:1421-1442, before calling). */
int32_t tsc_index_column_append_text(uint64_t handle, uint32_t column_id,
                                     uint64_t first_node_id, const uint16_t *units,
                                     const uint64_t *offsets, const uint8_t *is_null,
                                     uint64_t n);
/* tsc_index_filter_where for programs with leaves on text columns. Text operand t of the
 * program is text_units [text_offsets[t], text_offsets[t + 1]) (n_texts + 1 offsets,
 * <= 4096 operands). A leaf on a text column names its operand(s) by index: i_lo (and
 * i_hi for BETWEEN's end); the in_args entries of its IN / NOT IN list are operand
 * indices (uint64). LIKE / NOT LIKE on a numeric column are rejected
 * (TSC_ERR_UNSUPPORTED). */
int32_t tsc_index_filter_where_text(uint64_t handle, const tsc_where_op *ops, uint32_t n_ops,
                                    const void *in_args, uint32_t n_in_args,
                                    const uint16_t *text_units, const uint64_t *text_offsets,
                                    uint32_t n_texts, uint64_t *out_matched);

/* ---- search ---- */
/* queries: [nq, dims] fp32, already padded/truncated to dims and, for cosine,
 * already normalised (exactly what VectorIndexManager hands NghGraphEngine.search).
 * distance_threshold: NaN = none; results with distance > threshold are dropped.
 * out_ids [nq*k] (-1 padded), out_dist [nq*k] ascending (NaN padded),
 * out_counts [nq]. HOST buffers; blocks until the results are written. */
/* On a group handle, or on a shard whose exchange has been set up (tsc_comm_*), this
 * is the SHARDED search: out_* receive the global result (on the consumer ranks). */
int32_t tsc_search(uint64_t handle, const float *queries, uint32_t nq, uint32_t k,
                   double distance_threshold, int64_t *out_ids, double *out_dist,
                   uint32_t *out_counts);
/* non-blocking pair so a Dart isolate can yield between polls */
int32_t tsc_search_submit(uint64_t handle, const float *queries, uint32_t nq,
                          uint32_t k, double distance_threshold, int64_t *out_ids,
                          double *out_dist, uint32_t *out_counts,
                          uint64_t *out_ticket);
int32_t tsc_search_poll(uint64_t ticket, int32_t *out_done);   /* 0 / 1          */
int32_t tsc_search_wait(uint64_t ticket);                      /* blocks, frees  */
/* DEVICE buffers on the index's device, asynchronous on `cuda_stream`
 * (a cudaStream_t passed as void*, NULL = the index's own stream). */
int32_t tsc_search_device(uint64_t handle, const float *d_queries, uint32_t nq,
                          uint32_t k, double distance_threshold, int64_t *d_out_ids,
                          double *d_out_dist, uint32_t *d_out_counts,
                          void *cuda_stream);
/* Throughput mode for back-to-back DEVICE-buffer searches of up to 8 queries on ONE stream.
 * on != 0: the scan launch of search i+1 is made a programmatic dependent of search i's
 * kernels, so its HBM pass overlaps search i's tail (selection, exact re-rank, certificate,
 * shard exchange) and the launch gaps; it waits for search i to complete before it publishes
 * anything, and results still become visible in stream order. Contract while it is on:
 * (1) the query buffers of consecutive searches must not be produced by KERNELS enqueued
 * between those searches on the same stream (memcpys and anything enqueued before the
 * previous search are fine) — the overlapping launch reads its queries before the previous
 * kernels are formally complete; (2) the stream of the previous search must still exist
 * when a search is issued on another stream; (3) NO range pass runs in-stream: a query whose
 * exactness certificate fails keeps its best-effort result, tsc_search_flags reports 1 for it
 * and tsc_stats counts it under uncertified_queries — the caller re-issues such queries with
 * pipelining off (a launch between two scans would keep them from overlapping). Only every
 * 16th search carries the CUDA events of the hot-kernel timer (events would serialise the
 * launches). Host-buffer searches are unaffected. Off by default. */
int32_t tsc_index_set_pipelining(uint64_t handle, int32_t on);
/* Host mirror of VectorIndexManager.vectorSearch's arithmetic around the engine
 * call (core/vector_index_manager.dart:514-520, :576-587): _toFloat32
 * (truncate / zero-pad `values[len]` to dims), _normalizeFloat32 for cosine,
 * search, _distanceToScore. out_* are [k]. */
int32_t tsc_vector_search(uint64_t handle, const double *values, uint64_t len,
                          uint32_t k, double distance_threshold, int64_t *out_ids,
                          double *out_dist, double *out_score, uint32_t *out_count);

/* Batch form (additive; the reference's API is single-query): nq query vectors of `len`
 * values each [nq][len], prepared like single queries and searched in one call (the
 * tensor-core path from 5 queries on 16-bit columns, from 9 on fp32 columns). out_ids / out_dist / out_score are
 * [nq][k], out_counts [nq]. */
int32_t tsc_vector_search_batch(uint64_t handle, const double *values, uint64_t len, uint32_t nq,
                                uint32_t k, double distance_threshold, int64_t *out_ids,
                                double *out_dist, double *out_score, uint32_t *out_counts);

/* ---- nodeId -> primary key side table. Takes the place of the
 * `<index>__nid2pk` B+Tree lookups after the engine call
 * (core/vector_index_manager.dart:553-588; maintained at flush :1276-1293): a dense
 * host-memory table so that result assembly costs k array reads instead of k B+Tree
 * descents. key i = utf8[offsets[i], offsets[i+1]); an EMPTY key is the tombstone
 * mapping (value [1], :1283-1286): such rows are dropped from results. ---- */
int32_t tsc_index_set_primary_keys(uint64_t handle, uint64_t first_node_id,
                                   const uint8_t *utf8, const uint64_t *offsets, uint64_t n);
/* out_len = 0 when the node has no (or a tombstoned) mapping; TSC_ERR_BAD_ARG when the
 * key does not fit `capacity` (out_len still reports the needed size). */
int32_t tsc_index_get_primary_key(uint64_t handle, uint64_t node_id, uint8_t *out_utf8,
                                  uint32_t capacity, uint32_t *out_len);
/* WHERE prefilter from a SET OF PRIMARY KEYS — what any ToStore query returns, so every
 * condition the reference's own executor can evaluate becomes a vector prefilter: the rows
 * whose key (tsc_index_set_primary_keys) is in the set stay searchable, as after
 * tsc_index_set_filter with the corresponding bitmap. Replaces the `<index>__pk2nid` lookups
 * (core/vector_index_manager.dart:1350-1363) with a hash lookup in the library. Unknown and
 * empty keys are ignored; a key mapped by several node ids selects the highest. keys: n
 * utf-8 strings, key i = utf8 [offsets[i], offsets[i + 1]). out_matched (optional): rows
 * selected. HOST buffers. */
int32_t tsc_index_filter_primary_keys(uint64_t handle, const uint8_t *utf8,
                                      const uint64_t *offsets, uint64_t n,
                                      uint64_t *out_matched);
/* tsc_vector_search + the reference's result assembly (:576-587): results whose node
 * has no mapping are dropped, the rest stay in ascending distance order. out_pk_utf8
 * receives the concatenated keys, out_pk_offsets [k+1] their boundaries. */
int32_t tsc_vector_search_pk(uint64_t handle, const double *values, uint64_t len, uint32_t k,
                             double distance_threshold, int64_t *out_ids, double *out_dist,
                             double *out_score, uint8_t *out_pk_utf8, uint64_t pk_capacity,
                             uint64_t *out_pk_offsets, uint32_t *out_count);

/* Per-query verdict of the exactness certificate for the LAST search on the handle:
 * 0 = exact (certified, possibly after a range pass), 1 = a range pass is still owed
 * (device-buffer searches with more uncertified queries than the in-stream range
 * launches cover; the host-buffer calls never return this), 2 = uncertified (more
 * than 4096 rows tie with the k-th neighbour within the key error bound). */
int32_t tsc_search_flags(uint64_t handle, uint32_t nq, uint32_t *out_flags);

/* ---- row-range sharding across GPUs (one process per GPU) ---- */
/* Merge n_parts per-shard results ([n_parts, nq, k] ids/dist as produced by
 * tsc_search_device on each shard and all-gathered by the caller or by
 * tsc_comm_*) into the global top-k. DEVICE buffers, async on cuda_stream. */
int32_t tsc_merge_shards(uint64_t handle, const int64_t *d_part_ids,
                         const double *d_part_dist, uint32_t n_parts, uint32_t nq,
                         uint32_t k, int64_t *d_out_ids, double *d_out_dist,
                         uint32_t *d_out_counts, void *cuda_stream);
/* NCCL communicator owned by the library (libnccl.so.2 is dlopen'ed on first
 * use). unique_id: 128 bytes, created on rank 0 and distributed by the caller. */
int32_t tsc_comm_unique_id(uint8_t *out_id128);
int32_t tsc_comm_init(uint64_t handle, const uint8_t *id128, int32_t n_ranks,
                      int32_t rank);
/* search this shard, ncclAllGather the per-shard top-k, merge on every rank.
 * DEVICE buffers; out_* hold the global result on every rank. */
int32_t tsc_search_sharded(uint64_t handle, const float *d_queries, uint32_t nq,
                           uint32_t k, double distance_threshold,
                           int64_t *d_out_ids, double *d_out_dist,
                           uint32_t *d_out_counts, void *cuda_stream);

/* The exchange over NVLink peer memory (preferred; tsc_comm_init remains as the NCCL form):
 * every rank pushes its k pairs straight into the consumers' receive buffers and publishes
 * release flags; for nq <= 8 this happens in the scan kernel's last CTA, so scan + select
 * + re-rank + exchange + merge is ONE kernel launch. One process per GPU. export returns
 * this rank's 64-byte CUDA IPC handle; the caller all-gathers the n_ranks handles (any
 * host transport) and hands the [n_ranks][64] array to import; afterwards
 * tsc_search_sharded / tsc_search take this path. root = -1: every rank receives the
 * global result (all-gather semantics); root = r: only rank r does — the others never
 * wait and out_* on them is left untouched. */
int32_t tsc_comm_p2p_export(uint64_t handle, int32_t n_ranks, int32_t rank, uint8_t *out_ipc64);
int32_t tsc_comm_p2p_import(uint64_t handle, const uint8_t *all_ipc, int32_t root);

/* ---- observability ---- */
int32_t tsc_stats_get(uint64_t handle, tsc_stats *out);
int32_t tsc_stats_reset(uint64_t handle);        /* zero the hot_* accumulators  */
/* raw device pointers for zero-copy interop (bench / torch.distributed glue) */
int32_t tsc_index_device_rows(uint64_t handle, void **out_ptr, uint64_t *out_rows,
                              uint64_t *out_row_stride_bytes);

/* test hook: fp32 ranking keys the tcgen05 path computes for every (query,row);
 * out_keys [nq, rows] HOST. Only for 16-bit device dtypes. */
int32_t tsc_debug_gemm_keys(uint64_t handle, const float *queries, uint32_t nq,
                            float *out_keys);
/* self-test hook: the warp-sliced CRC-32 of the page validator, re-enacted on the
 * host (no GPU needed); must equal CRC-32/IEEE (Crc32.of, btree_page.dart:64-89). */
uint32_t tsc_selftest_crc32(const uint8_t *data, uint32_t len);

/* self-test hook — NOT a fallback: no product entry point calls it, and
 * tsc_index_filter_where always runs where_eval_kernel on the GPU (TSC_ERR_CUDA without
 * one). It re-runs the program translation and the per-row evaluator the kernel shares
 * (a __host__ __device__ function) on host arrays so the CPU test tier can pin them. col_values is
 * [n_cols][n_rows] raw 8-byte values, col_is_null [n_cols][n_rows] bytes, out_match
 * [n_rows] bytes. */
int32_t tsc_selftest_where(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                           uint32_t n_in_args, uint32_t n_cols, const uint32_t *col_ids,
                           const uint8_t *col_types, const uint64_t *col_values,
                           const uint8_t *col_is_null, uint64_t n_rows, uint8_t *out_match);
/* the same with text columns: row r of text column c is row_units
 * [row_offsets[c * (n_rows + 1) + r], row_offsets[c * (n_rows + 1) + r + 1]) (entries of
 * numeric columns are ignored, as are col_values of text columns). Runs the library's own
 * dictionary builder, the per-string test dict_match_kernel shares (text_leaf_match) and
 * the row evaluator — on the host, for the CPU test tier. */
int32_t tsc_selftest_where_text(const tsc_where_op *ops, uint32_t n_ops, const void *in_args,
                                uint32_t n_in_args, const uint16_t *text_units,
                                const uint64_t *text_offsets, uint32_t n_texts,
                                uint32_t n_cols, const uint32_t *col_ids,
                                const uint8_t *col_types, const uint64_t *col_values,
                                const uint8_t *col_is_null, const uint16_t *row_units,
                                const uint64_t *row_offsets, uint64_t n_rows,
                                uint8_t *out_match);

/* self-test hooks (no GPU, not fallbacks): the host-side query preparation (_toFloat32 +
 * _normalizeFloat32) and score mapping (_distanceToScore) of tsc_vector_search. */
int32_t tsc_selftest_query_prep(uint32_t dims, int32_t metric, const double *values, uint64_t len,
                                float *out_f32);
double tsc_selftest_distance_to_score(int32_t metric, double distance);
/* self-test hook (no GPU): the directory walk / chunking / double-buffered reader of
 * tsc_index_load_ngh with a host sink that records, per chunk, the first logical page, the
 * page count and the CRC-32 of the chunk's bytes. category 0 = rawvec, 1 = graph. */
int32_t tsc_selftest_ngh_walk(const char *index_dir, uint32_t category, uint64_t node_lo,
                              uint64_t node_hi, uint32_t chunk_pages, uint64_t *out_first_page,
                              uint64_t *out_n_pages, uint32_t *out_crc, uint32_t max_chunks,
                              uint32_t *out_n_chunks);
/* self-test hooks (no GPU): a host-only index object that carries only the primary-key
 * table (every compute entry point fails on it; release with tsc_index_destroy), and the
 * result-assembly step of tsc_vector_search_pk applied to caller-supplied hits
 * (ids / dist / score [k], *inout_count valid entries; compacted in place). */
int32_t tsc_selftest_host_index(uint64_t capacity_rows, uint64_t first_node_id,
                                uint64_t *out_handle);
int32_t tsc_selftest_pk_assemble(uint64_t handle, uint32_t k, int64_t *ids, double *dist,
                                 double *score, uint8_t *out_pk_utf8, uint64_t pk_capacity,
                                 uint64_t *out_pk_offsets, uint32_t *inout_count);
/* self-test hook (no GPU): the bitmap tsc_index_filter_primary_keys would install (64-bit
 * words, LSB first), computed over the handle's primary-key table; works on a host-only
 * index. */
int32_t tsc_selftest_pk_filter_bitmap(uint64_t handle, const uint8_t *utf8,
                                      const uint64_t *offsets, uint64_t n, uint64_t *out_words,
                                      uint64_t n_words, uint64_t *out_matched);

#ifdef __cplusplus
}
#endif
#endif /* TOSTORE_CUDA_H */
