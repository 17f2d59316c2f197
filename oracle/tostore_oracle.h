/*
 * tostore_oracle.h — CPU ORACLE for the ToStore vector-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This is a plain-C restatement of the reference's
 * (tocreator/tostore @ 130da06, pure Dart) exact-distance arithmetic, result
 * ordering, query preparation, score mapping and NGH page codec. Nothing in
 * the product path (tostore_b200/, libtostore_cuda.so) may include, link or
 * call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, as the checker or the reported CPU baseline.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference holds no test, golden vector
 * or fixture for vectorSearch, and no Dart SDK exists in
 * this image, so the oracle cannot be checked against reference output. It is
 * pinned instead by (1) an independent numpy-float64 restatement
 * (oracle/oracle_np.py) that must agree bit-for-bit, (2) closed-form values
 * for the reference's demo vectors (example/lib/tostore_example.dart:388-406),
 * (3) CRC-32/IEEE published check value 0xCBF43926.
 *
 * All file:line citations are relative to /root/reference/lib/src.
 * Build: gcc -O2 -ffp-contract=off -fopenmp (Dart never fuses a*b+c).
 */
#ifndef TOSTORE_ORACLE_H
#define TOSTORE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* VectorDistanceMetric enum order, model/table_schema.dart:2511-2531 */
enum { TSO_L2 = 0, TSO_INNER_PRODUCT = 1, TSO_COSINE = 2 };
/* VectorPrecision enum order, model/table_schema.dart:2481-2498 */
enum { TSO_F64 = 0, TSO_F32 = 1, TSO_I8 = 2 };
/* device storage dtypes of the new path (no reference equivalent) */
enum { TSO_DEV_F32 = 0, TSO_DEV_BF16 = 1, TSO_DEV_F16 = 2 };

/* ---- query preparation: core/vector_index_manager.dart ---- */
/* _toFloat32 :1385-1392 (== compute/vector_batch_prepare_compute.dart:79-86) */
void tso_to_float32(const double *values, uint64_t len, uint32_t dims, float *out);
/* _normalizeFloat32 :1395-1408; returns 0 when mag==0 (vector copied unchanged) */
int tso_normalize_f32(const float *v, uint32_t dims, float *out);
/* _distanceToScore :1411-1423 */
double tso_distance_to_score(double distance, int metric);

/* ---- exact distances: core/ngh_graph_engine.dart:908-946 ---- */
double tso_l2_distance(const float *a, const float *b, uint32_t d);
double tso_inner_product(const float *a, const float *b, uint32_t d);
double tso_cosine_similarity(const float *a, const float *b, uint32_t d);
double tso_exact_distance(const float *a, const float *b, uint32_t d, int metric);

/* double.compareTo total order (-0.0 < 0.0, NaN last), ties by node id. */
int tso_compare(double da, int64_t ia, double db, int64_t ib);

/*
 * Exhaustive search = the reference's re-rank semantics
 * (ngh_graph_engine.dart:122-134) applied to every live row:
 *   d = exact_distance(query,row); drop if d > threshold (threshold NaN = none);
 *   sort ascending by (d, nodeId); first k.
 * rows: [n, ld] fp32 row-major (ld >= dims; only the first dims are used).
 * deleted / filter: optional bitmaps, bit i of word i/64 (LSB first);
 *   a row is live iff !deleted[i] && (filter==NULL || filter[i]).
 * first_node_id is added to the row index to form the reported id.
 * threads <= 1: single thread (faithful to the reference's one isolate);
 * threads  > 1: OpenMP row-range split + merge.
 * Returns the number of results written (<= k).
 */
uint32_t tso_search(const float *rows, uint64_t n, uint32_t dims, uint64_t ld,
                    int64_t first_node_id, const uint64_t *deleted,
                    const uint64_t *filter, const float *query, int metric,
                    uint32_t k, double threshold, int threads,
                    int64_t *out_ids, double *out_dist);

/* Same scan over synthetic rows generated on the fly (no [n,d] array in RAM). */
uint32_t tso_search_synth(uint64_t seed, uint64_t n, uint32_t dims, int dev_dtype,
                          int64_t first_node_id, const uint64_t *deleted,
                          const uint64_t *filter, const float *query, int metric,
                          uint32_t k, double threshold, int threads,
                          int64_t *out_ids, double *out_dist);

/* ---- synthetic data shared bit-for-bit with the CUDA generator ---- */
/* element (row, col) -> fp32; sum of four int16 lanes of splitmix64 * 2^-15 */
float tso_synth_value(uint64_t seed, uint64_t flat_index);
void tso_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dims,
                    uint64_t ld, float *out);
/* round-to-nearest-even through a 16-bit storage type and back to fp32 */
float tso_round_bf16(float x);
float tso_round_f16(float x);
void tso_round_rows(float *rows, uint64_t count, int dev_dtype);

/* ---- page envelope: core/btree_page.dart ---- */
/* Crc32.of :64-89 (CRC-32/IEEE reflected 0xEDB88320, init/xorout 0xFFFFFFFF) */
uint32_t tso_crc32(const uint8_t *data, size_t len);
#define TSO_PAGE_MAGIC 0x32475054u /* 'TPG2', btree_page.dart:134 */
#define TSO_PAGE_HEADER 20u        /* btree_page.dart:133 */
#define TSO_PT_NGH_GRAPH 6u        /* BTreePageType index, btree_page.dart:14-55 */
#define TSO_PT_NGH_RAWVEC 8u
/* BTreePageIO.buildPageBytes :173-203; returns 0 or -1 on overflow */
int tso_build_page(uint8_t page_type, const uint8_t *payload, uint32_t payload_len,
                   uint32_t page_size, uint8_t *out_page);
/* BTreePageIO.parsePageBytes :206-226; returns payload_len or <0:
 * -1 bad magic/header, -2 bad length, -3 crc mismatch */
int64_t tso_parse_page(const uint8_t *page, uint32_t page_size, uint8_t *out_type,
                       const uint8_t **out_payload);

/* ---- NGH pages: core/ngh_page.dart ---- */
uint32_t tso_bytes_per_element(int precision);                 /* :331-340 */
/* NghPageSizer.vectorsPerRawPage :575-579 */
uint32_t tso_vectors_per_raw_page(uint32_t page_size, uint32_t dims, uint32_t bpe);
/* NghPageSizer.nodesPerGraphPage :559-566 */
uint32_t tso_nodes_per_graph_page(uint32_t page_size, uint32_t max_degree);
/* NghRawVectorPage.setVectorFromFloat32 :394-412 for one element */
void tso_encode_element(float v, int precision, uint8_t *out);
/* NghRawVectorPage.getVectorAsFloat32 :368-389 for one element */
float tso_decode_element(const uint8_t *in, int precision);
/* Full-capacity, zero-padded raw-vector page (NghRawVectorPage.empty :346-362 +
 * encodePayload :414-425 + page envelope). rows: [n_rows, dims] fp32,
 * n_rows <= capacity. Returns 0 / -1. */
int tso_build_rawvec_page(const float *rows, uint32_t n_rows, uint32_t dims,
                          int precision, uint32_t page_size, uint8_t *out_page);
/* tryDecodePayload :427-447 + getVectorAsFloat32; out: [vectorCount, dims].
 * Returns vectorCount or <0 (envelope error codes, -4 wrong type, -5 bad payload,
 * -6 dims mismatch). */
int32_t tso_parse_rawvec_page(const uint8_t *page, uint32_t page_size,
                              uint32_t expect_dims, float *out_rows,
                              uint32_t out_capacity);
/* NghGraphPage.encodePayload :166-187 restricted to flags (neighbours zero). */
int tso_build_graph_page(const uint8_t *flags, uint32_t n_slots, uint32_t max_degree,
                         uint32_t page_size, uint8_t *out_page);
/* tryDecodePayload :189-216, flags only. Returns slotCount or <0. */
int32_t tso_parse_graph_page_flags(const uint8_t *page, uint32_t page_size,
                                   uint8_t *out_flags, uint32_t out_capacity);

/* ---- addressing: model/ngh_index_meta.dart:451-490 ---- */
void tso_node_location(uint64_t node_id, uint32_t per_page, uint32_t pages_per_partition,
                       uint64_t *partition, uint32_t *local_page, uint32_t *slot);

/* text fields of the WHERE prefilter: String.compareTo (value_matcher.dart:211-240) and
 * ValueMatcher.matchesLike (:318-331) over UTF-16 code units; like: 1 / 0, -1 = out of memory */
int tso_string_compare(const uint16_t *a, uint32_t na, const uint16_t *b, uint32_t nb);
int tso_like_match(const uint16_t *s, uint32_t n, const uint16_t *pat, uint32_t m);
int tso_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
