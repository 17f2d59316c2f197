/*
 * tostore_oracle.c — CPU ORACLE (test infrastructure, never shipped, never on
 * the product path). See tostore_oracle.h for the contract and for the
 * "parity unpinned by the reference" statement. Citations: /root/reference/lib/src.
 */
#include "tostore_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* query preparation — core/vector_index_manager.dart                  */
/* ------------------------------------------------------------------ */

/* :1385-1392 — Float32List(dimensions) is zero-filled; min(len,dims) copied;
 * a Float32List store of a double rounds to nearest-even. */
void tso_to_float32(const double *values, uint64_t len, uint32_t dims, float *out) {
  uint64_t n = len < dims ? len : dims;
  for (uint32_t i = 0; i < dims; i++) out[i] = 0.0f;
  for (uint64_t i = 0; i < n; i++) out[i] = (float)values[i];
}

/* :1395-1408 — mag accumulated in double over the fp32 values. */
int tso_normalize_f32(const float *v, uint32_t dims, float *out) {
  double mag = 0;
  for (uint32_t i = 0; i < dims; i++) mag += (double)v[i] * (double)v[i];
  mag = sqrt(mag);
  if (mag == 0) {
    memcpy(out, v, (size_t)dims * sizeof(float));
    return 0;
  }
  double inv = 1.0 / mag;
  for (uint32_t i = 0; i < dims; i++) out[i] = (float)((double)v[i] * inv);
  return 1;
}

/* :1411-1423 */
double tso_distance_to_score(double distance, int metric) {
  switch (metric) {
    case TSO_L2:
      return 1.0 / (1.0 + distance);
    case TSO_INNER_PRODUCT:
      return 1.0 / (1.0 + exp(-(-distance)));
    default: { /* (1.0 - distance).clamp(0.0, 1.0) */
      double s = 1.0 - distance;
      if (s != s) return s;
      if (s < 0.0) return 0.0;
      if (s > 1.0) return 1.0;
      return s;
    }
  }
}

/* ------------------------------------------------------------------ */
/* exact distances — core/ngh_graph_engine.dart:908-946                */
/* Float32List reads widen to double; every op is a separate IEEE      */
/* double operation in index order (compile with -ffp-contract=off).   */
/* ------------------------------------------------------------------ */

double tso_l2_distance(const float *a, const float *b, uint32_t d) { /* :920-927 */
  double sum = 0;
  for (uint32_t i = 0; i < d; i++) {
    double diff = (double)a[i] - (double)b[i];
    sum += diff * diff;
  }
  return sqrt(sum);
}

double tso_inner_product(const float *a, const float *b, uint32_t d) { /* :929-935 */
  double sum = 0;
  for (uint32_t i = 0; i < d; i++) sum += (double)a[i] * (double)b[i];
  return sum;
}

double tso_cosine_similarity(const float *a, const float *b, uint32_t d) { /* :937-946 */
  double dot = 0, magA = 0, magB = 0;
  for (uint32_t i = 0; i < d; i++) {
    dot += (double)a[i] * (double)b[i];
    magA += (double)a[i] * (double)a[i];
    magB += (double)b[i] * (double)b[i];
  }
  double denom = sqrt(magA) * sqrt(magB);
  return denom > 0 ? dot / denom : 0;
}

double tso_exact_distance(const float *a, const float *b, uint32_t d, int metric) { /* :908-918 */
  switch (metric) {
    case TSO_L2:
      return tso_l2_distance(a, b, d);
    case TSO_INNER_PRODUCT:
      return -tso_inner_product(a, b, d);
    default:
      return 1.0 - tso_cosine_similarity(a, b, d);
  }
}

/* Dart double.compareTo (used by results.sort, ngh_graph_engine.dart:133 and
 * vector_index_manager.dart:587): numeric order, -0.0 before 0.0, NaN after
 * everything and equal to itself. Dart's sort is unstable, so ties are
 * undefined there; the oracle fixes them by ascending node id. */
int tso_compare(double da, int64_t ia, double db, int64_t ib) {
  if (da < db) return -1;
  if (da > db) return 1;
  if (da == db) {
    if (da == 0.0) {
      int sa = signbit(da) ? 1 : 0, sb = signbit(db) ? 1 : 0;
      if (sa != sb) return sa ? -1 : 1;
    }
  } else {
    int na = da != da, nb = db != db;
    if (na && !nb) return 1;
    if (!na && nb) return -1;
  }
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

/* ------------------------------------------------------------------ */
/* bounded max-heap top-k (role of _FixedHeap, ngh_graph_engine.dart   */
/* :1131-1227): root = current worst kept result.                      */
/* ------------------------------------------------------------------ */
typedef struct {
  double *d;
  int64_t *id;
  uint32_t n, cap;
} tso_heap;

static void heap_sift_down(tso_heap *h, uint32_t i) {
  for (;;) {
    uint32_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < h->n && tso_compare(h->d[l], h->id[l], h->d[m], h->id[m]) > 0) m = l;
    if (r < h->n && tso_compare(h->d[r], h->id[r], h->d[m], h->id[m]) > 0) m = r;
    if (m == i) return;
    double td = h->d[i]; h->d[i] = h->d[m]; h->d[m] = td;
    int64_t ti = h->id[i]; h->id[i] = h->id[m]; h->id[m] = ti;
    i = m;
  }
}

static void heap_offer(tso_heap *h, double d, int64_t id) {
  if (h->cap == 0) return;
  if (h->n < h->cap) {
    uint32_t i = h->n++;
    h->d[i] = d; h->id[i] = id;
    while (i > 0) {
      uint32_t p = (i - 1) / 2;
      if (tso_compare(h->d[i], h->id[i], h->d[p], h->id[p]) <= 0) break;
      double td = h->d[i]; h->d[i] = h->d[p]; h->d[p] = td;
      int64_t ti = h->id[i]; h->id[i] = h->id[p]; h->id[p] = ti;
      i = p;
    }
  } else if (tso_compare(d, id, h->d[0], h->id[0]) < 0) {
    h->d[0] = d; h->id[0] = id;
    heap_sift_down(h, 0);
  }
}

typedef struct { double d; int64_t id; } tso_pair;
static int pair_cmp(const void *a, const void *b) {
  const tso_pair *x = (const tso_pair *)a, *y = (const tso_pair *)b;
  return tso_compare(x->d, x->id, y->d, y->id);
}

static inline int bit_at(const uint64_t *bm, uint64_t i) {
  return (int)((bm[i >> 6] >> (i & 63)) & 1u);
}

/* row source: either an array or the synthetic generator */
typedef struct {
  const float *rows; uint64_t ld;          /* array mode */
  uint64_t seed; int dev_dtype; int synth; /* synth mode */
} tso_src;

static void fetch_row(const tso_src *s, uint64_t i, uint32_t dims, float *tmp,
                      const float **row) {
  if (!s->synth) { *row = s->rows + i * s->ld; return; }
  for (uint32_t j = 0; j < dims; j++) tmp[j] = tso_synth_value(s->seed, i * dims + j);
  if (s->dev_dtype != TSO_DEV_F32) tso_round_rows(tmp, dims, s->dev_dtype);
  *row = tmp;
}

static uint32_t search_impl(const tso_src *src, uint64_t n, uint32_t dims,
                            int64_t first_node_id, const uint64_t *deleted,
                            const uint64_t *filter, const float *query, int metric,
                            uint32_t k, double threshold, int threads,
                            int64_t *out_ids, double *out_dist) {
  if (k == 0 || n == 0) return 0;
  int has_thr = !(threshold != threshold);
  int nt = threads > 1 ? threads : 1;
#ifndef _OPENMP
  nt = 1;
#endif
  tso_pair *all = (tso_pair *)malloc(sizeof(tso_pair) * (size_t)k * (size_t)nt);
  uint32_t *counts = (uint32_t *)calloc((size_t)nt, sizeof(uint32_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
    int t = 0, tn = 1;
#ifdef _OPENMP
    t = omp_get_thread_num(); tn = omp_get_num_threads();
#endif
    uint64_t per = (n + (uint64_t)tn - 1) / (uint64_t)tn;
    uint64_t lo = per * (uint64_t)t, hi = lo + per; if (hi > n) hi = n;
    tso_heap h; h.n = 0; h.cap = k;
    h.d = (double *)malloc(sizeof(double) * k);
    h.id = (int64_t *)malloc(sizeof(int64_t) * k);
    float *tmp = src->synth ? (float *)malloc(sizeof(float) * dims) : NULL;
    for (uint64_t i = lo; i < hi; i++) {
      if (deleted && bit_at(deleted, i)) continue;   /* tombstone, ngh_page.dart:104-108 */
      if (filter && !bit_at(filter, i)) continue;
      const float *row;
      fetch_row(src, i, dims, tmp, &row);
      double dist = tso_exact_distance(query, row, dims, metric);
      if (has_thr && dist > threshold) continue;     /* ngh_graph_engine.dart:127 */
      heap_offer(&h, dist, first_node_id + (int64_t)i);
    }
    if (t < nt) {
      for (uint32_t j = 0; j < h.n; j++) {
        all[(size_t)t * k + j].d = h.d[j];
        all[(size_t)t * k + j].id = h.id[j];
      }
      counts[t] = h.n;
    }
    free(h.d); free(h.id); free(tmp);
  }
  /* compact + final sort, then cut at k (ngh_graph_engine.dart:133-134) */
  size_t total = 0;
  for (int t = 0; t < nt; t++) {
    if (total != (size_t)t * k) memmove(all + total, all + (size_t)t * k, sizeof(tso_pair) * counts[t]);
    total += counts[t];
  }
  qsort(all, total, sizeof(tso_pair), pair_cmp);
  uint32_t m = total < k ? (uint32_t)total : k;
  for (uint32_t j = 0; j < m; j++) { out_ids[j] = all[j].id; out_dist[j] = all[j].d; }
  free(all); free(counts);
  return m;
}

uint32_t tso_search(const float *rows, uint64_t n, uint32_t dims, uint64_t ld,
                    int64_t first_node_id, const uint64_t *deleted,
                    const uint64_t *filter, const float *query, int metric,
                    uint32_t k, double threshold, int threads,
                    int64_t *out_ids, double *out_dist) {
  tso_src s; memset(&s, 0, sizeof s);
  s.rows = rows; s.ld = ld;
  return search_impl(&s, n, dims, first_node_id, deleted, filter, query, metric, k,
                     threshold, threads, out_ids, out_dist);
}

uint32_t tso_search_synth(uint64_t seed, uint64_t n, uint32_t dims, int dev_dtype,
                          int64_t first_node_id, const uint64_t *deleted,
                          const uint64_t *filter, const float *query, int metric,
                          uint32_t k, double threshold, int threads,
                          int64_t *out_ids, double *out_dist) {
  tso_src s; memset(&s, 0, sizeof s);
  s.synth = 1; s.seed = seed; s.dev_dtype = dev_dtype;
  return search_impl(&s, n, dims, first_node_id, deleted, filter, query, metric, k,
                     threshold, threads, out_ids, out_dist);
}

/* ------------------------------------------------------------------ */
/* synthetic data (new; shared with csrc/synth.cuh bit for bit)        */
/* ------------------------------------------------------------------ */
float tso_synth_value(uint64_t seed, uint64_t flat_index) {
  uint64_t z = seed + (flat_index + 1) * 0x9E3779B97F4A7C15ull; /* splitmix64 */
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  int32_t s = (int32_t)(int16_t)(z & 0xFFFF) + (int32_t)(int16_t)((z >> 16) & 0xFFFF) +
              (int32_t)(int16_t)((z >> 32) & 0xFFFF) + (int32_t)(int16_t)((z >> 48) & 0xFFFF);
  return (float)s * 3.0517578125e-05f; /* 2^-15: exact */
}

void tso_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dims,
                    uint64_t ld, float *out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < (int64_t)n; r++) {
    float *o = out + (uint64_t)r * ld;
    uint64_t base = (first_row + (uint64_t)r) * dims;
    for (uint32_t j = 0; j < dims; j++) o[j] = tso_synth_value(seed, base + j);
    for (uint64_t j = dims; j < ld; j++) o[j] = 0.0f;
  }
}

float tso_round_bf16(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) { /* inf / nan: truncate, keep nan quiet */
    if (u & 0x007FFFFFu) u |= 0x00400000u;
    u &= 0xFFFF0000u;
  } else {
    u += 0x7FFFu + ((u >> 16) & 1u);
    u &= 0xFFFF0000u;
  }
  float r; memcpy(&r, &u, 4);
  return r;
}

float tso_round_f16(float x) { return (float)(_Float16)x; }

void tso_round_rows(float *rows, uint64_t count, int dev_dtype) {
  if (dev_dtype == TSO_DEV_BF16)
    for (uint64_t i = 0; i < count; i++) rows[i] = tso_round_bf16(rows[i]);
  else if (dev_dtype == TSO_DEV_F16)
    for (uint64_t i = 0; i < count; i++) rows[i] = tso_round_f16(rows[i]);
}

/* ------------------------------------------------------------------ */
/* page envelope — core/btree_page.dart                                */
/* ------------------------------------------------------------------ */
static uint32_t crc_table[256];
static int crc_ready = 0;
static void crc_init(void) { /* :67-78 */
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    crc_table[i] = c;
  }
  crc_ready = 1;
}

uint32_t tso_crc32(const uint8_t *data, size_t len) { /* :81-88 */
  if (!crc_ready) crc_init();
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < len; i++) c = crc_table[(c ^ data[i]) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

static void put_u16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put_u32(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
static uint32_t get_u16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t get_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

int tso_build_page(uint8_t page_type, const uint8_t *payload, uint32_t payload_len,
                   uint32_t page_size, uint8_t *out_page) { /* :148-160, :173-203 */
  if ((uint64_t)TSO_PAGE_HEADER + payload_len > page_size) return -1;
  memset(out_page, 0, page_size);
  put_u32(out_page + 0, TSO_PAGE_MAGIC);
  put_u16(out_page + 4, TSO_PAGE_HEADER);
  out_page[6] = page_type;
  out_page[7] = 0;
  put_u32(out_page + 8, payload_len);
  put_u32(out_page + 12, tso_crc32(payload, payload_len));
  put_u32(out_page + 16, 0);
  memcpy(out_page + TSO_PAGE_HEADER, payload, payload_len);
  return 0;
}

int64_t tso_parse_page(const uint8_t *page, uint32_t page_size, uint8_t *out_type,
                       const uint8_t **out_payload) { /* :163-181, :206-226 */
  if (page_size < TSO_PAGE_HEADER) return -1;
  if (get_u32(page) != TSO_PAGE_MAGIC) return -1;
  if (get_u16(page + 4) != TSO_PAGE_HEADER) return -1;
  uint8_t pt = page[6];
  if (pt >= 10) return -1; /* BTreePageType.values.length */
  uint32_t len = get_u32(page + 8);
  if ((uint64_t)TSO_PAGE_HEADER + len > page_size) return -2;
  if (tso_crc32(page + TSO_PAGE_HEADER, len) != get_u32(page + 12)) return -3;
  *out_type = pt;
  *out_payload = page + TSO_PAGE_HEADER;
  return (int64_t)len;
}

/* ------------------------------------------------------------------ */
/* NGH pages — core/ngh_page.dart                                      */
/* ------------------------------------------------------------------ */
uint32_t tso_bytes_per_element(int precision) { /* :331-340 */
  return precision == TSO_F64 ? 8u : (precision == TSO_I8 ? 1u : 4u);
}

uint32_t tso_vectors_per_raw_page(uint32_t page_size, uint32_t dims, uint32_t bpe) { /* :575-579 */
  int64_t usable = (int64_t)page_size - 20 - 8 - 64;
  int64_t vec = (int64_t)dims * bpe;
  return (usable > 0 && vec > 0) ? (uint32_t)(usable / vec) : 0u;
}

uint32_t tso_nodes_per_graph_page(uint32_t page_size, uint32_t max_degree) { /* :559-566 */
  int64_t slot = 2 + (int64_t)max_degree * 4;
  int64_t usable = (int64_t)page_size - 20 - 4 - 64;
  return usable > 0 ? (uint32_t)(usable / slot) : 0u;
}

void tso_encode_element(float v, int precision, uint8_t *out) { /* :394-412 */
  if (precision == TSO_F32) {
    memcpy(out, &v, 4);
  } else if (precision == TSO_F64) {
    double d = (double)v; memcpy(out, &d, 8);
  } else {
    double c = (double)v;
    if (c < -1.0) c = -1.0;
    if (c > 1.0) c = 1.0;
    long q = lround(c * 127.0); /* Dart .round(): half away from zero */
    out[0] = (uint8_t)(int8_t)q;
  }
}

float tso_decode_element(const uint8_t *in, int precision) { /* :368-389 */
  if (precision == TSO_F32) { float f; memcpy(&f, in, 4); return f; }
  if (precision == TSO_F64) { double d; memcpy(&d, in, 8); return (float)d; }
  return (float)((double)(int8_t)in[0] / 127.0);
}

int tso_build_rawvec_page(const float *rows, uint32_t n_rows, uint32_t dims,
                          int precision, uint32_t page_size, uint8_t *out_page) {
  uint32_t bpe = tso_bytes_per_element(precision);
  uint32_t cap = tso_vectors_per_raw_page(page_size, dims, bpe);
  if (cap == 0 || n_rows > cap || dims > 0xFFFFu) return -1;
  uint32_t payload_len = 8 + cap * dims * bpe;
  uint8_t *payload = (uint8_t *)calloc(payload_len, 1);
  put_u16(payload + 0, cap);          /* vectorCount = capacity (:346-362) */
  put_u16(payload + 2, dims);
  payload[4] = (uint8_t)precision;    /* [5..7] padding */
  for (uint32_t r = 0; r < n_rows; r++)
    for (uint32_t j = 0; j < dims; j++)
      tso_encode_element(rows[(size_t)r * dims + j], precision,
                         payload + 8 + ((size_t)r * dims + j) * bpe);
  int rc = tso_build_page(TSO_PT_NGH_RAWVEC, payload, payload_len, page_size, out_page);
  free(payload);
  return rc;
}

int32_t tso_parse_rawvec_page(const uint8_t *page, uint32_t page_size,
                              uint32_t expect_dims, float *out_rows,
                              uint32_t out_capacity) {
  uint8_t type; const uint8_t *p;
  int64_t len = tso_parse_page(page, page_size, &type, &p);
  if (len < 0) return (int32_t)len;
  if (type != TSO_PT_NGH_RAWVEC) return -4;
  if (len < 8) return -5;
  uint32_t vcount = get_u16(p), dims = get_u16(p + 2);
  int prec = p[4];
  if (dims == 0) return -5;
  uint32_t bpe = tso_bytes_per_element(prec);
  if ((uint64_t)len < 8 + (uint64_t)vcount * dims * bpe) return -5;
  if (dims != expect_dims) return -6;
  uint32_t m = vcount < out_capacity ? vcount : out_capacity;
  for (uint32_t r = 0; r < m; r++)
    for (uint32_t j = 0; j < dims; j++)
      out_rows[(size_t)r * dims + j] =
          tso_decode_element(p + 8 + ((size_t)r * dims + j) * bpe, prec);
  return (int32_t)vcount;
}

int tso_build_graph_page(const uint8_t *flags, uint32_t n_slots, uint32_t max_degree,
                         uint32_t page_size, uint8_t *out_page) {
  uint32_t slot = 2 + max_degree * 4;
  uint32_t cap = tso_nodes_per_graph_page(page_size, max_degree);
  if (cap == 0 || n_slots > cap) return -1;
  uint32_t payload_len = 4 + cap * slot;
  uint8_t *payload = (uint8_t *)calloc(payload_len, 1);
  put_u16(payload, cap);
  put_u16(payload + 2, max_degree);
  for (uint32_t i = 0; i < n_slots; i++) payload[4 + (size_t)i * slot] = flags[i];
  int rc = tso_build_page(TSO_PT_NGH_GRAPH, payload, payload_len, page_size, out_page);
  free(payload);
  return rc;
}

int32_t tso_parse_graph_page_flags(const uint8_t *page, uint32_t page_size,
                                   uint8_t *out_flags, uint32_t out_capacity) {
  uint8_t type; const uint8_t *p;
  int64_t len = tso_parse_page(page, page_size, &type, &p);
  if (len < 0) return (int32_t)len;
  if (type != TSO_PT_NGH_GRAPH) return -4;
  if (len < 4) return -5;
  uint32_t count = get_u16(p), deg = get_u16(p + 2);
  if (deg == 0) return -5;
  uint32_t slot = 2 + deg * 4;
  if ((uint64_t)len < 4 + (uint64_t)count * slot) return -5;
  uint32_t m = count < out_capacity ? count : out_capacity;
  for (uint32_t i = 0; i < m; i++) out_flags[i] = p[4 + (size_t)i * slot];
  return (int32_t)count;
}

/* model/ngh_index_meta.dart:451-490 (+ firstDataPageNo = 1, :232) */
void tso_node_location(uint64_t node_id, uint32_t per_page, uint32_t pages_per_partition,
                       uint64_t *partition, uint32_t *local_page, uint32_t *slot) {
  uint64_t logical = node_id / per_page;
  *partition = logical / pages_per_partition;
  *local_page = 1u + (uint32_t)(logical % pages_per_partition);
  *slot = (uint32_t)(node_id % per_page);
}

/* ---- text fields of the WHERE prefilter (handler/value_matcher.dart) ------------------ */
/* String.compareTo: lexicographic over UTF-16 code units (:211-240) */
int tso_string_compare(const uint16_t *a, uint32_t na, const uint16_t *b, uint32_t nb) {
  uint32_t n = na < nb ? na : nb;
  for (uint32_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na == nb ? 0 : (na < nb ? -1 : 1);
}

/* ValueMatcher.matchesLike (:318-331): ^pattern$ with % -> .* and _ -> . (a RegExp `.` is
 * one code unit other than \n \r U+2028 U+2029), every other character literal. Restated
 * as the textbook dynamic programme over (value prefix, pattern prefix) — deliberately a
 * different algorithm from both the library's two-pointer matcher and the regex restatement
 * in where_oracle.py, so the three pin each other. */
int tso_like_match(const uint16_t *s, uint32_t n, const uint16_t *pat, uint32_t m) {
  uint8_t *prev = (uint8_t *)calloc((size_t)m + 1, 1), *cur = (uint8_t *)calloc((size_t)m + 1, 1);
  if (!prev || !cur) { free(prev); free(cur); return -1; }
  /* row 0: the empty value is matched by a pattern prefix made of % only */
  prev[0] = 1;
  for (uint32_t j = 1; j <= m; j++) prev[j] = prev[j - 1] && pat[j - 1] == '%';
  for (uint32_t i = 1; i <= n; i++) {
    uint16_t c = s[i - 1];
    int lt = c == 0x000A || c == 0x000D || c == 0x2028 || c == 0x2029;
    cur[0] = 0;
    for (uint32_t j = 1; j <= m; j++) {
      uint16_t p = pat[j - 1];
      if (p == '%') cur[j] = cur[j - 1] || (!lt && prev[j]);       /* % takes nothing / takes c */
      else if (p == '_') cur[j] = !lt && prev[j - 1];
      else cur[j] = p == c && prev[j - 1];
    }
    uint8_t *t = prev; prev = cur; cur = t;
  }
  int r = prev[m];
  free(prev); free(cur);
  return r;
}

int tso_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
