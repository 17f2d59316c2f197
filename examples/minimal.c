/* minimal.c — the C ABI of libtostore_cuda.so end to end, as any FFI host would drive it
 * (this is what dart/tostore_cuda_bindings.dart does through dart:ffi).
 *
 *   gcc -std=c99 -I include examples/minimal.c -L tostore_b200 -ltostore_cuda \
 *       -Wl,-rpath,$PWD/tostore_b200 -lm -o /tmp/minimal && /tmp/minimal
 *
 * Needs a B200; without a CUDA device tsc_index_create fails with TSC_ERR_CUDA (there is
 * no CPU fallback) and the program says so and exits 1.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tostore_cuda.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int32_t rc_ = (call);                                                           \
    if (rc_ != TSC_OK) {                                                            \
      fprintf(stderr, "%s -> %s: %s\n", #call, tsc_status_name(rc_), tsc_last_error()); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int main(void) {
  enum { N = 4096, D = 128, K = 5 };
  /* rows: a deterministic ramp family, like the reference's demo vectors
   * (example/lib/tostore_example.dart:388-406) */
  float *rows = (float *)malloc(sizeof(float) * N * D);
  for (int r = 0; r < N; r++)
    for (int i = 0; i < D; i++) rows[r * D + i] = (float)(sin(0.37 * r + 0.11 * i) + 0.001 * r);

  tsc_index_desc desc;
  memset(&desc, 0, sizeof desc);
  desc.struct_size = sizeof desc;
  desc.dims = D;
  desc.metric = TSC_METRIC_COSINE;       /* VectorDistanceMetric.cosine */
  desc.src_precision = TSC_SRC_F32;      /* VectorPrecision.float32     */
  desc.dev_dtype = TSC_DEV_F32;
  desc.capacity_rows = N;
  desc.k_max = 16;
  desc.nq_max = 8;
  /* one process, all GPUs of the box: with >= 2 devices the handle is a GROUP (the column is
   * row-range sharded inside the library); every call below stays the same */
  {
    int ndev = tsc_device_count();
    if (ndev >= 2) {
      desc.n_devices = (uint32_t)(ndev < 8 ? ndev : 8);
      for (uint32_t i = 0; i < desc.n_devices; i++) desc.device_ids[i] = (int32_t)i;
    }
  }
  uint64_t h = 0;
  CHECK(tsc_index_create(&desc, &h));
  CHECK(tsc_index_append_rows(h, 0, rows, N));          /* flush-time hook */

  /* nodeId -> primary key side table (role of the __nid2pk B+Tree) */
  char *pk_bytes = (char *)malloc(16 * N);
  uint64_t *pk_off = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  uint64_t used = 0;
  for (int r = 0; r < N; r++) {
    pk_off[r] = used;
    used += (uint64_t)sprintf(pk_bytes + used, "doc-%d", r);
  }
  pk_off[N] = used;
  CHECK(tsc_index_set_primary_keys(h, 0, (const uint8_t *)pk_bytes, pk_off, N));

  /* a numeric table field mirrored column-wise, and WHERE year >= 2020 AND year != 2022 */
  int64_t *year = (int64_t *)malloc(sizeof(int64_t) * N);
  for (int r = 0; r < N; r++) year[r] = 2000 + r % 25;
  CHECK(tsc_index_column_create(h, 0, TSC_COL_I64));
  CHECK(tsc_index_column_append(h, 0, 0, year, NULL, N));
  /* a text field (UTF-16 code units, as a Dart String holds them): "en" / "de" / "en-GB" */
  static const uint16_t kLang[] = {'e', 'n', 'd', 'e', 'e', 'n', '-', 'G', 'B'};
  static const uint64_t kLangOff[4] = {0, 2, 4, 9};
  uint16_t *lang_units = (uint16_t *)malloc(sizeof(uint16_t) * 5 * N);
  uint64_t *lang_off = (uint64_t *)malloc(sizeof(uint64_t) * (N + 1));
  uint64_t lu = 0;
  for (int r = 0; r < N; r++) {
    lang_off[r] = lu;
    for (uint64_t i = kLangOff[r % 3]; i < kLangOff[r % 3 + 1]; i++) lang_units[lu++] = kLang[i];
  }
  lang_off[N] = lu;
  CHECK(tsc_index_column_create(h, 1, TSC_COL_TEXT));
  CHECK(tsc_index_column_append_text(h, 1, 0, lang_units, lang_off, NULL, N));
  /* WHERE year >= 2020 AND year != 2022 AND lang LIKE 'en%' (text operand 0 of the pool) */
  static const uint16_t kPattern[] = {'e', 'n', '%'};
  static const uint64_t kPatternOff[2] = {0, 3};
  tsc_where_op prog[4];
  memset(prog, 0, sizeof prog);
  prog[0].kind = TSC_W_LEAF; prog[0].op = TSC_OP_GE; prog[0].column_id = 0; prog[0].i_lo = 2020;
  prog[1].kind = TSC_W_LEAF; prog[1].op = TSC_OP_NE; prog[1].column_id = 0; prog[1].i_lo = 2022;
  prog[2].kind = TSC_W_LEAF; prog[2].op = TSC_OP_LIKE; prog[2].column_id = 1; prog[2].i_lo = 0;
  prog[3].kind = TSC_W_AND;  prog[3].n = 3;
  uint64_t matched = 0;
  CHECK(tsc_index_filter_where_text(h, prog, 4, NULL, 0, kPattern, kPatternOff, 1, &matched));
  printf("WHERE matched %llu of %d rows\n", (unsigned long long)matched, N);

  /* ToStore.vectorSearch: fp64 query of any length, topK, optional threshold */
  double query[D];
  for (int i = 0; i < D; i++) query[i] = sin(0.37 * 777 + 0.11 * i);
  int64_t ids[K];
  double dist[K], score[K];
  uint8_t pks[256];
  uint64_t offs[K + 1];
  uint32_t count = 0;
  CHECK(tsc_vector_search_pk(h, query, D, K, NAN, ids, dist, score, pks, sizeof pks, offs, &count));
  for (uint32_t j = 0; j < count; j++)
    printf("%u  %.*s  node=%lld year=%lld lang=%s distance=%.12g score=%.6f\n", j,
           (int)(offs[j + 1] - offs[j]), (const char *)pks + offs[j], (long long)ids[j],
           (long long)year[ids[j]], ids[j] % 3 == 0 ? "en" : (ids[j] % 3 == 1 ? "de" : "en-GB"), dist[j],
           score[j]);

  tsc_stats st;
  memset(&st, 0, sizeof st);
  st.struct_size = sizeof st;
  CHECK(tsc_stats_get(h, &st));
  printf("rows=%llu device_bytes=%llu kernel_launches=%llu last_search_ms=%.3f gpus=%u "
         "certified=%llu range_pass=%llu uncertified=%llu\n",
         (unsigned long long)st.rows, (unsigned long long)st.device_bytes,
         (unsigned long long)st.kernel_launches, st.last_search_ms, st.n_devices,
         (unsigned long long)st.certified_queries, (unsigned long long)st.retried_queries,
         (unsigned long long)st.uncertified_queries);
  CHECK(tsc_index_destroy(h));
  free(rows); free(pk_bytes); free(pk_off); free(year); free(lang_units); free(lang_off);
  return 0;
}
