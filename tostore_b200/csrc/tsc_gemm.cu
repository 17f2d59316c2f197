// tsc_gemm.cu — host launcher for K2 (tsc_gemm.cuh): TMA tensor maps, tile
// scheduling geometry, row norms (K4).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tsc_gemm.cuh"
#include "tsc_index.h"

namespace tsc {

// cuTensorMapEncodeTiled comes from the driver; resolve it at run time so the
// library links against nothing but the (static) runtime.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int32_t load_encode() {
  if (g_encode) return TSC_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TSC_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  return TSC_OK;
}

// 2-D map over a row-major [rows, ld] matrix, box = [box_rows, 128 bytes of K], 128B swizzle
// (64 elements of a 16-bit type, 32 of fp32). Encoded once per (pointer, extent, box) and
// kept in the index: the corpus map only changes when rows are appended.
static_assert(sizeof(CUtensorMap) == 128, "Index::TmapSlot holds a CUtensorMap");
static int32_t cached_map(Index::TmapSlot *slot, int dtype, const void *ptr, uint64_t rows,
                          uint32_t ld, uint32_t row_bytes, uint32_t box_rows, CUtensorMap *out) {
  if (slot->ptr == ptr && slot->rows == rows && slot->box_rows == box_rows) {
    memcpy(out, slot->bytes, sizeof(CUtensorMap));
    return TSC_OK;
  }
  cuuint64_t dims[2] = {ld, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(dtype == kF32 ? kGemmBK / 2 : kGemmBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, dtype == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                             : dtype == kF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                        2, const_cast<void *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu ld=%u", (int)r,
              (unsigned long long)rows, ld);
    return TSC_ERR_CUDA;
  }
  memcpy(slot->bytes, out, sizeof(CUtensorMap));
  slot->ptr = ptr;
  slot->rows = rows;
  slot->box_rows = box_rows;
  return TSC_OK;
}

bool gemm_supported(const Index *ix, uint32_t kprime) {
  // 16-bit columns: kind::f16 on a storage-type copy of the queries; fp32 columns: kind::tf32
  // on the fp32 queries as they are
  // K' beyond the kernel's list capacity: the lists are truncated to kGemmMaxKp entries and the
  // tail's certificate accounts for what they dropped (tsc_tail.cuh, trunc_len)
  return kprime <= kMaxRerank && ix->d_norm2 != nullptr && !ix->host_only;
}

int32_t gemm_update_norms(Index *ix, uint64_t first_row, uint64_t n, cudaStream_t st) {
  if (!ix->d_norm2 || n == 0) return TSC_OK;
  unsigned blocks = (unsigned)((n + 7) / 8 < (uint64_t)ix->sm_count * 16 ? (n + 7) / 8
                                                                       : ix->sm_count * 16);
  if (ix->desc.dev_dtype == TSC_DEV_F32)
    row_norms_kernel<kF32><<<blocks, 256, 0, st>>>(ix->d_rows, first_row, n, ix->ld,
                                                   ix->row_bytes, ix->d_norm2, ix->d_maxnorm);
  else if (ix->desc.dev_dtype == TSC_DEV_BF16)
    row_norms_kernel<kBF16><<<blocks, 256, 0, st>>>(ix->d_rows, first_row, n, ix->ld,
                                                    ix->row_bytes, ix->d_norm2, ix->d_maxnorm);
  else
    row_norms_kernel<kF16><<<blocks, 256, 0, st>>>(ix->d_rows, first_row, n, ix->ld,
                                                   ix->row_bytes, ix->d_norm2, ix->d_maxnorm);
  TSC_CUDA(cudaGetLastError());
  ix->launches++;
  // the certificate's |row| bound (every caller synchronises the stream before it returns)
  TSC_CUDA(cudaMemcpyAsync(&ix->max_norm2, ix->d_maxnorm, 4, cudaMemcpyDeviceToHost, st));
  return TSC_OK;
}

#ifdef TSC_DIAG
static int diag_env(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
#else
static int diag_env(const char *, int dflt) { return dflt; }   // shipping build: no switches
#endif

// SS kernel for one CTA (CG = 1) or a CTA pair (CG = 2) per tile
template <int CG, int KPR, int KIND>
static int32_t launch_ss(Index *ix, GemmParams p, uint32_t nq, uint32_t kprime, float *dbg_keys,
                         uint32_t *out_lists, cudaStream_t st, const void *a_rows) {
  const int dtype = ix->desc.dev_dtype;
  const uint32_t q_units = (p.q_tiles + CG - 1) / CG;
  uint32_t slices = (uint32_t)(ix->sm_count / CG) / q_units;
  if (slices == 0) {
    set_error("gemm: nq=%u needs more query tiles than SMs", nq);
    return TSC_ERR_BAD_ARG;
  }
  if (slices > p.n_tiles) slices = p.n_tiles;
  p.n_slices = slices;
  uint32_t stages = CG == 2 ? 5 : 4;
  if (diag_env("TSC_GEMM_STAGES", 0) >= 2) stages = (uint32_t)diag_env("TSC_GEMM_STAGES", 0);
  while (stages > 2 && gemm_smem_bytes<CG>(stages, kprime) > ix->smem_optin) stages--;
  p.stages = stages;
  const size_t smem = gemm_smem_bytes<CG>(stages, kprime);
  if (smem > ix->smem_optin) {
    set_error("gemm: shared memory %zu exceeds %zu", smem, ix->smem_optin);
    return TSC_ERR_UNSUPPORTED;
  }
  CUtensorMap map_q, map_b;
  // A operand: queries in the storage type (16-bit copy, or the caller's fp32 rows for tf32)
  int32_t rc = cached_map(&ix->tmap_q, dtype, a_rows, nq, ix->ld, ix->row_bytes, kGemmBM, &map_q);
  if (rc != TSC_OK) return rc;
  rc = cached_map(&ix->tmap_b[CG - 1], dtype, ix->d_rows, ix->rows, ix->ld, ix->row_bytes,
                  GemmGeom<CG>::kBRows, &map_b);
  if (rc != TSC_OK) return rc;

  static bool attr_done[64] = {false};   // per (CG, KPR, KIND) instantiation
  if (!attr_done[ix->device & 63]) {
    TSC_CUDA(cudaFuncSetAttribute(gemm_topk_kernel<false, CG, false, KPR, KIND>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_optin));
    TSC_CUDA(cudaFuncSetAttribute(gemm_topk_kernel<true, CG, false, KPR, KIND>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_optin));
#ifdef TSC_DIAG
    TSC_CUDA(cudaFuncSetAttribute(gemm_topk_kernel<false, CG, true, KPR, KIND>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ix->smem_optin));
#endif
    attr_done[ix->device & 63] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(q_units * CG * p.n_slices);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint32_t idesc = umma_idesc_f16(dtype, kGemmBM * CG, kGemmBN);
#ifdef TSC_DIAG
  // diagnostics build only (libtostore_cuda_diag.so, tools/): TSC_GEMM_EXP = experiment bit
  // mask (results invalid), TSC_GEMM_PROF=1 prints the per-role wait / work cycle averages
  p.exp_flags = (uint32_t)diag_env("TSC_GEMM_EXP", 0);
  const bool prof = diag_env("TSC_GEMM_PROF", 0) != 0 && ix->d_progress && !dbg_keys;
  const bool exp_kernel = (p.exp_flags != 0 || prof) && !dbg_keys;
  if (prof) {
    TSC_CUDA(cudaMemsetAsync(ix->d_progress, 0, kProfSlots * 8, st));
    p.prof = reinterpret_cast<unsigned long long *>(ix->d_progress);
  }
#endif
  int slot = 0;
  rc = hot_timer_begin(ix, st, &slot);
  if (rc != TSC_OK) return rc;
  if (dbg_keys)
    TSC_CUDA(cudaLaunchKernelEx(&cfg, gemm_topk_kernel<true, CG, false, KPR, KIND>, map_q, map_b, p,
                                idesc));
#ifdef TSC_DIAG
  else if (exp_kernel)
    TSC_CUDA(cudaLaunchKernelEx(&cfg, gemm_topk_kernel<false, CG, true, KPR, KIND>, map_q, map_b, p,
                                idesc));
#endif
  else
    TSC_CUDA(cudaLaunchKernelEx(&cfg, gemm_topk_kernel<false, CG, false, KPR, KIND>, map_q, map_b, p,
                                idesc));
  ix->launches++;
  *out_lists = p.n_slices * 2;
  // algorithmic work: 2 * nq * N * d flops; corpus bytes read once
  rc = hot_timer_end(ix, st, slot, (double)ix->rows * ix->desc.dims * ix->elem_bytes,
                     2.0 * nq * (double)ix->rows * ix->desc.dims);
#ifdef TSC_DIAG
  if (rc == TSC_OK && prof) {
    unsigned long long h[kProfSlots];
    TSC_CUDA(cudaMemcpyAsync(h, ix->d_progress, sizeof(h), cudaMemcpyDeviceToHost, st));
    TSC_CUDA(cudaStreamSynchronize(st));
    auto avg = [&](int slot_, int n_) { return h[n_] ? (double)h[slot_] / (double)h[n_] : 0.0; };
    fprintf(stderr,
            "gemm_prof cg=%d stages=%u exp=%u | prod: wait_empty %.0f of %.0f | mma: "
            "wait_full %.0f wait_tempty %.0f of %.0f | epi/warp: wait_tfull %.0f "
            "ldtm %.0f math %.0f of %.0f (cycles, mean per role instance; %u tiles x %u k-blocks "
            "per CTA)\n",
            CG, p.stages, p.exp_flags, avg(kProfProdWaitEmpty, kProfProdN),
            avg(kProfProdTotal, kProfProdN), avg(kProfMmaWaitFull, kProfMmaN),
            avg(kProfMmaWaitTempty, kProfMmaN),
            avg(kProfMmaTotal, kProfMmaN), avg(kProfEpiWaitTfull, kProfEpiN),
            avg(kProfEpiLdtm, kProfEpiN), avg(kProfEpiMath, kProfEpiN),
            avg(kProfEpiTotal, kProfEpiN), (p.n_tiles + p.n_slices - 1) / p.n_slices, p.k_blocks);
  }
#endif
  return rc;
}

template <int KPR, int KIND>
static int32_t launch_pair_or_single(Index *ix, const GemmParams &p, uint32_t nq, uint32_t kprime,
                                     float *dbg_keys, uint32_t *out_lists, cudaStream_t st,
                                     const void *a_rows, bool pair) {
  return pair ? launch_ss<2, KPR, KIND>(ix, p, nq, kprime, dbg_keys, out_lists, st, a_rows)
              : launch_ss<1, KPR, KIND>(ix, p, nq, kprime, dbg_keys, out_lists, st, a_rows);
}

// d_q: fp32 [nq, qld]. d_cand receives [nq][n_slices * 2][kprime]; *out_lists = n_slices * 2.
int32_t launch_gemm(Index *ix, const float *d_q, uint32_t nq, uint32_t kprime, uint64_t *d_cand,
                    uint32_t *out_lists, float *dbg_keys, cudaStream_t st) {
  int32_t rc = load_encode();
  if (rc != TSC_OK) return rc;
  const int dtype = ix->desc.dev_dtype;
  const bool tf32 = dtype == TSC_DEV_F32;
  if (!tf32) {
    convert_queries_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, ix->qld, ix->d_q16,
                                                                  ix->d_enorm, dtype);
    TSC_CUDA(cudaGetLastError());
    ix->launches++;
  }

  GemmParams p{};
  p.n_rows = ix->rows;
  p.nq = nq;
  const uint32_t bk = tf32 ? kGemmBK / 2 : kGemmBK;   // elements per 128-byte K block
  p.k_blocks = (ix->desc.dims + bk - 1) / bk;
  p.q_tiles = (nq + kGemmBM - 1) / kGemmBM;
  p.n_tiles = (uint32_t)((ix->rows + kGemmBN - 1) / kGemmBN);
  uint32_t slices = (uint32_t)ix->sm_count / p.q_tiles;
  if (slices == 0) {
    set_error("gemm: nq=%u needs more query tiles than SMs", nq);
    return TSC_ERR_BAD_ARG;
  }
  if (slices > p.n_tiles) slices = p.n_tiles;
  p.n_slices = slices;
  p.kprime = kprime;
  p.metric = ix->desc.metric;
  p.norm2 = ix->d_norm2;
  p.live_mask = (ix->has_deleted || ix->has_filter) ? ix->d_live : nullptr;
  p.cand = d_cand;
  p.dbg_keys = dbg_keys;
  // CTA pairs (cta_group::2) whenever the query tiles pair up evenly
  bool pair = (p.q_tiles % 2) == 0;
  if (diag_env("TSC_GEMM_2CTA", -1) >= 0) pair = diag_env("TSC_GEMM_2CTA", -1) != 0;
  if (tf32) {   // fp32 storage, tf32 multiply: the queries are used as they are (fp32, padded)
    if (kprime <= 20)
      return launch_pair_or_single<20, 1>(ix, p, nq, kprime, dbg_keys, out_lists, st, d_q, pair);
    return launch_pair_or_single<32, 1>(ix, p, nq, kprime, dbg_keys, out_lists, st, d_q, pair);
  }
  if (kprime <= 20)
    return launch_pair_or_single<20, 0>(ix, p, nq, kprime, dbg_keys, out_lists, st, ix->d_q16, pair);
  return launch_pair_or_single<32, 0>(ix, p, nq, kprime, dbg_keys, out_lists, st, ix->d_q16, pair);
}

}  // namespace tsc
