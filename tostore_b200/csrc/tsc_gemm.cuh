// tsc_gemm.cuh — K2: batched query x corpus contraction on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with a fused
// per-query running top-K' epilogue, so the nq x N score matrix never exists.
//
// It replaces, for batches of queries, the same candidate-generation stage as the
// scan kernel (tsc_scan.cuh; reference: ADC beam search,
// core/ngh_graph_engine.dart:98-113). The reference has no batch API and no 16-bit
// storage; this path is additive. Keys are fp32 accumulations of
// exact bf16/f16 products; the K' survivors are re-ranked in exact fp64 by
// tsc_select.cuh, which is what the caller sees.
//
// Shape of the computation (one CTA, 192 threads, persistent):
//   D[128 queries (TMEM lanes) x 256 corpus rows (TMEM columns)] += A x B^T
//   A = query tile  [128 x 64] bf16/f16, K-major, 128B swizzle (TMA)
//   B = corpus tile [256 x 64] bf16/f16, K-major, 128B swizzle (TMA)
//   warps 0-7: epilogue | warp 8: TMA producer | warp 9: TMEM alloc + MMA issuer
// A CTA keeps ONE query tile for its whole life and walks every (G/QT)-th corpus
// tile, so each epilogue thread owns one query: its threshold lives in a register
// and its sorted candidate list in shared memory, with no cross-thread reduction.
// CTAs that share a corpus tile (different query tiles) run side by side, so the
// corpus streams from HBM once and is re-read from L2.
#pragma once

#include <cuda.h>

#include "tsc_common.cuh"

namespace tsc {

constexpr int kGemmBM = 128;      // queries per tile (UMMA M)
constexpr int kGemmBN = 256;      // corpus rows per tile (UMMA N)
constexpr int kGemmBK = 64;       // K elements per stage (128 bytes = one swizzle row)
constexpr int kGemmUK = 16;       // UMMA K for 16-bit inputs
constexpr int kGemmThreads = 320;
constexpr int kGemmEpiThreads = 256;
// warp roles: 0-7 epilogue, 8 TMA producer, 9 MMA issuer. The issue arbiter favours
// the highest warp id on a scheduler, so the two latency-critical single-lane roles
// sit above the epilogue warps they share schedulers with.
constexpr int kWarpTma = 8, kWarpMma = 9;

struct GemmParams {
  uint64_t n_rows;           // corpus rows in the shard
  uint32_t nq;               // queries in the batch
  uint32_t k_blocks;         // ceil(dims / 64)
  uint32_t q_tiles;          // ceil(nq / 128)
  uint32_t n_slices;         // gridDim.x / q_tiles
  uint32_t n_tiles;          // ceil(n_rows / 256)
  uint32_t kprime;
  uint32_t stages;
  int metric;
  const float *norm2;        // [n_rows] sum of squares of the stored (rounded) row
  const uint32_t *live_mask; // optional
  uint64_t *cand;            // [nq][n_slices * 2][kprime] (two column halves per CTA)
  float *dbg_keys;           // optional [nq][n_rows] (tests only)
  uint32_t exp_flags;        // perf experiments (EXP kernels only; results invalid when non-zero):
                             // 1 = producer re-loads one corpus tile, 2 = epilogue skips the TMEM
                             // read and the arithmetic, 4 = no TMA at all (MMA issue rate only),
                             // 8 = skip A loads, 16 = epilogue TMEM reads only, 32 = arithmetic only
  unsigned long long *prof;  // EXP kernels: per-role wait / work cycle sums (kProf* slots)
};
// slots of GemmParams::prof (sums over all CTAs, atomicAdd by one lane per role)
enum : int { kProfProdWaitEmpty = 0, kProfProdTotal, kProfMmaWaitFull, kProfMmaWaitTempty,
             kProfMmaTotal, kProfEpiWaitTfull, kProfEpiLdtm, kProfEpiMath,
             kProfEpiTotal, kProfProdN, kProfMmaN, kProfEpiN, kProfSlots };

// ---- PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// pull one box into L2 only (no shared-memory slot, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// one lane of a converged warp; lets ptxas keep MMA / TMA operands in uniform
// registers instead of emitting a per-lane R2UR waterfall around every UTCHMMA
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] x B[smem desc]; issued by one thread
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with fp32 operands read as tf32 (10-bit mantissa, low 13 bits ignored): UMMA K = 8
__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile whose rows are 128 bytes (64 x 16-bit):
// 8-row swizzle atoms of 1024 B, SBO = 1024 B, LBO unused, descriptor version 1
// (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // SBO, bits [32,46)
  d |= (uint64_t)1 << 46;                               // version = 1
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, A/B both K-major
__host__ __device__ inline uint32_t umma_idesc_f16(int dtype, int m, int n) {
  uint32_t fmt = dtype == kBF16 ? 1u : (dtype == kF32 ? 2u : 0u);  // F16 = 0, BF16 = 1, TF32 = 2
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// Insert into one thread's sorted (ascending) list, column-major in smem
// ([kp][128 threads]); returns the new threshold. Equal keys keep arrival (row) order.
__device__ __noinline__ float gemm_list_insert(float *l_keys, uint32_t *l_rows, uint32_t kp,
                                               int qlane, float key, uint32_t row) {
  uint32_t pos = kp - 1;
  while (pos > 0 && key < l_keys[(pos - 1) * kGemmEpiThreads + qlane]) {
    l_keys[pos * kGemmEpiThreads + qlane] = l_keys[(pos - 1) * kGemmEpiThreads + qlane];
    l_rows[pos * kGemmEpiThreads + qlane] = l_rows[(pos - 1) * kGemmEpiThreads + qlane];
    pos--;
  }
  l_keys[pos * kGemmEpiThreads + qlane] = key;
  l_rows[pos * kGemmEpiThreads + qlane] = row;
  return l_keys[(kp - 1) * kGemmEpiThreads + qlane];
}

// ---- CTA-pair (cta_group::2) helpers ------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// TMA load whose completion bytes are signalled on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map,
                                                uint32_t cluster_bar, int32_t c0, int32_t c1,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ss2_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this smem
// offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

// ====================================================================================
// The SS kernel, for one CTA (CG = 1) or a CTA pair (CG = 2, cta_group::2).
//
// CG = 2: two SMs of a cluster cooperate on a 256 (queries) x 256 (corpus rows)
// tile. Each CTA stages its own 128 queries and only HALF of the corpus tile; one
// tcgen05.mma.cta_group::2 issued by the leader drives both tensor cores (M = 256:
// rows 0-127 from the leader's A, 128-255 from the peer's; N = 256: columns 0-127
// from the leader's B half, 128-255 from the peer's). The bytes an SM has to ingest
// per MMA cycle drop from 96 to 64 - the limit measured for the 1-CTA shape.
//
// Epilogue warps never synchronise with each other: each warp prepares the per-column
// coefficients of its own 128 columns, so a warp that is busy inserting candidates
// (frequent while the lists are still filling) does not hold the others back.
// ====================================================================================
template <int CG>
struct GemmGeom {
  static constexpr uint32_t kBRows = 256 / CG;                  // corpus rows this CTA stages
  static constexpr uint32_t kStageBytes = 16384 + kBRows * 128; // A 16 KB + B
};
// smem: [stages x (A | B)] [coeffs: 8 warps x 2 bufs x {scale,bias} x 128]
//       [candidate row ids: kprime x 256] [barriers]   (candidate keys live in registers)
template <int CG>
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t stages, uint32_t kprime) {
  return 1024 + (size_t)stages * GemmGeom<CG>::kStageBytes + 8 * 2 * 128 * 2 * 4 +
         (size_t)kprime * kGemmEpiThreads * 4 + (2 * stages + 4) * 8 + 16;
}

// dynamic pick of one of 32 registers: 31 selects, no local memory
__device__ __forceinline__ float pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
  for (int i = 0; i < 8; i++) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
  for (int i = 0; i < 4; i++) c[i] = (j & 4) ? b[i + 4] : b[i];
#pragma unroll
  for (int i = 0; i < 2; i++) d[i] = (j & 2) ? c[i + 2] : c[i];
  return __uint_as_float((j & 1) ? d[1] : d[0]);
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// KPR = capacity of the per-thread candidate list (registers): 20 or 32 >= kprime
// (10 warps put 3 on one scheduler: 16,384 / 96 caps a thread at 168 registers, which is
// why the epilogue reads its 128 columns in two batches of 64 instead of all at once.)
// KIND = 0: 16-bit storage (kind::f16, 64 elements per 128-byte K block);
// KIND = 1: fp32 storage multiplied as tf32 (kind::tf32, 32 elements per K block; opt-in,
// TSC_GEMM_TF32=1). The byte geometry of a stage is identical for both.
template <bool DBG, int CG, bool EXP, int KPR, int KIND = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_b,
                 const GemmParams p, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 needs 1024 B alignment
  uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t S = p.stages;
  constexpr uint32_t kStageBytes = GemmGeom<CG>::kStageBytes;
  float *s_coef = reinterpret_cast<float *>(sm + (size_t)S * kStageBytes);  // [8][2][2][128]
  uint32_t *l_rows = reinterpret_cast<uint32_t *>(s_coef + 8 * 2 * 2 * 128);  // [kp][256]
  uint64_t *bars = reinterpret_cast<uint64_t *>(l_rows + (size_t)p.kprime * kGemmEpiThreads);
  uint64_t *full = bars, *empty = bars + S, *tfull = bars + 2 * S, *tempty = bars + 2 * S + 2;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const uint32_t unit = blockIdx.x / CG;                      // CTA (pair) index
  const uint32_t q_units = (p.q_tiles + CG - 1) / CG;
  const uint32_t qt = (unit % q_units) * CG + crank;          // this CTA's query tile
  const uint32_t slice = unit / q_units;
  const uint32_t xf = EXP ? p.exp_flags : 0u;
  constexpr int kBKElems = KIND == 1 ? kGemmBK / 2 : kGemmBK;   // elements per 128-byte K block
  long long t_w0 = 0, t_w1 = 0, t_w2 = 0;   // EXP: cycle sums of this warp's role
  auto tick = [&]() -> long long { return EXP ? clock64() : 0ll; };

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_b);
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(&tfull[a]), 1);
      mbar_init(smem_u32(&tempty[a]), CG * 8);   // one arrival per epilogue warp (of both CTAs)
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) {
    if (CG == 2) tmem_alloc2(smem_u32(s_tmem), 512);
    else tmem_alloc(smem_u32(s_tmem), 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == kWarpTma) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    const uint64_t pol_q = policy_evict_normal();  // queries are re-read by every tile
    const uint64_t pol_b = policy_evict_normal();  // corpus tile is shared by the slice's CTAs
    uint32_t s = 0, ph = 0;
    const long long t_beg = tick();
    for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
      if (xf & 4u) break;
      for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
        const long long t0 = tick();
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        t_w0 += tick() - t0;
        if (elect_one()) {
          const uint32_t a_dst = base + s * kStageBytes, b_dst = a_dst + 16384;
          const uint32_t ctl = (xf & 1u) ? slice : ct;
          const bool skip_a = (xf & 8u) != 0;
          const uint32_t bytes = skip_a ? kStageBytes - 16384 : kStageBytes;
          if (CG == 2) {
            // bytes of BOTH CTAs are counted on the leader's barrier: its MMA eats both halves
            const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
            if (crank == 0) mbar_expect_tx(smem_u32(&full[s]), 2 * bytes);
            if (!skip_a)
              tma_load_2d_2sm(a_dst, &map_q, bar, (int32_t)(kb * kBKElems),
                              (int32_t)(qt * kGemmBM), pol_q);
            tma_load_2d_2sm(b_dst, &map_b, bar, (int32_t)(kb * kBKElems),
                            (int32_t)(ctl * kGemmBN + crank * GemmGeom<CG>::kBRows), pol_b);
          } else {
            const uint32_t bar = smem_u32(&full[s]);
            mbar_expect_tx(bar, bytes);
            if (!skip_a)
              tma_load_2d(a_dst, &map_q, bar, (int32_t)(kb * kBKElems), (int32_t)(qt * kGemmBM),
                          pol_q);
            tma_load_2d(b_dst, &map_b, bar, (int32_t)(kb * kBKElems), (int32_t)(ctl * kGemmBN),
                        pol_b);
          }
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
    if (EXP && lane == 0 && p.prof) {
      atomicAdd(p.prof + kProfProdWaitEmpty, (unsigned long long)t_w0);
      atomicAdd(p.prof + kProfProdTotal, (unsigned long long)(tick() - t_beg));
      atomicAdd(p.prof + kProfProdN, 1ull);
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (whole warp walks the loop, one elected lane issues; with a CTA
    // pair only the leader issues and its commits are multicast to both CTAs) =====
    if (crank == 0) {
      uint32_t s = 0, ph = 0, as = 0, aph = 0;
      const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
      const long long t_beg = tick();
      for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
        long long t0 = tick();
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1u);  // epilogues have drained this accumulator
        t_w1 += tick() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kGemmBN;
        for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
          t0 = tick();
          if (!(xf & 4u)) mbar_wait(smem_u32(&full[s]), ph);
          t_w0 += tick() - t0;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = base + s * kStageBytes, b_addr = a_addr + 16384;
            const uint32_t a_lo = (uint32_t)umma_desc_sw128(a_addr);
            const uint32_t b_lo = (uint32_t)umma_desc_sw128(b_addr);
#pragma unroll
            for (int k = 0; k < kGemmBK / kGemmUK; k++) {
              // +32 bytes per K step inside the 128-byte swizzle row (units of 16 B)
              if (KIND == 1 && CG == 2)
                umma_ss2_tf32(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                              (kb | k) != 0);
              else if (KIND == 1)
                umma_ss_tf32(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                             (kb | k) != 0);
              else if (CG == 2)
                umma_ss2(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                         (kb | k) != 0);
              else
                umma_ss(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                        (kb | k) != 0);
            }
            if (CG == 2) {
              umma_commit2(smem_u32(&empty[s]));  // frees the stage in both CTAs
              if (kb + 1 == p.k_blocks) umma_commit2(smem_u32(&tfull[as]));
            } else {
              umma_commit(smem_u32(&empty[s]));
              if (kb + 1 == p.k_blocks) umma_commit(smem_u32(&tfull[as]));
            }
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
      if (EXP && lane == 0 && p.prof) {
        atomicAdd(p.prof + kProfMmaWaitFull, (unsigned long long)t_w0);
        atomicAdd(p.prof + kProfMmaWaitTempty, (unsigned long long)t_w1);
        atomicAdd(p.prof + kProfMmaTotal, (unsigned long long)(tick() - t_beg));
        atomicAdd(p.prof + kProfMmaN, 1ull);
      }
    }
  } else {
    // ===== epilogue: 8 warps; thread <-> (query, half of the tile's columns) =====
    // TMEM lane quadrant = warp % 4 (hardware rule); warps 0-3 take columns [0,128) of
    // every tile, warps 4-7 columns [128,256). Each thread keeps the K' best keys it has
    // seen in REGISTERS (unordered; `thr` = their maximum) and the matching row ids in
    // shared memory, so a query has two candidate lists per CTA. An insertion replaces
    // the current maximum: ~80 register-only instructions and one fire-and-forget STS.
    // (A sorted list in shared memory cost ~2,000 cycles per insertion - a chain of
    // dependent LDS behind the tensor core's operand traffic - and with ~6,700
    // insertions per warp the epilogue, not the MMA, paced the kernel.)
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int qlane = quad * 32 + lane;              // query row inside the tile
    const int lidx = half * 128 + qlane;             // list slot, 0..255
    const uint32_t q = qt * kGemmBM + qlane;
    const uint32_t kp = p.kprime;
    const uint32_t col_base = (uint32_t)half * 128;
    float *w_coef = s_coef + (size_t)warp * (2 * 2 * 128);   // [buf][scale|bias][128]
    const uint32_t tempty_bar[2] = {
        CG == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : smem_u32(&tempty[0]),
        CG == 2 ? mapa_u32(smem_u32(&tempty[1]), 0) : smem_u32(&tempty[1])};
    float lk[KPR];   // slots >= kp hold -inf: never the maximum, never replaced, never emitted
#pragma unroll
    for (int j = 0; j < KPR; j++)
      lk[j] = __int_as_float((uint32_t)j < kp ? 0x7F800000u : 0xFF800000u);
    for (uint32_t j = 0; j < kp; j++) l_rows[j * kGemmEpiThreads + lidx] = kInvalidRow;
    float thr = __int_as_float(0x7F800000);
    uint32_t as = 0, aph = 0;

    // per-column scale / bias: key = fma(acc, scale, bias), scale <= 0 (-1, -2, -1/|b|),
    // bias >= 0 (0 or |b|^2); dead or out-of-range rows get a NaN scale -> NaN key. Lane l
    // prepares columns l, l+32, l+64, l+96 of the warp's half and prefetches the next
    // tile's values one tile ahead so the global-load latency is off the critical path.
    float sc_n[4], bi_n[4];
    auto column_coeffs = [&](uint32_t ct) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint64_t n = (uint64_t)ct * kGemmBN + col_base + j * 32 + lane;
        float sc = __int_as_float(0x7FC00000), bi = 0.0f;
        bool live = ct < p.n_tiles && n < p.n_rows;
        if (live && p.live_mask) live = (p.live_mask[n >> 5] >> (n & 31)) & 1u;
        if (live) {
          if (p.metric == kIP) {
            sc = -1.0f;
          } else {
            const float n2 = __ldg(p.norm2 + n);
            if (p.metric == kL2) { sc = -2.0f; bi = n2; }
            else { sc = n2 > 0.0f ? -rsqrtf(n2) : 0.0f; }
          }
        }
        sc_n[j] = sc;
        bi_n[j] = bi;
      }
    };
    column_coeffs(slice);

    // ---- the hot loop never touches shared memory ---------------------------------
    // Measured (profiles/r01b_gemm_roles.txt): with per-column coefficients read from
    // shared memory for every accumulator, the epilogue's LDS traffic competes with the
    // tensor core's operand reads and TMA's writes for the same port and the arithmetic
    // of one tile took 5,400 cycles - longer than the tile's MMAs. Instead each thread
    // tests the RAW accumulators against a bound R that is necessary for any column of
    // this warp's 128 to beat the thread's threshold:
    //     key = acc*sc + bi < thr, sc in [sc_min, 0], bi >= bi_min, thr < bi_min
    //       =>  acc > (bi_min - thr) / |sc_min|  =: R
    // (thr >= bi_min, e.g. while the list is still filling: R = -inf, everything is
    // examined). R is shrunk by 2^-18 so that fp32 rounding of R itself can never hide
    // a pass; fma(acc, sc, bi) is a single exactly rounded operation, so acc <= R
    // implies key >= thr. The common path is 16 three-input maxima per 32 columns;
    // only chunks with an accumulator above R read their coefficients.
    float bi_min = __int_as_float(0x7F800000), rcp_sc = 0.0f;   // per tile (warp-uniform)
    float R = __int_as_float(0xFF800000);
    auto bound_for = [&](float t) -> float {
      return t < bi_min ? (bi_min - t) * rcp_sc * (1.0f - 3.8146973e-06f)
                        : __int_as_float(0xFF800000);
    };
    auto list_insert = [&](float key, uint32_t row) {
      int slot = 0;
#pragma unroll
      for (int j = KPR - 1; j >= 0; j--)
        if (lk[j] == thr) slot = j;
#pragma unroll
      for (int j = 0; j < KPR; j++) lk[j] = (j == slot) ? key : lk[j];
      l_rows[slot * kGemmEpiThreads + lidx] = row;
      float m = fmaxf(lk[0], lk[1]);
#pragma unroll
      for (int j = 2; j + 1 < KPR; j += 2) m = fmax3(m, lk[j], lk[j + 1]);
      thr = m;
    };

    auto process = [&](const uint32_t (&v)[32], uint32_t c0, uint64_t row0) {
      const float *sc_s = w_coef + as * 256 + c0;
      const float *bi_s = sc_s + 128;
      const uint64_t r0 = row0 + col_base + c0;
      if (DBG) {
        if (q < p.nq)
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (r0 + j < p.n_rows)
              p.dbg_keys[(size_t)q * p.n_rows + r0 + j] =
                  fmaf(__uint_as_float(v[j]), sc_s[j], bi_s[j]) + 0.0f;
      }
      // Padded query rows (q >= nq: zero queries) would tie every column at one key, keep
      // R at -inf and send all 256 columns of every tile through the exact test: measured
      // (8 queries in a 128-row tile, profiles/r02_small_batches.txt) 7 x the instructions of
      // a full tile and a kernel at 23 % of HBM. They have nothing to find: skip them.
      if (q >= p.nq) return;
      const float *f = reinterpret_cast<const float *>(&v[0]);
      float g[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        g[i] = fmax3(fmax3(f[8 * i], f[8 * i + 1], f[8 * i + 2]),
                     fmax3(f[8 * i + 3], f[8 * i + 4], f[8 * i + 5]),
                     fmaxf(f[8 * i + 6], f[8 * i + 7]));
      if (fmaxf(fmax3(g[0], g[1], g[2]), g[3]) > R) {
        // rare path: which accumulators exceed the bound, then one exact test each
        uint32_t pass = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (g[i] > R) {
#pragma unroll
            for (int j = 8 * i; j < 8 * i + 8; j++)
              if (f[j] > R) pass |= 1u << j;
          }
        while (pass) {
          const int j = __ffs(pass) - 1;
          pass &= pass - 1;
          const float key = fmaf(pick32(v, j), sc_s[j], bi_s[j]);
          if (key < thr) {
            list_insert(key + 0.0f, (uint32_t)(r0 + j));
            R = bound_for(thr);
          }
        }
      }
    };

    const long long t_beg = tick();
    for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
      const uint64_t row0 = (uint64_t)ct * kGemmBN;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        w_coef[as * 256 + j * 32 + lane] = sc_n[j];
        w_coef[as * 256 + 128 + j * 32 + lane] = bi_n[j];
      }
      {
        // warp-wide bounds of this tile's live columns (NaN scale = dead column)
        float lo_sc = 0.0f, lo_bi = __int_as_float(0x7F800000);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          lo_sc = fminf(lo_sc, sc_n[j]);
          if (sc_n[j] == sc_n[j]) lo_bi = fminf(lo_bi, bi_n[j]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo_sc = fminf(lo_sc, __shfl_xor_sync(0xFFFFFFFFu, lo_sc, o));
          lo_bi = fminf(lo_bi, __shfl_xor_sync(0xFFFFFFFFu, lo_bi, o));
        }
        bi_min = lo_bi;
        rcp_sc = 1.0f / (-lo_sc);          // +inf when no live column has a non-zero scale
        R = bound_for(thr);
      }
      __syncwarp();
      column_coeffs(ct + p.n_slices);  // prefetch for the next tile
      long long t0 = tick();
      mbar_wait(smem_u32(&tfull[as]), aph);
      t_w0 += tick() - t0;
      tc_fence_after();
      const uint32_t t_addr =
          tmem_base + ((uint32_t)(quad * 32) << 16) + as * kGemmBN + col_base;
      if (EXP && (xf & (2u | 32u))) {   // experiments: no TMEM read (2), arithmetic on junk (32)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty_bar[as]);
          else mbar_arrive(tempty_bar[as]);
        }
        if (xf & 32u) {
          uint32_t va[32], vb[32];
#pragma unroll
          for (int j = 0; j < 32; j++) { va[j] = 0x3F800000u + j * 4099u + ct; vb[j] = va[j] ^ 0x12345u; }
          process(va, 0, row0);
          process(vb, 32, row0);
          process(va, 64, row0);
          process(vb, 96, row0);
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
        continue;
      }
      t0 = tick();
      // Two batches of 64 columns (a TMEM read costs ~35 cycles, measured); the buffer
      // goes back to the MMA issuer as soon as the second batch is in registers.
      uint32_t va[32], vb[32];
      tmem_ld32_nowait(t_addr, va);
      tmem_ld32_nowait(t_addr + 32, vb);
      tmem_ld_wait();
      t_w1 += tick() - t0;
      t0 = tick();
      const bool ld_only = EXP && (xf & 16u);   // experiment: TMEM reads only
      if (ld_only) {
        if ((va[0] ^ vb[31]) == 0x7FC12345u) thr = 0.0f;
      } else {
        process(va, 0, row0);
        process(vb, 32, row0);
      }
      t_w2 += tick() - t0;
      t0 = tick();
      tmem_ld32_nowait(t_addr + 64, va);
      tmem_ld32_nowait(t_addr + 96, vb);
      tmem_ld_wait();
      t_w1 += tick() - t0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_bar[as]);
        else mbar_arrive(tempty_bar[as]);
      }
      t0 = tick();
      if (ld_only) {
        if ((va[5] ^ vb[7]) == 0x7FC12345u) thr = 0.0f;
      } else {
        process(va, 64, row0);
        process(vb, 96, row0);
      }
      t_w2 += tick() - t0;
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (EXP && lane == 0 && p.prof) {
      atomicAdd(p.prof + kProfEpiWaitTfull, (unsigned long long)t_w0);
      atomicAdd(p.prof + kProfEpiLdtm, (unsigned long long)t_w1);
      atomicAdd(p.prof + kProfEpiMath, (unsigned long long)t_w2);
      atomicAdd(p.prof + kProfEpiTotal, (unsigned long long)(tick() - t_beg));
      atomicAdd(p.prof + kProfEpiN, 1ull);
    }
    if (q < p.nq) {
      uint64_t *out = p.cand + ((size_t)q * p.n_slices * 2 + slice * 2 + half) * kp;
#pragma unroll
      for (int j = 0; j < KPR; j++) {
        if ((uint32_t)j < kp) {
          const uint32_t r = l_rows[j * kGemmEpiThreads + lidx];
          out[j] = r == kInvalidRow ? ~0ull : (((uint64_t)ordered_key(lk[j]) << 32) | r);
        }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // peer arrivals on the leader's barriers must land first
  else __syncthreads();
  if (warp == kWarpMma) {
    if (CG == 2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// K4: sum of squares of every stored row (fp32, lane-strided fma + butterfly: the scan
// kernel's accumulation order, so the certificate's scan_eps covers it) and the running
// maximum of the shard (bits of a non-negative float order like the float).
template <int DTYPE>
__global__ void row_norms_kernel(const uint8_t *rows, uint64_t first, uint64_t n, uint32_t ld,
                                 uint32_t row_bytes, float *norm2, uint32_t *maxnorm) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  float mx = 0.0f;
  for (uint64_t r = first + w; r < first + n; r += nw) {
    const uint8_t *row = rows + r * row_bytes;
    float s = 0.0f;
    for (uint32_t i = lane; i < ld; i += 32) {
      float v = DTYPE == kF32 ? reinterpret_cast<const float *>(row)[i]
              : DTYPE == kBF16
                    ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t *>(row)[i] << 16)
                    : __half2float(reinterpret_cast<const __half *>(row)[i]);
      s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) norm2[r] = s;
    if (s == s) mx = fmaxf(mx, s); else mx = __int_as_float(0x7F800000);   // NaN row: no bound
  }
  if (lane == 0 && mx > 0.0f) atomicMax(maxnorm, __float_as_uint(mx));
}

// fp32 queries [nq, qld] -> 16-bit [nq, qld] in the corpus storage type (RNE), one warp per
// query; enorm[q] = |q16 - q|_2 rounded up: what the rounding can move a dot product by
// (Cauchy-Schwarz), the certificate's a_e term (tsc_tail.cuh).
__global__ void convert_queries_kernel(const float *src, uint32_t nq, uint32_t qld, uint16_t *dst,
                                       float *enorm, int dtype) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  float s = 0.0f;
  for (uint32_t i = lane; i < qld; i += 32) {
    const float v = src[(size_t)q * qld + i];
    float r;
    if (dtype == kBF16) {
      __nv_bfloat16 b = __float2bfloat16_rn(v);
      dst[(size_t)q * qld + i] = *reinterpret_cast<uint16_t *>(&b);
      r = __bfloat162float(b);
    } else {
      __half h = __float2half_rn(v);
      dst[(size_t)q * qld + i] = *reinterpret_cast<uint16_t *>(&h);
      r = __half2float(h);
    }
    const float e = v - r;   // exact: r is v rounded to fewer bits (or inf / NaN)
    s = fmaf(e, e, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) enorm[q] = sqrtf(s) * 1.0001f;
}

}  // namespace tsc
