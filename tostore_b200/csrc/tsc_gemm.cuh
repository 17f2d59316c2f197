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
//   warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer | warps 2-5: epilogue
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
constexpr int kGemmThreads = 192;
constexpr int kGemmEpiThreads = 128;
constexpr int kGemmMaxKp = 32;    // largest K' the in-smem lists support

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
  uint64_t *cand;            // [nq][n_slices][kprime]
  float *dbg_keys;           // optional [nq][n_rows] (tests only)
};

// ---- PTX wrappers -----------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int32_t c0, int32_t c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// mbarrier arrives once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// kind::f16 instruction descriptor: fp32 accumulate, A/B both K-major
__host__ __device__ inline uint32_t umma_idesc_f16(int dtype, int m, int n) {
  uint32_t fmt = dtype == kBF16 ? 1u : 0u;  // F16 = 0, BF16 = 1
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// smem: [stages x (A 16 KB | B 32 KB)] [scale 2x256 f32][bias 2x256 f32]
//       [lists: keys kp x 128 f32 | rows kp x 128 u32] [barriers] [tmem ptr]
__host__ __device__ inline size_t gemm_smem_bytes(uint32_t stages, uint32_t kprime) {
  return 1024 /* alignment slack */ + (size_t)stages * (16384 + 32768) + 2 * 2 * 256 * 4 +
         (size_t)kprime * 128 * 8 + (2 * stages + 4) * 8 + 16;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_b,
                 const GemmParams p, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 needs 1024 B alignment
  uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t S = p.stages;
  constexpr uint32_t kStageBytes = 16384 + 32768;
  float *s_scale = reinterpret_cast<float *>(sm + (size_t)S * kStageBytes);  // [2][256]
  float *s_bias = s_scale + 2 * 256;                                         // [2][256]
  float *l_keys = s_bias + 2 * 256;                                          // [kp][128]
  uint32_t *l_rows = reinterpret_cast<uint32_t *>(l_keys + (size_t)p.kprime * 128);
  uint64_t *bars = reinterpret_cast<uint64_t *>(l_rows + (size_t)p.kprime * 128);
  uint64_t *full = bars, *empty = bars + S, *tfull = bars + 2 * S, *tempty = bars + 2 * S + 2;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t qt = blockIdx.x % p.q_tiles;
  const uint32_t slice = blockIdx.x / p.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_b);
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(&tfull[a]), 1);
      mbar_init(smem_u32(&tempty[a]), kGemmEpiThreads);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint64_t pol_q = policy_evict_normal();  // queries are re-read by every tile
      const uint64_t pol_b = policy_evict_normal();  // corpus tile is shared by q_tiles CTAs
      uint32_t s = 0, ph = 0;
      for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
        for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
          const uint32_t bar = smem_u32(&full[s]);
          mbar_expect_tx(bar, kStageBytes);
          const uint32_t a_dst = base + s * kStageBytes, b_dst = a_dst + 16384;
          tma_load_2d(a_dst, &map_q, bar, (int32_t)(kb * kGemmBK), (int32_t)(qt * kGemmBM), pol_q);
          tma_load_2d(b_dst, &map_b, bar, (int32_t)(kb * kGemmBK), (int32_t)(ct * kGemmBN), pol_b);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0, as = 0, aph = 0;
      for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
        mbar_wait(smem_u32(&tempty[as]), aph ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kGemmBN;
        for (uint32_t kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(smem_u32(&full[s]), ph);
          tc_fence_after();
          const uint32_t a_addr = base + s * kStageBytes, b_addr = a_addr + 16384;
#pragma unroll
          for (int k = 0; k < kGemmBK / kGemmUK; k++) {
            umma_ss(d_tmem, umma_desc_sw128(a_addr + k * kGemmUK * 2),
                    umma_desc_sw128(b_addr + k * kGemmUK * 2), idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(&empty[s]));  // frees the smem stage when these MMAs retire
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        umma_commit(smem_u32(&tfull[as]));   // accumulator complete
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: thread <-> one query, running top-K' =====
    const int et = threadIdx.x - 64;                 // 0..127
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int qlane = quad * 32 + lane;              // query row inside the tile
    const uint32_t q = qt * kGemmBM + qlane;
    const uint32_t kp = p.kprime;
    for (uint32_t j = 0; j < kp; j++) {
      l_keys[j * 128 + qlane] = __int_as_float(0x7F800000);
      l_rows[j * 128 + qlane] = kInvalidRow;
    }
    float thr = __int_as_float(0x7F800000);
    uint32_t as = 0, aph = 0;
    for (uint32_t ct = slice; ct < p.n_tiles; ct += p.n_slices) {
      // per-column scale / bias for this tile (dead or out-of-range rows -> NaN key)
      const uint64_t row0 = (uint64_t)ct * kGemmBN;
      for (int c = et; c < kGemmBN; c += kGemmEpiThreads) {
        uint64_t n = row0 + c;
        float sc = __int_as_float(0x7FC00000), bi = 0.0f;
        bool live = n < p.n_rows;
        if (live && p.live_mask) live = (p.live_mask[n >> 5] >> (n & 31)) & 1u;
        if (live) {
          if (p.metric == kIP) {
            sc = -1.0f;
          } else {
            float n2 = p.norm2[n];
            if (p.metric == kL2) { sc = -2.0f; bi = n2; }
            else { sc = n2 > 0.0f ? -rsqrtf(n2) : 0.0f; }
          }
        }
        s_scale[as * 256 + c] = sc;
        s_bias[as * 256 + c] = bi;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      mbar_wait(smem_u32(&tfull[as]), aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * kGemmBN;
#pragma unroll 1
      for (int c0 = 0; c0 < kGemmBN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float sc = s_scale[as * 256 + c0 + j], bi = s_bias[as * 256 + c0 + j];
          float key = fmaf(__uint_as_float(v[j]), sc, bi) + 0.0f;
          if (p.dbg_keys && q < p.nq && row0 + c0 + j < p.n_rows)
            p.dbg_keys[(size_t)q * p.n_rows + row0 + c0 + j] = key;
          if (key < thr) {
            // insert into this thread's sorted list (ascending), dropping the last
            uint32_t pos = kp - 1;
            while (pos > 0 && key < l_keys[(pos - 1) * 128 + qlane]) {
              l_keys[pos * 128 + qlane] = l_keys[(pos - 1) * 128 + qlane];
              l_rows[pos * 128 + qlane] = l_rows[(pos - 1) * 128 + qlane];
              pos--;
            }
            l_keys[pos * 128 + qlane] = key;
            l_rows[pos * 128 + qlane] = (uint32_t)(row0 + c0 + j);
            thr = l_keys[(kp - 1) * 128 + qlane];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[as]));
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (q < p.nq) {
      uint64_t *out = p.cand + ((size_t)q * p.n_slices + slice) * kp;
      for (uint32_t j = 0; j < kp; j++) {
        uint32_t r = l_rows[j * 128 + qlane];
        out[j] = r == kInvalidRow ? ~0ull
                                  : (((uint64_t)ordered_key(l_keys[j * 128 + qlane]) << 32) | r);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- K4: per-row sum of squares of the stored values (fp32), one warp per row ------
template <int DTYPE>
__global__ void row_norms_kernel(const uint8_t *rows, uint64_t first, uint64_t n, uint32_t ld,
                                 uint32_t row_bytes, float *norm2) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
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
  }
}

// fp32 queries [nq, qld] -> 16-bit [nq, qld] in the corpus storage type (RNE)
__global__ void convert_queries_kernel(const float *src, uint32_t n, uint16_t *dst, int dtype) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (dtype == kBF16) {
      __nv_bfloat16 b = __float2bfloat16_rn(src[i]);
      dst[i] = *reinterpret_cast<uint16_t *>(&b);
    } else {
      __half h = __float2half_rn(src[i]);
      dst[i] = *reinterpret_cast<uint16_t *>(&h);
    }
  }
}

}  // namespace tsc
