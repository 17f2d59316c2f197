// tsc_common.cuh — shared device helpers for libtostore_cuda.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtostore_cuda targets sm_100a (B200) only"
#endif

namespace tsc {

constexpr int kWarp = 32;
constexpr uint32_t kInvalidRow = 0xFFFFFFFFu;
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;  // sorts after every real key
constexpr uint32_t kMaxRerank = 512;         // candidates the first pass of the tail re-ranks, at most
constexpr int kGemmMaxKp = 32;               // entries per candidate list of the tensor path, at most

// metric / dtype codes: include/tostore_cuda.h
enum : int { kL2 = 0, kIP = 1, kCos = 2 };
enum : int { kF32 = 0, kBF16 = 1, kF16 = 2 };

// ---- ordered keys ---------------------------------------------------------
// fp32 ranking key -> uint32 whose unsigned order equals the float order.
// NaN is folded into +inf first (such rows still rank, last; the exact fp64
// re-rank then restores the reference's NaN-last order among them).
__device__ __forceinline__ uint32_t ordered_key(float f) {
  if (!(f == f)) f = __int_as_float(0x7F800000);
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Dart double.compareTo total order as a uint64 key: -0.0 < +0.0, NaN last.
// (result ordering: core/ngh_graph_engine.dart:133, vector_index_manager.dart:587)
__device__ __forceinline__ uint64_t ordered_key64(double d) {
  if (!(d == d)) return 0xFFFFFFFFFFFFFFFFull;
  uint64_t b = (uint64_t)__double_as_longlong(d);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---- mbarrier / bulk-copy (TMA unit, SASS: UBLKCP) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// the memory of an mbarrier has to be invalidated before it is used for anything else
__device__ __forceinline__ void mbar_inval(uint32_t bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// 1-D bulk async copy global -> shared, completion signalled on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                         uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- 16-bit storage -> fp32 -------------------------------------------------
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float &lo, float &hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xFFFF0000u);
}

__device__ __forceinline__ void unpack_f16x2(uint32_t w, float &lo, float &hi) {
  __half2 h = *reinterpret_cast<__half2 *>(&w);
  float2 f = __half22float2(h);
  lo = f.x;
  hi = f.y;
}

// ---- synthetic corpus (bit-identical to oracle/tostore_oracle.c:tso_synth_value)
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t flat) {
  uint64_t z = seed + (flat + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  int32_t s = (int32_t)(int16_t)(z & 0xFFFF) + (int32_t)(int16_t)((z >> 16) & 0xFFFF) +
              (int32_t)(int16_t)((z >> 32) & 0xFFFF) + (int32_t)(int16_t)((z >> 48) & 0xFFFF);
  return (float)s * 3.0517578125e-05f;
}

}  // namespace tsc
