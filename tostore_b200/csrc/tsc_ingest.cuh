// tsc_ingest.cuh — K7 and friends: corpus ingestion kernels (HBM-bound byte work).
//
//  * convert_rows_kernel : host-format rows (f64 / f32 / i8) -> device storage
//    dtype with the reference's decode semantics
//    (NghRawVectorPage.getVectorAsFloat32, core/ngh_page.dart:368-389).
//  * page_check_kernel / page_decode_kernel : raw reference pages
//    (BTreePageHeader + NghRawVectorPage payload, core/btree_page.dart:132-234,
//    core/ngh_page.dart:414-447): header + CRC-32 validation on the GPU, then a
//    strided decode that strips the 28-byte headers and tail slack.
//  * graph_flags_kernel : tombstone bits from graph pages (ngh_page.dart:104-216).
//  * synth_rows_kernel : deterministic synthetic corpus for tests / benchmarks.
#pragma once

#include "tsc_common.cuh"

namespace tsc {

enum : int { kSrcF64 = 0, kSrcF32 = 1, kSrcI8 = 2 };

__device__ __forceinline__ float decode_src(const uint8_t *p, int prec) {
  if (prec == kSrcF32) {
    float f;
    memcpy(&f, p, 4);
    return f;
  }
  if (prec == kSrcF64) {
    double d;
    memcpy(&d, p, 8);
    return (float)d;  // Float32List store: round to nearest even
  }
  return (float)__ddiv_rn((double)(int8_t)p[0], 127.0);  // ngh_page.dart:384-386
}

__device__ __forceinline__ void store_dev(uint8_t *row, uint32_t i, float v, int dtype) {
  if (dtype == kF32)
    reinterpret_cast<float *>(row)[i] = v;
  else if (dtype == kBF16)
    reinterpret_cast<__nv_bfloat16 *>(row)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<__half *>(row)[i] = __float2half_rn(v);
}

// src: [n, dims] elements of `prec`, dense. dst rows: stride row_bytes, ld elems.
__global__ void convert_rows_kernel(const uint8_t *src, uint64_t n, uint32_t dims, int prec,
                                    uint32_t bpe, uint8_t *dst, uint32_t row_bytes, uint32_t ld,
                                    int dtype) {
  const uint64_t total = n * ld;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / ld;
    uint32_t c = (uint32_t)(i % ld);
    float v = (c < dims) ? decode_src(src + (r * dims + c) * bpe, prec) : 0.0f;
    store_dev(dst + r * row_bytes, c, v, dtype);
  }
}

__global__ void synth_rows_kernel(uint64_t seed, uint64_t first_row, uint64_t n, uint32_t dims,
                                  uint8_t *dst, uint32_t row_bytes, uint32_t ld, int dtype) {
  const uint64_t total = n * ld;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / ld;
    uint32_t c = (uint32_t)(i % ld);
    float v = (c < dims) ? synth_value(seed, (first_row + r) * dims + c) : 0.0f;
    store_dev(dst + r * row_bytes, c, v, dtype);
  }
}

// ---- pages -----------------------------------------------------------------
constexpr uint32_t kPageMagic = 0x32475054u;  // 'TPG2', btree_page.dart:134
constexpr uint32_t kPageHeader = 20;          // btree_page.dart:133
constexpr uint32_t kPtGraph = 6, kPtRawVec = 8;  // BTreePageType index, btree_page.dart:14-55

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}

// CRC-32/IEEE (Crc32.of, core/btree_page.dart:64-89), one warp per page: lane l
// folds slice l of the payload bytewise (table in shared memory; lane 0 starts
// from ~0, the others from 0), shifts its raw register past the bytes that
// follow the slice by multiplying with x^(8n) mod P in GF(2)[x] (reflected
// representation, bit 31 = x^0), and the 32 partial registers are XOR-ed: the
// register update is affine, so raw(A||B, i) = shift(raw(A, i), |B|) ^ raw(B, 0).
// status[page]: 0 ok, 1 bad magic/header, 2 bad length, 3 crc, 4 type, 5 payload, 6 dims,
// 7 precision
__host__ __device__ __forceinline__ uint32_t crc_bytes(const uint32_t *tab, uint32_t c,
                                                       const uint8_t *p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
  return c;
}
__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (uint32_t m = 0x80000000u; m != 0; m >>= 1) {
    if (a & m) p ^= b;
    b = (b & 1u) ? ((b >> 1) ^ 0xEDB88320u) : (b >> 1);
  }
  return p;
}
// x^(8 * nbytes) mod P
__host__ __device__ __forceinline__ uint32_t gf2_x8n(uint32_t nbytes) {
  uint32_t r = 0x80000000u;  // x^0
  uint32_t sq = 0x00800000u; // x^8
  while (nbytes) {
    if (nbytes & 1u) r = gf2_mulmod(sq, r);
    sq = gf2_mulmod(sq, sq);
    nbytes >>= 1;
  }
  return r;
}
__host__ __device__ __forceinline__ uint32_t crc_shift(uint32_t c, uint32_t nbytes) {
  return nbytes == 0 ? c : gf2_mulmod(gf2_x8n(nbytes), c);
}

__global__ void page_check_kernel(const uint8_t *pages, uint64_t n_pages, uint32_t page_size,
                                  uint32_t expect_type, uint32_t expect_dims, uint32_t expect_prec,
                                  uint32_t *status, uint32_t *bad_count) {
  __shared__ uint32_t tab[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    tab[i] = c;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t pg = warp; pg < n_pages; pg += nwarps) {
    const uint8_t *p = pages + pg * page_size;
    uint32_t st = 0, len = 0;
    if (ld_u32(p) != kPageMagic || ld_u16(p + 4) != kPageHeader || p[6] >= 10) {
      st = 1;
    } else {
      len = ld_u32(p + 8);
      if ((uint64_t)kPageHeader + len > page_size) st = 2;
    }
    if (st == 0) {
      const uint8_t *pay = p + kPageHeader;
      uint32_t per = (len + 31) / 32;
      uint32_t lo = min(len, per * lane), hi = min(len, per * (lane + 1));
      uint32_t c = crc_bytes(tab, lane == 0 ? 0xFFFFFFFFu : 0u, pay + lo, hi - lo);
      c = crc_shift(c, len - hi);  // past the bytes that follow my slice
      for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xFFFFFFFFu, c, o);
      if ((c ^ 0xFFFFFFFFu) != ld_u32(p + 12)) st = 3;
    }
    if (st == 0) {
      const uint8_t *pay = p + kPageHeader;
      if (p[6] != expect_type) {
        st = 4;
      } else if (expect_type == kPtRawVec) {
        if (len < 8) {
          st = 5;
        } else {
          uint32_t vcount = ld_u16(pay), dims = ld_u16(pay + 2), prec = pay[4];
          uint32_t bpe = prec == 0 ? 8u : (prec == 2 ? 1u : 4u);
          if (dims == 0 || (uint64_t)len < 8ull + (uint64_t)vcount * dims * bpe)
            st = 5;
          else if (dims != expect_dims)
            st = 6;
          else if (prec != expect_prec)   // the slot -> node mapping assumes the index's precision
            st = 7;
        }
      } else {
        if (len < 4) {
          st = 5;
        } else {
          uint32_t cnt = ld_u16(pay), deg = ld_u16(pay + 2);
          if (deg == 0 || (uint64_t)len < 4ull + (uint64_t)cnt * (2 + deg * 4)) st = 5;
        }
      }
    }
    if (lane == 0) {
      status[pg] = st;
      if (st) atomicAdd(bad_count, 1u);
    }
  }
}

// rows [first_row, first_row + n_rows) of the shard come from consecutive pages;
// page i holds rows_per_page slots, slot s at payload offset 8 + s*dims*bpe with
// bpe taken from the page's own precision byte (tryDecodePayload :427-447).
// Slots >= vectorCount decode as missing -> zero row (never reached for live ids).
__global__ void page_decode_kernel(const uint8_t *pages, uint32_t page_size, uint32_t rows_per_page,
                                   uint64_t slot_offset, uint64_t n_rows, uint32_t dims,
                                   uint8_t *dst, uint32_t row_bytes, uint32_t ld, int dtype) {
  const uint64_t total = n_rows * ld;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / ld;
    uint32_t c = (uint32_t)(i % ld);
    uint64_t slot_g = slot_offset + r;
    const uint8_t *pay = pages + (slot_g / rows_per_page) * page_size + kPageHeader;
    uint32_t slot = (uint32_t)(slot_g % rows_per_page);
    uint32_t vcount = ld_u16(pay), prec = pay[4];
    uint32_t bpe = prec == 0 ? 8u : (prec == 2 ? 1u : 4u);
    float v = 0.0f;
    if (c < dims && slot < vcount)
      v = decode_src(pay + 8 + ((size_t)slot * dims + c) * bpe, (int)prec);
    store_dev(dst + r * row_bytes, c, v, dtype);
  }
}

// graph pages -> deleted bitmap. Page p of the call holds node ids starting at first_node +
// p * per_page (per_page = NghPageSizer.nodesPerGraphPage, ngh_page.dart:559-566; a page
// may be filled to fewer slots than that); slot layout [flags:u8][degree:u8][maxDegree x u32];
// bit 0x01 = tombstone.
__global__ void graph_flags_kernel(const uint8_t *pages, uint64_t n_pages, uint32_t page_size,
                                   uint32_t per_page, uint64_t first_node, uint64_t shard_first,
                                   uint64_t shard_rows, uint32_t *deleted_bits, uint32_t *n_set) {
  for (uint64_t pg = blockIdx.x; pg < n_pages; pg += gridDim.x) {
    const uint8_t *pay = pages + pg * page_size + kPageHeader;
    uint32_t cnt = ld_u16(pay), deg = ld_u16(pay + 2);
    uint32_t slot = 2 + deg * 4;
    if (cnt > per_page) cnt = per_page;
    for (uint32_t s = threadIdx.x; s < cnt; s += blockDim.x) {
      uint64_t node = first_node + pg * per_page + s;
      if (node < shard_first || node >= shard_first + shard_rows) continue;
      uint64_t r = node - shard_first;
      uint32_t bit = 1u << (r & 31);
      if (pay[4 + (size_t)s * slot] & 0x01u) {
        uint32_t old = atomicOr(&deleted_bits[r >> 5], bit);
        if (!(old & bit)) atomicAdd(n_set, 1u);
      }
    }
  }
}

// ---- liveness bitmaps --------------------------------------------------------
__global__ void set_bits_kernel(const uint64_t *node_ids, uint64_t n, uint64_t shard_first,
                                uint64_t shard_rows, uint32_t *bits, int value, int *delta) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t id = node_ids[i];
    if (id < shard_first || id >= shard_first + shard_rows) continue;
    uint64_t r = id - shard_first;
    uint32_t bit = 1u << (r & 31);
    if (value) {
      uint32_t old = atomicOr(&bits[r >> 5], bit);
      if (!(old & bit)) atomicAdd(delta, 1);
    } else {
      uint32_t old = atomicAnd(&bits[r >> 5], ~bit);
      if (old & bit) atomicAdd(delta, -1);
    }
  }
}

// live = ~deleted & filter (filter may be NULL)
__global__ void combine_live_kernel(const uint32_t *deleted, const uint32_t *filter,
                                    uint32_t *live, uint64_t words, uint64_t n_rows,
                                    unsigned long long *live_count) {
  unsigned long long cnt = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t d = deleted ? deleted[i] : 0u;
    uint32_t f = filter ? filter[i] : 0xFFFFFFFFu;
    uint32_t l = ~d & f;
    live[i] = l;
    uint64_t base = i * 32;
    if (base + 32 > n_rows) l = base < n_rows ? (l & (uint32_t)((1ull << (n_rows - base)) - 1ull)) : 0u;
    cnt += __popc(l);
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(live_count, cnt);
}

// [nq, dims] -> [nq, qld] zero padded
__global__ void pad_queries_kernel(const float *src, uint32_t nq, uint32_t dims, float *dst,
                                   uint32_t qld) {
  uint32_t total = nq * qld;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint32_t q = i / qld, c = i % qld;
    dst[i] = c < dims ? src[(size_t)q * dims + c] : 0.0f;
  }
}

}  // namespace tsc
