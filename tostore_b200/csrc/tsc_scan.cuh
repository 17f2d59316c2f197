// tsc_scan.cuh — K1: HBM-bound exact scan with in-kernel top-K' selection.
//
// Replaces the candidate-generation half of NghGraphEngine.search
// (core/ngh_graph_engine.dart:98-113: ADC beam search) with an exhaustive pass
// over the row-major embedding block. The fp32 ranking key computed here plays
// the role the PQ/ADC distance plays in the reference: it only has to put the
// true top-k inside the K' = max(2k, 20) candidates (ngh_graph_engine.dart:115)
// that tsc_select.cu then re-ranks with the reference's exact fp64 arithmetic.
//
// Data movement: every warp owns a private ring of S stages in shared memory and
// keeps it full with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA
// unit) of R consecutive rows each, completion on a per-stage mbarrier. No
// CTA-wide barrier exists in the main loop. Lanes read 16-byte chunks of the
// staged rows (conflict-free LDS.128), accumulate in fp32, butterfly-reduce with
// warp shuffles, and insert into a per-warp sorted candidate list in shared memory.
#pragma once

#include "tsc_common.cuh"

namespace tsc {

struct ScanParams {
  const uint8_t *rows;       // [n_rows, row_bytes] device storage dtype, 16B aligned
  uint64_t n_rows;
  uint32_t row_bytes;        // row stride in bytes (multiple of 16)
  uint32_t chunks_per_row;   // row_bytes / 16
  const float *queries;      // [QB, qld] fp32, zero padded to qld
  uint32_t qld;              // padded dims = chunks_per_row * elems_per_chunk
  uint32_t nq;               // valid queries in this launch (<= QB)
  const uint32_t *live_mask; // optional: bit r = row r may be returned; NULL = all live
  uint32_t kprime;           // candidates kept per list
  uint32_t stages;           // S
  uint32_t stage_bytes;      // R * row_bytes
  uint64_t *cand;            // [QB][gridDim.x][kprime] (ordered key << 32 | shard row)
  uint32_t sort_cap;         // pow2 >= warps * kprime (block-level merge buffer)
};

// Insert (key,id) into a sorted (ascending) list of kp entries held in shared
// memory by one warp; the last entry falls off. All lanes pass identical
// arguments. Rows reach a warp in increasing id order, so an equal key is placed
// after the existing ones and (key,id) stays lexicographically sorted.
// Returns the new threshold (largest kept key).
__device__ __forceinline__ uint32_t list_insert(uint32_t *keys, uint32_t *ids, int kp,
                                                uint32_t key, uint32_t id, int lane) {
  int pos = 0;
  for (int base = 0; base < kp; base += 32) {
    int j = base + lane;
    bool le = (j < kp) && (keys[j] <= key);
    unsigned m = __ballot_sync(0xFFFFFFFFu, le);
    pos += __popc(m);
    if (m != 0xFFFFFFFFu) break;
  }
  for (int base = ((kp - 1) >> 5) << 5; base >= 0 && base + 31 >= pos; base -= 32) {
    int j = base + lane;
    bool mv = (j > pos) && (j < kp);
    uint32_t k1 = 0, i1 = 0;
    if (mv) {
      k1 = keys[j - 1];
      i1 = ids[j - 1];
    }
    __syncwarp();
    if (mv) {
      keys[j] = k1;
      ids[j] = i1;
    } else if (j == pos) {
      keys[j] = key;
      ids[j] = id;
    }
  }
  __syncwarp();
  return keys[kp - 1];
}

template <int DTYPE>
struct Chunk;  // 16 bytes of a stored row -> fp32 lanes

template <>
struct Chunk<kF32> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
};
template <>
struct Chunk<kBF16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[8]) {
    unpack_bf16x2(v.x, f[0], f[1]);
    unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]);
    unpack_bf16x2(v.w, f[6], f[7]);
  }
};
template <>
struct Chunk<kF16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &v, float (&f)[8]) {
    unpack_f16x2(v.x, f[0], f[1]);
    unpack_f16x2(v.y, f[2], f[3]);
    unpack_f16x2(v.z, f[4], f[5]);
    unpack_f16x2(v.w, f[6], f[7]);
  }
};

// Shared-memory footprint (must match the carve-up in the kernel).
__host__ __device__ inline size_t scan_smem_query_bytes(int qb, uint32_t qld) {
  return ((size_t)qb * qld * 4 + 127) & ~(size_t)127;
}
__host__ __device__ inline size_t scan_smem_sort_bytes(uint32_t sort_cap) {
  return ((size_t)sort_cap * 8 + 127) & ~(size_t)127;
}
__host__ __device__ inline size_t scan_smem_warp_bytes(int qb, uint32_t kprime, uint32_t stages,
                                                      uint32_t stage_bytes) {
  size_t lists = ((size_t)qb * kprime * 8 + 15) & ~(size_t)15;
  size_t bars = ((size_t)stages * 8 + 15) & ~(size_t)15;
  size_t ring = (size_t)stages * stage_bytes;
  return (lists + bars + ring + 127) & ~(size_t)127;
}

// ---- block-level merge: W sorted lists -> one list of kp per query -------------
// (bitonic sort of the composites in shared memory; every bulk copy this CTA
// issued has been waited on, so no async write is outstanding)
__device__ __forceinline__ void scan_block_merge(const ScanParams &p, uint64_t *sortbuf,
                                                 const uint32_t *lkeys, const uint32_t *lids,
                                                 uint32_t kp, int warp, int warps, int lane) {
  for (uint32_t q = 0; q < p.nq; q++) {
    __syncthreads();
    for (uint32_t i = lane; i < kp; i += 32)
      sortbuf[(size_t)warp * kp + i] =
          ((uint64_t)lkeys[(size_t)q * kp + i] << 32) | lids[(size_t)q * kp + i];
    for (uint32_t i = (uint32_t)warps * kp + threadIdx.x; i < p.sort_cap; i += blockDim.x)
      sortbuf[i] = ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= p.sort_cap; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = threadIdx.x; i < p.sort_cap; i += blockDim.x) {
          uint32_t x = i ^ j;
          if (x > i) {
            uint64_t a = sortbuf[i], b = sortbuf[x];
            bool up = (i & k) == 0;
            if ((a > b) == up) {
              sortbuf[i] = b;
              sortbuf[x] = a;
            }
          }
        }
        __syncthreads();
      }
    }
    uint64_t *out = p.cand + ((size_t)q * gridDim.x + blockIdx.x) * kp;
    for (uint32_t i = threadIdx.x; i < kp; i += blockDim.x) out[i] = sortbuf[i];
  }
}

template <int METRIC, int DTYPE, int QB, int R>
__global__ void __launch_bounds__(512, 1) scan_topk_kernel(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int E = Chunk<DTYPE>::kElems;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const uint32_t kp = p.kprime;
  const uint32_t S = p.stages;

  // ---- carve shared memory -------------------------------------------------
  float *qs = reinterpret_cast<float *>(smem);
  uint64_t *sortbuf = reinterpret_cast<uint64_t *>(smem + scan_smem_query_bytes(QB, p.qld));
  uint8_t *wbase = smem + scan_smem_query_bytes(QB, p.qld) + scan_smem_sort_bytes(p.sort_cap) +
                   (size_t)warp * scan_smem_warp_bytes(QB, kp, S, p.stage_bytes);
  uint32_t *lkeys = reinterpret_cast<uint32_t *>(wbase);            // [QB][kp]
  uint32_t *lids = lkeys + (size_t)QB * kp;                         // [QB][kp]
  size_t off = ((size_t)QB * kp * 8 + 15) & ~(size_t)15;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + off);
  off += ((size_t)S * 8 + 15) & ~(size_t)15;
  off = (off + 127) & ~(size_t)127;
  uint8_t *ring = wbase + off;

  // ---- queries -> smem (zero rows for q >= nq), lists -> empty --------------
  for (uint32_t i = threadIdx.x; i < (uint32_t)QB * p.qld; i += blockDim.x) {
    uint32_t q = i / p.qld;
    qs[i] = (q < p.nq) ? p.queries[i] : 0.0f;
  }
  for (uint32_t i = lane; i < (uint32_t)QB * kp; i += 32) {
    lkeys[i] = kEmptyKey;
    lids[i] = kInvalidRow;
  }
  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t policy = policy_evict_first();
  const uint64_t total_stages = (p.n_rows + R - 1) / R;
  const uint64_t gw = (uint64_t)blockIdx.x * warps + warp;
  const uint64_t GW = (uint64_t)gridDim.x * warps;

  auto issue = [&](uint32_t s, uint64_t st) {
    uint64_t row0 = st * R;
    uint64_t nrow = p.n_rows - row0;
    uint32_t bytes = (nrow >= (uint64_t)R) ? p.stage_bytes : (uint32_t)nrow * p.row_bytes;
    uint32_t bar = smem_u32(&bars[s]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(ring + (size_t)s * p.stage_bytes), p.rows + row0 * p.row_bytes, bytes, bar,
             policy);
  };

  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) {
      uint64_t st = gw + (uint64_t)s * GW;
      if (st < total_stages) issue(s, st);
    }
  }

  uint32_t thr[QB];
#pragma unroll
  for (int q = 0; q < QB; q++) thr[q] = kEmptyKey;

  uint32_t s = 0, parity = 0;
  const uint32_t cpr = p.chunks_per_row;
  for (uint64_t st = gw; st < total_stages; st += GW) {
    mbar_wait(smem_u32(&bars[s]), parity);

    float acc[QB][R];
    float bb[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      bb[r] = 0.0f;
#pragma unroll
      for (int q = 0; q < QB; q++) acc[q][r] = 0.0f;
    }
    const uint4 *stage = reinterpret_cast<const uint4 *>(ring + (size_t)s * p.stage_bytes);
#pragma unroll 2
    for (uint32_t c = lane; c < cpr; c += 32) {
      float b[R][E];
#pragma unroll
      for (int r = 0; r < R; r++) {
        uint4 v = stage[(uint32_t)r * cpr + c];
        Chunk<DTYPE>::unpack(v, b[r]);
      }
      if (METRIC == kCos) {
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int e = 0; e < E; e++) bb[r] = fmaf(b[r][e], b[r][e], bb[r]);
      }
#pragma unroll
      for (int q = 0; q < QB; q++) {
        float a[E];
        const float4 *qp = reinterpret_cast<const float4 *>(qs + (size_t)q * p.qld + (size_t)c * E);
#pragma unroll
        for (int h = 0; h < E / 4; h++) {
          float4 t = qp[h];
          a[4 * h + 0] = t.x;
          a[4 * h + 1] = t.y;
          a[4 * h + 2] = t.z;
          a[4 * h + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int e = 0; e < E; e++) {
            if (METRIC == kL2) {
              float d = a[e] - b[r][e];
              acc[q][r] = fmaf(d, d, acc[q][r]);
            } else {
              acc[q][r] = fmaf(a[e], b[r][e], acc[q][r]);
            }
          }
      }
    }
    __syncwarp();  // every lane is done reading this stage
    {
      uint64_t nst = st + (uint64_t)S * GW;
      if (lane == 0 && nst < total_stages) issue(s, nst);
    }

    // ---- butterfly reduction: every lane ends with the full sums -----------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (METRIC == kCos) bb[r] += __shfl_xor_sync(0xFFFFFFFFu, bb[r], o);
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q][r] += __shfl_xor_sync(0xFFFFFFFFu, acc[q][r], o);
      }
    }

    // ---- candidate insertion (warp-uniform control flow) --------------------
    const uint64_t row0 = st * R;
    uint32_t live = 0xFFFFFFFFu;
    if (p.live_mask != nullptr) {
      // R divides 32 and row0 is a multiple of R: the stage's bits sit in one word
      live = p.live_mask[row0 >> 5] >> (row0 & 31);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint64_t row = row0 + r;
      if (row >= p.n_rows || !((live >> r) & 1u)) continue;
#pragma unroll
      for (int q = 0; q < QB; q++) {
        float key;
        if (METRIC == kL2) {
          key = acc[q][r];
        } else if (METRIC == kIP) {
          key = -acc[q][r];
        } else {
          key = (bb[r] > 0.0f) ? -acc[q][r] * rsqrtf(bb[r]) : 0.0f;
        }
        key += 0.0f;  // -0.0 -> +0.0 so exact ties order by id only
        uint32_t uk = ordered_key(key);
        if (uk < thr[q] && (uint32_t)q < p.nq)
          thr[q] = list_insert(lkeys + (size_t)q * kp, lids + (size_t)q * kp, (int)kp, uk,
                               (uint32_t)row, lane);
      }
    }

    if (++s == S) {
      s = 0;
      parity ^= 1u;
    }
  }

  scan_block_merge(p, sortbuf, lkeys, lids, kp, warp, warps, lane);
}

// K6: scan of a sparsely live column (WHERE prefilter / heavy tombstoning).
// Same arithmetic and candidate lists as scan_topk_kernel, but the unit of data
// movement is ONE LIVE ROW: a warp walks its share of the liveness bitmap (one
// 32-bit word = 32 consecutive rows at a time) and issues a bulk copy only for rows
// whose bit is set, so dead rows cost no HBM bytes (rows are whole 16-byte-aligned
// byte ranges; at d=384 fp32 a row is exactly twelve 128-byte lines). The ring is a
// queue of S one-row stages; rows are consumed in issue (= increasing id) order.
//
// PF = false: a warp takes single bitmap words (32 rows) round-robin and loads each word
// when the previous one is exhausted — a dependent global load on the issue path.
// PF = true (opt-in, TSC_SCAN_SPARSE_PF=1; written after the round's GPU budget was spent,
// not yet measured): a warp takes BLOCKS of 32 consecutive words (1024 rows) round-robin;
// lane l holds word l of the block (one coalesced 128-byte load), the next block is
// prefetched while the current one is consumed, and all-zero words are skipped with a
// ballot — no load latency on the issue path and cheap skipping at low selectivity.
template <int METRIC, int DTYPE, int QB, bool PF = false>
__global__ void __launch_bounds__(512, 1) scan_topk_sparse_kernel(const ScanParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int E = Chunk<DTYPE>::kElems;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const uint32_t kp = p.kprime;
  const uint32_t S = p.stages;

  float *qs = reinterpret_cast<float *>(smem);
  uint64_t *sortbuf = reinterpret_cast<uint64_t *>(smem + scan_smem_query_bytes(QB, p.qld));
  uint8_t *wbase = smem + scan_smem_query_bytes(QB, p.qld) + scan_smem_sort_bytes(p.sort_cap) +
                   (size_t)warp * scan_smem_warp_bytes(QB, kp, S, p.stage_bytes);
  uint32_t *lkeys = reinterpret_cast<uint32_t *>(wbase);
  uint32_t *lids = lkeys + (size_t)QB * kp;
  size_t off = ((size_t)QB * kp * 8 + 15) & ~(size_t)15;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + off);
  off += ((size_t)S * 8 + 15) & ~(size_t)15;
  off = (off + 127) & ~(size_t)127;
  uint8_t *ring = wbase + off;

  for (uint32_t i = threadIdx.x; i < (uint32_t)QB * p.qld; i += blockDim.x) {
    uint32_t q = i / p.qld;
    qs[i] = (q < p.nq) ? p.queries[i] : 0.0f;
  }
  for (uint32_t i = lane; i < (uint32_t)QB * kp; i += 32) {
    lkeys[i] = kEmptyKey;
    lids[i] = kInvalidRow;
  }
  if (lane == 0) {
    for (uint32_t s = 0; s < S; s++) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t policy = policy_evict_first();
  const uint64_t n_words = (p.n_rows + 31) / 32;
  const uint64_t gw = (uint64_t)blockIdx.x * warps + warp;
  const uint64_t GW = (uint64_t)gridDim.x * warps;
  const uint32_t cpr = p.chunks_per_row;

  uint32_t thr[QB];
#pragma unroll
  for (int q = 0; q < QB; q++) thr[q] = kEmptyKey;

  // issue side: (word, remaining bits) cursor over this warp's bitmap words
  uint64_t word = gw;
  uint32_t bits = 0;
  auto load_word = [&]() {
    bits = 0;
    while (word < n_words) {
      bits = p.live_mask[word];
      uint64_t base_row = word * 32;
      if (base_row + 32 > p.n_rows) bits &= (uint32_t)((1ull << (p.n_rows - base_row)) - 1ull);
      if (bits) break;
      word += GW;
    }
  };
  // PF: block cursor. `cur` / `nxt` = this lane's word of the current / next block,
  // `nz` = lanes of the current block whose word is non-zero and not yet consumed.
  const uint64_t n_blocks = (n_words + 31) / 32;
  uint64_t blk = gw;
  uint32_t cur = 0, nxt = 0;
  unsigned nz = 0;
  auto load_block = [&](uint64_t b) -> uint32_t {
    const uint64_t w = b * 32 + (uint64_t)lane;
    uint32_t v = 0;
    if (b < n_blocks && w < n_words) {
      v = p.live_mask[w];
      const uint64_t base_row = w * 32;
      if (base_row + 32 > p.n_rows) v &= (uint32_t)((1ull << (p.n_rows - base_row)) - 1ull);
    }
    return v;
  };
  auto next_word = [&]() {   // warp-uniform: advance to the next non-zero word of this warp
    bits = 0;
    for (;;) {
      if (nz) {
        const int l = __ffs(nz) - 1;
        nz &= nz - 1;
        bits = __shfl_sync(0xFFFFFFFFu, cur, l);
        word = blk * 32 + (uint64_t)l;
        return;
      }
      blk += GW;
      if (blk >= n_blocks) return;
      cur = nxt;
      nxt = load_block(blk + GW);
      nz = __ballot_sync(0xFFFFFFFFu, cur != 0);
    }
  };
  if (PF) {
    cur = load_block(blk);
    nxt = load_block(blk + GW);
    nz = __ballot_sync(0xFFFFFFFFu, cur != 0);
    if (blk < n_blocks) next_word();
  } else {
    load_word();
  }
  uint32_t head = 0, tail = 0, inflight = 0, hpar = 0;
  // row ids of the stages in flight (uniform per warp): kept in registers of lane s
  uint32_t my_row = kInvalidRow;

  for (;;) {
    while (inflight < S && bits) {
      const uint32_t b = __ffs(bits) - 1;
      bits &= bits - 1;
      const uint32_t row = (uint32_t)(word * 32 + b);
      if (lane == (int)tail) my_row = row;
      if (lane == 0) {
        const uint32_t bar = smem_u32(&bars[tail]);
        mbar_expect_tx(bar, p.row_bytes);
        bulk_g2s(smem_u32(ring + (size_t)tail * p.stage_bytes), p.rows + (size_t)row * p.row_bytes,
                 p.row_bytes, bar, policy);
      }
      if (++tail == S) tail = 0;
      inflight++;
      if (!bits) {
        if (PF) {
          next_word();
        } else {
          word += GW;
          load_word();
        }
      }
    }
    if (inflight == 0) break;

    mbar_wait(smem_u32(&bars[head]), hpar);
    const uint32_t row = __shfl_sync(0xFFFFFFFFu, my_row, head);
    float acc[QB];
    float bb = 0.0f;
#pragma unroll
    for (int q = 0; q < QB; q++) acc[q] = 0.0f;
    const uint4 *stage = reinterpret_cast<const uint4 *>(ring + (size_t)head * p.stage_bytes);
#pragma unroll 2
    for (uint32_t c = lane; c < cpr; c += 32) {
      float b[E];
      uint4 v = stage[c];
      Chunk<DTYPE>::unpack(v, b);
      if (METRIC == kCos) {
#pragma unroll
        for (int e = 0; e < E; e++) bb = fmaf(b[e], b[e], bb);
      }
#pragma unroll
      for (int q = 0; q < QB; q++) {
        const float4 *qp = reinterpret_cast<const float4 *>(qs + (size_t)q * p.qld + (size_t)c * E);
#pragma unroll
        for (int h = 0; h < E / 4; h++) {
          float4 t = qp[h];
          const float a[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (METRIC == kL2) {
              float d = a[e] - b[4 * h + e];
              acc[q] = fmaf(d, d, acc[q]);
            } else {
              acc[q] = fmaf(a[e], b[4 * h + e], acc[q]);
            }
          }
        }
      }
    }
    __syncwarp();  // stage `head` is free again
    if (++head == S) {
      head = 0;
      hpar ^= 1u;
    }
    inflight--;

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (METRIC == kCos) bb += __shfl_xor_sync(0xFFFFFFFFu, bb, o);
#pragma unroll
      for (int q = 0; q < QB; q++) acc[q] += __shfl_xor_sync(0xFFFFFFFFu, acc[q], o);
    }
#pragma unroll
    for (int q = 0; q < QB; q++) {
      float key;
      if (METRIC == kL2) {
        key = acc[q];
      } else if (METRIC == kIP) {
        key = -acc[q];
      } else {
        key = (bb > 0.0f) ? -acc[q] * rsqrtf(bb) : 0.0f;
      }
      key += 0.0f;
      uint32_t uk = ordered_key(key);
      if (uk < thr[q] && (uint32_t)q < p.nq)
        thr[q] = list_insert(lkeys + (size_t)q * kp, lids + (size_t)q * kp, (int)kp, uk, row, lane);
    }
  }
  __syncwarp();
  scan_block_merge(p, sortbuf, lkeys, lids, kp, warp, warps, lane);
}

}  // namespace tsc
