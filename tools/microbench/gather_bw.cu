// gather_bw.cu — what bandwidth does HBM3e deliver for sparse row reads?
//
// The C5 path (WHERE prefilter, 10 % selectivity) reads isolated 1.5 KB rows. This
// microbenchmark measures the ceiling of that access pattern independently of the scan
// kernel's design: every warp reads whole rows with 128-bit loads, rows chosen by a
// Bernoulli(density) bitmap (or contiguous when density = 1), with enough warps in flight
// that only DRAM can be the limit. The number it prints for (row_bytes = 1536, density =
// 0.10) is the honest roofline denominator for C5.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/gather_bw.cu -o /tmp/gather_bw
//   /tmp/gather_bw [rows=12500000] [row_bytes=1536]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

// one warp per live row (list of row ids), 16 bytes per lane per step, evict-first loads
__global__ void gather_rows(const uint4 *__restrict__ base, const uint32_t *__restrict__ ids,
                            uint32_t n_ids, uint32_t chunks_per_row, uint32_t *__restrict__ sink) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t acc = 0;
  for (uint64_t i = warp; i < n_ids; i += warps) {
    const uint4 *row = base + (uint64_t)ids[i] * chunks_per_row;
    for (uint32_t c = lane; c < chunks_per_row; c += 32) {
      uint4 v = __ldcs(row + c);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x9E3779B9u) *sink = acc;   // keep the loads alive
}

int main(int argc, char **argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], nullptr, 10) : 12500000ull;
  const uint32_t row_bytes = argc > 2 ? (uint32_t)atoi(argv[2]) : 1536u;
  const uint32_t cpr = row_bytes / 16;
  uint4 *d = nullptr;
  uint32_t *d_ids = nullptr, *d_sink = nullptr;
  CK(cudaMalloc(&d, rows * row_bytes));
  CK(cudaMemset(d, 1, rows * row_bytes));
  CK(cudaMalloc(&d_ids, rows * 4));
  CK(cudaMalloc(&d_sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  std::printf("rows=%llu row_bytes=%u (%.2f GB column)\n", (unsigned long long)rows, row_bytes,
              rows * (double)row_bytes / 1e9);
  const double densities[] = {1.0, 0.5, 0.25, 0.10, 0.03, 0.01};
  for (double p : densities) {
    std::vector<uint32_t> ids;
    uint64_t s = 0x243F6A8885A308D3ull;
    for (uint64_t r = 0; r < rows; r++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;                 // xorshift64
      if (p >= 1.0 || (double)(s >> 11) * (1.0 / 9007199254740992.0) < p) ids.push_back((uint32_t)r);
    }
    CK(cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
    for (int ctas_per_sm : {8, 16}) {
      float best = 1e30f;
      for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        gather_rows<<<sms * ctas_per_sm, 128>>>(d, d_ids, (uint32_t)ids.size(), cpr, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < best) best = ms;
      }
      std::printf("density %.2f  live rows %zu  %d CTAs/SM x 4 warps: %.3f ms  %.0f GB/s of live-row bytes\n",
                  p, ids.size(), ctas_per_sm, best, ids.size() * (double)row_bytes / best / 1e6);
    }
  }
  return 0;
}
