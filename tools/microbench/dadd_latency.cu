// Dependent-issue latency of the fp64 add on this GPU: the exact re-rank (tsc_tail.cuh) is a
// sequential DADD chain per candidate by definition (Dart's `sum += a*b` loop), so d x this
// latency is the floor of the tail of every search. One warp, one chain per lane.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/dadd_latency.cu -o /tmp/dadd && /tmp/dadd
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(const double *src, double *out, long long *cycles, int n) {
  __shared__ double p[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) p[i] = src[i];
  __syncthreads();
  double s = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < n; r++) {
#pragma unroll 16
    for (int i = 0; i < 1024; i++) {
      if (OP == 0) s = __dadd_rn(s, p[i]);
      else if (OP == 1) s = __fma_rn(s, 1.0, p[i]);
      else s = __dmul_rn(s, p[i]);
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  double h[1024], *d_src, *d_out;
  long long *d_cyc, cyc;
  for (int i = 0; i < 1024; i++) h[i] = 1.0 + 1e-9 * i;
  cudaMalloc(&d_src, sizeof h);
  cudaMalloc(&d_out, 1024 * 8);
  cudaMalloc(&d_cyc, 8);
  cudaMemcpy(d_src, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = 64;
  const char *names[3] = {"DADD", "DFMA(s,1,p)", "DMUL"};
  for (int lanes = 1; lanes <= 32; lanes *= 32) {
    for (int op = 0; op < 3; op++) {
      for (int rep = 0; rep < 2; rep++) {
        if (op == 0) chain<0><<<1, lanes>>>(d_src, d_out, d_cyc, n);
        else if (op == 1) chain<1><<<1, lanes>>>(d_src, d_out, d_cyc, n);
        else chain<2><<<1, lanes>>>(d_src, d_out, d_cyc, n);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-12s %2d lane(s): %.2f cycles per dependent op\n", names[op], lanes,
             (double)cyc / (1024.0 * n));
    }
  }
  return 0;
}
