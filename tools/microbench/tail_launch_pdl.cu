// tail_launch_pdl.cu — can a programmatic dependent launch (PDL) overlap a parent grid's tail
// while the parent may still spawn a child grid with a device-side tail launch, and does the
// dependent's griddepcontrol.wait also wait for that child?
//
// Models the pipelined search: A = scan of query i (last CTA runs a tail and, when the
// certificate fails, must re-scan: child grid C), B = scan of query i+1 launched with
// programmatic stream serialization. Wanted: B starts while A's tail runs; B's wait returns
// only after C has finished; without a child nothing is added.
//
//   nvcc -O3 -rdc=true -gencode arch=compute_100a,code=sm_100a tools/microbench/tail_launch_pdl.cu \
//        -o /tmp/tail_launch_pdl -lcudadevrt && /tmp/tail_launch_pdl
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ void spin_us(unsigned us) {
  const unsigned long long t0 = now();
  while (now() - t0 < (unsigned long long)us * 1000ull) {}
}

struct Trace {
  unsigned long long a_start, a_tail_begin, a_tail_end, c_start, c_end, b_start, b_wait_done;
  unsigned int child_value_seen_by_b, ticket;
};

__global__ void child_kernel(Trace *t, unsigned int *out) {
  extern __shared__ unsigned char smem[];
  if (blockIdx.x == 0 && threadIdx.x == 0) t->c_start = now();
  spin_us(100);
  if (threadIdx.x == 0) atomicAdd(out, 1u);
  if (blockIdx.x == 0 && threadIdx.x == 0) t->c_end = now();
  smem[threadIdx.x] = 0;
}

__global__ void parent_kernel(Trace *t, unsigned int *out, int need_child, unsigned smem_bytes) {
  extern __shared__ unsigned char smem[];
  __shared__ unsigned s_ticket;
  if (blockIdx.x == 0 && threadIdx.x == 0) t->a_start = now();
  spin_us(200);
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&t->ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  if (threadIdx.x == 0) t->a_tail_begin = now();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  spin_us(40);   // the tail
  if (threadIdx.x == 0) {
    if (need_child) child_kernel<<<gridDim.x, blockDim.x, smem_bytes, cudaStreamTailLaunch>>>(t, out);
    t->a_tail_end = now();
    t->ticket = 0;
  }
  smem[threadIdx.x] = 0;
}

__global__ void next_kernel(Trace *t, unsigned int *out) {
  extern __shared__ unsigned char smem[];
  if (blockIdx.x == 0 && threadIdx.x == 0) t->b_start = now();
  spin_us(100);   // the main loop of the next scan (short: ends before A's tail + child do)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    t->b_wait_done = now();
    t->child_value_seen_by_b = *reinterpret_cast<volatile unsigned int *>(out);
  }
  smem[threadIdx.x] = 0;
}

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;          // 0 / 1: run only that need_child case
  const unsigned child_smem = argc > 2 ? (unsigned)atoi(argv[2]) : 100u * 1024u;
  Trace *t;
  unsigned int *out;
  CK(cudaMallocManaged(&t, sizeof(Trace)));
  CK(cudaMalloc(&out, 4));
  const unsigned smem = 100 * 1024;
  CK(cudaFuncSetAttribute(parent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(child_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(next_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(parent_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(child_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(next_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int need_child = 0; need_child < 2; need_child++) {
    if (only >= 0 && need_child != only) continue;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaMemset(t, 0, sizeof(Trace)));
      CK(cudaMemset(out, 0, 4));
      CK(cudaDeviceSynchronize());
      parent_kernel<<<148, 256, smem, st>>>(t, out, need_child, child_smem);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, next_kernel, t, out));
      CK(cudaStreamSynchronize(st));
      const double us = 1e-3;
      std::printf("need_child=%d rep=%d | A tail %.1f..%.1f us | C %.1f..%.1f | B start %.1f wait done %.1f | "
                  "B saw child value %u (want %u) | overlap %s, wait covers child %s\n",
                  need_child, rep, (t->a_tail_begin - t->a_start) * us, (t->a_tail_end - t->a_start) * us,
                  t->c_start ? (t->c_start - t->a_start) * us : 0.0, t->c_end ? (t->c_end - t->a_start) * us : 0.0,
                  (t->b_start - t->a_start) * us, (t->b_wait_done - t->a_start) * us,
                  t->child_value_seen_by_b, need_child ? 148u : 0u,
                  t->b_start < t->a_tail_end ? "YES" : "no",
                  (!need_child || t->child_value_seen_by_b == 148u) ? "YES" : "NO");
      std::fflush(stdout);
    }
  }
  return 0;
}
