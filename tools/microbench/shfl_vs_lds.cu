// Microbenchmark: do warp shuffles contend with shared-memory loads on B200?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256,1) k_shfl(int iters, int* out) {
  int v[16], acc = 0;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = threadIdx.x * 17 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __shfl_sync(0xffffffffu, v[q], (threadIdx.x + q + 1) & 31);  // 16 independent chains
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) acc += v[q];
  if (acc == 123456789) out[0] = acc;
}
__global__ void __launch_bounds__(256,1) k_lds(int iters, int* out) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_double2(i, i);
  __syncthreads();
  double acc = 0; int idx = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) { double2 v = sm[(idx * 5 + q * 32 + it) & 1023]; acc += v.x + v.y; }
  }
  if (acc == 123456789.0) out[0] = 1;
}
__global__ void __launch_bounds__(256,1) k_both(int iters, int* out) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_double2(i, i);
  __syncthreads();
  double acc = 0; int idx = threadIdx.x & 31; int v[16], iacc = 0;
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = threadIdx.x * 17 + q;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double2 d = sm[(idx * 5 + q * 32 + it) & 1023]; acc += d.x + d.y;
      v[q] = __shfl_sync(0xffffffffu, v[q], (threadIdx.x + q + 1) & 31);
    }
  }
#pragma unroll
  for (int q = 0; q < 16; ++q) iacc += v[q];
  if (acc == 123456789.0 || iacc == 123456789) out[0] = 1;
}
int main() {
  int* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000; float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_shfl<<<148, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("shfl only : %.3f ms  -> %.2f clk per warp-SHFL per SM\n", ms, ms * 1e-3 * 1.965e9 / (double(iters) * 16 * 8));
    cudaEventRecord(e0); k_lds<<<148, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("lds.128   : %.3f ms  -> %.2f clk per warp-LDS.128 per SM\n", ms, ms * 1e-3 * 1.965e9 / (double(iters) * 16 * 8));
    cudaEventRecord(e0); k_both<<<148, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("both      : %.3f ms  -> %.2f clk per (LDS.128 + SHFL) pair per SM\n", ms, ms * 1e-3 * 1.965e9 / (double(iters) * 16 * 8));
  }
  return 0;
}
