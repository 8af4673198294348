// Microbenchmark: FP64 tensor-core (mma.sync m8n8k4 f64) throughput on B200, for the record only --
// the shipped kernels use the FP64 FMA pipe (BASELINE north_star: no tensor cores).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_dmma(int iters, double* out) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  if (s == 123.456) out[0] = s;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000; float ms;
  for (int blocks : {148, 148 * 2, 148 * 4, 148 * 8}) {
    k_dmma<<<blocks, 256>>>(100, out);
    cudaEventRecord(e0); k_dmma<<<blocks, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (blocks * 8.0);
    printf("blocks=%d: %.3f ms, %.2f TFLOP/s\n", blocks, ms, flop / (ms * 1e-3) / 1e12);
  }
  return 0;
}
