// Microbenchmark: do DMMA (FP64 tensor) and DFMA (FP64 FMA pipe) share the same datapath on B200?
// Half of the warps of every block run DFMA chains, the other half DMMA; compare with each alone.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(int iters, int mode, double* out) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = (mode == 1) || (mode == 2 && (warp & 1));
  const bool do_fma = (mode == 0) || (mode == 2 && !(warp & 1));
  double s = 0;
  if (do_mma) {
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int t = 0; t < 8; ++t)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  }
  if (do_fma) {
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = 1.0 + q + threadIdx.x * 1e-6;
    const double m = 1.0 - 1e-9, cc = 1e-9;
    for (int it = 0; it < iters * 4; ++it)   // 16*4 = 64 DFMA per lane per outer iteration = 8 DMMA worth of FMAs
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] = fma(a[q], m, cc);
#pragma unroll
    for (int q = 0; q < 16; ++q) s += a[q];
  }
  if (s == 123.456) out[0] = s;
}
int main() {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000; float ms;
  const char* names[3] = {"DFMA only (all warps)", "DMMA only (all warps)", "half DFMA / half DMMA"};
  for (int mode = 0; mode < 3; ++mode) {
    k<<<148 * 4, 256>>>(100, mode, out);
    cudaEventRecord(e0); k<<<148 * 4, 256>>>(iters, mode, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    // every warp does 8*256 FMA-equivalents per iteration in all modes
    double flop = 2.0 * 8 * 256 * double(iters) * (148 * 4 * 8.0);
    printf("%-24s %.3f ms, %.2f TFLOP/s\n", names[mode], ms, flop / (ms * 1e-3) / 1e12);
  }
  return 0;
}
