// Roofline denominators that MEASURED_PEAKS.json does not carry: the FP64 FMA pipe and
// the L2-resident read bandwidth.  Used only by bench.py to compute `roofline.frac`.
#include "common.cuh"

namespace sa {

// 16 independent DFMA chains per thread: 2*16*iters flop per thread, no memory traffic.
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double *__restrict__ out) {
  double a[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) a[q] = seed + q + threadIdx.x * 1e-6;
  const double m = 1.0 - 1e-9, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = fma(a[q], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += a[q];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

// The tile kernel's instruction mix without any memory traffic: an 8x8 register outer-product update
// (64 DFMA with three distinct register operands + 8 DMUL per step).  Its throughput is the practical
// ceiling of the register-tiled design (register-file read bandwidth included).
__global__ void __launch_bounds__(256, 1) fp64_outer_kernel(int iters, double seed, double *__restrict__ out) {
  double acc[8][8], X[8], Y[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    X[a] = seed + a * 1e-3 + threadIdx.x * 1e-6;
    Y[a] = seed - a * 1e-3;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[a][k] = 0.0;
  }
  double w = 1.0 + 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const double wx = w * X[a];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][k] = fma(wx, Y[k], acc[a][k]);
    }
    w += 1e-12;  // keeps the multiplies inside the loop
  }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[a][k];
  if (s == 123.456) out[0] = s;
}

// Every thread block streams the same `bytes` buffer (L2 resident when bytes << 126 MB) `reps` times.
__global__ void __launch_bounds__(256)
l2_read_kernel(const double2 *__restrict__ buf, uint64_t n16, int reps, double *__restrict__ out) {
  double s = 0.0;
  for (int r = 0; r < reps; ++r) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (uint64_t)gridDim.x * blockDim.x) {
      const double2 v = __ldcg(buf + i);
      s += v.x + v.y;
    }
  }
  if (s == 123.456) out[0] = s;
}

}  // namespace sa

using namespace sa;

// Enqueue the FP64 peak kernel; flop = 2 * 16 * iters * threads.  Returns the thread count via *threads.
extern "C" int sa_microbench_fp64(void *stream, int iters, double *d_out, uint64_t *h_threads) {
  SA_REQUIRE(d_out && iters > 0, SA_ERR_ARG, "sa_microbench_fp64: bad arguments");
  const int blocks = sm_count() * 8;
  fp64_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, d_out);
  if (h_threads) *h_threads = (uint64_t)blocks * 256;
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// 8x8 outer-product mix: flop counted = 2 * 64 * iters * threads (the 8 DMUL per step are overhead, as
// in the tile kernel's algorithmic count).  One block of 256 threads per SM, like the tile kernel.
extern "C" int sa_microbench_fp64_outer(void *stream, int iters, double *d_out, uint64_t *h_threads) {
  SA_REQUIRE(d_out && iters > 0, SA_ERR_ARG, "sa_microbench_fp64_outer: bad arguments");
  const int blocks = sm_count();
  fp64_outer_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, d_out);
  if (h_threads) *h_threads = (uint64_t)blocks * 256;
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// Enqueue the L2 read kernel over `bytes` (multiple of 16); total bytes read = bytes * reps.
extern "C" int sa_microbench_l2_read(void *stream, const void *d_buf, uint64_t bytes, int reps, double *d_out) {
  SA_REQUIRE(d_buf && d_out && bytes >= 16 && reps > 0, SA_ERR_ARG, "sa_microbench_l2_read: bad arguments");
  const int blocks = sm_count() * 8;
  l2_read_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const double2 *)d_buf, bytes / 16, reps, d_out);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
