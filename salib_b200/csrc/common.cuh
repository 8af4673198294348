// Shared helpers for the salib_b200 CUDA library (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/salib_b200.h"

namespace sa {

void set_error(const char *fmt, ...);
// Bookkeeping for bench.py's "gpu_launches": every entry point reports the kernels it enqueued.
void note_launches(int n);

#define SA_REQUIRE(cond, code, ...)    \
  do {                                 \
    if (!(cond)) {                     \
      ::sa::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

#define SA_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::sa::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                      cudaGetErrorString(e__));                                        \
      return (int)e__;                                                                 \
    }                                                                                  \
  } while (0)

#define SA_LAUNCH_CHECK() SA_CUDA(cudaGetLastError())

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Number of SMs of the current device (cached per process; 148 on B200).
int sm_count();

__host__ __device__ inline uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// n / d for n*d' < 2^32 style bounded operands: q = umulhi(n, magic), magic = 2^32/d + 1.
// magic == 0 selects the plain division (host decides, see fastdiv_magic()).
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t d, uint32_t magic) {
  return magic ? __umulhi(n, magic) : n / d;
}
inline uint32_t fastdiv_magic(uint64_t n_max, uint32_t d) {
  // exact while n * (magic*d - 2^32) < 2^32; magic*d - 2^32 <= d
  if (d <= 1) return 0;
  if (n_max * (uint64_t)d >= (1ull << 32)) return 0;
  return (uint32_t)((1ull << 32) / d + 1);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

}  // namespace sa
