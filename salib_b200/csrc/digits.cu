// Opt-in "digits" engine of sobol.analyze: the bootstrap sums as an EXACT integer GEMM on the
// 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands by TMA).
//
// Reference behaviour replaced: the gathers + reductions of src/SALib/analyze/sobol.py:209-246 for all
// resamples at once -- the same sums sa_sobol_accumulate_f64 produces (analyze.cu), by another route.
//
// Why this is not "reshaping a gather into a GEMM": with multiplicities the sums ARE
//     psum[c][q] = sum_i w[c][i] * F[i][q]          w: small non-negative integers (uint8), F: float64
// and the integer operand makes an error-free split possible.  Every feature column is scaled by a power of
// two (from its exact max |F[:, q]|) and rounded once to a (8*S-2)-bit integer t, which is written as S
// balanced radix-256 digits  t = sum_s d_s * 256^s,  d_s in [-128, 127]  (int8).  Then
//     sum_i w_i * t_i = sum_s 256^s * (sum_i w_i * d_{s,i})
// and each inner sum is a u8 x s8 dot product accumulated in int32 WITHOUT ANY ROUNDING: |sum| <= 128 *
// sum_i w_i = 128 * N < 2^31 for N <= 2^24.  The digit sums are recombined in 128-bit integers and rounded
// to float64 once.  The only error is the one rounding of each feature value to 8*S-2 bits relative to its
// column maximum (S = 6: 2^-47 of the column max per element, unbiased) -- the summation itself is exact,
// which FP64 accumulation is not.
//
// Kernels
//   feature_table_kernel / feature_max_kernel / feature_scale_kernel / digit_split_kernel
//       Y -> Q int8, K-major, column q*S+s = digit s of feature q, stored K-blocked [Kp/128][nfeat*S][128]:
//       128 consecutive rows of Y of all columns are contiguous (streaming writes, contiguous TMA boxes)
//   digit_gemm_kernel     C[ks][count][ldc] int32 = W[count][K] (u8) * Q[ncols][K]^T (s8) per K split
//       warp 0: TMA producer (128-byte-swizzled 128 x 128 B and BN x 128 B boxes, STAGES-deep mbarrier ring)
//       warp 1: allocates TMEM, one lane issues tcgen05.mma.cta_group::1.kind::i8 (M=128, N=BN, K=32)
//       warps 2-5: epilogue, tcgen05.ld 32x32b -> int32 rows to global
//   digit_combine_kernel  psum[count][nfeat] float64 from the digit sums
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "counts.cuh"

namespace sa {
namespace dg {

constexpr int BM = 128;        // resamples per tile = UMMA M (TMEM lanes)
constexpr int BK = 128;        // rows of Y per pipeline stage = bytes per swizzle row
constexpr int UK = 32;         // K of one tcgen05.mma kind::i8
constexpr int kThreads = 192;  // 6 warps
constexpr int kMaxDigits = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// A barrier that never completes would hang the GPU: give up after 20 s (wall clock, %globaltimer) and trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin) {
    if ((spin & 1023u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}
// The same wait for warps that sit out a whole tile (the epilogue warps wait ~6 ms for the accumulators): sleep
// between polls.  Warps that spin take issue slots from younger warps on the SM -- measured with the slim
// multiplicity kernel running beside the GEMM: 5x slower when the GEMM's CTAs were resident first
// (tools/exp_overlap.py).  A microsecond of wake-up latency per tile is nothing here.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin) {
    __nanosleep(1000);
    if ((spin & 1023u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1, int32_t c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Arrives on `bar` once every tcgen05.mma this thread has issued so far has completed.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x s8 -> s32, shapes from the instruction descriptor.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major operand whose rows are 128 bytes,
// 128-byte swizzle, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4)  // start address, 16-byte units          bits [0,14)
         | (1ull << 16)                      // leading byte offset (unused, K-major)  bits [16,30)
         | ((uint64_t)(1024 >> 4) << 32)     // stride byte offset: 8 rows * 128 B     bits [32,46)
         | (1ull << 46)                      // descriptor version (sm_100)            bits [46,48)
         | (2ull << 61);                     // layout: SWIZZLE_128B                   bits [61,64)
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: D = s32, A = u8, B = s8, both K-major.
__host__ __device__ constexpr uint32_t instr_desc_i8(int M, int N) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
digit_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmQ,
                  int32_t *__restrict__ C, uint32_t count, uint32_t ldc, uint32_t tiles_m, uint32_t tiles_n,
                  uint32_t ksteps, uint32_t ksteps_per_split, uint32_t band, uint32_t q_blocked) {
  // q_blocked: Q is stored [K / 128][ncols][128] (tmQ is 3-D) instead of [ncols][pitch]
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_base_sh;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  constexpr uint32_t kABytes = BM * BK, kBBytes = BN * BK;
  auto sA = [&](uint32_t s) { return tiles_base + s * kABytes; };
  auto sB = [&](uint32_t s) { return tiles_base + STAGES * kABytes + s * kBBytes; };
  auto full = [&](uint32_t s) { return smem_u32(&bars[s]); };
  auto empty = [&](uint32_t s) { return smem_u32(&bars[STAGES + s]); };
  const uint32_t tmem_full = smem_u32(&bars[2 * STAGES]);

  // tile of this block: K split outermost, then bands of `band` m-tiles swept over all n-tiles, so the
  // blocks resident at the same time share their W rows and Q columns in L2
  const uint32_t per_split = tiles_m * tiles_n;
  const uint32_t ks = blockIdx.x / per_split;
  const uint32_t t = blockIdx.x - ks * per_split;
  const uint32_t b = t / (band * tiles_n);
  const uint32_t in_band = t - b * band * tiles_n;
  const uint32_t bm = min(band, tiles_m - b * band);
  const uint32_t tile_m = b * band + in_band % bm, tile_n = in_band / bm;
  const uint32_t k0 = ks * ksteps_per_split;
  const uint32_t nk = min(ksteps_per_split, ksteps - k0);

  if (tid == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {  // TMA producer
      for (uint32_t i = 0; i < nk; ++i) {
        const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        mbar_arrive_expect_tx(full(s), kABytes + kBBytes);
        const int32_t kbyte = (int32_t)((k0 + i) * BK);
        tma_load_2d(sA(s), &tmW, kbyte, (int32_t)(tile_m * BM), full(s));
        if (q_blocked) tma_load_3d(sB(s), &tmQ, 0, (int32_t)(tile_n * BN), (int32_t)(k0 + i), full(s));
        else tma_load_2d(sB(s), &tmQ, kbyte, (int32_t)(tile_n * BN), full(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {  // MMA issuer
      constexpr uint32_t idesc = instr_desc_i8(BM, BN);
      for (uint32_t i = 0; i < nk; ++i) {
        const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(full(s), ph);
        tc_fence_after();
        const uint64_t ad = smem_desc_sw128(sA(s)), bd = smem_desc_sw128(sB(s));
#pragma unroll
        for (uint32_t k = 0; k < BK / UK; ++k)  // +32 bytes along K inside the swizzle row: +2 in 16-byte units
          umma_i8(tmem, ad + 2 * k, bd + 2 * k, idesc, (i | k) != 0);
        tc_commit(empty(s));  // frees the stage when these MMAs have read it
      }
      tc_commit(tmem_full);
    }
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; lane = resample row of the tile
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t quarter = warp & 3;
    const uint32_t row = tile_m * BM + quarter * 32 + lane;
    int32_t *crow = C + ((uint64_t)ks * count + row) * ldc + (uint64_t)tile_n * BN;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < BN; c0 += 32) {
      if (tile_n * BN + c0 >= ldc) break;  // warp-uniform
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((quarter * 32u) << 16) + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < count) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4 *>(crow + c0 + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)BN) : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant: tcgen05.mma.cta_group::2, pair tile 256 resamples x (NACC * 256) digit columns.
//
// The single-CTA kernel above is bound by L2 -> SM delivery, not by the tensor cores: a 128 x 256 tile does
// 128*256/(128+256) = 85 MAC per operand byte, the L2 feeds ~42 B/clk per SM (6300 B/clk per chip,
// B300_MICROARCH.md), the tensor cores want 8192 MAC/clk -- 44 % of the int8 peak, which is what it measures.
// Two CTAs of one cluster (the two SMs of a TPC) share their operands: each loads 128 of the 256 resample rows
// and half of the digit columns, the leader issues M = 256 MMAs that read both CTAs' shared memory and write
// both CTAs' TMEM.  With two 256-column accumulators per CTA (all 512 TMEM columns) the pair tile is
// 256 x 512: 171 MAC per byte.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Plain arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}
// TMA load of a CTA pair: the bytes are counted on the LEADER's barrier (peer bit of the address cleared).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1,
                                                 uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, int32_t c0, int32_t c1,
                                                 int32_t c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar & 0xFEFFFFFFu)
      : "memory");
}
// Arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed.
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int NACC, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
digit_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmQ,
                       int32_t *__restrict__ C, uint32_t count, uint32_t ldc, uint32_t tiles_m, uint32_t tiles_n,
                       uint32_t ksteps, uint32_t ksteps_per_split, uint32_t band, uint32_t no_reload,
                       uint32_t q_blocked) {
  // no_reload (sa_microbench_i8_tensor only): the ring is filled once and the MMAs keep re-reading it -- the
  // tensor-core rate of this instruction mix with no operand traffic, i.e. the roofline denominator.
  constexpr uint32_t PM = 256, PN = 256 * NACC;          // pair tile
  constexpr uint32_t kABytes = 128 * BK, kBBytes = NACC * 128 * BK;  // per CTA and stage
  constexpr uint32_t kCols = NACC * 256;                  // TMEM columns per CTA
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_base_sh;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sA = [&](uint32_t s) { return tiles_base + s * kABytes; };
  auto sB = [&](uint32_t s) { return tiles_base + STAGES * kABytes + s * kBBytes; };
  auto full = [&](uint32_t s) { return smem_u32(&bars[s]); };
  auto empty = [&](uint32_t s) { return smem_u32(&bars[STAGES + s]); };
  const uint32_t tmem_full = smem_u32(&bars[2 * STAGES]);

  const uint32_t pair = blockIdx.x >> 1;
  const uint32_t per_split = tiles_m * tiles_n;
  const uint32_t ks = pair / per_split;
  const uint32_t t = pair - ks * per_split;
  const uint32_t b = t / (band * tiles_n);
  const uint32_t in_band = t - b * band * tiles_n;
  const uint32_t bm = min(band, tiles_m - b * band);
  const uint32_t tile_m = b * band + in_band % bm, tile_n = in_band / bm;
  const uint32_t k0 = ks * ksteps_per_split;
  const uint32_t nk = min(ksteps_per_split, ksteps - k0);

  if (tid == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 2);   // leader: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(empty(s), 1);  // one multicast commit
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the same warp of both CTAs allocates
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {  // TMA producer of this CTA's halves
      const int32_t wrow = (int32_t)(tile_m * PM + rank * 128);
      for (uint32_t i = 0; i < (no_reload ? min(nk, (uint32_t)STAGES) : nk); ++i) {
        const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(full(s), 2 * (kABytes + kBBytes));
        else mbar_arrive_remote(full(s), 0);
        const int32_t kbyte = (int32_t)((k0 + i) * BK);
        tma_load_2d_pair(sA(s), &tmW, kbyte, wrow, full(s));
#pragma unroll
        for (uint32_t j = 0; j < NACC; ++j) {  // accumulator j: columns [j*256, j*256+256) of the pair tile, half each
          const int32_t col = (int32_t)(tile_n * PN + j * 256 + rank * 128);
          if (q_blocked) tma_load_3d_pair(sB(s) + j * 128 * BK, &tmQ, 0, col, (int32_t)(k0 + i), full(s));
          else tma_load_2d_pair(sB(s) + j * 128 * BK, &tmQ, kbyte, col, full(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {  // the leader issues the MMAs of the pair
      constexpr uint32_t idesc = instr_desc_i8(256, 256);
      for (uint32_t i = 0; i < nk; ++i) {
        const uint32_t s = i % STAGES, ph = (i / STAGES) & 1;
        if (!no_reload || i < STAGES) mbar_wait(full(s), ph);
        tc_fence_after();
        const uint64_t ad = smem_desc_sw128(sA(s));
#pragma unroll
        for (uint32_t k = 0; k < BK / UK; ++k)
#pragma unroll
          for (uint32_t j = 0; j < NACC; ++j)
            umma_i8_pair(tmem + j * 256, ad + 2 * k, smem_desc_sw128(sB(s) + j * 128 * BK) + 2 * k, idesc,
                         (i | k) != 0);
        tc_commit_pair(empty(s));
      }
      tc_commit_pair(tmem_full);
    }
    __syncwarp();
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t quarter = warp & 3;
    const uint32_t row = tile_m * PM + rank * 128 + quarter * 32 + lane;
    int32_t *crow = C + ((uint64_t)ks * count + row) * ldc + (uint64_t)tile_n * PN;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < PN; c0 += 32) {
      if (tile_n * PN + c0 >= ldc) break;  // warp-uniform
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((quarter * 32u) << 16) + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < count) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4 *>(crow + c0 + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  cluster_sync();  // the peer's shared memory and TMEM stay alive until the pair is done
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols) : "memory");
}

// ------------------------------------------------------------------------------------------
// Persistent, K-synchronised CTA-pair kernel (the production kernel of the integer engine).
//
// Round-1 ncu of the one-tile-per-pair kernel above at the headline: 436 GB of DRAM reads for 26 GB of
// operands (16x), L2 hit rate 47 % -- the 74 co-resident pairs share a band of W rows / Q columns but drift
// apart in K, so each re-reads from HBM what a neighbour fetched a few hundred microseconds earlier; that
// HBM power is what holds the SM clock (and with it the tensor pipe) at 1.3-1.4 GHz.  Here
//   * P pairs (all that fit the GPU at once, cudaOccupancyMaxActiveClusters) stay resident and walk the tile
//     list in WAVES: in wave w pair p owns tile w*P + p (band-major order, so a wave is ~band m-tiles x
//     P/band n-tiles); ring, barriers and TMEM live across tiles, the epilogue of a tile (TMEM -> global,
//     ~10 us of a ~6 ms tile) hands TMEM back through a `tmem_empty` barrier both CTAs' epilogue warps arrive
//     on, while the producers already refill the ring for the next tile;
//   * the leaders' producers keep the pairs of a wave within `slack` epochs of `epoch` K-steps of each other:
//     at every epoch boundary a producer bumps a global counter and waits (bounded: it falls through after
//     kSyncTimeoutNs and stops trying after kSyncStrikes time-outs, so correctness never depends on
//     co-residency) until every pair has reached the boundary `slack` epochs back.  The operands of a wave's
//     K-step (~0.85 MB) are then fetched from HBM once and served to all pairs from L2;
//   * the tiles left over after the full waves are cut along K into `rsplit` chunks spread over all pairs
//     (stream-K style remainder; int32 atomics into C, which the launcher zeroes for those tiles -- integer
//     addition is associative, the result is bit-identical).  Small launches are all remainder.
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kSyncTimeoutNs = 200000ull;  // 0.2 ms per wait, ~30 tile K-steps
constexpr int kSyncStrikes = 4;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The next batch's multiplicities, drawn by spare warps of the GEMM CTAs (counts.cuh) while the tensor cores work on
// the current batch: CTA x draws resamples c0 + x, c0 + x + gridDim.x, ... of `count` over the row window.
struct DrawArgs {
  uint64_t seed, N, c0, count, row0, nrows, Npad;
  uint8_t *w;
  int L;
};
constexpr int kDrawThreads = 512;

struct PairPlan {
  uint32_t tiles_m, tiles_n, band;   // tile grid (pair tiles of 256 x PN), band of m-tiles swept over n
  uint32_t ksteps;                   // K-steps (128 rows) of a whole tile
  uint32_t pairs;                    // P: resident pairs = gridDim.x / 2
  uint32_t full_waves;               // waves of P whole tiles
  uint32_t rem_tiles, rsplit, rper;  // left-over tiles, K chunks per left-over tile, K-steps per chunk
  uint32_t epoch, slack;             // K-sync: K-steps per epoch, epochs a pair may run ahead (0 = off)
};

template <int NACC, int STAGES, bool DRAW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DRAW ? kThreads + kDrawThreads : kThreads, 1)
digit_gemm_pair_persistent_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmQ,
                                  int32_t *__restrict__ C, uint32_t count, uint32_t ldc, const PairPlan pl,
                                  uint32_t *__restrict__ sync_ctr, uint32_t no_reload, uint32_t q_blocked,
                                  const DrawArgs draw) {
  constexpr uint32_t PM = 256, PN = 256 * NACC;
  constexpr uint32_t kABytes = 128 * BK, kBBytes = NACC * 128 * BK;
  constexpr uint32_t kCols = NACC * 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 2];
  __shared__ uint32_t tmem_base_sh;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sA = [&](uint32_t s) { return tiles_base + s * kABytes; };
  auto sB = [&](uint32_t s) { return tiles_base + STAGES * kABytes + s * kBBytes; };
  auto full = [&](uint32_t s) { return smem_u32(&bars[s]); };
  auto empty = [&](uint32_t s) { return smem_u32(&bars[STAGES + s]); };
  const uint32_t tmem_full = smem_u32(&bars[2 * STAGES]);
  const uint32_t tmem_empty = smem_u32(&bars[2 * STAGES + 1]);

  const uint32_t pair = blockIdx.x >> 1;
  // work list of this pair: full_waves whole tiles, then its share of the K-chunked left-over tiles
  const uint32_t rem_units = pl.rem_tiles * pl.rsplit;
  const uint32_t my_rem = pair < rem_units ? (rem_units - pair + pl.pairs - 1) / pl.pairs : 0;
  const uint32_t n_units = pl.full_waves + my_rem;
  struct Unit {
    uint32_t tile_m, tile_n, k0, nk, atomic;
  };
  auto unit = [&](uint32_t idx) {
    Unit u;
    uint32_t t;
    if (idx < pl.full_waves) {
      t = idx * pl.pairs + pair;
      u.k0 = 0;
      u.nk = pl.ksteps;
      u.atomic = 0;
    } else {
      const uint32_t r = pair + (idx - pl.full_waves) * pl.pairs;  // units running side by side share a K chunk
      t = pl.full_waves * pl.pairs + r % pl.rem_tiles;
      u.k0 = (r / pl.rem_tiles) * pl.rper;
      u.nk = min(pl.rper, pl.ksteps - u.k0);
      u.atomic = pl.rsplit > 1;
    }
    const uint32_t b = t / (pl.band * pl.tiles_n);
    const uint32_t in_band = t - b * pl.band * pl.tiles_n;
    const uint32_t bm = min(pl.band, pl.tiles_m - b * pl.band);
    u.tile_m = b * pl.band + in_band % bm;
    u.tile_n = in_band / bm;
    return u;
  };

  if (tid == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(full(s), 2);   // leader: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(empty(s), 1);  // one multicast commit
    }
    mbar_init(tmem_full, 1);   // one multicast commit per unit
    mbar_init(tmem_empty, 8);  // leader: 4 epilogue warps of each CTA per unit
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {  // TMA producer of this CTA's halves
      uint32_t it = 0;  // running K-step of this pair: ring stage and phase
      int strikes = 0;
      const bool pace = rank == 0 && pl.slack > 0 && sync_ctr != nullptr && !no_reload;
      for (uint32_t idx = 0; idx < n_units; ++idx) {
        const Unit u = unit(idx);
        const int32_t wrow = (int32_t)(u.tile_m * PM + rank * 128);
        const uint32_t steps = no_reload ? min(u.nk, (uint32_t)STAGES) : u.nk;
        for (uint32_t i = 0; i < steps; ++i, ++it) {
          if (pace && idx < pl.full_waves && strikes < kSyncStrikes) {
            const uint32_t g = idx * pl.ksteps + i;  // the same for every pair in a full wave
            if (g % pl.epoch == 0) {
              const uint32_t e = g / pl.epoch;
              atomicAdd(sync_ctr + e, 1u);
              if (e >= pl.slack && ld_relaxed_u32(sync_ctr + e - pl.slack) < pl.pairs) {
                uint64_t t0, now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                while (ld_relaxed_u32(sync_ctr + e - pl.slack) < pl.pairs) {
                  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                  if (now - t0 > kSyncTimeoutNs) {
                    ++strikes;
                    break;
                  }
                }
              }
            }
          }
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty(s), ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full(s), 2 * (kABytes + kBBytes));
          else mbar_arrive_remote(full(s), 0);
          const int32_t kbyte = (int32_t)((u.k0 + i) * BK);
          tma_load_2d_pair(sA(s), &tmW, kbyte, wrow, full(s));
#pragma unroll
          for (uint32_t j = 0; j < NACC; ++j) {
            const int32_t col = (int32_t)(u.tile_n * PN + j * 256 + rank * 128);
            if (q_blocked) tma_load_3d_pair(sB(s) + j * 128 * BK, &tmQ, 0, col, (int32_t)(u.k0 + i), full(s));
            else tma_load_2d_pair(sB(s) + j * 128 * BK, &tmQ, kbyte, col, full(s));
          }
        }
        if (no_reload) break;  // microbench: one ring fill, ever
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {  // the leader issues the MMAs of the pair
      constexpr uint32_t idesc = instr_desc_i8(256, 256);
      uint32_t it = 0;
      for (uint32_t idx = 0; idx < n_units; ++idx) {
        const Unit u = unit(idx);
        if (idx > 0) {  // the epilogue warps of both CTAs have drained the previous unit's accumulators
          mbar_wait(tmem_empty, (idx - 1) & 1);
          tc_fence_after();
        }
        for (uint32_t i = 0; i < u.nk; ++i, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          if (!no_reload || it < STAGES) mbar_wait(full(s), ph);
          tc_fence_after();
          const uint64_t ad = smem_desc_sw128(sA(s));
#pragma unroll
          for (uint32_t k = 0; k < BK / UK; ++k)
#pragma unroll
            for (uint32_t j = 0; j < NACC; ++j)
              umma_i8_pair(tmem + j * 256, ad + 2 * k, smem_desc_sw128(sB(s) + j * 128 * BK) + 2 * k, idesc,
                           (i | k) != 0);
          tc_commit_pair(empty(s));
        }
        tc_commit_pair(tmem_full);
      }
    }
    __syncwarp();
  } else if (DRAW && warp >= kThreads / 32) {
    // warps 6..21: the multiplicities of the next batch (they never touch the GEMM's barriers or TMEM).  Their
    // window histogram + split tree sit behind the operand ring.
    uint32_t *wsm = reinterpret_cast<uint32_t *>(smem_raw + (tiles_base - smem_u32(smem_raw)) +
                                                 (size_t)STAGES * (kABytes + kBBytes));
    const int dtid = (int)tid - kThreads;
    for (uint64_t i = blockIdx.x; i < draw.count; i += gridDim.x)
      counts_window_resample<kDrawThreads, 1>(wsm, draw.seed, draw.c0 + i, draw.N, draw.L, draw.row0, draw.nrows,
                                              draw.w + i * draw.Npad, dtid);
  } else {
    const uint32_t quarter = warp & 3;
    for (uint32_t idx = 0; idx < n_units; ++idx) {
      const Unit u = unit(idx);
      mbar_wait_relaxed(tmem_full, idx & 1);
      tc_fence_after();
      const uint32_t row = u.tile_m * PM + rank * 128 + quarter * 32 + lane;
      int32_t *crow = C + (uint64_t)row * ldc + (uint64_t)u.tile_n * PN;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < PN; c0 += 32) {
        if (u.tile_n * PN + c0 >= ldc) break;  // warp-uniform
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem + ((quarter * 32u) << 16) + c0)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < count) {
          if (u.atomic) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (v[q]) atomicAdd(crow + c0 + q, (int32_t)v[q]);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4 *>(crow + c0 + 4 * q) =
                  make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
      // hand the accumulators back to the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty) : "memory");
        else mbar_arrive_remote(tmem_empty, 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync();  // the peer's shared memory and TMEM stay alive until the pair is done
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kCols) : "memory");
}

// C[rows of tile][columns of tile] = 0 for the K-chunked left-over tiles (they are accumulated with atomics)
__global__ void __launch_bounds__(256)
zero_tiles_kernel(int32_t *__restrict__ C, uint32_t count, uint32_t ldc, uint32_t pn, uint32_t tiles_m,
                  uint32_t tiles_n, uint32_t band, uint32_t first_tile) {
  const uint32_t t = first_tile + blockIdx.x;
  const uint32_t b = t / (band * tiles_n);
  const uint32_t in_band = t - b * band * tiles_n;
  const uint32_t bm = min(band, tiles_m - b * band);
  const uint32_t tile_m = b * band + in_band % bm, tile_n = in_band / bm;
  const uint32_t c_lo = tile_n * pn, c_hi = min(ldc, c_lo + pn);
  const uint32_t r_lo = tile_m * 256, r_hi = min(count, r_lo + 256);
  if (c_lo >= c_hi) return;
  const uint32_t w4 = (c_hi - c_lo) >> 2;  // ldc and pn are multiples of 32
  for (uint32_t r = r_lo + (threadIdx.x / 64); r < r_hi; r += 4) {
    uint4 *dst = reinterpret_cast<uint4 *>(C + (uint64_t)r * ldc + c_lo);
    for (uint32_t k = threadIdx.x % 64; k < w4; k += 64) dst[k] = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------
// features -> digits
// ------------------------------------------------------------------------------------------
// Feature q of a normalised row x[0..step) in the reference order [A, AB_0.., (BA_0..,) B], SA_NACC order:
//   type 0: x[u]          type 1: x[u] * x[v]          type 2: x[u] * (x[v] - x[w])          type 3: (x[u] - x[v])^2
struct FeatDesc {
  uint16_t type, u, v, w;
};

__global__ void __launch_bounds__(256)
feature_table_kernel(int Dg, int step, int nfeat, FeatDesc *__restrict__ tab) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nfeat) return;
  FeatDesc d{0, 0, 0, 0};
  const int A = 0, B = step - 1;
  if (q == 0) d = {0, (uint16_t)A, 0, 0};
  else if (q == 1) d = {0, (uint16_t)B, 0, 0};
  else if (q == 2) d = {1, (uint16_t)A, (uint16_t)A, 0};
  else if (q == 3) d = {1, (uint16_t)B, (uint16_t)B, 0};
  else if (q == 4) d = {1, (uint16_t)A, (uint16_t)B, 0};
  else if (q < 5 + Dg) d = {2, (uint16_t)B, (uint16_t)(1 + q - 5), (uint16_t)A};       // B * (AB_j - A)
  else if (q < 5 + 2 * Dg) d = {3, (uint16_t)A, (uint16_t)(1 + q - 5 - Dg), 0};        // (A - AB_j)^2
  else {
    int p = q - 5 - 2 * Dg, j = 0;  // pair p -> (j, k), row-major over the strict upper triangle
    while (p >= Dg - 1 - j) {
      p -= Dg - 1 - j;
      ++j;
    }
    const int k = j + 1 + p;
    d = {1, (uint16_t)(1 + Dg + j), (uint16_t)(1 + k), 0};                             // BA_j * AB_k
  }
  tab[q] = d;
}

// Same operations, in the same order, as normalize_features_kernel (analyze.cu).
template <class X>
__device__ __forceinline__ double feature_value(const FeatDesc d, X x) {
  const double a = x(d.u);
  if (d.type == 0) return a;
  const double b = x(d.v);
  if (d.type == 1) return __dmul_rn(a, b);
  if (d.type == 2) return __dmul_rn(a, __dsub_rn(b, x(d.w)));
  const double t = __dsub_rn(a, b);
  return __dmul_rn(t, t);
}

// Normalised rows [i0, i0 + rows) of Y -> xs[col * pitch + slot(r)] (column-major in shared memory); rows >= N
// are 0.  INTERLEAVE: slot(r) = (r % 4) * (rows / 4) + r / 4, so that lanes which own 4 consecutive rows each
// read consecutive doubles for a given r % 4 (a plain row order would make those reads 4-way bank conflicts).
template <bool INTERLEAVE>
__device__ __forceinline__ void load_tile(const double *__restrict__ Y, uint64_t N, int step, uint64_t i0, int rows,
                                          int pitch, double mean, double sd, double *xs) {
  const int total = rows * step;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int r = e / step, c = e - r * step;
    const uint64_t i = i0 + r;
    const int slot = INTERLEAVE ? (r & 3) * (rows >> 2) + (r >> 2) : r;
    xs[c * pitch + slot] = i < N ? __ddiv_rn(__dsub_rn(Y[i * (uint64_t)step + c], mean), sd) : 0.0;
  }
}

// fmax_bits[q] (bit pattern of a non-negative double, zeroed by the caller) = max_i |F[i][q]|.
// big_count (zeroed by the caller) += the number of normalised values within kGuardBits bits of the largest |x| of
// the whole design -- the heavy-tail guard of the integer engine: the digits of a column are relative to its
// maximum, so if only a handful of rows carry the large values (an outlier row), a resample that happens to miss
// them is left with entries whose low bits were cut; "auto" then repeats the call with the FP64 engines
// (analyze/sobol.py), which is what keeps 1e-10 against the reference on such data.
constexpr int kGuardBits = 7;

__global__ void __launch_bounds__(256)
feature_max_kernel(const double *__restrict__ Y, uint64_t N, int step, int nfeat, const FeatDesc *__restrict__ tab,
                   const double *__restrict__ stats, int tile_rows, unsigned long long *__restrict__ fmax_bits,
                   unsigned long long *__restrict__ big_count) {
  extern __shared__ double dsm[];
  const int pitch = tile_rows + 1;
  double *xs = dsm;                       // [step][pitch]
  double *best = dsm + step * pitch;      // [nfeat], entry q owned by thread q % 256
  const double mean = stats[0], sd = stats[1];
  const double thr = ldexp(fmax(fabs(__ddiv_rn(__dsub_rn(stats[2], mean), sd)),
                                fabs(__ddiv_rn(__dsub_rn(stats[3], mean), sd))), -kGuardBits);
  unsigned big = 0;
  for (int q = threadIdx.x; q < nfeat; q += blockDim.x) best[q] = 0.0;
  const uint64_t tiles = ceil_div_u64(N, tile_rows);
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    load_tile<false>(Y, N, step, tile * tile_rows, tile_rows, pitch, mean, sd, xs);
    __syncthreads();
    {
      const uint64_t i0 = tile * tile_rows;
      const int rows = (int)((N - i0 < (uint64_t)tile_rows) ? N - i0 : (uint64_t)tile_rows);
      for (int e = threadIdx.x; e < rows * step; e += blockDim.x) {
        const int r = e / step, c = e - r * step;
        big += fabs(xs[c * pitch + r]) > thr;
      }
    }
    for (int q = threadIdx.x; q < nfeat; q += blockDim.x) {
      const FeatDesc d = tab[q];
      uint32_t m = (uint32_t)__double2hiint(best[q]);
      for (int r = 0; r < tile_rows; ++r)
        m = max(m, (uint32_t)__double2hiint(feature_value(d, [&](int c) { return xs[c * pitch + r]; })) & 0x7fffffffu);
      best[q] = __hiloint2double((int)m, 0);
    }
  }
  for (int q = threadIdx.x; q < nfeat; q += blockDim.x)
    atomicMax(fmax_bits + q, (unsigned long long)__double_as_longlong(best[q]));
  big = __reduce_add_sync(kFull, big);
  if ((threadIdx.x & 31) == 0 && big) atomicAdd(big_count, (unsigned long long)big);
}

// The same maxima with the pair products BA_j * AB_k (94 % of the columns at Dg = 64) on 4 x 4 register tiles:
// a thread owns four BA columns and four AB columns, reads two rows of each with one LDS.128 and keeps the 16
// running maxima in registers -- 8 loads per 32 products instead of 64, and no per-value descriptor decode (the
// first version spent 24 instructions per value, 2.5 ms at the headline).  Tiles on or above the diagonal of the
// (j, k) triangle are enumerated row-major; entries with k <= j are computed and dropped.  Used whenever the
// tile (even pitch, for the 16-byte loads) and the maxima fit 54 KB of shared memory; else the kernel above.
constexpr int kFmaxThreads = 160;
constexpr int kFmaxLoadBatch = 4;  // the kernel sits at 96 of its 102 registers: four loads in flight, not eight

// Only the EXPONENT of a column's largest |F| enters the scale (feature_scale_kernel), and the high 32 bits of
// |x| are monotone in |x|: the kernels track max(hi32(|F|)) with one integer max per value.  (A double-precision
// max is DSETP + two selects + moves on this architecture, ~6 instructions; ncu/SASS of the first version.)
__device__ __forceinline__ uint32_t abs_hi(double x) { return (uint32_t)__double2hiint(x) & 0x7fffffffu; }

// Position of model-output column c in the shared-memory tile: the AB columns (c = 1 + k) are stored as
// (k mod 4) * TB + k / 4, so that the columns 4 kb + b that consecutive lanes (consecutive kb) read with one
// instruction are neighbours -- 16-byte loads at a stride of one column pitch (68 words = 4 banks) are conflict-free,
// at four column pitches they were 4-way conflicts.  A, the BA columns and B keep their order behind them.
__device__ __forceinline__ int fmax_colpos(int c, int Dg, int TB) {
  if (c >= 1 && c <= Dg) {
    const int k = c - 1;
    return 1 + (k & 3) * TB + (k >> 2);
  }
  return c == 0 ? 0 : c + 4 * TB - Dg;
}

__global__ void __launch_bounds__(kFmaxThreads, 4)
feature_max_tiled_kernel(const double *__restrict__ Y, uint64_t N, int step, int Dg, int nhead, int nfeat,
                         const FeatDesc *__restrict__ tab, const double *__restrict__ stats, int tile_rows,
                         unsigned long long *__restrict__ fmax_bits, unsigned long long *__restrict__ big_count) {
  extern __shared__ __align__(16) double dsm[];
  const int pitch = tile_rows + 2;
  double *xs = dsm;                       // [step][pitch], column-major, rows in order
  const int TB = (Dg + 3) >> 2;
  const int cols = step + 4 * TB - Dg;    // the permuted AB block is padded to 4 TB columns
  uint32_t *best = reinterpret_cast<uint32_t *>(dsm + cols * pitch);  // [nfeat] max hi32(|F|)
  const double mean = stats[0], sd = stats[1];
  const double thr = ldexp(fmax(fabs(__ddiv_rn(__dsub_rn(stats[2], mean), sd)),
                                fabs(__ddiv_rn(__dsub_rn(stats[3], mean), sd))), -kGuardBits);
  unsigned big = 0;
  for (int q = threadIdx.x; q < nfeat; q += blockDim.x) best[q] = 0u;
  const int ntb = nfeat > nhead ? TB * (TB + 1) / 2 : 0;
  const uint64_t tiles = ceil_div_u64(N, tile_rows);
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    {
      const uint64_t i0 = tile * tile_rows;
      const int total = tile_rows * step;
      const int dr = (int)blockDim.x / step, dc = (int)blockDim.x - dr * step;
      int r = (int)threadIdx.x / step, c = (int)threadIdx.x - r * step;
      for (int e = threadIdx.x; e < total; e += kFmaxLoadBatch * (int)blockDim.x) {
        double y[kFmaxLoadBatch];  // loads in flight before the first use (see load_tile_split)
        int r1 = r, c1 = c;
#pragma unroll
        for (int k = 0; k < kFmaxLoadBatch; ++k) {
          const uint64_t i = i0 + r1;
          y[k] = (e + k * (int)blockDim.x < total && i < N) ? __ldg(Y + i * (uint64_t)step + c1) : 0.0;
          r1 += dr;
          c1 += dc;
          if (c1 >= step) {
            c1 -= step;
            ++r1;
          }
        }
#pragma unroll
        for (int k = 0; k < kFmaxLoadBatch; ++k) {
          if (e + k * (int)blockDim.x < total) {
            double x = 0.0;
            if (i0 + r < N) {
              x = __ddiv_rn(__dsub_rn(y[k], mean), sd);
              big += fabs(x) > thr;
            }
            xs[fmax_colpos(c, Dg, TB) * pitch + r] = x;
          }
          r += dr;
          c += dc;
          if (c >= step) {
            c -= step;
            ++r;
          }
        }
      }
    }
    __syncthreads();
    // scalar, square, first- and total-order features (type switch outside the row loop, two rows per load)
    for (int q = threadIdx.x; q < nhead; q += blockDim.x) {
      const FeatDesc d = tab[q];
      const double *pa = xs + fmax_colpos(d.u, Dg, TB) * pitch, *pb = xs + fmax_colpos(d.v, Dg, TB) * pitch,
                   *pc = xs + fmax_colpos(d.w, Dg, TB) * pitch;
      uint32_t m = best[q];
      if (d.type == 0) {
        for (int r = 0; r < tile_rows; r += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(pa + r);
          m = max(m, max(abs_hi(a.x), abs_hi(a.y)));
        }
      } else if (d.type == 1) {
        for (int r = 0; r < tile_rows; r += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(pa + r), b = *reinterpret_cast<const double2 *>(pb + r);
          m = max(m, max(abs_hi(__dmul_rn(a.x, b.x)), abs_hi(__dmul_rn(a.y, b.y))));
        }
      } else if (d.type == 2) {
        for (int r = 0; r < tile_rows; r += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(pa + r), b = *reinterpret_cast<const double2 *>(pb + r);
          const double2 c = *reinterpret_cast<const double2 *>(pc + r);
          m = max(m, max(abs_hi(__dmul_rn(a.x, __dsub_rn(b.x, c.x))), abs_hi(__dmul_rn(a.y, __dsub_rn(b.y, c.y)))));
        }
      } else {
        for (int r = 0; r < tile_rows; r += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(pa + r), b = *reinterpret_cast<const double2 *>(pb + r);
          const double t0 = __dsub_rn(a.x, b.x), t1 = __dsub_rn(a.y, b.y);
          m = max(m, max(abs_hi(__dmul_rn(t0, t0)), abs_hi(__dmul_rn(t1, t1))));
        }
      }
      best[q] = m;
    }
    // pair products on 4 x 4 tiles; |BA_j * AB_k| = |BA_j| * |AB_k| exactly, so the sign is dropped once per load
    for (int tb = threadIdx.x; tb < ntb; tb += blockDim.x) {
      int jb = 0, rest = tb;
      while (rest >= TB - jb) {
        rest -= TB - jb;
        ++jb;
      }
      const int kb = jb + rest;
      const double *pu[4], *pv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        pu[a] = xs + fmax_colpos(1 + Dg + min(4 * jb + a, Dg - 1), Dg, TB) * pitch;
        pv[a] = xs + fmax_colpos(1 + min(4 * kb + a, Dg - 1), Dg, TB) * pitch;
      }
      uint32_t m[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) m[a][b] = 0u;
#pragma unroll 2
      for (int r = 0; r < tile_rows; r += 2) {
        double2 u[4], v[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          u[a] = *reinterpret_cast<const double2 *>(pu[a] + r);
          v[a] = *reinterpret_cast<const double2 *>(pv[a] + r);
          u[a].x = fabs(u[a].x);
          u[a].y = fabs(u[a].y);
          v[a].x = fabs(v[a].x);
          v[a].y = fabs(v[a].y);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            m[a][b] = max(m[a][b], max((uint32_t)__double2hiint(__dmul_rn(u[a].x, v[b].x)),
                                       (uint32_t)__double2hiint(__dmul_rn(u[a].y, v[b].y))));
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int j = 4 * jb + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = 4 * kb + b;
          if (j < k && k < Dg) {
            const int q = nhead + j * (2 * Dg - j - 1) / 2 + (k - j - 1);
            best[q] = max(best[q], m[a][b]);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nfeat; q += blockDim.x)
    atomicMax(fmax_bits + q, (unsigned long long)best[q] << 32);
  big = __reduce_add_sync(kFull, big);
  if ((threadIdx.x & 31) == 0 && big) atomicAdd(big_count, (unsigned long long)big);
}

// scale[q] = 2^(bits - e), inv[q] = 2^(e - bits) with max |F[:, q]| < 2^e: |F * scale| < 2^bits
__global__ void __launch_bounds__(256)
feature_scale_kernel(const unsigned long long *__restrict__ fmax, int nfeat, int bits, double *__restrict__ scale,
                     double *__restrict__ inv, const unsigned long long *__restrict__ big_count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) inv[nfeat] = (double)*big_count;  // the heavy-tail guard rides behind the column scales
  if (q >= nfeat) return;
  const double m = __longlong_as_double((long long)fmax[q]);
  int e = 0;
  if (m > 0.0 && m < INFINITY) e = ilogb(m) + 1;
  e = max(-900, min(900, e));
  scale[q] = ldexp(1.0, bits - e);
  inv[q] = ldexp(1.0, e - bits);
}

// Q[i / 128][q*S + s][i % 128] = digit s of rint(F[i][q] * scale[q]).  Blocks take tiles of TR rows; LPF = TR/4
// lanes share a feature (4 consecutive rows each -> one 32-bit store per digit, TR contiguous bytes per column).
//
// Round-1 ncu of the first version: 5.1 ms for 15.8 GB (3.1 TB/s), 40 instructions per value, warps waiting on
// the per-feature descriptor / scale loads (long scoreboard 5.6 per issue) at 24 warps per SM.  This version
//   * walks the pair products BA_j * AB_k (94 % of the columns at Dg = 64) with (j, k) advanced incrementally --
//     no descriptor loads -- and the next feature's scale fetched one iteration ahead,
//   * reads the four rows of a lane with two LDS.128 per operand (tile layout below) and keeps two features in
//     flight per lane,
//   * writes the digits with streaming stores (Q is read next by the GEMM's TMA, 15.8 GB later).
// Tile layout: column-major, pitch = rows + 2 doubles; row r of the tile sits in slot
//   ((r >> 1) & 1) * (rows / 2) + 2 * (r >> 2) + (r & 1)
// so lane `sub` (rows 4 sub .. 4 sub + 3) finds rows (0, 1) at slot 2 sub and rows (2, 3) at rows / 2 + 2 sub:
// consecutive lanes read consecutive 16-byte chunks.
__device__ __forceinline__ int split_slot(int r, int rows) {
  return ((r >> 1) & 1) * (rows >> 1) + 2 * (r >> 2) + (r & 1);
}

constexpr int kLoadBatch = 8;  // global loads in flight per thread while a tile of Y is staged

__device__ __forceinline__ void load_tile_split(const double *__restrict__ Y, uint64_t N, int step, uint64_t i0,
                                                int rows, int pitch, double mean, double sd, double *xs) {
  // (row, column) of element e = threadIdx.x + k * blockDim.x, advanced without a division per element
  // kLoadBatch loads are issued back to back before the first is used: with one load in flight per thread the
  // 33 dependent trips to HBM per tile were 30 % of the kernel's stall samples (ncu source page, round 2)
  const int total = rows * step;
  const int dr = (int)blockDim.x / step, dc = (int)blockDim.x - dr * step;
  int r = (int)threadIdx.x / step, c = (int)threadIdx.x - r * step;
  for (int e = threadIdx.x; e < total; e += kLoadBatch * (int)blockDim.x) {
    double y[kLoadBatch];
    int r1 = r, c1 = c;
#pragma unroll
    for (int k = 0; k < kLoadBatch; ++k) {
      const uint64_t i = i0 + r1;
      y[k] = (e + k * (int)blockDim.x < total && i < N) ? __ldcs(Y + i * (uint64_t)step + c1) : 0.0;
      r1 += dr;
      c1 += dc;
      if (c1 >= step) {
        c1 -= step;
        ++r1;
      }
    }
#pragma unroll
    for (int k = 0; k < kLoadBatch; ++k) {
      if (e + k * (int)blockDim.x < total)
        xs[c * pitch + split_slot(r, rows)] = i0 + r < N ? __ddiv_rn(__dsub_rn(y[k], mean), sd) : 0.0;
      r += dr;
      c += dc;
      if (c >= step) {
        c -= step;
        ++r;
      }
    }
  }
}

// The four rows of this lane of column c.
__device__ __forceinline__ void lds_rows4(const double *xs, int c, int pitch, int half, int sub, double v[4]) {
  const double2 a = *reinterpret_cast<const double2 *>(xs + c * pitch + 2 * sub);
  const double2 b = *reinterpret_cast<const double2 *>(xs + c * pitch + half + 2 * sub);
  v[0] = a.x;
  v[1] = a.y;
  v[2] = b.x;
  v[3] = b.y;
}

// Balanced digits without a carry chain: with u_s = d_s + 128 in [0, 255],  t + sum_s 128 * 256^s =
// sum_s u_s * 256^s  is the plain base-256 form of a non-negative number below 256^S (|t| < 2^(8S-2)), and d_s
// as an int8 bit pattern is u_s with the top bit flipped.  One add and one xor per value give all digits; a
// 4 x 4 byte transpose (PRMT) per 32-bit half packs the four rows of each digit into a word.
template <int S>
__device__ __forceinline__ void emit_digits4(const double f[4], double sc, int8_t *dst) {
  constexpr unsigned long long kBias = 0x8080808080808080ull >> (8 * (8 - S));
  uint32_t lo[4], hi[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const unsigned long long u =
        ((unsigned long long)__double2ll_rn(__dmul_rn(f[rr], sc)) + kBias) ^ 0x8080808080808080ull;
    lo[rr] = (uint32_t)u;
    hi[rr] = (uint32_t)(u >> 32);
  }
  uint32_t packed[8];
  {
    const uint32_t a = __byte_perm(lo[0], lo[1], 0x5140), b = __byte_perm(lo[0], lo[1], 0x7362);
    const uint32_t c = __byte_perm(lo[2], lo[3], 0x5140), e = __byte_perm(lo[2], lo[3], 0x7362);
    packed[0] = __byte_perm(a, c, 0x5410);
    packed[1] = __byte_perm(a, c, 0x7632);
    packed[2] = __byte_perm(b, e, 0x5410);
    packed[3] = __byte_perm(b, e, 0x7632);
  }
  if (S > 4) {
    const uint32_t a = __byte_perm(hi[0], hi[1], 0x5140), b = __byte_perm(hi[0], hi[1], 0x7362);
    const uint32_t c = __byte_perm(hi[2], hi[3], 0x5140), e = __byte_perm(hi[2], hi[3], 0x7362);
    packed[4] = __byte_perm(a, c, 0x5410);
    packed[5] = __byte_perm(a, c, 0x7632);
    packed[6] = __byte_perm(b, e, 0x5410);
    packed[7] = __byte_perm(b, e, 0x7632);
  }
#pragma unroll
  for (int s = 0; s < S; ++s) __stcs(reinterpret_cast<uint32_t *>(dst + s * 128), packed[s]);
}

// Pair p + step of the strict upper triangle (row-major) from pair p = (j, k); the caller guarantees it exists.
__device__ __forceinline__ void advance_pair(int Dg, int step, int &j, int &k) {
  int pos = k - j - 1 + step;
  while (pos >= Dg - 1 - j) {
    pos -= Dg - 1 - j;
    ++j;
  }
  k = j + 1 + pos;
}

template <int S>
__global__ void __launch_bounds__(256)
digit_split_kernel(const double *__restrict__ Y, uint64_t N, int step, int Dg, int nhead, int nfeat,
                   const FeatDesc *__restrict__ tab, const double *__restrict__ stats,
                   const double *__restrict__ scale, int tile_rows, uint64_t Kp, int8_t *__restrict__ Q) {
  extern __shared__ __align__(16) double dsm[];
  const int pitch = tile_rows + 2, half = tile_rows >> 1;
  double *xs = dsm;
  const double mean = stats[0], sd = stats[1];
  const int lpf = tile_rows >> 2, fpw = 32 / lpf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane % lpf, g = warp * fpw + lane / lpf, G = nwarps * fpw;  // feature group of this lane
  const int npairs = nfeat - nhead;
  // first pair of this lane group
  int j0 = 0, k0 = 1;
  if (g < npairs) {
    int pos = g;
    while (pos >= Dg - 1 - j0) {
      pos -= Dg - 1 - j0;
      ++j0;
    }
    k0 = j0 + 1 + pos;
  }
  const uint64_t tiles = Kp / tile_rows;
  for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    __syncthreads();
    load_tile_split(Y, N, step, tile * tile_rows, tile_rows, pitch, mean, sd, xs);
    __syncthreads();
    // K-blocked layout: block kb = 128 consecutive rows of Y holds [ncols][128] bytes contiguously, so a tile's
    // digits form one streaming write (and one GEMM operand box = 128 columns x 128 bytes is contiguous too)
    const uint64_t row0 = tile * tile_rows;
    int8_t *qtile = Q + (row0 >> 7) * ((uint64_t)nfeat * S) * 128 + (row0 & 127) + 4 * sub;
    // scalar, square, first- and total-order features: descriptors from the table
    for (int q = g; q < nhead; q += G) {
      const FeatDesc d = tab[q];
      const double sc = scale[q];
      double a[4], b[4], f[4];
      lds_rows4(xs, d.u, pitch, half, sub, a);
      if (d.type == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = a[i];
      } else {
        lds_rows4(xs, d.v, pitch, half, sub, b);
        if (d.type == 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) f[i] = __dmul_rn(a[i], b[i]);
        } else if (d.type == 2) {
          double c[4];
          lds_rows4(xs, d.w, pitch, half, sub, c);
#pragma unroll
          for (int i = 0; i < 4; ++i) f[i] = __dmul_rn(a[i], __dsub_rn(b[i], c[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double t = __dsub_rn(a[i], b[i]);
            f[i] = __dmul_rn(t, t);
          }
        }
      }
      emit_digits4<S>(f, sc, qtile + (uint64_t)q * S * 128);
    }
    // pair products BA_j * AB_k, two per lane in flight
    if (g < npairs) {
      int j = j0, k = k0;
      double sc1 = scale[nhead + g];
      double sc2 = g + G < npairs ? scale[nhead + g + G] : 0.0;
      for (int p = g; p < npairs; p += 2 * G) {
        const bool two = p + G < npairs;
        int j2 = j, k2 = k;
        if (two) advance_pair(Dg, G, j2, k2);
        const double s1 = sc1, s2 = sc2;
        if (p + 2 * G < npairs) sc1 = scale[nhead + p + 2 * G];
        if (p + 3 * G < npairs) sc2 = scale[nhead + p + 3 * G];
        double ba1[4], ab1[4], ba2[4], ab2[4], f1[4], f2[4];
        lds_rows4(xs, 1 + Dg + j, pitch, half, sub, ba1);
        lds_rows4(xs, 1 + k, pitch, half, sub, ab1);
        lds_rows4(xs, 1 + Dg + j2, pitch, half, sub, ba2);
        lds_rows4(xs, 1 + k2, pitch, half, sub, ab2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f1[i] = __dmul_rn(ba1[i], ab1[i]);
          f2[i] = __dmul_rn(ba2[i], ab2[i]);
        }
        emit_digits4<S>(f1, s1, qtile + (uint64_t)(nhead + p) * S * 128);
        if (two) emit_digits4<S>(f2, s2, qtile + (uint64_t)(nhead + p + G) * S * 128);
        j = j2;
        k = k2;
        if (p + 2 * G < npairs) advance_pair(Dg, G, j, k);
      }
    }
  }
}

__device__ __forceinline__ double i128_to_double(__int128 v) {
  const bool neg = v < 0;
  const unsigned __int128 a = neg ? (unsigned __int128)0 - (unsigned __int128)v : (unsigned __int128)v;
  const uint64_t hi = (uint64_t)(a >> 64), lo = (uint64_t)a;
  double r;
  if (hi == 0) {
    r = __ull2double_rn(lo);
  } else {
    const int sh = 64 - __clzll((long long)hi);  // 1..64 bits above the low word
    uint64_t top = (uint64_t)(a >> sh);
    if ((a & (((unsigned __int128)1 << sh) - 1)) != 0) top |= 1;  // sticky bit: one correct rounding
    r = ldexp(__ull2double_rn(top), sh);
  }
  return neg ? -r : r;
}

// psum[c][q] = inv[q] * sum_s 256^s * sum_ks C[ks][c][q*S + s]
__global__ void __launch_bounds__(256)
digit_combine_kernel(const int32_t *__restrict__ C, uint64_t count, uint64_t ldc, int ksplit, int nfeat, int S,
                     const double *__restrict__ inv, double *__restrict__ psum) {
  const uint64_t total = count * (uint64_t)nfeat;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t c = idx / nfeat;
    const int q = (int)(idx - c * nfeat);
    __int128 acc = 0;
    for (int s = S - 1; s >= 0; --s) {
      long long part = 0;
      for (int ks = 0; ks < ksplit; ++ks) part += C[((uint64_t)ks * count + c) * ldc + (uint64_t)q * S + s];
      acc = acc * 256 + part;
    }
    psum[idx] = __dmul_rn(i128_to_double(acc), inv[q]);
  }
}

// The same for one C layer (ksplit == 1, what the persistent kernel leaves), a block per resample row: the row's
// digit sums are staged in shared memory with 16-byte loads, then thread t combines features t, t + 256, ... --
// their S words sit S apart, an odd stride for the default S = 7, so the reads do not collide on banks.  No
// 64-bit division per value and every sector of C is fetched once: the first kernel ran at 2.4 TB/s of C + psum
// traffic between two GEMM batches (67 us per batch at the headline).
__global__ void __launch_bounds__(256)
digit_combine_rows_kernel(const int32_t *__restrict__ C, uint64_t count, uint64_t ldc, int nfeat, int S,
                          const double *__restrict__ inv, double *__restrict__ psum) {
  extern __shared__ int32_t row_sh[];
  const uint32_t words4 = (uint32_t)(ldc >> 2);  // ldc is a multiple of 32
  for (uint64_t c = blockIdx.x; c < count; c += gridDim.x) {
    const uint4 *src = reinterpret_cast<const uint4 *>(C + c * ldc);
    uint4 *dst = reinterpret_cast<uint4 *>(row_sh);
    for (uint32_t k = threadIdx.x; k < words4; k += 256) dst[k] = __ldcs(src + k);
    __syncthreads();
    for (int q = threadIdx.x; q < nfeat; q += 256) {
      __int128 acc = 0;
      for (int s = S - 1; s >= 0; --s) acc = acc * 256 + (long long)row_sh[q * S + s];
      psum[c * (uint64_t)nfeat + q] = __dmul_rn(i128_to_double(acc), inv[q]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Byte matrix [rows][K] with row pitch `pitch`: boxes of box_rows x 128 bytes, 128-byte swizzle, zero fill
// outside [0, K) x [0, rows).
static int make_map(CUtensorMap *map, const void *base, uint64_t K, uint64_t rows, uint64_t pitch, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  SA_REQUIRE(fn, SA_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {pitch};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SA_REQUIRE(r == CUDA_SUCCESS, SA_ERR_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SA_OK;
}

// K-blocked byte matrix [nkb][cols][128]: boxes of box_cols x 128 bytes of one block, zero fill beyond `cols`.
static int make_map_blocked(CUtensorMap *map, const void *base, uint64_t cols, uint64_t nkb, int box_cols) {
  EncodeTiledFn fn = encode_tiled();
  SA_REQUIRE(fn, SA_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)BK, cols, nkb};
  const cuuint64_t strides[2] = {(cuuint64_t)BK, cols * BK};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_cols, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SA_REQUIRE(r == CUDA_SUCCESS, SA_ERR_ARG, "cuTensorMapEncodeTiled (blocked) failed with CUresult %d", (int)r);
  return SA_OK;
}

static uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// K splits so that small launches still fill the GPU (at least ~2 blocks per SM, at least 8 K-steps each).
static int choose_ksplit(uint64_t count, uint64_t ncols, uint64_t K, int bm, int bn) {
  const uint64_t tiles = ceil_div_u64(count, bm) * ceil_div_u64(ncols, bn) * (bm / BM);  // CTAs
  const uint64_t ksteps = ceil_div_u64(K, BK);
  const uint64_t want = 2 * (uint64_t)sm_count();
  uint64_t ks = tiles >= want ? 1 : ceil_div_u64(want, tiles);
  if (const char *e = getenv("SALIB_B200_DIGIT_KSPLIT")) ks = (uint64_t)atoi(e) > ks ? (uint64_t)atoi(e) : ks;
  const uint64_t cap = ksteps / 8 ? ksteps / 8 : 1;
  if (ks > cap) ks = cap;
  const uint64_t per = ceil_div_u64(ksteps, ks);
  return (int)ceil_div_u64(ksteps, per);  // no empty split
}

template <int BN, int STAGES>
static int launch_gemm(cudaStream_t st, const CUtensorMap &mw, const CUtensorMap &mq, int32_t *d_C, uint64_t count,
                       uint64_t ldc, uint64_t ncols, uint64_t K, int ksplit, uint32_t q_blocked) {
  const uint64_t tm = ceil_div_u64(count, BM), tn = ceil_div_u64(ncols, BN);
  const uint64_t ksteps = ceil_div_u64(K, BK), per = ceil_div_u64(ksteps, ksplit);
  const uint64_t blocks = tm * tn * (uint64_t)ksplit;
  SA_REQUIRE(blocks < (1ull << 31), SA_ERR_UNSUPPORTED, "digit GEMM: too many tiles");
  const size_t smem = (size_t)STAGES * (BM + BN) * BK + 1024;
  auto kern = digit_gemm_kernel<BN, STAGES>;
  SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, kThreads, smem, st>>>(mw, mq, d_C, (uint32_t)count, (uint32_t)ldc, (uint32_t)tm,
                                                (uint32_t)tn, (uint32_t)ksteps, (uint32_t)per, 16u, q_blocked);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

template <int NACC, int STAGES>
static int launch_gemm_pair_oneshot(cudaStream_t st, const CUtensorMap &mw, const CUtensorMap &mq, int32_t *d_C,
                                    uint64_t count, uint64_t ldc, uint64_t ncols, uint64_t K, int ksplit,
                                    uint32_t q_blocked) {
  const uint64_t tm = ceil_div_u64(count, 256), tn = ceil_div_u64(ncols, 256 * NACC);
  const uint64_t ksteps = ceil_div_u64(K, BK), per = ceil_div_u64(ksteps, ksplit);
  const uint64_t blocks = 2 * tm * tn * (uint64_t)ksplit;
  SA_REQUIRE(blocks < (1ull << 31), SA_ERR_UNSUPPORTED, "digit GEMM: too many tiles");
  const size_t smem = (size_t)STAGES * (128 + NACC * 128) * BK + 1024;
  auto kern = digit_gemm_pair_kernel<NACC, STAGES>;
  SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)blocks, kThreads, smem, st>>>(mw, mq, d_C, (uint32_t)count, (uint32_t)ldc, (uint32_t)tm,
                                                (uint32_t)tn, (uint32_t)ksteps, (uint32_t)per, 8u, 0u, q_blocked);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

static int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

// Pairs that can be resident at once (clusters of 2 CTAs, one CTA per SM): 74 on a full B200.
template <class Kern>
static int resident_pairs(Kern kern, size_t smem) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  static int cache[64];
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)sm_count(), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void *)kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = sm_count() / 2;
  }
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

// Ring of K-sync counters: slots of kSyncRing counters handed out round-robin, so launches in flight on
// different streams pace themselves on different counters.  One allocation per device and process.
constexpr uint32_t kSyncRing = 1u << 15;
constexpr int kSyncSlots = 8;
static uint32_t *sync_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  static uint32_t *base[64];
  static unsigned next[64];
  if (!base[dev]) {
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)kSyncSlots * kSyncRing * sizeof(uint32_t)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    base[dev] = (uint32_t *)p;
  }
  return base[dev] + (size_t)(next[dev]++ % kSyncSlots) * kSyncRing;
}

// Work plan of the persistent pair kernel (pure host logic, also exported for the tests): full waves of
// `pairs` tiles, then the left-over tiles cut into `rsplit` K chunks so that the last partial wave costs
// ~rem/pairs of a tile time instead of a whole one.
static PairPlan plan_pairs(uint64_t tm, uint64_t tn, uint64_t ksteps, int max_pairs, int band) {
  PairPlan pl = {};
  const uint64_t tiles = tm * tn;
  pl.tiles_m = (uint32_t)tm;
  pl.tiles_n = (uint32_t)tn;
  pl.band = (uint32_t)(band < 1 ? 1 : band);
  pl.ksteps = (uint32_t)ksteps;
  pl.full_waves = (uint32_t)(tiles / max_pairs);
  pl.rem_tiles = (uint32_t)(tiles % max_pairs);
  pl.rsplit = 1;
  pl.rper = (uint32_t)ksteps;
  if (pl.rem_tiles) {
    // rounds(s) = ceil(rem * s / P) chunks of ksteps / s each; a chunk keeps >= 16 K-steps; the small penalty
    // per split accounts for the extra epilogues (atomics)
    double best = 1e30;
    const uint32_t s_max = env_int("SALIB_B200_GEMM_RSPLIT_MAX", 16);
    for (uint32_t s = 1; s <= s_max && ksteps / s >= 16; ++s) {
      const uint64_t rounds = ceil_div_u64((uint64_t)pl.rem_tiles * s, max_pairs);
      const double cost = (double)rounds * (double)ceil_div_u64(ksteps, s) / (double)ksteps + 0.004 * s;
      if (cost < best - 1e-12) {
        best = cost;
        pl.rsplit = s;
      }
    }
    pl.rper = (uint32_t)ceil_div_u64(ksteps, pl.rsplit);
    pl.rsplit = (uint32_t)ceil_div_u64(ksteps, pl.rper);  // no empty chunk
  }
  const uint64_t rem_units = (uint64_t)pl.rem_tiles * pl.rsplit;
  pl.pairs = (uint32_t)(pl.full_waves ? (uint64_t)max_pairs : (rem_units < (uint64_t)max_pairs ? rem_units : max_pairs));
  pl.epoch = (uint32_t)env_int("SALIB_B200_GEMM_EPOCH", 16);
  pl.slack = (uint32_t)env_int("SALIB_B200_GEMM_SLACK", 1);
  if (pl.epoch < 1) pl.epoch = 1;
  if (pl.full_waves == 0 || pl.pairs < 2) pl.slack = 0;
  return pl;
}

// m-tiles per band: a wave of P tiles covers `band` m-tiles x P / band n-tiles and reads 256 * band W rows +
// 512 * P / band Q columns per K byte -- least for band = sqrt(2 P) (12 at P = 74); bands of equal height.
static int default_band(uint64_t tm, int pairs) {
  int ideal = 1;
  while ((ideal + 1) * (ideal + 1) <= 2 * pairs) ++ideal;
  const uint64_t nb = (tm + ideal / 2) / ideal;
  return (int)ceil_div_u64(tm, nb ? nb : 1);
}

// Does the launch of this shape go through the persistent pair kernel (which can host a draw)?
static bool pair_persistent(int variant, int ksplit) {
  return variant >= 2 && ksplit == 1 && env_int("SALIB_B200_GEMM_PERSIST", 1) != 0;
}

template <int NACC, int STAGES>
static int launch_gemm_pair(cudaStream_t st, const CUtensorMap &mw, const CUtensorMap &mq, int32_t *d_C,
                            uint64_t count, uint64_t ldc, uint64_t ncols, uint64_t K, int ksplit,
                            uint32_t q_blocked, uint32_t no_reload = 0, int force_pairs = 0,
                            const DrawArgs *draw = nullptr) {
  if (!no_reload && (ksplit > 1 || env_int("SALIB_B200_GEMM_PERSIST", 1) == 0)) {
    SA_REQUIRE(draw == nullptr, SA_ERR_ARG, "digit GEMM: a draw needs the persistent kernel");
    return launch_gemm_pair_oneshot<NACC, STAGES>(st, mw, mq, d_C, count, ldc, ncols, K, ksplit, q_blocked);
  }
  const uint64_t tm = ceil_div_u64(count, 256), tn = ceil_div_u64(ncols, 256 * NACC);
  const uint64_t ksteps = ceil_div_u64(K, BK);
  SA_REQUIRE(tm * tn < (1ull << 31), SA_ERR_UNSUPPORTED, "digit GEMM: too many tiles");
  const size_t ring = (size_t)STAGES * (128 + NACC * 128) * BK + 1024;
  const size_t smem = ring + (draw ? window_smem_bytes(draw->N, draw->L) : 0);
  SA_REQUIRE(smem <= 227 * 1024, SA_ERR_UNSUPPORTED, "digit GEMM: draw does not fit next to the operand ring");
  auto kern = draw ? digit_gemm_pair_persistent_kernel<NACC, STAGES, true>
                   : digit_gemm_pair_persistent_kernel<NACC, STAGES, false>;
  SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int max_pairs = force_pairs > 0 ? force_pairs
                                  : resident_pairs(digit_gemm_pair_persistent_kernel<NACC, STAGES, false>, ring);
  if (const int cap = env_int("SALIB_B200_GEMM_PAIRS", 0)) max_pairs = cap < max_pairs ? cap : max_pairs;
  SA_REQUIRE(max_pairs >= 1, SA_ERR_UNSUPPORTED, "digit GEMM: no CTA pair fits this device");
  PairPlan pl = plan_pairs(tm, tn, ksteps, max_pairs, env_int("SALIB_B200_GEMM_BAND", default_band(tm, max_pairs)));
  uint32_t *ctr = nullptr;
  if (pl.slack > 0 && !no_reload) {
    // counters are used as a ring: epochs per launch must fit, else widen the epoch
    const uint64_t steps = (uint64_t)pl.full_waves * pl.ksteps;
    if (ceil_div_u64(steps, pl.epoch) + 1 > kSyncRing) pl.epoch = (uint32_t)ceil_div_u64(steps, kSyncRing - 1);
    ctr = sync_slot();
    if (ctr) SA_CUDA(cudaMemsetAsync(ctr, 0, (size_t)kSyncRing * sizeof(uint32_t), st));
    else pl.slack = 0;
  }
  if (pl.rem_tiles && pl.rsplit > 1) {
    zero_tiles_kernel<<<pl.rem_tiles, 256, 0, st>>>(d_C, (uint32_t)count, (uint32_t)ldc, 256u * NACC, pl.tiles_m,
                                                   pl.tiles_n, pl.band, pl.full_waves * pl.pairs);
    note_launches(1);
  }
  const DrawArgs none = {};
  kern<<<2 * pl.pairs, draw ? kThreads + kDrawThreads : kThreads, smem, st>>>(
      mw, mq, d_C, (uint32_t)count, (uint32_t)ldc, pl, ctr, no_reload, q_blocked, draw ? *draw : none);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// Kernel variant: 0 = one CTA per 128 x 128 tile, 1 = one CTA per 128 x 256 tile, 2 = CTA pair per 256 x 256
// tile, 3 = CTA pair per 256 x 512 tile.
static int pick_variant(uint64_t count, uint64_t ncols) {
  if (const char *e = getenv("SALIB_B200_DIGIT_GEMM")) return atoi(e);
  if (ncols <= 128) return 0;
  if (ncols <= 256 || count <= 128) return 1;
  return 3;
}
static int variant_bm(int v) { return v >= 2 ? 256 : 128; }
static int variant_bn(int v) { return v == 0 ? 128 : (v == 3 ? 512 : 256); }

// Rows per shared-memory tile of the feature kernels: step * (rows + 1) doubles (+ `extra` doubles) within
// `budget` bytes, so that several blocks are resident per SM (the kernels are latency bound otherwise).
static int feature_tile_rows(int step, int extra, size_t budget, int max_rows) {
  for (int rows = max_rows; rows >= 8; rows >>= 1)
    if (((size_t)step * (rows + 1) + extra) * 8 <= budget) return rows;
  for (int rows = max_rows; rows >= 8; rows >>= 1)  // long rows: one block per SM
    if (((size_t)step * (rows + 1) + extra) * 8 <= 200 * 1024) return rows;
  return 0;
}

// Rows per tile of the split kernel: step * (rows + 2) doubles; 64 rows while three blocks fit an SM, then fewer.
static int split_tile_rows(int step) {
  if (const int r = env_int("SALIB_B200_SPLIT_ROWS", 0))  // experiments: 8, 16, 32 or 64
    if ((r == 8 || r == 16 || r == 32 || r == 64) && (size_t)step * (r + 2) * 8 <= 200 * 1024) return r;
  for (int rows = 64; rows >= 8; rows >>= 1)
    if ((size_t)step * (rows + 2) * 8 <= 72 * 1024) return rows;
  for (int rows = 64; rows >= 8; rows >>= 1)  // long rows: one block per SM
    if ((size_t)step * (rows + 2) * 8 <= 200 * 1024) return rows;
  return 0;
}

}  // namespace dg
}  // namespace sa

using namespace sa;
using namespace sa::dg;

// 1 if this process can run the integer engine: tcgen05/TMEM hardware (compute capability 10.x) and a driver
// that exports cuTensorMapEncodeTiled.
extern "C" int sa_digit_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 && encode_tiled() != nullptr;
}

extern "C" uint64_t sa_digit_row_pitch(uint64_t N) { return round_up(N, 128); }

extern "C" uint64_t sa_digit_ldc(int nfeat, int ndigits) { return round_up((uint64_t)nfeat * ndigits, 32); }

extern "C" int sa_digit_ksplit(uint64_t N, int nfeat, int ndigits, uint64_t count) {
  const uint64_t ncols = (uint64_t)nfeat * ndigits;
  const int v = pick_variant(count, ncols);
  // the persistent pair kernel cuts K by itself (left-over tiles, atomics into one C layer)
  if (v >= 2 && env_int("SALIB_B200_GEMM_PERSIST", 1) != 0 && !getenv("SALIB_B200_DIGIT_KSPLIT")) return 1;
  return choose_ksplit(count, ncols, N, variant_bm(v), variant_bn(v));
}

// Work plan of the persistent CTA-pair kernel for `tiles_m x tiles_n` pair tiles of `ksteps` K-steps on
// `max_pairs` resident pairs (pure host logic; tests).  out[0..7] = pairs, full_waves, rem_tiles, rsplit, rper,
// band, epoch, slack.
extern "C" int sa_digit_pair_plan(uint64_t tiles_m, uint64_t tiles_n, uint64_t ksteps, int max_pairs, int band,
                                  uint32_t *out) {
  SA_REQUIRE(out && tiles_m >= 1 && tiles_n >= 1 && ksteps >= 1 && max_pairs >= 1, SA_ERR_ARG,
             "sa_digit_pair_plan: bad arguments");
  const PairPlan pl = plan_pairs(tiles_m, tiles_n, ksteps, max_pairs, band);
  out[0] = pl.pairs; out[1] = pl.full_waves; out[2] = pl.rem_tiles; out[3] = pl.rsplit;
  out[4] = pl.rper; out[5] = pl.band; out[6] = pl.epoch; out[7] = pl.slack;
  return SA_OK;
}

extern "C" size_t sa_digit_features_workspace_bytes(int nfeat) {
  return (size_t)round_up((uint64_t)nfeat * (sizeof(FeatDesc) + 8 + 8) + 8, 256);
}

extern "C" size_t sa_digit_sums_workspace_bytes(uint64_t N, int nfeat, int ndigits, uint64_t count) {
  return (size_t)sa_digit_ksplit(N, nfeat, ndigits, count) * count * sa_digit_ldc(nfeat, ndigits) * sizeof(int32_t);
}

static int digit_gemm_impl(void *stream, const uint8_t *d_w, uint64_t w_pitch, uint64_t count, const int8_t *d_Q,
                           uint64_t q_pitch, uint64_t ncols, uint64_t K, int ksplit, int32_t *d_C, uint64_t ldc,
                           const DrawArgs *draw) {
  SA_REQUIRE(d_w && d_Q && d_C, SA_ERR_ARG, "sa_digit_gemm_i32: null pointer");
  SA_REQUIRE(count >= 1 && ncols >= 1 && K >= 1 && ksplit >= 1, SA_ERR_ARG, "sa_digit_gemm_i32: bad shape");
  SA_REQUIRE(K <= (1ull << 24), SA_ERR_UNSUPPORTED, "sa_digit_gemm_i32: more than 2^24 rows could overflow int32");
  const uint32_t blocked = q_pitch == 0;  // Q stored [ceil(K / 128)][ncols][128]
  SA_REQUIRE(w_pitch % 16 == 0 && w_pitch >= K && (blocked || (q_pitch % 16 == 0 && q_pitch >= K)), SA_ERR_ARG,
             "sa_digit_gemm_i32: row pitches must be multiples of 16 bytes and >= K");
  SA_REQUIRE(((uintptr_t)d_w & 15) == 0 && ((uintptr_t)d_Q & 15) == 0 && ((uintptr_t)d_C & 15) == 0, SA_ERR_ARG,
             "sa_digit_gemm_i32: pointers must be 16-byte aligned");
  SA_REQUIRE(ldc % 32 == 0 && ldc >= ncols, SA_ERR_ARG, "sa_digit_gemm_i32: ldc must be a multiple of 32 >= ncols");
  SA_REQUIRE((uint64_t)ksplit <= ceil_div_u64(K, BK), SA_ERR_ARG, "sa_digit_gemm_i32: more K splits than K steps");
  {
    const uint64_t ksteps = ceil_div_u64(K, BK), per = ceil_div_u64(ksteps, ksplit);
    SA_REQUIRE(per * (uint64_t)(ksplit - 1) < ksteps, SA_ERR_ARG, "sa_digit_gemm_i32: ksplit leaves an empty split");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int v = pick_variant(count, ncols);
  CUtensorMap mw, mq;
  int rc = make_map(&mw, d_w, K, count, w_pitch, BM);  // every variant loads W in boxes of 128 rows
  if (rc) return rc;
  rc = blocked ? make_map_blocked(&mq, d_Q, ncols, ceil_div_u64(K, BK), v == 1 ? 256 : 128)
               : make_map(&mq, d_Q, K, ncols, q_pitch, v == 1 ? 256 : 128);
  if (rc) return rc;
  SA_REQUIRE(draw == nullptr || pair_persistent(v, ksplit), SA_ERR_ARG,
             "digit GEMM: only the persistent pair kernel hosts a draw");
  switch (v) {
    case 0: return launch_gemm<128, 6>(st, mw, mq, d_C, count, ldc, ncols, K, ksplit, blocked);
    case 1: return launch_gemm<256, 4>(st, mw, mq, d_C, count, ldc, ncols, K, ksplit, blocked);
    case 2: return launch_gemm_pair<1, 6>(st, mw, mq, d_C, count, ldc, ncols, K, ksplit, blocked, 0, 0, draw);
    case 3: return launch_gemm_pair<2, 4>(st, mw, mq, d_C, count, ldc, ncols, K, ksplit, blocked, 0, 0, draw);
  }
  SA_REQUIRE(false, SA_ERR_ARG, "sa_digit_gemm_i32: unknown kernel variant %d", v);
}

extern "C" int sa_digit_gemm_i32(void *stream, const uint8_t *d_w, uint64_t w_pitch, uint64_t count,
                                 const int8_t *d_Q, uint64_t q_pitch, uint64_t ncols, uint64_t K, int ksplit,
                                 int32_t *d_C, uint64_t ldc) {
  return digit_gemm_impl(stream, d_w, w_pitch, count, d_Q, q_pitch, ncols, K, ksplit, d_C, ldc, nullptr);
}

extern "C" int sa_digit_features_i8(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                                    const double *d_stats, int ndigits, int8_t *d_Q, double *d_inv, void *d_work,
                                    size_t work_bytes) {
  SA_REQUIRE(d_Y && d_stats && d_Q && d_inv && d_work, SA_ERR_ARG, "sa_digit_features_i8: null pointer");
  SA_REQUIRE(N >= 1 && Dg >= 1 && Dg <= 16000, SA_ERR_ARG, "sa_digit_features_i8: bad shape");
  SA_REQUIRE(ndigits >= 2 && ndigits <= kMaxDigits, SA_ERR_ARG, "sa_digit_features_i8: 2 <= ndigits <= 8");
  SA_REQUIRE(N <= (1ull << 24), SA_ERR_UNSUPPORTED, "sa_digit_features_i8: N > 2^24");
  const int nfeat = SA_NACC(Dg, second_order);
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  SA_REQUIRE(work_bytes >= sa_digit_features_workspace_bytes(nfeat), SA_ERR_WORKSPACE,
             "sa_digit_features_i8: workspace too small");
  const int rows_max = feature_tile_rows(step, nfeat, 54 * 1024, 32);   // 4 blocks per SM
  const int rows = split_tile_rows(step);                              // 3 blocks per SM while step allows
  SA_REQUIRE(rows >= 8 && rows_max >= 8, SA_ERR_UNSUPPORTED,
             "sa_digit_features_i8: %d outputs per base row do not fit shared memory", step);
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t Kp = sa_digit_row_pitch(N);
  unsigned long long *fmax = (unsigned long long *)d_work;   // [nfeat] + the guard counter
  double *scale = (double *)(fmax + nfeat + 1);
  FeatDesc *tab = (FeatDesc *)(scale + nfeat);
  const int fb = (nfeat + 255) / 256;
  SA_CUDA(cudaMemsetAsync(fmax, 0, ((size_t)nfeat + 1) * 8, st));
  feature_table_kernel<<<fb, 256, 0, st>>>(Dg, step, nfeat, tab);
  const size_t smem_max = ((size_t)step * (rows_max + 1) + nfeat) * 8;
  const size_t smem_split = (size_t)step * (rows + 2) * 8;
  const int nhead = 5 + 2 * Dg;
  SA_CUDA(cudaFuncSetAttribute(feature_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  const uint64_t tiles = ceil_div_u64(N, rows_max), slots = 4 * (uint64_t)sm_count();
  const unsigned grid = (unsigned)(tiles < slots ? tiles : slots);
  // 4 x 4 register tiles when the tile (even pitch) + the maxima fit 54 KB (4 blocks per SM); else thread per feature
  int rows_t = 0;
  const size_t cols_t = (size_t)step + 4 * ((Dg + 3) / 4) - Dg;
  for (int r = 32; r >= 8 && !rows_t; r >>= 1)
    if ((cols_t * (r + 2) + nfeat) * 8 <= 54 * 1024) rows_t = r;
  if (rows_t && !getenv("SALIB_B200_FMAX_PLAIN")) {
    const size_t smem_t = (cols_t * (rows_t + 2) + nfeat) * 8;
    SA_CUDA(cudaFuncSetAttribute(feature_max_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const uint64_t tiles_t = ceil_div_u64(N, rows_t);
    int per_sm = 0;  // one full wave: a partial second wave would run at a fraction of the occupancy
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, feature_max_tiled_kernel, kFmaxThreads, smem_t) !=
            cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    const uint64_t wave = (uint64_t)per_sm * sm_count();
    const unsigned grid_t = (unsigned)(tiles_t < wave ? tiles_t : wave);
    feature_max_tiled_kernel<<<grid_t, kFmaxThreads, smem_t, st>>>(d_Y, N, step, Dg, 5 + 2 * Dg, nfeat, tab, d_stats,
                                                                  rows_t, fmax, fmax + nfeat);
  } else {
    feature_max_kernel<<<grid, 256, smem_max, st>>>(d_Y, N, step, nfeat, tab, d_stats, rows_max, fmax, fmax + nfeat);
  }
  feature_scale_kernel<<<fb, 256, 0, st>>>(fmax, nfeat, 8 * ndigits - 2, scale, d_inv, fmax + nfeat);
  int split_per_sm = 3;  // resident blocks per SM of the split kernel (one full wave of blocks walks the tiles)
  if (env_int("SALIB_B200_SPLIT_ROWS", 0)) {
    int per_sm = 0;
    auto kern = ndigits == 7 ? (const void *)digit_split_kernel<7> : (const void *)digit_split_kernel<6>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_split);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem_split) == cudaSuccess && per_sm >= 1)
      split_per_sm = per_sm;
    else
      cudaGetLastError();
  }
  const uint64_t tiles2 = Kp / rows, slots2 = (uint64_t)split_per_sm * sm_count();
  const unsigned grid2 = (unsigned)(tiles2 < slots2 ? tiles2 : slots2);
#define SA_SPLIT(S)                                                                                               \
  case S:                                                                                                         \
    SA_CUDA(cudaFuncSetAttribute(digit_split_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                 (int)smem_split));                                                               \
    digit_split_kernel<S><<<grid2, 256, smem_split, st>>>(d_Y, N, step, Dg, nhead, nfeat, tab, d_stats, scale,    \
                                                          rows, Kp, d_Q);                                        \
    break;
  switch (ndigits) {
    SA_SPLIT(2) SA_SPLIT(3) SA_SPLIT(4) SA_SPLIT(5) SA_SPLIT(6) SA_SPLIT(7) SA_SPLIT(8)
  }
#undef SA_SPLIT
  note_launches(4);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_counts_philox_rows_ex(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                                        uint64_t row0, uint64_t nrows, uint8_t *d_w, int32_t *d_overflow, int slim);

extern "C" int sa_digit_sums_draw_f64(void *stream, const int8_t *d_Q, const double *d_inv, uint64_t N, int nfeat,
                                      int ndigits, const uint8_t *d_w, uint64_t w_pitch, uint64_t count, void *d_work,
                                      size_t work_bytes, double *d_psum, uint64_t draw_seed, uint64_t draw_N,
                                      uint64_t draw_c0, uint64_t draw_count, uint64_t draw_row0, uint64_t draw_nrows,
                                      uint8_t *d_w_next) {
  SA_REQUIRE(d_Q && d_inv && d_w && d_work && d_psum, SA_ERR_ARG, "sa_digit_sums_f64: null pointer");
  SA_REQUIRE(draw_count == 0 || (d_w_next && draw_N >= 1 && draw_nrows >= 1 && draw_row0 + draw_nrows <= draw_N &&
                                 ((uintptr_t)d_w_next & 3) == 0),
             SA_ERR_ARG, "sa_digit_sums_draw_f64: bad draw arguments");
  SA_REQUIRE(N >= 1 && nfeat >= 1 && count >= 1, SA_ERR_ARG, "sa_digit_sums_f64: bad shape");
  SA_REQUIRE(ndigits >= 2 && ndigits <= kMaxDigits, SA_ERR_ARG, "sa_digit_sums_f64: 2 <= ndigits <= 8");
  SA_REQUIRE(work_bytes >= sa_digit_sums_workspace_bytes(N, nfeat, ndigits, count), SA_ERR_WORKSPACE,
             "sa_digit_sums_f64: workspace too small");
  const uint64_t ncols = (uint64_t)nfeat * ndigits, ldc = sa_digit_ldc(nfeat, ndigits);
  const int ksplit = sa_digit_ksplit(N, nfeat, ndigits, count);
  int32_t *C = (int32_t *)d_work;
  cudaStream_t st = (cudaStream_t)stream;
  DrawArgs draw = {};
  const DrawArgs *hosted = nullptr;
  if (draw_count) {
    const int L = window_levels(draw_N);
    const uint64_t Npad = sa_pad_rows(draw_nrows);
    const size_t ring = (size_t)4 * (128 + 2 * 128) * BK + 1024;
    if (L >= 0 && pair_persistent(pick_variant(count, ncols), ksplit) && !getenv("SALIB_B200_NO_HOSTED_DRAW") &&
        ring + window_smem_bytes(draw_N, L) <= 227 * 1024) {
      // hosted by the GEMM: spare warps of its CTAs draw while the tensor cores work (same stream, no events)
      if (Npad > draw_nrows)
        SA_CUDA(cudaMemset2DAsync(d_w_next + draw_nrows, Npad, 0, Npad - draw_nrows, draw_count, st));
      draw = DrawArgs{draw_seed, draw_N, draw_c0, draw_count, draw_row0, draw_nrows, Npad, d_w_next, L};
      hosted = &draw;
    } else {
      // shapes that do not run the persistent pair kernel are small: draw first, then multiply
      const int rc = sa_counts_philox_rows_ex(stream, draw_seed, draw_N, draw_c0, draw_count, draw_row0, draw_nrows,
                                              d_w_next, nullptr, 0);
      if (rc) return rc;
    }
  }
  const int rc = digit_gemm_impl(stream, d_w, w_pitch, count, d_Q, 0 /* K-blocked */, ncols, N, ksplit, C, ldc, hosted);
  if (rc) return rc;
  const uint64_t total = count * (uint64_t)nfeat;
  const unsigned grid = (unsigned)(ceil_div_u64(total, 256) < 148ull * 16 ? ceil_div_u64(total, 256) : 148ull * 16);
  const size_t row_bytes = ldc * sizeof(int32_t);
  if (ksplit == 1 && row_bytes <= 100 * 1024 && !getenv("SALIB_B200_COMBINE_PLAIN")) {
    SA_CUDA(cudaFuncSetAttribute(digit_combine_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)row_bytes));
    const unsigned grid_r = (unsigned)(count < 148ull * 8 ? count : 148ull * 8);
    digit_combine_rows_kernel<<<grid_r, 256, row_bytes, st>>>(C, count, ldc, nfeat, ndigits, d_inv, d_psum);
  } else {
    digit_combine_kernel<<<grid, 256, 0, st>>>(C, count, ldc, ksplit, nfeat, ndigits, d_inv, d_psum);
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_digit_sums_f64(void *stream, const int8_t *d_Q, const double *d_inv, uint64_t N, int nfeat,
                                 int ndigits, const uint8_t *d_w, uint64_t w_pitch, uint64_t count, void *d_work,
                                 size_t work_bytes, double *d_psum) {
  return sa_digit_sums_draw_f64(stream, d_Q, d_inv, N, nfeat, ndigits, d_w, w_pitch, count, d_work, work_bytes, d_psum,
                                0, 0, 0, 0, 0, 0, nullptr);
}

// Tensor-core roofline denominator: the CTA-pair kernel on one 256 x 512 tile per SM pair with the operand ring
// filled once (no TMA / L2 / HBM traffic in the loop).  ops = 2 * 256 * 512 * 128 * ksteps per pair.
extern "C" int sa_microbench_i8_tensor(void *stream, const uint8_t *d_w, const int8_t *d_Q, int32_t *d_C,
                                       int ksteps, double *h_ops) {
  SA_REQUIRE(d_w && d_Q && d_C && h_ops && ksteps >= 8, SA_ERR_ARG, "sa_microbench_i8_tensor: bad arguments");
  // buffers: d_w [pairs*256][1024] u8, d_Q [512][1024] s8, d_C [pairs*256][512] int32
  const uint64_t pairs = (uint64_t)sm_count() / 2, K = 1024;
  CUtensorMap mw, mq;
  int rc = make_map(&mw, d_w, K, pairs * 256, K, BM);
  if (rc) return rc;
  rc = make_map(&mq, d_Q, K, 512, K, 128);
  if (rc) return rc;
  // the K coordinate runs past the 1024-byte rows after 8 steps: those boxes are zero-filled, and never loaded
  // anyway (no_reload stops the producer after the first ring fill)
  rc = launch_gemm_pair<2, 4>((cudaStream_t)stream, mw, mq, d_C, pairs * 256, 512, 512, (uint64_t)ksteps * BK, 1, 0u,
                              1u, (int)pairs);
  if (rc) return rc;
  *h_ops = 2.0 * 256 * 512 * 128 * (double)ksteps * (double)pairs;
  return SA_OK;
}
