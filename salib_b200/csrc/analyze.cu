// sobol.analyze on the device: normalisation, resample multiplicities, the weighted
// first/total/second-order sums, estimator algebra and bootstrap confidence intervals.
//
// Reference behaviour replaced: src/SALib/analyze/sobol.py:141-182,209-246,267-279.
//
// Core idea: a bootstrap resample c is the multiset of rows r[:, c]; every sum the
// reference takes over the gathered arrays A[r], AB[r, j], BA[r, j], B[r] equals the same sum over
// ALL rows weighted by the multiplicity w_c[i] = #{t : r[t, c] == i}.  So the kernels stream the
// normalised output matrix in row order (coalesced, shared by all resamples in flight, L1/L2
// resident) instead of gathering N*R random 1 KB rows, and rows with w == 0 (36.8 %) are skipped.
//
//   second-order kernel  (s2_accum_smem_kernel): one warp per resample; the D(D-1)/2 pair sums
//     sum_i w*BA_j*AB_k live in registers as one 8x8 (j,k) tile per lane, i.e. a register-tiled
//     FP64 rank-1 update per row (16 operand doubles feed 64 DFMAs); rows are staged through shared
//     memory by cp.async.bulk.  FP64-pipe / shared-memory-bandwidth bound.
//   first-order kernel   (fo_gemv_kernel): all first/total-order sums are linear in w, so they are a
//     skinny GEMM w[R x N] * F[N x (5+2D)] over a per-row feature matrix; lanes <-> resamples.
//   generic kernel       (accum_generic_kernel): any shape, one accumulator per thread; fallback
//     for Dg > 64 with second order and an independent cross-check of the two fast kernels.
#include <type_traits>

#include <stdlib.h>

#include "common.cuh"
#include "counts.cuh"

namespace sa {

__host__ __device__ inline int round_up8(int x) { return (x + 7) & ~7; }
// Analysis layout: every block of 8 columns sits at a pitch of 10 doubles (80 B): 5*b mod 8 is a
// bijection, so the 16-byte chunks of 8 different blocks read with the same instruction offset fall
// into 8 different bank groups -- the tile kernel's LDS.128 are conflict free with plain offsets.
constexpr int kBlockPitch = 10;
__host__ __device__ inline int part_len(int Dg) { return (round_up8(Dg) >> 3) * kBlockPitch; }
__host__ __device__ inline int row_stride(int Dg, int second_order) {
  return (second_order ? 2 : 1) * part_len(Dg) + 2;
}
// Row pitch of the multiplicity matrix w[resample][row] (bytes).  A multiple of 128, but never of 4096:
// the first-order kernel reads 32 bytes of 32 different resamples at once, and a power-of-two pitch
// would put all of them into the same HBM channel/bank.
__host__ __device__ inline uint64_t pad_rows(uint64_t N) {
  uint64_t p = (N + 127) & ~(uint64_t)127;
  if ((p & 4095) == 0) p += 128;
  return p;
}
// Position of logical column j inside the AB / BA part of an analysis-layout row.
__host__ __device__ inline int col_pos(int j) { return (j >> 3) * kBlockPitch + (j & 7); }
__host__ __device__ inline int pair_index(int j, int k, int Dg) {
  return j * Dg - (j * (j + 1)) / 2 + (k - j - 1);
}

// ------------------------------------------------------------------------------------------
// moments: mean, population std, min/max (all values and A/B columns only)
// ------------------------------------------------------------------------------------------
constexpr int kMomBlocks = 1024;
constexpr int kMomThreads = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], const int (&op)[NV], double *dst) {
  // op: 0 sum, 1 min, 2 max
  __shared__ double sh[NV][kMomThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = v[q];
    x = (op[q] == 0) ? warp_sum(x) : (op[q] == 1 ? warp_min(x) : warp_max(x));
    if (lane == 0) sh[q][warp] = x;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      double x;
      if (lane < kMomThreads / 32) x = sh[q][lane];
      else x = (op[q] == 0) ? 0.0 : (op[q] == 1 ? INFINITY : -INFINITY);
      x = (op[q] == 0) ? warp_sum(x) : (op[q] == 1 ? warp_min(x) : warp_max(x));
      if (lane == 0) dst[q] = x;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMomThreads)
moments_pass1_kernel(const double *__restrict__ Y, uint64_t total, int step, double *__restrict__ part) {
  const uint64_t S = (uint64_t)gridDim.x * blockDim.x;
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int col = (int)(idx % (uint64_t)step);
  const int colstep = (int)(S % (uint64_t)step);
  double v[5] = {0.0, INFINITY, -INFINITY, INFINITY, -INFINITY};
  for (; idx < total; idx += S) {
    const double y = Y[idx];
    v[0] += y;
    v[1] = fmin(v[1], y);
    v[2] = fmax(v[2], y);
    if (col == 0 || col == step - 1) {
      v[3] = fmin(v[3], y);
      v[4] = fmax(v[4], y);
    }
    col += colstep;
    if (col >= step) col -= step;
  }
  const int op[5] = {0, 1, 2, 1, 2};
  block_reduce_store<5>(v, op, part + (size_t)blockIdx.x * 5);
}

__global__ void __launch_bounds__(kMomThreads)
moments_final1_kernel(const double *__restrict__ part, int nblocks, uint64_t total, double *__restrict__ stats) {
  double v[5] = {0.0, INFINITY, -INFINITY, INFINITY, -INFINITY};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const double *p = part + (size_t)b * 5;
    v[0] += p[0];
    v[1] = fmin(v[1], p[1]);
    v[2] = fmax(v[2], p[2]);
    v[3] = fmin(v[3], p[3]);
    v[4] = fmax(v[4], p[4]);
  }
  const int op[5] = {0, 1, 2, 1, 2};
  __shared__ double out[5];
  block_reduce_store<5>(v, op, out);
  if (threadIdx.x == 0) {
    stats[0] = out[0] / (double)total;
    stats[2] = out[1];
    stats[3] = out[2];
    stats[4] = out[3];
    stats[5] = out[4];
    stats[6] = 0.0;
    stats[7] = 0.0;
  }
}

__global__ void __launch_bounds__(kMomThreads)
moments_pass2_kernel(const double *__restrict__ Y, uint64_t total, const double *__restrict__ stats,
                     double *__restrict__ part) {
  const double mean = stats[0];
  const uint64_t S = (uint64_t)gridDim.x * blockDim.x;
  double v[1] = {0.0};
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += S) {
    const double d = Y[idx] - mean;
    v[0] = fma(d, d, v[0]);
  }
  const int op[1] = {0};
  block_reduce_store<1>(v, op, part + blockIdx.x);
}

__global__ void __launch_bounds__(kMomThreads)
moments_final2_kernel(const double *__restrict__ part, int nblocks, uint64_t total, double *__restrict__ stats) {
  double v[1] = {0.0};
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v[0] += part[b];
  const int op[1] = {0};
  __shared__ double out[1];
  block_reduce_store<1>(v, op, out);
  if (threadIdx.x == 0) stats[1] = sqrt(out[0] / (double)total);
}

// ------------------------------------------------------------------------------------------
// normalise + re-layout
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const double *__restrict__ Y, uint64_t N, int Dg, int step, int second_order,
                      const double *__restrict__ stats, double *__restrict__ Yn) {
  const double mean = stats[0], sd = stats[1];
  const int PL = part_len(Dg);
  const int stride = row_stride(Dg, second_order);
  const uint64_t total = N * (uint64_t)stride;
  const uint64_t S = (uint64_t)gridDim.x * blockDim.x;
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = idx / (uint64_t)stride;
  int p = (int)(idx - i * (uint64_t)stride);
  const uint64_t di = S / (uint64_t)stride;
  const int dp = (int)(S - di * (uint64_t)stride);
  for (; idx < total; idx += S) {
    int src = -1;
    const int nparts = second_order ? 2 : 1;
    if (p < nparts * PL) {
      const int part = (p >= PL) ? 1 : 0;
      const int q = p - part * PL;
      const int blk = q / kBlockPitch, e = q - blk * kBlockPitch;
      const int j = 8 * blk + e;
      if (e < 8 && j < Dg) src = 1 + part * Dg + j;
    } else {
      const int q = p - nparts * PL;
      if (q == 0) src = 0;
      else if (q == 1) src = step - 1;
    }
    double v = 0.0;
    if (src >= 0) v = __ddiv_rn(__dsub_rn(Y[i * (uint64_t)step + src], mean), sd);
    Yn[idx] = v;
    i += di;
    p += dp;
    if (p >= stride) {
      p -= stride;
      ++i;
    }
  }
}

// Column-major copy: Yt[col][Npad], cols = A, B, AB_0..AB_{Dg-1}; rows >= N are zero.
__global__ void __launch_bounds__(256)
normalize_cols_kernel(const double *__restrict__ Y, uint64_t N, uint64_t Npad, int Dg, int step,
                      const double *__restrict__ stats, double *__restrict__ Yt) {
  const double mean = stats[0], sd = stats[1];
  const int col = blockIdx.y;
  const int src = (col == 0) ? 0 : (col == 1 ? step - 1 : col - 1);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Npad;
       i += (uint64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (i < N) v = __ddiv_rn(__dsub_rn(Y[i * (uint64_t)step + src], mean), sd);
    Yt[(uint64_t)col * Npad + i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// resample multiplicities
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
counts_from_indices_kernel(const int64_t *__restrict__ r, uint64_t N, uint64_t R, uint64_t c0,
                           uint64_t count, uint64_t row0, uint64_t nrows, uint64_t Npad,
                           uint32_t *__restrict__ w32, int32_t *overflow) {
  const uint64_t total = N * count;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = t / count;
    const uint64_t cl = t - i * count;
    const int64_t idx = r[i * R + c0 + cl];
    if (idx < 0 || (uint64_t)idx >= N) {
      if (overflow) atomicOr(overflow, 2);
      continue;
    }
    const uint64_t local = (uint64_t)idx - row0;  // rows outside [row0, row0+nrows) belong to other shards
    if (local >= nrows) continue;
    const uint64_t pos = cl * Npad + local;
    const unsigned sh = 8u * (unsigned)(pos & 3);
    const uint32_t old = atomicAdd(w32 + (pos >> 2), 1u << sh);
    if (((old >> sh) & 0xffu) == 0xffu && overflow) atomicOr(overflow, 1);
  }
}

// Draw t of resample c: u64 = philox(counter = (t/2, c), key = seed) word pair (t & 1);
// index = floor(u64 * N / 2^64).
__global__ void __launch_bounds__(256)
counts_philox_kernel(uint64_t seed, uint64_t N, uint64_t c0, uint64_t row0, uint64_t nrows, uint64_t Npad,
                     uint32_t *__restrict__ w32) {
  const uint64_t c = c0 + blockIdx.y;
  uint32_t *wrow = w32 + ((uint64_t)blockIdx.y * Npad >> 2);
  const uint64_t npairs = (N + 1) >> 1;
  for (uint64_t t2 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t2 < npairs;
       t2 += (uint64_t)gridDim.x * blockDim.x) {
    const Philox4 p = philox4x32_10((uint32_t)t2, (uint32_t)(t2 >> 32), (uint32_t)c, (uint32_t)(c >> 32),
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t u0 = ((uint64_t)p.y << 32) | p.x;
    const uint64_t u1 = ((uint64_t)p.w << 32) | p.z;
    const uint64_t i0 = __umul64hi(u0, N) - row0;
    if (i0 < nrows) atomicAdd(wrow + (i0 >> 2), 1u << (8u * (unsigned)(i0 & 3)));
    if (2 * t2 + 1 < N) {
      const uint64_t i1 = __umul64hi(u1, N) - row0;
      if (i1 < nrows) atomicAdd(wrow + (i1 >> 2), 1u << (8u * (unsigned)(i1 & 3)));
    }
  }
}

// Same distribution without global atomics (the kernel above is bound by the L2 RED rate, 2.2e11/s): the
// binomial split tree + window histograms of counts.cuh, one block per resample.
// Two launch shapes: "wide" -- 256 threads, one block per resample, up to 8 blocks per SM (fastest when the GPU is
// otherwise idle) -- and "slim" -- one block of 512 threads per SM looping over resamples with a single 16 KB
// window: it always fits NEXT TO a CTA of the persistent GEMM kernel (198 KB of shared memory, 192 threads), so
// the next batch's multiplicities are drawn on a side stream while the tensor cores work on the current batch.
constexpr int kWinThreads = 256;
constexpr int kWinThreadsSlim = 512;

template <int NT>
__global__ void __launch_bounds__(NT)
counts_window_kernel(uint64_t seed, uint64_t N, uint64_t c0, uint64_t count, int L, uint64_t row0, uint64_t nrows,
                     uint64_t Npad, uint8_t *__restrict__ w) {
  extern __shared__ __align__(16) uint32_t wsm[];
  for (uint64_t i = blockIdx.x; i < count; i += gridDim.x)
    counts_window_resample<NT, 0>(wsm, seed, c0 + i, N, L, row0, nrows, w + i * Npad, (int)threadIdx.x);
}

// ------------------------------------------------------------------------------------------
// second-order accumulate (production kernel for Dg <= 64): warp <-> resample,
// lane <-> 8x8 (j,k) register tile, rows staged in shared memory
// ------------------------------------------------------------------------------------------
// For resample c the D(D-1)/2 sums  S_jk = sum_i w_i * BA_j[i] * AB_k[i]  (j < k) are a weighted
// rank-1 update per row.  One warp owns one resample; the (j,k) triangle is cut into 8x8 tiles:
// 28 off-diagonal tiles + 4 diagonal tiles are full register tiles on the 32 lanes (64 DFMA per row
// fed by 16 operand doubles), the other 4 diagonal triangles (112 pairs) are spread 4 per lane.
// First/total-order sums ride along (2 columns per lane).  Rows with multiplicity 0 are skipped.
//
// Data path: tiles of 32 contiguous rows (one cp.async.bulk each, 41 KB at Dg = 64) go through a
// 5-stage shared-memory ring with one mbarrier per stage; the 8 warps of a block (8 resamples)
// consume the same tiles at their own pace and the last warp to leave a stage re-arms it and issues
// the copy of tile t+5 -- no block-wide barrier in the loop.  The analysis layout puts the 8-column
// blocks at an 80-byte pitch (col_pos), so the chunks that different lanes read with the SAME
// instruction offset fall into different bank groups: all operand loads are conflict-free LDS.128
// from one base register per operand set.
//
// Software pipeline: the operands of row r+1 (8 LDS.128) are fetched while the 64 DFMAs of row r are
// issued -- the BA chunks into the registers the update has just finished with, the AB values into
// a second buffer -- so only the first row of a tile waits for shared memory.
// ncu history of this kernel is in profiles/ (L1 version: 118 wavefronts/row, FP64 pipe 38 %).
constexpr int kS2Warps = 8;

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// The 28 pairs (jj < kk) of an 8x8 strict upper triangle on 8 lanes, 4 slots per lane: slots 0,1 are
// rows (ra, ra+1) of column ca, slots 2,3 rows (rb, rb+1) of column cb.  Lanes 0..5 carry two full
// row pairs, lanes 6,7 carry the four left-over single pairs (their slots 1 and 3 are unused):
//   g:      0        1        2        3        4        5        6          7
//   ca/ra:  7/0      7/2      7/4      5/0      5/2      3/0      7/6 (one)  3/2 (one)
//   cb/rb:  6/0      6/2      6/4      4/0      4/2      2/0      5/4 (one)  1/0 (one)
// Every lane uses column ca for slots 0,1 and cb for slots 2,3: no per-slot operand selection, and
// each row pair is one aligned LDS.128.
__device__ __forceinline__ void diag_group(int g, int &ca, int &ra, int &cb, int &rb, bool &single) {
  const int sh = 4 * g;  // one nibble per group, group 0 in the low nibble
  ca = (0x37355777u >> sh) & 15;
  ra = (0x26020420u >> sh) & 15;
  cb = (0x15244666u >> sh) & 15;
  rb = (0x04020420u >> sh) & 15;
  single = g >= 6;
}

// sum_i w*A, w*B, w*A^2, w*B^2, w*A*B.  Kept out of the tile kernel (10 registers and 5 DFMA per row
// there).  One block per 4 resamples and row split; its 256 threads stride over the rows of the
// column-major A/B copy (coalesced, L2 resident), two rows in flight per thread; fixed-order
// block reduction, so the result is deterministic.
constexpr int kScC = 4;
__global__ void __launch_bounds__(256)
scalar_sums_kernel(const double *__restrict__ Yt, uint64_t Npad, uint64_t N, const uint8_t *__restrict__ w,
                   uint64_t count, uint64_t rows_per_split, double *__restrict__ psum, int nacc) {
  __shared__ double red[8][kScC * 5];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t cbase = (uint64_t)blockIdx.x * kScC;
  const int sp = blockIdx.y;
  const uint64_t row0 = (uint64_t)sp * rows_per_split;
  uint64_t row1 = row0 + rows_per_split;
  if (row1 > N) row1 = N;
  double sc[kScC][5];
#pragma unroll
  for (int c = 0; c < kScC; ++c)
#pragma unroll
    for (int q = 0; q < 5; ++q) sc[c][q] = 0.0;
  const double *colA = Yt, *colB = Yt + Npad;
  auto add_row = [&](uint64_t i) {
    const double a = colA[i], b = colB[i];
    const double aa = a * a, bb = b * b, ab = a * b;
#pragma unroll
    for (int c = 0; c < kScC; ++c) {
      double wv = 1.0;
      if (w) wv = (cbase + c < count) ? (double)w[(cbase + c) * Npad + i] : 0.0;
      sc[c][0] = fma(wv, a, sc[c][0]);
      sc[c][1] = fma(wv, b, sc[c][1]);
      sc[c][2] = fma(wv, aa, sc[c][2]);
      sc[c][3] = fma(wv, bb, sc[c][3]);
      sc[c][4] = fma(wv, ab, sc[c][4]);
    }
  };
  uint64_t i = row0 + threadIdx.x;
  for (; i + 256 < row1; i += 512) {
    add_row(i);
    add_row(i + 256);
  }
  if (i < row1) add_row(i);
#pragma unroll
  for (int c = 0; c < kScC; ++c)
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double v = warp_sum(sc[c][q]);
      if (lane == 0) red[warp][c * 5 + q] = v;
    }
  __syncthreads();
  if (threadIdx.x < kScC * 5) {
    const int c = threadIdx.x / 5, q = threadIdx.x % 5;
    if (cbase + c < count) {
      double v = 0.0;
      for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
      psum[((uint64_t)sp * count + cbase + c) * (uint64_t)nacc + q] = v;
    }
  }
}

template <int kTileRows, int kStagesT, bool GROUPED>  // kStagesT == 0: ring depth given at run time
__global__ void __launch_bounds__(kS2Warps * 32, 1)
s2_accum_smem_kernel(const double *__restrict__ Yn, int stride, int Dp, uint64_t N, int Dg,
                     const uint8_t *__restrict__ w, uint64_t Npad, uint64_t count,
                     uint64_t rows_per_split, double *__restrict__ psum, int nacc, int stages_rt) {
  const int kStages = kStagesT ? kStagesT : stages_rt;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t c = (uint64_t)blockIdx.x * kS2Warps + warp;
  const bool live = c < count;
  const int sp = blockIdx.y;
  const uint64_t row0 = (uint64_t)sp * rows_per_split;
  uint64_t row1 = row0 + rows_per_split;
  if (row1 > N) row1 = N;
  const int ntiles = (row1 > row0) ? (int)((row1 - row0 + kTileRows - 1) / kTileRows) : 0;
  const int TB = Dp >> 3;
  const uint32_t row_bytes = (uint32_t)stride * 8u;
  const uint32_t stage_bytes = kTileRows * row_bytes;
  const uint32_t sm_base = smem_u32(smraw);
  const uint32_t bar_base = sm_base + kStages * stage_bytes;  // kStages x u64
  uint32_t *cnt = reinterpret_cast<uint32_t *>(smraw + kStages * stage_bytes + kStages * 8);

  auto issue = [&](int t) {  // one elected thread: arm the barrier and start the bulk copy of tile t
    const int s = t % kStages;
    const uint64_t r = row0 + (uint64_t)t * kTileRows;
    const uint32_t rows = (uint32_t)((row1 - r < (uint64_t)kTileRows) ? (row1 - r) : kTileRows);
    const uint32_t bytes = rows * row_bytes;
    mbar_arrive_expect_tx(bar_base + 8u * s, bytes);
    bulk_g2s(sm_base + s * stage_bytes, Yn + r * (uint64_t)stride, bytes, bar_base + 8u * s);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_base + 8u * s, 1);
      cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);
  }

  // ---- lane roles ------------------------------------------------------------------------
  // Tile list: off-diagonal tiles (tj < tk) in row-major order, then the diagonal tiles.  Up to Dg = 64
  // (TB <= 8) one warp carries a resample: 32 full tiles, the diagonal tiles that do not fit are spread
  // over the lanes 4 pairs each (below).  For larger Dg the list is cut into groups of 32 tiles and
  // blockIdx.z selects the group: a resample is carried by ceil(TB(TB+1)/64) warps in different blocks.
  constexpr bool grouped = GROUPED;  // host: TB > 8
  constexpr bool W1FAST = false;     // weight-1 fast path: measured no gain (shared-memory bound)
  const int group = grouped ? blockIdx.z : 0;
  const int ntOff = TB * (TB - 1) / 2;
  const int ndiag_full = grouped ? TB : ((32 - ntOff < TB) ? (32 - ntOff) : TB);  // diagonal tiles done as full tiles
  const int tile_id = group * 32 + lane;
  int tj = 0, tk = 0;
  bool actA = false;
  if (tile_id < ntOff) {
    int idx = tile_id;
    for (int t = 0; t < TB; ++t) {
      const int n = TB - 1 - t;
      if (idx < n) {
        tj = t;
        tk = t + 1 + idx;
        actA = true;
        break;
      }
      idx -= n;
    }
  } else if (tile_id - ntOff < ndiag_full) {
    tj = tk = tile_id - ntOff;
    actA = true;
  }
  const int PL = TB * kBlockPitch;                                              // doubles per AB / BA part
  const uint32_t offX = actA ? (uint32_t)(PL + kBlockPitch * tj) * 8u : 0u;  // BA block tj
  const uint32_t offY = actA ? (uint32_t)(kBlockPitch * tk) * 8u : 0u;       // AB block tk
  // remaining diagonal tiles (only when TB == 8: tiles 4..7): 8 lanes per tile, see diag_group
  const bool hasB = ndiag_full < TB;
  const int dB = ndiag_full + (lane >> 3);
  const int gB = lane & 7;
  const bool actB = hasB && dB < TB;
  int bca, bra, bcb, brb;
  bool bsingle;
  diag_group(gB, bca, bra, bcb, brb, bsingle);
  const uint32_t offBxa = actB ? (uint32_t)(PL + col_pos(8 * dB + bra)) * 8u : 0u;  // rows ra, ra+1
  const uint32_t offBxb = actB ? (uint32_t)(PL + col_pos(8 * dB + brb)) * 8u : 0u;  // rows rb, rb+1
  const uint32_t offBya = actB ? (uint32_t)col_pos(8 * dB + bca) * 8u : 0u;
  const uint32_t offByb = actB ? (uint32_t)col_pos(8 * dB + bcb) * 8u : 0u;
  // first/total order: two columns per lane; block order {0,4,1,5} / {2,6,3,7} over the quarter-warps
  // keeps the two blocks a half-warp reads with one LDS.64 in disjoint banks
  const int cblk0 = ((lane >> 3) & 1) * 4 + (lane >> 4);
  const int jC0 = 64 * group + 8 * cblk0 + (lane & 7), jC1 = jC0 + 16;  // group g: columns 64g .. 64g+63
  const uint32_t offC0 = (jC0 < Dp) ? (uint32_t)col_pos(jC0) * 8u : 0u;
  const uint32_t offC1 = (jC1 < Dp) ? (uint32_t)col_pos(jC1) * 8u : 0u;
  const uint32_t offAB = 2u * PL * 8u;

  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
  double accB[4] = {0.0, 0.0, 0.0, 0.0};
  double s1_0 = 0.0, s1_1 = 0.0, st_0 = 0.0, st_1 = 0.0;

  const uint8_t *wrow = (w && live) ? w + c * Npad : nullptr;
  auto load_w = [&](int t) -> unsigned {
    const uint64_t i = row0 + (uint64_t)t * kTileRows + lane;
    if (!live || t >= ntiles || lane >= kTileRows || i >= row1) return 0u;
    return wrow ? (unsigned)__ldg(wrow + i) : 1u;
  };
  unsigned wl_next = load_w(0);

  // Operand registers: P = the 8 BA values of the lane's tile, Q0/Q1 = the 8 AB values of the current
  // and of the next row (roles swap every row).  While the 64 DFMAs of row r are issued, the 8
  // LDS.128 of row r+1 are already in flight: P chunks are refilled as soon as their 16 DFMAs are
  // out, the other AB buffer is free.  Only the first row of a tile waits for its operands.
  double P[8], Q0[8], Q1[8];
  auto ldP = [&](uint32_t row, int q) {
    const double2 v = lds_f64x2(row + offX + 16u * q);
    P[2 * q] = v.x;
    P[2 * q + 1] = v.y;
  };
  auto ldQ = [&](double(&Q)[8], uint32_t row, int q) {
    const double2 v = lds_f64x2(row + offY + 16u * q);
    Q[2 * q] = v.x;
    Q[2 * q + 1] = v.y;
  };
  // one row: rank-1 update with (P, Qc); prefetch the next row into (P, Qn); diagonal leftovers and
  // first/total order ("side work") are loaded after the second chunk and consumed at the end
  auto do_row = [&](uint32_t row, uint32_t nrow, double wv, double(&Qc)[8], double(&Qn)[8], auto scale_tag) {
    constexpr bool SCALE = decltype(scale_tag)::value;  // false: weight 1, no multiplications by it
    double ya = 0.0, yb = 0.0, t0, t1;
    double2 xa = make_double2(0.0, 0.0), xb = xa, ab;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = 2 * q + e;
        const double wx = SCALE ? wv * P[a] : P[a];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = fma(wx, Qc[k], acc[a][k]);
      }
      ldP(nrow, q);
      ldQ(Qn, nrow, q);
      if (q == 1) {  // side loads for THIS row
        if (hasB) {
          ya = lds_f64(row + offBya);
          yb = lds_f64(row + offByb);
          xa = lds_f64x2(row + offBxa);
          xb = lds_f64x2(row + offBxb);
        }
        ab = lds_f64x2(row + offAB);
        t0 = lds_f64(row + offC0);
        t1 = lds_f64(row + offC1);
      }
    }
    if (hasB) {  // warp-uniform
      if (SCALE) {
        ya *= wv;
        yb *= wv;
      }
      accB[0] = fma(xa.x, ya, accB[0]);
      accB[1] = fma(xa.y, ya, accB[1]);
      accB[2] = fma(xb.x, yb, accB[2]);
      accB[3] = fma(xb.y, yb, accB[3]);
    }
    const double wb = SCALE ? wv * ab.y : ab.y;
    t0 -= ab.x;
    t1 -= ab.x;
    s1_0 = fma(wb, t0, s1_0);
    s1_1 = fma(wb, t1, s1_1);
    st_0 = fma(SCALE ? wv * t0 : t0, t0, st_0);
    st_1 = fma(SCALE ? wv * t1 : t1, t1, st_1);
  };
  auto row_any = [&](uint32_t row, uint32_t nrow, int hi, double(&Qc)[8], double(&Qn)[8]) {
    if (W1FAST && hi == 0x3FF00000) {  // multiplicity 1 (58 % of the processed rows)
      do_row(row, nrow, 1.0, Qc, Qn, std::false_type{});
    } else {
      do_row(row, nrow, __hiloint2double(hi, 0), Qc, Qn, std::true_type{});
    }
  };

  // Row stream.  The operands of the upcoming row always sit in (P, Q[phase]); the last row of a tile
  // prefetches the first row of the NEXT tile (whose barrier is waited for just before), so a warp meets a
  // cold start only at its very first row (or after a tile in which its resample has no row at all).
  int phase = 0;        // which AB buffer holds the upcoming row: 0 -> Q0, 1 -> Q1
  bool have = false;    // operands of the upcoming row already loaded
  bool waited = false;  // barrier of the upcoming tile already passed
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kStages;
    const unsigned wl = wl_next;
    wl_next = load_w(t + 1);
    unsigned mask = __ballot_sync(kFull, wl != 0);
    // multiplicities are small integers: as doubles their low word is 0, so one 32-bit shuffle of
    // the high word moves them between lanes and no int->double conversion sits in the row loop
    const int whi = __double2hiint((double)wl);
    if (!waited) {
      const uint32_t bar = bar_base + 8u * s;
      const uint32_t parity = (uint32_t)(t / kStages) & 1u;
      while (!mbar_try_wait(bar, parity)) {
      }
    }
    waited = false;
    const uint32_t tile = sm_base + s * stage_bytes;
    if (mask && !have) {  // cold start
      const uint32_t row = tile + (uint32_t)(__ffs(mask) - 1) * row_bytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ldP(row, q);
        if (phase) ldQ(Q1, row, q);
        else ldQ(Q0, row, q);
      }
    }
    if (mask) have = false;
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const int hi = __shfl_sync(kFull, whi, b);
      const uint32_t row = tile + (uint32_t)b * row_bytes;
      uint32_t nrow = row;  // nothing to prefetch: reload this row (harmless)
      if (mask) {
        nrow = tile + (uint32_t)(__ffs(mask) - 1) * row_bytes;
      } else if (t + 1 < ntiles) {  // last row of the tile: chain into the next tile
        const unsigned mask_n = __ballot_sync(kFull, wl_next != 0);
        if (mask_n) {
          const int sn = (t + 1) % kStages;
          const uint32_t bar = bar_base + 8u * sn;
          const uint32_t parity = (uint32_t)((t + 1) / kStages) & 1u;
          while (!mbar_try_wait(bar, parity)) {
          }
          waited = true;
          have = true;
          nrow = sm_base + sn * stage_bytes + (uint32_t)(__ffs(mask_n) - 1) * row_bytes;
        }
      }
      if (phase) row_any(row, nrow, hi, Q1, Q0);
      else row_any(row, nrow, hi, Q0, Q1);
      phase ^= 1;
    }
    // release the stage; the last warp out refills it with tile t + kStages
    __syncwarp();
    if (lane == 0) {
      uint32_t old;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "r"(smem_u32(cnt + s))
                   : "memory");
      if (old == kS2Warps - 1) {
        cnt[s] = 0;
        if (t + kStages < ntiles) issue(t + kStages);
      }
    }
  }

  if (!live) return;
  double *out = psum + ((uint64_t)sp * count + c) * (uint64_t)nacc;
  if (jC0 < Dg) {
    out[5 + jC0] = s1_0;
    out[5 + Dg + jC0] = st_0;
  }
  if (jC1 < Dg) {
    out[5 + jC1] = s1_1;
    out[5 + Dg + jC1] = st_1;
  }
  double *o2 = out + 5 + 2 * Dg;
  if (actA) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int j = 8 * tj + a;
        const int kk = 8 * tk + k;
        if (j < kk && kk < Dg) o2[pair_index(j, kk, Dg)] = acc[a][k];
      }
  }
  if (actB) {
    const int base = 8 * dB;
    if (base + bca < Dg) {
      o2[pair_index(base + bra, base + bca, Dg)] = accB[0];
      if (!bsingle) o2[pair_index(base + bra + 1, base + bca, Dg)] = accB[1];
    }
    if (base + bcb < Dg) {
      o2[pair_index(base + brb, base + bcb, Dg)] = accB[2];
      if (!bsingle) o2[pair_index(base + brb + 1, base + bcb, Dg)] = accB[3];
    }
  }
}

// ------------------------------------------------------------------------------------------
// EXPERIMENTAL, opt-in (algo 5): second-order accumulate on the FP64 tensor cores
// ------------------------------------------------------------------------------------------
// BASELINE's north_star rules tensor cores out for this path, so the shipped default is the DFMA tile
// kernel above.  This variant exists to quantify what that rule costs: the weighted Gram matrix
// G = BA^T diag(w) AB is fed to mma.sync.m8n8k4.f64 (DMMA, 37.2 TFLOP/s measured, the same peak as the
// FMA pipe) four nonzero rows at a time.  A warp still owns one resample; per K-step a lane loads 8 + 8
// fragment doubles for 36 tile MMAs, i.e. 4x less shared-memory traffic per flop than the register tiles.
// Same ring, weights, first/total-order side work and output layout as s2_accum_smem_kernel.  Dg <= 64.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int kStages>
__global__ void __launch_bounds__(kS2Warps * 32, 1)
s2_accum_dmma_kernel(const double *__restrict__ Yn, int stride, int Dp, uint64_t N, int Dg,
                     const uint8_t *__restrict__ w, uint64_t Npad, uint64_t count,
                     uint64_t rows_per_split, double *__restrict__ psum, int nacc) {
  constexpr int kTileRows = 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint64_t c = (uint64_t)blockIdx.x * kS2Warps + warp;
  const bool live = c < count;
  const int sp = blockIdx.y;
  const uint64_t row0 = (uint64_t)sp * rows_per_split;
  uint64_t row1 = row0 + rows_per_split;
  if (row1 > N) row1 = N;
  const int ntiles = (row1 > row0) ? (int)((row1 - row0 + kTileRows - 1) / kTileRows) : 0;
  const int TB = Dp >> 3;
  const uint32_t row_bytes = (uint32_t)stride * 8u;
  const uint32_t stage_bytes = kTileRows * row_bytes;
  const uint32_t sm_base = smem_u32(smraw);
  const uint32_t bar_base = sm_base + kStages * stage_bytes;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(smraw + kStages * stage_bytes + kStages * 8);
  auto issue = [&](int t) {
    const int s = t % kStages;
    const uint64_t r = row0 + (uint64_t)t * kTileRows;
    const uint32_t rows = (uint32_t)((row1 - r < (uint64_t)kTileRows) ? (row1 - r) : kTileRows);
    mbar_arrive_expect_tx(bar_base + 8u * s, rows * row_bytes);
    bulk_g2s(sm_base + s * stage_bytes, Yn + r * (uint64_t)stride, rows * row_bytes, bar_base + 8u * s);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_base + 8u * s, 1);
      cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);

  const int PL = TB * kBlockPitch;
  const int gid = lane >> 2, kq = lane & 3;  // fragment coordinates: M/N index and K index of this lane
  const uint32_t offA = (uint32_t)(PL + gid) * 8u;  // + 80 B per j-block
  const uint32_t offB = (uint32_t)gid * 8u;         // + 80 B per k-block
  const int cblk0 = ((lane >> 3) & 1) * 4 + (lane >> 4);
  const int jC0 = 8 * cblk0 + (lane & 7), jC1 = jC0 + 16;
  const uint32_t offC0 = (jC0 < Dp) ? (uint32_t)col_pos(jC0) * 8u : 0u;
  const uint32_t offC1 = (jC1 < Dp) ? (uint32_t)col_pos(jC1) * 8u : 0u;
  const uint32_t offAB = 2u * PL * 8u;

  double acc[36][2];
#pragma unroll
  for (int t = 0; t < 36; ++t) acc[t][0] = acc[t][1] = 0.0;
  double s1_0 = 0.0, s1_1 = 0.0, st_0 = 0.0, st_1 = 0.0;

  const uint8_t *wrow = (w && live) ? w + c * Npad : nullptr;
  auto load_w = [&](int t) -> unsigned {
    const uint64_t i = row0 + (uint64_t)t * kTileRows + lane;
    if (!live || t >= ntiles || i >= row1) return 0u;
    return wrow ? (unsigned)__ldg(wrow + i) : 1u;
  };
  auto release = [&](int t) {  // this warp is done with tile t; the last warp out refills its stage
    __syncwarp();
    if (lane == 0) {
      const int s = t % kStages;
      uint32_t old;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "r"(smem_u32(cnt + s))
                   : "memory");
      if (old == kS2Warps - 1) {
        cnt[s] = 0;
        if (t + kStages < ntiles) issue(t + kStages);
      }
    }
  };

  // Row stream state: current tile t, the not yet consumed nonzero rows of it (mask) and the weights (whi).
  int t = 0;
  unsigned wl = load_w(0), wl_next = load_w(1);
  unsigned mask = __ballot_sync(kFull, wl != 0);
  int whi = __double2hiint((double)wl);
  uint32_t tile = sm_base;
  bool tile_ready = false;
  int done_tiles = 0;  // tiles [0, done_tiles) have been released
  // A K-step = the next (up to) four nonzero rows; rows never mix tiles.
  struct Step {
    uint32_t rowaddr[4];
    double wv[4];
    uint32_t myrow;  // row of this lane's K slot
    double wq;       // its weight (0 for padding)
    int tile_id;
    bool valid;
  };
  auto pick = [&](Step &st) {
    while (mask == 0 && t + 1 < ntiles) {  // advance to the next tile that has rows of this resample
      ++t;
      wl = wl_next;
      wl_next = load_w(t + 1);
      mask = __ballot_sync(kFull, wl != 0);
      whi = __double2hiint((double)wl);
      tile = sm_base + (t % kStages) * stage_bytes;
      tile_ready = false;
    }
    st.valid = mask != 0;
    st.tile_id = t;
    if (!st.valid) return;
    if (!tile_ready) {
      while (!mbar_try_wait(bar_base + 8u * (t % kStages), (uint32_t)(t / kStages) & 1u)) {
      }
      tile_ready = true;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = mask ? __ffs(mask) - 1 : 0;
      st.rowaddr[q] = tile + (uint32_t)r * row_bytes;
      st.wv[q] = mask ? __hiloint2double(__shfl_sync(kFull, whi, r), 0) : 0.0;
      if (mask) mask &= mask - 1;
    }
    st.myrow = (kq == 0) ? st.rowaddr[0] : (kq == 1 ? st.rowaddr[1] : (kq == 2 ? st.rowaddr[2] : st.rowaddr[3]));
    st.wq = (kq == 0) ? st.wv[0] : (kq == 1 ? st.wv[1] : (kq == 2 ? st.wv[2] : st.wv[3]));
  };

  double fa[8], fb[8];  // fragments of the current step; refilled in place for the next step as the MMAs retire them
  Step cur, nxt;
  pick(cur);
  if (cur.valid) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      fa[q] = (q < TB) ? lds_f64(cur.myrow + offA + 80u * q) : 0.0;
      fb[q] = (q < TB) ? lds_f64(cur.myrow + offB + 80u * q) : 0.0;
    }
  }
  while (cur.valid) {
#pragma unroll
    for (int q = 0; q < 8; ++q) fa[q] *= cur.wq;  // fragments were loaded unscaled
    pick(nxt);
    double2 sab = make_double2(0.0, 0.0);
    double sc0 = 0.0, sc1 = 0.0;
    int ti = 0;
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
#pragma unroll
      for (int kb = jb; kb < 8; ++kb) {
        if (kb < TB) dmma884(acc[ti], fa[jb], fb[kb]);  // uniform
        ++ti;
      }
      if (nxt.valid && jb < TB) {  // block jb is retired: its registers take the next step's fragments
        fa[jb] = lds_f64(nxt.myrow + offA + 80u * jb);
        fb[jb] = lds_f64(nxt.myrow + offB + 80u * jb);
      }
      // first/total order of the step's rows: row q is loaded after MMA group q and consumed after group q+1
      if (jb >= 1 && jb <= 4) {
        const double wv = cur.wv[jb - 1];
        const double wb = wv * sab.y;
        const double t0 = sc0 - sab.x, t1 = sc1 - sab.x;
        s1_0 = fma(wb, t0, s1_0);
        s1_1 = fma(wb, t1, s1_1);
        st_0 = fma(wv * t0, t0, st_0);
        st_1 = fma(wv * t1, t1, st_1);
      }
      if (jb <= 3) {
        sab = lds_f64x2(cur.rowaddr[jb] + offAB);
        sc0 = lds_f64(cur.rowaddr[jb] + offC0);
        sc1 = lds_f64(cur.rowaddr[jb] + offC1);
      }
    }
    // tiles left behind by this step can go back to the ring
    const int upto = nxt.valid ? nxt.tile_id : ntiles;
    while (done_tiles < upto) release(done_tiles++);
    cur = nxt;
  }
  while (done_tiles < ntiles) release(done_tiles++);

  if (!live) return;
  double *out = psum + ((uint64_t)sp * count + c) * (uint64_t)nacc;
  if (jC0 < Dg) {
    out[5 + jC0] = s1_0;
    out[5 + Dg + jC0] = st_0;
  }
  if (jC1 < Dg) {
    out[5 + jC1] = s1_1;
    out[5 + Dg + jC1] = st_1;
  }
  double *o2 = out + 5 + 2 * Dg;
  int ti = 0;
#pragma unroll
  for (int jb = 0; jb < 8; ++jb)
#pragma unroll
    for (int kb = jb; kb < 8; ++kb) {
      // C fragment: row = lane / 4, columns 2 * (lane % 4) and + 1
      const int j = 8 * jb + gid;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * kb + 2 * kq + e;
        if (kb < TB && j < k && k < Dg) o2[pair_index(j, k, Dg)] = acc[ti][e];
      }
      ++ti;
    }
}

// ------------------------------------------------------------------------------------------
// first/total-order accumulate: a skinny GEMM  acc[c][q] += w[c][i] * F[i][q]
// ------------------------------------------------------------------------------------------
// Every first/total-order sum is linear in the multiplicities: with the per-row features
//   F[i] = [ A, B, A^2, B^2, A*B,  B*(AB_j - A) (j < Dg),  (A - AB_j)^2 (j < Dg) ]
// (written once by normalize_features_kernel) the SA_NACC sums of resample c are sum_i w_c[i] * F[i][q],
// in exactly the order of the psum layout.  Lanes <-> resamples: a warp owns 32 resamples, every lane
// keeps its NF (<= 32 per grid.z chunk) accumulators in registers, the feature row is warp-uniform
// (one broadcast load per 16 bytes for 32 resamples) and the only per-lane data is one weight byte per
// row, read 32 rows at a time.  One DFMA per (resample, row, feature): 5 + 2*Dg flop-pairs instead of
// the 7 + 8*Dg operations of evaluating the estimator terms per resample.
constexpr int kFoWarps = 8;
#ifndef SA_MULTI_CHUNK
#define SA_MULTI_CHUNK 16
#endif
constexpr int kMultiChunk = SA_MULTI_CHUNK;  // chunk width when the features do not fit one chunk of <= 24
constexpr int kFoRows = 32;   // rows per weight fetch (32 bytes per resample)

// Feature matrix layout: chunks of CW features, F[chunk][row][CW] (each chunk is what one grid.z slice of
// fo_gemv_kernel streams).  CW = 8/16/24 when all 5 + 2*Dg features fit one chunk, else 32.
__host__ __device__ inline int chunk_width(int nfeat) { return nfeat <= 24 ? ((nfeat + 7) & ~7) : kMultiChunk; }
__host__ __device__ inline int chunk_count(int nfeat) {
  const int cw = chunk_width(nfeat);
  return (nfeat + cw - 1) / cw;
}
__host__ __device__ inline int feature_width(int Dg) { return chunk_width(5 + 2 * Dg); }
__host__ __device__ inline int feature_chunks(int Dg) { return chunk_count(5 + 2 * Dg); }

// second_order: the row also carries BA_j (columns 1+Dg+j) and the features continue with the Dg(Dg-1)/2 pair
// products BA_j * AB_k (j < k, row-major over the strict upper triangle) -- SA_NACC order throughout, so the
// GEMV output is exactly the accumulator vector finalize_kernel expects.
__global__ void __launch_bounds__(256)
normalize_features_kernel(const double *__restrict__ Y, uint64_t N, int Dg, int step, int second_order,
                          const double *__restrict__ stats, double *__restrict__ F) {
  const double mean = stats[0], sd = stats[1];
  const int nfeat = 5 + 2 * Dg + (second_order ? Dg * (Dg - 1) / 2 : 0);
  const int CW = chunk_width(nfeat), NZ = chunk_count(nfeat);
  const uint64_t per_chunk = N * (uint64_t)CW;
  const uint64_t total = per_chunk * NZ;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const int z = (int)(idx / per_chunk);
    const uint64_t rem = idx - (uint64_t)z * per_chunk;
    const uint64_t i = rem / (uint64_t)CW;
    const int q = z * CW + (int)(rem - i * (uint64_t)CW);  // global feature index
    const double *row = Y + i * (uint64_t)step;
    double v = 0.0;
    if (q < 5 + 2 * Dg) {
      const double a = __ddiv_rn(__dsub_rn(row[0], mean), sd);
      const double b = __ddiv_rn(__dsub_rn(row[step - 1], mean), sd);
      if (q == 0) v = a;
      else if (q == 1) v = b;
      else if (q == 2) v = __dmul_rn(a, a);
      else if (q == 3) v = __dmul_rn(b, b);
      else if (q == 4) v = __dmul_rn(a, b);
      else {
        const int j = (q - 5 < Dg) ? q - 5 : q - 5 - Dg;
        const double ab = __ddiv_rn(__dsub_rn(row[1 + j], mean), sd);
        if (q - 5 < Dg) v = __dmul_rn(b, __dsub_rn(ab, a));          // B * (AB_j - A)
        else { const double t = __dsub_rn(a, ab); v = __dmul_rn(t, t); }  // (A - AB_j)^2
      }
    } else if (q < nfeat) {
      // pair p -> (j, k): p = j*(2Dg - j - 1)/2 + (k - j - 1)
      const int p = q - 5 - 2 * Dg;
      const double t = 2.0 * Dg - 1.0;
      int j = (int)((t - sqrt(t * t - 8.0 * p)) * 0.5);
      while (j > 0 && j * (2 * Dg - j - 1) / 2 > p) --j;
      while ((j + 1) * (2 * Dg - j - 2) / 2 <= p) ++j;
      const int k = j + 1 + (p - j * (2 * Dg - j - 1) / 2);
      const double ba = __ddiv_rn(__dsub_rn(row[1 + Dg + j], mean), sd);
      const double ab = __ddiv_rn(__dsub_rn(row[1 + k], mean), sd);
      v = __dmul_rn(ba, ab);
    }
    F[idx] = v;
  }
}

constexpr int kFoStages = 4;

// CH: features per chunk (= chunk width of the feature matrix: 8, 16, 24 or 32); RPL: resamples per lane;
// FPW: features per warp (CH / FPW warps share a group of 32 * RPL resamples, each taking FPW features).
// A warp-uniform LDS.128 costs 2 crossbar wavefronts (of the SM's 1 per clock) for 2 features; the 2 DFMAs it
// feeds with one resample per lane cost 4 pipe-clocks on one of 4 schedulers = 1 SM-clock: <CH, 1, CH> is
// crossbar-bound at half the FP64 rate (measured 89 % of that bound; used for launches of <= 256
// resamples).  <CH, 2, CH> reuses every broadcast for 4 DFMAs, which balances the two pipes (cfg2: 3.1 -> 2.65
// ms; FP64 pipe 55 %, crossbar 55 % -- the rest is the lane-strided weight fetch, see launch_feature_gemv).
template <int CH, int RPL, int FPW, bool WEIGHTED>
__global__ void __launch_bounds__(kFoWarps * 32)
fo_gemv_kernel(const double *__restrict__ Fall, uint64_t N, const uint8_t *__restrict__ w, uint64_t Npad,
               uint64_t count, uint64_t rows_per_split, double *__restrict__ psum, int nacc) {
  constexpr int NF = CH;                                           // row length of this chunk
  const double *__restrict__ F = Fall + (uint64_t)blockIdx.z * N * CH;  // chunk blockIdx.z: [N][CH]
  // Feature rows are streamed through a shared-memory ring exactly like the tile kernel's rows:
  // 32-row tiles (contiguous in HBM) by cp.async.bulk, one mbarrier per stage, last warp out refills.
  extern __shared__ __align__(128) unsigned char smraw[];
  const int lane = threadIdx.x & 31;
  constexpr int FG = CH / FPW;               // warps per resample group
  const uint32_t nwarps = blockDim.x >> 5;   // a multiple of FG
  const uint32_t warp = threadIdx.x >> 5;
  const int fq = (int)(warp % FG) * FPW;     // first feature (within the chunk) of this warp
  const uint64_t cbase = ((uint64_t)blockIdx.x * (nwarps / FG) + warp / FG) * (32 * RPL) + lane;
  const int sp = blockIdx.y;
  const int q0 = blockIdx.z * CH;
  const uint64_t row0 = (uint64_t)sp * rows_per_split;
  uint64_t row1 = row0 + rows_per_split;
  if (row1 > N) row1 = N;
  const int ntiles = (row1 > row0) ? (int)((row1 - row0 + kFoRows - 1) / kFoRows) : 0;
  const uint32_t row_bytes = (uint32_t)NF * 8u;
  const uint32_t stage_bytes = kFoRows * row_bytes;
  const uint32_t sm_base = smem_u32(smraw);
  const uint32_t bar_base = sm_base + kFoStages * stage_bytes;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(smraw + kFoStages * stage_bytes + kFoStages * 8);
  auto issue = [&](int t) {
    const int s = t % kFoStages;
    const uint64_t r = row0 + (uint64_t)t * kFoRows;
    const uint32_t rows = (uint32_t)((row1 - r < (uint64_t)kFoRows) ? (row1 - r) : kFoRows);
    mbar_arrive_expect_tx(bar_base + 8u * s, rows * row_bytes);
    bulk_g2s(sm_base + s * stage_bytes, F + r * (uint64_t)NF, rows * row_bytes, bar_base + 8u * s);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kFoStages; ++s) {
      mbar_init(bar_base + 8u * s, 1);
      cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 0; t < kFoStages && t < ntiles; ++t) issue(t);

  double acc[RPL][FPW];
  const uint8_t *wrow[RPL];
  double ones[RPL];
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
#pragma unroll
    for (int q = 0; q < FPW; ++q) acc[k][q] = 0.0;
    const uint64_t c = cbase + 32u * k;
    wrow[k] = (WEIGHTED && c < count) ? w + c * Npad : nullptr;
    ones[k] = (!WEIGHTED && c < count) ? 1.0 : 0.0;  // point estimate: weight 1 for every row
  }
  // this lane's weights, 16 rows (one uint4 per resample) at a time: half-tile u covers rows [16u, 16u+16) of
  // the split (rows_per_split is a multiple of 32, Npad of 128: in bounds); fetched one half-tile ahead
  auto load_w = [&](int u, uint4 (&wn)[RPL]) {
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
      wn[k] = make_uint4(0, 0, 0, 0);
      if (wrow[k] && u < 2 * ntiles) wn[k] = __ldg(reinterpret_cast<const uint4 *>(wrow[k] + row0 + 16u * (uint64_t)u));
    }
  };
  uint4 wn[RPL];
  load_w(0, wn);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t % kFoStages;
    const uint64_t i0 = row0 + (uint64_t)t * kFoRows;
    const int nrows = (int)((row1 - i0 < (uint64_t)kFoRows) ? (row1 - i0) : kFoRows);
    while (!mbar_try_wait(bar_base + 8u * s, (uint32_t)(t / kFoStages) & 1u)) {
    }
    const uint32_t tile = sm_base + s * stage_bytes;
    // 4 rows per iteration; the loop is kept rolled (the weight words rotate through wreg[k][0]) so that the
    // body stays in the instruction cache.  Full tiles run a branch-free body (WEIGHTED is a template
    // parameter, no per-row bounds test), so the 4 rows are one basic block and ptxas issues the next rows'
    // broadcasts ahead of the DFMAs that consume the current ones; only the last tile of a split is checked.
    uint32_t wreg[RPL][4];
    auto four_rows = [&](int wi, uint32_t (&cur)[RPL], auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int r = 4 * wi + bb;
        if (FULL || r < nrows) {
          double wv[RPL];
#pragma unroll
          for (int k = 0; k < RPL; ++k) {
            const uint32_t byte = (cur[k] >> (8 * bb)) & 0xffu;
            // small integer -> double without the conversion pipe: 2^52 + n has n in the low mantissa bits
            wv[k] = WEIGHTED ? (__hiloint2double(0x43300000, (int)byte) - 4503599627370496.0) : ones[k];
          }
          const uint32_t f = tile + (uint32_t)r * row_bytes + 8u * (uint32_t)fq;
#pragma unroll
          for (int q = 0; q < FPW; q += 2) {
            const double2 v = lds_f64x2(f + 8u * q);  // warp-uniform address: one broadcast
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
              acc[k][q] = fma(wv[k], v.x, acc[k][q]);
              acc[k][q + 1] = fma(wv[k], v.y, acc[k][q + 1]);
            }
          }
        }
      }
    };
#pragma unroll 1
    for (int wi = 0; wi < 8; ++wi) {
      if (WEIGHTED && (wi & 3) == 0) {  // uniform: start of a half-tile
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          wreg[k][0] = wn[k].x; wreg[k][1] = wn[k].y; wreg[k][2] = wn[k].z; wreg[k][3] = wn[k].w;
        }
        load_w(2 * t + (wi >> 2) + 1, wn);
      }
      uint32_t cur[RPL];
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        cur[k] = WEIGHTED ? wreg[k][0] : 0u;
        if (WEIGHTED) {
          wreg[k][0] = wreg[k][1]; wreg[k][1] = wreg[k][2]; wreg[k][2] = wreg[k][3];
        }
      }
      if (nrows == kFoRows) four_rows(wi, cur, std::true_type{});
      else four_rows(wi, cur, std::false_type{});
    }
    __syncwarp();
    if (lane == 0) {
      uint32_t old;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;"
                   : "=r"(old)
                   : "r"(smem_u32(cnt + s))
                   : "memory");
      if (old == nwarps - 1) {
        cnt[s] = 0;
        if (t + kFoStages < ntiles) issue(t + kFoStages);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const uint64_t c = cbase + 32u * k;
    if (c >= count) continue;
    double *out = psum + ((uint64_t)sp * count + c) * (uint64_t)nacc;
#pragma unroll
    for (int q = 0; q < FPW; ++q)
      if (q0 + fq + q < nacc) out[q0 + fq + q] = acc[k][q];
  }
}

// ------------------------------------------------------------------------------------------
// DGSM (analyze/dgsm.py): the bootstrap of mean((dY/dx_j)^2) is the same skinny GEMM with other features
// ------------------------------------------------------------------------------------------
// F[chunk][i][CW], features q < D: ((Y_j[i] - base[i]) / (X_j[i][j] - X_base[i][j]))^2 (dgsm.py:118-120),
// D <= q < 2D: its square (for vi_std).  X is the finite_diff matrix [N*(D+1), D], Y the model output.
__global__ void __launch_bounds__(256)
dgsm_features_kernel(const double *__restrict__ X, const double *__restrict__ Y, uint64_t N, int D,
                     double *__restrict__ F) {
  const int nfeat = 2 * D, CW = chunk_width(nfeat), NZ = chunk_count(nfeat);
  const uint64_t per_chunk = N * (uint64_t)CW, total = per_chunk * NZ;
  const int step = D + 1;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const int z = (int)(idx / per_chunk);
    const uint64_t rem = idx - (uint64_t)z * per_chunk;
    const uint64_t i = rem / (uint64_t)CW;
    const int q = z * CW + (int)(rem - i * (uint64_t)CW);
    double v = 0.0;
    if (q < nfeat) {
      const int j = q < D ? q : q - D;
      const uint64_t r0 = i * (uint64_t)step, rj = r0 + 1 + j;
      const double dy = __dsub_rn(Y[rj], Y[r0]);
      const double dx = __dsub_rn(X[rj * (uint64_t)D + j], X[r0 * (uint64_t)D + j]);
      const double g = __ddiv_rn(dy, dx);
      v = __dmul_rn(g, g);
      if (q >= D) v = __dmul_rn(v, v);
    }
    F[idx] = v;
  }
}

// out[c][q] = scale * sum_sp psum[sp][c][q]
__global__ void __launch_bounds__(256)
reduce_splits_kernel(const double *__restrict__ psum, uint64_t count, int nfeat, int nsplit, double scale,
                     double *__restrict__ out) {
  const uint64_t total = count * (uint64_t)nfeat;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) v += psum[(uint64_t)sp * total + idx];
    out[idx] = v * scale;
  }
}

// finite_diff rows (sample/finite_diff.py:76-91): per base point its D+1 rows -- the point, then the point
// with column j moved by base_j * delta and clipped to [lo_j, hi_j].
__global__ void __launch_bounds__(256)
finite_diff_expand_kernel(const double *__restrict__ base, uint64_t n, int D, double delta,
                          const double *__restrict__ lo, const double *__restrict__ hi, double *__restrict__ out) {
  const uint64_t rowlen = (uint64_t)(D + 1) * D, total = n * rowlen;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = idx / rowlen;
    const uint32_t q = (uint32_t)(idx - i * rowlen);
    const int s = q / D, j = q - s * D;
    double v = base[i * (uint64_t)D + j];
    if (s == j + 1) {
      v = __dadd_rn(v, __dmul_rn(v, delta));
      v = fmax(fmin(v, hi[j]), lo[j]);
    }
    out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------
// generic accumulate: block <-> (resample, split), thread <-> accumulator
// ------------------------------------------------------------------------------------------
constexpr int kGenThreads = 256;
constexpr int kGenRows = 32;

__global__ void __launch_bounds__(kGenThreads)
accum_generic_kernel(const double *__restrict__ Yn, int stride, int Dp, uint64_t N, int Dg,
                     int second_order, const uint8_t *__restrict__ w, uint64_t Npad, uint64_t count,
                     uint64_t rows_per_split, double *__restrict__ psum, int nacc) {
  const uint64_t c = blockIdx.x;
  const int sp = blockIdx.y;
  const int PL = part_len(Dg);
  const uint64_t row0 = (uint64_t)sp * rows_per_split;
  uint64_t row1 = row0 + rows_per_split;
  if (row1 > N) row1 = N;
  const uint8_t *wrow = w ? w + c * Npad : nullptr;
  const int offA = (second_order ? 2 : 1) * PL, offB = offA + 1;
  __shared__ double sw[kGenRows];
  double *out = psum + ((uint64_t)sp * count + c) * (uint64_t)nacc;

  for (int a0 = 0; a0 < nacc; a0 += kGenThreads) {
    const int a = a0 + threadIdx.x;
    // decode accumulator a -> kind, operand columns
    int kind = -1, cx = 0, cy = 0;
    if (a < nacc) {
      if (a < 5) kind = a;
      else if (a < 5 + Dg) { kind = 5; cx = col_pos(a - 5); }
      else if (a < 5 + 2 * Dg) { kind = 6; cx = col_pos(a - 5 - Dg); }
      else {
        kind = 7;
        int p = a - 5 - 2 * Dg, j = 0;
        while (p >= Dg - 1 - j) { p -= Dg - 1 - j; ++j; }
        cx = PL + col_pos(j);       // BA_j
        cy = col_pos(j + 1 + p);    // AB_k
      }
    }
    double accv = 0.0;
    for (uint64_t i0 = row0; i0 < row1; i0 += kGenRows) {
      __syncthreads();
      if (threadIdx.x < kGenRows) {
        const uint64_t i = i0 + threadIdx.x;
        sw[threadIdx.x] = (i < row1) ? (wrow ? (double)wrow[i] : 1.0) : 0.0;
      }
      __syncthreads();
      if (kind < 0) continue;
      for (int r = 0; r < kGenRows; ++r) {
        const double wv = sw[r];
        if (wv == 0.0) continue;
        const double *row = Yn + (i0 + r) * (uint64_t)stride;
        const double A = row[offA], B = row[offB];
        double term;
        switch (kind) {
          case 0: term = A; break;
          case 1: term = B; break;
          case 2: term = A * A; break;
          case 3: term = B * B; break;
          case 4: term = A * B; break;
          case 5: term = B * (row[cx] - A); break;
          case 6: { const double t = A - row[cx]; term = t * t; break; }
          default: term = row[cx] * row[cy]; break;
        }
        accv = fma(wv, term, accv);
      }
    }
    if (a < nacc) out[a] = accv;
  }
}

// ------------------------------------------------------------------------------------------
// finalize: sums -> S1, ST, S2 per resample
// ------------------------------------------------------------------------------------------
template <bool STAGE>  // STAGE: sums over the row splits staged in shared memory (Dg + nacc doubles fit)
__global__ void __launch_bounds__(128)
finalize_kernel(const double *__restrict__ psum, uint64_t N, int Dg, int second_order, uint64_t count,
                int nsplit, const double *__restrict__ stats, double *__restrict__ est, int nacc, int nest) {
  const uint64_t c = blockIdx.x;
  extern __shared__ double fsh[];
  double *s1sh = fsh;                          // [Dg]
  double *tot_sh = STAGE ? fsh + Dg : nullptr; // [nacc] sums over the row splits (fixed order: deterministic)
  __shared__ double head[5];
  auto direct = [&](int a) {
    double v = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) v += psum[((uint64_t)sp * count + c) * (uint64_t)nacc + a];
    return v;
  };
  if (STAGE) {
    if (nsplit == 1) {
      for (int a = threadIdx.x; a < nacc; a += blockDim.x) tot_sh[a] = psum[c * (uint64_t)nacc + a];
    } else {
      // one warp per accumulator, lanes stride over the splits: the loads of a sum are in flight together
      const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int a = threadIdx.x >> 5; a < nacc; a += nw) {
        double v = 0.0;
        for (int sp = lane; sp < nsplit; sp += 32) v += psum[((uint64_t)sp * count + c) * (uint64_t)nacc + a];
        v = warp_sum(v);
        if (lane == 0) tot_sh[a] = v;
      }
    }
    __syncthreads();
  }
  auto tot = [&](int a) { return STAGE ? tot_sh[a] : direct(a); };
  if (threadIdx.x < 5) head[threadIdx.x] = tot(threadIdx.x);
  __syncthreads();
  const bool constant = !(stats[4] < stats[5]);  // ptp([A;B]) == 0 (or NaN)
  const double n = (double)N;
  const double m = (head[0] + head[1]) / (2.0 * n);
  const double V = (head[2] + head[3]) / (2.0 * n) - m * m;
  double *o = est + c * (uint64_t)nest;
  for (int j = threadIdx.x; j < Dg; j += blockDim.x) {
    const double s1 = constant ? 0.0 : (tot(5 + j) / n) / V;
    const double st = constant ? 0.0 : 0.5 * (tot(5 + Dg + j) / n) / V;
    s1sh[j] = s1;
    o[j] = s1;
    o[Dg + j] = st;
  }
  __syncthreads();
  if (second_order) {
    const int npairs = Dg * (Dg - 1) / 2;
    const double mab = head[4] / n;
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
      int q = p, j = 0;
      while (q >= Dg - 1 - j) { q -= Dg - 1 - j; ++j; }
      const int k = j + 1 + q;
      const double vjk = (tot(5 + 2 * Dg + p) / n - mab) / V;
      o[2 * Dg + p] = constant ? 0.0 : vjk - s1sh[j] - s1sh[k];
    }
  }
}

// conf[col] = z * std(est[:, col], ddof=1): two passes (mean, then squared deviations)
__global__ void __launch_bounds__(512)
conf_kernel(const double *__restrict__ est, uint64_t R, int ncols, double z, double *__restrict__ conf) {
  __shared__ double sh[16][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  double s = 0.0;
  if (col < ncols)
    for (uint64_t c = ty; c < R; c += 16) s += est[c * (uint64_t)ncols + col];
  sh[ty][threadIdx.x] = s;
  __syncthreads();
  double mean = 0.0;
  for (int q = 0; q < 16; ++q) mean += sh[q][threadIdx.x];
  mean /= (double)R;
  __syncthreads();
  double ss = 0.0;
  if (col < ncols)
    for (uint64_t c = ty; c < R; c += 16) {
      const double d = est[c * (uint64_t)ncols + col] - mean;
      ss = fma(d, d, ss);
    }
  sh[ty][threadIdx.x] = ss;
  __syncthreads();
  if (ty == 0 && col < ncols) {
    double t = 0.0;
    for (int q = 0; q < 16; ++q) t += sh[q][threadIdx.x];
    conf[col] = z * sqrt(t / (double)(R - 1));  // R == 1 -> NaN like numpy's ddof=1
  }
}

// out[q] = sum_i (F[i][q] - center[q])^2 over the rows of the chunked feature matrix (q < ncols): the second pass of
// a two-pass standard deviation (np.std of analyze/dgsm.py:109).  One block per column, fixed-order reduction.
__global__ void __launch_bounds__(256)
feature_centered_ss_kernel(const double *__restrict__ F, uint64_t N, int nfeat, const double *__restrict__ center,
                           double *__restrict__ out) {
  __shared__ double sh[8];
  const int q = blockIdx.x;
  const int CW = chunk_width(nfeat);
  const double *col = F + (uint64_t)(q / CW) * N * CW + (q % CW);
  const double c = center[q];
  double s = 0.0;
  for (uint64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double d = col[i * (uint64_t)CW] - c;
    s = fma(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[q] = t;
  }
}

// The two passes of conf_kernel as separate column reductions, for estimates that are spread over ranks:
//   power 1: out[col] = sum_r est[r][col]                     (all-reduced, / R -> mean)
//   power 2: out[col] = sum_r (est[r][col] - mean[col])^2     (all-reduced -> conf_finish_kernel)
__global__ void __launch_bounds__(512)
col_moment_kernel(const double *__restrict__ est, uint64_t R, int ncols, const double *__restrict__ mean,
                  double *__restrict__ out) {
  __shared__ double sh[16][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  double s = 0.0;
  if (col < ncols) {
    if (mean == nullptr) {
      for (uint64_t c = ty; c < R; c += 16) s += est[c * (uint64_t)ncols + col];
    } else {
      const double m = mean[col];
      for (uint64_t c = ty; c < R; c += 16) {
        const double d = est[c * (uint64_t)ncols + col] - m;
        s = fma(d, d, s);
      }
    }
  }
  sh[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty == 0 && col < ncols) {
    double t = 0.0;
    for (int q = 0; q < 16; ++q) t += sh[q][threadIdx.x];
    out[col] = t;
  }
}

// x[col] = x[col] * scale   or   x[col] = z * sqrt(x[col] / (R - 1))  (the last step of the CI)
__global__ void __launch_bounds__(256)
conf_finish_kernel(double *__restrict__ x, int ncols, double scale, double z, uint64_t R, int finish) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  x[col] = finish ? z * sqrt(x[col] / (double)(R - 1)) : x[col] * scale;
}

// The CI over estimates that are spread over the ranks of a row-sharded analyze with ONE collective: every rank
// reduces its own estimates to (n_g, mean_g[col], M2_g[col] = sum (x - mean_g)^2) -- two passes inside one kernel --,
// the records are all-gathered, and every rank combines them (Chan, Golub & LeVeque):
//   M2 = sum_g M2_g + sum_g n_g (mean_g - mean)^2,   conf = z * sqrt(M2 / (n - 1))      (analyze/sobol.py:159,173)
// rec: [0] = n_g, [1 .. ncols] = mean_g, [1 + ncols .. 2 ncols] = M2_g.
__global__ void __launch_bounds__(512)
conf_local_kernel(const double *__restrict__ est, uint64_t R_all, int ncols, double *__restrict__ rec_all) {
  __shared__ double sh[16][33];
  __shared__ double mean_sh[32];
  // blockIdx.y = one of gridDim.y contiguous row slices, each with its own record (they are combined like ranks)
  const uint64_t per = (R_all + gridDim.y - 1) / gridDim.y;
  const uint64_t r_lo = min(R_all, per * blockIdx.y), r_hi = min(R_all, r_lo + per);
  const uint64_t R = r_hi - r_lo;
  est += r_lo * (uint64_t)ncols;
  double *rec = rec_all + (size_t)blockIdx.y * (1 + 2 * (size_t)ncols);
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  if (blockIdx.x == 0 && threadIdx.x == 0 && ty == 0) rec[0] = (double)R;
  double s = 0.0;
  if (col < ncols)
    for (uint64_t c = ty; c < R; c += 16) s += est[c * (uint64_t)ncols + col];
  sh[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty == 0) {
    double t = 0.0;
    for (int q = 0; q < 16; ++q) t += sh[q][threadIdx.x];
    mean_sh[threadIdx.x] = R ? t / (double)R : 0.0;
  }
  __syncthreads();
  const double m = mean_sh[threadIdx.x];
  s = 0.0;
  if (col < ncols)
    for (uint64_t c = ty; c < R; c += 16) {
      const double d = est[c * (uint64_t)ncols + col] - m;
      s = fma(d, d, s);
    }
  __syncthreads();
  sh[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty == 0 && col < ncols) {
    double t = 0.0;
    for (int q = 0; q < 16; ++q) t += sh[q][threadIdx.x];
    rec[1 + col] = m;
    rec[1 + ncols + col] = t;
  }
}

// recs: k records `pitch` doubles apart, each [n_g, mean_g[ncols], M2_g[ncols], extra[nextra]];
// conf[col] as above, extra_sum[e] = sum_g extra_g[e] (the point estimates -- zero on every rank but one -- and
// the guard counts ride in the same all-gather).
__global__ void __launch_bounds__(256)
conf_combine_kernel(const double *__restrict__ recs, int k, size_t pitch, int ncols, int nextra, double z,
                    double *__restrict__ conf, double *__restrict__ extra_sum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncols) {
    double n = 0.0, sum = 0.0;
    for (int g = 0; g < k; ++g) {
      const double c = recs[g * pitch];
      n += c;
      sum += c * recs[g * pitch + 1 + i];
    }
    const double mean = sum / n;
    double m2 = 0.0;
    for (int g = 0; g < k; ++g) {
      const double c = recs[g * pitch];
      if (!(c > 0.0)) continue;
      const double d = recs[g * pitch + 1 + i] - mean;
      m2 += recs[g * pitch + 1 + ncols + i] + c * d * d;
    }
    conf[i] = z * sqrt(m2 / (n - 1.0));
  } else if (i < ncols + nextra) {
    const int e = i - ncols;
    double t = 0.0;
    for (int g = 0; g < k; ++g) t += recs[g * pitch + 1 + 2 * (size_t)ncols + e];
    extra_sum[e] = t;
  }
}

// Statistics of the union of k row shards from per-shard (n_values, stats8) records (Chan, Golub & LeVeque), on
// the device so that a row-sharded analyze needs no host round trip between the moments and the features:
// out[0..7] = mean, population std, min, max, min/max over the A and B columns, 0, 0; out[8] = total n_values.
__device__ void combine_moment_records(const double *__restrict__ parts, int k, size_t pitch, double *out) {
  double n = 0.0, sum = 0.0;
  for (int i = 0; i < k; ++i) {
    const double c = parts[pitch * i];
    if (c > 0.0) {
      n += c;
      sum += c * parts[pitch * i + 1];
    }
  }
  const double mean = sum / n;
  double m2 = 0.0, mn = INFINITY, mx = -INFINITY, amn = INFINITY, amx = -INFINITY;
  for (int i = 0; i < k; ++i) {
    const double *p = parts + pitch * i;
    if (!(p[0] > 0.0)) continue;
    const double d = p[1] - mean;
    m2 += p[0] * (p[2] * p[2] + d * d);
    mn = fmin(mn, p[3]);
    mx = fmax(mx, p[4]);
    amn = fmin(amn, p[5]);
    amx = fmax(amx, p[6]);
    // a NaN anywhere must survive (fmin/fmax drop it): the caller's degenerate-output test looks at these
    if (p[3] != p[3] || p[4] != p[4]) mn = mx = p[3] + p[4];
  }
  out[0] = mean;
  out[1] = sqrt(m2 / n);
  out[2] = mn;
  out[3] = mx;
  out[4] = amn;
  out[5] = amx;
  out[6] = 0.0;
  out[7] = 0.0;
  out[8] = n;
}

__global__ void combine_moments_kernel(const double *__restrict__ parts, int k, size_t pitch,
                                       double *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  combine_moment_records(parts, k, pitch, out);
}

// One record per rank for the single collective in front of a row-sharded analyze: the rank's k shard records
// combined into (n_values, stats8) -- a copy when k == 1 --, followed by three numbers every rank must agree on
// (all-gathered with the moments, reduced with max on the host).
__global__ void rank_record_kernel(const double *__restrict__ parts, int k, double f0, double f1, double f2,
                                   double *__restrict__ rec) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (k == 1) {
    for (int i = 0; i < 9; ++i) rec[i] = parts[i];
  } else {
    double st[9];
    combine_moment_records(parts, k, 9, st);
    rec[0] = st[8];
    for (int i = 0; i < 8; ++i) rec[1 + i] = st[i];
  }
  rec[9] = f0;
  rec[10] = f1;
  rec[11] = f2;
}

}  // namespace sa

using namespace sa;

// psum[nsplit][count][nfeat] = sum over the rows of each split of w[c][i] * F[i][q]
static int launch_feature_gemv(cudaStream_t st, const double *d_F, uint64_t N, int nfeat, const uint8_t *d_w,
                               uint64_t Npad, uint64_t count, int nsplit, uint64_t rps, double *d_psum) {
  const int ch = chunk_width(nfeat), nz = chunk_count(nfeat);
  SA_REQUIRE(nz <= 65535, SA_ERR_UNSUPPORTED, "shape too large for the feature GEMV kernel");
  const size_t smem = (size_t)kFoStages * kFoRows * ch * 8 + kFoStages * 12 + 16;
  if (count > (uint64_t)kFoWarps * 32 && ch == 16) {
    // <16, 4, 16>: 4 resamples per lane on a 16-wide chunk -- 8 warp-uniform LDS.128 (16 wavefronts) feed 64
    // DFMAs (32 SM-clocks): FP64-bound.  (The same tile with the two halves of a 32-wide chunk on two warps
    // was no faster than <32, 2, 32>: a warp-dependent offset turns the broadcast into a per-lane address,
    // 4 wavefronts instead of 2.)
    const uint64_t bx = ceil_div_u64(count, 512);
    SA_REQUIRE(bx < (1ull << 31), SA_ERR_UNSUPPORTED, "shape too large for the feature GEMV kernel");
    dim3 grid((unsigned)bx, (unsigned)nsplit, (unsigned)nz);
    fo_gemv_kernel<16, 4, 16, true><<<grid, 128, smem, st>>>(d_F, N, d_w, Npad, count, rps, d_psum, nfeat);
  } else if (count > (uint64_t)kFoWarps * 32) {
    // <CH, 2, CH>: two resamples per lane; ~165 registers, so blocks of 4 warps (256 resamples), 3 per SM.
    // (<CH, 4, 8> -- 4 per lane, CH/8 warps per resample group -- was measured slower, 3.55 vs 2.65 ms on cfg2:
    // every warp of a group re-reads the group's weight bytes with lane-strided 16-byte loads.)
    const uint64_t bx = ceil_div_u64(count, 256);
    SA_REQUIRE(bx < (1ull << 31), SA_ERR_UNSUPPORTED, "shape too large for the feature GEMV kernel");
    dim3 grid((unsigned)bx, (unsigned)nsplit, (unsigned)nz);
    auto kern = ch == 8 ? fo_gemv_kernel<8, 2, 8, true> : (ch == 24 ? fo_gemv_kernel<24, 2, 24, true> : fo_gemv_kernel<32, 2, 32, true>);
    kern<<<grid, 128, smem, st>>>(d_F, N, d_w, Npad, count, rps, d_psum, nfeat);
  } else {
    dim3 grid(1, (unsigned)nsplit, (unsigned)nz);
    const unsigned threads = (unsigned)(32 * ceil_div_u64(count, 32));  // only warps that own a resample
    if (d_w) {
      auto kern = ch == 8 ? fo_gemv_kernel<8, 1, 8, true> : (ch == 16 ? fo_gemv_kernel<16, 1, 16, true> : (ch == 24 ? fo_gemv_kernel<24, 1, 24, true> : fo_gemv_kernel<32, 1, 32, true>));
      kern<<<grid, threads, smem, st>>>(d_F, N, d_w, Npad, count, rps, d_psum, nfeat);
    } else {
      auto kern = ch == 8 ? fo_gemv_kernel<8, 1, 8, false> : (ch == 16 ? fo_gemv_kernel<16, 1, 16, false> : (ch == 24 ? fo_gemv_kernel<24, 1, 24, false> : fo_gemv_kernel<32, 1, 32, false>));
      kern<<<grid, threads, smem, st>>>(d_F, N, d_w, Npad, count, rps, d_psum, nfeat);
    }
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_row_stride(int Dg, int second_order) { return row_stride(Dg, second_order); }
extern "C" uint64_t sa_pad_rows(uint64_t N) { return pad_rows(N); }
extern "C" int sa_feature_count(int Dg) { return feature_width(Dg) * feature_chunks(Dg); }
extern "C" size_t sa_moments_workspace_bytes(void) { return (size_t)kMomBlocks * 6 * sizeof(double); }

extern "C" int sa_moments_f64(void *stream, const double *d_Y, uint64_t N, int step, double *d_stats,
                              void *d_work, size_t work_bytes) {
  SA_REQUIRE(d_Y && d_stats && d_work, SA_ERR_ARG, "sa_moments_f64: null pointer");
  SA_REQUIRE(N >= 1 && step >= 1, SA_ERR_ARG, "sa_moments_f64: bad N/step");
  SA_REQUIRE(work_bytes >= sa_moments_workspace_bytes(), SA_ERR_WORKSPACE, "sa_moments_f64: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t total = N * (uint64_t)step;
  int blocks = (int)ceil_div_u64(total, kMomThreads * 4);
  if (blocks > kMomBlocks) blocks = kMomBlocks;
  if (blocks < 1) blocks = 1;
  double *part = (double *)d_work;
  moments_pass1_kernel<<<blocks, kMomThreads, 0, st>>>(d_Y, total, step, part);
  moments_final1_kernel<<<1, kMomThreads, 0, st>>>(part, blocks, total, d_stats);
  moments_pass2_kernel<<<blocks, kMomThreads, 0, st>>>(d_Y, total, d_stats, part + (size_t)kMomBlocks * 5);
  moments_final2_kernel<<<1, kMomThreads, 0, st>>>(part + (size_t)kMomBlocks * 5, blocks, total, d_stats);
  note_launches(4);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_normalize_f64(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                                const double *d_stats, double *d_Yn, double *d_Yt) {
  SA_REQUIRE(d_Y && d_stats && (d_Yn || d_Yt), SA_ERR_ARG, "sa_normalize_f64: null pointer");
  SA_REQUIRE(N >= 1 && Dg >= 1, SA_ERR_ARG, "sa_normalize_f64: bad N/Dg");
  cudaStream_t st = (cudaStream_t)stream;
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  const int sms = sm_count();
  if (d_Yn) {
    const uint64_t total = N * (uint64_t)row_stride(Dg, second_order);
    uint64_t blocks = ceil_div_u64(total, 256);
    if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
    normalize_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_Y, N, Dg, step, second_order, d_stats, d_Yn);
    note_launches(1);
  }
  if (d_Yt) {
    if (second_order) {  // A and B columns for scalar_sums_kernel
      const uint64_t Npad = pad_rows(N);
      uint64_t bx = ceil_div_u64(Npad, 256);
      if (bx > 4096) bx = 4096;
      dim3 grid((unsigned)bx, 2u);
      normalize_cols_kernel<<<grid, 256, 0, st>>>(d_Y, N, Npad, Dg, step, d_stats, d_Yt);
    } else {             // feature matrix for fo_gemv_kernel
      const uint64_t total = N * (uint64_t)feature_width(Dg) * feature_chunks(Dg);
      uint64_t blocks = ceil_div_u64(total, 256);
      if (blocks > (uint64_t)sms * 32) blocks = (uint64_t)sms * 32;
      normalize_features_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_Y, N, Dg, step, 0, d_stats, d_Yt);
    }
    note_launches(1);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// Feature matrix of the GEMV kernel for either order: SA_NACC(Dg, second_order) features per row, chunked.
extern "C" int sa_normalize_features_f64(void *stream, const double *d_Y, uint64_t N, int Dg, int second_order,
                                         const double *d_stats, double *d_F) {
  SA_REQUIRE(d_Y && d_stats && d_F, SA_ERR_ARG, "sa_normalize_features_f64: null pointer");
  SA_REQUIRE(N >= 1 && Dg >= 1 && Dg <= (second_order ? 2048 : (1 << 24)), SA_ERR_ARG,
             "sa_normalize_features_f64: bad N/Dg (second order: Dg <= 2048)");
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  const int nfeat = SA_NACC(Dg, second_order);
  const uint64_t total = N * (uint64_t)chunk_width(nfeat) * chunk_count(nfeat);
  uint64_t blocks = ceil_div_u64(total, 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  normalize_features_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_Y, N, Dg, step, second_order,
                                                                               d_stats, d_F);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_counts_from_indices_rows(void *stream, const int64_t *d_r, uint64_t N, uint64_t R, uint64_t c0,
                                           uint64_t count, uint64_t row0, uint64_t nrows, uint8_t *d_w,
                                           int32_t *d_overflow) {
  SA_REQUIRE(d_r && d_w, SA_ERR_ARG, "sa_counts_from_indices: null pointer");
  SA_REQUIRE(N >= 1 && c0 + count <= R, SA_ERR_ARG, "sa_counts_from_indices: resample range out of bounds");
  SA_REQUIRE(nrows >= 1 && row0 + nrows <= N, SA_ERR_ARG, "sa_counts_from_indices: row window out of bounds");
  SA_REQUIRE(((uintptr_t)d_w & 3) == 0, SA_ERR_ARG, "sa_counts_from_indices: d_w not 4-byte aligned");
  if (count == 0) return SA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t Npad = pad_rows(nrows);
  SA_CUDA(cudaMemsetAsync(d_w, 0, count * Npad, st));
  uint64_t blocks = ceil_div_u64(N * count, 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  counts_from_indices_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_r, N, R, c0, count, row0, nrows, Npad,
                                                              (uint32_t *)d_w, d_overflow);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_counts_from_indices(void *stream, const int64_t *d_r, uint64_t N, uint64_t R, uint64_t c0,
                                      uint64_t count, uint8_t *d_w, int32_t *d_overflow) {
  return sa_counts_from_indices_rows(stream, d_r, N, R, c0, count, 0, N, d_w, d_overflow);
}

// slim != 0: the low-footprint launch (one 1024-thread block per SM) that co-resides with the persistent GEMM.
extern "C" int sa_counts_philox_rows_ex(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                                        uint64_t row0, uint64_t nrows, uint8_t *d_w, int32_t *d_overflow, int slim) {
  (void)d_overflow;  // P(multiplicity > 255) < 1e-500 for N iid uniform draws over N rows, N >= 256
  SA_REQUIRE(d_w, SA_ERR_ARG, "sa_counts_philox: null pointer");
  SA_REQUIRE(N >= 1, SA_ERR_ARG, "sa_counts_philox: bad N");
  SA_REQUIRE(nrows >= 1 && row0 + nrows <= N, SA_ERR_ARG, "sa_counts_philox: row window out of bounds");
  SA_REQUIRE(((uintptr_t)d_w & 3) == 0, SA_ERR_ARG, "sa_counts_philox: d_w not 4-byte aligned");
  if (count == 0) return SA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t Npad = pad_rows(nrows);
  const int L = window_levels(N);
  if (L >= 0) {
    // every byte of the row window is written by the kernel: only the pad bytes of each row are cleared
    // (a full memset of the 10.5 GB headline buffer costs 3 ms)
    if (Npad > nrows) SA_CUDA(cudaMemset2DAsync(d_w + nrows, Npad, 0, Npad - nrows, count, st));
    const size_t smem = window_smem_bytes(N, L);
    SA_REQUIRE(smem <= 200 * 1024, SA_ERR_ARG, "sa_counts_philox: N too large");
    // per device and cheap: no process-wide cache (a second GPU of the process would miss the opt-in)
    const int opt_in = (int)(smem > 48 * 1024 ? smem : 48 * 1024);
    if (slim) {
      // threads of the one block per SM: more threads draw faster but take more of the shared-memory bandwidth the
      // tensor cores read their operands with (tools/exp_overlap.py); SALIB_B200_SLIM_THREADS overrides
      int nt = kWinThreadsSlim;
      if (const char *e = getenv("SALIB_B200_SLIM_THREADS")) nt = atoi(e);
      const uint64_t nb = count < (uint64_t)sm_count() ? count : (uint64_t)sm_count();
#define SA_SLIM(NT)                                                                                              \
  {                                                                                                              \
    SA_CUDA(cudaFuncSetAttribute(counts_window_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, opt_in)); \
    /* keep the SM's carveout at "all shared memory" so that a GEMM CTA (198 KB) can join while this runs */    \
    SA_CUDA(cudaFuncSetAttribute(counts_window_kernel<NT>, cudaFuncAttributePreferredSharedMemoryCarveout,       \
                                 (int)cudaSharedmemCarveoutMaxShared));                                          \
    counts_window_kernel<NT><<<(unsigned)nb, NT, smem, st>>>(seed, N, c0, count, L, row0, nrows, Npad, d_w);     \
  }
      // 512 threads: 5.6 ms per 2560 resamples at N = 2^20, hidden behind a 22 ms GEMM batch at no measurable
      // cost to it; 1024 threads draw no faster and slow the GEMM by ~15 % while they run
      if (nt <= 128) SA_SLIM(128)
      else if (nt <= 256) SA_SLIM(256)
      else if (nt <= 512) SA_SLIM(512)
      else SA_SLIM(1024)
#undef SA_SLIM
    } else {
      SA_CUDA(cudaFuncSetAttribute(counts_window_kernel<kWinThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   opt_in));
      const uint64_t nb = count < (1u << 30) ? count : (1u << 30);
      counts_window_kernel<kWinThreads><<<(unsigned)nb, kWinThreads, smem, st>>>(seed, N, c0, count, L, row0, nrows,
                                                                                Npad, d_w);
    }
    note_launches(1);
    SA_LAUNCH_CHECK();
    return SA_OK;
  }
  SA_CUDA(cudaMemsetAsync(d_w, 0, count * Npad, st));
  const uint64_t npairs = (N + 1) / 2;
  uint64_t bx = ceil_div_u64(npairs, 256 * 4);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  for (uint64_t done = 0; done < count; done += 65535) {
    const uint64_t nb = (count - done < 65535) ? count - done : 65535;
    dim3 grid((unsigned)bx, (unsigned)nb);
    counts_philox_kernel<<<grid, 256, 0, st>>>(seed, N, c0 + done, row0, nrows, Npad,
                                               (uint32_t *)(d_w + done * Npad));
    note_launches(1);
  }
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_counts_philox_rows(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                                     uint64_t row0, uint64_t nrows, uint8_t *d_w, int32_t *d_overflow) {
  return sa_counts_philox_rows_ex(stream, seed, N, c0, count, row0, nrows, d_w, d_overflow, 0);
}

extern "C" int sa_counts_philox(void *stream, uint64_t seed, uint64_t N, uint64_t c0, uint64_t count,
                                uint8_t *d_w, int32_t *d_overflow) {
  return sa_counts_philox_rows(stream, seed, N, c0, count, 0, N, d_w, d_overflow);
}

// algo: 0 auto, 1 generic, 3 first/total order, 4 second-order 8x8 register tiles, 5 experimental DMMA
extern "C" int sa_sobol_accumulate_ex_f64(void *stream, const double *d_Yn, const double *d_Yt, uint64_t N,
                                          int Dg, int second_order, const uint8_t *d_w, uint64_t count,
                                          int nsplit, double *d_psum, int algo) {
  SA_REQUIRE(d_psum, SA_ERR_ARG, "sa_sobol_accumulate_f64: null pointer");
  SA_REQUIRE(N >= 1 && Dg >= 1 && count >= 1 && nsplit >= 1, SA_ERR_ARG, "sa_sobol_accumulate_f64: bad shape");
  SA_REQUIRE(d_w != nullptr || count == 1, SA_ERR_ARG, "sa_sobol_accumulate_f64: point estimate needs count == 1");
  SA_REQUIRE(nsplit <= 65535, SA_ERR_ARG, "sa_sobol_accumulate_f64: nsplit too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int Dp = round_up8(Dg);
  const int stride = row_stride(Dg, second_order);
  const uint64_t Npad = pad_rows(N);
  const int nacc = SA_NACC(Dg, second_order);
  uint64_t rps = ceil_div_u64(N, nsplit);
  rps = (rps + 31) & ~(uint64_t)31;
  if (algo == 0) algo = second_order ? (Dg <= 512 ? 4 : 1) : 3;
  if (algo == 4) {
    SA_REQUIRE(second_order && Dg <= 512, SA_ERR_UNSUPPORTED, "tile kernel needs second order and Dg <= 512");
    SA_REQUIRE(d_Yn && d_Yt, SA_ERR_ARG, "sa_sobol_accumulate_f64: d_Yn / d_Yt is null");
    SA_REQUIRE(((uintptr_t)d_Yn & 15) == 0, SA_ERR_ARG, "sa_sobol_accumulate_f64: d_Yn not 16-byte aligned");
    const uint64_t bx = ceil_div_u64(count, kS2Warps);
    const uint64_t bsc = ceil_div_u64(count, kScC);
    SA_REQUIRE(bx < (1ull << 31), SA_ERR_UNSUPPORTED, "too many resamples in one call");
    // Ring geometry: 32-row tiles and up to 5 stages while they fit 224 KB (Dg <= 64: 5 x 41 KB);
    // longer rows get fewer stages, then 16- or 8-row tiles.  The weight-1 fast path (no multiplications
    // for 58 % of the rows) is compiled out: measured no gain, the kernel is bound by shared-memory wavefronts.
    const size_t row_bytes = (size_t)stride * 8, budget = 224 * 1024;
    int tile_rows = 32;
    while (tile_rows > 8 && 2 * tile_rows * row_bytes > budget) tile_rows >>= 1;
    int stages = (int)(budget / (tile_rows * row_bytes));
    if (stages > 5) stages = 5;
    SA_REQUIRE(stages >= 2, SA_ERR_UNSUPPORTED, "tile kernel: rows of %zu bytes do not fit shared memory", row_bytes);
    const int TBh = Dp >> 3;
    const int ngroups = TBh > 8 ? (TBh * (TBh + 1) / 2 + 31) / 32 : 1;
    auto kern = TBh <= 8 ? (stages == 5 ? s2_accum_smem_kernel<32, 5, false> : s2_accum_smem_kernel<32, 0, false>)
                         : (tile_rows == 32 ? s2_accum_smem_kernel<32, 0, true>
                                            : (tile_rows == 16 ? s2_accum_smem_kernel<16, 0, true>
                                                               : s2_accum_smem_kernel<8, 0, true>));
    const size_t smem = (size_t)stages * tile_rows * row_bytes + stages * 12 + 16;
    SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scalar_sums_kernel<<<dim3((unsigned)bsc, (unsigned)nsplit), 256, 0, st>>>(d_Yt, Npad, N, d_w, count, rps, d_psum,
                                                                               nacc);
    kern<<<dim3((unsigned)bx, (unsigned)nsplit, (unsigned)ngroups), kS2Warps * 32, smem, st>>>(
        d_Yn, stride, Dp, N, Dg, d_w, Npad, count, rps, d_psum, nacc, stages);
    note_launches(1);
  } else if (algo == 5) {  // experimental FP64 tensor-core variant (not the default, see the kernel header)
    SA_REQUIRE(second_order && Dg <= 64, SA_ERR_UNSUPPORTED, "DMMA kernel needs second order and Dg <= 64");
    SA_REQUIRE(d_Yn && d_Yt, SA_ERR_ARG, "sa_sobol_accumulate_f64: d_Yn / d_Yt is null");
    const uint64_t bx = ceil_div_u64(count, kS2Warps);
    SA_REQUIRE(bx < (1ull << 31), SA_ERR_UNSUPPORTED, "too many resamples in one call");
    constexpr int stages = 5;
    const size_t smem = (size_t)stages * 32 * stride * 8 + stages * 12 + 16;
    SA_CUDA(cudaFuncSetAttribute(s2_accum_dmma_kernel<stages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    scalar_sums_kernel<<<dim3((unsigned)ceil_div_u64(count, kScC), (unsigned)nsplit), 256, 0, st>>>(
        d_Yt, Npad, N, d_w, count, rps, d_psum, nacc);
    s2_accum_dmma_kernel<stages><<<dim3((unsigned)bx, (unsigned)nsplit), kS2Warps * 32, smem, st>>>(
        d_Yn, stride, Dp, N, Dg, d_w, Npad, count, rps, d_psum, nacc);
    note_launches(1);
  } else if (algo == 3) {
    // second order: d_Yt must be the pair-feature matrix of sa_normalize_features_f64(second_order = 1)
    SA_REQUIRE(d_Yt, SA_ERR_ARG, "sa_sobol_accumulate_f64: d_Yt (feature matrix) is null");
    return launch_feature_gemv(st, d_Yt, N, nacc, d_w, Npad, count, nsplit, rps, d_psum);
  } else if (algo == 1) {
    SA_REQUIRE(d_Yn, SA_ERR_ARG, "sa_sobol_accumulate_f64: d_Yn is null");
    SA_REQUIRE(count < (1ull << 31), SA_ERR_UNSUPPORTED, "too many resamples in one call");
    dim3 grid((unsigned)count, (unsigned)nsplit);
    accum_generic_kernel<<<grid, kGenThreads, 0, st>>>(d_Yn, stride, Dp, N, Dg, second_order, d_w, Npad, count,
                                                      rps, d_psum, nacc);
  } else {
    SA_REQUIRE(false, SA_ERR_ARG, "sa_sobol_accumulate_f64: unknown algo %d", algo);
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_sobol_accumulate_f64(void *stream, const double *d_Yn, const double *d_Yt, uint64_t N, int Dg,
                                       int second_order, const uint8_t *d_w, uint64_t count, int nsplit,
                                       double *d_psum) {
  return sa_sobol_accumulate_ex_f64(stream, d_Yn, d_Yt, N, Dg, second_order, d_w, count, nsplit, d_psum, 0);
}

extern "C" int sa_sobol_finalize_f64(void *stream, const double *d_psum, uint64_t N, int Dg, int second_order,
                                     uint64_t count, int nsplit, const double *d_stats, double *d_est) {
  SA_REQUIRE(d_psum && d_stats && d_est, SA_ERR_ARG, "sa_sobol_finalize_f64: null pointer");
  SA_REQUIRE(N >= 1 && Dg >= 1 && count >= 1 && nsplit >= 1, SA_ERR_ARG, "sa_sobol_finalize_f64: bad shape");
  SA_REQUIRE(count < (1ull << 31), SA_ERR_UNSUPPORTED, "too many resamples in one call");
  const size_t fin_smem = ((size_t)Dg + SA_NACC(Dg, second_order)) * sizeof(double);
  if (fin_smem <= 200 * 1024) {
    // the attribute is per device: set it whenever the opt-in is needed (no process-wide cache)
    if (fin_smem > 48 * 1024)
      SA_CUDA(cudaFuncSetAttribute(finalize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    finalize_kernel<true><<<(unsigned)count, 128, fin_smem, (cudaStream_t)stream>>>(
        d_psum, N, Dg, second_order, count, nsplit, d_stats, d_est, SA_NACC(Dg, second_order),
        SA_NEST(Dg, second_order));
  } else {  // very wide problems: read the partial sums in place
    SA_REQUIRE((size_t)Dg * sizeof(double) <= 48 * 1024, SA_ERR_UNSUPPORTED, "sa_sobol_finalize_f64: Dg too large");
    finalize_kernel<false><<<(unsigned)count, 128, (size_t)Dg * sizeof(double), (cudaStream_t)stream>>>(
        d_psum, N, Dg, second_order, count, nsplit, d_stats, d_est, SA_NACC(Dg, second_order),
        SA_NEST(Dg, second_order));
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_sobol_conf_f64(void *stream, const double *d_est, uint64_t R, int ncols, double z,
                                 double *d_conf) {
  SA_REQUIRE(d_est && d_conf, SA_ERR_ARG, "sa_sobol_conf_f64: null pointer");
  SA_REQUIRE(R >= 1 && ncols >= 1, SA_ERR_ARG, "sa_sobol_conf_f64: bad shape");
  dim3 block(32, 16);
  conf_kernel<<<(unsigned)((ncols + 31) / 32), block, 0, (cudaStream_t)stream>>>(d_est, R, ncols, z, d_conf);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_col_moment_f64(void *stream, const double *d_est, uint64_t R, int ncols, const double *d_mean,
                                 double *d_out) {
  SA_REQUIRE(d_out && ncols >= 1 && (d_est || R == 0), SA_ERR_ARG, "sa_col_moment_f64: bad arguments");
  dim3 block(32, 16);
  col_moment_kernel<<<(unsigned)((ncols + 31) / 32), block, 0, (cudaStream_t)stream>>>(d_est, R, ncols, d_mean, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_conf_finish_f64(void *stream, double *d_x, int ncols, double scale, double z, uint64_t R,
                                  int finish) {
  SA_REQUIRE(d_x && ncols >= 1, SA_ERR_ARG, "sa_conf_finish_f64: bad arguments");
  conf_finish_kernel<<<(unsigned)((ncols + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, ncols, scale, z, R, finish);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_conf_local_f64(void *stream, const double *d_est, uint64_t R, int ncols, int slices, double *d_rec) {
  SA_REQUIRE(d_rec && ncols >= 1 && slices >= 1 && slices <= 65535 && (d_est || R == 0), SA_ERR_ARG,
             "sa_conf_local_f64: bad arguments");
  dim3 block(32, 16), grid((unsigned)((ncols + 31) / 32), (unsigned)slices);
  conf_local_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_est, R, ncols, d_rec);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_conf_combine_f64(void *stream, const double *d_recs, int k, uint64_t pitch, int ncols, int nextra,
                                   double z, double *d_conf, double *d_extra_sum) {
  SA_REQUIRE(d_recs && d_conf && k >= 1 && ncols >= 1 && nextra >= 0 && pitch >= (uint64_t)(1 + 2 * ncols + nextra) &&
                 (d_extra_sum || nextra == 0),
             SA_ERR_ARG, "sa_conf_combine_f64: bad arguments");
  conf_combine_kernel<<<(unsigned)((ncols + nextra + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_recs, k, (size_t)pitch, ncols, nextra, z, d_conf, d_extra_sum);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_combine_moments_pitch_f64(void *stream, const double *d_parts, int k, uint64_t pitch, double *d_out) {
  SA_REQUIRE(d_parts && d_out && k >= 1 && pitch >= 9, SA_ERR_ARG, "sa_combine_moments_f64: bad arguments");
  combine_moments_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_parts, k, (size_t)pitch, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_combine_moments_f64(void *stream, const double *d_parts, int k, double *d_out) {
  return sa_combine_moments_pitch_f64(stream, d_parts, k, 9, d_out);
}

extern "C" int sa_rank_record_f64(void *stream, const double *d_parts, int k, double f0, double f1, double f2,
                                  double *d_rec) {
  SA_REQUIRE(d_parts && d_rec && k >= 1, SA_ERR_ARG, "sa_rank_record_f64: bad arguments");
  rank_record_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_parts, k, f0, f1, f2, d_rec);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

// ---- DGSM / finite_diff ------------------------------------------------------------
extern "C" int sa_chunked_feature_doubles(int nfeat) { return chunk_width(nfeat) * chunk_count(nfeat); }

extern "C" int sa_finite_diff_expand_f64(void *stream, const double *d_base, uint64_t n, int D, double delta,
                                         const double *d_lo, const double *d_hi, double *d_out) {
  SA_REQUIRE(d_base && d_lo && d_hi && d_out && D >= 1, SA_ERR_ARG, "sa_finite_diff_expand_f64: bad arguments");
  if (n == 0) return SA_OK;
  uint64_t blocks = ceil_div_u64(n * (uint64_t)(D + 1) * D, 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  finite_diff_expand_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_base, n, D, delta, d_lo, d_hi, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_dgsm_features_f64(void *stream, const double *d_X, const double *d_Y, uint64_t N, int D,
                                    double *d_F) {
  SA_REQUIRE(d_X && d_Y && d_F && N >= 1 && D >= 1, SA_ERR_ARG, "sa_dgsm_features_f64: bad arguments");
  uint64_t blocks = ceil_div_u64(N * (uint64_t)sa_chunked_feature_doubles(2 * D), 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  dgsm_features_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, d_Y, N, D, d_F);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_feature_sums_f64(void *stream, const double *d_F, uint64_t N, int nfeat, const uint8_t *d_w,
                                   uint64_t count, int nsplit, double *d_psum) {
  SA_REQUIRE(d_F && d_psum && N >= 1 && nfeat >= 1 && count >= 1 && nsplit >= 1 && nsplit <= 65535, SA_ERR_ARG,
             "sa_feature_sums_f64: bad arguments");
  SA_REQUIRE(d_w != nullptr || count == 1, SA_ERR_ARG, "sa_feature_sums_f64: unweighted sums need count == 1");
  uint64_t rps = ceil_div_u64(N, nsplit);
  rps = (rps + 31) & ~(uint64_t)31;
  return launch_feature_gemv((cudaStream_t)stream, d_F, N, nfeat, d_w, pad_rows(N), count, nsplit, rps, d_psum);
}

extern "C" int sa_feature_centered_ss_f64(void *stream, const double *d_F, uint64_t N, int nfeat, int ncols,
                                          const double *d_center, double *d_out) {
  SA_REQUIRE(d_F && d_center && d_out && N >= 1 && nfeat >= 1 && ncols >= 1 && ncols <= nfeat, SA_ERR_ARG,
             "sa_feature_centered_ss_f64: bad arguments");
  feature_centered_ss_kernel<<<(unsigned)ncols, 256, 0, (cudaStream_t)stream>>>(d_F, N, nfeat, d_center, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_reduce_splits_f64(void *stream, const double *d_psum, uint64_t count, int nfeat, int nsplit,
                                    double scale, double *d_out) {
  SA_REQUIRE(d_psum && d_out && count >= 1 && nfeat >= 1 && nsplit >= 1, SA_ERR_ARG, "sa_reduce_splits_f64: bad arguments");
  uint64_t blocks = ceil_div_u64(count * (uint64_t)nfeat, 256);
  if (blocks > (uint64_t)sm_count() * 8) blocks = (uint64_t)sm_count() * 8;
  reduce_splits_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_psum, count, nfeat, nsplit, scale, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
