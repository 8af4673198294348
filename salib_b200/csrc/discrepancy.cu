// Discrepancy sensitivity indices: the reference's analyze/discrepancy.py:88-116 calls
// scipy.stats.qmc.discrepancy on the 2-D projections [X_i, Y] (unit-scaled) for every input i.  All four
// measures are  c0 - c1/n * sum_a g(x_ai) g(y_a) + 1/n^2 * sum_a sum_b k(x_ai, x_bi) k(y_a, y_b)  with a
// per-method pair kernel k and point kernel g (Hickernell 1998; Zhou et al. 2013; Warnock 1972, as published
// in scipy's qmc.discrepancy documentation).  The double sum is the work: N^2 pairs x D inputs of FP64.
//
// Mapping: thread <-> point a (256 per block), IC inputs of a chunk in registers; points b stream through a
// shared-memory tile of 64 and are read with warp-uniform LDS.128 broadcasts, so a (pair, input) costs 3-5
// FP64 instructions and ~1/32 of a shared-memory wavefront.  k(y_a, y_b) is computed once per pair and reused
// by the IC inputs.  The pair sum is symmetric: only b-tiles at or above the a-tile are visited, tiles
// strictly above count twice.  Per-block partial sums are reduced in a fixed order (deterministic).
#include <math.h>

#include "common.cuh"

namespace sa {

constexpr int kDiscTA = 256;  // points a per block
constexpr int kDiscTB = 64;   // points b per shared-memory tile (kDiscTA is a multiple of it)

enum { kWD = 0, kCD = 1, kMD = 2, kL2 = 3 };

// per-point term that the pair kernel of CD / MD adds for u and for v
template <int M>
__device__ __forceinline__ double disc_aux(double u) {
  if (M == kCD) return 0.5 + 0.5 * fabs(u - 0.5);
  if (M == kMD) return 0.9375 - 0.25 * fabs(u - 0.5);
  return 0.0;
}

template <int M>
__device__ __forceinline__ double disc_pair(double u, double v, double au, double av) {
  if (M == kWD) {  // 3/2 - |u-v| (1 - |u-v|)
    const double t = fabs(u - v);
    return fma(-t, 1.0 - t, 1.5);
  }
  if (M == kCD) {  // 1 + |u-1/2|/2 + |v-1/2|/2 - |u-v|/2
    return fma(-0.5, fabs(u - v), au + av);
  }
  if (M == kMD) {  // 15/8 - |u-1/2|/4 - |v-1/2|/4 - 3|u-v|/4 + |u-v|^2/2
    const double t = fabs(u - v);
    return fma(t, fma(0.5, t, -0.75), au + av);
  }
  return 1.0 - fmax(u, v);  // L2-star
}

template <int M>
__device__ __forceinline__ double disc_point(double u) {
  const double d = fabs(u - 0.5);
  if (M == kCD) return 1.0 + 0.5 * d - 0.5 * d * d;
  if (M == kMD) return 5.0 / 3.0 - 0.25 * d - 0.25 * d * d;
  if (M == kL2) return 1.0 - u * u;
  return 0.0;
}

struct DiscPlan {
  int ic, nchunk, dpad, nz;
  uint64_t np, ntile_a, nbt;
};

static DiscPlan disc_plan(uint64_t N, int D) {
  DiscPlan p;
  p.ic = D >= 13 ? 16 : 4;
  p.nchunk = (D + p.ic - 1) / p.ic;
  p.dpad = p.nchunk * p.ic;
  p.np = ceil_div_u64(N, kDiscTA) * kDiscTA;
  p.ntile_a = p.np / kDiscTA;
  p.nbt = ceil_div_u64(N, kDiscTB);
  const uint64_t base = p.ntile_a * (uint64_t)p.nchunk;
  uint64_t nz = ceil_div_u64((uint64_t)4 * sm_count(), base);
  if (nz > p.nbt) nz = p.nbt;
  if (nz < 1) nz = 1;
  p.nz = (int)nz;
  return p;
}

// ---- min / max of Y (one block; N is at most a few million here) ---------------------------------------
__global__ void __launch_bounds__(1024) disc_minmax_kernel(const double *__restrict__ Y, uint64_t N,
                                                           double *__restrict__ mm) {
  double lo = INFINITY, hi = -INFINITY;
  bool nan = false;
  for (uint64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double y = Y[i];
    nan |= (y != y);
    lo = fmin(lo, y);
    hi = fmax(hi, y);
  }
  __shared__ double slo[32], shi[32];
  __shared__ int snan;
  if (threadIdx.x == 0) snan = 0;
  __syncthreads();
  if (nan) atomicOr(&snan, 1);
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = warp_min(slo[threadIdx.x]);
    hi = warp_max(shi[threadIdx.x]);
    if (threadIdx.x == 0) {
      mm[0] = snan ? NAN : lo;
      mm[1] = snan ? NAN : hi;
    }
  }
}

// ---- unit scaling + transpose: Xt[i][a] = (X[a][i] - lb_i) / (ub_i - lb_i), Yu[a] = (Y - min) / (max - min) ----
__global__ void __launch_bounds__(256)
disc_prepare_kernel(const double *__restrict__ X, const double *__restrict__ Y, uint64_t N, int D,
                    const double *__restrict__ lb, const double *__restrict__ ub, const double *__restrict__ mm,
                    uint64_t Np, double *__restrict__ Xt, double *__restrict__ Yu, int32_t *__restrict__ flag) {
  const uint64_t total = N * (uint64_t)D;
  const double ymin = mm[0], ymax = mm[1];
  if (blockIdx.x == 0 && threadIdx.x == 0 && !(ymin < ymax) && flag) atomicOr(flag, 2);
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t a = idx / (uint64_t)D;
    const int i = (int)(idx - a * (uint64_t)D);
    const double x = X[idx], l = lb[i], u = ub[i];
    if (!(x >= l && x <= u) && flag) atomicOr(flag, 1);  // qmc.scale(reverse=True): "Sample is out of bounds"
    Xt[(uint64_t)i * Np + a] = (x - l) / (u - l);
    if (i == 0) Yu[a] = (Y[a] - ymin) / (ymax - ymin);
  }
}

// ---- the pair sums ---------------------------------------------------------------------------------------
template <int M, int IC>
__global__ void __launch_bounds__(kDiscTA)
disc_pairs_kernel(const double *__restrict__ Xt, const double *__restrict__ Yu, uint64_t N, uint64_t Np, int Dpad,
                  double *__restrict__ part) {
  constexpr bool kAux = (M == kCD || M == kMD);
  __shared__ __align__(16) double xb[IC / 2][kDiscTB][2];
  __shared__ __align__(16) double ab[kAux ? IC / 2 : 1][kAux ? kDiscTB : 1][2];
  __shared__ __align__(16) double yb[kDiscTB][2];  // {y_b, aux(y_b)}
  __shared__ double red[kDiscTA / 32][IC];

  const int tid = threadIdx.x;
  const uint64_t tile_a = blockIdx.x;
  const int chunk = blockIdx.y;
  const int nz = gridDim.z;
  const uint64_t a = tile_a * kDiscTA + tid;
  const double *Xc = Xt + (uint64_t)chunk * IC * Np;

  double xa[IC], aa[IC], acc[IC];
#pragma unroll
  for (int q = 0; q < IC; ++q) {
    xa[q] = Xc[(uint64_t)q * Np + a];
    aa[q] = disc_aux<M>(xa[q]);
    acc[q] = 0.0;
  }
  const double ya = Yu[a];
  const double aya = disc_aux<M>(ya);

  const uint64_t nbt = ceil_div_u64(N, kDiscTB);
  const uint64_t first = tile_a * (kDiscTA / kDiscTB);
  for (uint64_t jb = first + blockIdx.z; jb < nbt; jb += nz) {
    const uint64_t B0 = jb * kDiscTB;
    __syncthreads();
    for (int e = tid; e < IC * kDiscTB; e += kDiscTA) {
      const int q = e / kDiscTB, b = e - q * kDiscTB;
      const double v = Xc[(uint64_t)q * Np + B0 + b];
      xb[q >> 1][b][q & 1] = v;
      if (kAux) ab[q >> 1][b][q & 1] = disc_aux<M>(v);
    }
    if (tid < kDiscTB) {
      const double v = Yu[B0 + tid];
      yb[tid][0] = v;
      yb[tid][1] = disc_aux<M>(v);
    }
    __syncthreads();
    const double w = (jb >= first + kDiscTA / kDiscTB) ? 2.0 : 1.0;  // tiles strictly above the diagonal block
    const int nb = (int)((N - B0 < (uint64_t)kDiscTB) ? N - B0 : (uint64_t)kDiscTB);
#pragma unroll 2
    for (int b = 0; b < nb; ++b) {
      const double2 yv = *reinterpret_cast<const double2 *>(yb[b]);
      const double ky = w * disc_pair<M>(ya, yv.x, aya, yv.y);
#pragma unroll
      for (int qq = 0; qq < IC / 2; ++qq) {
        const double2 xv = *reinterpret_cast<const double2 *>(xb[qq][b]);
        double2 av = make_double2(0.0, 0.0);
        if (kAux) av = *reinterpret_cast<const double2 *>(ab[qq][b]);
        acc[2 * qq] = fma(disc_pair<M>(xa[2 * qq], xv.x, aa[2 * qq], av.x), ky, acc[2 * qq]);
        acc[2 * qq + 1] = fma(disc_pair<M>(xa[2 * qq + 1], xv.y, aa[2 * qq + 1], av.y), ky, acc[2 * qq + 1]);
      }
    }
  }

  const bool valid = a < N;
#pragma unroll
  for (int q = 0; q < IC; ++q) {
    const double v = warp_sum(valid ? acc[q] : 0.0);
    if ((tid & 31) == 0) red[tid >> 5][q] = v;
  }
  __syncthreads();
  if (tid < IC) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < kDiscTA / 32; ++wv) v += red[wv][tid];
    part[(tile_a * (uint64_t)nz + blockIdx.z) * (uint64_t)Dpad + (uint64_t)chunk * IC + tid] = v;
  }
}

// ---- per input: reduce the partial sums, add the single sum, apply the closed form (d = 2) ----------------
template <int M>
__global__ void __launch_bounds__(256)
disc_final_kernel(const double *__restrict__ part, uint64_t nparts, int Dpad, const double *__restrict__ Xt,
                  const double *__restrict__ Yu, uint64_t N, uint64_t Np, double c0, double c1,
                  double *__restrict__ disc) {
  const int i = blockIdx.x;
  double s2 = 0.0, s1 = 0.0;
  for (uint64_t p = threadIdx.x; p < nparts; p += blockDim.x) s2 += part[p * (uint64_t)Dpad + i];
  if (M != kWD)
    for (uint64_t a = threadIdx.x; a < N; a += blockDim.x)
      s1 = fma(disc_point<M>(Xt[(uint64_t)i * Np + a]), disc_point<M>(Yu[a]), s1);
  __shared__ double sh[2][8];
  s2 = warp_sum(s2);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s2;
    sh[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s2 = s1 = 0.0;
    for (int wv = 0; wv < 8; ++wv) {
      s2 += sh[0][wv];
      s1 += sh[1][wv];
    }
    const double n = (double)N;
    double v = c0 - c1 / n * s1 + s2 / (n * n);
    if (M == kWD) v = -c0 + s2 / (n * n);
    if (M == kL2) v = sqrt(v);
    disc[i] = v;
  }
}

// s_i = disc_i / sum(disc) (analyze/discrepancy.py:98); groups: mean over the members (:100-108)
__global__ void disc_normalize_kernel(const double *__restrict__ disc, int D, int Dg, const int32_t *__restrict__ gcol,
                                      double *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double tot = 0.0;
  for (int i = 0; i < D; ++i) tot += disc[i];
  if (!gcol) {
    for (int i = 0; i < D; ++i) out[i] = disc[i] / tot;
    return;
  }
  for (int g = 0; g < Dg; ++g) {
    double s = 0.0;
    int cnt = 0;
    for (int i = 0; i < D; ++i)
      if (gcol[i] == g) {
        s += disc[i] / tot;
        ++cnt;
      }
    out[g] = s / (double)cnt;
  }
}

template <int M>
static void launch_pairs(const DiscPlan &p, cudaStream_t st, const double *Xt, const double *Yu, uint64_t N,
                         double *part) {
  dim3 grid((unsigned)p.ntile_a, (unsigned)p.nchunk, (unsigned)p.nz);
  if (p.ic == 16)
    disc_pairs_kernel<M, 16><<<grid, kDiscTA, 0, st>>>(Xt, Yu, N, p.np, p.dpad, part);
  else
    disc_pairs_kernel<M, 4><<<grid, kDiscTA, 0, st>>>(Xt, Yu, N, p.np, p.dpad, part);
}

template <int M>
static void launch_final(const DiscPlan &p, cudaStream_t st, const double *part, const double *Xt, const double *Yu,
                         uint64_t N, int D, double c0, double c1, double *disc) {
  disc_final_kernel<M><<<D, 256, 0, st>>>(part, p.ntile_a * (uint64_t)p.nz, p.dpad, Xt, Yu, N, p.np, c0, c1, disc);
}

static size_t disc_work_doubles(const DiscPlan &p, int D) {
  return (size_t)p.dpad * p.np + p.np + (size_t)p.ntile_a * p.nz * p.dpad + (size_t)D + 2;
}

}  // namespace sa

using namespace sa;

extern "C" size_t sa_discrepancy_workspace_bytes(uint64_t N, int D) {
  if (N < 1 || D < 1) return 0;
  return disc_work_doubles(disc_plan(N, D), D) * sizeof(double);
}

extern "C" int sa_discrepancy_f64(void *stream, const double *d_X, const double *d_Y, uint64_t N, int D,
                                  const double *d_lb, const double *d_ub, int method, int Dg,
                                  const int32_t *d_gcol, double *d_out, double *d_disc, void *d_work,
                                  size_t work_bytes, int32_t *d_flag) {
  SA_REQUIRE(d_X && d_Y && d_lb && d_ub && d_out && d_work, SA_ERR_ARG, "sa_discrepancy_f64: null pointer");
  SA_REQUIRE(N >= 1 && D >= 1 && D <= 65535, SA_ERR_ARG, "sa_discrepancy_f64: bad N/D");
  SA_REQUIRE(method >= 0 && method <= 3, SA_ERR_ARG, "sa_discrepancy_f64: method must be 0 (WD), 1 (CD), 2 (MD), 3 (L2-star)");
  SA_REQUIRE(d_gcol ? (Dg >= 1 && Dg <= D) : Dg == D, SA_ERR_ARG, "sa_discrepancy_f64: bad group count");
  SA_REQUIRE(N <= ((uint64_t)1 << 26), SA_ERR_ARG, "sa_discrepancy_f64: N too large for the N^2 pair sum");
  const DiscPlan p = disc_plan(N, D);
  SA_REQUIRE(work_bytes >= disc_work_doubles(p, D) * sizeof(double), SA_ERR_WORKSPACE,
             "sa_discrepancy_f64: workspace too small");
  SA_REQUIRE(((uintptr_t)d_work & 15) == 0, SA_ERR_ARG, "sa_discrepancy_f64: workspace not 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  double *Xt = (double *)d_work;
  double *Yu = Xt + (size_t)p.dpad * p.np;
  double *part = Yu + p.np;
  double *disc = part + (size_t)p.ntile_a * p.nz * p.dpad;
  double *mm = disc + D;
  SA_CUDA(cudaMemsetAsync(Xt, 0, ((size_t)p.dpad * p.np + p.np) * sizeof(double), st));
  disc_minmax_kernel<<<1, 1024, 0, st>>>(d_Y, N, mm);
  uint64_t blocks = ceil_div_u64(N * (uint64_t)D, 256);
  if (blocks > (uint64_t)sm_count() * 16) blocks = (uint64_t)sm_count() * 16;
  disc_prepare_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_X, d_Y, N, D, d_lb, d_ub, mm, p.np, Xt, Yu, d_flag);
  double c0 = 0.0, c1 = 2.0;
  switch (method) {
    case kWD:
      c0 = pow(4.0 / 3.0, 2.0);
      launch_pairs<kWD>(p, st, Xt, Yu, N, part);
      launch_final<kWD>(p, st, part, Xt, Yu, N, D, c0, c1, disc);
      break;
    case kCD:
      c0 = pow(13.0 / 12.0, 2.0);
      launch_pairs<kCD>(p, st, Xt, Yu, N, part);
      launch_final<kCD>(p, st, part, Xt, Yu, N, D, c0, c1, disc);
      break;
    case kMD:
      c0 = pow(19.0 / 12.0, 2.0);
      launch_pairs<kMD>(p, st, Xt, Yu, N, part);
      launch_final<kMD>(p, st, part, Xt, Yu, N, D, c0, c1, disc);
      break;
    default:
      c0 = pow(3.0, -2.0);
      c1 = pow(2.0, -1.0);
      launch_pairs<kL2>(p, st, Xt, Yu, N, part);
      launch_final<kL2>(p, st, part, Xt, Yu, N, D, c0, c1, disc);
      break;
  }
  disc_normalize_kernel<<<1, 32, 0, st>>>(disc, D, Dg, d_gcol, d_out);
  if (d_disc) SA_CUDA(cudaMemcpyAsync(d_disc, disc, (size_t)D * sizeof(double), cudaMemcpyDeviceToDevice, st));
  note_launches(5);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
