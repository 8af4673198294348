// Saltelli sample generation: Sobol' draw (Joe-Kuo direction vectors, Gray-code skip-ahead)
// fused with the A/AB/BA/B cross-matrix build and the uniform bounds scaling.
//
// Reference behaviour replaced: src/SALib/sample/sobol.py:107-188 (and saltelli.py:147-203).
//   point k:   x_k[d] = shift[d] ^ XOR_{b in gray(k)} sv[d][b];  value = double(x) * 2^-30
//   step k->k+1: x ^= sv[d][ctz(k+1)]                           (sobol_sequence.py:86-89)
//   scaling:   round(round(v * span) + lb)                       (util/__init__.py:49-53)
//   layout:    base row i -> rows i*step + {0: A, 1+k: AB_k, 1+Dg+k: BA_k, step-1: B}
//
// The kernel is bound by HBM writes (step*D*8 bytes per base row, nothing read but a
// [2D][30] u32 table), so the design is about issuing 16-byte, fully coalesced, streaming
// stores with as little integer work per store as possible:
//   * "register path" (D even and D/2 | 32 or 32 | D/2): a warp owns a run of consecutive base
//     rows; each lane owns one column pair, keeps the four Sobol' states (a0,a1,b0,b1) in
//     registers, advances them with one XOR per row and writes every slot with st.global.cs.v2.f64.
//   * "generic path" (any D): the block stages the scaled A/B rows of a chunk in shared memory and
//     writes the chunk's contiguous output region flat, decoding (row, slot, col) with a
//     multiply-shift division.
#include <type_traits>

#include "common.cuh"

namespace sa {

constexpr double kTwoPowMinus30 = 1.0 / 1073741824.0;

// Closed form for point k in dimension d.
__device__ __forceinline__ uint32_t sobol_point(const uint32_t *__restrict__ sv, uint32_t shift,
                                                int d, uint64_t k) {
  uint32_t g = (uint32_t)(k ^ (k >> 1));
  uint32_t x = shift;
  const uint32_t *v = sv + (size_t)d * SA_SOBOL_BITS;
#pragma unroll
  for (int b = 0; b < SA_SOBOL_BITS; ++b) {
    if ((g >> b) & 1u) x ^= __ldg(v + b);
  }
  return x;
}

__device__ __forceinline__ double scale_value(uint32_t x, double span, double lb) {
  // two roundings, never an FMA
  return __dadd_rn(__dmul_rn((double)x * kTwoPowMinus30, span), lb);
}

// Per-column scaling: uniform bounds (code == nullptr; util/__init__.py:49-53) or the inverse CDF of
// the column's distribution (util/__init__.py:126-289, scipy.stats ppf restated).  Distribution
// codes and the 4 parameters per column are prepared by the host (salib_b200/sample/sobol.py):
//   0 unif      lb, span                        u*span + lb
//   1 triang    loc, scale, c                   loc + scale * (u < c ? sqrt(c u) : 1 - sqrt((1-c)(1-u)))
//   2 logunif   log a, log b - log a            exp(log a + u (log b - log a))
//   3 norm      loc, scale                      loc + scale * ndtri(u)
//   4 truncnorm (a < 0)  Phi(a),  mass, loc, scale    loc + scale * ndtri(Phi(a) + u mass)
//   5 truncnorm (a >= 0) Phi(-b), mass, loc, scale    loc - scale * ndtri(Phi(-b) + (1-u) mass)
//   6 lognorm   loc, scale                      exp(loc + scale * ndtri(u))
//   7 weibull   1/c, scale, loc                 loc + scale * (-log1p(-u))^(1/c)
struct ColScale {
  const double *lb, *span;
  const int32_t *code;  // nullptr: every column uniform
  const double *par;    // [D][4]
};

__device__ __forceinline__ double dist_transform(double u, int code, const double *__restrict__ p) {
  switch (code) {
    case 1: {
      const double c = p[2];
      const double t = (u < c) ? sqrt(__dmul_rn(c, u)) : __dsub_rn(1.0, sqrt(__dmul_rn(__dsub_rn(1.0, c), __dsub_rn(1.0, u))));
      return __dadd_rn(__dmul_rn(t, p[1]), p[0]);
    }
    case 2: return exp(__dadd_rn(p[0], __dmul_rn(u, p[1])));
    case 3: return __dadd_rn(__dmul_rn(normcdfinv(u), p[1]), p[0]);
    case 4: return __dadd_rn(__dmul_rn(normcdfinv(__dadd_rn(p[0], __dmul_rn(u, p[1]))), p[3]), p[2]);
    case 5: return __dadd_rn(__dmul_rn(-normcdfinv(__dadd_rn(p[0], __dmul_rn(__dsub_rn(1.0, u), p[1]))), p[3]), p[2]);
    case 6: return exp(__dadd_rn(__dmul_rn(normcdfinv(u), p[1]), p[0]));
    case 7: return __dadd_rn(__dmul_rn(pow(-log1p(-u), p[0]), p[1]), p[2]);
    default: return __dadd_rn(__dmul_rn(u, p[1]), p[0]);
  }
}

__device__ __forceinline__ double scale_col(uint32_t x, int col, const ColScale &cs) {
  if (cs.code == nullptr) return scale_value(x, cs.span[col], cs.lb[col]);
  return dist_transform((double)x * kTwoPowMinus30, cs.code[col], cs.par + 4 * (size_t)col);
}

// Does slot s take column (with group g) from B?  sample/sobol.py:149-186
__device__ __forceinline__ bool slot_uses_b(int s, int g, int Dg, int step, int second_order) {
  const int k1 = s - 1;
  const int k2 = s - 1 - Dg;
  const bool in_ab = (unsigned)k1 < (unsigned)Dg;
  const bool in_ba = second_order && ((unsigned)k2 < (unsigned)Dg);
  const bool last = (s == step - 1);
  return last | (in_ab & (g == k1)) | (in_ba & (g != k2));
}

// ------------------------------------------------------------------------------------------
// Register path
// ------------------------------------------------------------------------------------------
template <bool GROUPED, bool DISTS, int MINB>  // MINB: resident blocks per SM asked of the compiler
__global__ void __launch_bounds__(256, MINB)
saltelli_reg_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
                    uint64_t first_index, uint64_t n, int D, int Dg,
                    const int32_t *__restrict__ gcol, int step, int second_order, const ColScale cs,
                    double *__restrict__ out, int rows_per_warp, int cb_split) {
  // cb_split > 1 (few rows of a wide design): the 64-column blocks of a row run are dealt to cb_split
  // neighbouring warps instead of being walked by one -- a chunk of 1024 rows at D = 1024 is 16 K warps
  // instead of 1 K, and the warps of a block write neighbouring 512-byte pieces of the same rows
  const int lane = threadIdx.x & 31;
  const uint64_t gwc = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t gw = gwc / (uint64_t)cb_split;
  const int cb_part = (int)(gwc - gw * (uint64_t)cb_split);
  const uint64_t r0 = gw * (uint64_t)rows_per_warp;
  if (r0 >= n) return;
  const uint64_t r1 = (r0 + rows_per_warp < n) ? r0 + rows_per_warp : n;

  const int halfD = D >> 1;
  int ncb, nsub, sub, jp_lane;
  if (halfD >= 32) {
    ncb = halfD >> 5;
    nsub = 1;
    sub = 0;
    jp_lane = lane;
  } else {
    ncb = 1;
    nsub = 32 / halfD;
    sub = lane / halfD;
    jp_lane = lane - sub * halfD;
  }

  const int cb_per = ncb / cb_split;  // the launcher picks a divisor of ncb
  for (int cb = cb_part * cb_per; cb < (cb_part + 1) * cb_per; ++cb) {
    const int j0 = 2 * (cb * 32 + jp_lane);
    const int j1 = j0 + 1;
    const int g0 = GROUPED ? gcol[j0] : j0;
    const int g1 = GROUPED ? gcol[j1] : j1;
    double lb0 = 0.0, lb1 = 0.0, sp0 = 0.0, sp1 = 0.0;
    if (!DISTS) {
      lb0 = cs.lb[j0];
      lb1 = cs.lb[j1];
      sp0 = cs.span[j0];
      sp1 = cs.span[j1];
    }
    const uint32_t *va0 = sv + (size_t)j0 * SA_SOBOL_BITS;
    const uint32_t *va1 = sv + (size_t)j1 * SA_SOBOL_BITS;
    const uint32_t *vb0 = sv + (size_t)(D + j0) * SA_SOBOL_BITS;
    const uint32_t *vb1 = sv + (size_t)(D + j1) * SA_SOBOL_BITS;
    uint64_t k = first_index + r0;
    uint32_t xa0 = sobol_point(sv, shift ? shift[j0] : 0u, j0, k);
    uint32_t xa1 = sobol_point(sv, shift ? shift[j1] : 0u, j1, k);
    uint32_t xb0 = sobol_point(sv, shift ? shift[D + j0] : 0u, D + j0, k);
    uint32_t xb1 = sobol_point(sv, shift ? shift[D + j1] : 0u, D + j1, k);

    for (uint64_t r = r0; r < r1; ++r) {
      const double a0 = DISTS ? scale_col(xa0, j0, cs) : scale_value(xa0, sp0, lb0);
      const double a1 = DISTS ? scale_col(xa1, j1, cs) : scale_value(xa1, sp1, lb1);
      const double b0 = DISTS ? scale_col(xb0, j0, cs) : scale_value(xb0, sp0, lb0);
      const double b1 = DISTS ? scale_col(xb1, j1, cs) : scale_value(xb1, sp1, lb1);
      double *row = out + (r * (uint64_t)step) * (uint64_t)D + j0;
#pragma unroll 4
      for (int s = sub; s < step; s += nsub) {
        const bool u0 = slot_uses_b(s, g0, Dg, step, second_order);
        const bool u1 = slot_uses_b(s, g1, Dg, step, second_order);
        double2 v = make_double2(u0 ? b0 : a0, u1 ? b1 : a1);
        __stcs(reinterpret_cast<double2 *>(row + (size_t)s * D), v);
      }
      // advance to point k+1: flip direction number ctz(k+1)
      ++k;
      const int c = __ffsll((long long)k) - 1;
      if (c < SA_SOBOL_BITS) {
        xa0 ^= __ldg(va0 + c);
        xa1 ^= __ldg(va1 + c);
        xb0 ^= __ldg(vb0 + c);
        xb1 ^= __ldg(vb1 + c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Generic path
// ------------------------------------------------------------------------------------------
// Stage the scaled A/B values of base rows [base, base+nv) in shared memory: sa/sb [BR][D].
__device__ __forceinline__ void stage_rows(const uint32_t *__restrict__ sv,
                                           const uint32_t *__restrict__ shift, uint64_t k0, int nv,
                                           int D, const ColScale &cs, double *sa, double *sb, int ld = 0) {
  if (ld == 0) ld = D;  // row pitch of sa / sb in doubles
  const int dims = 2 * D;
  int nseg = blockDim.x / dims;
  if (nseg < 1) nseg = 1;
  if (nseg > nv) nseg = nv;
  const int seglen = (nv + nseg - 1) / nseg;
  for (int item = threadIdx.x; item < dims * nseg; item += blockDim.x) {
    const int seg = item / dims;
    const int d = item - seg * dims;
    const int rs = seg * seglen;
    const int re = (rs + seglen < nv) ? rs + seglen : nv;
    if (rs >= re) continue;
    const int col = (d < D) ? d : d - D;
    double *dst = (d < D) ? sa : sb;
    const uint32_t *v = sv + (size_t)d * SA_SOBOL_BITS;
    uint64_t k = k0 + rs;
    uint32_t x = sobol_point(sv, shift ? shift[d] : 0u, d, k);
    for (int r = rs; r < re; ++r) {
      dst[(size_t)r * ld + col] = scale_col(x, col, cs);
      ++k;
      const int c = __ffsll((long long)k) - 1;
      if (c < SA_SOBOL_BITS) x ^= __ldg(v + c);
    }
  }
}

template <bool GROUPED>
__global__ void __launch_bounds__(256)
saltelli_generic_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
                        uint64_t first_index, uint64_t n, int D, int Dg,
                        const int32_t *__restrict__ gcol, int step, int second_order, const ColScale cs,
                        double *__restrict__ out, int BR, uint32_t magic_rowlen, uint32_t magic_D) {
  extern __shared__ double smem[];
  double *sa = smem;
  double *sb = smem + (size_t)BR * D;
  const uint32_t rowlen = (uint32_t)step * (uint32_t)D;
  const uint64_t nchunks = ceil_div_u64(n, BR);
  for (uint64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const uint64_t base = chunk * (uint64_t)BR;
    const int nv = (int)((n - base < (uint64_t)BR) ? (n - base) : BR);
    stage_rows(sv, shift, first_index + base, nv, D, cs, sa, sb);
    __syncthreads();
    const uint32_t total = (uint32_t)nv * rowlen;
    double *dst = out + base * (uint64_t)rowlen;
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
      const uint32_t r = fastdiv(e, rowlen, magic_rowlen);
      const uint32_t q = e - r * rowlen;
      const uint32_t s = fastdiv(q, (uint32_t)D, magic_D);
      const uint32_t j = q - s * (uint32_t)D;
      const int g = GROUPED ? gcol[j] : (int)j;
      const bool ub = slot_uses_b((int)s, g, Dg, step, second_order);
      const double v = ub ? sb[(size_t)r * D + j] : sa[(size_t)r * D + j];
      __stcs(dst + e, v);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Test functions (X -> Y)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double ishigami_value(double x0, double x1, double x2, double A, double B) {
  // Ishigami.py:55-59: sin(x0) + A*sin(x1)^2 + B*x2^4*sin(x0), evaluated left to right
  const double s0 = sin(x0);
  const double s1 = sin(x1);
  const double x2sq = __dmul_rn(x2, x2);
  const double t1 = __dmul_rn(A, __dmul_rn(s1, s1));
  const double t2 = __dmul_rn(__dmul_rn(B, __dmul_rn(x2sq, x2sq)), s0);
  return __dadd_rn(__dadd_rn(s0, t1), t2);
}

__global__ void eval_ishigami_kernel(const double *__restrict__ X, uint64_t rows, double A, double B,
                                     double *__restrict__ Y) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const double *x = X + 3 * i;
    Y[i] = ishigami_value(x[0], x[1], x[2], A, B);
  }
}

// One factor of the G function, Sobol_G.py:79-83:
//   ((1+alpha) * |2*frac(x+delta) - 1|^alpha + a) / (1 + a)
__device__ __forceinline__ double g_factor(double x, double a, double delta, double alpha) {
  const double sx = __dadd_rn(x, delta);
  const double fx = __dsub_rn(sx, trunc(sx));
  double t = fabs(__dsub_rn(__dmul_rn(2.0, fx), 1.0));
  if (alpha != 1.0) t = pow(t, alpha);
  return __ddiv_rn(__dadd_rn(__dmul_rn(__dadd_rn(1.0, alpha), t), a), __dadd_rn(1.0, a));
}

// LPR lanes cooperate on one row; the product is taken in column order within a lane and
// combined across lanes by a shuffle tree (rounding differs from the reference's strictly
// sequential product by a few ulp; the fused kernel below is strictly sequential).
template <int LPR>
__global__ void __launch_bounds__(256)
eval_sobol_g_kernel(const double *__restrict__ X, uint64_t rows, int D, const double *__restrict__ a,
                    const double *__restrict__ delta, const double *__restrict__ alpha,
                    double *__restrict__ Y, int32_t *flag) {
  const int sub = threadIdx.x % LPR;
  const uint64_t rows_per_block = blockDim.x / LPR;
  int bad = 0;
  for (uint64_t row = (uint64_t)blockIdx.x * rows_per_block + threadIdx.x / LPR;
       row < ceil_div_u64(rows, rows_per_block) * rows_per_block;
       row += (uint64_t)gridDim.x * rows_per_block) {
    double p = 1.0;
    if (row < rows) {
      const double *x = X + row * (uint64_t)D;
      for (int j = sub; j < D; j += LPR) {
        const double v = x[j];
        bad |= (v < 0.0) ? 1 : 0;
        bad |= (v > 1.0) ? 2 : 0;
        p = __dmul_rn(p, g_factor(v, a[j], delta ? delta[j] : 0.0, alpha ? alpha[j] : 1.0));
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) p = __dmul_rn(p, __shfl_xor_sync(kFull, p, o));
    if (sub == 0 && row < rows) Y[row] = p;
  }
  if (flag && bad) atomicOr(flag, bad);
}

// ------------------------------------------------------------------------------------------
// Fused: Sobol' indices -> Y (X never written)
// ------------------------------------------------------------------------------------------
// Block stages BR base rows (scaled a/b), converts them in place to G factors fa/fb, then each
// thread owns one (base row, slot) output and multiplies D factors in column order -- the
// same order as the reference's row.prod().
template <bool GROUPED>
__global__ void __launch_bounds__(256)
saltelli_eval_g_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
                       uint64_t first_index, uint64_t n, int D, int Dg,
                       const int32_t *__restrict__ gcol, int step, int second_order, const ColScale cs,
                       const double *__restrict__ a, const double *__restrict__ delta,
                       const double *__restrict__ alpha, double *__restrict__ Y, int32_t *flag,
                       int BR) {
  extern __shared__ double smem[];
  double *sa = smem;
  double *sb = smem + (size_t)BR * D;
  int *sg = reinterpret_cast<int *>(smem + 2 * (size_t)BR * D);
  for (int j = threadIdx.x; j < D; j += blockDim.x) sg[j] = GROUPED ? gcol[j] : j;
  const uint64_t nchunks = ceil_div_u64(n, BR);
  int bad = 0;
  for (uint64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const uint64_t base = chunk * (uint64_t)BR;
    const int nv = (int)((n - base < (uint64_t)BR) ? (n - base) : BR);
    stage_rows(sv, shift, first_index + base, nv, D, cs, sa, sb);
    __syncthreads();
    for (int e = threadIdx.x; e < nv * D; e += blockDim.x) {
      const int j = e % D;
      const double aj = a[j], dj = delta ? delta[j] : 0.0, al = alpha ? alpha[j] : 1.0;
      const double va = sa[e], vb = sb[e];
      bad |= (va < 0.0 || vb < 0.0) ? 1 : 0;
      bad |= (va > 1.0 || vb > 1.0) ? 2 : 0;
      sa[e] = g_factor(va, aj, dj, al);
      sb[e] = g_factor(vb, aj, dj, al);
    }
    __syncthreads();
    const int total = nv * step;
    for (int o = threadIdx.x; o < total; o += blockDim.x) {
      const int r = o / step;
      const int s = o - r * step;
      const double *fa = sa + (size_t)r * D;
      const double *fb = sb + (size_t)r * D;
      double p = 1.0;
      if (GROUPED || D < 32) {  // (short rows: the slot test per factor is cheaper than three divergent runs)
        for (int j = 0; j < D; ++j) {
          const bool ub = slot_uses_b(s, sg[j], Dg, step, second_order);
          p = __dmul_rn(p, ub ? fb[j] : fa[j]);
        }
      } else {
        // no groups: slot s takes every factor from one matrix except column k from the other
        // (A: k = -1; AB_k; BA_k; B) -- three plain runs, no per-factor slot test, same product order
        const bool from_b = (s == step - 1) || (second_order && s > D);
        const int k = (s == 0 || s == step - 1) ? -1 : (s <= D ? s - 1 : s - 1 - D);
        const double *main_f = from_b ? fb : fa;
        const double *other_f = from_b ? fa : fb;
        const int kk = k < 0 ? D : k;
        for (int j = 0; j < kk; ++j) p = __dmul_rn(p, main_f[j]);
        if (k >= 0) {
          p = __dmul_rn(p, other_f[k]);
          for (int j = k + 1; j < D; ++j) p = __dmul_rn(p, main_f[j]);
        }
      }
      __stcs(Y + base * (uint64_t)step + o, p);
    }
    __syncthreads();
  }
  if (flag && bad) atomicOr(flag, bad);
}

// Ungrouped problems: slot AB_k is row A with column k taken from B (BA_k the other way round), so the
// reference's left-to-right product of slot AB_k starts with the prefix product of A up to column k -- shared by
// all slots and bit-identical to the sequential loop -- and only the tail after the exchanged column is
// specific.  A thread takes the slot pair (k, D-1-k) (tails of D-k-1 and k factors: the same work for every
// thread) for kGRows base rows side by side (independent dependency chains).  Half the multiplications of the
// slot-by-slot loop and no slot test per factor.
constexpr int kGRows = 4;
__global__ void __launch_bounds__(256)
saltelli_eval_g_pairs_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
                             uint64_t first_index, uint64_t n, int D, int step, int second_order,
                             const ColScale cs, const double *__restrict__ a, const double *__restrict__ delta,
                             const double *__restrict__ alpha, double *__restrict__ Y, int32_t *flag, int BR) {
  extern __shared__ double smem[];
  const int LD = D | 1;  // odd row pitch: the row-per-lane prefix pass is free of bank conflicts
  double *fa = smem;
  double *fb = fa + (size_t)BR * LD;
  double *pa = fb + (size_t)BR * LD;
  double *pb = pa + (size_t)BR * LD;
  const uint64_t nchunks = ceil_div_u64(n, BR);
  const int nm = second_order ? 2 : 1;
  const int npair = (D + 1) / 2;
  int bad = 0;
  for (uint64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const uint64_t base = chunk * (uint64_t)BR;
    const int nv = (int)((n - base < (uint64_t)BR) ? (n - base) : BR);
    stage_rows(sv, shift, first_index + base, nv, D, cs, fa, fb, LD);
    __syncthreads();
    for (int e = threadIdx.x; e < nv * D; e += blockDim.x) {
      const int r = e / D, j = e - r * D;
      const int pos = r * LD + j;
      const double aj = a[j], dj = delta ? delta[j] : 0.0, al = alpha ? alpha[j] : 1.0;
      const double va = fa[pos], vb = fb[pos];
      bad |= (va < 0.0 || vb < 0.0) ? 1 : 0;
      bad |= (va > 1.0 || vb > 1.0) ? 2 : 0;
      fa[pos] = g_factor(va, aj, dj, al);
      fb[pos] = g_factor(vb, aj, dj, al);
    }
    __syncthreads();
    // prefix products p[j] = f[0] * ... * f[j-1] (left to right) and the A / B rows themselves
    for (int e = threadIdx.x; e < 2 * nv; e += blockDim.x) {
      const int m = e / nv, r = e - m * nv;  // lanes of a warp: consecutive rows of one matrix
      const double *f = (m ? fb : fa) + (size_t)r * LD;
      double *pp = (m ? pb : pa) + (size_t)r * LD;
      double p = 1.0;
      for (int j = 0; j < D; ++j) {
        pp[j] = p;
        p = __dmul_rn(p, f[j]);
      }
      __stcs(Y + (base + r) * (uint64_t)step + (m ? step - 1 : 0), p);
    }
    __syncthreads();
    const int ngrp = (nv + kGRows - 1) / kGRows;
    const int items = ngrp * nm * npair;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int q = it % npair;
      const int t = it / npair;
      const int m = t % nm, grp = t / nm;
      const double *main_f = m ? fb : fa;
      const double *other_f = m ? fa : fb;
      const double *pm = m ? pb : pa;
      int rc[kGRows];
#pragma unroll
      for (int u = 0; u < kGRows; ++u) {
        const int r = grp * kGRows + u;
        rc[u] = (r < nv ? r : nv - 1) * LD;  // clamp: rows past the chunk repeat the last one (not stored)
      }
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        const int k = which ? D - 1 - q : q;
        if (which && k == q) break;  // odd D: the middle column pairs with itself
        double p[kGRows];
#pragma unroll
        for (int u = 0; u < kGRows; ++u) p[u] = __dmul_rn(pm[rc[u] + k], other_f[rc[u] + k]);
        for (int j = k + 1; j < D; ++j) {
#pragma unroll
          for (int u = 0; u < kGRows; ++u) p[u] = __dmul_rn(p[u], main_f[rc[u] + j]);
        }
#pragma unroll
        for (int u = 0; u < kGRows; ++u) {
          const int r = grp * kGRows + u;
          if (r < nv) __stcs(Y + (base + r) * (uint64_t)step + 1 + m * D + k, p[u]);
        }
      }
    }
    __syncthreads();
  }
  if (flag && bad) atomicOr(flag, bad);
}

// One thread per run of kIshRun consecutive base rows: closed form once, then one XOR per coordinate and row.
// Every output row of a base row combines coordinates of A_i and B_i only, so the three terms of the model are
// evaluated once per source (4 sines per base row instead of 2 per output row) and combined per row in exactly
// the order of ishigami_value: (s0 + A*s1^2) + (B*x2^4)*s0 -- bit-identical to evaluating the rows one by one.
constexpr int kIshRun = 8;
template <bool GROUPED>
__global__ void __launch_bounds__(256)
saltelli_eval_ishigami_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift,
                              uint64_t first_index, uint64_t n, int Dg,
                              const int32_t *__restrict__ gcol, int step, int second_order, const ColScale cs,
                              double A, double B, double *__restrict__ Y) {
  const uint64_t nruns = ceil_div_u64(n, kIshRun);
  int g[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) g[j] = GROUPED ? gcol[j] : j;
  for (uint64_t run = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; run < nruns;
       run += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i0 = run * kIshRun;
    const int len = (int)((n - i0 < (uint64_t)kIshRun) ? (n - i0) : kIshRun);
    uint64_t k = first_index + i0;
    uint32_t x[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) x[d] = sobol_point(sv, shift ? shift[d] : 0u, d, k);
    for (int r = 0; r < len; ++r) {
      double s0[2], t1[2], q[2];  // [0]: from A_i, [1]: from B_i
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double x0 = scale_col(x[3 * h + 0], 0, cs);
        const double x1 = scale_col(x[3 * h + 1], 1, cs);
        const double x2 = scale_col(x[3 * h + 2], 2, cs);
        const double s1 = sin(x1);
        const double x2sq = __dmul_rn(x2, x2);
        s0[h] = sin(x0);
        t1[h] = __dmul_rn(A, __dmul_rn(s1, s1));
        q[h] = __dmul_rn(B, __dmul_rn(x2sq, x2sq));
      }
      double *dst = Y + (i0 + r) * (uint64_t)step;
      for (int s = 0; s < step; ++s) {
        const int c0 = slot_uses_b(s, g[0], Dg, step, second_order) ? 1 : 0;
        const int c1 = slot_uses_b(s, g[1], Dg, step, second_order) ? 1 : 0;
        const int c2 = slot_uses_b(s, g[2], Dg, step, second_order) ? 1 : 0;
        const double u0 = c0 ? s0[1] : s0[0];
        const double u1 = c1 ? t1[1] : t1[0];
        const double u2 = c2 ? q[1] : q[0];
        dst[s] = __dadd_rn(__dadd_rn(u0, u1), __dmul_rn(u2, u0));
      }
      ++k;
      const int c = __ffsll((long long)k) - 1;
      if (c < SA_SOBOL_BITS) {
#pragma unroll
        for (int d = 0; d < 6; ++d) x[d] ^= __ldg(sv + (size_t)d * SA_SOBOL_BITS + c);
      }
    }
  }
}

// Plain Sobol' points [n, dims] (no cross-matrix, no scaling): sample/sobol_sequence.py:49-91.
// Block <-> run of 64 consecutive points, thread <-> dimension: closed form once, then one XOR per point;
// a warp writes 32 consecutive dimensions of a row.
constexpr int kPointRows = 64;
__global__ void __launch_bounds__(256)
sobol_points_kernel(const uint32_t *__restrict__ sv, const uint32_t *__restrict__ shift, uint64_t first_index,
                    uint64_t n, int dims, double *__restrict__ out) {
  const uint64_t r0 = (uint64_t)blockIdx.x * kPointRows;
  const uint64_t r1 = (r0 + kPointRows < n) ? r0 + kPointRows : n;
  for (int d = threadIdx.x; d < dims; d += blockDim.x) {
    const uint32_t *v = sv + (size_t)d * SA_SOBOL_BITS;
    uint64_t k = first_index + r0;
    uint32_t x = sobol_point(sv, shift ? shift[d] : 0u, d, k);
    for (uint64_t r = r0; r < r1; ++r) {
      out[r * (uint64_t)dims + d] = (double)x * kTwoPowMinus30;
      ++k;
      const int c = __ffsll((long long)k) - 1;
      if (c < SA_SOBOL_BITS) x ^= __ldg(v + c);
    }
  }
}

// Stand-alone scaling of a [rows, D] matrix of unit-cube values (util/__init__.py:56-90): the same
// per-column transforms as the fused sample kernels, in place.
__global__ void __launch_bounds__(256)
scale_rows_kernel(double *__restrict__ X, uint64_t total, int D, const ColScale cs) {
  const uint64_t S = (uint64_t)gridDim.x * blockDim.x;
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int col = (int)(idx % (uint64_t)D);
  const int dcol = (int)(S % (uint64_t)D);
  for (; idx < total; idx += S) {
    const double u = X[idx];
    X[idx] = cs.code ? dist_transform(u, cs.code[col], cs.par + 4 * (size_t)col)
                     : __dadd_rn(__dmul_rn(u, cs.span[col]), cs.lb[col]);
    col += dcol;
    if (col >= D) col -= D;
  }
}

static bool reg_path_ok(int D) {
  if (D < 2 || (D & 1)) return false;
  const int h = D / 2;
  return (h >= 32) ? (h % 32 == 0) : (32 % h == 0);
}

static int check_sample_args(const char *fn, const uint32_t *d_sv, uint64_t first_index, uint64_t n,
                             int D, int Dg, const int32_t *gcol, const double *lb, const double *span,
                             const int32_t *code, const double *par, const void *out) {
  SA_REQUIRE(d_sv && out, SA_ERR_ARG, "%s: null pointer", fn);
  SA_REQUIRE((code && par) || (!code && lb && span), SA_ERR_ARG,
             "%s: need either lb/span or dist_code/dist_par", fn);
  SA_REQUIRE(D >= 1 && Dg >= 1 && Dg <= D, SA_ERR_ARG, "%s: bad D=%d Dg=%d", fn, D, Dg);
  SA_REQUIRE(gcol != nullptr || Dg == D, SA_ERR_ARG, "%s: Dg != D needs group_of_col", fn);
  SA_REQUIRE(2 * (int64_t)D <= SA_SOBOL_MAXDIM, SA_ERR_RANGE, "%s: 2*D=%d exceeds %d Sobol' dimensions",
             fn, 2 * D, SA_SOBOL_MAXDIM);
  SA_REQUIRE(first_index + n <= (1ull << SA_SOBOL_BITS) && first_index + n >= first_index, SA_ERR_RANGE,
             "%s: index range [%llu, %llu) exceeds 2^30 points", fn, (unsigned long long)first_index,
             (unsigned long long)(first_index + n));
  return SA_OK;
}

}  // namespace sa

using namespace sa;

extern "C" int sa_saltelli_sample_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                      uint64_t first_index, uint64_t n, int D, int Dg,
                                      const int32_t *d_group_of_col, int second_order,
                                      const double *d_lb, const double *d_span, const int32_t *d_dist_code,
                                      const double *d_dist_par, double *d_out) {
  int rc = check_sample_args("sa_saltelli_sample_f64", d_sv, first_index, n, D, Dg, d_group_of_col, d_lb,
                             d_span, d_dist_code, d_dist_par, d_out);
  const ColScale cs{d_lb, d_span, d_dist_code, d_dist_par};
  if (rc) return rc;
  if (n == 0) return SA_OK;
  SA_REQUIRE(((uintptr_t)d_out & 15) == 0, SA_ERR_ARG, "sa_saltelli_sample_f64: d_out not 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  const bool grouped = d_group_of_col != nullptr;
  const int sms = sm_count();
  if (reg_path_ok(D) && D >= 16) {  // D <= 8: the flat generic path is faster (4.5 vs 3.6 TB/s at D = 8)
    // enough warps to fill the machine, long enough row runs to amortise the closed form
    uint64_t rpw = n / ((uint64_t)sms * 8 * 8);
    if (rpw < 1) rpw = 1;
    if (rpw > 16) rpw = 16;
    // few rows of a wide design: deal the 64-column blocks of a row over neighbouring warps (the smallest divisor
    // of ncb that brings the grid to ~64 warps per SM, else all of them)
    const int ncb = D >= 64 ? D / 64 : 1;
    int cb_split = 1;
    for (int d = 1; d <= ncb; ++d) {
      if (ncb % d) continue;
      cb_split = d;
      if (ceil_div_u64(n, rpw) * (uint64_t)d >= (uint64_t)sms * 64) break;
    }
    const uint64_t warps = ceil_div_u64(n, rpw) * (uint64_t)cb_split;
    const uint64_t blocks = ceil_div_u64(warps, 8);
    SA_REQUIRE(blocks < (1ull << 31), SA_ERR_UNSUPPORTED, "sa_saltelli_sample_f64: grid too large");
    // D < 64: more of the time goes into the per-row Sobol' update, two resident blocks per SM (<= 128
    // registers) hide it better (+16 % at D = 16, +24 % at D = 8); D >= 64 is faster with one.
    auto pick = [&](auto minb) {
      constexpr int M = decltype(minb)::value;
      return grouped ? (d_dist_code ? saltelli_reg_kernel<true, true, M> : saltelli_reg_kernel<true, false, M>)
                     : (d_dist_code ? saltelli_reg_kernel<false, true, M> : saltelli_reg_kernel<false, false, M>);
    };
    auto kern = D < 64 ? pick(std::integral_constant<int, 2>{}) : pick(std::integral_constant<int, 1>{});
    kern<<<(unsigned)blocks, 256, 0, st>>>(d_sv, d_shift, first_index, n, D, Dg, d_group_of_col, step,
                                          second_order, cs, d_out, (int)rpw, cb_split);
  } else {
    const size_t row_bytes = 2 * (size_t)D * sizeof(double);
    SA_REQUIRE(row_bytes <= 200 * 1024, SA_ERR_UNSUPPORTED,
               "sa_saltelli_sample_f64: D=%d needs %zu B of shared memory per base row", D, row_bytes);
    const uint64_t rowlen = (uint64_t)step * D;
    // base rows per chunk: 32 for wide rows; for short rows (small D) enough of them that a chunk writes
    // ~24 K values -- with 32 the per-chunk closed-form start and the two barriers capped every D <= 5 at
    // 84 Grows/s (20-50 % of the HBM roofline)
    int BR = (int)((48 * 1024) / row_bytes);
    int cap = (int)(24576 / rowlen);
    if (cap < 32) cap = 32;
    if (cap > 1024) cap = 1024;
    if (BR < 1) BR = 1;
    if (BR > cap) BR = cap;
    while (BR > 1 && (uint64_t)BR * rowlen >= (1ull << 31)) BR >>= 1;
    SA_REQUIRE((uint64_t)BR * rowlen < (1ull << 31), SA_ERR_UNSUPPORTED,
               "sa_saltelli_sample_f64: row of %llu values too long", (unsigned long long)rowlen);
    const size_t smem = (size_t)BR * row_bytes;
    const uint32_t m_rowlen = fastdiv_magic((uint64_t)BR * rowlen, (uint32_t)rowlen);
    const uint32_t m_D = fastdiv_magic(rowlen, (uint32_t)D);
    uint64_t blocks = ceil_div_u64(n, BR);
    if (blocks > (uint64_t)sms * 16) blocks = (uint64_t)sms * 16;
    auto kern = grouped ? saltelli_generic_kernel<true> : saltelli_generic_kernel<false>;
    if (smem > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, 256, smem, st>>>(d_sv, d_shift, first_index, n, D, Dg, d_group_of_col, step,
                                             second_order, cs, d_out, BR, m_rowlen, m_D);
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_scale_samples_f64(void *stream, double *d_X, uint64_t rows, int D, const double *d_lb,
                                    const double *d_span, const int32_t *d_dist_code, const double *d_dist_par) {
  SA_REQUIRE(d_X && D >= 1, SA_ERR_ARG, "sa_scale_samples_f64: bad arguments");
  SA_REQUIRE((d_dist_code && d_dist_par) || (!d_dist_code && d_lb && d_span), SA_ERR_ARG,
             "sa_scale_samples_f64: need either lb/span or dist_code/dist_par");
  if (rows == 0) return SA_OK;
  const uint64_t total = rows * (uint64_t)D;
  uint64_t blocks = ceil_div_u64(total, 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  const ColScale cs{d_lb, d_span, d_dist_code, d_dist_par};
  scale_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, total, D, cs);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_sobol_points_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                   uint64_t first_index, uint64_t n, int dims, double *d_out) {
  SA_REQUIRE(d_sv && d_out, SA_ERR_ARG, "sa_sobol_points_f64: null pointer");
  SA_REQUIRE(dims >= 1 && dims <= SA_SOBOL_MAXDIM, SA_ERR_RANGE, "sa_sobol_points_f64: dims=%d out of range", dims);
  SA_REQUIRE(first_index + n <= (1ull << SA_SOBOL_BITS), SA_ERR_RANGE,
             "sa_sobol_points_f64: index range exceeds 2^30 points");
  if (n == 0) return SA_OK;
  const uint64_t blocks = ceil_div_u64(n, kPointRows);
  SA_REQUIRE(blocks < (1ull << 31), SA_ERR_UNSUPPORTED, "sa_sobol_points_f64: grid too large");
  sobol_points_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_sv, d_shift, first_index, n, dims, d_out);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_eval_ishigami_f64(void *stream, const double *d_X, uint64_t rows, double A, double B,
                                    double *d_Y) {
  SA_REQUIRE(d_X && d_Y, SA_ERR_ARG, "sa_eval_ishigami_f64: null pointer");
  if (rows == 0) return SA_OK;
  uint64_t blocks = ceil_div_u64(rows, 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  eval_ishigami_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_X, rows, A, B, d_Y);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_eval_sobol_g_f64(void *stream, const double *d_X, uint64_t rows, int D, const double *d_a,
                                   const double *d_delta, const double *d_alpha, double *d_Y,
                                   int32_t *d_flag) {
  SA_REQUIRE(d_X && d_Y && d_a, SA_ERR_ARG, "sa_eval_sobol_g_f64: null pointer");
  SA_REQUIRE(D >= 1, SA_ERR_ARG, "sa_eval_sobol_g_f64: bad D=%d", D);
  if (rows == 0) return SA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int lpr = 1;
  while (lpr < 32 && lpr < D) lpr <<= 1;
  const uint64_t rows_per_block = 256 / lpr;
  uint64_t blocks = ceil_div_u64(rows, rows_per_block);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
#define SA_G_CASE(L)                                                                                   \
  case L:                                                                                              \
    eval_sobol_g_kernel<L><<<(unsigned)blocks, 256, 0, st>>>(d_X, rows, D, d_a, d_delta, d_alpha, d_Y, \
                                                            d_flag);                                   \
    break;
  switch (lpr) {
    SA_G_CASE(1) SA_G_CASE(2) SA_G_CASE(4) SA_G_CASE(8) SA_G_CASE(16) SA_G_CASE(32)
  }
#undef SA_G_CASE
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_saltelli_eval_sobol_g_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                            uint64_t first_index, uint64_t n, int D, int Dg,
                                            const int32_t *d_group_of_col, int second_order,
                                            const double *d_lb, const double *d_span,
                                            const int32_t *d_dist_code, const double *d_dist_par,
                                            const double *d_a, const double *d_delta, const double *d_alpha,
                                            double *d_Y, int32_t *d_flag) {
  int rc = check_sample_args("sa_saltelli_eval_sobol_g_f64", d_sv, first_index, n, D, Dg, d_group_of_col,
                             d_lb, d_span, d_dist_code, d_dist_par, d_Y);
  const ColScale cs{d_lb, d_span, d_dist_code, d_dist_par};
  if (rc) return rc;
  SA_REQUIRE(d_a != nullptr, SA_ERR_ARG, "sa_saltelli_eval_sobol_g_f64: null a");
  if (n == 0) return SA_OK;
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  const size_t row_bytes = 2 * (size_t)D * sizeof(double);
  SA_REQUIRE(row_bytes + D * sizeof(int) <= 200 * 1024, SA_ERR_UNSUPPORTED,
             "sa_saltelli_eval_sobol_g_f64: D=%d too large", D);
  if (!d_group_of_col && D >= 32 && 4 * (size_t)(D | 1) * sizeof(double) * kGRows <= 96 * 1024) {
    // prefix-sharing kernel: A/B factors and their prefix products of BR base rows in shared memory
    const size_t pitch_bytes = 4 * (size_t)(D | 1) * sizeof(double);  // fa, fb, pa, pb rows
    int BR = (int)((64 * 1024) / pitch_bytes);
    BR -= BR % kGRows;
    if (BR < kGRows) BR = kGRows;
    if (BR > 64) BR = 64;
    const size_t smem = (size_t)BR * pitch_bytes;
    uint64_t blocks = ceil_div_u64(n, BR);
    if (blocks > (uint64_t)sm_count() * 16) blocks = (uint64_t)sm_count() * 16;
    if (smem > 48 * 1024)  // per device: no process-wide cache of the opt-in
      SA_CUDA(cudaFuncSetAttribute(saltelli_eval_g_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    saltelli_eval_g_pairs_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
        d_sv, d_shift, first_index, n, D, step, second_order, cs, d_a, d_delta, d_alpha, d_Y, d_flag, BR);
  } else {
    int BR = (int)((32 * 1024) / row_bytes);
    if (BR < 1) BR = 1;
    if (BR > 64) BR = 64;
    const size_t smem = (size_t)BR * row_bytes + (size_t)D * sizeof(int);
    uint64_t blocks = ceil_div_u64(n, BR);
    if (blocks > (uint64_t)sm_count() * 16) blocks = (uint64_t)sm_count() * 16;
    auto kern = d_group_of_col ? saltelli_eval_g_kernel<true> : saltelli_eval_g_kernel<false>;
    if (smem > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(d_sv, d_shift, first_index, n, D, Dg,
                                                               d_group_of_col, step, second_order, cs, d_a,
                                                               d_delta, d_alpha, d_Y, d_flag, BR);
  }
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_saltelli_eval_ishigami_f64(void *stream, const uint32_t *d_sv, const uint32_t *d_shift,
                                             uint64_t first_index, uint64_t n, int Dg,
                                             const int32_t *d_group_of_col, int second_order,
                                             const double *d_lb, const double *d_span,
                                             const int32_t *d_dist_code, const double *d_dist_par, double A,
                                             double B, double *d_Y) {
  int rc = check_sample_args("sa_saltelli_eval_ishigami_f64", d_sv, first_index, n, 3, Dg, d_group_of_col,
                             d_lb, d_span, d_dist_code, d_dist_par, d_Y);
  const ColScale cs{d_lb, d_span, d_dist_code, d_dist_par};
  if (rc) return rc;
  if (n == 0) return SA_OK;
  const int step = second_order ? 2 * Dg + 2 : Dg + 2;
  uint64_t blocks = ceil_div_u64(ceil_div_u64(n, kIshRun), 256);
  if (blocks > (uint64_t)sm_count() * 32) blocks = (uint64_t)sm_count() * 32;
  if (d_group_of_col)
    saltelli_eval_ishigami_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_sv, d_shift, first_index, n, Dg, d_group_of_col, step, second_order, cs, A, B, d_Y);
  else
    saltelli_eval_ishigami_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_sv, d_shift, first_index, n, Dg, nullptr, step, second_order, cs, A, B, d_Y);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
