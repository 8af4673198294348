// Text output of sample matrices / model outputs: what the reference's CLI does with
// np.savetxt(X, fmt="%.8e") (sample/sobol.py:247-251, scripts via common_args) -- byte-identical on the device.
//
// "%.Pe" of a double is its exact binary value rounded half-to-even to P+1 significant decimal digits
// (CPython's float formatting and glibc printf are both exact).  One thread formats one value:
//   x = m * 2^e exactly (m < 2^53);  digits D = round(x / 10^s), s = k - P, k = floor(log10 x).
//   s <= 0:  x / 10^s = (m * 5^|s|) * 2^(e - s)   -- a big integer shifted right; the dropped bits decide
//   s >  0:  x / 10^s = m * 2^(e - s) / 5^s        -- restoring division (the quotient has < 64 bits)
// Big integers are arrays of 32-bit limbs sized for the whole double range (1152 bits); sample values use 2-4
// limbs.  Pass 1 writes each value into a 32-byte slot and its length; the caller scans the row lengths; pass 2
// packs slots, delimiters and newlines into the final text.
#include "common.cuh"

namespace sa {

constexpr int kBigLimbs = 36;
constexpr int kSlot = 32;

struct Big {
  uint32_t d[kBigLimbs];
  int n;  // limbs in use (d[n-1] != 0 unless n == 0)
};

__device__ __forceinline__ void big_set(Big &a, uint64_t v) {
  a.d[0] = (uint32_t)v;
  a.d[1] = (uint32_t)(v >> 32);
  a.n = a.d[1] ? 2 : (a.d[0] ? 1 : 0);
}

__device__ inline void big_mul_small(Big &a, uint32_t v) {
  uint64_t carry = 0;
  for (int i = 0; i < a.n; ++i) {
    const uint64_t t = (uint64_t)a.d[i] * v + carry;
    a.d[i] = (uint32_t)t;
    carry = t >> 32;
  }
  if (carry && a.n < kBigLimbs) a.d[a.n++] = (uint32_t)carry;
}

__device__ inline void big_mul_pow5(Big &a, int p) {
  while (p >= 13) {
    big_mul_small(a, 1220703125u);  // 5^13
    p -= 13;
  }
  uint32_t v = 1;
  for (int i = 0; i < p; ++i) v *= 5u;
  if (v > 1) big_mul_small(a, v);
}

__device__ inline void big_shl(Big &a, int bits) {
  if (a.n == 0 || bits == 0) return;
  const int limbs = bits >> 5, sh = bits & 31;
  int n = a.n + limbs + 1;
  if (n > kBigLimbs) n = kBigLimbs;
  for (int i = n - 1; i >= 0; --i) {
    const int j = i - limbs;
    uint32_t v = 0;
    if (j >= 0 && j < a.n) v = a.d[j] << sh;
    if (sh && j - 1 >= 0 && j - 1 < a.n) v |= a.d[j - 1] >> (32 - sh);
    a.d[i] = v;
  }
  while (n > 0 && a.d[n - 1] == 0) --n;
  a.n = n;
}

__device__ inline void big_shr1(Big &a) {
  for (int i = 0; i < a.n; ++i) {
    uint32_t v = a.d[i] >> 1;
    if (i + 1 < a.n) v |= a.d[i + 1] << 31;
    a.d[i] = v;
  }
  while (a.n > 0 && a.d[a.n - 1] == 0) --a.n;
}

__device__ inline int big_cmp(const Big &a, const Big &b) {
  if (a.n != b.n) return a.n < b.n ? -1 : 1;
  for (int i = a.n - 1; i >= 0; --i)
    if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
  return 0;
}

__device__ inline void big_sub(Big &a, const Big &b) {  // a -= b, a >= b
  uint64_t borrow = 0;
  for (int i = 0; i < a.n; ++i) {
    const uint64_t bi = (i < b.n ? b.d[i] : 0u) + borrow;
    const uint64_t ai = a.d[i];
    a.d[i] = (uint32_t)(ai - bi);
    borrow = ai < bi ? 1 : 0;
  }
  while (a.n > 0 && a.d[a.n - 1] == 0) --a.n;
}

__device__ inline int big_bits(const Big &a) { return a.n ? 32 * (a.n - 1) + (32 - __clz(a.d[a.n - 1])) : 0; }

constexpr uint64_t kTooBig = ~0ull;

// q = floor(m * 2^e / 10^s); rel = -1 / 0 / +1: the dropped fraction is below / exactly / above one half.
// Returns kTooBig when the quotient needs 63 bits or more (the caller's decimal exponent was too small).
__device__ uint64_t scaled_digits(uint64_t m, int e, int s, int &rel) {
  Big a;
  big_set(a, m);
  const int t = e - s;
  if (s <= 0) {
    big_mul_pow5(a, -s);
    if (t >= 0) {
      big_shl(a, t);
      rel = -1;
      if (big_bits(a) > 62) return kTooBig;
      return (uint64_t)a.d[0] | ((uint64_t)(a.n > 1 ? a.d[1] : 0u) << 32);
    }
    const int sh = -t;
    const int bits = big_bits(a);
    if (bits - sh > 62) return kTooBig;
    if (sh > bits) {  // everything is dropped and the fraction is below one half
      rel = -1;
      return 0;
    }
    // half bit = bit sh-1, sticky = any bit below it
    const int hb = sh - 1;
    const bool half = (a.d[hb >> 5] >> (hb & 31)) & 1u;
    bool sticky = false;
    for (int i = 0; i < (hb >> 5); ++i) sticky |= a.d[i] != 0;
    if (hb & 31) sticky |= (a.d[hb >> 5] & ((1u << (hb & 31)) - 1u)) != 0;
    rel = half ? (sticky ? 1 : 0) : -1;
    uint64_t q = 0;
    for (int b = 0; b < 63; ++b) {
      const int pos = sh + b;
      if (pos >= bits) break;
      q |= (uint64_t)((a.d[pos >> 5] >> (pos & 31)) & 1u) << b;
    }
    return q;
  }
  Big d;
  big_set(d, 1);
  big_mul_pow5(d, s);
  if (t >= 0) big_shl(a, t);
  else big_shl(d, -t);
  const int qbits = big_bits(a) - big_bits(d) + 1;  // the quotient has at most this many bits
  if (qbits > 62) return kTooBig;
  uint64_t q = 0;
  if (qbits > 0) {
    big_shl(d, qbits - 1);
    for (int b = qbits - 1; b >= 0; --b) {
      if (big_cmp(a, d) >= 0) {
        big_sub(a, d);
        q |= 1ull << b;
      }
      if (b) big_shr1(d);
    }
  }
  // remainder a vs d / 2: compare 2a with d
  big_shl(a, 1);
  rel = big_cmp(a, d);
  return q;
}

__device__ __forceinline__ uint64_t pow10_u64(int p) {
  uint64_t v = 1;
  for (int i = 0; i < p; ++i) v *= 10ull;
  return v;
}

// "%.Pe" of x into out[0..kSlot); returns the length.  0 <= P <= 17.
__device__ int format_e(double x, int P, uint8_t *out) {
  const uint64_t bits = (uint64_t)__double_as_longlong(x);
  const bool neg = bits >> 63;
  const int be = (int)((bits >> 52) & 0x7ff);
  const uint64_t frac = bits & 0xfffffffffffffull;
  int len = 0;
  if (be == 0x7ff) {
    if (frac) {
      out[0] = 'n'; out[1] = 'a'; out[2] = 'n';
      return 3;
    }
    if (neg) out[len++] = '-';
    out[len++] = 'i'; out[len++] = 'n'; out[len++] = 'f';
    return len;
  }
  uint64_t D = 0;
  int k = 0;
  if (be != 0 || frac != 0) {
    const uint64_t m = be ? (frac | (1ull << 52)) : frac;
    const int e = be ? be - 1075 : -1074;
    const int top = 63 - __clzll((long long)m);                  // m in [2^top, 2^(top+1))
    k = (int)floor((double)(e + top) * 0.30102999566398120);     // floor(log10 x) or one less
    const uint64_t lo = pow10_u64(P), hi = lo * 10ull;
    for (int iter = 0; iter < 4; ++iter) {
      int rel;
      uint64_t q = scaled_digits(m, e, k - P, rel);
      if (q == kTooBig || q >= hi) {
        ++k;
        continue;
      }
      if (q < lo) {
        --k;
        continue;
      }
      if (rel > 0 || (rel == 0 && (q & 1ull))) ++q;
      if (q == hi) {  // carry into a new digit: 9.99..e k -> 1.00..e k+1
        q = lo;
        ++k;
      }
      D = q;
      break;
    }
  }
  if (neg) out[len++] = '-';
  // digits of D, most significant first (P + 1 of them)
  char dig[20];
  for (int i = P; i >= 0; --i) {
    dig[i] = (char)('0' + (int)(D % 10ull));
    D /= 10ull;
  }
  out[len++] = (uint8_t)dig[0];
  if (P > 0) {
    out[len++] = '.';
    for (int i = 1; i <= P; ++i) out[len++] = (uint8_t)dig[i];
  }
  out[len++] = 'e';
  out[len++] = k < 0 ? '-' : '+';
  const int ak = k < 0 ? -k : k;
  if (ak >= 100) out[len++] = (uint8_t)('0' + ak / 100);
  out[len++] = (uint8_t)('0' + (ak / 10) % 10);
  out[len++] = (uint8_t)('0' + ak % 10);
  return len;
}

// pass 1: slots[i][32], len[i]; rowlen[r] = sum of the row's value lengths + cols (delimiters + newline)
__global__ void __launch_bounds__(128)
format_values_kernel(const double *__restrict__ x, uint64_t n, int P, uint8_t *__restrict__ slots,
                     uint8_t *__restrict__ lens) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    __align__(16) uint8_t buf[kSlot];
    const int len = format_e(x[i], P, buf);
    uint4 *dst = reinterpret_cast<uint4 *>(slots + i * kSlot);
    const uint4 *src = reinterpret_cast<const uint4 *>(buf);
    dst[0] = src[0];
    dst[1] = src[1];
    lens[i] = (uint8_t)len;
  }
}

__global__ void __launch_bounds__(256)
row_lengths_kernel(const uint8_t *__restrict__ lens, uint64_t rows, int cols, int64_t *__restrict__ rowlen) {
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (uint64_t)gridDim.x * blockDim.x) {
    int64_t s = cols;  // cols - 1 delimiters + 1 newline
    for (int c = 0; c < cols; ++c) s += lens[r * (uint64_t)cols + c];
    rowlen[r] = s;
  }
}

// pass 2: one warp per row; rowoff = exclusive scan of rowlen
__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t *__restrict__ slots, const uint8_t *__restrict__ lens, uint64_t rows, int cols,
                 uint8_t delim, const int64_t *__restrict__ rowoff, uint8_t *__restrict__ text) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < rows; r += nwarps) {
    int64_t pos = rowoff[r];
    for (int c0 = 0; c0 < cols; c0 += 32) {
      const int c = c0 + lane;
      const uint64_t i = r * (uint64_t)cols + c;
      const int len = c < cols ? lens[i] + 1 : 0;  // value + delimiter / newline
      int incl = len;                               // inclusive scan of the lengths over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (c < cols) {
        uint8_t *dst = text + pos + incl - len;
        const uint8_t *src = slots + i * kSlot;
        for (int b = 0; b < len - 1; ++b) dst[b] = src[b];
        dst[len - 1] = (c == cols - 1) ? (uint8_t)'\n' : delim;
      }
      pos += __shfl_sync(kFull, incl, 31);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Text input: what np.loadtxt does for the CLI (analyze/sobol.py:476-491, dgsm.py, discrepancy.py cli_action):
// rectangular numeric text -> float64 [rows, cols], every token converted like Python's float() (correctly
// rounded).  Lines are located with a two-level newline count; one thread then walks one line (comments,
// delimiters, tokens), first to count its tokens, then to convert them.
//   value = M * 10^E (M: the significant digits as an integer).  M < 2^53 and |E| <= 22: one IEEE operation
//   on exactly representable operands (Clinger's fast path; every "%.8e" token of ordinary magnitude).
//   Otherwise big integers: E >= 0: (M * 5^E) * 2^E, the top 53 bits with half/sticky from the rest;
//   E < 0: M * 2^sh / 5^|E| by restoring division to 55-56 quotient bits + sticky remainder.  One rounding,
//   at the precision the result's exponent allows (subnormals).
// ------------------------------------------------------------------------------------------------------------
constexpr int kParseChunk = 4096;   // bytes per block of the newline count
constexpr int kMaxDigits = 40;      // significant digits converted exactly; longer tokens are reported

__device__ inline void big_add_small(Big &a, uint32_t v) {
  uint64_t carry = v;
  for (int i = 0; i < a.n && carry; ++i) {
    const uint64_t t = (uint64_t)a.d[i] + carry;
    a.d[i] = (uint32_t)t;
    carry = t >> 32;
  }
  if (carry && a.n < kBigLimbs) a.d[a.n++] = (uint32_t)carry;
}

// round the integer `mant` (nb significant bits; `sticky`: nonzero bits below it) * 2^ex to a double
__device__ double assemble_double(const Big &a, int ex, bool sticky_in, bool neg) {
  // value = a * 2^ex (+ a sticky fraction of the last unit).  Position of the top bit: eb = bits(a) - 1 + ex.
  const int nb = big_bits(a);
  const int eb = nb - 1 + ex;
  if (eb > 1023) return neg ? -INFINITY : INFINITY;
  int prec = 53;
  if (eb < -1022) prec = 53 - (-1022 - eb);  // subnormal: fewer bits
  if (prec < 0) return neg ? -0.0 : 0.0;      // below half of the smallest subnormal
  const int drop = nb - prec;                 // bits of a to drop (prec may be 0: everything, rounding on the top bit)
  uint64_t mant = 0;
  bool half = false, sticky = sticky_in;
  if (drop <= 0) {
    for (int b = 0; b < nb; ++b) mant |= (uint64_t)((a.d[b >> 5] >> (b & 31)) & 1u) << b;
    mant <<= -drop;
  } else {
    const int hb = drop - 1;
    half = (a.d[hb >> 5] >> (hb & 31)) & 1u;
    for (int i = 0; i < (hb >> 5); ++i) sticky |= a.d[i] != 0;
    if (hb & 31) sticky |= (a.d[hb >> 5] & ((1u << (hb & 31)) - 1u)) != 0;
    for (int b = 0; b < prec; ++b) {
      const int pos = drop + b;
      mant |= (uint64_t)((a.d[pos >> 5] >> (pos & 31)) & 1u) << b;
    }
  }
  if (half && (sticky || (mant & 1ull))) ++mant;
  // mant * 2^(eb - prec + 1); mant may have carried to 2^prec (handled by the arithmetic below)
  const int e2 = eb - prec + 1;
  // build with exact scaling: mant < 2^54, result representable (normal, subnormal or overflow to inf)
  double v = (double)mant;
  // split the scaling so that intermediate results stay exact
  int r = e2;
  while (r > 1000) { v *= 8.98846567431158e307; r -= 1023; }   // 2^1023
  while (r < -1000) { v *= 2.2250738585072014e-308; r += 1022; }  // 2^-1022 (exact: v has <= prec bits)
  v = ldexp(v, r);
  return neg ? -v : v;
}

__device__ double decimal_to_double(uint64_t m64, const Big *mbig, int E, bool neg) {
  // fast path
  if (!mbig && m64 < (1ull << 53) && E >= -22 && E <= 22) {
    const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                            1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const double v = E >= 0 ? __dmul_rn((double)m64, p10[E]) : __ddiv_rn((double)m64, p10[-E]);
    return neg ? -v : v;
  }
  Big a;
  if (mbig) a = *mbig;
  else big_set(a, m64);
  if (a.n == 0) return neg ? -0.0 : 0.0;
  // magnitude guards keep the big integers inside their 1152 bits
  const int digits10 = (int)((double)big_bits(a) * 0.30103) + 1;
  if (digits10 + E > 311) return neg ? -INFINITY : INFINITY;
  if (digits10 + E < -330) return neg ? -0.0 : 0.0;
  if (E >= 0) {
    big_mul_pow5(a, E);
    return assemble_double(a, E, false, neg);
  }
  Big d;
  big_set(d, 1);
  big_mul_pow5(d, -E);
  // scale the numerator so that the quotient has 55 or 56 bits
  int sh = big_bits(d) + 55 - big_bits(a);
  if (sh < 0) sh = 0;
  big_shl(a, sh);
  const int qbits = big_bits(a) - big_bits(d) + 1;  // >= 56; more only when the numerator was already long
  Big q;
  q.n = (qbits + 31) >> 5;
  for (int i = 0; i < q.n; ++i) q.d[i] = 0;
  big_shl(d, qbits - 1);
  for (int b = qbits - 1; b >= 0; --b) {
    if (big_cmp(a, d) >= 0) {
      big_sub(a, d);
      q.d[b >> 5] |= 1u << (b & 31);
    }
    if (b) big_shr1(d);
  }
  while (q.n > 0 && q.d[q.n - 1] == 0) --q.n;
  if (q.n == 0) return neg ? -0.0 : 0.0;
  return assemble_double(q, E - sh, a.n != 0, neg);
}

__device__ __forceinline__ uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

// Python float() of the token p[0..len): returns false if it is not a decimal float literal
__device__ bool parse_token(const uint8_t *p, int len, double &out, int &err) {
  int i = 0;
  bool neg = false;
  if (i < len && (p[i] == '+' || p[i] == '-')) neg = p[i++] == '-';
  if (len - i == 3 && lower(p[i]) == 'n' && lower(p[i + 1]) == 'a' && lower(p[i + 2]) == 'n') {
    out = neg ? -NAN : NAN;
    return true;
  }
  if ((len - i == 3 || len - i == 8) && lower(p[i]) == 'i' && lower(p[i + 1]) == 'n' && lower(p[i + 2]) == 'f') {
    bool ok = true;
    if (len - i == 8) {
      const char *rest = "inity";
      for (int j = 0; j < 5; ++j) ok &= lower(p[i + 3 + j]) == (uint8_t)rest[j];
    }
    if (ok) {
      out = neg ? -INFINITY : INFINITY;
      return true;
    }
    return false;
  }
  uint64_t m = 0;
  Big mb;
  bool use_big = false;
  int nd = 0;          // significant digits taken
  int dropped = 0;     // integer-part digits beyond kMaxDigits
  int frac_digits = 0; // digits taken after the decimal point
  bool any = false, seen_dot = false, toolong = false;
  for (; i < len; ++i) {
    const uint8_t c = p[i];
    if (c == '.') {
      if (seen_dot) return false;
      seen_dot = true;
      continue;
    }
    if (c < '0' || c > '9') break;
    any = true;
    const uint32_t dgt = c - '0';
    if (nd == 0 && dgt == 0) {  // leading zero: only moves the decimal point
      if (seen_dot) ++frac_digits;
      continue;
    }
    if (nd >= kMaxDigits) {
      if (dgt) toolong = true;   // a nonzero digit we cannot take exactly
      if (!seen_dot) ++dropped;
      continue;
    }
    if (nd < 19) m = m * 10ull + dgt;
    else {
      if (!use_big) {
        big_set(mb, m);
        use_big = true;
      }
      big_mul_small(mb, 10u);
      big_add_small(mb, dgt);
    }
    ++nd;
    if (seen_dot) ++frac_digits;
  }
  if (!any) return false;
  int ex = 0;
  if (i < len && (p[i] == 'e' || p[i] == 'E')) {
    ++i;
    bool eneg = false;
    if (i < len && (p[i] == '+' || p[i] == '-')) eneg = p[i++] == '-';
    if (i >= len) return false;
    for (; i < len; ++i) {
      if (p[i] < '0' || p[i] > '9') return false;
      if (ex < 100000) ex = ex * 10 + (p[i] - '0');
    }
    if (eneg) ex = -ex;
  }
  if (i != len) return false;
  if (toolong) {
    err |= 4;
    return true;
  }
  out = decimal_to_double(m, use_big ? &mb : nullptr, ex - frac_digits + dropped, neg);
  return true;
}

// ---- line structure ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_newlines_kernel(const uint8_t *__restrict__ text, uint64_t len, int64_t *__restrict__ blockcount) {
  const uint64_t base = (uint64_t)blockIdx.x * kParseChunk;
  int c = 0;
  for (int k = threadIdx.x; k < kParseChunk; k += 256)
    if (base + k < len && text[base + k] == '\n') ++c;
  c = __reduce_add_sync(kFull, c);
  __shared__ int sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    blockcount[blockIdx.x] = t;
  }
}

// line l (l >= 1) starts after the (l-1)-th newline; blockoff = exclusive scan of blockcount
__global__ void __launch_bounds__(256)
line_starts_kernel(const uint8_t *__restrict__ text, uint64_t len, const int64_t *__restrict__ blockoff,
                   int64_t *__restrict__ linestart) {
  const uint64_t base = (uint64_t)blockIdx.x * kParseChunk;
  __shared__ int warpsum[8];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < kParseChunk; k0 += 256) {
    const uint64_t pos = base + k0 + threadIdx.x;
    const bool nl = pos < len && text[pos] == '\n';
    const unsigned ballot = __ballot_sync(kFull, nl);
    if (lane == 0) warpsum[warp] = __popc(ballot);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += warpsum[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (nl) linestart[blockoff[blockIdx.x] + before + 1] = (int64_t)pos + 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += warpsum[w];
      running += t;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) linestart[0] = 0;
}

// Walk one line: tokens separated by `delim` (0: runs of blanks/tabs), '#' starts a comment, '\r' ignored at
// the end.  mode 0: ntok[line] = token count.  mode 1: convert tokens into out[rowidx[line]][0..cols).
__global__ void __launch_bounds__(128)
parse_lines_kernel(const uint8_t *__restrict__ text, uint64_t len, const int64_t *__restrict__ linestart,
                   uint64_t nlines, int delim, int mode, int32_t *__restrict__ ntok,
                   const int64_t *__restrict__ rowidx, int cols, double *__restrict__ out,
                   int32_t *__restrict__ flag) {
  for (uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines;
       l += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t p = (uint64_t)linestart[l];
    uint64_t end = (l + 1 < nlines) ? (uint64_t)linestart[l + 1] - 1 : len;  // excludes the newline
    if (end > len) end = len;
    // cut the comment and trailing blanks / CR
    for (uint64_t q = p; q < end; ++q)
      if (text[q] == '#') {
        end = q;
        break;
      }
    while (end > p && (text[end - 1] == '\r' || text[end - 1] == ' ' || text[end - 1] == '\t')) --end;
    while (p < end && (text[p] == ' ' || text[p] == '\t') && delim != ' ' && delim != '\t') ++p;
    int count = 0, err = 0;
    const int64_t row = mode ? rowidx[l] : 0;
    if (p < end) {
      uint64_t tstart = p;
      for (uint64_t q = p; q <= end; ++q) {
        const bool sep = q == end || (delim ? text[q] == (uint8_t)delim : (text[q] == ' ' || text[q] == '\t'));
        if (!sep) continue;
        if (delim == 0 && q == tstart) {  // run of blanks
          tstart = q + 1;
          continue;
        }
        if (mode) {
          // strip blanks around the field (np.loadtxt does for explicit delimiters)
          uint64_t a = tstart, b = q;
          while (a < b && (text[a] == ' ' || text[a] == '\t')) ++a;
          while (b > a && (text[b - 1] == ' ' || text[b - 1] == '\t')) --b;
          double v = 0.0;
          if (b - a > 400 || !parse_token(text + a, (int)(b - a), v, err)) err |= 1;
          if (count < cols) out[row * (int64_t)cols + count] = v;
        }
        ++count;
        tstart = q + 1;
      }
    }
    if (mode == 0) ntok[l] = count;
    else if (count && count != cols) err |= 2;
    if (err && flag) atomicOr(flag, err);
  }
}

}  // namespace sa

using namespace sa;

extern "C" int sa_text_slot_bytes(void) { return kSlot; }

extern "C" int sa_format_values_f64(void *stream, const double *d_x, uint64_t rows, int cols, int precision,
                                    uint8_t *d_slots, uint8_t *d_lens, int64_t *d_rowlen) {
  SA_REQUIRE(d_x && d_slots && d_lens && d_rowlen, SA_ERR_ARG, "sa_format_values_f64: null pointer");
  SA_REQUIRE(cols >= 1, SA_ERR_ARG, "sa_format_values_f64: bad column count");
  SA_REQUIRE(precision >= 0 && precision <= 17, SA_ERR_ARG, "sa_format_values_f64: precision must be 0..17");
  SA_REQUIRE(((uintptr_t)d_slots & 15) == 0, SA_ERR_ARG, "sa_format_values_f64: d_slots not 16-byte aligned");
  if (rows == 0) return SA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint64_t n = rows * (uint64_t)cols;
  uint64_t blocks = ceil_div_u64(n, 128);
  if (blocks > (uint64_t)sm_count() * 16) blocks = (uint64_t)sm_count() * 16;
  format_values_kernel<<<(unsigned)blocks, 128, 0, st>>>(d_x, n, precision, d_slots, d_lens);
  blocks = ceil_div_u64(rows, 256);
  if (blocks > (uint64_t)sm_count() * 8) blocks = (uint64_t)sm_count() * 8;
  row_lengths_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_lens, rows, cols, d_rowlen);
  note_launches(2);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_pack_text_rows(void *stream, const uint8_t *d_slots, const uint8_t *d_lens, uint64_t rows,
                                 int cols, int delimiter, const int64_t *d_rowoff, uint8_t *d_text) {
  SA_REQUIRE(d_slots && d_lens && d_rowoff && d_text, SA_ERR_ARG, "sa_pack_text_rows: null pointer");
  SA_REQUIRE(cols >= 1 && delimiter > 0 && delimiter < 256, SA_ERR_ARG, "sa_pack_text_rows: bad arguments");
  if (rows == 0) return SA_OK;
  uint64_t blocks = ceil_div_u64(rows, 8);
  if (blocks > (uint64_t)sm_count() * 8) blocks = (uint64_t)sm_count() * 8;
  pack_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_slots, d_lens, rows, cols,
                                                                        (uint8_t)delimiter, d_rowoff, d_text);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_text_parse_chunk_bytes(void) { return kParseChunk; }

extern "C" int sa_text_count_newlines(void *stream, const uint8_t *d_text, uint64_t len, int64_t *d_blockcount) {
  SA_REQUIRE(d_text && d_blockcount, SA_ERR_ARG, "sa_text_count_newlines: null pointer");
  if (len == 0) return SA_OK;
  const uint64_t blocks = ceil_div_u64(len, kParseChunk);
  SA_REQUIRE(blocks < (1ull << 31), SA_ERR_UNSUPPORTED, "sa_text_count_newlines: text too long for one call");
  count_newlines_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_text, len, d_blockcount);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_text_line_starts(void *stream, const uint8_t *d_text, uint64_t len, const int64_t *d_blockoff,
                                   int64_t *d_linestart) {
  SA_REQUIRE(d_text && d_blockoff && d_linestart, SA_ERR_ARG, "sa_text_line_starts: null pointer");
  if (len == 0) return SA_OK;
  const uint64_t blocks = ceil_div_u64(len, kParseChunk);
  line_starts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_text, len, d_blockoff, d_linestart);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}

extern "C" int sa_text_parse_lines(void *stream, const uint8_t *d_text, uint64_t len, const int64_t *d_linestart,
                                   uint64_t nlines, int delimiter, int mode, int32_t *d_ntok,
                                   const int64_t *d_rowidx, int cols, double *d_out, int32_t *d_flag) {
  SA_REQUIRE(d_text && d_linestart, SA_ERR_ARG, "sa_text_parse_lines: null pointer");
  SA_REQUIRE(delimiter >= 0 && delimiter < 256 && delimiter != '\n' && delimiter != '#', SA_ERR_ARG,
             "sa_text_parse_lines: bad delimiter");
  SA_REQUIRE(mode == 0 ? d_ntok != nullptr : (d_rowidx && d_out && cols >= 1), SA_ERR_ARG,
             "sa_text_parse_lines: missing buffers for this mode");
  if (nlines == 0) return SA_OK;
  uint64_t blocks = ceil_div_u64(nlines, 128);
  if (blocks > (uint64_t)sm_count() * 16) blocks = (uint64_t)sm_count() * 16;
  parse_lines_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_text, len, d_linestart, nlines, delimiter,
                                                                         mode, d_ntok, d_rowidx, cols, d_out, d_flag);
  note_launches(1);
  SA_LAUNCH_CHECK();
  return SA_OK;
}
