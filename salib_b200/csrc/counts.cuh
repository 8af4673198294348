// Device-side Philox multiplicities (internal header): the resample draw of src/SALib/analyze/sobol.py:144 taken as
// row multiplicities  w_c ~ Multinomial(N; 1/N, ..., 1/N)  without materialising the N indices.
//
// A multinomial splits exactly: the number of the n draws of a row range that land in its left half is
// Binomial(n, 1/2) -- the popcount of n random bits -- and the draws inside a range are iid uniform over it.  A
// thread group owns one resample: it walks L levels of that tree in shared memory (N random bits per level) down
// to windows of W = N / 2^L <= kWinMax rows, then fills one window at a time: n_w uniform draws into a W-byte
// shared-memory histogram (shared atomics), copied out as coalesced words.  No global atomics.
//
// Philox4x32-10 counters are keyed by (seed, resample id, tree node | window, word): the multiplicities of a
// resample do not depend on the batch, the GPU, the thread group or the row window they are generated for.
// Row shards (`row0`, `nrows`) walk only the tree nodes above their own windows.
//
// The body is a device function over a thread group of NT threads synchronised by `Sync` so that both the
// stand-alone kernel (one block per resample) and the spare warps of the persistent GEMM kernel (digits.cu: the
// next batch's multiplicities are drawn while the tensor cores work on the current one) run the same code.
#pragma once
#include "common.cuh"

namespace sa {

constexpr int kWinMax = 16384;  // rows per window: 16 KB of shared memory, fits beside the GEMM's operand ring

// Philox4x32-10 (Salmon et al. 2011), counter-based: 128-bit counter, 64-bit key.
struct Philox4 {
  uint32_t x, y, z, w;
};
__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// The same function with the ten round keys of one (seed) precomputed: 2 of the 6 integer instructions per round.
struct PhiloxKeys {
  uint32_t a[10], b[10];
  __device__ __forceinline__ PhiloxKeys(uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      a[r] = k0 + (uint32_t)r * 0x9E3779B9u;
      b[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
  }
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const PhiloxKeys &key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0;
    const uint32_t h1 = __umulhi(M1, c2), l1 = M1 * c2;
    c0 = h1 ^ c1 ^ key.a[round];
    c1 = l1;
    c2 = h0 ^ c3 ^ key.b[round];
    c3 = l0;
  }
  return Philox4{c0, c1, c2, c3};
}

// levels of the split tree for N rows; -1 if the window scheme does not apply (2^L does not divide N)
__host__ __device__ inline int window_levels(uint64_t N) {
  int L = 0;
  while ((N >> L) > (uint64_t)kWinMax) ++L;
  if (L > 16 || (N & ((1ull << L) - 1)) != 0) return -1;
  return L;
}
// shared memory of one thread group: window histogram + scratch word + tree heap
__host__ __device__ inline size_t window_smem_bytes(uint64_t N, int L) {
  const uint32_t W = (uint32_t)(N >> L);
  return ((size_t)((W + 3) >> 2) + 1 + ((size_t)2 << L)) * sizeof(uint32_t);
}

// Thread groups: a whole block (barrier 0) or NT threads of a block on a named barrier.
template <int NT, int BAR>
struct GroupSync {
  static __device__ __forceinline__ void sync() {
    if (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory");
  }
};

// popcount of the first min(128, n - 128*word) bits of the 128-bit Philox block `word` of tree node (level, j)
__device__ __forceinline__ uint32_t tree_bits(const PhiloxKeys &key, uint64_t c, int level, uint32_t j, uint32_t word,
                                              uint32_t n) {
  const Philox4 p = philox4x32_10(word, 0x80000000u | ((uint32_t)level << 20) | j, (uint32_t)c, (uint32_t)(c >> 32),
                                  key);
  const uint32_t left = n - (word << 7);
  if (left >= 128) return __popc(p.x) + __popc(p.y) + __popc(p.z) + __popc(p.w);
  uint32_t v[4] = {p.x, p.y, p.z, p.w};
  uint32_t s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t lo = 32u * q;
    if (left >= lo + 32) s += __popc(v[q]);
    else if (left > lo) s += __popc(v[q] & ((1u << (left - lo)) - 1u));
  }
  return s;
}

// One resample: multiplicities of rows [row0, row0 + nrows) -> wrow[0 .. nrows) (the caller clears the pad bytes
// beyond nrows).  wsm: window_smem_bytes(N, L) bytes of shared memory owned by the group; tid in [0, NT).
template <int NT, int BAR>
__device__ __forceinline__ void counts_window_resample(uint32_t *wsm, uint64_t seed, uint64_t c, uint64_t N, int L,
                                                       uint64_t row0, uint64_t nrows, uint8_t *__restrict__ wrow,
                                                       int tid) {
  using G = GroupSync<NT, BAR>;
  constexpr int kWarps = NT / 32;
  const uint32_t W = (uint32_t)(N >> L);
  uint32_t *hist = wsm;                       // W bytes (rounded up to words)
  uint32_t *scratch = wsm + ((W + 3) >> 2);
  uint32_t *heap = wsm + ((W + 3) >> 2) + 1;  // heap[k]: draws in node k (root 1, children 2k, 2k+1)
  const PhiloxKeys key((uint32_t)seed, (uint32_t)(seed >> 32));
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t nwin = 1u << L;
  const uint32_t win_lo = (uint32_t)(row0 / W);
  uint32_t win_hi = (uint32_t)((row0 + nrows - 1) / W);
  if (win_hi >= nwin) win_hi = nwin - 1;

  if (tid == 0) heap[1] = (uint32_t)N;
  G::sync();
  for (int l = 0; l < L; ++l) {
    // only the nodes above this call's windows: node j of level l covers windows [j << (L-l), (j+1) << (L-l))
    const uint32_t nodes = 1u << l;
    const uint32_t jlo = win_lo >> (L - l), jhi = win_hi >> (L - l);
    if (jhi - jlo + 1 < (uint32_t)kWarps) {
      for (uint32_t j = jlo; j <= jhi; ++j) {
        const uint32_t n = heap[nodes + j];
        const uint32_t words = (n + 127) >> 7;
        uint32_t sum = 0;
        for (uint32_t wd = tid; wd < words; wd += NT) sum += tree_bits(key, c, l, j, wd, n);
        sum = __reduce_add_sync(kFull, sum);
        G::sync();
        if (tid == 0) *scratch = 0;
        G::sync();
        if (lane == 0 && sum) atomicAdd(scratch, sum);
        G::sync();
        if (tid == 0) {
          const uint32_t left = *scratch;
          heap[2 * (nodes + j)] = left;
          heap[2 * (nodes + j) + 1] = n - left;
        }
      }
      G::sync();
    } else {
      for (uint32_t j = jlo + warp; j <= jhi; j += kWarps) {
        const uint32_t n = heap[nodes + j];
        const uint32_t words = (n + 127) >> 7;
        uint32_t sum = 0;
        for (uint32_t wd = lane; wd < words; wd += 32) sum += tree_bits(key, c, l, j, wd, n);
        sum = __reduce_add_sync(kFull, sum);
        if (lane == 0) {
          heap[2 * (nodes + j)] = sum;
          heap[2 * (nodes + j) + 1] = n - sum;
        }
      }
      G::sync();
    }
  }

  const bool pow2 = (W & (W - 1)) == 0 && W >= 4;
  // W = 2^b (5 <= b <= 14): a 16-bit field f of the Philox output is a row of the window -- bits [2, b) of f, taken
  // where they are, are the BYTE offset of the row's histogram word, bits [b, b+2) its byte lane; both are uniform
  // and independent, so rows stay uniform.  A draw is then 6-7 instructions (mask, add the window base, shift/mask
  // the lane to 8 * lane, 1 << that, red.shared) -- the first version spent 13 per draw on a predicate, a branch,
  // the shared-window base and a shifted word index, 20 per draw with Philox against 12 now (SASS, cuobjdump).
  const int b = 31 - __clz(W);
  const uint32_t off_mask = (W - 1) & ~3u;
  const int lane_lo = b - 3, lane_hi = 16 + b - 3;  // (f >> (b - 3)) & 24 == 8 * ((f >> b) & 3)
  const uint32_t hist_sh = (uint32_t)__cvta_generic_to_shared(hist);
  auto hit = [&](uint32_t off, uint32_t shift8) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_sh + off), "r"(1u << shift8) : "memory");
  };
  for (uint32_t v = win_lo; v <= win_hi; ++v) {
    const uint32_t n = heap[nwin + v];
    for (uint32_t k = tid; k < ((W + 3) >> 2); k += NT) hist[k] = 0;
    G::sync();
    if (pow2 && b >= 5 && b <= 14) {
      // Philox call q serves draws [8q, 8q + 8): whole calls without a predicate, the last partial one by its owner
      const uint32_t whole = n >> 3;
      for (uint32_t q = tid; q < whole; q += NT) {
        const Philox4 p = philox4x32_10(q, 0x40000000u | v, (uint32_t)c, (uint32_t)(c >> 32), key);
        const uint32_t f[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          hit(f[d] & off_mask, (f[d] >> lane_lo) & 24u);
          hit((f[d] >> 16) & off_mask, (f[d] >> lane_hi) & 24u);
        }
      }
      const uint32_t left = n & 7u;
      if (left && (uint32_t)tid == whole % NT) {
        const Philox4 p = philox4x32_10(whole, 0x40000000u | v, (uint32_t)c, (uint32_t)(c >> 32), key);
        const uint32_t f[4] = {p.x, p.y, p.z, p.w};
        for (uint32_t d = 0; d < left; ++d) {
          const uint32_t fv = (d & 1) ? (f[d >> 1] >> 16) : (f[d >> 1] & 0xffffu);
          hit(fv & off_mask, (fv >> lane_lo) & 24u);
        }
      }
    } else {
      const uint32_t calls = (n + 1) >> 1;
      for (uint32_t q = tid; q < calls; q += NT) {
        const Philox4 p = philox4x32_10(q, 0x40000000u | v, (uint32_t)c, (uint32_t)(c >> 32), key);
        const uint32_t i0 = (uint32_t)__umul64hi(((uint64_t)p.y << 32) | p.x, (uint64_t)W);
        atomicAdd(hist + (i0 >> 2), 1u << (8u * (i0 & 3)));
        if (2 * q + 1 < n) {
          const uint32_t i1 = (uint32_t)__umul64hi(((uint64_t)p.w << 32) | p.z, (uint64_t)W);
          atomicAdd(hist + (i1 >> 2), 1u << (8u * (i1 & 3)));
        }
      }
    }
    G::sync();
    // rows [lo, lo+W) of the window, clipped to the caller's row range, -> wrow[row - row0]
    const uint64_t lo = (uint64_t)v * W;
    const uint32_t r0 = (uint32_t)(row0 > lo ? row0 - lo : 0);
    const uint32_t r1 = (uint32_t)((row0 + nrows < lo + W) ? row0 + nrows - lo : W);
    const int64_t dst0 = (int64_t)lo - (int64_t)row0;  // destination of window row 0
    if (((dst0 | (int64_t)r0 | (int64_t)(uintptr_t)wrow) & 3) == 0) {
      uint32_t *dw = reinterpret_cast<uint32_t *>(wrow + dst0);
      const uint32_t full = r1 >> 2;
      for (uint32_t k = (r0 >> 2) + tid; k < full; k += NT) dw[k] = hist[k];
      const uint8_t *hb = reinterpret_cast<const uint8_t *>(hist);
      for (uint32_t r = (full << 2) + tid; r < r1; r += NT) wrow[dst0 + r] = hb[r];
    } else {
      const uint8_t *hb = reinterpret_cast<const uint8_t *>(hist);
      for (uint32_t r = r0 + tid; r < r1; r += NT) wrow[dst0 + (int64_t)r] = hb[r];
    }
    G::sync();
  }
}

}  // namespace sa
