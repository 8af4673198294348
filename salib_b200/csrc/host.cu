// Host-side pieces of the C ABI: error string, version, SM count, Joe-Kuo direction vectors.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace sa {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace sa

extern "C" int sa_version(void) { return 100; }

extern "C" long long sa_launch_count(int reset) { return sa::launches(reset); }

extern "C" const char *sa_last_error(void) { return sa::g_err; }

// Direction vectors V_1..V_30 per dimension, 30-bit words.
// Algorithm: sample/sobol_sequence.py:62-84 of the reference (Joe & Kuo 2008 recurrence);
// dimension 0 has m_j = 1 for all j.
extern "C" int sa_direction_vectors(const uint32_t *h_a, const uint8_t *h_s, const uint32_t *h_m,
                                    int dims, uint32_t *h_sv) {
  SA_REQUIRE(h_sv != nullptr && dims >= 1, SA_ERR_ARG, "sa_direction_vectors: bad arguments");
  SA_REQUIRE(dims <= SA_SOBOL_MAXDIM, SA_ERR_RANGE, "sa_direction_vectors: dims %d > %d", dims,
             SA_SOBOL_MAXDIM);
  SA_REQUIRE(dims == 1 || (h_a && h_s && h_m), SA_ERR_ARG, "sa_direction_vectors: null table");
  const int B = SA_SOBOL_BITS;
  for (int j = 1; j <= B; ++j) h_sv[j - 1] = 1u << (B - j);
  for (int d = 1; d < dims; ++d) {
    const uint32_t a = h_a[d - 1];
    const int s = h_s[d - 1];
    const uint32_t *m = h_m + (size_t)(d - 1) * 18;
    uint32_t V[SA_SOBOL_BITS + 1];
    SA_REQUIRE(s >= 1 && s <= 18, SA_ERR_ARG, "sa_direction_vectors: degree %d at dim %d", s, d);
    for (int j = 1; j <= B && j <= s; ++j) V[j] = m[j - 1] << (B - j);
    for (int j = s + 1; j <= B; ++j) {
      uint32_t v = V[j - s] ^ (V[j - s] >> s);
      for (int k = 1; k < s; ++k)
        if ((a >> (s - 1 - k)) & 1u) v ^= V[j - k];
      V[j] = v;
    }
    memcpy(h_sv + (size_t)d * B, V + 1, sizeof(uint32_t) * B);
  }
  return SA_OK;
}
