// div_by_invariant() of csrc/pf_common.cuh -- the product source, compiled for the host -- against the IEEE division,
// bit for bit, on random and adversarial operand pairs (tests only; see tests/test_exact_division.py).
//   usage: divcheck PAIRS [SEED]      exit code 1 on the first mismatches
#include "cuda_runtime.h"

#include <stdio.h>
#include <string.h>

#include <thread>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
emu::Team emu::team;
pthread_barrier_t* emu::grid_barrier = nullptr;
thread_local emu::Team* emu::mine = &emu::team;
long long clock64() { return 0; }

#include "pf_common.cuh"

static inline uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double frombits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static inline uint64_t rng(uint64_t* s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return *s; }

static void worker(long n, uint64_t seed, long* bad, long* fast) {
  uint64_t s = seed;
  for (long k = 0; k < n; ++k) {
    // divisor: random mantissa; all ones; a power of two; just above / below one; few significant bits (quotients that
    // fall next to rounding boundaries); exponents inside and outside the fast path's range
    uint64_t mh = rng(&s) & ((1ULL << 52) - 1);
    const int kind = (int)(rng(&s) % 16);
    if (kind == 0) mh = (1ULL << 52) - 1;
    else if (kind == 1) mh = 0;
    else if (kind == 2) mh = rng(&s) & 0xff;
    else if (kind == 3) mh = ((1ULL << 52) - 1) ^ (rng(&s) & 0xff);
    else if (kind == 4 || kind == 5) mh = (rng(&s) & 0xfff) << 40;
    else if (kind == 6) mh = (rng(&s) & 0xffffff) << 28;
    int eh = (int)(rng(&s) % 201) - 100;
    if (rng(&s) % 64 == 0) eh = (int)(rng(&s) % 2000) - 1000;              // beyond the guard: the IEEE path
    double h = frombits(((uint64_t)(1023 + eh) << 52) | mh);
    if (rng(&s) % 256 == 0) h = -h;
    double a;
    const int akind = (int)(rng(&s) % 8);
    if (akind < 3) {                         // random magnitude over the whole double range, random sign
      const uint64_t ma = rng(&s) & ((1ULL << 52) - 1);
      const int ea = (int)(rng(&s) % 2046) - 1022;
      a = frombits(((rng(&s) & 1) << 63) | ((uint64_t)(1023 + ea) << 52) | ma);
    } else if (akind < 6) {                  // the quotient next to a midpoint between two doubles
      const uint64_t mq = rng(&s) & ((1ULL << 52) - 1);
      const int eq = (int)(rng(&s) % 801) - 400;
      const double q = frombits(((uint64_t)(1023 + eq) << 52) | mq);
      const long double mid = (long double)q + (long double)ldexp(1.0, eq - 53);
      a = (double)(mid * (long double)h);
      a = frombits(bits(a) + (uint64_t)(int64_t)((int)(rng(&s) % 5) - 2));
      if (rng(&s) & 1) a = -a;
    } else {                                 // zeros, infinities, NaN, subnormals, the largest finite numbers
      static const double special[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, 4.9e-324, -4.9e-324, 2.2250738585072014e-308,
                                       1.7976931348623157e308, -1.7976931348623157e308, 1.0, -1.0};
      a = special[rng(&s) % 12];
    }
    const double want = a / h, got = pf::div_by_invariant(a, h);
    const bool both_nan = (want != want) && (got != got);
    if (!both_nan && bits(want) != bits(got)) {
      if (++*bad < 5) printf("MISMATCH a=%a h=%a want=%a got=%a\n", a, h, want, got);
    }
    const double aa = fabs(a);
    if (h >= 0x1p-100 && h <= 0x1p100 && ((aa >= 0x1p-600 && aa <= 0x1p600) || aa == 0.0)) ++*fast;
  }
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 10000000L;
  const uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
  const int T = (int)std::thread::hardware_concurrency() > 0 ? (int)std::thread::hardware_concurrency() : 1;
  std::vector<long> bad(T, 0), fast(T, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < T; ++t)
    pool.emplace_back(worker, n / T, 0x9E3779B97F4A7C15ULL * (uint64_t)(seed * 131 + t + 1), &bad[t], &fast[t]);
  for (auto& th : pool) th.join();
  long b = 0, f = 0;
  for (int t = 0; t < T; ++t) { b += bad[t]; f += fast[t]; }
  printf("pairs %ld fast_path %ld mismatches %ld\n", n / T * T, f, b);
  return b != 0;
}
