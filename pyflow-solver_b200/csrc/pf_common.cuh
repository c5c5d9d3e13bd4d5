// pf_common.cuh -- shared device/host definitions for libpyflow_b200.
//
// Parity rules (SURVEY.md section 7 "hard parts" 2): every element-wise
// expression is evaluated in the reference's (numpy's) order, in IEEE double,
// with FMA contraction OFF (the TU is compiled with -fmad=false) so results
// are bit-identical to the numpy float64 path.  Only reductions differ in
// summation order (as any BLAS does).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/pyflow_b200.h"

namespace pf {

constexpr int HALO = 2;            // halo rows below/above the local slab (QUICK needs j-2)
constexpr int PITCH_ALIGN = 16;    // row pitch in doubles is a multiple of 16 (128 B)

// Device layout of one field: (ny_local + 2*HALO) rows x pitch doubles.
// Element (local row jl in [-HALO, nyl+HALO), column i in [0, nx)):
//     base[(jl + HALO) * pitch + i]
// Columns [nx, pitch) are padding: never stored to by kernels, never used.
struct Slab {
  int nx;      // interior columns
  int nyl;     // local interior rows
  int nyg;     // global interior rows
  int row0;    // global row index of local row 0
  int pitch;   // doubles per row
};

__host__ __device__ __forceinline__ size_t cell(const Slab& s, int jl, int i) {
  return (size_t)(jl + HALO) * (size_t)s.pitch + (size_t)i;
}

// Slots of the device scalar block (doubles).
enum Scalar {
  // [0,1] all-reduced with MAX, [2..8] with SUM at the end of a step (multi-GPU)
  S_UMAX = 0, S_VMAX,
  S_DIVC_SUMSQ, S_DIVA_SUMSQ, S_RU_SUMSQ, S_RV_SUMSQ, S_NONFINITE, S_UNAN, S_VNAN,
  S_DIVB_SUMSQ, S_RHS_SUM,                    // rhs pass 1
  S_BB, S_BZ,                                 // rhs pass 2: b.b, b.(M b)
  S_RES_SUMSQ, S_X0,                          // ||b - A x||^2, x[0]
  S_RR, S_INFO, S_ITERS,                      // CG results
  S_PQ,                                       // split-phase CG (multi-GPU) state
  S_RR_NEW, S_RZ_NEW,
  S_RHO, S_RHO_PREV, S_DONE, S_ATOL,
  S_TMP0, S_TMP1,                             // divergence statistics of the alternate driver
  S_COUNT = 32
};

// ---- exact-order arithmetic helpers ---------------------------------------
// a / h, CORRECTLY ROUNDED (bit for bit what the IEEE division gives), for a divisor that does not change inside a
// kernel (dx, dy, 2 dx, dx^2 ...): r = RN(1 / h) is loop-invariant (one real division per thread, hoisted out of the
// sweeps), then q0 = RN(a r), two residual corrections e = a - q h (exact in an FMA), q' = RN(q + e r).  With r the
// correctly rounded reciprocal and q1 a faithful quotient, the second correction returns RN(a / h) (Markstein's
// theorem); the guards keep every intermediate away from underflow / overflow, and everything else -- infinities,
// NaNs, magnitudes beyond 2^+-600, divisors that are not positive and moderate -- takes the real division.  Zeros are
// common (fields at rest) and stay on the fast path: +-0 / h = +-0 for h > 0.  1 DMUL + 4 DFMA instead of the ~14
// dependent instructions + MUFU of a double division: the general-spacing CG sweeps (12 divisions per cell and
// iteration, the reference's arithmetic) went from 0.19 of the copy peak to the number in DESIGN 4.1.
// Checked against `a / h` on 2e10 random and near-midpoint pairs (tests/divcheck.c, tests/test_exact_division.py)
// and by every bit-exactness test with non-unit spacing, on the host emulation and on the GPU.
#ifdef PF_HOST_EMULATION
static inline double div_ieee(double a, double h) { return a / h; }
#else
// the rare path, ONE copy per kernel (a double division inlined at every call site doubled the size of the sweeps)
static __device__ __noinline__ double div_ieee(double a, double h) { return a / h; }
#endif

__device__ __forceinline__ double div_by_invariant(double a, double h) {
  const double r = 1.0 / h;
  const bool h_ok = h >= 0x1p-100 && h <= 0x1p100;
  const double aa = fabs(a);
  if (h_ok && ((aa >= 0x1p-600 && aa <= 0x1p600) || aa == 0.0)) {
    const double q0 = a * r;
    const double e0 = fma(-q0, h, a);
    const double q1 = fma(e0, r, q0);
    const double e1 = fma(-q1, h, a);
    const double q2 = fma(e1, r, q1);
    return aa == 0.0 ? a : q2;
  }
  return div_ieee(a, h);
}

template <bool UNIT>
__device__ __forceinline__ double dv(double a, double h) {
  // x / 1.0 == x exactly, so the unit-spacing path skips the division
  // (dx = dy = 1.0 on the config path: solvers/solver.py:19-22).
  if (UNIT) return a;
#ifdef PF_IEEE_DIVIDE
  return a / h;
#else
  return div_by_invariant(a, h);
#endif
}

// -laplacian_coherent at one cell (core_ops_coherent.py:45-84 composed;
// SURVEY 8a row a9): ge/gw/gn/gs are the forward-difference gradients on the
// four faces (0 on the high Neumann faces and beyond the low ones), then
// lap = (0 + (ge-gw)/dx) + (gn-gs)/dy.
template <bool UNIT>
__device__ __forceinline__ double neg_lap(double c, double w, double e, double s, double n,
                                          bool hw, bool he, bool hs, bool hn, double dx, double dy) {
  const double ge = he ? dv<UNIT>(e - c, dx) : 0.0;
  const double gw = hw ? dv<UNIT>(c - w, dx) : 0.0;
  const double gn = hn ? dv<UNIT>(n - c, dy) : 0.0;
  const double gs = hs ? dv<UNIT>(c - s, dy) : 0.0;
  const double xt = dv<UNIT>(ge - gw, dx);
  const double yt = dv<UNIT>(gn - gs, dy);
  const double lap = (0.0 + xt) + yt;
  return -lap;
}

// One row of the matrix the reference assembles (pressure_solver.py:15-40) applied to a vector: identity on the
// boundary ring, else the five products summed left to right in CSR column order (south, west, centre, east, north),
// which is how scipy's csr_matvec evaluates it.
__device__ __forceinline__ double ring5_row(double c, double w, double e, double s, double n, bool ring, const double* coef) {
  if (ring) return c;
  double sum = 0.0;
  sum += coef[0] * s;
  sum += coef[1] * w;
  sum += coef[2] * c;
  sum += coef[3] * e;
  sum += coef[4] * n;
  return sum;
}

// ---- raw global memory access (never the non-coherent path: buffers are
// rewritten between phases of the persistent kernel) -------------------------
#ifdef PF_HOST_EMULATION
// tests/hostemu compiles these sources with g++ and executes kernel bodies on the host (one OS thread per CUDA
// thread): the PTX accessors below become plain loads and stores there.  Never defined in the product build.
__device__ __forceinline__ double2 ld2(const double* p) { return make_double2(p[0], p[1]); }
__device__ __forceinline__ double ld1(const double* p) { return *(const volatile double*)p; }
__device__ __forceinline__ uint64_t policy_evict_first() { return 0; }
__device__ __forceinline__ uint64_t policy_evict_last() { return 0; }
__device__ __forceinline__ uint64_t policy_evict_normal() { return 0; }
__device__ __forceinline__ double2 ld2_hint(const double* p, uint64_t) { return make_double2(p[0], p[1]); }
__device__ __forceinline__ double2 ld2_hint_nol1(const double* p, uint64_t) { return make_double2(p[0], p[1]); }
__device__ __forceinline__ void st2_hint(double* p, double a, double b, uint64_t) { p[0] = a; p[1] = b; }
__device__ __forceinline__ void st2(double* p, double a, double b) { p[0] = a; p[1] = b; }
__device__ __forceinline__ void st1(double* p, double a) { p[0] = a; }
__device__ __forceinline__ double ld_cg(const double* p) { return *(const volatile double*)p; }
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
#else
__device__ __forceinline__ double2 ld2(const double* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld1(const double* p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// Streaming accesses: an L2 evict-first cache policy (createpolicy) attached to
// the access, for data nobody re-reads before it would be evicted anyway (x in
// phase B, the old search direction in phase A), so that it does not displace
// r and p, which the next (reversed) sweep re-reads from L2.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld2_hint(const double* p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld2_hint_nol1(const double* p, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st2_hint(double* p, double a, double b, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void st2(double* p, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st1(double* p, double a) {
  asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}
__device__ __forceinline__ double ld_cg(const double* p) {  // L2-only (partials, scalars)
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#endif  // PF_HOST_EMULATION

// ---- LL words: self-validating 64-bit flag+data words for the lock-free exchanges ------------------
// Every value that crosses CTAs / GPUs without a fence between "flag" and "payload" travels as LL words
// (the encoding of NCCL's LL protocol): a 64-bit word = {32-bit tag, 32 payload bits}; a double is two
// words.  The PTX memory model guarantees single-copy atomicity for scalar accesses of up to 64 bits and
// models a vector access as independent scalar accesses in unspecified order -- so a reader that finds
// the expected tag in a word holds the payload of THAT store, whatever happens to the neighbouring words;
// a packet is accepted only when every one of its words carries the tag.  No 128/256-bit atomicity is
// assumed anywhere (round 1 relied on STG.256 / LDG.256 being indivisible, which is an implementation
// property).  Tags are derived from the monotonic exchange epoch and are never 0, so zero-initialised
// memory never validates.
constexpr int LL_GPU = 0, LL_SYS = 1;

__host__ __device__ __forceinline__ unsigned ll_tag(unsigned long long e) { return (unsigned)(e % 0xffffffffULL) + 1u; }

struct LLVal {                       // one double as two LL words
  unsigned long long w0, w1;
};
__device__ __forceinline__ LLVal ll_encode(double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  LLVal o;
  o.w0 = ((unsigned long long)tag << 32) | (b & 0xffffffffULL);
  o.w1 = ((unsigned long long)tag << 32) | (b >> 32);
  return o;
}
__device__ __forceinline__ bool ll_ok(const LLVal& k, unsigned tag) {
  return (unsigned)(k.w0 >> 32) == tag && (unsigned)(k.w1 >> 32) == tag;
}
__device__ __forceinline__ double ll_value(const LLVal& k) {
  return __longlong_as_double((long long)((k.w1 << 32) | (k.w0 & 0xffffffffULL)));
}

#ifdef PF_HOST_EMULATION
// Host emulation: every word is its own relaxed atomic access and the words of a packet are written in
// DESCENDING address order with a yield in between -- a reader regularly sees torn packets, which is
// exactly what the encoding has to survive (tests/test_kernels_on_host.py).
template <int SCOPE>
__device__ __forceinline__ void st_ll2(void* p, const LLVal& a) {
  unsigned long long* q = (unsigned long long*)p;
  __atomic_store_n(q + 1, a.w1, __ATOMIC_RELAXED);
  emu_yield();
  __atomic_store_n(q + 0, a.w0, __ATOMIC_RELAXED);
}
template <int SCOPE>
__device__ __forceinline__ void st_ll4(void* p, const LLVal& a, const LLVal& b) {
  st_ll2<SCOPE>((unsigned long long*)p + 2, b);
  emu_yield();
  st_ll2<SCOPE>(p, a);
}
template <int SCOPE>
__device__ __forceinline__ LLVal ld_ll2(const void* p) {
  const unsigned long long* q = (const unsigned long long*)p;
  emu_yield();
  LLVal k;
  k.w0 = __atomic_load_n(q + 0, __ATOMIC_RELAXED);
  k.w1 = __atomic_load_n(q + 1, __ATOMIC_RELAXED);
  return k;
}
template <int SCOPE>
__device__ __forceinline__ void ld_ll4(const void* p, LLVal& a, LLVal& b) {
  a = ld_ll2<SCOPE>(p);
  b = ld_ll2<SCOPE>((const unsigned long long*)p + 2);
}
#else
template <int SCOPE>
__device__ __forceinline__ void st_ll2(void* p, const LLVal& a) {
  if (SCOPE == LL_SYS) asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a.w0), "l"(a.w1) : "memory");
  else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a.w0), "l"(a.w1) : "memory");
}
template <int SCOPE>
__device__ __forceinline__ void st_ll4(void* p, const LLVal& a, const LLVal& b) {
  if (SCOPE == LL_SYS)
    asm volatile("st.relaxed.sys.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a.w0), "l"(a.w1), "l"(b.w0), "l"(b.w1) : "memory");
  else
    asm volatile("st.relaxed.gpu.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a.w0), "l"(a.w1), "l"(b.w0), "l"(b.w1) : "memory");
}
template <int SCOPE>
__device__ __forceinline__ LLVal ld_ll2(const void* p) {
  LLVal k;
  if (SCOPE == LL_SYS) asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(k.w0), "=l"(k.w1) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(k.w0), "=l"(k.w1) : "l"(p) : "memory");
  return k;
}
template <int SCOPE>
__device__ __forceinline__ void ld_ll4(const void* p, LLVal& a, LLVal& b) {
  if (SCOPE == LL_SYS)
    asm volatile("ld.relaxed.sys.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.w0), "=l"(a.w1), "=l"(b.w0), "=l"(b.w1) : "l"(p) : "memory");
  else
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a.w0), "=l"(a.w1), "=l"(b.w0), "=l"(b.w1) : "l"(p) : "memory");
}
#endif  // PF_HOST_EMULATION

// Acquire / release fences.  __threadfence() and __threadfence_system() compile to SEQUENTIALLY CONSISTENT fences
// (MEMBAR.SC.GPU / MEMBAR.SC.SYS); every pattern in this library is a release (fence, then relaxed store / atomic)
// or an acquire (relaxed load / atomic, then fence), for which fence.acq_rel is what the PTX model asks for.
#ifdef PF_HOST_EMULATION
__device__ __forceinline__ void fence_acq_rel_gpu() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
__device__ __forceinline__ void fence_acq_rel_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#else
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
#endif

// A packet of three doubles = 6 LL words in a 64-byte slot (reduction partials across GPUs / CTAs).
constexpr int LL_SLOT_BYTES = 64;
template <int SCOPE>
__device__ __forceinline__ void st_ll_triple(void* slot, unsigned long long e, double v0, double v1, double v2) {
  const unsigned t = ll_tag(e);
  st_ll4<SCOPE>(slot, ll_encode(v0, t), ll_encode(v1, t));
  st_ll2<SCOPE>((char*)slot + 32, ll_encode(v2, t));
}
// One attempt: true (and the three values) iff every word carries the tag of epoch e.
template <int SCOPE>
__device__ __forceinline__ bool ld_ll_triple(const void* slot, unsigned long long e, double (&v)[3]) {
  const unsigned t = ll_tag(e);
  LLVal a, b;
  ld_ll4<SCOPE>(slot, a, b);
  const LLVal c = ld_ll2<SCOPE>((const char*)slot + 32);
  v[0] = ll_value(a); v[1] = ll_value(b); v[2] = ll_value(c);
  return ll_ok(a, t) && ll_ok(b, t) && ll_ok(c, t);
}

// ---- deterministic reductions ----------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the block; every thread returns the same value.  `sm` holds >= 32
// doubles.  Fixed order => bitwise reproducible for a fixed blockDim.
__device__ __forceinline__ double block_sum(double v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();                 // protect sm from a previous use
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += sm[w];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = sm[0];
  for (int w = 1; w < nw; ++w) t = fmax(t, sm[w]);
  return t;
}

}  // namespace pf
