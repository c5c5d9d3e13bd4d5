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
template <bool UNIT>
__device__ __forceinline__ double dv(double a, double h) {
  // x / 1.0 == x exactly, so the unit-spacing path skips the division
  // (dx = dy = 1.0 on the config path: solvers/solver.py:19-22).
  if (UNIT) return a;
  return a / h;
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
#endif  // PF_HOST_EMULATION

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
