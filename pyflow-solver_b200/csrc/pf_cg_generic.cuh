// pf_cg_generic.cuh -- straightforward persistent CG (q = A p stored, one thread
// per cell, grid-stride) for the operators that are NOT the hot path:
//   MODE_PINNED    x -> -(laplacian_coherent(x)), row 0 = identity
//                  (pressure_solver.py:97-105; used as an independent cross-check
//                  of the tuned kernels)
//   MODE_DIRICHLET x -> +laplacian_coherent(x), identity rows on the boundary ring
//                  (manufactured / verification mode, pressure_solver.py:61-93)
//   MODE_RING5     the matrix the reference ASSEMBLES (assemble_negative_laplacian,
//                  pressure_solver.py:15-40): identity rows on the boundary ring, the plain
//                  5-point -Laplacian inside, evaluated in CSR column order like scipy's
//                  csr_matvec -- what linear_solvers.solve(A, b) gets when a caller hands it
//                  that sparse matrix instead of the matrix-free operator
// Same scipy recurrence and stopping rule as pf_cg.cuh; no preconditioner other
// than the ring Jacobi of the pinned operator (the reference skips the automatic
// Jacobi in manufactured mode).  Single-GPU contexts only.
#pragma once

#include <cooperative_groups.h>

#include "pf_cg.cuh"

namespace pf {

enum { MODE_PINNED = 0, MODE_DIRICHLET = 1, MODE_RING5 = 2 };

struct GenArgs {
  Slab s;
  const double* b;
  double* x;
  double* r;
  double* p;
  double* q;
  double* part;       // [3][CG_MAX_GRID]
  double* scal;
  double dx, dy, invd, tol;
  int maxiter;
  int mode;
  double coef[5];     // MODE_RING5: interior row {south, west, centre, east, north} = {-1/dy^2, -1/dx^2, 2/dx^2+2/dy^2, -1/dx^2, -1/dy^2}
};

template <bool UNIT>
__device__ __forceinline__ double gen_apply(const GenArgs& a, const double* v, int jl, int i) {
  const Slab& s = a.s;
  const int jg = s.row0 + jl;
  const bool hw = i > 0, he = i < s.nx - 1, hs = jg > 0, hn = jg < s.nyg - 1;
  const double c = v[cell(s, jl, i)];
  const double w = hw ? v[cell(s, jl, i - 1)] : 0.0;
  const double e = he ? v[cell(s, jl, i + 1)] : 0.0;
  const double so = hs ? v[cell(s, jl - 1, i)] : 0.0;
  const double n = hn ? v[cell(s, jl + 1, i)] : 0.0;
  const bool ring = !hw || !he || !hs || !hn;
  if (a.mode == MODE_RING5) return ring5_row(c, w, e, so, n, ring, a.coef);
  const double nl = neg_lap<UNIT>(c, w, e, so, n, hw, he, hs, hn, a.dx, a.dy);
  if (a.mode == MODE_PINNED) return (jg == 0 && i == 0) ? c : nl;
  return ring ? c : -nl;
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS) k_cg_generic(GenArgs a) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  __shared__ double sm[32];
  const Slab& s = a.s;
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const long long ncell = (long long)s.nx * s.nyl;
  const long long stride = (long long)G * CG_THREADS;
  const long long first_cell = (long long)bid * CG_THREADS + tid;

  // x = 0, r = b, ||b||^2, b.(M b)
  double v0 = 0.0, v1 = 0.0;
  for (long long k = first_cell; k < ncell; k += stride) {
    const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
    const int jg = s.row0 + jl;
    const double bv = a.b[cell(s, jl, i)];
    a.x[cell(s, jl, i)] = 0.0;
    a.r[cell(s, jl, i)] = bv;
    const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
    v0 += bv * bv;
    v1 += bv * ((ring ? 1.0 : a.invd) * bv);
  }
  v0 = block_sum(v0, sm);
  v1 = block_sum(v1, sm);
  if (tid == 0) { a.part[bid] = v0; a.part[CG_MAX_GRID + bid] = v1; }
  grid.sync();
  const double bb = fold_partials(a.part, G, sm);
  double rho = fold_partials(a.part + CG_MAX_GRID, G, sm);
  grid.sync();
  const double bnorm = sqrt(bb);
  const double t0 = a.tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;
  double rr = bb, rho_prev = 0.0;
  int iters = 0, info = a.maxiter;
  if (bnorm == 0.0) {
    info = 0;
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) { info = 0; break; }
      const double beta = (it == 0) ? 0.0 : rho / rho_prev;
      for (long long k = first_cell; k < ncell; k += stride) {       // p = z (first) | p = p*beta + z
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const int jg = s.row0 + jl;
        const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
        const double z = (ring ? 1.0 : a.invd) * a.r[cell(s, jl, i)];
        const size_t o = cell(s, jl, i);
        a.p[o] = (it == 0) ? z : (a.p[o] * beta) + z;
      }
      grid.sync();
      double pq = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // q = A p ; p.q
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const double qv = gen_apply<UNIT>(a, a.p, jl, i);
        a.q[cell(s, jl, i)] = qv;
        pq += a.p[cell(s, jl, i)] * qv;
      }
      pq = block_sum(pq, sm);
      if (tid == 0) a.part[bid] = pq;
      grid.sync();
      pq = fold_partials(a.part, G, sm);
      const double alpha = rho / pq;
      double arr = 0.0, arz = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // x += alpha p ; r -= alpha q
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const int jg = s.row0 + jl;
        const size_t o = cell(s, jl, i);
        a.x[o] = a.x[o] + alpha * a.p[o];
        const double rn = a.r[o] - alpha * a.q[o];
        a.r[o] = rn;
        const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
        arr += rn * rn;
        arz += rn * ((ring ? 1.0 : a.invd) * rn);
      }
      arr = block_sum(arr, sm);
      arz = block_sum(arz, sm);
      if (tid == 0) { a.part[CG_MAX_GRID + bid] = arr; a.part[2 * CG_MAX_GRID + bid] = arz; }
      grid.sync();
      rr = fold_partials(a.part + CG_MAX_GRID, G, sm);
      rho_prev = rho;
      rho = fold_partials(a.part + 2 * CG_MAX_GRID, G, sm);
      ++iters;
    }
  }
  // true residual ||b - A x||^2 (interface.py:48-49; == ||A x - b||^2 of pressure_solver.py:174-183)
  grid.sync();
  double res = 0.0;
  for (long long k = first_cell; k < ncell; k += stride) {
    const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
    const double d = a.b[cell(s, jl, i)] - gen_apply<UNIT>(a, a.x, jl, i);
    res += d * d;
  }
  res = block_sum(res, sm);
  if (tid == 0) a.part[bid] = res;
  grid.sync();
  res = fold_partials(a.part, G, sm);
  if (bid == 0 && tid == 0) {
    a.scal[S_INFO] = (double)info;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    a.scal[S_BB] = bb;
    a.scal[S_RES_SUMSQ] = res;
    a.scal[S_X0] = a.x[cell(s, 0, 0)];
  }
}

}  // namespace pf
