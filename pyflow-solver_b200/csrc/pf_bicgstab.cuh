// pf_bicgstab.cuh -- BiCGSTAB on the pinned Neumann operator, for the `method='bicgstab'` branch of the
// reference's linear-solve seam (linear_solvers/interface.py:44-45; it is the DEFAULT method of `solve`, though
// the time step always asks for 'cg': off the hot path).
//
// The recurrence is scipy.sparse.linalg.bicgstab's (scipy/sparse/linalg/_isolve/iterative.py, pinned 1.13.1,
// same loop in 1.18), with x0 = 0 as interface.py:33 passes it, statement for statement:
//
//   atol = max(0, rtol*||b||);  ||b|| == 0 -> return b, 0;  r = b;  rtilde = r
//   loop:  ||r|| < atol -> 0;  rho = rtilde.r;  |rho| < eps^2 -> -10;
//          it > 0: |omega| < eps^2 -> -11;  beta = (rho/rho_prev)*(alpha/omega);  p = ((p - omega*v)*beta) + r
//          it = 0: p = r
//          phat = M p;  v = A phat;  rv = rtilde.v;  rv == 0 -> -11;  alpha = rho/rv;  r = r - alpha*v  (= s)
//          ||s|| < atol -> x = x + alpha*phat, 0
//          shat = M s;  t = A shat;  omega = (t.s)/(t.t);  x = x + alpha*phat;  x = x + omega*shat;  r = r - omega*t
//   exhausted -> maxiter
//
// M is the ring-identity Jacobi diagonal (1 on the boundary ring, invd inside), the identity (invd = 1), or any
// diagonal the caller supplies as an array of inverse entries (minv; a user-supplied diagonal preconditioner):
// phat / shat are never stored, the scaling is applied to the five stencil inputs (same roundings as a stored
// product).  rtilde is b itself (x0 = 0), s aliases r.  Vectors: b, x, r, p, v, t -- t borrows the context's
// previous-step velocity buffer, which is dead between time steps.
//
// One thread per cell, grid-stride, cooperative grid; every reduction goes through per-CTA partials folded in CTA
// order by every CTA (fold_partials), so all CTAs hold bit-identical scalars and take the same exits.  Each of
// the seven reductions of an iteration has its own partial slot: a slot is rewritten one whole iteration (several
// grid syncs) after it was folded.  Clarity over speed: this is not a hot-path kernel.  Single-GPU contexts only.
#pragma once

#include <cooperative_groups.h>

#include "pf_cg.cuh"

namespace pf {

struct BiArgs {
  Slab s;
  const double* b;
  double* x;
  double* r;
  double* p;
  double* v;
  double* t;
  double* part;       // [7][CG_MAX_GRID]
  double* scal;
  const double* minv; // general inverse diagonal (device field layout) or nullptr = ring identity / invd
  double dx, dy, invd, tol;
  int maxiter;
  int ring5;          // != 0: the assembled ring-identity 5-point matrix (MODE_RING5 of pf_cg_generic.cuh) instead of the pinned operator
  double coef[5];
};

// M[cell]: the caller's diagonal if there is one, else 1 on the boundary ring and invd inside
__device__ __forceinline__ double bi_scale(const BiArgs& a, bool ring, size_t o) {
  return a.minv ? a.minv[o] : (ring ? 1.0 : a.invd);
}

static_assert(7 * CG_MAX_GRID <= PF_TRACE_OFFSET, "seven partial slots fit in front of the trace region");

// (A (M w))[jl, i] for the pinned operator: -(laplacian_coherent(M w)), row 0 replaced by the identity;
// SCALED = false: plain (A w)[jl, i]
template <bool UNIT, bool SCALED>
__device__ __forceinline__ double bi_apply_scaled(const BiArgs& a, const double* w, int jl, int i) {
  const Slab& s = a.s;
  const int jg = s.row0 + jl;
  const bool hw = i > 0, he = i < s.nx - 1, hs = jg > 0, hn = jg < s.nyg - 1;
  // ring membership of the four neighbours (the cell's own follows from the h* flags)
  const bool ring_c = !hw || !he || !hs || !hn;
  const bool ring_w = (i - 1 == 0) || !hs || !hn;
  const bool ring_e = (i + 1 == s.nx - 1) || !hs || !hn;
  const bool ring_s = (jg - 1 == 0) || !hw || !he;
  const bool ring_n = (jg + 1 == s.nyg - 1) || !hw || !he;
  const size_t oc = cell(s, jl, i);
  double c = w[oc];
  double ww = hw ? w[cell(s, jl, i - 1)] : 0.0;
  double e = he ? w[cell(s, jl, i + 1)] : 0.0;
  double so = hs ? w[cell(s, jl - 1, i)] : 0.0;
  double n = hn ? w[cell(s, jl + 1, i)] : 0.0;
  if (SCALED) {
    c = bi_scale(a, ring_c, oc) * c;
    if (hw) ww = bi_scale(a, ring_w, cell(s, jl, i - 1)) * ww;
    if (he) e = bi_scale(a, ring_e, cell(s, jl, i + 1)) * e;
    if (hs) so = bi_scale(a, ring_s, cell(s, jl - 1, i)) * so;
    if (hn) n = bi_scale(a, ring_n, cell(s, jl + 1, i)) * n;
  }
  if (a.ring5) return ring5_row(c, ww, e, so, n, !hw || !he || !hs || !hn, a.coef);
  const double nl = neg_lap<UNIT>(c, ww, e, so, n, hw, he, hs, hn, a.dx, a.dy);
  return (jg == 0 && i == 0) ? c : nl;
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS) k_bicgstab(BiArgs a) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  __shared__ double sm[32];
  const Slab& s = a.s;
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const long long ncell = (long long)s.nx * s.nyl;
  const long long stride = (long long)G * CG_THREADS;
  const long long first_cell = (long long)bid * CG_THREADS + tid;
  double* const P0 = a.part;
  double* const P1 = a.part + 1 * CG_MAX_GRID;
  double* const P2 = a.part + 2 * CG_MAX_GRID;
  double* const P3 = a.part + 3 * CG_MAX_GRID;
  double* const P4 = a.part + 4 * CG_MAX_GRID;
  double* const P5 = a.part + 5 * CG_MAX_GRID;
  double* const P6 = a.part + 6 * CG_MAX_GRID;
  const double breakdown = 2.220446049250313e-16 * 2.220446049250313e-16;     // np.finfo(float64).eps ** 2

  // x = 0, r = b, ||b||^2 (= rtilde.r of the first iteration: rtilde = r = b)
  double v0 = 0.0;
  for (long long k = first_cell; k < ncell; k += stride) {
    const size_t o = cell(s, (int)(k / s.nx), (int)(k % s.nx));
    const double bv = a.b[o];
    a.x[o] = 0.0;
    a.r[o] = bv;
    v0 += bv * bv;
  }
  v0 = block_sum(v0, sm);
  if (tid == 0) P6[bid] = v0;
  grid.sync();
  const double bb = fold_partials(P6, G, sm);
  const double bnorm = sqrt(bb);
  const double t0 = a.tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;            // max(0.0, rtol*||b||) with Python's NaN behaviour
  double rr = bb, rho = bb, rho_prev = 0.0, alpha = 0.0, omega = 0.0;
  int iters = 0, info = a.maxiter;
  if (bnorm == 0.0) {
    info = 0;                                           // "return b, 0": x was zeroed above
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) { info = 0; break; }
      if (fabs(rho) < breakdown) { info = -10; break; }
      if (it > 0) {
        if (fabs(omega) < breakdown) { info = -11; break; }
        const double beta = (rho / rho_prev) * (alpha / omega);
        for (long long k = first_cell; k < ncell; k += stride) {     // p -= omega*v; p *= beta; p += r
          const size_t o = cell(s, (int)(k / s.nx), (int)(k % s.nx));
          a.p[o] = ((a.p[o] - omega * a.v[o]) * beta) + a.r[o];
        }
      } else {
        for (long long k = first_cell; k < ncell; k += stride) {     // p = r.copy()
          const size_t o = cell(s, (int)(k / s.nx), (int)(k % s.nx));
          a.p[o] = a.r[o];
        }
      }
      grid.sync();
      double rv = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // v = A (M p);  rtilde.v
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const double vv = bi_apply_scaled<UNIT, true>(a, a.p, jl, i);
        a.v[cell(s, jl, i)] = vv;
        rv += a.b[cell(s, jl, i)] * vv;
      }
      rv = block_sum(rv, sm);
      if (tid == 0) P0[bid] = rv;
      grid.sync();
      rv = fold_partials(P0, G, sm);
      if (rv == 0.0) { info = -11; break; }
      alpha = rho / rv;
      double ss = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // r -= alpha*v  (s = r)
        const size_t o = cell(s, (int)(k / s.nx), (int)(k % s.nx));
        const double sv = a.r[o] - alpha * a.v[o];
        a.r[o] = sv;
        ss += sv * sv;
      }
      ss = block_sum(ss, sm);
      if (tid == 0) P1[bid] = ss;
      grid.sync();
      ss = fold_partials(P1, G, sm);
      if (sqrt(ss) < atol) {                                         // x += alpha*phat; return x, 0
        for (long long k = first_cell; k < ncell; k += stride) {
          const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
          const int jg = s.row0 + jl;
          const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
          const size_t o = cell(s, jl, i);
          a.x[o] = a.x[o] + alpha * (bi_scale(a, ring, o) * a.p[o]);
        }
        rr = ss;
        info = 0;
        break;
      }
      double ts = 0.0, tt = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // t = A (M s);  t.s, t.t
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const double tv = bi_apply_scaled<UNIT, true>(a, a.r, jl, i);
        a.t[cell(s, jl, i)] = tv;
        ts += tv * a.r[cell(s, jl, i)];
        tt += tv * tv;
      }
      ts = block_sum(ts, sm);
      tt = block_sum(tt, sm);
      if (tid == 0) { P2[bid] = ts; P3[bid] = tt; }
      grid.sync();
      ts = fold_partials(P2, G, sm);
      tt = fold_partials(P3, G, sm);
      omega = ts / tt;
      double nrr = 0.0, nrho = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // x += alpha*phat; x += omega*shat; r -= omega*t
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const int jg = s.row0 + jl;
        const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
        const size_t o = cell(s, jl, i);
        const double m = bi_scale(a, ring, o);
        const double sv = a.r[o];
        double xv = a.x[o] + alpha * (m * a.p[o]);
        xv = xv + omega * (m * sv);
        a.x[o] = xv;
        const double rn = sv - omega * a.t[o];
        a.r[o] = rn;
        nrr += rn * rn;
        nrho += a.b[o] * rn;
      }
      nrr = block_sum(nrr, sm);
      nrho = block_sum(nrho, sm);
      if (tid == 0) { P4[bid] = nrr; P5[bid] = nrho; }
      grid.sync();
      rr = fold_partials(P4, G, sm);
      rho_prev = rho;
      rho = fold_partials(P5, G, sm);
      ++iters;
    }
  }
  // true residual ||b - A x||^2 (interface.py:48-49)
  grid.sync();
  double res = 0.0;
  for (long long k = first_cell; k < ncell; k += stride) {
    const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
    const double d = a.b[cell(s, jl, i)] - bi_apply_scaled<UNIT, false>(a, a.x, jl, i);   // plain A x
    res += d * d;
  }
  res = block_sum(res, sm);
  if (tid == 0) P6[bid] = res;
  grid.sync();
  res = fold_partials(P6, G, sm);
  if (bid == 0 && tid == 0) {
    a.scal[S_INFO] = (double)info;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    a.scal[S_BB] = bb;
    a.scal[S_RES_SUMSQ] = res;
    a.scal[S_X0] = a.x[cell(s, 0, 0)];
  }
}

// ---- CG on the pinned operator with a caller-supplied diagonal preconditioner ------------------------------
// scipy's cg recurrence exactly as in pf_cg_generic.cuh (x0 = 0, strict test at the top of the loop, info =
// maxiter when exhausted), z = minv * r read from an array instead of the analytic ring-identity diagonal:
// `solve(A, b, method='cg', M=<any diagonal operator>)` of the linear-solve seam (interface.py:43).  q = A p is
// stored in a.v; a.t is unused.
template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS) k_cg_diag(BiArgs a) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  __shared__ double sm[32];
  const Slab& s = a.s;
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const long long ncell = (long long)s.nx * s.nyl;
  const long long stride = (long long)G * CG_THREADS;
  const long long first_cell = (long long)bid * CG_THREADS + tid;
  double* const P0 = a.part;
  double* const P1 = a.part + 1 * CG_MAX_GRID;
  double* const P2 = a.part + 2 * CG_MAX_GRID;
  double* const P5 = a.part + 5 * CG_MAX_GRID;
  double* const P6 = a.part + 6 * CG_MAX_GRID;

  double v0 = 0.0, v1 = 0.0;                            // x = 0, r = b, ||b||^2, b.(M b)
  for (long long k = first_cell; k < ncell; k += stride) {
    const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
    const int jg = s.row0 + jl;
    const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
    const size_t o = cell(s, jl, i);
    const double bv = a.b[o];
    a.x[o] = 0.0;
    a.r[o] = bv;
    v0 += bv * bv;
    v1 += bv * (bi_scale(a, ring, o) * bv);
  }
  v0 = block_sum(v0, sm);
  v1 = block_sum(v1, sm);
  if (tid == 0) { P6[bid] = v0; P5[bid] = v1; }
  grid.sync();
  const double bb = fold_partials(P6, G, sm);
  double rho = fold_partials(P5, G, sm);
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
        const size_t o = cell(s, jl, i);
        const double z = bi_scale(a, ring, o) * a.r[o];
        a.p[o] = (it == 0) ? z : (a.p[o] * beta) + z;
      }
      grid.sync();
      double pq = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // q = A p ; p.q
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const double qv = bi_apply_scaled<UNIT, false>(a, a.p, jl, i);
        a.v[cell(s, jl, i)] = qv;
        pq += a.p[cell(s, jl, i)] * qv;
      }
      pq = block_sum(pq, sm);
      if (tid == 0) P0[bid] = pq;
      grid.sync();
      pq = fold_partials(P0, G, sm);
      const double alpha = rho / pq;
      double arr = 0.0, arz = 0.0;
      for (long long k = first_cell; k < ncell; k += stride) {       // x += alpha p ; r -= alpha q
        const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
        const int jg = s.row0 + jl;
        const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
        const size_t o = cell(s, jl, i);
        a.x[o] = a.x[o] + alpha * a.p[o];
        const double rn = a.r[o] - alpha * a.v[o];
        a.r[o] = rn;
        arr += rn * rn;
        arz += rn * (bi_scale(a, ring, o) * rn);
      }
      arr = block_sum(arr, sm);
      arz = block_sum(arz, sm);
      if (tid == 0) { P1[bid] = arr; P2[bid] = arz; }
      grid.sync();
      rr = fold_partials(P1, G, sm);
      rho_prev = rho;
      rho = fold_partials(P2, G, sm);
      ++iters;
    }
  }
  grid.sync();
  double res = 0.0;                                      // true residual ||b - A x||^2 (interface.py:48-49)
  for (long long k = first_cell; k < ncell; k += stride) {
    const int jl = (int)(k / s.nx), i = (int)(k % s.nx);
    const double d = a.b[cell(s, jl, i)] - bi_apply_scaled<UNIT, false>(a, a.x, jl, i);
    res += d * d;
  }
  res = block_sum(res, sm);
  if (tid == 0) P6[bid] = res;
  grid.sync();
  res = fold_partials(P6, G, sm);
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
