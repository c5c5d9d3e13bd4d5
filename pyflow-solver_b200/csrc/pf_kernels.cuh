// pf_kernels.cuh -- per-step (non-CG) kernels and operator-level kernels.
//
// These touch each field once or twice per STEP (200 B/cell/step
// against 64 B/cell per CG iteration x ~300 iterations), so they are written
// for exactness and clarity: one thread per cell, rows walked by blockIdx.y,
// neighbours through L1.  Reductions: warp shuffle -> block -> per-block
// partial -> the last block to finish folds the partials in a fixed order
// (deterministic, single launch).
#pragma once

#include "pf_common.cuh"

namespace pf {

constexpr int KT = 256;          // threads per block of the simple kernels
constexpr int MAX_RED_BLOCKS = 4096;
constexpr int MAX_RED_VALS = 10;

// Reduction scratch shared by the simple kernels (one launch at a time).
struct Red {
  double* part;        // [MAX_RED_VALS][MAX_RED_BLOCKS]
  unsigned* counter;   // ticket, self-resetting
  double* scal;        // device scalar block
};

__device__ __forceinline__ int lin_block() { return blockIdx.y * gridDim.x + blockIdx.x; }
__device__ __forceinline__ int num_blocks() { return gridDim.x * gridDim.y; }

// Fold NV per-thread values (bit k of MAXMASK set => max, else sum) over the
// whole grid and store result k into scal[slots[k]].
template <int NV, unsigned MAXMASK>
__device__ __forceinline__ void grid_reduce_store(double (&vals)[NV], const Red& red, const int (&slots)[NV]) {
  __shared__ double sm[32];
  __shared__ bool is_last;
  const int nb = num_blocks();
  const int b = lin_block();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double t = ((MAXMASK >> k) & 1u) ? block_max(vals[k], sm) : block_sum(vals[k], sm);
    if (threadIdx.x == 0) red.part[(size_t)k * MAX_RED_BLOCKS + b] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(red.counter, 1u);
    is_last = (ticket == (unsigned)nb - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const bool mx = (MAXMASK >> k) & 1u;
    double acc = mx ? -INFINITY : 0.0;
    for (int q = threadIdx.x; q < nb; q += blockDim.x) {
      const double t = ld_cg(&red.part[(size_t)k * MAX_RED_BLOCKS + q]);
      acc = mx ? fmax(acc, t) : acc + t;
    }
    const double t = mx ? block_max(acc, sm) : block_sum(acc, sm);
    if (threadIdx.x == 0) red.scal[slots[k]] = t;
  }
  if (threadIdx.x == 0) *red.counter = 0u;
}

// ---- stencils --------------------------------------------------------------
// divergence_coherent at one cell (core_ops_coherent.py:63-79)
template <bool UNIT>
__device__ __forceinline__ double div_coh(const Slab& s, const double* u, const double* v,
                                          int jl, int i, double dx, double dy) {
  const int jg = s.row0 + jl;
  const double uc = u[cell(s, jl, i)];
  const double vc = v[cell(s, jl, i)];
  const double xt = (i > 0) ? dv<UNIT>(uc - u[cell(s, jl, i - 1)], dx) : dv<UNIT>(uc, dx);
  const double yt = (jg > 0) ? dv<UNIT>(vc - v[cell(s, jl - 1, i)], dy) : dv<UNIT>(vc, dy);
  return (0.0 + xt) + yt;
}

// fluid_ops._divergence_impl at one cell (numerics/fluid_ops.py:8-28).  The
// right column away from the corners is never written by the reference and
// stays 0.0 -- reproduced.
template <bool UNIT>
__device__ __forceinline__ double div_central(const Slab& s, const double* u, const double* v,
                                              int jl, int i, double dx, double dy) {
  const int jg = s.row0 + jl;
  const int nx = s.nx, ny = s.nyg;
  const bool left = (i == 0), right = (i == nx - 1);
  const bool bot = (jg == 0), top = (jg == ny - 1);
  if (right && !bot && !top) return 0.0;
  const double uc = u[cell(s, jl, i)];
  const double vc = v[cell(s, jl, i)];
  const double tdx = 2.0 * dx, tdy = 2.0 * dy;
  double xt, yt;
  if (left) {
    xt = dv<UNIT>(u[cell(s, jl, i + 1)] - uc, dx);
  } else if (right) {
    xt = dv<UNIT>(uc - u[cell(s, jl, i - 1)], dx);
  } else {
    const double d = u[cell(s, jl, i + 1)] - u[cell(s, jl, i - 1)];
    xt = UNIT ? d * 0.5 : dv<false>(d, tdx);      // d/2.0 == d*0.5 exactly
  }
  if (bot) {
    yt = dv<UNIT>(v[cell(s, jl + 1, i)] - vc, dy);
  } else if (top) {
    yt = dv<UNIT>(vc - v[cell(s, jl - 1, i)], dy);
  } else {
    const double d = v[cell(s, jl + 1, i)] - v[cell(s, jl - 1, i)];
    yt = UNIT ? d * 0.5 : dv<false>(d, tdy);
  }
  return xt + yt;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Convective derivative at one cell.  SCHEME = PF_ADV_UPWIND: advection.py:17-43
// (strict > 0, clamped neighbour).  PF_ADV_QUICK: advection.py:45-84 (>= 0,
// np.pad(field,2,'edge'), the reference's literal negative-velocity branch).
template <bool UNIT, int SCHEME>
__device__ __forceinline__ double advect_cell(const Slab& s, const double* f, double uc, double vc,
                                              int jl, int i, double dx, double dy) {
  const int jg = s.row0 + jl;
  const int nx = s.nx, ny = s.nyg;
  // row index clamped in GLOBAL coordinates, then mapped back to local
  auto row = [&](int k) { return clampi(jg + k, 0, ny - 1) - s.row0; };
  auto col = [&](int k) { return clampi(i + k, 0, nx - 1); };
  const double P = f[cell(s, jl, i)];
  if (SCHEME == PF_ADV_UPWIND) {
    const double fx = (uc > 0.0) ? dv<UNIT>(P - f[cell(s, jl, col(-1))], dx)
                                 : dv<UNIT>(f[cell(s, jl, col(+1))] - P, dx);
    const double fy = (vc > 0.0) ? dv<UNIT>(P - f[cell(s, row(-1), i)], dy)
                                 : dv<UNIT>(f[cell(s, row(+1), i)] - P, dy);
    return uc * fx + vc * fy;
  } else {
    const double E = f[cell(s, jl, col(+1))], W = f[cell(s, jl, col(-1))], WW = f[cell(s, jl, col(-2))];
    const double N = f[cell(s, row(+1), i)], S = f[cell(s, row(-1), i)], SS = f[cell(s, row(-2), i)];
    const double a = 0.375, b = 0.75, c = 0.125;
    double fe, fw, fn, fs;
    if (uc >= 0.0) {
      fe = a * E + b * P - c * W;
      fw = a * P + b * W - c * WW;
    } else {
      fe = a * W + b * P - c * E;
      fw = a * WW + b * W - c * P;
    }
    if (vc >= 0.0) {
      fn = a * N + b * P - c * S;
      fs = a * P + b * S - c * SS;
    } else {
      fn = a * S + b * P - c * N;
      fs = a * SS + b * S - c * P;
    }
    return dv<UNIT>(uc * fe - uc * fw, dx) + dv<UNIT>(vc * fn - vc * fs, dy);
  }
}

// ---- pre-step reductions (solver.py:25-26,45) -------------------------------
// vals: max|u|, max|v|, #nan(u), #nan(v), sum div_central(u,v;1,1)^2
__global__ void __launch_bounds__(KT) k_prestep(Slab s, const double* u, const double* v, Red red) {
  double vals[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const double uc = u[cell(s, jl, i)], vc = v[cell(s, jl, i)];
      vals[0] = fmax(vals[0], fabs(uc));
      vals[1] = fmax(vals[1], fabs(vc));
      vals[2] += (uc != uc) ? 1.0 : 0.0;
      vals[3] += (vc != vc) ? 1.0 : 0.0;
      const double d = div_central<true>(s, u, v, jl, i, 1.0, 1.0);
      vals[4] += d * d;
    }
  }
  const int slots[5] = {S_UMAX, S_VMAX, S_UNAN, S_VNAN, S_DIVC_SUMSQ};
  grid_reduce_store<5, 0x3u>(vals, red, slots);
}

// ---- predictor (solver.py:71-95): u* = u_old - dt*conv ------------------------
template <bool UNIT, int SCHEME>
__global__ void __launch_bounds__(KT) k_predict(Slab s, const double* uo, const double* vo,
                                                double* un, double* vn, double dt, double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    const double uc = uo[cell(s, jl, i)], vc = vo[cell(s, jl, i)];
    double us = uc, vs = vc;
    if (SCHEME != PF_ADV_NONE) {
      const double cu = advect_cell<UNIT, SCHEME>(s, uo, uc, vc, jl, i, dx, dy);
      const double cv = advect_cell<UNIT, SCHEME>(s, vo, uc, vc, jl, i, dx, dy);
      us = uc - dt * cu;
      vs = vc - dt * cv;
    }
    un[cell(s, jl, i)] = us;
    vn[cell(s, jl, i)] = vs;
  }
}

// ---- rhs pass 1 (pressure_solver.py:116-123) --------------------------------
// b = (-div_coherent(u*,v*)) / max(dt,1e-14);  sums: div^2, b
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_rhs1(Slab s, const double* u, const double* v, double* b,
                                             double dt_safe, double dx, double dy, Red red) {
  double vals[2] = {0.0, 0.0};
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const double d = div_coh<UNIT>(s, u, v, jl, i, dx, dy);
      const double rb = dv<false>(-d, dt_safe);       // invariant divisor: exact, see pf_common.cuh
      b[cell(s, jl, i)] = rb;
      vals[0] += d * d;
      vals[1] += rb;
    }
  }
  const int slots[2] = {S_DIVB_SUMSQ, S_RHS_SUM};
  grid_reduce_store<2, 0u>(vals, red, slots);
}

// ---- rhs pass 2 (pressure_solver.py:126-128 + CG init x0=0, r=b) ------------
// b -= mean; b[0] = 0; r = b; x = 0; sums: b.b, b.(M b)
__global__ void __launch_bounds__(KT) k_rhs2(Slab s, double* b, double* x, double* r, double ncells,
                                             int subtract_mean, int pin, double invd, Red red) {
  double vals[2] = {0.0, 0.0};
  // np.mean(rhs) = add.reduce(rhs) / N  (pressure_solver.py:126)
  const double mean = subtract_mean ? ld_cg(red.scal + S_RHS_SUM) / ncells : 0.0;
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const int jg = s.row0 + jl;
      double bv = b[cell(s, jl, i)];
      if (subtract_mean) bv = bv - mean;
      if (pin && jg == 0 && i == 0) bv = 0.0;
      b[cell(s, jl, i)] = bv;
      r[cell(s, jl, i)] = bv;
      x[cell(s, jl, i)] = 0.0;
      const bool ring = (i == 0) || (i == s.nx - 1) || (jg == 0) || (jg == s.nyg - 1);
      const double z = (ring ? 1.0 : invd) * bv;
      vals[0] += bv * bv;
      vals[1] += bv * z;
    }
  }
  const int slots[2] = {S_BB, S_BZ};
  grid_reduce_store<2, 0u>(vals, red, slots);
}

// ---- true residual (interface.py:48-49): sum (b - A x)^2, A pinned -lap ------
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_resid(Slab s, const double* b, const double* x,
                                              double dx, double dy, Red red) {
  double vals[1] = {0.0};
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const int jg = s.row0 + jl;
      const bool hw = i > 0, he = i < s.nx - 1, hs = jg > 0, hn = jg < s.nyg - 1;
      const double c = x[cell(s, jl, i)];
      const double w = hw ? x[cell(s, jl, i - 1)] : 0.0;
      const double e = he ? x[cell(s, jl, i + 1)] : 0.0;
      const double so = hs ? x[cell(s, jl - 1, i)] : 0.0;
      const double n = hn ? x[cell(s, jl + 1, i)] : 0.0;
      double ax = neg_lap<UNIT>(c, w, e, so, n, hw, he, hs, hn, dx, dy);
      if (jg == 0 && i == 0) {
        ax = c;
        red.scal[S_X0] = c;
      }
      const double res = b[cell(s, jl, i)] - ax;
      vals[0] += res * res;
    }
  }
  const int slots[1] = {S_RES_SUMSQ};
  grid_reduce_store<1, 0u>(vals, red, slots);
}

// ---- correction (pressure_solver.py:147-157) --------------------------------
// p = x - x[0]; u -= dt*dpdx; v -= dt*dpdy   (x doubles as the p field)
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_correct(Slab s, const double* x, double* pout, double* u, double* v,
                                                double dt, double dx, double dy, const double* scal) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  const double x0 = scal[S_X0];
  const bool shift = !(x0 == 0.0);      // exact no-op unless x[0] != 0 (or NaN)
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    const int jg = s.row0 + jl;
    const double pc = x[cell(s, jl, i)] - x0;
    double gx = 0.0, gy = 0.0;
    if (i < s.nx - 1) gx = dv<UNIT>((x[cell(s, jl, i + 1)] - x0) - pc, dx);
    if (jg < s.nyg - 1) gy = dv<UNIT>((x[cell(s, jl + 1, i)] - x0) - pc, dy);
    u[cell(s, jl, i)] = u[cell(s, jl, i)] - dt * gx;
    v[cell(s, jl, i)] = v[cell(s, jl, i)] - dt * gy;
    if (shift) pout[cell(s, jl, i)] = pc;
  }
}

// ---- divergence after correction (pressure_solver.py:165-169) ---------------
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_divnorm(Slab s, const double* u, const double* v,
                                                double dx, double dy, Red red, int slot) {
  double vals[1] = {0.0};
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const double d = div_coh<UNIT>(s, u, v, jl, i, dx, dy);
      vals[0] += d * d;
    }
  }
  const int slots[1] = {slot};
  grid_reduce_store<1, 0u>(vals, red, slots);
}

// ---- lid BC (solver.py:104-105): last interior row, every interior column ---
__global__ void k_lid(Slab s, double* u, double lid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int jl = s.nyg - 1 - s.row0;
  if (i < s.nx && jl >= 0 && jl < s.nyl) u[cell(s, jl, i)] = lid;
}

// ---- post-step pass (solver.py:108-121 + driver :89-96 + next step's :25-26) -
// sums: (u-uo)^2, (v-vo)^2, div_central^2, #nonfinite(u,v,p), #nan u, #nan v;
// max: |u|, |v|;  centerline[jl] = u[jl, nx/2]
__global__ void __launch_bounds__(KT) k_post(Slab s, const double* u, const double* v, const double* uo,
                                             const double* vo, const double* p, double* center, Red red) {
  double vals[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const double uc = u[cell(s, jl, i)], vc = v[cell(s, jl, i)], pc = p[cell(s, jl, i)];
      const double du = uc - uo[cell(s, jl, i)];
      const double dvv = vc - vo[cell(s, jl, i)];
      vals[0] = fmax(vals[0], fabs(uc));
      vals[1] = fmax(vals[1], fabs(vc));
      vals[2] += du * du;
      vals[3] += dvv * dvv;
      const double d = div_central<true>(s, u, v, jl, i, 1.0, 1.0);
      vals[4] += d * d;
      vals[5] += ((isfinite(uc) && isfinite(vc) && isfinite(pc)) ? 0.0 : 1.0);
      vals[6] += (uc != uc) ? 1.0 : 0.0;
      vals[7] += (vc != vc) ? 1.0 : 0.0;
      if (i == s.nx / 2) center[jl] = uc;
    }
  }
  const int slots[8] = {S_UMAX, S_VMAX, S_RU_SUMSQ, S_RV_SUMSQ, S_DIVC_SUMSQ, S_NONFINITE, S_UNAN, S_VNAN};
  grid_reduce_store<8, 0x3u>(vals, red, slots);
}

// ---- divergence statistics of the alternate driver (driver/simulation_driver.py:126-130, :78-81):
// pass 0: sum of div_central(u,v;dx,dy);  pass 1: sum of (div - mean)^2 with mean = S_TMP0 / ncells
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_divc_stats(Slab s, const double* u, const double* v, double dx, double dy,
                                                   int pass, double ncells, Red red) {
  double vals[1] = {0.0};
  const double mean = pass ? ld_cg(red.scal + S_TMP0) / ncells : 0.0;
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i < s.nx) {
    for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
      const double d = div_central<UNIT>(s, u, v, jl, i, dx, dy);
      vals[0] += pass ? (d - mean) * (d - mean) : d;
    }
  }
  const int slots[1] = {pass ? S_TMP1 : S_TMP0};
  grid_reduce_store<1, 0u>(vals, red, slots);
}

__global__ void __launch_bounds__(KT) k_centerline(Slab s, const double* u, double* center) {
  const int jl = blockIdx.x * KT + threadIdx.x;
  if (jl < s.nyl) center[jl] = u[cell(s, jl, s.nx / 2)];
}

// ---- operator-level kernels (write whole fields) ----------------------------
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_op_grad(Slab s, const double* p, double* gx, double* gy, double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    const int jg = s.row0 + jl;
    const double c = p[cell(s, jl, i)];
    gx[cell(s, jl, i)] = (i < s.nx - 1) ? dv<UNIT>(p[cell(s, jl, i + 1)] - c, dx) : 0.0;
    gy[cell(s, jl, i)] = (jg < s.nyg - 1) ? dv<UNIT>(p[cell(s, jl + 1, i)] - c, dy) : 0.0;
  }
}

template <bool UNIT>
__global__ void __launch_bounds__(KT) k_op_div(Slab s, const double* u, const double* v, double* out,
                                               double dx, double dy, int central) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y)
    out[cell(s, jl, i)] = central ? div_central<UNIT>(s, u, v, jl, i, dx, dy) : div_coh<UNIT>(s, u, v, jl, i, dx, dy);
}

template <bool UNIT>
__global__ void __launch_bounds__(KT) k_op_lap(Slab s, const double* p, double* out, double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    const int jg = s.row0 + jl;
    const bool hw = i > 0, he = i < s.nx - 1, hs = jg > 0, hn = jg < s.nyg - 1;
    const double c = p[cell(s, jl, i)];
    const double w = hw ? p[cell(s, jl, i - 1)] : 0.0;
    const double e = he ? p[cell(s, jl, i + 1)] : 0.0;
    const double so = hs ? p[cell(s, jl - 1, i)] : 0.0;
    const double n = hn ? p[cell(s, jl + 1, i)] : 0.0;
    out[cell(s, jl, i)] = -neg_lap<UNIT>(c, w, e, so, n, hw, he, hs, hn, dx, dy);
  }
}

// ---- legacy central operators of fluid_ops.py (off the time-step path; the
// reference's own projection tests import them) ------------------------------
// fluid_ops.gradient at one cell (numerics/fluid_ops.py:43-67): the two clamped
// neighbours' difference over 2*h inside, over h in the first / last column (row).
// DIR 0: d/dx, DIR 1: d/dy.
template <bool UNIT, int DIR>
__device__ __forceinline__ double grad_central(const Slab& s, const double* p, int jl, int i, double h) {
  double d;
  bool inner;
  if (DIR == 0) {
    inner = (i > 0) && (i < s.nx - 1);
    d = p[cell(s, jl, clampi(i + 1, 0, s.nx - 1))] - p[cell(s, jl, clampi(i - 1, 0, s.nx - 1))];
  } else {
    const int jg = s.row0 + jl;
    inner = (jg > 0) && (jg < s.nyg - 1);
    d = p[cell(s, clampi(jg + 1, 0, s.nyg - 1) - s.row0, i)] - p[cell(s, clampi(jg - 1, 0, s.nyg - 1) - s.row0, i)];
  }
  if (inner) return UNIT ? d * 0.5 : dv<false>(d, 2.0 * h);      // d/2.0 == d*0.5 exactly
  return dv<UNIT>(d, h);
}

template <bool UNIT>
__global__ void __launch_bounds__(KT) k_op_grad_central(Slab s, const double* p, double* gx, double* gy,
                                                        double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    gx[cell(s, jl, i)] = grad_central<UNIT, 0>(s, p, jl, i, dx);
    gy[cell(s, jl, i)] = grad_central<UNIT, 1>(s, p, jl, i, dy);
  }
}

// fluid_ops.laplacian (numerics/fluid_ops.py:69-91): 5-point second differences inside;
// every wall cell copies its nearest interior cell (rows are copied first, then columns,
// so a corner takes its diagonal neighbour); all zero unless nx > 2 and ny > 2.
template <bool UNIT>
__global__ void __launch_bounds__(KT) k_op_lap_central(Slab s, const double* p, double* out, double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  const bool any = (s.nx > 2) && (s.nyg > 2);
  const double dx2 = dx * dx, dy2 = dy * dy;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y) {
    double r = 0.0;
    if (any) {
      const int ic = clampi(i, 1, s.nx - 2);
      const int jc = clampi(s.row0 + jl, 1, s.nyg - 2) - s.row0;
      const double tc = 2.0 * p[cell(s, jc, ic)];
      const double xt = dv<UNIT>((p[cell(s, jc, ic + 1)] - tc) + p[cell(s, jc, ic - 1)], dx2);
      const double yt = dv<UNIT>((p[cell(s, jc + 1, ic)] - tc) + p[cell(s, jc - 1, ic)], dy2);
      r = xt + yt;
    }
    out[cell(s, jl, i)] = r;
  }
}

// one component of fluid_ops.apply_pressure_correction (numerics/fluid_ops.py:108-125):
// out = f - dt * d(p)/d(DIR) with the central gradient above.
template <bool UNIT, int DIR>
__global__ void __launch_bounds__(KT) k_op_correct_central(Slab s, const double* p, const double* f, double* out,
                                                           double dt, double h) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y)
    out[cell(s, jl, i)] = f[cell(s, jl, i)] - dt * grad_central<UNIT, DIR>(s, p, jl, i, h);
}

template <bool UNIT, int SCHEME>
__global__ void __launch_bounds__(KT) k_op_advect(Slab s, const double* u, const double* v, const double* f,
                                                  double* out, double dx, double dy) {
  const int i = blockIdx.x * KT + threadIdx.x;
  if (i >= s.nx) return;
  for (int jl = blockIdx.y; jl < s.nyl; jl += gridDim.y)
    out[cell(s, jl, i)] = advect_cell<UNIT, SCHEME>(s, f, u[cell(s, jl, i)], v[cell(s, jl, i)], jl, i, dx, dy);
}

}  // namespace pf
