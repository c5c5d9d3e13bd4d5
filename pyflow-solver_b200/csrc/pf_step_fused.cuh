// pf_step_fused.cuh -- the per-step work around the CG solve in TWO passes (single-GPU path of pf_step).
//
//   k_front  = predictor (solver.py:71-95) + rhs pass 1 (pressure_solver.py:116-123)
//              reads u, v once; writes u*, v*, b                                         40 B / cell
//   k_finish = true residual (interface.py:48-49) + correction (pressure_solver.py:147-157) + divergence after
//              (:165-169) + lid row (solver.py:104-105) + residual / continuity / finite pass (solver.py:108-121)
//              + centre line + next step's max|u|, max|v|
//              reads b, x, u*, v*, u_old, v_old once; writes u, v                        64 B / cell
//
// against seven passes and 168 B / cell for the same work in pf_kernels.cuh (which stay: slab runs on several GPUs,
// whose halo rows arrive between the passes, the operator-level seams and tiny grids use them, and the GPU tests hold
// the two paths against each other).  Every value is formed by the SAME expression, in the same operation order, as in
// the separate kernels -- intermediate fields (u after the correction, v after the correction, p - p[0]) simply never
// travel through memory; where a neighbour's intermediate value is needed it is recomputed from the inputs by that
// expression or fetched from the neighbouring lane.  Results are therefore bit-identical to the separate kernels
// (element-wise fields) and differ only in summation order (reductions).
//
// Work distribution: a warp marches up a band of rows.  A lane holds TWO adjacent columns (16-byte loads / stores); the
// rows a stencil needs below and above live in a register window that slides one row per step, so every operand is
// loaded once per band.  Lanes 2..30 OWN their pair (store it, add it to the sums); lanes 0, 1 and 31 hold the pairs
// just left / right of the warp's 58 columns and only supply neighbours through shuffles -- no lane ever needs a value
// from another warp, no shared memory, no block barrier inside the sweep, and all 32 lanes run the same code.  Grid:
// x = column blocks of 58 * (warps per block) columns, y = bands.  u and v are written OUT OF PLACE (into the buffers
// that held the previous step's fields, each cell after its owner read the old value): a neighbour's u*, v* can be read
// at any time.
//
// Latency: with ~110 registers per thread only 16 warps fit on an SM, so a sweep that waited for its loads where it
// uses them ran at a quarter of the issue rate and no faster than the seven kernels together (first B200 run).  Hence
// (i) every load of a row step is UNCONDITIONAL and stands at the top of the step: a lane without a column, or a row
// beyond the wall, reads a clamped, valid address and the value is discarded by the "neighbour exists" selects that
// guard every use anyway; (ii) the loads are issued one row step AHEAD of their use (PF_FUSED_PREFETCH).
// Measured on a B200 at 4096^2 (profiles/r2b_step_passes_*.txt): k_finish 217 us against 432 us for the five kernels it
// replaces (DRAM traffic 1.06 GB = the 64 B / cell, 76 % of the copy peak; still latency-bound: long-scoreboard 3.7 per
// issue at 16 warps per SM), k_front 223 us against 253 us (issue-bound: 63 % issue utilisation, QUICK's arithmetic and
// the two IEEE divisions of the rhs are 530 instructions per 58 cells); the non-CG part of a step 810 -> 553 us.
// Build knobs tried and left at their defaults: no prefetch 570 us, 128-thread blocks x 3 per SM (140 registers) 682 us,
// row loop unrolled by 2 / 4 (window rotation by renaming) 545 / 560 us, QUICK with operand selects 543 us.
#pragma once

#include "pf_kernels.cuh"

namespace pf {

#ifndef PF_FUSED_FT
#define PF_FUSED_FT 256
#endif
#ifndef PF_FUSED_MINB
#define PF_FUSED_MINB 2
#endif
#ifndef PF_FUSED_PREFETCH
#define PF_FUSED_PREFETCH 1
#endif
#ifndef PF_FUSED_UNROLL
#define PF_FUSED_UNROLL 1
#endif
constexpr int FT = PF_FUSED_FT;     // threads per block (at most; narrow grids launch fewer warps)
constexpr int F_MINB = PF_FUSED_MINB;            // blocks per SM the register budget is set for
constexpr bool F_PREFETCH = PF_FUSED_PREFETCH;   // loads run one row step ahead of their use
constexpr int F_UNROLL = PF_FUSED_UNROLL;        // row steps per loop body (> 1: the window rotation becomes renaming)
constexpr int F_OWN = 58;           // columns a warp owns (lanes 2..30)

// Launch geometry of the two passes (host side; shared by pf_api.cu and the host emulation of the tests): a block is
// up to FT / 32 warps side by side (as few column blocks as cover the row, warps spread evenly over them), a band as
// many rows as leave about four blocks per SM -- at least 2, every band pays a few halo rows.  threads = 0: the grid is
// too small (or too wide for the reduction scratch) and the separate kernels of pf_kernels.cuh are used.
struct FusedGeometry {
  int gx, gy, threads, rows;
};
inline FusedGeometry fused_geometry(int nx, int nyl, int sm_count) {
  FusedGeometry g = {0, 0, 0, 0};
  if (nx < 4 || nyl < 4) return g;
  const int nwarps = (nx + F_OWN - 1) / F_OWN;
  const int gx = (nwarps + FT / 32 - 1) / (FT / 32);
  const int wpb = (nwarps + gx - 1) / gx;
  const int sms = sm_count > 0 ? sm_count : 1;
  const int gy_target = (4 * sms / gx > 1) ? 4 * sms / gx : 1;
  const int rows_even = (nyl + gy_target - 1) / gy_target;
  const int rows = rows_even > 2 ? rows_even : 2;
  const int gy = (nyl + rows - 1) / rows;
  if ((long long)gx * gy > MAX_RED_BLOCKS) return g;
  g.gx = gx; g.gy = gy; g.threads = 32 * wpb; g.rows = rows;
  return g;
}

struct FCols {
  int c0;        // first column of the lane's pair (even; negative on lanes 0 and 1 of the first warp of a row)
  int cl;        // the column the lane LOADS from: c0 clamped into the row (even, so the pair lies inside the pitch)
  bool va, vb;   // the grid has column c0 / c0 + 1
  bool own;      // lanes 2..30
};

__device__ __forceinline__ FCols fused_cols(const Slab& s) {
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
  const int wpb = (int)blockDim.x >> 5;
  FCols c;
  c.c0 = F_OWN * ((int)blockIdx.x * wpb + warp) + 2 * (lane - 2);
  c.cl = clampi(c.c0, 0, (s.nx - 1) & ~1);
  c.va = c.c0 >= 0 && c.c0 < s.nx;
  c.vb = c.c0 >= 0 && c.c0 + 1 < s.nx;
  c.own = lane >= 2 && lane <= 30;
  return c;
}

// Offset (in doubles) of the lane's load column in local row jl.
__device__ __forceinline__ size_t fused_off(const Slab& s, int jl, const FCols& c) { return cell(s, jl, c.cl); }

// Stores go to the lane's own pair (caller: own && va); padding columns are never stored to.
__device__ __forceinline__ void st_pair(double* f, size_t off, const FCols& c, double a, double b) {
  if (c.vb) st2(f + off, a, b);
  else st1(f + off, a);
}

__device__ __forceinline__ double shfl_left(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_right(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// advect_cell of pf_kernels.cuh on values: P = f[j,i]; E, W, WW, N, S, SS its neighbours with np.pad(..., 'edge')
// clamping already applied; uc, vc the advecting velocity at the cell.  (advection.py:17-84)
template <bool UNIT, int SCHEME>
__device__ __forceinline__ double advect_vals(double P, double E, double W, double WW, double N, double S, double SS,
                                              double uc, double vc, double dx, double dy) {
  if (SCHEME == PF_ADV_UPWIND) {
    const double fx = (uc > 0.0) ? dv<UNIT>(P - W, dx) : dv<UNIT>(E - P, dx);
    const double fy = (vc > 0.0) ? dv<UNIT>(P - S, dy) : dv<UNIT>(N - P, dy);
    return uc * fx + vc * fy;
  } else {
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

// div_central of pf_kernels.cuh on values (fluid_ops.py:8-28; the right column away from the corners stays 0.0).
template <bool UNIT>
__device__ __forceinline__ double div_central_vals(double uc, double uw, double ue, double vc, double vs, double vn,
                                                   bool left, bool right, bool bot, bool top, double dx, double dy) {
  if (right && !bot && !top) return 0.0;
  const double tdx = 2.0 * dx, tdy = 2.0 * dy;
  double xt, yt;
  if (left) {
    xt = dv<UNIT>(ue - uc, dx);
  } else if (right) {
    xt = dv<UNIT>(uc - uw, dx);
  } else {
    const double d = ue - uw;
    xt = UNIT ? d * 0.5 : dv<false>(d, tdx);
  }
  if (bot) {
    yt = dv<UNIT>(vn - vc, dy);
  } else if (top) {
    yt = dv<UNIT>(vc - vs, dy);
  } else {
    const double d = vn - vs;
    yt = UNIT ? d * 0.5 : dv<false>(d, tdy);
  }
  return xt + yt;
}

// div_coh of pf_kernels.cuh on values (core_ops_coherent.py:63-79).
template <bool UNIT>
__device__ __forceinline__ double div_coh_vals(double uc, double uw, double vc, double vs, bool hw, bool hs,
                                               double dx, double dy) {
  const double xt = hw ? dv<UNIT>(uc - uw, dx) : dv<UNIT>(uc, dx);
  const double yt = hs ? dv<UNIT>(vc - vs, dy) : dv<UNIT>(vc, dy);
  return (0.0 + xt) + yt;
}

// ---- pass 1: predictor + rhs pass 1 -----------------------------------------------------------------------------------
// (u, v) -> (us, vs) = u*, v*;  b = (-div_coherent(u*, v*)) / max(dt, 1e-14);  sums: div^2, b.
// The band's first row needs v* of the row below it: one warm-up row is computed (not stored) before the band.
// Row step j loads row j + 2 (QUICK's window: j - 2 .. j + 1), i.e. one step ahead of its use.
template <bool UNIT, int SCHEME>
__global__ void __launch_bounds__(FT, F_MINB)
k_front(Slab s, const double* u, const double* v, double* us, double* vs, double* b, double dt, double dt_safe,
        double dx, double dy, int band_rows, Red red) {
  const FCols c = fused_cols(s);
  const int top = s.nyg - 1 - s.row0;                      // local index of the global top row
  const int bot = -s.row0;                                 // ... and of the global bottom row
  const int y0 = (int)blockIdx.y * band_rows;
  const int y1 = (y0 + band_rows < s.nyl) ? y0 + band_rows : s.nyl;
  const bool has_w = c.c0 > 0;                             // column c0 - 1 exists
  const bool has_e = c.c0 + 2 < s.nx;                      // column c0 + 2 exists
  const size_t pitch = (size_t)s.pitch;
  double vals[2] = {0.0, 0.0};
  if (y0 < y1) {
    const int js = (y0 > bot) ? y0 - 1 : y0;
    auto rowc = [&](int jl) { return clampi(jl, bot, top); };       // np.pad(..., 'edge') in y
    double2 u2, u1, uc, un, v2, v1, vc, vn;
    uc = ld2(u + fused_off(s, js, c));
    vc = ld2(v + fused_off(s, js, c));
    u2 = u1 = un = uc;
    v2 = v1 = vn = vc;
    if (SCHEME != PF_ADV_NONE) {
      u2 = ld2(u + fused_off(s, rowc(js - 2), c)); u1 = ld2(u + fused_off(s, rowc(js - 1), c));
      un = ld2(u + fused_off(s, rowc(js + 1), c));
      v2 = ld2(v + fused_off(s, rowc(js - 2), c)); v1 = ld2(v + fused_off(s, rowc(js - 1), c));
      vn = ld2(v + fused_off(s, rowc(js + 1), c));
    }
    double2 vs_south = make_double2(0.0, 0.0);             // v* of row j - 1
    size_t off = fused_off(s, js, c);                      // row j, the lane's load column
#pragma unroll(F_UNROLL)
    for (int j = js; j < y1; ++j, off += pitch) {
      // the row the NEXT step needs: j + 2 (the top row again beyond the wall), resp. j + 1 without advection
      const int ahead = (SCHEME != PF_ADV_NONE) ? (top - j < 2 ? top - j : 2) : (j + 1 < y1 ? 1 : 0);
      const double2 nu = ld2(u + off + (size_t)ahead * pitch);
      const double2 nv = ld2(v + off + (size_t)ahead * pitch);
      double2 ust = uc, vst = vc;
      if (SCHEME != PF_ADV_NONE) {
        // the pair to the left (x = column c0 - 2, y = c0 - 1) and the first cell to the right (column c0 + 2)
        const double ulx = shfl_left(uc.x), uly = shfl_left(uc.y), urx = shfl_right(uc.x);
        const double vlx = shfl_left(vc.x), vly = shfl_left(vc.y), vrx = shfl_right(vc.x);
        // np.pad(field, 2, 'edge'): a neighbour beyond a wall is the wall cell itself
        const double uW_a = has_w ? uly : uc.x, uWW_a = has_w ? ulx : uc.x, uE_a = c.vb ? uc.y : uc.x;
        const double uW_b = uc.x, uWW_b = has_w ? uly : uc.x, uE_b = has_e ? urx : uc.y;
        const double vW_a = has_w ? vly : vc.x, vWW_a = has_w ? vlx : vc.x, vE_a = c.vb ? vc.y : vc.x;
        const double vW_b = vc.x, vWW_b = has_w ? vly : vc.x, vE_b = has_e ? vrx : vc.y;
        const double cu_a = advect_vals<UNIT, SCHEME>(uc.x, uE_a, uW_a, uWW_a, un.x, u1.x, u2.x, uc.x, vc.x, dx, dy);
        const double cv_a = advect_vals<UNIT, SCHEME>(vc.x, vE_a, vW_a, vWW_a, vn.x, v1.x, v2.x, uc.x, vc.x, dx, dy);
        const double cu_b = advect_vals<UNIT, SCHEME>(uc.y, uE_b, uW_b, uWW_b, un.y, u1.y, u2.y, uc.y, vc.y, dx, dy);
        const double cv_b = advect_vals<UNIT, SCHEME>(vc.y, vE_b, vW_b, vWW_b, vn.y, v1.y, v2.y, uc.y, vc.y, dx, dy);
        ust.x = uc.x - dt * cu_a; vst.x = vc.x - dt * cv_a;
        ust.y = uc.y - dt * cu_b; vst.y = vc.y - dt * cv_b;
      }
      const double ust_w = shfl_left(ust.y);               // u* of column c0 - 1
      if (j >= y0 && c.own && c.va) {
        const bool hs = j > bot;
        const size_t so = off + (size_t)(c.c0 - c.cl);      // own && va: c0 == cl
        const double d_a = div_coh_vals<UNIT>(ust.x, ust_w, vst.x, vs_south.x, has_w, hs, dx, dy);
        const double rb_a = dv<false>(-d_a, dt_safe);      // invariant divisor: exact, see pf_common.cuh
        double rb_b = 0.0;
        vals[0] += d_a * d_a;
        vals[1] += rb_a;
        if (c.vb) {
          const double d_b = div_coh_vals<UNIT>(ust.y, ust.x, vst.y, vs_south.y, true, hs, dx, dy);
          rb_b = dv<false>(-d_b, dt_safe);
          vals[0] += d_b * d_b;
          vals[1] += rb_b;
        }
        st_pair(us, so, c, ust.x, ust.y);
        st_pair(vs, so, c, vst.x, vst.y);
        st_pair(b, so, c, rb_a, rb_b);
      }
      vs_south = vst;
      if (SCHEME != PF_ADV_NONE) {
        u2 = u1; u1 = uc; uc = un; un = nu;
        v2 = v1; v1 = vc; vc = vn; vn = nv;
      } else {
        uc = nu;
        vc = nv;
      }
    }
  }
  const int slots[2] = {S_DIVB_SUMSQ, S_RHS_SUM};
  grid_reduce_store<2, 0u>(vals, red, slots);
}

// ---- pass 2: everything after the solve ---------------------------------------------------------------------------------
// inputs: b, x (the solution; doubles as the pressure field), us / vs = u*, v*; uout / vout hold the PREVIOUS step's
// fields on entry (POST) and the new fields on return.
// vals: ||b - A x||^2, ||div_coherent(u, v)||^2 after the correction; POST adds max|u|, max|v|, sum (u-uo)^2,
// sum (v-vo)^2, sum div_central^2, #nonfinite, #nan u, #nan v (the slots k_post writes).
struct FinRow {              // what row step j loads
  double2 X2;                // x   of row j + 2
  double2 VS1;               // v*  of row j + 1
  double2 U, B, UO, VO;      // u*, b, previous u, previous v of row j
};

template <bool UNIT, bool POST>
__global__ void __launch_bounds__(FT, F_MINB)
k_finish(Slab s, const double* b, double* x, const double* us, const double* vs, double* uout, double* vout, double dt,
         double dx, double dy, int lid_set, double lid, double* center, int band_rows, Red red) {
  constexpr int NV = POST ? 10 : 2;
  const FCols c = fused_cols(s);
  const int nx = s.nx;
  const int top = s.nyg - 1 - s.row0;
  const int bot = -s.row0;
  const int y0 = (int)blockIdx.y * band_rows;
  const int y1 = (y0 + band_rows < s.nyl) ? y0 + band_rows : s.nyl;
  const bool has_w = c.c0 > 0;                 // column c0 - 1 exists
  const bool he_b = c.c0 + 2 < nx;             // column c0 + 2 exists (east neighbour of the second cell)
  const size_t pitch = (size_t)s.pitch;
  double vals[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) vals[k] = 0.0;
  if (y0 < y1) {
    // p = x - x[0] (pressure_solver.py:148): an exact no-op unless x[0] != 0 (x[0] stays 0 under the pin; NaN poisons all)
    const double x0 = ld1(x + cell(s, bot, 0));
    const bool shift = !(x0 == 0.0);
    // v after the correction at local row jl: v* - dt * dp/dy, dp/dy = 0 in the top row (gradient_coherent)
    auto v_corrected = [&](const double2& vstar, const double2& xr, const double2& xabove, int jl) {
      double2 o;
      const bool inner = jl < top;
      const double gya = inner ? dv<UNIT>((xabove.x - x0) - (xr.x - x0), dy) : 0.0;
      const double gyb = inner ? dv<UNIT>((xabove.y - x0) - (xr.y - x0), dy) : 0.0;
      o.x = vstar.x - dt * gya;
      o.y = vstar.y - dt * gyb;
      return o;
    };
    // all loads of a row step, unconditional: rows beyond the top wall read the top row (discarded by the selects)
    auto load_row = [&](int j) {
      const size_t o = fused_off(s, j, c);
      const int d1 = top - j < 1 ? top - j : 1, d2 = top - j < 2 ? top - j : 2;
      FinRow r;
      r.X2 = ld2(x + o + (size_t)d2 * pitch);
      r.VS1 = ld2(vs + o + (size_t)d1 * pitch);
      r.U = ld2(us + o);
      r.B = ld2(b + o);
      if (POST) {
        r.UO = ld2(uout + o);
        r.VO = ld2(vout + o);
      } else {
        r.UO = r.VO = r.U;
      }
      return r;
    };
    const bool below = y0 > bot;
    const int ym = below ? y0 - 1 : y0, yp = y0 < top ? y0 + 1 : y0;
    double2 Xm = ld2(x + fused_off(s, ym, c));
    double2 Xc = ld2(x + fused_off(s, y0, c));
    double2 Xp = ld2(x + fused_off(s, yp, c));
    const double2 vs_m = ld2(vs + fused_off(s, ym, c)), vs_c = ld2(vs + fused_off(s, y0, c));
    FinRow cur = load_row(y0);
    double2 Vm = v_corrected(vs_m, Xm, Xc, y0 - 1);          // unused when there is no row below
    double2 Vc = v_corrected(vs_c, Xc, Xp, y0);
    size_t off = fused_off(s, y0, c);
#pragma unroll(F_UNROLL)
    for (int j = y0; j < y1; ++j, off += pitch) {
      FinRow nxt;
      if (F_PREFETCH) nxt = load_row(j + 1 < y1 ? j + 1 : j);      // the next step's operands, before this step's math
      else if (j > y0) cur = load_row(j);
      const bool hs = j > bot, hn = j < top;
      const double2 Xpp = cur.X2;
      const double2 Vp = v_corrected(cur.VS1, Xp, Xpp, j + 1);     // unused in the top row
      const double2 U = cur.U, B = cur.B, UO = cur.UO, VO = cur.VO;
      const double xw = shfl_left(Xc.y);        // x at column c0 - 1
      const double xe = shfl_right(Xc.x);       // x at column c0 + 2
      // u after the correction: u* - dt * dp/dx, dp/dx = 0 in the last column
      const double pa = Xc.x - x0, pb = Xc.y - x0;
      const double gx_a = c.vb ? dv<UNIT>(pb - pa, dx) : 0.0;
      const double gx_b = he_b ? dv<UNIT>((xe - x0) - pb, dx) : 0.0;
      const double un_a = U.x - dt * gx_a, un_b = U.y - dt * gx_b;
      const double un_w = shfl_left(un_b);      // corrected u at column c0 - 1
      const double un_e = shfl_right(un_a);     // corrected u at column c0 + 2
      if (c.own && c.va) {
        const int jg = s.row0 + j;
        // true residual b - A x, A = pinned -laplacian_coherent
        double ax_a = neg_lap<UNIT>(Xc.x, has_w ? xw : 0.0, c.vb ? Xc.y : 0.0, hs ? Xm.x : 0.0, hn ? Xp.x : 0.0,
                                    has_w, c.vb, hs, hn, dx, dy);
        if (jg == 0 && c.c0 == 0) {
          ax_a = Xc.x;
          red.scal[S_X0] = Xc.x;
        }
        const double res_a = B.x - ax_a;
        vals[0] += res_a * res_a;
        // divergence after the correction (before the lid row is imposed)
        const double da = div_coh_vals<UNIT>(un_a, un_w, Vc.x, Vm.x, has_w, hs, dx, dy);
        vals[1] += da * da;
        double uf_a = un_a, uf_b = un_b, uf_w = un_w, uf_e = un_e;
        if (POST && lid_set && jg == s.nyg - 1) uf_a = uf_b = uf_w = uf_e = lid;
        if (c.vb) {
          const double ax_b = neg_lap<UNIT>(Xc.y, Xc.x, he_b ? xe : 0.0, hs ? Xm.y : 0.0, hn ? Xp.y : 0.0,
                                            true, he_b, hs, hn, dx, dy);
          const double res_b = B.y - ax_b;
          vals[0] += res_b * res_b;
          const double db = div_coh_vals<UNIT>(un_b, un_a, Vc.y, Vm.y, true, hs, dx, dy);
          vals[1] += db * db;
        }
        if (POST) {
          const bool botr = jg == 0, topr = jg == s.nyg - 1;
          const double du = uf_a - UO.x, dvv = Vc.x - VO.x;
          vals[2] = fmax(vals[2], fabs(uf_a));
          vals[3] = fmax(vals[3], fabs(Vc.x));
          vals[4] += du * du;
          vals[5] += dvv * dvv;
          const double dc = div_central_vals<true>(uf_a, uf_w, uf_b, Vc.x, Vm.x, Vp.x, c.c0 == 0, c.c0 == nx - 1, botr,
                                                   topr, 1.0, 1.0);
          vals[6] += dc * dc;
          vals[7] += ((isfinite(uf_a) && isfinite(Vc.x) && isfinite(pa)) ? 0.0 : 1.0);
          vals[8] += (uf_a != uf_a) ? 1.0 : 0.0;
          vals[9] += (Vc.x != Vc.x) ? 1.0 : 0.0;
          if (c.c0 == nx / 2) center[j] = uf_a;
          if (c.vb) {
            const double du2 = uf_b - UO.y, dv2 = Vc.y - VO.y;
            vals[2] = fmax(vals[2], fabs(uf_b));
            vals[3] = fmax(vals[3], fabs(Vc.y));
            vals[4] += du2 * du2;
            vals[5] += dv2 * dv2;
            const double dc2 = div_central_vals<true>(uf_b, uf_a, uf_e, Vc.y, Vm.y, Vp.y, false, c.c0 + 1 == nx - 1,
                                                      botr, topr, 1.0, 1.0);
            vals[6] += dc2 * dc2;
            vals[7] += ((isfinite(uf_b) && isfinite(Vc.y) && isfinite(pb)) ? 0.0 : 1.0);
            vals[8] += (uf_b != uf_b) ? 1.0 : 0.0;
            vals[9] += (Vc.y != Vc.y) ? 1.0 : 0.0;
            if (c.c0 + 1 == nx / 2) center[j] = uf_b;
          }
        }
        st_pair(uout, off, c, uf_a, uf_b);       // own && va: the load column IS the lane's column
        st_pair(vout, off, c, Vc.x, Vc.y);
        if (shift) st_pair(x, off, c, pa, pb);
      }
      Xm = Xc; Xc = Xp; Xp = Xpp;
      Vm = Vc; Vc = Vp;
      if (F_PREFETCH) cur = nxt;
    }
  }
  if constexpr (POST) {
    const int slots[NV] = {S_RES_SUMSQ, S_DIVA_SUMSQ, S_UMAX, S_VMAX, S_RU_SUMSQ, S_RV_SUMSQ, S_DIVC_SUMSQ, S_NONFINITE,
                           S_UNAN, S_VNAN};
    grid_reduce_store<NV, 0xCu>(vals, red, slots);       // bits 2, 3: max|u|, max|v|
  } else {
    const int slots[NV] = {S_RES_SUMSQ, S_DIVA_SUMSQ};
    grid_reduce_store<NV, 0u>(vals, red, slots);
  }
}

}  // namespace pf
