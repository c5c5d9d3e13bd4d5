// pf_cg_dyn.cuh -- the production CG kernel: persistent, cooperative, dynamically
// scheduled, single- or multi-GPU (see pf_cg.cuh for the algorithm, the traffic
// accounting and the row/strip machinery it shares).
//
// What this adds over k_cg_persistent / k_cg_persistent_multi (kept as
// cross-checks, selectable with pf_set_option "cg_kernel"):
//   * the slab is cut into ~4 tiles per CTA (column chunk x ~28-row band) and
//     CTAs DRAW tiles from an atomic ticket counter -- ascending in phase A,
//     descending in phase B, so the sweep order is a grid-wide LIFO (the tiles
//     phase A finished last are the first phase B touches: L2 reuse) and the
//     tail of a phase is one small tile instead of a 110-row band.  The ncu /
//     globaltimer timeline of the static version showed 10 us (A) and 29 us (B)
//     between the first and the last CTA finishing a ~70 / ~110 us phase;
//   * partial sums are stored PER TILE and folded in tile order, so the result
//     does not depend on which CTA processed which tile: still bit-reproducible
//     run to run although the schedule is dynamic;
//   * tiles that do not touch the slab edge run mask-free from the first to the
//     last row (HEAD / LEAN / TAIL batches differ only in compile-time
//     store / centre-row control); only slab-edge tiles evaluate the Neumann,
//     ring, pin and peer-halo logic (GEN batches);
//   * one barrier implementation for 1..8 GPUs: last CTA to arrive folds the
//     tile partials, (R > 1) trades them with the other ranks through NVLink
//     mailboxes, and releases the local CTAs; one fence per CTA and side.
#pragma once

#include "pf_cg.cuh"

namespace pf {

constexpr int CG_MAX_TILES = 8192;
static_assert(PF_TRACE_OFFSET == 2 * CG_MAX_TILES, "trace region sits behind the tile partials");
static_assert(PF_PART_DOUBLES >= 3 * CG_MAX_GRID, "legacy kernels use 3 x CG_MAX_GRID partials");

enum BatchMode { B_GEN = 0, B_HEAD = 1, B_LEAN = 2, B_TAIL = 3 };

// ---- phase A batch ---------------------------------------------------------------
template <bool UNIT, int XE, int MODE, int DIR>
__device__ __forceinline__ void dynA_batch(const CgArgs& a, const Cols& c, const March& m, const double* r,
                                           const double* pold, double* pnew, double* peer_lo, double* peer_hi,
                                           bool first, double beta, uint64_t pol, int t0, int n, Row& prev, Row& cur,
                                           double& acc) {
  const Slab& s = a.s;
  double2 rr[CG_U], pp[CG_U];
  double rh[CG_U], ph[CG_U];
#pragma unroll
  for (int u = 0; u < CG_U; ++u) {
    const int jl = m.jl0 + DIR * (t0 + u);
    const int jg = s.row0 + jl;
    rr[u] = make_double2(0.0, 0.0);
    pp[u] = make_double2(0.0, 0.0);
    rh[u] = 0.0;
    ph[u] = 0.0;
    bool ok = true;
    if (MODE == B_GEN) ok = (u < n) && jg >= 0 && jg < s.nyg;
    if (MODE == B_TAIL) ok = (u < n);
    if (ok) {
      if (c.va) {
        const size_t o = cell(s, jl, c.ia);
        rr[u] = ld2(r + o);
        if (!first) pp[u] = ld2_hint(pold + o, pol);
      }
      if (c.vh) {
        const size_t oh = cell(s, jl, c.ih);
        rh[u] = ld1(r + oh);
        if (!first) ph[u] = ld1(pold + oh);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CG_U; ++u) {
    const bool live = (MODE == B_GEN || MODE == B_TAIL) ? (u < n) : true;   // warp-uniform
    if (live) {
      const int t = t0 + u;
      const int jl = m.jl0 + DIR * t;
      const int jg = s.row0 + jl;
      double da = a.invd, db = a.invd, dh = a.invd;
      if (MODE == B_GEN || XE != X_IN) {
        const bool ring_row = (MODE == B_GEN) && ((jg == 0) || (jg == s.nyg - 1));
        da = (ring_row || c.ring_a) ? 1.0 : a.invd;
        db = (ring_row || c.ring_b) ? 1.0 : a.invd;
        dh = (ring_row || c.ring_h) ? 1.0 : a.invd;
      }
      const double za = da * rr[u].x, zb = db * rr[u].y, zh = dh * rh[u];
      Row nw;
      if (first) {
        nw.a = za; nw.b = zb; nw.h = zh;
      } else {
        nw.a = (pp[u].x * beta) + za;
        nw.b = (pp[u].y * beta) + zb;
        nw.h = (ph[u] * beta) + zh;
      }
      bool do_store = true, do_center = true;
      if (MODE == B_GEN) { do_store = (t >= 1 && t <= m.TH); do_center = (t >= 2); }
      if (MODE == B_HEAD) { do_store = (u >= 1); do_center = (u >= 2); }       // compile-time after unrolling
      if (MODE == B_TAIL) { do_store = (t <= m.TH); }
      if (do_store) {
        double* dst = pnew + cell(s, jl, c.ia);
        if (c.vb) st2(dst, nw.a, nw.b);
        else if (c.va) st1(dst, nw.a);
        if (MODE == B_GEN) peer_row_store(s, c, jl, peer_lo, peer_hi, nw.a, nw.b);
      }
      if (do_center) {
        const int jgc = jg - DIR;
        const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
        double la, lb;
        if (DIR > 0) lap_pair<UNIT, XE, MODE == B_GEN>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, MODE == B_GEN>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (MODE == B_GEN && jgc == 0 && c.ia == 0) la = -cur.a;      // pinned row: q[0] = p[0]
        if (XE == X_PART) {
          if (c.va) acc += cur.a * la;
          if (c.vb) acc += cur.b * lb;
        } else {
          acc += cur.a * la;
          acc += cur.b * lb;
        }
      }
      prev = cur;
      cur = nw;
    }
  }
}

// ---- phase B batch ---------------------------------------------------------------
template <bool UNIT, int XE, int MODE, int DIR>
__device__ __forceinline__ void dynB_batch(const CgArgs& a, const Cols& c, const March& m, const double* p, double* x,
                                           double* r, double* peer_lo, double* peer_hi, double alpha, uint64_t pol,
                                           int t0, int n, Row& prev, Row& cur, double& acc_rr, double& acc_rz) {
  const Slab& s = a.s;
  double2 pr[CG_U], xx[CG_U], rr[CG_U];
  double ph[CG_U];
#pragma unroll
  for (int u = 0; u < CG_U; ++u) {
    const int t = t0 + u;
    const int jl = m.jl0 + DIR * t;
    const int jg = s.row0 + jl;
    pr[u] = make_double2(0.0, 0.0);
    xx[u] = make_double2(0.0, 0.0);
    rr[u] = make_double2(0.0, 0.0);
    ph[u] = 0.0;
    const bool live = (MODE == B_GEN || MODE == B_TAIL) ? (u < n) : true;
    if (live) {
      const bool exists = (MODE == B_GEN) ? (jg >= 0 && jg < s.nyg) : true;
      if (exists) {
        if (c.va) pr[u] = ld2(p + cell(s, jl, c.ia));
        if (c.vh) ph[u] = ld1(p + cell(s, jl, c.ih));
      }
      bool do_center = true;
      if (MODE == B_GEN) do_center = (t >= 2);
      if (MODE == B_HEAD) do_center = (u >= 2);
      if (do_center && c.va) {
        const size_t o = cell(s, jl - DIR, c.ia);
        xx[u] = ld2_hint_nol1(x + o, pol);
        rr[u] = ld2(r + o);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CG_U; ++u) {
    const bool live = (MODE == B_GEN || MODE == B_TAIL) ? (u < n) : true;
    if (live) {
      const int t = t0 + u;
      const int jl = m.jl0 + DIR * t;
      const int jg = s.row0 + jl;
      Row nw = {pr[u].x, pr[u].y, ph[u]};
      bool do_center = true;
      if (MODE == B_GEN) do_center = (t >= 2);
      if (MODE == B_HEAD) do_center = (u >= 2);
      if (do_center) {
        const int jc = jl - DIR;
        const int jgc = jg - DIR;
        const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
        double la, lb;
        if (DIR > 0) lap_pair<UNIT, XE, MODE == B_GEN>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, MODE == B_GEN>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (MODE == B_GEN && jgc == 0 && c.ia == 0) la = -cur.a;
        const double xa = xx[u].x + alpha * cur.a;
        const double xb = xx[u].y + alpha * cur.b;
        const double ra = rr[u].x + alpha * la;
        const double rb = rr[u].y + alpha * lb;
        const size_t o = cell(s, jc, c.ia);
        if (c.vb) {
          st2_hint(x + o, xa, xb, pol);
          st2(r + o, ra, rb);
        } else if (c.va) {
          st1(x + o, xa);
          st1(r + o, ra);
        }
        if (MODE == B_GEN) peer_row_store(s, c, jc, peer_lo, peer_hi, ra, rb);
        double da = a.invd, db = a.invd;
        if (MODE == B_GEN || XE != X_IN) {
          const bool ring_row = (MODE == B_GEN) && ((jgc == 0) || (jgc == s.nyg - 1));
          da = (ring_row || c.ring_a) ? 1.0 : a.invd;
          db = (ring_row || c.ring_b) ? 1.0 : a.invd;
        }
        if (XE != X_PART || c.va) {
          acc_rr += ra * ra;
          acc_rz += ra * (da * ra);
        }
        if (XE != X_PART || c.vb) {
          acc_rr += rb * rb;
          acc_rz += rb * (db * rb);
        }
      }
      prev = cur;
      cur = nw;
    }
  }
}

// ---- strip drivers -----------------------------------------------------------------
// YE: the tile touches the slab edge (global boundary rows, Jacobi ring rows, the
// pinned cell, peer halo rows live there): everything through GEN batches.
template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ double dynA_strip(const CgArgs& a, const double* r, const double* pold, double* pnew,
                                             double* peer_lo, double* peer_hi, bool first, double beta,
                                             const WarpTile& w) {
  const Cols c = strip_cols<XE>(a.s, w.i0);
  const March m = make_march<DIR>(w);
  const uint64_t pol = policy_evict_first();
  const int nrows = m.TH + 2;
  const bool ye = (w.j0 == 0) || (w.j1 == a.s.nyl);
  double acc = 0.0;
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  if (ye || nrows < 2 * CG_U) {
    for (int t = 0; t < nrows; t += CG_U) {
      const int n = (nrows - t) < CG_U ? (nrows - t) : CG_U;
      dynA_batch<UNIT, XE, B_GEN, DIR>(a, c, m, r, pold, pnew, peer_lo, peer_hi, first, beta, pol, t, n, prev, cur, acc);
    }
    return acc;
  }
  dynA_batch<UNIT, XE, B_HEAD, DIR>(a, c, m, r, pold, pnew, nullptr, nullptr, first, beta, pol, 0, CG_U, prev, cur, acc);
  int t = CG_U;
  for (; t + CG_U <= m.TH + 1; t += CG_U)       // every row of the batch is a band row (t <= TH)
    dynA_batch<UNIT, XE, B_LEAN, DIR>(a, c, m, r, pold, pnew, nullptr, nullptr, first, beta, pol, t, CG_U, prev, cur, acc);
  dynA_batch<UNIT, XE, B_TAIL, DIR>(a, c, m, r, pold, pnew, nullptr, nullptr, first, beta, pol, t, nrows - t, prev, cur, acc);
  return acc;
}

template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ void dynB_strip(const CgArgs& a, const double* p, double* x, double* r, double* peer_lo,
                                           double* peer_hi, double alpha, const WarpTile& w, double& acc_rr,
                                           double& acc_rz) {
  const Cols c = strip_cols<XE>(a.s, w.i0);
  const March m = make_march<DIR>(w);
  const uint64_t pol = policy_evict_first();
  const int nrows = m.TH + 2;
  const bool ye = (w.j0 == 0) || (w.j1 == a.s.nyl);
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  if (ye || nrows < 2 * CG_U) {
    for (int t = 0; t < nrows; t += CG_U) {
      const int n = (nrows - t) < CG_U ? (nrows - t) : CG_U;
      dynB_batch<UNIT, XE, B_GEN, DIR>(a, c, m, p, x, r, peer_lo, peer_hi, alpha, pol, t, n, prev, cur, acc_rr, acc_rz);
    }
    return;
  }
  dynB_batch<UNIT, XE, B_HEAD, DIR>(a, c, m, p, x, r, nullptr, nullptr, alpha, pol, 0, CG_U, prev, cur, acc_rr, acc_rz);
  int t = CG_U;
  for (; t + CG_U <= m.TH + 2; t += CG_U)       // phase B stores the CENTRE row (t-1 <= TH): one more lean row than A
    dynB_batch<UNIT, XE, B_LEAN, DIR>(a, c, m, p, x, r, nullptr, nullptr, alpha, pol, t, CG_U, prev, cur, acc_rr, acc_rz);
  if (t < nrows)
    dynB_batch<UNIT, XE, B_TAIL, DIR>(a, c, m, p, x, r, nullptr, nullptr, alpha, pol, t, nrows - t, prev, cur, acc_rr, acc_rz);
}

// ---- barrier + all-reduce over tile partials ---------------------------------------
// Ends a phase: every CTA arrives; the last one folds the per-TILE partials in tile
// order, trades the sums with the other ranks (R > 1) and releases everybody.
template <int NV>
__device__ __forceinline__ bool dyn_exchange(const CgArgs& a, unsigned long long e, double (&out)[NV], double* sm,
                                             unsigned long long e0) {
  unsigned long long* tr = trace_slot(a, e0, e);
  __shared__ int s_last;
  __shared__ double s_out[NV + 1];
  const CgSync& y = a.y;
  const int G = gridDim.x, tid = threadIdx.x, R = y.nranks;
  const int par = (int)(e & 1ULL);
  __syncthreads();                   // every thread's stores happen-before thread 0's fence
  if (tid == 0) {
    if (R > 1) __threadfence_system(); else __threadfence();
    if (tr) {
      const unsigned long long t = gtime();
      atomicMin(tr + 0, t);
      atomicMax(tr + 1, t);
      const unsigned long long kb = e - e0 - 401;
      if (kb < 2) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        ((unsigned long long*)(a.part + 4096))[kb * 512 + blockIdx.x] = (t << 10) | (unsigned long long)(smid & 1023u);
      }
    }
    s_last = (atomicAdd(y.arrive, 1ULL) == e * (unsigned long long)G - 1ULL) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    if (tid == 0) __threadfence();
    __syncthreads();
    double loc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) loc[k] = 0.0;
    for (int q = tid; q < a.ntiles; q += CG_THREADS) {
#pragma unroll
      for (int k = 0; k < NV; ++k) loc[k] += ld_cg(a.part + k * CG_MAX_TILES + q);
    }
    block_sum_n<NV>(loc, sm);
    if (tid == 0 && tr) tr[2] = gtime();
    if (R > 1) {
      if (tid == 0) {
        for (int q = 0; q < R; ++q)
          for (int k = 0; k < NV; ++k) st_sys_f64(y.peer_mbox[q] + (par * CG_MAX_RANKS + y.rank) * 4 + k, loc[k]);
        for (int q = 0; q < R; ++q) st_release_sys(y.peer_flags[q] + y.rank, e);   // release orders the mailbox stores
        if (tr) tr[3] = gtime();
      }
      if (tid < R) {
        const long long t0 = clock64();
        while (ld_relaxed_sys_u64(y.flags + tid) < e) {
          if (clock64() - t0 > CG_SPIN_LIMIT) {
            atomicExch(y.err, 1ULL);
            break;
          }
        }
        __threadfence_system();      // acquire: the peers' mailbox and halo-row stores are visible now
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (tr) tr[4] = gtime();
      for (int k = 0; k < NV; ++k) {
        double sum = loc[k];
        if (R > 1) {                 // rank order, identical on every GPU
          sum = 0.0;
          for (int q = 0; q < R; ++q) sum += ld_sys_f64(y.mbox + (par * CG_MAX_RANKS + q) * 4 + k);
        }
        y.gsum[par * 4 + k] = sum;
      }
      __threadfence();
      st_release_gpu(y.release, e);
      if (tr) tr[5] = gtime();
    }
  }
  if (tid == 0) {
    const long long t0 = clock64();
    while (ld_relaxed_u64(y.release) < e) {
      if (clock64() - t0 > CG_SPIN_LIMIT) {
        atomicExch(y.err, 1ULL);
        break;
      }
    }
    __threadfence();                 // acquire side (also drops stale L1 lines of this SM)
#pragma unroll
    for (int k = 0; k < NV; ++k) s_out[k] = ld_cg(y.gsum + par * 4 + k);
    s_out[NV] = (ld_relaxed_u64(y.err) == 0ULL) ? 1.0 : 0.0;
    if (tr && blockIdx.x == 0) tr[6] = gtime();
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) out[k] = s_out[k];
  return s_out[NV] != 0.0;
}

// Tile tickets: a monotonic counter, phase n owns tickets [n*(ntiles+G), (n+1)*(ntiles+G));
// every CTA draws until it gets a ticket past the last tile, so exactly ntiles+G are drawn.
struct TileFeed {
  unsigned long long base;
  int ntiles;
};
__device__ __forceinline__ int draw_tile(const CgArgs& a, const TileFeed& f) {
  const unsigned long long tk = atomicAdd(a.y.tickets, 1ULL) - f.base;
  return tk < (unsigned long long)f.ntiles ? (int)tk : -1;
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS, 2) k_cg_dyn(CgArgs a) {
  __shared__ double sm[3 * 32];
  __shared__ int s_tile[2];
  const int G = gridDim.x;
  const int tid = threadIdx.x;
  const CgSync& y = a.y;

  const double bb = ld_cg(a.scal + S_BB);            // all-reduced by the host sequence when R > 1
  const double bnorm = sqrt(bb);
  const double tol0 = a.tol * bnorm;
  const double atol = (tol0 > 0.0) ? tol0 : 0.0;     // max(0.0, rtol*||b||) with Python's NaN behaviour
  int iters = 0;
  int info = a.maxiter;
  double rr = bb;
  double rho = ld_cg(a.scal + S_BZ);
  double rho_prev = 0.0;
  double* pcur = a.p0;
  double* pnext = a.p1;
  unsigned long long e = *((volatile unsigned long long*)y.epoch);   // barrier epochs used by earlier launches
  const unsigned long long e0 = e;
  const unsigned long long tbase = *((volatile unsigned long long*)y.ticket_base);
  const unsigned long long per_phase = (unsigned long long)a.ntiles + (unsigned long long)G;
  bool ok = true;

  if (bnorm == 0.0) {                                 // scipy: "return b, 0"; x was zeroed by k_rhs2
    info = 0;
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) {                          // top-of-loop test on the recursive residual
        info = 0;
        break;
      }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      // ---------------- phase A: tiles ascending, rows top -> bottom
      {
        const TileFeed feed = {tbase + (e - e0) * per_phase, a.ntiles};
        double* plo = (pnext == a.p0) ? y.below_p0 : y.below_p1;
        double* phi = (pnext == a.p0) ? y.above_p0 : y.above_p1;
        if (tid == 0) s_tile[0] = draw_tile(a, feed);
        __syncthreads();
        int cur = s_tile[0], slot = 0;
        while (cur >= 0) {
          if (tid == 0) s_tile[slot ^ 1] = draw_tile(a, feed);     // next ticket: latency hidden by this tile
          const WarpTile w = warp_tile(a, cur);
          double v[1] = {0.0};
          if (w.active) {
            const int xk = strip_kind(a.s, w.i0);
            if (xk == X_PART) v[0] = dynA_strip<UNIT, X_PART, +1>(a, a.r, pcur, pnext, plo, phi, first, beta, w);
            else if (xk == X_WALL) v[0] = dynA_strip<UNIT, X_WALL, +1>(a, a.r, pcur, pnext, plo, phi, first, beta, w);
            else v[0] = dynA_strip<UNIT, X_IN, +1>(a, a.r, pcur, pnext, plo, phi, first, beta, w);
          }
          block_sum_n<1>(v, sm);
          if (tid == 0) a.part[cur] = -v[0];                       // p.q = -(sum p*lap)
          slot ^= 1;
          cur = s_tile[slot];                                       // written before the barriers of block_sum_n
        }
      }
      double oa[1];
      ok = dyn_exchange<1>(a, ++e, oa, sm, e0);
      if (!ok) break;
      const double alpha = rho / oa[0];
      // ---------------- phase B: tiles descending, rows bottom -> top
      {
        const TileFeed feed = {tbase + (e - e0) * per_phase, a.ntiles};
        if (tid == 0) s_tile[0] = draw_tile(a, feed);
        __syncthreads();
        int cur = s_tile[0], slot = 0;
        while (cur >= 0) {
          if (tid == 0) s_tile[slot ^ 1] = draw_tile(a, feed);
          const int tile = a.ntiles - 1 - cur;
          const WarpTile w = warp_tile(a, tile);
          double v[2] = {0.0, 0.0};
          if (w.active) {
            const int xk = strip_kind(a.s, w.i0);
            if (xk == X_PART) dynB_strip<UNIT, X_PART, -1>(a, pnext, a.x, a.r, y.below_r, y.above_r, alpha, w, v[0], v[1]);
            else if (xk == X_WALL) dynB_strip<UNIT, X_WALL, -1>(a, pnext, a.x, a.r, y.below_r, y.above_r, alpha, w, v[0], v[1]);
            else dynB_strip<UNIT, X_IN, -1>(a, pnext, a.x, a.r, y.below_r, y.above_r, alpha, w, v[0], v[1]);
          }
          block_sum_n<2>(v, sm);
          if (tid == 0) {
            a.part[tile] = v[0];
            a.part[CG_MAX_TILES + tile] = v[1];
          }
          slot ^= 1;
          cur = s_tile[slot];
        }
      }
      double ob[2];
      ok = dyn_exchange<2>(a, ++e, ob, sm, e0);
      if (!ok) break;
      rr = ob[0];
      rho_prev = rho;
      rho = ob[1];
      double* tmp = pcur; pcur = pnext; pnext = tmp;
      ++iters;
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.scal[S_INFO] = ok ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    *y.epoch = e;                    // every rank executed the same number of phases
    *y.ticket_base = tbase + (e - e0) * per_phase;
  }
}

}  // namespace pf
