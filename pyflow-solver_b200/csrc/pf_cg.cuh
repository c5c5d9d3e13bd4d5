// pf_cg.cuh -- matrix-free Jacobi-preconditioned Conjugate Gradient on the
// pinned Neumann operator, as ONE persistent cooperative kernel.
//
// What it replaces: scipy.sparse.linalg.cg as called from
// linear_solvers/interface.py:43 with A = _NegLaplacianPinned
// (pressure_solver.py:97-105) and M = Jacobi of the ring-identity matrix
// (pressure_solver.py:15-40, preconditioners.py:8-25).  The recurrence,
// stopping rule and operation order are scipy's; the
// classical (non-pipelined) form is kept so results match.
//
// Traffic per cell per iteration: 8 doubles = 64 B (SURVEY 8d contract):
//   phase A: read r, p_old  -> write p_new = (p_old*beta) + M r ; q = A p_new is
//            formed on the fly from a 3-row register window and only p.q kept
//   phase B: read x, r, p   -> recompute q = A p ; x += alpha p ; r -= alpha q ;
//            accumulate r.r and r.(M r)
// Two grid-wide reductions per iteration (p.q, then {r.r, r.z}): per-CTA
// partials, grid barrier, every CTA folds the partials in the same fixed
// order => all CTAs hold bit-identical scalars and take the same branch.
//
// Work decomposition: the slab is cut into column chunks of 64*wx columns and
// row bands; a CTA tile is chunk x band, its 8 warps are arranged wx x wy.  A
// warp owns a 64-column strip (2 columns per lane, 16-byte vector accesses)
// and marches along the rows of its sub-band keeping rows j-1, j, j+1 in
// registers; left/right neighbours come from warp shuffles, strip-edge columns
// from one extra scalar load on lanes 0 / 31 (an L1 hit: the neighbouring warp
// of the same CTA streams that line).  Each value is fetched from HBM once per
// phase.  Phase A marches top->bottom and phase B bottom->top so the rows one
// phase touched last are the first the next one reads (L2 reuse, 126 MB);
// x and the old search direction, which nobody re-reads soon, are streamed
// with evict-first hints so they do not displace r and p in L2.
//
// Rows are processed in batches (CG_UA / CG_UB rows): all loads of a batch are issued, then
// consumed.  Rows strictly inside a band (the vast majority) take a LEAN path
// with no boundary masks; only the first/last rows of a band (which may touch
// the domain boundary, the Jacobi ring or the pinned cell) and strips on the
// left/right walls (template XE) evaluate the Neumann / ring predicates.
//
// Sign bookkeeping that keeps results bit-identical while saving negations:
// q = -lap exactly, so r - alpha*q == r + alpha*lap and sum p*q == -(sum p*lap)
// (IEEE negation is exact and round-to-nearest is sign-symmetric).
#pragma once

#include <cooperative_groups.h>

#include "pf_common.cuh"

namespace pf {
namespace cg = cooperative_groups;

constexpr int CG_THREADS = 256;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int STRIP = 64;     // columns per warp
#ifndef PF_CG_UA
#define PF_CG_UA 6
#endif
#ifndef PF_CG_UB
#define PF_CG_UB 4
#endif
#ifndef PF_CG_MINB
#define PF_CG_MINB 2
#endif
constexpr int CG_UA = PF_CG_UA;   // rows fetched per batch in phase A (loads in flight per lane: 2 arrays)
constexpr int CG_UB = PF_CG_UB;   // ... in phase B (3 arrays)
constexpr int CG_MAX_GRID = 2048;

constexpr int CG_MAX_RANKS = 8;
// Layout of the CG partial-sum scratch (doubles): [2][CG_MAX_TILES][2] per-tile partials of the TMA kernels as LL
// words (the other kernels use [3][CG_MAX_GRID] per-CTA partials in the same space), 64 doubles of barrier
// timeline, [2][CG_MAX_GROUPS] per-group partials and CG_MAX_GROUPS u32 group counters (hierarchical fold,
// pf_cg_tma.cuh).
constexpr int CG_MAX_TILES = 16384;         // per-tile partial sums of the dynamically scheduled kernel
constexpr int CG_TILE_GROUP = 64;           // tiles per fold group
constexpr int CG_MAX_GROUPS = CG_MAX_TILES / CG_TILE_GROUP;
constexpr int PF_TRACE_OFFSET = 4 * CG_MAX_TILES;
constexpr int PF_GPART_OFFSET = PF_TRACE_OFFSET + 64;
constexpr int PF_GCOUNT_OFFSET = PF_GPART_OFFSET + 2 * CG_MAX_GROUPS;
constexpr int PF_PART_DOUBLES = PF_GCOUNT_OFFSET + CG_MAX_GROUPS / 2;

// Cross-GPU exchange state of the persistent kernel (multi-GPU slabs).  The
// `peer_*` pointers are the SAME buffers of the other ranks, mapped into this
// process through symmetric memory (NVLink / NVSwitch peer access).
struct CgSync {
  unsigned long long* arrive;    // local: monotonic count of CTA arrivals
  unsigned long long* release;   // local: last epoch released to the local CTAs
  unsigned long long* flags;     // local, written by peers: epoch published by rank q
  double* mbox;                  // local, written by peers: [2][CG_MAX_RANKS] 64-byte LL slots of partial sums
  double* gsum;                  // local: one 64-byte LL slot, sums over all ranks for the local CTAs
  unsigned long long* epoch;     // local: epoch counter persisted between launches
  unsigned long long* err;       // local: != 0 after a spin-wait timeout
  unsigned long long* tickets;   // local: monotonic tile-ticket counter (k_cg_tmad)
  unsigned long long* ticket_base;  // local: first ticket of the next launch (persisted)
  unsigned long long* peer_flags[CG_MAX_RANKS];
  double* peer_mbox[CG_MAX_RANKS];
  // neighbour halo rows (null when there is no neighbour): my bottom row goes to
  // the upper halo of the slab below, my top row to the lower halo of the slab above
  double* below_r; double* below_p0; double* below_p1;
  double* above_r; double* above_p0; double* above_p1;
  int rank, nranks;
  int opt;                       // experiment bits (pf_set_option "cg_sync_opt")
};

struct CgArgs {
  Slab s;
  const double* b;
  double* x;
  double* r;
  double* p0;
  double* p1;
  double* part;       // [3][CG_MAX_GRID] per-CTA partials: p.q, r.r, r.z
  double* scal;       // device scalar block
  double dx, dy;
  double invd;        // interior inverse diagonal (1.0 when M = I)
  double tol;
  int maxiter;
  int wx, wy;         // warp arrangement inside a CTA tile, wx*wy == CG_WARPS
  int nchunks, nbands, ntiles;
  int tile_rows;      // rows per (big) tile of the dynamically scheduled TMA kernel
  int small_rows;     // rows per small tile: nb_small bands of them at EACH end of the slab (0: uniform tiles), so that
  int nb_small;       // the last tickets of both sweeps are short and the CTAs finish a phase close together
  CgSync y;           // only used by the multi-GPU persistent kernel
};

struct Row {
  double a, b;        // this lane's two columns
  double h;           // strip halo column (lane 0: left, lane 31: right)
};

struct WarpTile {
  int i0, j0, j1;
  bool active;
};

__device__ __forceinline__ WarpTile warp_tile(const CgArgs& a, int t) {
  const int warp = threadIdx.x >> 5;
  const int wxi = warp % a.wx, wyi = warp / a.wx;
  const int c = t % a.nchunks, bd = t / a.nchunks;
  WarpTile w;
  w.i0 = (c * a.wx + wxi) * STRIP;
  const int bj0 = (int)((long long)bd * a.s.nyl / a.nbands);
  const int bj1 = (int)((long long)(bd + 1) * a.s.nyl / a.nbands);
  const int th = bj1 - bj0;
  w.j0 = bj0 + (int)((long long)wyi * th / a.wy);
  w.j1 = bj0 + (int)((long long)(wyi + 1) * th / a.wy);
  w.active = (w.i0 < a.s.nx) && (w.j1 > w.j0);
  return w;
}

// Per-lane column geometry of a strip.
struct Cols {
  int ia, ih;
  bool va, vb, vh;          // column exists
  bool hw_a, he_a, he_b;    // has west / east neighbour inside the domain
  bool ring_a, ring_b, ring_h;
};

// Strip kinds (template XE): 0 = strictly inside the domain in x (no column
// predicates at all); 1 = WALL strip whose 64 columns all exist (only lane 0's
// west face / lane 31's east face and the two ring columns are special: two
// selects per row); 2 = partially outside the domain (every predicate).
constexpr int X_IN = 0, X_WALL = 1, X_PART = 2;

template <int XE>
__device__ __forceinline__ Cols strip_cols(const Slab& s, int i0) {
  const int lane = threadIdx.x & 31;
  Cols c;
  c.ia = i0 + 2 * lane;
  c.ih = (lane == 0) ? i0 - 1 : i0 + STRIP;
  if (XE == X_WALL) {         // all 64 columns exist; nx >= 64 so a lane's two columns are never both walls
    const int ib = c.ia + 1;
    c.va = c.vb = true;
    c.vh = (lane == 0 || lane == 31) && c.ih >= 0 && c.ih < s.nx;
    c.hw_a = c.ia > 0;
    c.he_a = true;
    c.he_b = ib < s.nx - 1;
    c.ring_a = (c.ia == 0);
    c.ring_b = (ib == s.nx - 1);
    c.ring_h = (c.ih == 0) || (c.ih == s.nx - 1);
  } else if (XE == X_PART) {
    const int ib = c.ia + 1;
    c.va = c.ia < s.nx;
    c.vb = ib < s.nx;
    c.vh = (lane == 0 || lane == 31) && c.ih >= 0 && c.ih < s.nx;
    c.hw_a = c.ia > 0;
    c.he_a = c.ia < s.nx - 1;
    c.he_b = ib < s.nx - 1;
    c.ring_a = (c.ia == 0) || (c.ia == s.nx - 1);
    c.ring_b = (ib == 0) || (ib == s.nx - 1);
    c.ring_h = (c.ih == 0) || (c.ih == s.nx - 1);
  } else {                    // strip strictly inside the domain in x
    c.va = c.vb = true;
    c.vh = (lane == 0 || lane == 31);
    c.hw_a = c.he_a = c.he_b = true;
    c.ring_a = c.ring_b = c.ring_h = false;
  }
  return c;
}

__device__ __forceinline__ int strip_kind(const Slab& s, int i0) {
  if (i0 + STRIP > s.nx) return X_PART;                 // some of the 64 columns do not exist
  // wall strips -- and the strip whose east HALO column is the wall (nx % 64 == 1): its halo
  // value needs the ring-column Jacobi factor, which only the WALL path evaluates
  if (i0 == 0 || i0 + STRIP >= s.nx - 1) return X_WALL;
  return X_IN;
}

// laplacian_coherent (NOT negated) at the two columns of the centre row.
// GEN: evaluate the Neumann predicates of this row / strip; otherwise every
// neighbour exists.
template <bool UNIT, int XE, bool GEN>
__device__ __forceinline__ void lap_pair(const CgArgs& a, const Cols& c, const Row& south, const Row& cur,
                                         const Row& north, bool hs, bool hn, double& la, double& lb) {
  const int lane = threadIdx.x & 31;
  double w_a = __shfl_up_sync(0xffffffffu, cur.b, 1);
  double e_b = __shfl_down_sync(0xffffffffu, cur.a, 1);
  if (lane == 0) w_a = cur.h;
  if (lane == 31) e_b = cur.h;
  if (XE == X_PART || (GEN && XE != X_IN)) {
    la = -neg_lap<UNIT>(cur.a, w_a, cur.b, south.a, north.a, c.hw_a, c.he_a, hs, hn, a.dx, a.dy);
    lb = -neg_lap<UNIT>(cur.b, cur.a, e_b, south.b, north.b, true, c.he_b, hs, hn, a.dx, a.dy);
  } else if (GEN) {           // interior strip, boundary row: only the row predicates
    la = -neg_lap<UNIT>(cur.a, w_a, cur.b, south.a, north.a, true, true, hs, hn, a.dx, a.dy);
    lb = -neg_lap<UNIT>(cur.b, cur.a, e_b, south.b, north.b, true, true, hs, hn, a.dx, a.dy);
  } else if (XE == X_WALL) {  // wall strip, interior row: only lane 0's west / lane 31's east face
    la = -neg_lap<UNIT>(cur.a, w_a, cur.b, south.a, north.a, c.hw_a, true, true, true, a.dx, a.dy);
    lb = -neg_lap<UNIT>(cur.b, cur.a, e_b, south.b, north.b, true, c.he_b, true, true, a.dx, a.dy);
  } else {
    la = -neg_lap<UNIT>(cur.a, w_a, cur.b, south.a, north.a, true, true, true, true, a.dx, a.dy);
    lb = -neg_lap<UNIT>(cur.b, cur.a, e_b, south.b, north.b, true, true, true, true, a.dx, a.dy);
  }
}

// Visiting order of a warp tile: index t = 0 is the halo row in front of the
// band, t = 1..TH the band rows, t = TH+1 the halo row behind it.
struct March {
  int jl0;      // local row of t = 0
  int dj;       // +1 / -1
  int TH;
};
template <int DIR>
__device__ __forceinline__ March make_march(const WarpTile& w) {
  March m;
  m.TH = w.j1 - w.j0;
  m.dj = DIR;
  m.jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  return m;
}

// ---- phase A -------------------------------------------------------------------
// One batch of n (<= CG_UA) consecutive visits starting at t0.  GEN rows evaluate
// existence / ring / pin predicates; lean rows are strictly inside the band.
// Mirror a freshly written slab-boundary row into the neighbour's halo row
// (peer memory over NVLink).  Only ever reached from GEN rows.
__device__ __forceinline__ void peer_row_store(const Slab& s, const Cols& c, int jl, double* below, double* above,
                                               double va, double vb) {
  double* dst = nullptr;
  if (jl == 0 && below) dst = below;                 // -> upper halo row of the slab below
  if (jl == s.nyl - 1 && above) {                    // -> lower halo row of the slab above
    if (dst) {                                       // one-row slab: feeds both neighbours
      if (c.vb) st2(dst + c.ia, va, vb);
      else if (c.va) st1(dst + c.ia, va);
    }
    dst = above;
  }
  if (dst) {
    if (c.vb) st2(dst + c.ia, va, vb);
    else if (c.va) st1(dst + c.ia, va);
  }
}

template <bool UNIT, int XE, bool GEN, int DIR>
__device__ __forceinline__ void phaseA_batch(const CgArgs& a, const Cols& c, const March& m, const double* r,
                                             const double* pold, double* pnew, double* peer_lo, double* peer_hi,
                                             bool first, double beta, int t0, int n,
                                             Row& prev, Row& cur, double& acc) {
  const Slab& s = a.s;
  const uint64_t pol = policy_evict_first();
  double2 rr[CG_UA], pp[CG_UA];
  double rh[CG_UA], ph[CG_UA];
#pragma unroll
  for (int u = 0; u < CG_UA; ++u) {
    const int jl = m.jl0 + DIR * (t0 + u);
    const int jg = s.row0 + jl;
    rr[u] = make_double2(0.0, 0.0);
    pp[u] = make_double2(0.0, 0.0);
    rh[u] = 0.0;
    ph[u] = 0.0;
    const bool ok = GEN ? (u < n && jg >= 0 && jg < s.nyg) : true;
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
  for (int u = 0; u < CG_UA; ++u) {
    if (!GEN || u < n) {               // warp-uniform
      const int t = t0 + u;
      const int jl = m.jl0 + DIR * t;
      const int jg = s.row0 + jl;
      double da = a.invd, db = a.invd, dh = a.invd;
      if (GEN || XE != X_IN) {
        const bool ring_row = GEN && ((jg == 0) || (jg == s.nyg - 1));
        da = (ring_row || c.ring_a) ? 1.0 : a.invd;
        db = (ring_row || c.ring_b) ? 1.0 : a.invd;
        dh = (ring_row || c.ring_h) ? 1.0 : a.invd;
      }
      const double za = da * rr[u].x, zb = db * rr[u].y, zh = dh * rh[u];
      Row nw;
      if (first) {                     // p = z
        nw.a = za; nw.b = zb; nw.h = zh;
      } else {                         // p *= beta ; p += z
        nw.a = (pp[u].x * beta) + za;
        nw.b = (pp[u].y * beta) + zb;
        nw.h = (ph[u] * beta) + zh;
      }
      if (!GEN || (t >= 1 && t <= m.TH)) {
        double* dst = pnew + cell(s, jl, c.ia);
        if (c.vb) st2(dst, nw.a, nw.b);
        else if (c.va) st1(dst, nw.a);
        if (GEN) peer_row_store(s, c, jl, peer_lo, peer_hi, nw.a, nw.b);
      }
      if (!GEN || t >= 2) {
        const int jgc = jg - DIR;
        const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
        double la, lb;
        if (DIR > 0) lap_pair<UNIT, XE, GEN>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, GEN>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (GEN && jgc == 0 && c.ia == 0) la = -cur.a;      // pinned row: q[0] = p[0]
        // acc holds sum p*lap = -(p.q)
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

template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ double phaseA_strip(const CgArgs& a, const double* r, const double* pold, double* pnew,
                                               double* peer_lo, double* peer_hi, bool first, double beta,
                                               const WarpTile& w) {
  const Cols c = strip_cols<XE>(a.s, w.i0);
  const March m = make_march<DIR>(w);
  const int nrows = m.TH + 2;
  double acc = 0.0;
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  // generic head: t = 0, 1, 2 (halo row, first band row, and the row that
  // completes the first centre); lean body: t = 3 .. TH-1; generic tail.
  const int head = nrows < 3 ? nrows : 3;
  phaseA_batch<UNIT, XE, true, DIR>(a, c, m, r, pold, pnew, peer_lo, peer_hi, first, beta, 0, head, prev, cur, acc);
  int t = head;
  const int lean_end = m.TH;                       // lean rows are t in [3, TH)
  for (; t + CG_UA <= lean_end; t += CG_UA)
    phaseA_batch<UNIT, XE, false, DIR>(a, c, m, r, pold, pnew, peer_lo, peer_hi, first, beta, t, CG_UA, prev, cur, acc);
  for (; t < nrows; t += CG_UA) {
    const int n = (nrows - t) < CG_UA ? (nrows - t) : CG_UA;
    phaseA_batch<UNIT, XE, true, DIR>(a, c, m, r, pold, pnew, peer_lo, peer_hi, first, beta, t, n, prev, cur, acc);
  }
  return acc;
}

// ---- phase B -------------------------------------------------------------------
template <bool UNIT, int XE, bool GEN, int DIR>
__device__ __forceinline__ void phaseB_batch(const CgArgs& a, const Cols& c, const March& m, const double* p,
                                             double* x, double* r, double* peer_lo, double* peer_hi, double alpha,
                                             int t0, int n, Row& prev, Row& cur, double& acc_rr, double& acc_rz) {
  const Slab& s = a.s;
  const uint64_t pol = policy_evict_first();
  double2 pr[CG_UB], xx[CG_UB], rr[CG_UB];
  double ph[CG_UB];
#pragma unroll
  for (int u = 0; u < CG_UB; ++u) {
    const int t = t0 + u;
    const int jl = m.jl0 + DIR * t;
    const int jg = s.row0 + jl;
    pr[u] = make_double2(0.0, 0.0);
    xx[u] = make_double2(0.0, 0.0);
    rr[u] = make_double2(0.0, 0.0);
    ph[u] = 0.0;
    if (!GEN || u < n) {
      if (!GEN || (jg >= 0 && jg < s.nyg)) {
        if (c.va) pr[u] = ld2(p + cell(s, jl, c.ia));
        if (c.vh) ph[u] = ld1(p + cell(s, jl, c.ih));
      }
      if ((!GEN || t >= 2) && c.va) {          // centre row of this visit: always a band row
        const size_t o = cell(s, jl - DIR, c.ia);
        xx[u] = ld2_hint_nol1(x + o, pol);
        rr[u] = ld2(r + o);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CG_UB; ++u) {
    if (!GEN || u < n) {
      const int t = t0 + u;
      const int jl = m.jl0 + DIR * t;
      const int jg = s.row0 + jl;
      Row nw = {pr[u].x, pr[u].y, ph[u]};
      if (!GEN || t >= 2) {
        const int jc = jl - DIR;
        const int jgc = jg - DIR;
        const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
        double la, lb;
        if (DIR > 0) lap_pair<UNIT, XE, GEN>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, GEN>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (GEN && jgc == 0 && c.ia == 0) la = -cur.a;    // pinned row: q[0] = p[0]
        const double xa = xx[u].x + alpha * cur.a;        // x += alpha*p
        const double xb = xx[u].y + alpha * cur.b;
        const double ra = rr[u].x + alpha * la;           // r -= alpha*q,  q = -lap
        const double rb = rr[u].y + alpha * lb;
        const size_t o = cell(s, jc, c.ia);
        if (c.vb) {
          st2_hint(x + o, xa, xb, pol);
          st2(r + o, ra, rb);
        } else if (c.va) {
          st1(x + o, xa);
          st1(r + o, ra);
        }
        if (GEN) peer_row_store(s, c, jc, peer_lo, peer_hi, ra, rb);
        double da = a.invd, db = a.invd;
        if (GEN || XE != X_IN) {
          const bool ring_row = GEN && ((jgc == 0) || (jgc == s.nyg - 1));
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

template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ void phaseB_strip(const CgArgs& a, const double* p, double* x, double* r, double* peer_lo,
                                             double* peer_hi, double alpha, const WarpTile& w, double& acc_rr,
                                             double& acc_rz) {
  const Cols c = strip_cols<XE>(a.s, w.i0);
  const March m = make_march<DIR>(w);
  const int nrows = m.TH + 2;
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  const int head = nrows < 3 ? nrows : 3;
  phaseB_batch<UNIT, XE, true, DIR>(a, c, m, p, x, r, peer_lo, peer_hi, alpha, 0, head, prev, cur, acc_rr, acc_rz);
  int t = head;
  const int lean_end = m.TH;
  for (; t + CG_UB <= lean_end; t += CG_UB)
    phaseB_batch<UNIT, XE, false, DIR>(a, c, m, p, x, r, peer_lo, peer_hi, alpha, t, CG_UB, prev, cur, acc_rr, acc_rz);
  for (; t < nrows; t += CG_UB) {
    const int n = (nrows - t) < CG_UB ? (nrows - t) : CG_UB;
    phaseB_batch<UNIT, XE, true, DIR>(a, c, m, p, x, r, peer_lo, peer_hi, alpha, t, n, prev, cur, acc_rr, acc_rz);
  }
}

// Warp-level dispatch: strips touching the left / right wall (or only partly
// inside the domain) take the XE instantiation.
template <bool UNIT, int DIR>
__device__ __forceinline__ double phaseA_warp(const CgArgs& a, const double* r, const double* pold, double* pnew,
                                              bool first, double beta, const WarpTile& w, double* peer_lo = nullptr,
                                              double* peer_hi = nullptr) {
  const int xk = strip_kind(a.s, w.i0);
  if (xk == X_PART) return phaseA_strip<UNIT, X_PART, DIR>(a, r, pold, pnew, peer_lo, peer_hi, first, beta, w);
  if (xk == X_WALL) return phaseA_strip<UNIT, X_WALL, DIR>(a, r, pold, pnew, peer_lo, peer_hi, first, beta, w);
  return phaseA_strip<UNIT, X_IN, DIR>(a, r, pold, pnew, peer_lo, peer_hi, first, beta, w);
}
template <bool UNIT, int DIR>
__device__ __forceinline__ void phaseB_warp(const CgArgs& a, const double* p, double* x, double* r, double alpha,
                                            const WarpTile& w, double& acc_rr, double& acc_rz,
                                            double* peer_lo = nullptr, double* peer_hi = nullptr) {
  const int xk = strip_kind(a.s, w.i0);
  if (xk == X_PART) phaseB_strip<UNIT, X_PART, DIR>(a, p, x, r, peer_lo, peer_hi, alpha, w, acc_rr, acc_rz);
  else if (xk == X_WALL) phaseB_strip<UNIT, X_WALL, DIR>(a, p, x, r, peer_lo, peer_hi, alpha, w, acc_rr, acc_rz);
  else phaseB_strip<UNIT, X_IN, DIR>(a, p, x, r, peer_lo, peer_hi, alpha, w, acc_rr, acc_rz);
}

// Sum of NV doubles per thread over the block with ONE pair of barriers; every thread
// returns the totals.  sm holds >= 32*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
    for (int q = 0; q < nw; ++q) t += sm[k * 32 + q];
    v[k] = t;
  }
}

// Fold per-CTA partials; identical in every CTA.
__device__ __forceinline__ double fold_partials(const double* part, int n, double* sm) {
  double acc = 0.0;
  for (int q = threadIdx.x; q < n; q += blockDim.x) acc += ld_cg(part + q);
  return block_sum(acc, sm);
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS, PF_CG_MINB) k_cg_persistent(CgArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[2 * 32];
  const int G = gridDim.x;
  const int bid = blockIdx.x;

  const double bb = ld_cg(a.scal + S_BB);
  const double bnorm = sqrt(bb);
  const double t0 = a.tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;        // max(0.0, rtol*||b||) with Python's NaN behaviour
  int iters = 0;
  int info = a.maxiter;
  double rr = bb;
  double rho = ld_cg(a.scal + S_BZ);
  double rho_prev = 0.0;
  double* pcur = a.p0;
  double* pnext = a.p1;

  if (bnorm == 0.0) {                                 // scipy: "return b, 0"; x was zeroed by k_rhs2
    info = 0;
  } else {
    // number of tiles this CTA owns: t = bid + k*G
    const int ktiles = (a.ntiles > bid) ? (a.ntiles - bid + G - 1) / G : 0;
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) {                          // top-of-loop test on the recursive residual
        info = 0;
        break;
      }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      // ---- phase A: top -> bottom
      double acc = 0.0;
      for (int k = 0; k < ktiles; ++k) {
        const WarpTile w = warp_tile(a, bid + k * G);
        if (w.active) acc += phaseA_warp<UNIT, +1>(a, a.r, pcur, pnext, first, beta, w);
      }
      acc = block_sum(acc, sm);
      if (threadIdx.x == 0) a.part[bid] = -acc;       // p.q = -(sum p*lap)
      grid.sync();
      const double pq = fold_partials(a.part, G, sm);
      const double alpha = rho / pq;
      // ---- phase B: bottom -> top
      double arr = 0.0, arz = 0.0;
      for (int k = ktiles - 1; k >= 0; --k) {
        const WarpTile w = warp_tile(a, bid + k * G);
        if (w.active) phaseB_warp<UNIT, -1>(a, pnext, a.x, a.r, alpha, w, arr, arz);
      }
      {                                              // both partials with one block reduction
        double v2[2] = {arr, arz};
        block_sum_n<2>(v2, sm);
        arr = v2[0];
        arz = v2[1];
      }
      if (threadIdx.x == 0) {
        a.part[CG_MAX_GRID + bid] = arr;
        a.part[2 * CG_MAX_GRID + bid] = arz;
      }
      grid.sync();
      double rz;
      {                                              // both folds in one pass, one block reduction
        double f[2] = {0.0, 0.0};
        for (int q = threadIdx.x; q < G; q += blockDim.x) {
          f[0] += ld_cg(a.part + CG_MAX_GRID + q);
          f[1] += ld_cg(a.part + 2 * CG_MAX_GRID + q);
        }
        block_sum_n<2>(f, sm);
        rr = f[0];
        rz = f[1];
      }
      rho_prev = rho;
      rho = rz;
      double* tmp = pcur; pcur = pnext; pnext = tmp;
      ++iters;
    }
  }
  if (bid == 0 && threadIdx.x == 0) {
    a.scal[S_INFO] = (double)info;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
  }
}


// ---- multi-GPU persistent variant: compute + collective in ONE kernel ----------
// Every rank runs the same persistent kernel on its slab.  Slab-boundary rows of
// p (phase A) and r (phase B) are stored straight into the neighbour's halo rows
// through NVLink peer pointers while the sweep runs.  The grid barrier that ends
// a phase doubles as the cross-GPU collective: the last CTA to arrive on a GPU
// folds the per-CTA partials, publishes them into every rank's mailbox (peer
// stores + system fence + epoch flag), waits for every rank's flag, sums the
// mailbox in RANK ORDER (so all ranks get bit-identical scalars) and releases
// the local CTAs.  One barrier per phase covers halo visibility and the
// all-reduce; nothing returns to the host until the solve is over.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long CG_SPIN_LIMIT = 20000000000LL;   // ~10 s of SM clocks: never hang the GPU

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Optional timeline of a few barriers (opt bit 4): 8 u64 per traced barrier in the
// unused upper half of the first partials array.
__device__ __forceinline__ unsigned long long* trace_slot(const CgArgs& a, unsigned long long e0, unsigned long long e) {
  if (!(a.y.opt & 4)) return nullptr;
  const unsigned long long k = e - e0;              // barrier index within this launch (1-based)
  if (k < 401 || k >= 409) return nullptr;
  return (unsigned long long*)(a.part + PF_TRACE_OFFSET) + (k - 401) * 8;
}

// Packets of the exchanges below: three doubles as six LL words (pf_common.cuh) in a 64-byte slot; a
// reader accepts a packet when every word carries the tag of the epoch it waits for, so flag and payload
// need no fence between them and no access wider than 64 bits has to be indivisible.
// Ends a phase on 1..8 GPUs.  Returns false if a spin-wait timed out (error flag).
// Chain after the last CTA arrives: fold the per-CTA partials (one pass, one block
// reduction) -> (R > 1: one system fence, one LL packet per peer, wait for the
// peers' packets, sum in RANK order) -> one LL release packet -> the other CTAs
// find tag and sums in the same words -> one acquire load per CTA.
template <int NV>
__device__ __forceinline__ bool grid_exchange(const CgArgs& a, unsigned long long e, double (&vals)[NV],
                                              double (&out)[NV], double* sm, unsigned long long e0 = 0) {
  static_assert(NV <= 3, "packet carries three values");
  unsigned long long* tr = trace_slot(a, e0, e);
  __shared__ int s_last;
  __shared__ double s_out[4];
  const CgSync& y = a.y;
  const int G = gridDim.x, tid = threadIdx.x, R = y.nranks;
  const int par = (int)(e & 1ULL);
  block_sum_n<NV>(vals, sm);         // its barriers also order every thread's phase stores before thread 0's fence
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) a.part[k * CG_MAX_GRID + blockIdx.x] = vals[k];
    // cumulative release: publishes the CTA's field / partial / peer-halo stores
    if (R > 1) fence_acq_rel_sys(); else fence_acq_rel_gpu();
    if (tr) {
      const unsigned long long t = gtime();
      atomicMin(tr + 0, t);
      atomicMax(tr + 1, t);
      const unsigned long long kb = e - e0 - 401;          // per-CTA stamps of the first two traced barriers
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
    if (tid == 0) fence_acq_rel_gpu();   // acquire the other CTAs' partials (read below with ld.cg)
    __syncthreads();
    double loc[3] = {0.0, 0.0, 0.0};
    for (int q = tid; q < G; q += CG_THREADS) {
#pragma unroll
      for (int k = 0; k < NV; ++k) loc[k] += ld_cg(a.part + k * CG_MAX_GRID + q);
    }
    block_sum_n<3>(loc, sm);
    if (tid == 0 && tr) tr[2] = gtime();
    if (R > 1) {
      const size_t slot_of_me = (size_t)(par * CG_MAX_RANKS + y.rank) * LL_SLOT_BYTES;
      if (tid == 0) {
        fence_acq_rel_sys();         // the halo rows pushed by the other CTAs of this GPU go out before the packet
        for (int q = 0; q < R; ++q) st_ll_triple<LL_SYS>((char*)y.peer_mbox[q] + slot_of_me, e, loc[0], loc[1], loc[2]);
        if (tr) tr[3] = gtime();
      }
      __shared__ double s_peer[CG_MAX_RANKS][3];
      if (tid < R) {
        const long long t0 = clock64();
        double k[3];
        const char* src = (const char*)y.mbox + (size_t)(par * CG_MAX_RANKS + tid) * LL_SLOT_BYTES;
        while (!ld_ll_triple<LL_SYS>(src, e, k)) {
          if (clock64() - t0 > CG_SPIN_LIMIT) {
            atomicExch(y.err, 1ULL);
            break;
          }
        }
        (void)ld_acquire_sys((const unsigned long long*)src);   // acquire: the peers' halo-row stores are visible now
        s_peer[tid][0] = k[0]; s_peer[tid][1] = k[1]; s_peer[tid][2] = k[2];
      }
      __syncthreads();
      if (tid == 0) {
        for (int k = 0; k < 3; ++k) {   // rank order, identical on every GPU
          double sum = 0.0;
          for (int q = 0; q < R; ++q) sum += s_peer[q][k];
          loc[k] = sum;
        }
      }
    }
    if (tid == 0) {
      if (tr) tr[4] = gtime();
      fence_acq_rel_gpu();           // (R > 1) peers' halo rows, made visible to this thread, go before the release
      st_ll_triple<LL_GPU>(y.gsum, e, loc[0], loc[1], loc[2]);
      if (tr) tr[5] = gtime();
    }
  }
  if (tid == 0) {
    const long long t0 = clock64();
    double k[3];
    while (!ld_ll_triple<LL_GPU>(y.gsum, e, k)) {
      if (clock64() - t0 > CG_SPIN_LIMIT) {
        atomicExch(y.err, 1ULL);
        break;
      }
    }
    (void)ld_acquire_gpu((const unsigned long long*)y.gsum);   // acquire: later loads (incl. halo rows written by peers) see fresh data
    s_out[0] = k[0]; s_out[1] = k[1]; s_out[2] = k[2];
    s_out[3] = (ld_relaxed_u64(y.err) == 0ULL) ? 1.0 : 0.0;
    if (tr && blockIdx.x == 0) tr[6] = gtime();
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) out[k] = s_out[k];
  return s_out[3] != 0.0;
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS, PF_CG_MINB) k_cg_persistent_multi(CgArgs a) {
  __shared__ double sm[3 * 32];
  const int G = gridDim.x;
  const int bid = blockIdx.x;
  const CgSync& y = a.y;

  const double bb = ld_cg(a.scal + S_BB);            // already all-reduced by the host sequence
  const double bnorm = sqrt(bb);
  const double t0 = a.tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;
  int iters = 0;
  int info = a.maxiter;
  double rr = bb;
  double rho = ld_cg(a.scal + S_BZ);
  double rho_prev = 0.0;
  double* pcur = a.p0;
  double* pnext = a.p1;
  unsigned long long e = *((volatile unsigned long long*)y.epoch);   // written only by the previous launch
  const unsigned long long e0 = e;
  bool ok = true;

  if (bnorm == 0.0) {
    info = 0;
  } else {
    const int ktiles = (a.ntiles > bid) ? (a.ntiles - bid + G - 1) / G : 0;
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) {
        info = 0;
        break;
      }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      double* plo = (pnext == a.p0) ? y.below_p0 : y.below_p1;
      double* phi = (pnext == a.p0) ? y.above_p0 : y.above_p1;
      double va[1] = {0.0}, oa[1];
      for (int k = 0; k < ktiles; ++k) {
        const WarpTile w = warp_tile(a, bid + k * G);
        if (w.active) va[0] += phaseA_warp<UNIT, +1>(a, a.r, pcur, pnext, first, beta, w, plo, phi);
      }
      va[0] = -va[0];
      ok = grid_exchange<1>(a, ++e, va, oa, sm, e0);
      if (!ok) break;
      const double alpha = rho / oa[0];
      double vb[2] = {0.0, 0.0}, ob[2];
      for (int k = ktiles - 1; k >= 0; --k) {
        const WarpTile w = warp_tile(a, bid + k * G);
        if (w.active) phaseB_warp<UNIT, -1>(a, pnext, a.x, a.r, alpha, w, vb[0], vb[1], y.below_r, y.above_r);
      }
      ok = grid_exchange<2>(a, ++e, vb, ob, sm, e0);
      if (!ok) break;
      rr = ob[0];
      rho_prev = rho;
      rho = ob[1];
      double* tmp = pcur; pcur = pnext; pnext = tmp;
      ++iters;
    }
  }
  if (bid == 0 && threadIdx.x == 0) {
    a.scal[S_INFO] = ok ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    *y.epoch = e;                    // every rank executed the same number of exchanges
  }
}

// ---- split-phase variant (multi-GPU slabs) -----------------------------------
// Same tile functions, one launch per phase, so that the host can put the
// neighbour halo exchange and the scalar all-reduce (NCCL, same stream) between
// the phases.  All per-iteration scalars live in the device scalar block; a
// device-side `done` flag turns the remaining launches into no-ops once the
// stop test fires, the host only polls it now and then.
//   begin -> [A -> halo(p) + allreduce(p.q) -> B -> halo(r) + allreduce(r.r, r.z) -> update] x maxiter
struct FoldScratch {
  unsigned* counter;
};

// Last CTA to finish folds the per-CTA partials of `nv` values in fixed order.
template <int NV>
__device__ __forceinline__ void split_fold_store(const CgArgs& a, double (&vals)[NV], unsigned* counter,
                                                 const int (&slots)[NV], double* sm) {
  __shared__ bool is_last;
  const int G = gridDim.x;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double t = block_sum(vals[k], sm);
    if (threadIdx.x == 0) a.part[k * CG_MAX_GRID + blockIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == (unsigned)G - 1u);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double t = fold_partials(a.part + k * CG_MAX_GRID, G, sm);
    if (threadIdx.x == 0) a.scal[slots[k]] = t;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

__global__ void k_cg_begin(double* scal, double tol) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double bb = scal[S_BB];
  const double bnorm = sqrt(bb);
  const double t0 = tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;
  scal[S_ATOL] = atol;
  scal[S_RR] = bb;
  scal[S_RHO] = scal[S_BZ];
  scal[S_RHO_PREV] = 0.0;
  scal[S_ITERS] = 0.0;
  double done = 0.0, info = -1.0;          // info < 0: still running (host turns it into maxiter)
  if (bnorm == 0.0 || sqrt(bb) < atol) {   // "return b, 0" / converged before the first iteration
    done = 1.0;
    info = 0.0;
  }
  scal[S_DONE] = done;
  scal[S_INFO] = info;
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS, PF_CG_MINB) k_cg_phaseA(CgArgs a, int it, unsigned* counter) {
  __shared__ double sm[32];
  if (ld_cg(a.scal + S_DONE) != 0.0) return;
  const int G = gridDim.x, bid = blockIdx.x;
  const bool first = (it == 0);
  const double beta = first ? 0.0 : ld_cg(a.scal + S_RHO) / ld_cg(a.scal + S_RHO_PREV);
  double* pcur = (it & 1) ? a.p1 : a.p0;
  double* pnext = (it & 1) ? a.p0 : a.p1;
  const int ktiles = (a.ntiles > bid) ? (a.ntiles - bid + G - 1) / G : 0;
  double acc = 0.0;
  for (int k = 0; k < ktiles; ++k) {
    const WarpTile w = warp_tile(a, bid + k * G);
    if (w.active) acc += phaseA_warp<UNIT, +1>(a, a.r, pcur, pnext, first, beta, w);
  }
  double vals[1] = {-acc};
  const int slots[1] = {S_PQ};
  split_fold_store<1>(a, vals, counter, slots, sm);
}

template <bool UNIT>
__global__ void __launch_bounds__(CG_THREADS, PF_CG_MINB) k_cg_phaseB(CgArgs a, int it, unsigned* counter) {
  __shared__ double sm[32];
  if (ld_cg(a.scal + S_DONE) != 0.0) return;
  const int G = gridDim.x, bid = blockIdx.x;
  const double alpha = ld_cg(a.scal + S_RHO) / ld_cg(a.scal + S_PQ);
  double* pnext = (it & 1) ? a.p0 : a.p1;
  const int ktiles = (a.ntiles > bid) ? (a.ntiles - bid + G - 1) / G : 0;
  double vals[2] = {0.0, 0.0};
  for (int k = ktiles - 1; k >= 0; --k) {
    const WarpTile w = warp_tile(a, bid + k * G);
    if (w.active) phaseB_warp<UNIT, -1>(a, pnext, a.x, a.r, alpha, w, vals[0], vals[1]);
  }
  const int slots[2] = {S_RR_NEW, S_RZ_NEW};
  split_fold_store<2>(a, vals, counter, slots, sm);
}

__global__ void k_cg_finish(double* scal, int maxiter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (scal[S_INFO] < 0.0) scal[S_INFO] = (double)maxiter;     // loop exhausted: info = maxiter
}

__global__ void k_cg_update(double* scal, int it, int maxiter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (scal[S_DONE] != 0.0) return;
  scal[S_RHO_PREV] = scal[S_RHO];
  scal[S_RHO] = scal[S_RZ_NEW];
  const double rr = scal[S_RR_NEW];
  scal[S_RR] = rr;
  scal[S_ITERS] = (double)(it + 1);
  // top-of-loop test of the NEXT iteration; none after the last one (scipy)
  if (it + 1 < maxiter && sqrt(rr) < scal[S_ATOL]) {
    scal[S_DONE] = 1.0;
    scal[S_INFO] = 0.0;
  }
}

}  // namespace pf
