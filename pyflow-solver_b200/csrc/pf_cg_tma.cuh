// pf_cg_tma.cuh -- the CG sweeps fed by the TMA engine (bulk asynchronous copies)
// through a shared-memory ring, for grids whose width is a multiple of 1024.
//
// Why: the register-window kernel (pf_cg.cuh) keeps at most ~100 KB of loads in
// flight per SM and only while its warps sit in the load part of a batch; ncu
// shows it latency-bound (long-scoreboard 12 per issue, 21 % issue slots, 5.3 TB/s
// of DRAM traffic).  Here one producer thread per CTA streams whole 8 KB row
// segments with cp.async.bulk (global -> shared, completion on an mbarrier) up to
// 8 rows (~200 KB) ahead, independent of registers; 16 compute warps consume rows
// from shared memory.  Same arithmetic, same operation order, same reductions as
// pf_cg.cuh -- only where the operands come from changes.
//
//   CTA  = 16 compute warps (64 columns each, 2 per lane) + 1 producer warp, 1 CTA/SM
//   tile = 1024 columns x band; grid = #SMs, tile = blockIdx (static)
//   ring = 8 stages; stage = one row visit: phase A {r, p_old} (2 x 8224 B incl. 2+2
//          halo columns), phase B {p (with halo), x, r of the centre row} (24608 B)
//   full[s]  mbarrier: producer arrive.expect_tx + TMA complete_tx
//   empty[s] mbarrier: one arrive per compute warp after its reads
// Row sequence numbers run on across phases and iterations, so stage = n % 8 and the
// mbarrier parity = (n / 8) & 1 need no per-phase reset.
#pragma once

#include <cooperative_groups.h>

#include "pf_cg.cuh"

namespace pf {


constexpr int TM_WARPS = 16;
constexpr int TM_THREADS = (TM_WARPS + 1) * 32;
constexpr int TM_CW = TM_WARPS * STRIP;          // 1024 columns per CTA tile
constexpr int TM_HALO = 2;                       // halo columns kept on each side (16-byte alignment)
constexpr int TM_ROW = TM_CW + 2 * TM_HALO;      // doubles per haloed smem row
constexpr int TM_STAGES = 8;
constexpr int TM_STAGE_BYTES = (TM_ROW + 2 * TM_CW) * 8;            // 24608: phase B layout is the larger one
constexpr int TM_SMEM_BYTES = TM_STAGES * TM_STAGE_BYTES + 2 * TM_STAGES * 8 + 64;
static_assert(TM_STAGE_BYTES % 16 == 0, "bulk copies need 16-byte aligned destinations");

// ---- mbarrier / bulk-copy PTX ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// Address-based variants for the hot loops.  The 32-bit shared address of the ring is pinned in a register once
// per kernel (an opaque move): left to itself the compiler re-derives it from the CTA's shared window (S2R
// SR_CgaCtaId + MOV + VIADD + LEA) at every mbarrier operation of every row visit.
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr_pin(const void* p) {
  uint32_t o;
  asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(smem_u32(p)));
  return o;
}
__device__ __forceinline__ bool mbar_try_wait_at(smem_addr_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_at(smem_addr_t bar, uint32_t parity) {
  if (mbar_try_wait_at(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait_at(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ void mbar_arrive_at(smem_addr_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double2 lds2_at(smem_addr_t p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(p));
  return v;
}
__device__ __forceinline__ double lds1_at(smem_addr_t p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(p));
  return v;
}

struct TmaTile {
  int i0;            // first column of the tile
  int j0, j1;        // local rows [j0, j1)
  int cs, ce;        // columns actually copied [cs, ce) (halo clipped at the walls)
  bool active;
};

__device__ __forceinline__ TmaTile tma_tile(const CgArgs& a, int t) {
  TmaTile w;
  const int c = t % a.nchunks, bd = t / a.nchunks;
  w.active = t < a.ntiles;
  w.i0 = c * TM_CW;
  w.j0 = (int)((long long)bd * a.s.nyl / a.nbands);
  w.j1 = (int)((long long)(bd + 1) * a.s.nyl / a.nbands);
  w.cs = w.i0 - TM_HALO < 0 ? 0 : w.i0 - TM_HALO;
  w.ce = w.i0 + TM_CW + TM_HALO > a.s.nx ? a.s.nx : w.i0 + TM_CW + TM_HALO;
  return w;
}

struct Ring {
  unsigned char* base;
  uint64_t* full;
  uint64_t* empty;
  smem_addr_t base_at, full_at, empty_at;      // the same three as pinned shared addresses (lean loops)
  __device__ __forceinline__ double* stage(int s) const { return (double*)(base + (size_t)s * TM_STAGE_BYTES); }
};

// ---- producer: one thread streams the rows of a tile in visiting order ------------
// Visit t of a tile (DIR = +1 top->bottom, -1 bottom->top): local row jl = jl0 + DIR*t.
// Rows outside the global domain are not copied (the consumers mask them anyway).
template <int DIR>
__device__ __forceinline__ bool produce_A(const CgArgs& a, const Ring& rg, const TmaTile& w, const double* r,
                                          const double* pold, bool first, uint64_t pol, unsigned long long& n) {
  const Slab& s = a.s;
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  const uint32_t bytes = (uint32_t)(w.ce - w.cs) * 8u;
  const int soff = w.cs - (w.i0 - TM_HALO);                    // where the copied segment starts inside the haloed row
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.empty + st, (uint32_t)(((n / TM_STAGES) & 1ULL) ^ 1ULL))) return false;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    double* dst = rg.stage(st);
    if (jg >= 0 && jg < s.nyg) {
      mbar_expect_tx(rg.full + st, first ? bytes : 2u * bytes);
      bulk_g2s(dst + soff, r + cell(s, jl, w.cs), bytes, rg.full + st);
      if (!first) bulk_g2s_hint(dst + TM_ROW + soff, pold + cell(s, jl, w.cs), bytes, rg.full + st, pol);
    } else {
      mbar_expect_tx(rg.full + st, 0u);
    }
  }
  return true;
}

template <int DIR>
__device__ __forceinline__ bool produce_B(const CgArgs& a, const Ring& rg, const TmaTile& w, const double* p,
                                          const double* x, const double* r, uint64_t pol, unsigned long long& n) {
  const Slab& s = a.s;
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  const uint32_t pbytes = (uint32_t)(w.ce - w.cs) * 8u;
  const uint32_t cbytes = (uint32_t)TM_CW * 8u;
  const int soff = w.cs - (w.i0 - TM_HALO);
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.empty + st, (uint32_t)(((n / TM_STAGES) & 1ULL) ^ 1ULL))) return false;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    double* dst = rg.stage(st);
    const bool prow = jg >= 0 && jg < s.nyg;
    const bool centre = t >= 2;                                 // x, r of the centre row jl - DIR (always a band row)
    mbar_expect_tx(rg.full + st, (prow ? pbytes : 0u) + (centre ? 2u * cbytes : 0u));
    if (prow) bulk_g2s(dst + soff, p + cell(s, jl, w.cs), pbytes, rg.full + st);
    if (centre) {
      bulk_g2s_hint(dst + TM_ROW, x + cell(s, jl - DIR, w.i0), cbytes, rg.full + st, pol);
      bulk_g2s(dst + TM_ROW + TM_CW, r + cell(s, jl - DIR, w.i0), cbytes, rg.full + st);
    }
  }
  return true;
}

// ---- consumers ------------------------------------------------------------------------
// One compute warp = one 64-column strip of the tile.  Values come from the shared
// ring; the 3-row window of p lives in registers as in pf_cg.cuh.
template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ bool consume_A(const CgArgs& a, const Ring& rg, const TmaTile& w, int warp, double* pnew,
                                          bool first, double beta, unsigned long long& n, double& acc) {
  const Slab& s = a.s;
  const int lane = threadIdx.x & 31;
  const int i0 = w.i0 + warp * STRIP;
  const Cols c = strip_cols<XE>(s, i0);
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  const int k = TM_HALO + warp * STRIP + 2 * lane;                // my first column inside the haloed smem row
  const int kh = (lane == 0) ? k - 1 : k + 2;                    // strip halo column (lanes 0 / 31)
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const double* sr = rg.stage(st);
    const double* sp = sr + TM_ROW;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    const bool exists = jg >= 0 && jg < s.nyg;
    double2 rr = make_double2(0.0, 0.0), pp = make_double2(0.0, 0.0);
    double rh = 0.0, ph = 0.0;
    if (exists) {
      rr = *reinterpret_cast<const double2*>(sr + k);
      if (c.vh) rh = sr[kh];
      if (!first) {
        pp = *reinterpret_cast<const double2*>(sp + k);
        if (c.vh) ph = sp[kh];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(rg.empty + st);                    // this warp is done with the stage
    // generic treatment whenever the produced row or the centre row is a global boundary row
    // (or lies outside the domain) -- decided from global row numbers, not from the tile position:
    // with one-row bands the halo row of an interior tile can be the Jacobi-ring row
    const int jgc_ = jg - DIR;
    const bool gen = (jg <= 0) || (jg >= s.nyg - 1) || (jgc_ <= 0) || (jgc_ >= s.nyg - 1);
    const bool ring_row = (jg == 0) || (jg == s.nyg - 1);
    double da = a.invd, db = a.invd, dh = a.invd;
    if (XE != X_IN || gen) {
      da = (ring_row || c.ring_a) ? 1.0 : a.invd;
      db = (ring_row || c.ring_b) ? 1.0 : a.invd;
      dh = (ring_row || c.ring_h) ? 1.0 : a.invd;
    }
    const double za = da * rr.x, zb = db * rr.y, zh = dh * rh;
    Row nw;
    if (first) {
      nw.a = za; nw.b = zb; nw.h = zh;
    } else {
      nw.a = (pp.x * beta) + za;
      nw.b = (pp.y * beta) + zb;
      nw.h = (ph * beta) + zh;
    }
    if (t >= 1 && t <= TH) st2(pnew + cell(s, jl, c.ia), nw.a, nw.b);
    if (t >= 2) {
      const int jgc = jg - DIR;
      const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
      double la, lb;
      if (gen) {
        if (DIR > 0) lap_pair<UNIT, XE, true>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, true>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (jgc == 0 && c.ia == 0) la = -cur.a;                   // pinned row: q[0] = p[0]
      } else {
        if (DIR > 0) lap_pair<UNIT, XE, false>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, false>(a, c, nw, cur, prev, hs, hn, la, lb);
      }
      acc += cur.a * la;                                          // sum p*lap = -(p.q)
      acc += cur.b * lb;
    }
    prev = cur;
    cur = nw;
  }
  return true;
}

template <bool UNIT, int XE, int DIR>
__device__ __forceinline__ bool consume_B(const CgArgs& a, const Ring& rg, const TmaTile& w, int warp, double* x,
                                          double* r, double alpha, uint64_t pol, unsigned long long& n, double& acc_rr,
                                          double& acc_rz) {
  const Slab& s = a.s;
  const int lane = threadIdx.x & 31;
  const int i0 = w.i0 + warp * STRIP;
  const Cols c = strip_cols<XE>(s, i0);
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  const int k = TM_HALO + warp * STRIP + 2 * lane;
  const int kh = (lane == 0) ? k - 1 : k + 2;
  const int kc = warp * STRIP + 2 * lane;                         // column inside the un-haloed x / r rows
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const double* sp = rg.stage(st);
    const double* sx = sp + TM_ROW;
    const double* sr = sx + TM_CW;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    const bool exists = jg >= 0 && jg < s.nyg;
    Row nw = {0.0, 0.0, 0.0};
    if (exists) {
      const double2 pv = *reinterpret_cast<const double2*>(sp + k);
      nw.a = pv.x; nw.b = pv.y;
      if (c.vh) nw.h = sp[kh];
    }
    double2 xx = make_double2(0.0, 0.0), rr = make_double2(0.0, 0.0);
    if (t >= 2) {
      xx = *reinterpret_cast<const double2*>(sx + kc);
      rr = *reinterpret_cast<const double2*>(sr + kc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(rg.empty + st);
    if (t >= 2) {
      const int jc = jl - DIR;
      const int jgc = jg - DIR;
      const bool gen = (jg <= 0) || (jg >= s.nyg - 1) || (jgc <= 0) || (jgc >= s.nyg - 1);
      const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
      double la, lb;
      if (gen) {
        if (DIR > 0) lap_pair<UNIT, XE, true>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, true>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (jgc == 0 && c.ia == 0) la = -cur.a;
      } else {
        if (DIR > 0) lap_pair<UNIT, XE, false>(a, c, prev, cur, nw, hs, hn, la, lb);
        else lap_pair<UNIT, XE, false>(a, c, nw, cur, prev, hs, hn, la, lb);
      }
      const double xa = xx.x + alpha * cur.a;                     // x += alpha*p
      const double xb = xx.y + alpha * cur.b;
      const double ra = rr.x + alpha * la;                        // r -= alpha*q, q = -lap
      const double rb = rr.y + alpha * lb;
      const size_t o = cell(s, jc, c.ia);
      st2_hint(x + o, xa, xb, pol);
      st2(r + o, ra, rb);
      const bool ring_row = (jgc == 0) || (jgc == s.nyg - 1);
      double da = a.invd, db = a.invd;
      if (XE != X_IN || gen) {
        da = (ring_row || c.ring_a) ? 1.0 : a.invd;
        db = (ring_row || c.ring_b) ? 1.0 : a.invd;
      }
      acc_rr += ra * ra;
      acc_rz += ra * (da * ra);
      acc_rr += rb * rb;
      acc_rz += rb * (db * rb);
    }
    prev = cur;
    cur = nw;
  }
  return true;
}

template <bool UNIT>
__global__ void __launch_bounds__(TM_THREADS, 1) k_cg_tma(CgArgs a) {
  namespace cgr = cooperative_groups;
  cgr::grid_group grid = cgr::this_grid();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double sm[32];
  __shared__ int s_fail;
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = (warp == TM_WARPS);
  Ring rg;
  rg.base = smem_raw;
  rg.full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TM_STAGES * TM_STAGE_BYTES);
  rg.empty = rg.full + TM_STAGES;
  if (tid == 0) {
    for (int q = 0; q < TM_STAGES; ++q) {
      mbar_init(rg.full + q, 1u);
      mbar_init(rg.empty + q, (uint32_t)TM_WARPS);
    }
    s_fail = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const TmaTile w = tma_tile(a, bid);
  const int xk = strip_kind(a.s, w.i0 + (warp < TM_WARPS ? warp : 0) * STRIP);
  const uint64_t pol = policy_evict_first();
  unsigned long long n = 0;                          // row sequence number (same in every thread of the CTA)

  const double bb = ld_cg(a.scal + S_BB);
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
  bool ok = true;

  if (bnorm == 0.0) {
    info = 0;
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) {
        info = 0;
        break;
      }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      // ---------------- phase A: top -> bottom
      double acc = 0.0;
      if (w.active) {
        bool good = true;
        if (producer) {
          if (lane == 0) good = produce_A<+1>(a, rg, w, a.r, pcur, first, pol, n);
          else n += (unsigned long long)(w.j1 - w.j0 + 2);
        } else if (xk == X_WALL) {
          good = consume_A<UNIT, X_WALL, +1>(a, rg, w, warp, pnext, first, beta, n, acc);
        } else {
          good = consume_A<UNIT, X_IN, +1>(a, rg, w, warp, pnext, first, beta, n, acc);
        }
        if (!good) s_fail = 1;
      }
      acc = block_sum(acc, sm);
      if (tid == 0) a.part[bid] = -acc;
      if ((a.y.opt & 4) && it == 200 && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        ((unsigned long long*)(a.part + 4096))[bid] = (gtime() << 10) | (unsigned long long)(smid & 1023u);
      }
      grid.sync();
      if ((a.y.opt & 4) && it == 200 && tid == 0 && bid == 0) ((unsigned long long*)(a.part + 4096))[1024] = gtime();
      const double pq = fold_partials(a.part, G, sm);
      const double alpha = rho / pq;
      // ---------------- phase B: bottom -> top
      double arr = 0.0, arz = 0.0;
      if (w.active) {
        bool good = true;
        if (producer) {
          if (lane == 0) good = produce_B<-1>(a, rg, w, pnext, a.x, a.r, pol, n);
          else n += (unsigned long long)(w.j1 - w.j0 + 2);
        } else if (xk == X_WALL) {
          good = consume_B<UNIT, X_WALL, -1>(a, rg, w, warp, a.x, a.r, alpha, pol, n, arr, arz);
        } else {
          good = consume_B<UNIT, X_IN, -1>(a, rg, w, warp, a.x, a.r, alpha, pol, n, arr, arz);
        }
        if (!good) s_fail = 1;
      }
      arr = block_sum(arr, sm);
      arz = block_sum(arz, sm);
      if (tid == 0) {
        a.part[CG_MAX_GRID + bid] = arr;
        a.part[2 * CG_MAX_GRID + bid] = arz;
        a.part[3 * CG_MAX_GRID + bid] = s_fail ? 1.0 : 0.0;
        if ((a.y.opt & 4) && it == 200) {
          unsigned smid;
          asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
          ((unsigned long long*)(a.part + 4096))[512 + bid] = (gtime() << 10) | (unsigned long long)(smid & 1023u);
        }
      }
      grid.sync();
      if ((a.y.opt & 4) && it == 200 && tid == 0 && bid == 0) ((unsigned long long*)(a.part + 4096))[1025] = gtime();
      if ((a.y.opt & 4) && it == 199 && tid == 0 && bid == 0) ((unsigned long long*)(a.part + 4096))[1026] = gtime();
      rr = fold_partials(a.part + CG_MAX_GRID, G, sm);
      const double rz = fold_partials(a.part + 2 * CG_MAX_GRID, G, sm);
      if (fold_partials(a.part + 3 * CG_MAX_GRID, G, sm) != 0.0) {   // a pipeline wait timed out somewhere
        ok = false;
        break;
      }
      rho_prev = rho;
      rho = rz;
      double* tmp = pcur; pcur = pnext; pnext = tmp;
      ++iters;
    }
  }
  if (bid == 0 && tid == 0) {
    a.scal[S_INFO] = ok ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
  }
}


// =====================================================================================
// Dynamically scheduled variant: the production kernel for every grid at least 1024 columns wide, on one GPU and
// (MULTI) on the row slabs of several.
//
// The globaltimer timeline of k_cg_tma (static: one 1024 x ~110-row tile per SM)
// showed the first CTA finishing phase B after 62 us and the last after 106 us: SMs
// do not get equal shares of the memory system (neighbouring SM ids are slow
// together: GPC / TPC level sharing), so equal static tiles leave a third of the
// chip idle at the end of every sweep.  Here tiles are 1024 columns (the last column
// chunk narrower when nx is not a multiple of 1024) x 12-16 rows; a CTA's FIRST tile
// of a phase is static, the others are DRAWN from an atomic ticket counter --
// ascending in phase A, descending in phase B (grid-wide LIFO: what A wrote last B
// reads first, from L2).  The producer keeps the shared-memory ring full ACROSS tile
// boundaries and across the phase barrier (tmad_prefetch), so a small tile costs two
// extra row visits, not a pipeline drain.  Interior tiles run the lean row loops
// (no per-row predicates, pinned shared addresses, ~85 instructions per row visit
// instead of ~180).  Partial sums are kept per tile (the last compute warp to finish
// a tile adds the 16 warp partials in warp order), folded per group of 64 tiles by
// whichever warp completes the group (fence-free: LL-tagged words) and the <= 256
// group partials in fixed order at the phase end (tile_exchange): the result does not
// depend on which SM processed which tile.
constexpr int TD_ROWS = 16;            // default rows per tile (CgArgs::tile_rows; larger for huge slabs)
constexpr int TD_SLOTS = 4;            // tile descriptors in flight (ring depth 8 visits < 1 tile)

struct DynShared {
  int tile[TD_SLOTS];                  // tile id of tile-sequence slot (-1: end of phase)
  unsigned count[TD_SLOTS];            // compute warps that finished the tile
  double wpart[TD_SLOTS][2][TM_WARPS];
  unsigned long long n_end, seq_end;   // producer's counters at the end of a phase
};

// Tile t = band * nchunks + chunk.  Bands: nb_small small ones (small_rows rows) at the bottom of the slab, then
// big ones (tile_rows rows), then nb_small small ones at the top.  Phase A draws tiles in ascending and phase B
// in descending order, so both sweeps END on small tiles: the spread between the first and the last CTA to
// finish a phase is about one tile duration (10-12 us with uniform 16-row tiles at 4096^2, timeline in
// profiles/r2_barrier_timeline_2gpu_v1.txt).
__host__ __device__ __forceinline__ void tmad_band_rows(const CgArgs& a, int bd, int& j0, int& j1) {
  const int nyl = a.s.nyl;
  const int lo = a.nb_small * a.small_rows;           // rows covered by the small bands at each end
  const int nbig = a.nbands - 2 * a.nb_small;
  if (bd < a.nb_small) {
    j0 = bd * a.small_rows;
    j1 = j0 + a.small_rows;
  } else if (bd < a.nb_small + nbig) {
    j0 = lo + (bd - a.nb_small) * a.tile_rows;
    j1 = j0 + a.tile_rows;
    if (j1 > nyl - lo) j1 = nyl - lo;
  } else {
    j0 = nyl - lo + (bd - a.nb_small - nbig) * a.small_rows;
    j1 = j0 + a.small_rows;
  }
}

// Host side: tile geometry for a slab of nyl rows on a grid of G CTAs (pf_api.cu, tests/hostemu).
// tile_rows_min > 0 forces the big-tile height (tests); small tiles only when the slab is tall enough.
inline void tmad_geometry(CgArgs& a, int G, int force_tile_rows = 0, int force_small_rows = -1, int force_nb_small = -1) {
  const int nyl = a.s.nyl;
  a.nchunks = (a.s.nx + TM_CW - 1) / TM_CW;        // the last chunk may be narrower (any nx >= TM_CW)
  int hb = force_tile_rows > 0 ? force_tile_rows : 16;
  if (force_tile_rows <= 0) {
    while ((long long)a.nchunks * ((nyl + hb - 1) / hb) > CG_MAX_TILES / 2) hb *= 2;
    // 16-row tiles when that gives every CTA >= 8 of them; else 12-row tiles while a CTA still gets >= 4 (4096^2: 9 per
    // CTA; the 8192 x 1024 slab of the 8-GPU strong-scaling run: 4.6 per CTA); shorter tiles only for very short slabs.
    // Measured (one GPU, profiles/r2_tile_rows_sweep_v3.txt, r2_short_slab_tile_rows.txt): with the lean row loops a
    // tile's two extra row visits are cheap enough that 12 rows beat 16 at 4096^2, but on the short slab 6-row tiles
    // (a third more row visits) lose 5 % against 12-row ones even though they balance better.
    const int G8 = 8 * G, G4 = 4 * G, G2 = 2 * G;
    auto tiles = [&](int h) { return (long long)a.nchunks * ((nyl + h - 1) / h); };
    if (hb <= 16 && tiles(16) < G8) {
      if (tiles(12) >= G4) hb = 12;
      else if (tiles(8) >= G4) hb = 8;
      else if (tiles(6) >= G2) hb = 6;
      else hb = 4;
    }
  }
  int hs = force_small_rows >= 0 ? force_small_rows : (hb >= 8 ? hb / 4 : 0);
  // small end bands are OFF by default: measured at 4096^2 they cost more (two extra row visits per tile) than the
  // shorter tail of a sweep saves -- 59.1 vs 55.5 ms per 300-iteration launch (profiles/r2_tile_geometry_sweep.txt)
  int ks = force_nb_small >= 0 ? force_nb_small : 0;
  if (hs <= 0 || ks <= 0 || 2LL * ks * hs > nyl / 2) { hs = 0; ks = 0; }
  a.tile_rows = hb;
  a.small_rows = hs;
  a.nb_small = ks;
  const int mid = nyl - 2 * ks * hs;
  a.nbands = 2 * ks + (mid + hb - 1) / hb;
  a.ntiles = a.nchunks * a.nbands;
  a.wx = TM_WARPS; a.wy = 1;
}

__device__ __forceinline__ TmaTile tmad_tile(const CgArgs& a, int t) {
  TmaTile w;
  const int c = t % a.nchunks, bd = t / a.nchunks;
  w.active = true;
  w.i0 = c * TM_CW;
  tmad_band_rows(a, bd, w.j0, w.j1);
  w.cs = w.i0 - TM_HALO < 0 ? 0 : w.i0 - TM_HALO;
  w.ce = w.i0 + TM_CW + TM_HALO > a.s.nx ? a.s.nx : w.i0 + TM_CW + TM_HALO;
  w.ce += w.ce & 1;                 // bulk copies move multiples of 16 bytes: an odd width takes the padding column along
  return w;                         // (pitch is a multiple of 16 doubles; padding is never stored to, never used)
}

// Producer side.  PHASE_B selects the stage layout / tile order.
//
// Every CTA's FIRST tile of a phase is static (tile b in phase A, ntiles-1-b in phase B for CTA b; the ticket
// counter hands out the rest), so it is known before the phase starts -- and the operands of its first visits that
// the CURRENT phase does not write can be fetched BEFORE the barrier that separates the phases:
//   next phase B: x and r of the centre rows (phase A writes neither);   not p (phase A writes it)
//   next phase A: p_old = this iteration's new p (phase B only reads it); not r (phase B writes it)
// The producer reserves up to TM_STAGES stages for those visits (arrive.expect_tx with the visit's FULL byte count),
// issues the safe copies, and after the barrier only adds the missing ones -- which were written moments ago and
// come from L2.  While the CTAs sit in the barrier (1.5-2 us on one GPU, 4-6 us across GPUs) the memory system keeps
// streaming ~100 KB per SM that the sweep needs anyway, and the compute warps start the phase on a full ring
// instead of waiting for a ticket and a DRAM round trip.
struct Prefetch {
  int tile;      // first tile of the coming phase whose leading visits are reserved (-1: nothing pending)
  int npref;     // number of reserved visits (the first npref row-sequence numbers of the coming phase)
  bool phase_b;  // the coming phase is a phase B
};

constexpr int TM_ISSUE_ALL = 0, TM_ISSUE_SAFE = 1, TM_ISSUE_REST = 2;

// One row visit of tile w.  MODE ALL: expect_tx + every copy; SAFE: expect_tx for the full visit, only the copies the
// running phase does not write; REST: the copies SAFE left out (the stage is already reserved).
template <bool PHASE_B, int MODE>
__device__ __forceinline__ void tmad_issue_visit(const CgArgs& a, const Ring& rg, const TmaTile& w, int t, int st,
                                                 const double* f0, const double* f1, const double* f2, bool first,
                                                 uint64_t pol, uint64_t keep) {
  const Slab& s = a.s;
  constexpr int DIR = PHASE_B ? -1 : +1;
  const int jl0 = DIR > 0 ? w.j0 - 1 : w.j1;
  const uint32_t hbytes = (uint32_t)(w.ce - w.cs) * 8u;
  int cw = a.s.nx - w.i0 < TM_CW ? a.s.nx - w.i0 : TM_CW;       // centre-row segment, clipped in the last chunk
  cw += cw & 1;
  const uint32_t cbytes = (uint32_t)cw * 8u;
  const int soff = w.cs - (w.i0 - TM_HALO);
  const int jl = jl0 + DIR * t;
  const int jg = s.row0 + jl;
  const bool exists = jg >= 0 && jg < s.nyg;
  double* dst = rg.stage(st);
  if (!PHASE_B) {                                     // {r (written by phase B), p_old (stable)} of the visited row, haloed
    if (MODE != TM_ISSUE_REST) mbar_expect_tx(rg.full + st, exists ? (first ? hbytes : 2u * hbytes) : 0u);
    if (exists) {
      if (MODE != TM_ISSUE_SAFE) bulk_g2s_hint(dst + soff, f0 + cell(s, jl, w.cs), hbytes, rg.full + st, keep);
      if (MODE != TM_ISSUE_REST && !first)
        bulk_g2s_hint(dst + TM_ROW + soff, f1 + cell(s, jl, w.cs), hbytes, rg.full + st, pol);
    }
  } else {                                            // p of the visited row (haloed; written by phase A) + x, r of the centre row (stable)
    const bool centre = t >= 2;
    if (MODE != TM_ISSUE_REST) mbar_expect_tx(rg.full + st, (exists ? hbytes : 0u) + (centre ? 2u * cbytes : 0u));
    if (MODE != TM_ISSUE_SAFE && exists) bulk_g2s_hint(dst + soff, f0 + cell(s, jl, w.cs), hbytes, rg.full + st, keep);
    if (MODE != TM_ISSUE_REST && centre) {
      bulk_g2s_hint(dst + TM_ROW, f1 + cell(s, jl - DIR, w.i0), cbytes, rg.full + st, pol);
      bulk_g2s_hint(dst + TM_ROW + TM_CW, f2 + cell(s, jl - DIR, w.i0), cbytes, rg.full + st, keep);
    }
  }
}

__device__ __forceinline__ int tmad_first_tile(const CgArgs& a, bool phase_b) {
  const int b = (int)blockIdx.x;
  if (b >= a.ntiles) return -1;
  return phase_b ? a.ntiles - 1 - b : b;
}
// tickets of one phase: the tiles beyond the static first ones + one failing draw per CTA
__host__ __device__ __forceinline__ unsigned long long tmad_tickets_per_phase(int ntiles, int G) {
  return (unsigned long long)(ntiles > G ? ntiles - G : 0) + (unsigned long long)G;
}

// Before the barrier that precedes a phase: reserve the leading visits of this CTA's first tile of that phase and
// start the copies the running phase cannot invalidate.  n is the row sequence number the phase will start with.
template <bool PHASE_B>
__device__ __forceinline__ bool tmad_prefetch(const CgArgs& a, const Ring& rg, DynShared* ds, const double* f0,
                                              const double* f1, const double* f2, uint64_t pol, unsigned long long n,
                                              unsigned long long seq, Prefetch& pf) {
  pf.tile = -1;
  pf.npref = 0;
  if (a.y.opt & 32) return true;                      // experiment switch: no prefetch across the barrier
  const int tile = tmad_first_tile(a, PHASE_B);
  if (tile < 0) return true;
  const uint64_t keep = (a.y.opt & 8) ? policy_evict_last() : policy_evict_normal();
  const TmaTile w = tmad_tile(a, tile);
  const int nrows = w.j1 - w.j0 + 2;
  const int np = nrows < TM_STAGES ? nrows : TM_STAGES;
  for (int t = 0; t < np; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.empty + st, (uint32_t)(((n / TM_STAGES) & 1ULL) ^ 1ULL))) return false;
    if (t == 0) ds->tile[(int)(seq % TD_SLOTS)] = tile;
    tmad_issue_visit<PHASE_B, TM_ISSUE_SAFE>(a, rg, w, t, st, f0, f1, f2, false, pol, keep);
  }
  pf.tile = tile;
  pf.npref = np;
  pf.phase_b = PHASE_B;
  return true;
}

// A phase that will not run after all (the solve stopped): complete the reserved visits and wait for the copies,
// so that no bulk copy is in flight into this CTA's shared memory when it exits.
template <bool PHASE_B>
__device__ __forceinline__ void tmad_drain(const CgArgs& a, const Ring& rg, const double* f0, const double* f1,
                                           const double* f2, uint64_t pol, unsigned long long n, Prefetch& pf) {
  if (pf.tile < 0) return;
  const uint64_t keep = policy_evict_normal();
  const TmaTile w = tmad_tile(a, pf.tile);
  for (int t = 0; t < pf.npref; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    tmad_issue_visit<PHASE_B, TM_ISSUE_REST>(a, rg, w, t, st, f0, f1, f2, false, pol, keep);
    (void)mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL));
  }
  pf.tile = -1;
}

template <bool PHASE_B>
__device__ __forceinline__ bool tmad_produce(const CgArgs& a, const Ring& rg, DynShared* ds, unsigned long long tbase,
                                             const double* f0, const double* f1, const double* f2, bool first,
                                             uint64_t pol, unsigned long long& n, unsigned long long& seq,
                                             Prefetch& pf) {
  // r and p are re-read by the next (reversed) sweep.  Marking them evict_last (opt bit 8) was measured:
  // DRAM reads 173.0 GB vs 168.7 GB with the default priority, no gain -- default stays evict_normal.
  const uint64_t keep = (a.y.opt & 8) ? policy_evict_last() : policy_evict_normal();
  const int G = (int)gridDim.x;
  int tile = tmad_first_tile(a, PHASE_B);             // static first tile, then tickets
  int npref = (pf.tile == tile) ? pf.npref : 0;       // its leading visits may already be reserved (tmad_prefetch)
  pf.tile = -1;
  for (;;) {
    if (tile < 0) {
      const unsigned long long tk = (unsigned long long)G + atomicAdd(a.y.tickets, 1ULL) - tbase;
      tile = tk < (unsigned long long)a.ntiles ? (PHASE_B ? a.ntiles - 1 - (int)tk : (int)tk) : -1;
    }
    const int slot = (int)(seq % TD_SLOTS);
    if (tile < 0) {                                   // end marker: an empty visit carrying tile = -1
      const int st = (int)(n % TM_STAGES);
      if (!mbar_wait(rg.empty + st, (uint32_t)(((n / TM_STAGES) & 1ULL) ^ 1ULL))) return false;
      ds->tile[slot] = -1;
      mbar_expect_tx(rg.full + st, 0u);
      ++n;
      ++seq;
      return true;
    }
    const TmaTile w = tmad_tile(a, tile);
    const int nrows = w.j1 - w.j0 + 2;
    for (int t = 0; t < nrows; ++t, ++n) {
      const int st = (int)(n % TM_STAGES);
      if (t < npref) {                                // reserved before the barrier: add what was left out
        tmad_issue_visit<PHASE_B, TM_ISSUE_REST>(a, rg, w, t, st, f0, f1, f2, first, pol, keep);
        continue;
      }
      if (!mbar_wait(rg.empty + st, (uint32_t)(((n / TM_STAGES) & 1ULL) ^ 1ULL))) return false;
      if (t == 0) ds->tile[slot] = tile;             // published by the release of the arrive below
      tmad_issue_visit<PHASE_B, TM_ISSUE_ALL>(a, rg, w, t, st, f0, f1, f2, first, pol, keep);
    }
    ++seq;
    npref = 0;
    tile = -1;
  }
}

// Last compute warp to finish a tile adds the warp partials in warp order.  Tiles are accounted in groups of
// CG_TILE_GROUP consecutive ids: the warp that completes the last tile of a group in this phase folds the
// group's tile partials (fixed shape: lane i adds tiles i and i + 32 of the group, then the xor tree) into one
// group partial -- during the sweep, off the critical path -- so that the phase-end fold reads ntiles / 64
// values instead of ntiles (the last-arriving CTA of the multi-GPU exchange spent 2.3 us at 4096^2 and 10 us at
// 16384^2 folding alone, profiles/r2_barrier_timeline_2gpu_v1.txt).  NO FENCE inside the sweep: a fence makes the
// warp wait for its outstanding field stores and the ring couples all warps of the CTA to the slowest one
// (measured: +3.7 us per phase).  Instead the tile partials travel as LL words tagged with the phase (pf_common.cuh):
// the group counter only ELECTS the folding warp (relaxed atomic, orders nothing), which then polls the group's
// words until each carries the tag.  The result does not depend on which SM processed which tile, nor on which warp
// folds a group; the group partials are plain stores published by the barrier that ends the phase.
__device__ __forceinline__ unsigned long long* tmad_ll_partial(const CgArgs& a, int k, int tile) {
  return (unsigned long long*)a.part + 2 * ((size_t)k * CG_MAX_TILES + (size_t)tile);
}

template <int NV>
__device__ __forceinline__ void tmad_tile_partial(const CgArgs& a, DynShared* ds, int slot, int tile, int warp,
                                                  double (&v)[NV], bool negate, unsigned phase_seq, unsigned tag) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  unsigned prev = 0;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) ds->wpart[slot][k][warp] = v[k];
    __threadfence_block();
    prev = atomicAdd(&ds->count[slot], 1u);
  }
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev == (unsigned)TM_WARPS - 1u) {              // every warp's partial of this tile is in shared memory
    __threadfence_block();
    const int g = tile / CG_TILE_GROUP;
    const int gfirst = g * CG_TILE_GROUP;
    const int gsize = (a.ntiles - gfirst) < CG_TILE_GROUP ? (a.ntiles - gfirst) : CG_TILE_GROUP;
    unsigned done = 0;
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int q = 0; q < TM_WARPS; ++q) t += ds->wpart[slot][k][q];
        st_ll2<LL_GPU>(tmad_ll_partial(a, k, tile), ll_encode(negate ? -t : t, tag));
      }
      ds->count[slot] = 0u;
      done = atomicAdd((unsigned*)(a.part + PF_GCOUNT_OFFSET) + g, 1u);
    }
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done == phase_seq * (unsigned)gsize + (unsigned)gsize - 1u) {      // this warp folds the group
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double t = 0.0;
#pragma unroll
        for (int h = 0; h < CG_TILE_GROUP / 32; ++h) {
          const int q = lane + 32 * h;
          if (q < gsize) {
            const unsigned long long* src = tmad_ll_partial(a, k, gfirst + q);
            LLVal w = ld_ll2<LL_GPU>(src);
            const long long t0 = clock64();
            while (!ll_ok(w, tag)) {
              if (clock64() - t0 > CG_SPIN_LIMIT) {
                atomicExch(a.y.err, 1ULL);
                break;
              }
              w = ld_ll2<LL_GPU>(src);
            }
            t += ll_value(w);
          }
        }
        t = warp_sum(t);
        if (lane == 0) a.part[PF_GPART_OFFSET + k * CG_MAX_GROUPS + g] = t;
      }
    }
  }
}

// ---- lean row loops ---------------------------------------------------------------------------------------
// Tiles whose visited rows and centre rows are all strictly inside the global domain (every tile except those of
// the first and last band) need none of the per-row predicates of the generic loops below: every row exists, no
// Neumann / ring / pin treatment in y, the Jacobi factors are per-lane constants.  The visits are peeled (t = 0: load
// only; t = 1: first store; t = 2 .. TH: store + stencil; t = TH + 1: stencil only) so that the steady loop carries no
// conditions on t either, addresses advance by pointer increments, the ring is addressed through pinned 32-bit
// shared addresses.  Same arithmetic, same operation order as the generic loops: results are bit-identical.
// ncu on the round-1 kernel: 181 warp instructions per 64-column row visit, 48 % issue utilisation, time
// proportional to row VISITS (halo visits that hit L2 cost as much as streamed ones) -- the SM side, not DRAM, set
// the pace at 4096^2.
struct LeanCtx {
  smem_addr_t ring_at, full_at, empty_at;
  uint32_t koff, khoff, kcoff;     // byte offsets of this lane's pair / halo column in a haloed row, pair in an un-haloed row
  bool vh;
  int lane;
};

template <bool UNIT, int XE>
__device__ __forceinline__ bool tmad_lean_A(const CgArgs& a, const LeanCtx& L, const Cols& c, int TH, double* pdst,
                                            double beta, uint64_t keep, unsigned long long& n, double& acc) {
  const size_t pitch = (size_t)a.s.pitch;
  const double da = (XE != X_IN && c.ring_a) ? 1.0 : a.invd;
  const double db = (XE != X_IN && c.ring_b) ? 1.0 : a.invd;
  const double dh = (XE != X_IN && c.ring_h) ? 1.0 : a.invd;
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  auto visit = [&](Row& nw) -> bool {
    const uint32_t st = (uint32_t)n & (TM_STAGES - 1);
    if (!mbar_wait_at(L.full_at + 8u * st, (uint32_t)(n / TM_STAGES) & 1u)) return false;
    const smem_addr_t sa = L.ring_at + st * (uint32_t)TM_STAGE_BYTES;
    const double2 rr = lds2_at(sa + L.koff);
    const double2 pp = lds2_at(sa + (uint32_t)(TM_ROW * 8) + L.koff);
    double rh = 0.0, ph = 0.0;
    if (L.vh) {
      rh = lds1_at(sa + L.khoff);
      ph = lds1_at(sa + (uint32_t)(TM_ROW * 8) + L.khoff);
    }
    __syncwarp();
    if (L.lane == 0) mbar_arrive_at(L.empty_at + 8u * st);
    ++n;
    nw.a = (pp.x * beta) + da * rr.x;
    nw.b = (pp.y * beta) + db * rr.y;
    nw.h = (ph * beta) + dh * rh;
    return true;
  };
  auto stencil = [&](const Row& south, const Row& centre, const Row& north) {
    double la, lb;
    lap_pair<UNIT, XE, false>(a, c, south, centre, north, true, true, la, lb);
    if (XE == X_PART) {                               // strip cut by the last column: per-lane validity
      if (c.va) acc += centre.a * la;
      if (c.vb) acc += centre.b * lb;
    } else {
      acc += centre.a * la;
      acc += centre.b * lb;
    }
  };
  auto store = [&](double* dst, const Row& v) {
    if (XE != X_PART || c.vb) st2_hint(dst, v.a, v.b, keep);
    else if (c.va) st1(dst, v.a);
  };
  Row nw;
  if (!visit(nw)) return false;                       // t = 0: the row below the tile
  cur = nw;
  if (!visit(nw)) return false;                       // t = 1: first tile row
  store(pdst, nw);
  pdst += pitch;
  prev = cur; cur = nw;
#pragma unroll 2
  for (int t = 2; t <= TH; ++t) {
    if (!visit(nw)) return false;
    store(pdst, nw);
    pdst += pitch;
    stencil(prev, cur, nw);
    prev = cur; cur = nw;
  }
  if (!visit(nw)) return false;                       // t = TH + 1: the row above the tile
  stencil(prev, cur, nw);
  return true;
}

template <bool UNIT, int XE>
__device__ __forceinline__ bool tmad_lean_B(const CgArgs& a, const LeanCtx& L, const Cols& c, int TH, double* xdst,
                                            double* rdst, double alpha, uint64_t pol, uint64_t keep,
                                            unsigned long long& n, double& acc_rr, double& acc_rz) {
  const size_t pitch = (size_t)a.s.pitch;
  const double da = (XE != X_IN && c.ring_a) ? 1.0 : a.invd;
  const double db = (XE != X_IN && c.ring_b) ? 1.0 : a.invd;
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  // rows are visited top -> bottom: prev is the row ABOVE cur
  auto visit_p = [&](Row& nw, double2& xx, double2& rr, bool centre) -> bool {
    const uint32_t st = (uint32_t)n & (TM_STAGES - 1);
    if (!mbar_wait_at(L.full_at + 8u * st, (uint32_t)(n / TM_STAGES) & 1u)) return false;
    const smem_addr_t sa = L.ring_at + st * (uint32_t)TM_STAGE_BYTES;
    const double2 pv = lds2_at(sa + L.koff);
    nw.a = pv.x; nw.b = pv.y; nw.h = 0.0;
    if (L.vh) nw.h = lds1_at(sa + L.khoff);
    if (centre) {
      xx = lds2_at(sa + (uint32_t)(TM_ROW * 8) + L.kcoff);
      rr = lds2_at(sa + (uint32_t)((TM_ROW + TM_CW) * 8) + L.kcoff);
    }
    __syncwarp();
    if (L.lane == 0) mbar_arrive_at(L.empty_at + 8u * st);
    ++n;
    return true;
  };
  Row nw;
  double2 xx = make_double2(0.0, 0.0), rr = make_double2(0.0, 0.0);
  if (!visit_p(nw, xx, rr, false)) return false;      // t = 0: the row above the tile
  cur = nw;
  if (!visit_p(nw, xx, rr, false)) return false;      // t = 1: top tile row
  prev = cur; cur = nw;
#pragma unroll 2
  for (int t = 2; t <= TH + 1; ++t) {                 // centre row = cur, visited row = the one below it
    if (!visit_p(nw, xx, rr, true)) return false;
    double la, lb;
    lap_pair<UNIT, XE, false>(a, c, nw, cur, prev, true, true, la, lb);
    const double xa = xx.x + alpha * cur.a;
    const double xb = xx.y + alpha * cur.b;
    const double ra = rr.x + alpha * la;
    const double rb = rr.y + alpha * lb;
    if (XE != X_PART || c.vb) {
      st2_hint(xdst, xa, xb, pol);
      st2_hint(rdst, ra, rb, keep);
    } else if (c.va) {
      st1(xdst, xa);
      st1(rdst, ra);
    }
    xdst -= pitch;
    rdst -= pitch;
    if (XE != X_PART || c.va) {
      acc_rr += ra * ra;
      acc_rz += ra * (da * ra);
    }
    if (XE != X_PART || c.vb) {
      acc_rr += rb * rb;
      acc_rz += rb * (db * rb);
    }
    prev = cur; cur = nw;
  }
  return true;
}

// A warp whose 64-column strip lies entirely beyond the last column (tiles of a clipped last chunk): it only keeps
// the ring protocol going.
__device__ __forceinline__ bool tmad_tile_idle(const Ring& rg, int nrows, int lane, unsigned long long& n) {
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    __syncwarp();
    if (lane == 0) mbar_arrive(rg.empty + st);
  }
  return true;
}

// One tile of phase A for one warp.  XE: strip kind (X_IN / X_WALL: all 64 columns exist; X_PART: the strip is cut by
// the last column -- per-lane validity in the stores and the sums).
template <bool UNIT, int XE, bool MULTI>
__device__ __forceinline__ bool tmad_tile_A(const CgArgs& a, const Ring& rg, LeanCtx& L, const TmaTile& w, int warp,
                                            double* pnew, double* peer_lo, double* peer_hi, bool first, double beta,
                                            uint64_t keep, unsigned long long& n, double& acc) {
  const Slab& s = a.s;
  constexpr int DIR = +1;
  const int lane = L.lane;
  const int k = TM_HALO + warp * STRIP + 2 * lane;                // my first column inside the haloed smem row
  const int kh = (lane == 0) ? k - 1 : k + 2;                    // strip halo column (lanes 0 / 31)
  const int i0 = w.i0 + warp * STRIP;
  const Cols c = strip_cols<XE>(s, i0);
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = w.j0 - 1;
  // lean: no visited or centre row is a global boundary row; with several GPUs also none that is mirrored to a peer
  const bool lean = !first && !(a.y.opt & 16) && (s.row0 + w.j0 - 1 > 0) &&
                    (s.row0 + w.j1 < s.nyg - 1) && !(MULTI && (w.j0 == 0 || w.j1 == s.nyl));
  if (lean) {
    L.vh = c.vh;
    return tmad_lean_A<UNIT, XE>(a, L, c, TH, pnew + cell(s, w.j0, c.ia), beta, keep, n, acc);
  }
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const double* sr = rg.stage(st);
    const double* sp = sr + TM_ROW;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    const bool exists = jg >= 0 && jg < s.nyg;
    double2 rr = make_double2(0.0, 0.0), pp = make_double2(0.0, 0.0);
    double rh = 0.0, ph = 0.0;
    if (exists) {
      rr = *reinterpret_cast<const double2*>(sr + k);
      if (c.vh) rh = sr[kh];
      if (!first) {
        pp = *reinterpret_cast<const double2*>(sp + k);
        if (c.vh) ph = sp[kh];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(rg.empty + st);
    const int jgc = jg - DIR;
    const bool gen = (jg <= 0) || (jg >= s.nyg - 1) || (jgc <= 0) || (jgc >= s.nyg - 1);
    const bool ring_row = (jg == 0) || (jg == s.nyg - 1);
    double da = a.invd, db = a.invd, dh = a.invd;
    if (XE != X_IN || gen) {
      da = (ring_row || c.ring_a) ? 1.0 : a.invd;
      db = (ring_row || c.ring_b) ? 1.0 : a.invd;
      dh = (ring_row || c.ring_h) ? 1.0 : a.invd;
    }
    const double za = da * rr.x, zb = db * rr.y, zh = dh * rh;
    Row nw;
    if (first) {
      nw.a = za; nw.b = zb; nw.h = zh;
    } else {
      nw.a = (pp.x * beta) + za;
      nw.b = (pp.y * beta) + zb;
      nw.h = (ph * beta) + zh;
    }
    if (t >= 1 && t <= TH) {
      double* dst = pnew + cell(s, jl, c.ia);
      if (XE != X_PART || c.vb) st2_hint(dst, nw.a, nw.b, keep);
      else if (c.va) st1(dst, nw.a);
      if (MULTI && (jl == 0 || jl == s.nyl - 1)) peer_row_store(s, c, jl, peer_lo, peer_hi, nw.a, nw.b);
    }
    if (t >= 2) {
      const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
      double la, lb;
      if (gen) {
        lap_pair<UNIT, XE, true>(a, c, prev, cur, nw, hs, hn, la, lb);
        if (jgc == 0 && c.ia == 0) la = -cur.a;
      } else {
        lap_pair<UNIT, XE, false>(a, c, prev, cur, nw, hs, hn, la, lb);
      }
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
  return true;
}

template <bool UNIT, bool MULTI>
__device__ __forceinline__ bool tmad_consume_A(const CgArgs& a, const Ring& rg, DynShared* ds, int warp, double* pnew,
                                               double* peer_lo, double* peer_hi, bool first, double beta,
                                               unsigned long long& n, unsigned long long& seq, unsigned phase_seq,
                                               unsigned ptag) {
  const Slab& s = a.s;
  const int lane = threadIdx.x & 31;
  const uint64_t keep = (a.y.opt & 8) ? policy_evict_last() : policy_evict_normal();
  const int k = TM_HALO + warp * STRIP + 2 * lane;
  const int kh = (lane == 0) ? k - 1 : k + 2;
  LeanCtx L;
  L.ring_at = rg.base_at; L.full_at = rg.full_at; L.empty_at = rg.empty_at;
  L.koff = (uint32_t)k * 8u; L.khoff = (uint32_t)kh * 8u; L.kcoff = (uint32_t)(warp * STRIP + 2 * lane) * 8u;
  L.lane = lane; L.vh = false;
  for (;;) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const int slot = (int)(seq % TD_SLOTS);
    const int tile = ds->tile[slot];
    if (tile < 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(rg.empty + st);
      ++n;
      ++seq;
      return true;
    }
    const TmaTile w = tmad_tile(a, tile);
    const int i0 = w.i0 + warp * STRIP;
    double acc = 0.0;
    bool good;
    if (i0 >= s.nx) {
      good = tmad_tile_idle(rg, w.j1 - w.j0 + 2, lane, n);
    } else {
      const int kind = strip_kind(s, i0);
      if (kind == X_IN) good = tmad_tile_A<UNIT, X_IN, MULTI>(a, rg, L, w, warp, pnew, peer_lo, peer_hi, first, beta, keep, n, acc);
      else if (kind == X_WALL) good = tmad_tile_A<UNIT, X_WALL, MULTI>(a, rg, L, w, warp, pnew, peer_lo, peer_hi, first, beta, keep, n, acc);
      else good = tmad_tile_A<UNIT, X_PART, MULTI>(a, rg, L, w, warp, pnew, peer_lo, peer_hi, first, beta, keep, n, acc);
    }
    if (!good) return false;
    double v[1] = {acc};
    tmad_tile_partial<1>(a, ds, slot, tile, warp, v, /*negate=*/true, phase_seq, ptag);     // p.q = -(sum p*lap)
    ++seq;
  }
}

template <bool UNIT, int XE, bool MULTI>
__device__ __forceinline__ bool tmad_tile_B(const CgArgs& a, const Ring& rg, LeanCtx& L, const TmaTile& w, int warp,
                                            double* x, double* r, double* peer_lo, double* peer_hi, double alpha,
                                            uint64_t pol, uint64_t keep, unsigned long long& n, double& acc_rr,
                                            double& acc_rz) {
  const Slab& s = a.s;
  constexpr int DIR = -1;
  const int lane = L.lane;
  const int k = TM_HALO + warp * STRIP + 2 * lane;
  const int kh = (lane == 0) ? k - 1 : k + 2;
  const int kc = warp * STRIP + 2 * lane;                         // column inside the un-haloed x / r rows
  const int i0 = w.i0 + warp * STRIP;
  const Cols c = strip_cols<XE>(s, i0);
  const int TH = w.j1 - w.j0, nrows = TH + 2;
  const int jl0 = w.j1;
  const bool lean = !(a.y.opt & 16) && (s.row0 + w.j0 - 1 > 0) && (s.row0 + w.j1 < s.nyg - 1) &&
                    !(MULTI && (w.j0 == 0 || w.j1 == s.nyl));
  if (lean) {
    L.vh = c.vh;
    const size_t o = cell(s, w.j1 - 1, c.ia);
    return tmad_lean_B<UNIT, XE>(a, L, c, TH, x + o, r + o, alpha, pol, keep, n, acc_rr, acc_rz);
  }
  Row prev = {0.0, 0.0, 0.0}, cur = {0.0, 0.0, 0.0};
  for (int t = 0; t < nrows; ++t, ++n) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const double* sp = rg.stage(st);
    const double* sx = sp + TM_ROW;
    const double* sr = sx + TM_CW;
    const int jl = jl0 + DIR * t;
    const int jg = s.row0 + jl;
    const bool exists = jg >= 0 && jg < s.nyg;
    Row nw = {0.0, 0.0, 0.0};
    if (exists) {
      const double2 pv = *reinterpret_cast<const double2*>(sp + k);
      nw.a = pv.x; nw.b = pv.y;
      if (c.vh) nw.h = sp[kh];
    }
    double2 xx = make_double2(0.0, 0.0), rr = make_double2(0.0, 0.0);
    if (t >= 2) {
      xx = *reinterpret_cast<const double2*>(sx + kc);
      rr = *reinterpret_cast<const double2*>(sr + kc);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(rg.empty + st);
    if (t >= 2) {
      const int jc = jl - DIR;
      const int jgc = jg - DIR;
      const bool gen = (jg <= 0) || (jg >= s.nyg - 1) || (jgc <= 0) || (jgc >= s.nyg - 1);
      const bool hs = jgc > 0, hn = jgc < s.nyg - 1;
      double la, lb;
      if (gen) {
        lap_pair<UNIT, XE, true>(a, c, nw, cur, prev, hs, hn, la, lb);
        if (jgc == 0 && c.ia == 0) la = -cur.a;
      } else {
        lap_pair<UNIT, XE, false>(a, c, nw, cur, prev, hs, hn, la, lb);
      }
      const double xa = xx.x + alpha * cur.a;
      const double xb = xx.y + alpha * cur.b;
      const double ra = rr.x + alpha * la;
      const double rb = rr.y + alpha * lb;
      const size_t o = cell(s, jc, c.ia);
      if (XE != X_PART || c.vb) {
        st2_hint(x + o, xa, xb, pol);
        st2_hint(r + o, ra, rb, keep);
      } else if (c.va) {
        st1(x + o, xa);
        st1(r + o, ra);
      }
      if (MULTI && (jc == 0 || jc == s.nyl - 1)) peer_row_store(s, c, jc, peer_lo, peer_hi, ra, rb);
      const bool ring_row = (jgc == 0) || (jgc == s.nyg - 1);
      double da = a.invd, db = a.invd;
      if (XE != X_IN || gen) {
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
  return true;
}

template <bool UNIT, bool MULTI>
__device__ __forceinline__ bool tmad_consume_B(const CgArgs& a, const Ring& rg, DynShared* ds, int warp, double* x,
                                               double* r, double* peer_lo, double* peer_hi, double alpha, uint64_t pol,
                                               unsigned long long& n, unsigned long long& seq, unsigned phase_seq,
                                               unsigned ptag) {
  const Slab& s = a.s;
  const int lane = threadIdx.x & 31;
  const uint64_t keep = (a.y.opt & 8) ? policy_evict_last() : policy_evict_normal();
  const int k = TM_HALO + warp * STRIP + 2 * lane;
  const int kh = (lane == 0) ? k - 1 : k + 2;
  LeanCtx L;
  L.ring_at = rg.base_at; L.full_at = rg.full_at; L.empty_at = rg.empty_at;
  L.koff = (uint32_t)k * 8u; L.khoff = (uint32_t)kh * 8u; L.kcoff = (uint32_t)(warp * STRIP + 2 * lane) * 8u;
  L.lane = lane; L.vh = false;
  for (;;) {
    const int st = (int)(n % TM_STAGES);
    if (!mbar_wait(rg.full + st, (uint32_t)((n / TM_STAGES) & 1ULL))) return false;
    const int slot = (int)(seq % TD_SLOTS);
    const int tile = ds->tile[slot];
    if (tile < 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(rg.empty + st);
      ++n;
      ++seq;
      return true;
    }
    const TmaTile w = tmad_tile(a, tile);
    const int i0 = w.i0 + warp * STRIP;
    double v[2] = {0.0, 0.0};
    bool good;
    if (i0 >= s.nx) {
      good = tmad_tile_idle(rg, w.j1 - w.j0 + 2, lane, n);
    } else {
      const int kind = strip_kind(s, i0);
      if (kind == X_IN) good = tmad_tile_B<UNIT, X_IN, MULTI>(a, rg, L, w, warp, x, r, peer_lo, peer_hi, alpha, pol, keep, n, v[0], v[1]);
      else if (kind == X_WALL) good = tmad_tile_B<UNIT, X_WALL, MULTI>(a, rg, L, w, warp, x, r, peer_lo, peer_hi, alpha, pol, keep, n, v[0], v[1]);
      else good = tmad_tile_B<UNIT, X_PART, MULTI>(a, rg, L, w, warp, x, r, peer_lo, peer_hi, alpha, pol, keep, n, v[0], v[1]);
    }
    if (!good) return false;
    tmad_tile_partial<2>(a, ds, slot, tile, warp, v, /*negate=*/false, phase_seq, ptag);
    ++seq;
  }
}

// Phase end: barrier + all-reduce in one, on one GPU (SYS = false) or across the GPUs of a box (SYS = true).
// Every CTA arrives (release fence + atomic on the GPU's counter: its stores, peer halo rows included, happen
// before the CTA that arrives last); the last CTA folds the group partials (<= CG_MAX_GROUPS values) while the
// thread that will publish runs the one fence that makes everything visible at the scope of the exchange
// (cumulative: it also covers the halo rows pushed to the peers by the other CTAs), then writes one LL packet
// (three doubles as six self-validating words) into the mailbox of EVERY rank, its own included.  All CTAs of all
// ranks poll their local mailbox directly -- no separate local release hop -- and add the packets in rank order, so
// every CTA of every GPU holds bit-identical sums.  On one GPU this replaces cooperative-groups grid.sync() followed
// by a fold in every CTA: one atomic, one 1-2 sector fold, one packet, instead of two sequentially consistent fences
// per CTA, a polled barrier word and 148 folds.
template <int NV, bool SYS>
__device__ __forceinline__ bool tile_exchange(const CgArgs& a, unsigned long long e, double (&out)[NV],
                                              unsigned long long e0 = 0) {
  constexpr int SCOPE = SYS ? LL_SYS : LL_GPU;
  unsigned long long* tr = trace_slot(a, e0, e);
  __shared__ int s_last;
  __shared__ double s_out[4];
  __shared__ double s_peer[CG_MAX_RANKS][3];
  const CgSync& y = a.y;
  const int G = gridDim.x, tid = threadIdx.x, R = SYS ? y.nranks : 1;
  const int par = (int)(e & 1ULL);
  __syncthreads();                   // every thread's stores (fields, group partials, peer halos) before thread 0's fence
  if (tid == 0) {
    fence_acq_rel_gpu();             // release
    if (tr) { const unsigned long long t = gtime(); atomicMin(tr + 0, t); atomicMax(tr + 1, t); }
    s_last = (atomicAdd(y.arrive, 1ULL) == e * (unsigned long long)G - 1ULL) ? 1 : 0;
    if (s_last) fence_acq_rel_gpu(); // acquire: every CTA's group partials and peer halo stores happen-before what follows
  }
  __syncthreads();
  if (s_last) {
    // the fence of the publishing thread runs WHILE warp 0 folds the group partials
    const bool publisher = (tid == TM_WARPS * 32);
    if (publisher) {
      if (SYS) fence_acq_rel_sys(); else fence_acq_rel_gpu();
    }
    if (tid < 32) {
      const int ng = (a.ntiles + CG_TILE_GROUP - 1) / CG_TILE_GROUP;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double t = 0.0;
        if (k < NV) {
          for (int q = tid; q < ng; q += 32) t += ld_cg(a.part + PF_GPART_OFFSET + k * CG_MAX_GROUPS + q);
          t = warp_sum(t);
        }
        if (tid == 0) s_out[k] = t;
      }
      if (tid == 0 && tr) tr[2] = gtime();
    }
    __syncthreads();
    if (publisher) {
      const size_t slot = (size_t)(par * CG_MAX_RANKS + (SYS ? y.rank : 0)) * LL_SLOT_BYTES;
      if (SYS) {
        for (int q = 0; q < R; ++q) st_ll_triple<LL_SYS>((char*)y.peer_mbox[q] + slot, e, s_out[0], s_out[1], s_out[2]);
      } else {
        st_ll_triple<LL_GPU>((char*)y.mbox + slot, e, s_out[0], s_out[1], s_out[2]);
      }
      if (tr) tr[3] = gtime();
    }
    __syncthreads();                 // s_out is reused for the totals below
  }
  if (tid < R) {
    const long long t0 = clock64();
    double k[3];
    const char* src = (const char*)y.mbox + (size_t)(par * CG_MAX_RANKS + tid) * LL_SLOT_BYTES;
    while (!ld_ll_triple<SCOPE>(src, e, k)) {
      if (clock64() - t0 > CG_SPIN_LIMIT) {
        atomicExch(y.err, 1ULL);
        break;
      }
    }
    if (tr && blockIdx.x == 0) tr[4 + (tid & 1)] = gtime();      // when block 0 saw rank tid's packet
    // acquire (an acquire load: LDG.STRONG + L1 invalidate; a full fence on the wait side cost ~4 us across GPUs):
    // the field data of this GPU's other CTAs and the peers' halo rows are visible to what follows
    if (SYS) (void)ld_acquire_sys((const unsigned long long*)src);
    else (void)ld_acquire_gpu((const unsigned long long*)src);
    s_peer[tid][0] = k[0]; s_peer[tid][1] = k[1]; s_peer[tid][2] = k[2];
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < 3; ++k) {    // rank order, identical everywhere
      double sum = 0.0;
      for (int q = 0; q < R; ++q) sum += s_peer[q][k];
      s_out[k] = sum;
    }
    s_out[3] = (ld_relaxed_u64(y.err) == 0ULL) ? 1.0 : 0.0;
    if (tr && blockIdx.x == 0) tr[6] = gtime();
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) out[k] = s_out[k];
  return s_out[3] != 0.0;
}

template <bool UNIT, bool MULTI>
// __launch_bounds__(544, 1) makes ptxas stop at 96 registers (a few bytes of spills in the general-spacing and multi-GPU
// instantiations).  Measured alternative, -DPF_TMAD_MAXNREG: __maxnreg__(112) removes the spills and is 7 % (4096^2) to
// 19 % (16384^2) SLOWER (profiles/r2_ab_maxnreg.txt).
#ifdef PF_TMAD_MAXNREG
__global__ void __maxnreg__(112) k_cg_tmad(CgArgs a) {
#else
__global__ void __launch_bounds__(TM_THREADS, 1) k_cg_tmad(CgArgs a) {
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double sm[3 * 32];
  __shared__ DynShared ds;
  const int G = gridDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = (warp == TM_WARPS);
  Ring rg;
  rg.base = smem_raw;
  rg.full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TM_STAGES * TM_STAGE_BYTES);
  rg.empty = rg.full + TM_STAGES;
  rg.base_at = smem_addr_pin(rg.base);
  rg.full_at = smem_addr_pin(rg.full);
  rg.empty_at = smem_addr_pin(rg.empty);
  if (tid == 0) {
    for (int q = 0; q < TM_STAGES; ++q) {
      mbar_init(rg.full + q, 1u);
      mbar_init(rg.empty + q, (uint32_t)TM_WARPS);
    }
    for (int q = 0; q < TD_SLOTS; ++q) { ds.tile[q] = -1; ds.count[q] = 0u; }
    ds.n_end = 0; ds.seq_end = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t pol = policy_evict_first();
  unsigned long long n = 0, seq = 0;                  // row / tile sequence numbers of this CTA
  const unsigned long long tbase0 = *((volatile unsigned long long*)a.y.ticket_base);
  const unsigned long long per_phase = tmad_tickets_per_phase(a.ntiles, G);
  Prefetch pf;                                        // producer lane only
  pf.tile = -1; pf.npref = 0; pf.phase_b = false;
  unsigned long long phase = 0;
  unsigned long long e = *((volatile unsigned long long*)a.y.epoch);   // cross-GPU barrier epochs (MULTI)
  const unsigned long long e0 = e;
  // strip kind of this warp is the same for every tile of a chunk COLUMN position, but chunks differ:
  // it depends on the tile, so it is resolved per phase call below from the warp's column only when
  // there is a single chunk; in general X_WALL applies to warp 0 of chunk 0 and warp 15 of the last chunk.

  const double bb = ld_cg(a.scal + S_BB);
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
  bool ok = true;

  if (bnorm == 0.0) {
    info = 0;
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) {
        info = 0;
        break;
      }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      // ---------------- phase A
      {
        bool good = true;
        const unsigned long long tbase = tbase0 + phase * per_phase;
        if (producer) {
          if (lane == 0) {
            good = tmad_produce<false>(a, rg, &ds, tbase, a.r, pcur, nullptr, first, pol, n, seq, pf);
            ds.n_end = n; ds.seq_end = seq;
            // phase B follows unconditionally: fetch x and r of its first visits while the barrier runs
            if (good) good = tmad_prefetch<true>(a, rg, &ds, pnext, a.x, a.r, pol, n, seq, pf);
          }
        } else {
          // the strip kind (interior / wall / cut by the last column / beyond it) is resolved per tile
          double* plo = (pnext == a.p0) ? a.y.below_p0 : a.y.below_p1;
          double* phi = (pnext == a.p0) ? a.y.above_p0 : a.y.above_p1;
          good = tmad_consume_A<UNIT, MULTI>(a, rg, &ds, warp, pnext, plo, phi, first, beta, n, seq, (unsigned)phase, ll_tag(tbase + 1ULL));
        }
        if (!good) atomicExch(a.y.err, 1ULL);
        ++phase;
      }
      __syncthreads();
      n = ds.n_end; seq = ds.seq_end;
      double pq;
      {
        double o1[1];
        ok = tile_exchange<1, MULTI>(a, ++e, o1, e0);
        if (!ok) break;
        pq = o1[0];
      }
      const double alpha = rho / pq;
      // ---------------- phase B
      {
        bool good = true;
        const unsigned long long tbase = tbase0 + phase * per_phase;
        if (producer) {
          if (lane == 0) {
            good = tmad_produce<true>(a, rg, &ds, tbase, pnext, a.x, a.r, false, pol, n, seq, pf);
            ds.n_end = n; ds.seq_end = seq;
            // phase A of the next iteration (if the solve goes on): its p_old is this iteration's new p
            if (good && it + 1 < a.maxiter) good = tmad_prefetch<false>(a, rg, &ds, a.r, pnext, nullptr, pol, n, seq, pf);
          }
        } else {
          good = tmad_consume_B<UNIT, MULTI>(a, rg, &ds, warp, a.x, a.r, a.y.below_r, a.y.above_r, alpha, pol, n, seq, (unsigned)phase, ll_tag(tbase + 1ULL));
        }
        if (!good) atomicExch(a.y.err, 1ULL);
        ++phase;
      }
      __syncthreads();
      n = ds.n_end; seq = ds.seq_end;
      double rz;
      {
        double o2[2];
        ok = tile_exchange<2, MULTI>(a, ++e, o2, e0);
        if (!ok) break;
        rr = o2[0];
        rz = o2[1];
      }
      rho_prev = rho;
      rho = rz;
      double* tmp = pcur; pcur = pnext; pnext = tmp;
      ++iters;
    }
  }
  // the solve stopped with a phase A reserved (converged, or an error): complete and retire those copies.  After the
  // swap at the end of the iteration pcur is the buffer the prefetch was reading from.
  // (An error exit in the middle of an iteration can leave a phase B reserved instead: p is pnext then.)
  if (producer && lane == 0) {
    if (pf.phase_b) tmad_drain<true>(a, rg, pnext, a.x, a.r, pol, n, pf);
    else tmad_drain<false>(a, rg, a.r, pcur, nullptr, pol, n, pf);
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.scal[S_INFO] = ok ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    *a.y.ticket_base = tbase0 + phase * per_phase;
    *a.y.epoch = e;
  }
}

}  // namespace pf
