// pf_cg_small.cuh -- CG for grids small enough to LIVE ON THE SMs (config 2: 512^2).
//
// At 512^2 a vector is 2 MiB: the streaming kernels run out of L2 and an iteration is
// two grid barriers plus two latency-bound sweeps (10.8 us).  Here every CTA (1 per SM,
// a band of rows x all columns) keeps its part of the problem on chip for the WHOLE
// solve: x, r, q and the CTA's own p cells live in REGISTERS (each thread owns the same
// NP column pairs from the first iteration to the last; their geometry -- shared-memory
// offset, Neumann / ring / pin predicates -- is worked out once), p with its two halo
// rows lives in shared memory because the stencil needs the neighbours.  Only the two
// boundary rows of r leave the SM per iteration -- the halo rows of p are recomputed
// redundantly from them with the owner's exact arithmetic, so no third barrier is needed.
// The two reductions per iteration (classical CG) are fence-free all-to-all exchanges of
// self-validating LL words (pf_common.cuh): every CTA posts its partials, warp 0 of every
// CTA polls all of them and adds them in CTA order (bit-identical everywhere); the
// boundary rows of r travel the same way, posted straight from the owners' registers.
// Same operation order per cell as pf_cg.cuh / scipy.
//
// Round-1 version kept x, r, p in shared memory and re-derived every cell's geometry
// (integer divisions, eight predicated pair slots) in every phase: ncu showed 986 warp
// instructions per warp per iteration and 35 % issue utilisation on a problem that is
// pure latency (profiles/r2_cg_small_v1_ncu.txt).
#pragma once

#include "pf_cg.cuh"

namespace pf {

constexpr int SM_THREADS = 512;
constexpr int SM_PAIRS = 8;                       // column pairs per thread, at most
constexpr int SM_MAX_SMEM = 200 * 1024;

struct SmallArgs {
  Slab s;
  const double* b;          // rhs (= initial r), device field layout
  double* x;                // solution out (field layout)
  double* rhalo;            // [G][2][nx/2][4]: first / last row of r of every CTA, one LL packet per column pair
  double* part;             // [2][SM_MAX_GRID] reduction packets (4 LL words each, SM_SLOT_DOUBLES apart), double-buffered by barrier parity
  double* scal;
  unsigned long long* epoch;    // persisted barrier epoch
  unsigned long long* err;      // != 0 after a spin-wait timeout
  double dx, dy, invd, tol;
  int maxiter;
  int rows;                 // rows per CTA (last CTA may own fewer)
  unsigned long long* trace;    // timeline of iteration 100 (pf_set_option cg_sync_opt bit 4), else null: [G][8] globaltimer stamps
};

// shared memory: p with halo rows [(rows+2)][nx] + the neighbours' boundary rows of r [2][nx]
__host__ __device__ __forceinline__ size_t small_smem_bytes(int nx, int rows) {
  return sizeof(double) * (size_t)nx * (size_t)(rows + 4);
}

constexpr int SM_POLL = 5;                        // reduce packets per lane of warp 0
constexpr int SM_MAX_GRID = 32 * SM_POLL;         // => at most 160 CTAs (B200: 148 SMs)
// Reduction packets are packed 32 bytes apart.  (Measured: spreading them 1 KiB apart, to put the G x G sector
// reads of a polling round on different L2 slices, made the 512^2 solve 9 % SLOWER -- 1.757 vs 1.607 ms -- so
// the L2 is not the bottleneck of the exchange; profiles/README.md.)
constexpr int SM_SLOT_DOUBLES = 4;

// packets per boundary row: one per column pair, 4 LL words = 32 bytes each
__host__ __device__ __forceinline__ int small_row_packets(int nx) { return nx / 2; }

// geometry bits of a thread's column pair
constexpr unsigned SMF_VALID = 1u, SMF_HS = 2u, SMF_HN = 4u, SMF_HW = 8u, SMF_HEB = 16u, SMF_PIN = 32u,
                   SMF_RING_A = 64u, SMF_RING_B = 128u, SMF_FIRST = 256u, SMF_LAST = 512u;

// All-reduce of NV (<= 2) doubles over the grid; barrier included; every thread returns the sums.
//
// No counter, no fence: every CTA publishes its partials as 4 LL words (self-validating, pf_common.cuh)
// into its slot, warp 0 of every CTA polls all slots -- its loads are in flight together -- and adds
// them in CTA order (same order everywhere => identical sums).  Slots are double-buffered by epoch
// parity: a fast CTA can post e+1 while a slow one still reads e, but cannot reach e+2 before the slow
// one posted e+1.
//
// HALO: warps 1.. fetch the neighbours' boundary rows of r (posted by their owners before this call)
// into rh_s while warp 0 runs the reduction.  A packet validates word by word, so no release / acquire
// fence is needed for the rows either; a neighbour overwrites its row packets only after the NEXT
// reduction, which needs this CTA's packet, which is posted after these reads.
template <int NV, bool HALO>
__device__ __forceinline__ void small_allreduce(const SmallArgs& a, unsigned long long e, double (&v)[NV], double* sm,
                                                double* s_res2, double* rh_s, unsigned long long* tr = nullptr) {
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bid = blockIdx.x;
  const int nx = a.s.nx;
  const unsigned tag = ll_tag(e);
  double* slots = a.part + (size_t)(e & 1ULL) * (size_t)SM_MAX_GRID * SM_SLOT_DOUBLES;
  double* s_res = s_res2 + (e & 1ULL) * 4;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm[k * 32 + warp] = v[k];
  }
  __syncthreads();                                     // warp partials are in shared memory
  if (warp == 0) {
    if (lane == 0) {
      double t[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < NV; ++k)
        for (int q = 0; q < SM_THREADS / 32; ++q) t[k] += sm[k * 32 + q];
      st_ll4<LL_GPU>(slots + (size_t)bid * SM_SLOT_DOUBLES, ll_encode(t[0], tag), ll_encode(t[1], tag));
      if (tr) tr[0] = gtime();                           // posted
    }
    double g0[SM_POLL], g1[SM_POLL];                   // decoded partials of this lane's slots
    unsigned pending = 0u;
#pragma unroll
    for (int j = 0; j < SM_POLL; ++j) {
      g0[j] = 0.0; g1[j] = 0.0;
      if (lane + 32 * j < G) pending |= 1u << j;
    }
    const long long t0 = clock64();
    while (pending) {
#pragma unroll
      for (int j = 0; j < SM_POLL; ++j) {                // the loads of one round are in flight together
        if (pending & (1u << j)) {
          LLVal wa, wb;
          ld_ll4<LL_GPU>(slots + (size_t)(lane + 32 * j) * SM_SLOT_DOUBLES, wa, wb);
          if (ll_ok(wa, tag) && ll_ok(wb, tag)) {
            g0[j] = ll_value(wa); g1[j] = ll_value(wb);
            pending &= ~(1u << j);
          }
        }
      }
      if (pending && clock64() - t0 > CG_SPIN_LIMIT) {   // never hang the GPU
        atomicExch(a.err, 1ULL);
        break;
      }
    }
    if (tr && lane == 0) tr[1] = gtime();                // every packet seen
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < SM_POLL; ++j) {                  // CTA order within a lane, fixed xor tree across lanes
      if (lane + 32 * j < G) {
        acc[0] += g0[j];
        if (NV > 1) acc[1] += g1[j];
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc[k] = warp_sum(acc[k]);
      if (lane == 0) s_res[k] = acc[k];
    }
  } else if (HALO) {
    const int hp = nx / 2;
    for (int idx = tid - 32; idx < 2 * hp; idx += SM_THREADS - 32) {           // fetch the neighbours' rows
      const int side = idx >= hp ? 1 : 0, k = idx - side * hp;
      const int nb = side == 0 ? bid - 1 : bid + 1;
      if (nb < 0 || nb >= G) continue;
      const double* src = a.rhalo + (((size_t)nb * 2 + (side == 0 ? 1 : 0)) * hp + (size_t)k) * 4;
      LLVal h0, h1;
      ld_ll4<LL_GPU>(src, h0, h1);
      const long long t0 = clock64();
      while (!(ll_ok(h0, tag) && ll_ok(h1, tag))) {
        if (clock64() - t0 > CG_SPIN_LIMIT) {
          atomicExch(a.err, 1ULL);
          break;
        }
        ld_ll4<LL_GPU>(src, h0, h1);
      }
      *reinterpret_cast<double2*>(rh_s + (size_t)side * nx + 2 * k) = make_double2(ll_value(h0), ll_value(h1));
    }
  }
  __syncthreads();                                     // sums (and halo rows) are in shared memory
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_res[k];
}

template <bool UNIT, int NP>
__global__ void __launch_bounds__(SM_THREADS, 1) k_cg_small(SmallArgs a) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sm[3 * 32];
  __shared__ double s_res[8];                       // sums, double-buffered by barrier parity
  const Slab& s = a.s;
  const int nx = s.nx, G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const int j0 = bid * a.rows;
  const int j1 = (j0 + a.rows > s.nyl) ? s.nyl : j0 + a.rows;
  const int R = j1 - j0;                              // >= 1 by construction
  // shared layout: p_s[(rows+2)][nx] (row 0 / R+1 = halo), rh_s[2][nx] (r of the rows below / above the band)
  double* p_s = smem_d;
  double* rh_s = p_s + (size_t)(a.rows + 2) * nx;
  const int hp = nx / 2;                              // column pairs per row
  const int npairs = R * hp;                          // own cells, in pairs
  const int jg_lo = s.row0 + j0 - 1, jg_hi = s.row0 + j1;       // global rows of the two halo rows
  const bool lo_exists = jg_lo >= 0, hi_exists = jg_hi < s.nyg;

  // ---- this thread's column pairs: geometry once, r = b, x = 0 in registers
  int off[NP];                                        // offset of the pair inside p_s
  unsigned flg[NP];
  double ra[NP], rb[NP], xa[NP], xb[NP], pa[NP], pb[NP], qa[NP], qb[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int idx = tid + k * SM_THREADS;
    off[k] = nx; flg[k] = 0u;
    ra[k] = rb[k] = xa[k] = xb[k] = pa[k] = pb[k] = qa[k] = qb[k] = 0.0;
    if (idx < npairs) {
      const int lr = idx / hp, c = (idx - lr * hp) * 2;
      const int jg = s.row0 + j0 + lr;
      off[k] = (lr + 1) * nx + c;
      unsigned f = SMF_VALID;
      if (jg > 0) f |= SMF_HS;
      if (jg < s.nyg - 1) f |= SMF_HN;
      if (c > 0) f |= SMF_HW;
      if (c + 1 < nx - 1) f |= SMF_HEB;
      if (jg == 0 && c == 0) f |= SMF_PIN;
      const bool ring_row = (jg == 0) || (jg == s.nyg - 1);
      if (ring_row || c == 0) f |= SMF_RING_A;
      if (ring_row || c + 1 == nx - 1) f |= SMF_RING_B;
      if (lr == 0) f |= SMF_FIRST;
      if (lr == R - 1) f |= SMF_LAST;
      flg[k] = f;
      const double2 bv = ld2(a.b + cell(s, j0 + lr, c));
      ra[k] = bv.x; rb[k] = bv.y;
    }
  }
  // halo rows: r of the neighbouring rows (from b), p = 0
  for (int idx = tid; idx < 2 * hp; idx += SM_THREADS) {
    const int side = idx >= hp ? 1 : 0, c = (idx - side * hp) * 2;
    const int jl = side == 0 ? j0 - 1 : j1;
    double2 v = make_double2(0.0, 0.0);
    if (jl >= 0 && jl < s.nyl) v = ld2(a.b + cell(s, jl, c));
    *reinterpret_cast<double2*>(rh_s + (size_t)side * nx + c) = v;
    *reinterpret_cast<double2*>(p_s + (size_t)(side == 0 ? 0 : R + 1) * nx + c) = make_double2(0.0, 0.0);
  }
  __syncthreads();

  const double bb = ld_cg(a.scal + S_BB);
  const double bnorm = sqrt(bb);
  const double t0 = a.tol * bnorm;
  const double atol = (t0 > 0.0) ? t0 : 0.0;
  int iters = 0, info = a.maxiter;
  double rr = bb, rho = ld_cg(a.scal + S_BZ), rho_prev = 0.0;
  unsigned long long e = *((volatile unsigned long long*)a.epoch);

  if (bnorm == 0.0) {
    info = 0;
  } else {
    for (int it = 0; it < a.maxiter; ++it) {
      if (sqrt(rr) < atol) { info = 0; break; }
      const bool first = (it == 0);
      const double beta = first ? 0.0 : rho / rho_prev;
      unsigned long long* tr = (a.trace && it == 100) ? a.trace + (size_t)bid * 8 : nullptr;
      if (tr && tid == 0) tr[0] = gtime();                // iteration start
      // ---- phase A.1: p = (p*beta) + M r on own cells (registers -> shared) AND on the two halo rows
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (flg[k] & SMF_VALID) {
          const double za = ((flg[k] & SMF_RING_A) ? 1.0 : a.invd) * ra[k];
          const double zb = ((flg[k] & SMF_RING_B) ? 1.0 : a.invd) * rb[k];
          if (first) { pa[k] = za; pb[k] = zb; }
          else { pa[k] = (pa[k] * beta) + za; pb[k] = (pb[k] * beta) + zb; }
          *reinterpret_cast<double2*>(p_s + off[k]) = make_double2(pa[k], pb[k]);
        }
      }
      for (int idx = tid; idx < 2 * hp; idx += SM_THREADS) {
        const int side = idx >= hp ? 1 : 0, c = (idx - side * hp) * 2;
        if (side == 0 ? !lo_exists : !hi_exists) continue;
        const int jg = side == 0 ? jg_lo : jg_hi;
        const bool ring_row = (jg == 0) || (jg == s.nyg - 1);
        const double2 rv = *reinterpret_cast<const double2*>(rh_s + (size_t)side * nx + c);
        const double za = ((ring_row || c == 0) ? 1.0 : a.invd) * rv.x;
        const double zb = ((ring_row || c + 1 == nx - 1) ? 1.0 : a.invd) * rv.y;
        double* ph = p_s + (size_t)(side == 0 ? 0 : R + 1) * nx + c;
        double2 pv;
        if (first) {
          pv = make_double2(za, zb);
        } else {
          const double2 po = *reinterpret_cast<const double2*>(ph);
          pv = make_double2((po.x * beta) + za, (po.y * beta) + zb);
        }
        *reinterpret_cast<double2*>(ph) = pv;
      }
      __syncthreads();
      // ---- phase A.2: q = A p on own cells (kept in registers), p.q
      double v1[1] = {0.0};
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (flg[k] & SMF_VALID) {
          const unsigned f = flg[k];
          const double* pc = p_s + off[k];
          const bool hs = f & SMF_HS, hn = f & SMF_HN, hw = f & SMF_HW, he_b = f & SMF_HEB;
          const double w = hw ? pc[-1] : 0.0, ee = he_b ? pc[2] : 0.0;
          double2 sv = make_double2(0.0, 0.0), nv = make_double2(0.0, 0.0);
          if (hs) sv = *reinterpret_cast<const double2*>(pc - nx);
          if (hn) nv = *reinterpret_cast<const double2*>(pc + nx);
          double q0 = neg_lap<UNIT>(pa[k], w, pb[k], sv.x, nv.x, hw, true, hs, hn, a.dx, a.dy);
          const double q1 = neg_lap<UNIT>(pb[k], pa[k], ee, sv.y, nv.y, true, he_b, hs, hn, a.dx, a.dy);
          if (f & SMF_PIN) q0 = pa[k];                  // pinned row
          qa[k] = q0; qb[k] = q1;
          v1[0] += pa[k] * q0;
          v1[0] += pb[k] * q1;
        }
      }
      if (tr && tid == 0) tr[1] = gtime();                // A.1 + A.2 done
      small_allreduce<1, false>(a, ++e, v1, sm, s_res, rh_s, tr ? tr + 2 : nullptr);      // tr[2] posted, tr[3] all seen
      if (tr && tid == 0) tr[4] = gtime();                // exchange 1 done
      const double alpha = rho / v1[0];
      // ---- phase B: x += alpha p ; r -= alpha q (registers) ; the owners of the band's first / last row
      //      post them for the neighbours, tagged with the epoch of the exchange that follows
      ++e;
      const unsigned tag = ll_tag(e);
      double v2[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        if (flg[k] & SMF_VALID) {
          const unsigned f = flg[k];
          xa[k] = xa[k] + alpha * pa[k];
          xb[k] = xb[k] + alpha * pb[k];
          ra[k] = ra[k] - alpha * qa[k];
          rb[k] = rb[k] - alpha * qb[k];
          if (f & (SMF_FIRST | SMF_LAST)) {
            const int c2 = (off[k] % nx) >> 1;          // column pair index (rare path: boundary rows only)
            const LLVal w0 = ll_encode(ra[k], tag), w1 = ll_encode(rb[k], tag);
            if (f & SMF_FIRST) st_ll4<LL_GPU>(a.rhalo + (((size_t)bid * 2 + 0) * hp + c2) * 4, w0, w1);
            if (f & SMF_LAST) st_ll4<LL_GPU>(a.rhalo + (((size_t)bid * 2 + 1) * hp + c2) * 4, w0, w1);
          }
          v2[0] += ra[k] * ra[k];
          v2[1] += ra[k] * (((f & SMF_RING_A) ? 1.0 : a.invd) * ra[k]);
          v2[0] += rb[k] * rb[k];
          v2[1] += rb[k] * (((f & SMF_RING_B) ? 1.0 : a.invd) * rb[k]);
        }
      }
      if (tr && tid == 0) tr[5] = gtime();                // phase B done
      small_allreduce<2, true>(a, e, v2, sm, s_res, rh_s, tr ? tr + 6 : nullptr);   // also fetches the neighbours' boundary rows of r; tr[6] posted, tr[7] all seen
      rr = v2[0];
      rho_prev = rho;
      rho = v2[1];
      ++iters;
    }
  }
  // ---- write x back
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    if (flg[k] & SMF_VALID) {
      const int lr = off[k] / nx - 1, c = off[k] % nx;
      st2(a.x + cell(s, j0 + lr, c), xa[k], xb[k]);
    }
  }
  if (bid == 0 && tid == 0) {
    a.scal[S_INFO] = (ld_relaxed_u64(a.err) == 0ULL) ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    *a.epoch = e;
  }
}

}  // namespace pf
