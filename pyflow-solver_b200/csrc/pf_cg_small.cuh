// pf_cg_small.cuh -- CG for grids small enough to LIVE IN SHARED MEMORY (config 2: 512^2).
//
// At 512^2 a vector is 2 MiB: the streaming kernels run out of L2 and an iteration is
// two grid barriers plus two latency-bound sweeps (10.8 us).  Here x, r and p stay in
// the shared memory of the CTA that owns their rows for the WHOLE solve (1 CTA per SM,
// a band of rows x all columns each); A p is a pure shared-memory stencil; q stays in
// registers between the two phases; only the two boundary rows of r travel through
// global memory (L2) per iteration -- the halo rows of p are recomputed redundantly
// from them with the owner's exact arithmetic, so no third barrier is needed.  The
// two reductions per iteration (classical CG) are fence-free: every CTA posts one
// self-validating 32-byte packet {epoch, partials}, every CTA's warp 0 polls all of
// them and adds them in CTA order (bit-identical everywhere); the boundary rows of r
// travel as the same kind of tagged packets.  Same operation order as pf_cg.cuh / scipy.
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
  double* rhalo;            // [G][2][packets per row][4]: first / last row of r of every CTA as tagged packets
  double* part;             // [2][G][4] per-CTA reduction packets, double-buffered by barrier parity
  double* scal;
  unsigned long long* epoch;    // persisted barrier epoch
  unsigned long long* err;      // != 0 after a spin-wait timeout
  double dx, dy, invd, tol;
  int maxiter;
  int rows;                 // rows per CTA (last CTA may own fewer)
};

__device__ __forceinline__ int small_smem_bytes(int nx, int rows) {
  return (int)sizeof(double) * nx * ((rows + 2) * 2 + rows);
}

constexpr int SM_POLL = 5;                        // reduce packets per lane of warp 0
constexpr int SM_MAX_GRID = 32 * SM_POLL;         // => at most 160 CTAs (B200: 148 SMs)

// packets per boundary row: {epoch, 3 doubles} each
__host__ __device__ __forceinline__ int small_row_packets(int nx) { return (nx + 2) / 3; }

// All-reduce of NV doubles over the grid; barrier included; every thread returns the sums.
//
// No counter, no fence: every CTA publishes ONE self-validating 32-byte packet {epoch, partials}
// (a single 256-bit store) into its slot, warp 0 of every CTA polls all slots -- its loads are in
// flight together -- and adds them in CTA order (same order everywhere => identical sums).
// Slots are double-buffered by epoch parity: a fast CTA can post e+1 while a slow one still reads
// e, but cannot reach e+2 before the slow one posted e+1.
//
// HALO: the CTA also publishes its first / last row of r (a.r_s rows 1 and R) as tagged packets, and
// warps 1.. fetch the neighbours' rows into the halo rows of r_s while warp 0 runs the reduction.
// A packet is complete or absent, so no release / acquire fence is needed for the rows either; a
// neighbour overwrites its row packets only after the NEXT reduction, which needs this CTA's packet,
// which is posted after these reads.
template <int NV, bool HALO>
__device__ __forceinline__ void small_allreduce(const SmallArgs& a, unsigned long long e, double (&v)[NV], double* sm,
                                                double* s_res2, double* r_s, int R) {
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, bid = blockIdx.x;
  const int nx = a.s.nx;
  double* slots = a.part + (size_t)(e & 1ULL) * (size_t)G * 4;
  double* s_res = s_res2 + (e & 1ULL) * 4;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sm[k * 32 + warp] = v[k];
  }
  __syncthreads();                                     // warp partials and (HALO) the CTA's new r rows are in shared memory
  if (warp == 0) {
    if (lane == 0) {
      double t[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < NV; ++k)
        for (int q = 0; q < SM_THREADS / 32; ++q) t[k] += sm[k * 32 + q];
      st_packet_gpu(slots + (size_t)bid * 4, e, t[0], t[1], 0.0);
    }
    Packet pk[SM_POLL];
    const long long t0 = clock64();
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int j = 0; j < SM_POLL; ++j) {
        const int q = lane + 32 * j;
        if (q < G) {
          pk[j] = ld_packet_gpu(slots + (size_t)q * 4);
          ok = ok && (pk[j].epoch == e);
        }
      }
      if (!ok && clock64() - t0 > CG_SPIN_LIMIT) {       // never hang the GPU
        atomicExch(a.err, 1ULL);
        break;
      }
    } while (!ok);
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
#pragma unroll
    for (int j = 0; j < SM_POLL; ++j) {                  // CTA order within a lane, fixed xor tree across lanes
      if (lane + 32 * j < G) {
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += pk[j].v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      acc[k] = warp_sum(acc[k]);
      if (lane == 0) s_res[k] = acc[k];
    }
  } else if (HALO) {
    const int npk = small_row_packets(nx);
    const size_t hstride = (size_t)npk * 4;
    for (int idx = tid - 32; idx < 2 * npk; idx += SM_THREADS - 32) {          // publish own boundary rows
      const int side = idx / npk, k = idx - side * npk, c = 3 * k;
      const double* row = r_s + (size_t)(side == 0 ? 1 : R) * nx;
      st_packet_gpu(a.rhalo + ((size_t)bid * 2 + side) * hstride + (size_t)k * 4, e, row[c],
                    c + 1 < nx ? row[c + 1] : 0.0, c + 2 < nx ? row[c + 2] : 0.0);
    }
    for (int idx = tid - 32; idx < 2 * npk; idx += SM_THREADS - 32) {          // fetch the neighbours' rows
      const int side = idx / npk, k = idx - side * npk, c = 3 * k;
      const int nb = side == 0 ? bid - 1 : bid + 1;
      if (nb < 0 || nb >= G) continue;
      const double* src = a.rhalo + ((size_t)nb * 2 + (side == 0 ? 1 : 0)) * hstride + (size_t)k * 4;
      Packet h = ld_packet_gpu(src);
      const long long t0 = clock64();
      while (h.epoch != e) {
        if (clock64() - t0 > CG_SPIN_LIMIT) {
          atomicExch(a.err, 1ULL);
          break;
        }
        h = ld_packet_gpu(src);
      }
      double* dst = r_s + (size_t)(side == 0 ? 0 : R + 1) * nx + c;
      dst[0] = h.v[0];
      if (c + 1 < nx) dst[1] = h.v[1];
      if (c + 2 < nx) dst[2] = h.v[2];
    }
  }
  __syncthreads();                                     // sums (and halo rows) are in shared memory
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = s_res[k];
}

template <bool UNIT>
__global__ void __launch_bounds__(SM_THREADS, 1) k_cg_small(SmallArgs a) {
  extern __shared__ __align__(16) double smem_d[];
  __shared__ double sm[3 * 32];
  __shared__ double s_res[8];                       // sums, double-buffered by barrier parity
  const Slab& s = a.s;
  const int nx = s.nx, G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  const int j0 = bid * a.rows;
  const int j1 = (j0 + a.rows > s.nyl) ? s.nyl : j0 + a.rows;
  const int R = j1 - j0;                              // >= 1 by construction
  // shared layout: r_s[(rows+2)][nx], p_s[(rows+2)][nx] (row 0 / R+1 = halo), x_s[rows][nx]
  double* r_s = smem_d;
  double* p_s = r_s + (size_t)(a.rows + 2) * nx;
  double* x_s = p_s + (size_t)(a.rows + 2) * nx;
  const int hp = nx / 2;                              // column pairs per row
  const int npairs = R * hp;                          // own cells, in pairs
  const int hpairs = (R + 2) * hp;                    // incl. the two halo rows

  auto invd_of = [&](int jg, int i) {
    const bool ring = (i == 0) || (i == nx - 1) || (jg == 0) || (jg == s.nyg - 1);
    return ring ? 1.0 : a.invd;
  };

  // ---- load r = b (own rows + halo rows), x = 0, p = 0
  for (int idx = tid; idx < hpairs; idx += SM_THREADS) {
    const int lr = idx / hp, c = (idx % hp) * 2;      // lr: 0 = halo below, 1..R own, R+1 = halo above
    const int jl = j0 + lr - 1;
    double2 v = make_double2(0.0, 0.0);
    if (jl >= 0 && jl < s.nyl) v = ld2(a.b + cell(s, jl, c));
    *reinterpret_cast<double2*>(r_s + (size_t)lr * nx + c) = v;
    *reinterpret_cast<double2*>(p_s + (size_t)lr * nx + c) = make_double2(0.0, 0.0);
  }
  for (int idx = tid; idx < npairs; idx += SM_THREADS)
    *reinterpret_cast<double2*>(x_s + (size_t)(idx / hp) * nx + (idx % hp) * 2) = make_double2(0.0, 0.0);
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
      // ---- phase A.1: p = (p*beta) + M r on own rows AND the two halo rows
      for (int idx = tid; idx < hpairs; idx += SM_THREADS) {
        const int lr = idx / hp, c = (idx % hp) * 2;
        const int jg = s.row0 + j0 + lr - 1;
        if (jg < 0 || jg >= s.nyg) continue;
        const size_t o = (size_t)lr * nx + c;
        const double2 rv = *reinterpret_cast<const double2*>(r_s + o);
        const double za = invd_of(jg, c) * rv.x, zb = invd_of(jg, c + 1) * rv.y;
        double2 pv;
        if (first) {
          pv = make_double2(za, zb);
        } else {
          const double2 po = *reinterpret_cast<const double2*>(p_s + o);
          pv = make_double2((po.x * beta) + za, (po.y * beta) + zb);
        }
        *reinterpret_cast<double2*>(p_s + o) = pv;
      }
      __syncthreads();
      // ---- phase A.2: q = A p on own rows (kept in registers), p.q
      double qa[SM_PAIRS], qb[SM_PAIRS];
      double v1[1] = {0.0};
#pragma unroll
      for (int k = 0; k < SM_PAIRS; ++k) {
        const int idx = tid + k * SM_THREADS;
        qa[k] = 0.0; qb[k] = 0.0;
        if (idx < npairs) {
          const int lr = idx / hp + 1, c = (idx % hp) * 2;
          const int jg = s.row0 + j0 + lr - 1;
          const double* pc = p_s + (size_t)lr * nx + c;
          const bool hs = jg > 0, hn = jg < s.nyg - 1;
          const bool hw = c > 0, he_b = c + 1 < nx - 1;
          const double pa = pc[0], pb = pc[1];
          const double w = hw ? pc[-1] : 0.0, ee = he_b ? pc[2] : 0.0;
          const double sa = hs ? pc[-nx] : 0.0, sb = hs ? pc[-nx + 1] : 0.0;
          const double na = hn ? pc[nx] : 0.0, nb = hn ? pc[nx + 1] : 0.0;
          double q0 = neg_lap<UNIT>(pa, w, pb, sa, na, hw, true, hs, hn, a.dx, a.dy);
          double q1 = neg_lap<UNIT>(pb, pa, ee, sb, nb, true, he_b, hs, hn, a.dx, a.dy);
          if (jg == 0 && c == 0) q0 = pa;               // pinned row
          qa[k] = q0; qb[k] = q1;
          v1[0] += pa * q0;
          v1[0] += pb * q1;
        }
      }
      small_allreduce<1, false>(a, ++e, v1, sm, s_res, r_s, R);
      const double alpha = rho / v1[0];
      // ---- phase B: x += alpha p ; r -= alpha q ; publish the boundary rows of r
      double v2[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < SM_PAIRS; ++k) {
        const int idx = tid + k * SM_THREADS;
        if (idx < npairs) {
          const int lr = idx / hp + 1, c = (idx % hp) * 2;
          const int jg = s.row0 + j0 + lr - 1;
          const size_t o = (size_t)lr * nx + c;
          const double2 pv = *reinterpret_cast<const double2*>(p_s + o);
          double2 xv = *reinterpret_cast<double2*>(x_s + (size_t)(lr - 1) * nx + c);
          double2 rv = *reinterpret_cast<double2*>(r_s + o);
          xv.x = xv.x + alpha * pv.x;
          xv.y = xv.y + alpha * pv.y;
          rv.x = rv.x - alpha * qa[k];
          rv.y = rv.y - alpha * qb[k];
          *reinterpret_cast<double2*>(x_s + (size_t)(lr - 1) * nx + c) = xv;
          *reinterpret_cast<double2*>(r_s + o) = rv;
          v2[0] += rv.x * rv.x;
          v2[1] += rv.x * (invd_of(jg, c) * rv.x);
          v2[0] += rv.y * rv.y;
          v2[1] += rv.y * (invd_of(jg, c + 1) * rv.y);
        }
      }
      small_allreduce<2, true>(a, ++e, v2, sm, s_res, r_s, R);   // also exchanges the boundary rows of r
      rr = v2[0];
      rho_prev = rho;
      rho = v2[1];
      ++iters;
    }
  }
  // ---- write x back
  for (int idx = tid; idx < npairs; idx += SM_THREADS) {
    const int lr = idx / hp, c = (idx % hp) * 2;
    const double2 xv = *reinterpret_cast<const double2*>(x_s + (size_t)lr * nx + c);
    st2(a.x + cell(s, j0 + lr, c), xv.x, xv.y);
  }
  if (bid == 0 && tid == 0) {
    a.scal[S_INFO] = (ld_relaxed_u64(a.err) == 0ULL) ? (double)info : -2.0;
    a.scal[S_ITERS] = (double)iters;
    a.scal[S_RR] = rr;
    *a.epoch = e;
  }
}

}  // namespace pf
