// pf_api.cu -- C-ABI of libpyflow_b200 (see include/pyflow_b200.h).
//
// Host side of the device step: owns the context, carves the caller's
// workspace into fields, sequences the kernels of one projection step on the
// caller's stream and returns the step scalars with one small D2H.
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "pf_cg.cuh"
#include "pf_cg_generic.cuh"
#include "pf_bicgstab.cuh"
#include "pf_cg_small.cuh"
#include "pf_cg_tma.cuh"
#include "pf_comm.h"
#include "pf_kernels.cuh"
#include "pf_step_fused.cuh"

using namespace pf;

struct pf_ctx {
  int device = 0;
  int sm_count = 0;
  Slab s{};
  size_t field_elems = 0;
  // fields (device): velocity ping-pong, rhs, CG vectors
  double *u = nullptr, *v = nullptr, *uo = nullptr, *vo = nullptr;
  double *b = nullptr, *x = nullptr, *r = nullptr, *p0 = nullptr, *p1 = nullptr;
  // small device buffers
  double* red_part = nullptr;
  unsigned* red_counter = nullptr;
  double* scal = nullptr;
  double* cg_part = nullptr;
  double* center = nullptr;
  // pinned host mirrors
  double* h_scal = nullptr;
  double* h_center = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int cg_grid = 0;
  int gen_grid = 0;                // co-resident CTAs of k_cg_generic
  int tma_grid = 0;                // co-resident CTAs of k_cg_tma (0: unavailable)
  bool small_ok = false;           // k_cg_small usable (dynamic shared memory attribute set)
  double* small_halo = nullptr;    // [SM_MAX_GRID][2][nx/2][4] boundary rows of r as LL packets, one per column pair (k_cg_small)
  double* small_part = nullptr;    // [2][SM_MAX_GRID] reduction packets, SM_SLOT_DOUBLES apart, double-buffered by epoch parity
  long long launches = 0;          // kernels of this library launched so far
  bool center_valid = false;       // h_center holds u[:, nx/2] of the current device u
  dim3 sgrid;                      // grid of the simple kernels
  dim3 fgrid;                      // grid of the two fused per-step passes (pf_step_fused.cuh); x = 0: not usable
  int fthreads = 0;                //   threads per block
  int fband = 0;                   //   rows per band
  pf::Comm* comm = nullptr;        // multi-GPU slabs (null when nranks == 1)
  char* ws_base = nullptr;         // device workspace (field k lives at k * field_stride)
  size_t field_stride = 0;
  size_t sync_off = 0;             // byte offset of the cross-GPU sync block inside the workspace
  char* peer_ws[pf::CG_MAX_RANKS] = {nullptr};   // the other ranks' workspaces mapped into this process
  bool peers = false;
  int opt_multi_single = 0;        // debug/tuning: run the multi-GPU persistent kernel with one rank
  int opt_sync = 0;
  int opt_rows_min = 4;            // minimum rows per warp tile of the register-window kernel
  int opt_tile_rows = 0;           // TMA kernel: force the big-tile height (0: automatic)
  int opt_small_rows = -1;         // TMA kernel: force the small-tile height (-1: automatic, 0: uniform tiles)
  int opt_nb_small = -1;           // TMA kernel: force the number of small bands at each end (-1: automatic)
  int opt_step_fused = 1;          // 1: one GPU steps with the two fused passes k_front / k_finish; 0: the seven separate kernels
  int opt_kernel = 3;              // 3: auto = shared-memory-resident kernel for small grids, TMA-fed dynamically scheduled
                                   //    kernel when nx % 1024 == 0, else the register-window kernels; 2: static TMA tiles;
                                   //    1: register-window only
  char err[512] = {0};
};

namespace {

constexpr size_t ALIGN = 256;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int fail(pf_ctx* c, int code, const char* what, const char* detail) {
  if (c) snprintf(c->err, sizeof(c->err), "%s: %s", what, detail ? detail : "");
  return code;
}

#define PF_CUDA(ctx, call)                                                         \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) return fail((ctx), PF_ERR_CUDA, #call, cudaGetErrorString(e__)); \
  } while (0)

// after every kernel launch: surface launch errors, count the launch
#define PF_LAUNCHED(ctx)                      \
  do {                                        \
    PF_CUDA((ctx), cudaGetLastError());       \
    (ctx)->launches++;                        \
  } while (0)

#define PF_TRY(expr)                \
  do {                              \
    int rc__ = (expr);              \
    if (rc__ != PF_OK) return rc__; \
  } while (0)

int pitch_of(int64_t nx) { return (int)align_up((size_t)nx, PITCH_ALIGN); }
size_t field_elems_of(int64_t nx, int64_t nyl) { return (size_t)(nyl + 2 * HALO) * (size_t)pitch_of(nx); }

struct SmallLayout {
  size_t red_part, red_counter, scal, cg_part, sync, center, small_halo, small_part, total;
};
SmallLayout small_layout(int64_t nyl, int64_t nx = 0) {
  SmallLayout L{};
  size_t o = 0;
  L.red_part = o;    o = align_up(o + sizeof(double) * MAX_RED_VALS * MAX_RED_BLOCKS, ALIGN);
  L.red_counter = o; o = align_up(o + sizeof(unsigned) * 4, ALIGN);
  L.scal = o;        o = align_up(o + sizeof(double) * S_COUNT, ALIGN);
  L.cg_part = o;     o = align_up(o + sizeof(double) * PF_PART_DOUBLES, ALIGN);
  L.sync = o;        o = align_up(o + 2048, ALIGN);
  L.center = o;      o = align_up(o + sizeof(double) * (size_t)nyl, ALIGN);
  L.small_halo = o;  o = align_up(o + sizeof(double) * 2 * SM_MAX_GRID * (size_t)(nx <= 2048 ? 4 * small_row_packets((int)nx) : 0), ALIGN);
  L.small_part = o;  o = align_up(o + sizeof(double) * 2 * SM_MAX_GRID * SM_SLOT_DOUBLES, ALIGN);
  L.total = o;
  return L;
}

constexpr int NFIELDS = 9;

dim3 simple_grid(const Slab& s) {
  const int gx = (s.nx + KT - 1) / KT;
  int gy = std::max(1, 2368 / gx);
  gy = std::min(gy, s.nyl);
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

// Grid of the fused per-step passes (pf_step_fused.cuh).
void fused_geometry(pf_ctx* c) {
  const FusedGeometry g = pf::fused_geometry(c->s.nx, c->s.nyl, c->sm_count);
  c->fgrid = dim3((unsigned)g.gx, (unsigned)(g.gy > 0 ? g.gy : 1), 1);      // gx = 0: not usable
  c->fthreads = g.threads;
  c->fband = g.rows;
}

bool is_unit(double dx, double dy) { return dx == 1.0 && dy == 1.0; }

// Geometry of the CG tiles for this slab and grid size (see pf_cg.cuh).
void cg_geometry(const pf_ctx* c, CgArgs& a) {
  const Slab& s = c->s;
  const int nstrips = (s.nx + STRIP - 1) / STRIP;
  int wx = 1;
  while (wx * 2 <= CG_WARPS && wx * 2 <= nstrips) wx *= 2;
  a.wx = wx;
  a.wy = CG_WARPS / wx;
  a.nchunks = (nstrips + wx - 1) / wx;
  const int G = c->cg_grid;
  // Bands: aim for one tile per CTA (a multiple of the grid when the slab is
  // tall enough), at least ~4 rows per warp so halo rows stay a small share.
  const int rows_min = c->opt_rows_min * a.wy;
  const int max_bands = std::max(1, s.nyl / rows_min);
  int want = std::max(1, G / a.nchunks);
  a.nbands = std::min(want, max_bands);
  a.ntiles = a.nchunks * a.nbands;
}

// Sync block layout (2 KiB, zeroed at creation): u64 arrive, release, epoch, err, tickets, ticket_base, pad, small-kernel
// epoch, flags[8]; then the mailbox ([2][CG_MAX_RANKS] 64-byte LL slots = 1 KiB) and one 64-byte LL slot of folded sums.
void fill_sync(const pf_ctx* c, CgSync& y) {
  auto u64 = [](char* p, int i) { return (unsigned long long*)p + i; };
  char* mine = c->ws_base + c->sync_off;
  y.arrive = u64(mine, 0);
  y.release = u64(mine, 1);
  y.epoch = u64(mine, 2);
  y.err = u64(mine, 3);
  y.tickets = u64(mine, 4);
  y.ticket_base = u64(mine, 5);
  y.flags = u64(mine, 8);
  y.mbox = (double*)(mine + 128);
  y.gsum = (double*)(mine + 128 + 2 * CG_MAX_RANKS * LL_SLOT_BYTES);
  y.opt = c->opt_sync;
  y.rank = c->comm ? comm_rank(c->comm) : 0;
  y.nranks = c->comm ? comm_nranks(c->comm) : 1;
  for (int q = 0; q < CG_MAX_RANKS; ++q) {
    y.peer_flags[q] = nullptr;
    y.peer_mbox[q] = nullptr;
    if (q < y.nranks) {
      char* theirs = (c->comm ? c->peer_ws[q] : c->ws_base) + c->sync_off;
      y.peer_flags[q] = u64(theirs, 8);
      y.peer_mbox[q] = (double*)(theirs + 128);
    }
  }
  const Slab& s = c->s;
  auto field = [&](int q, int k, int jl) { return (double*)(c->peer_ws[q] + (size_t)k * c->field_stride) + cell(s, jl, 0); };
  const bool has_below = y.rank > 0, has_above = y.rank < y.nranks - 1;
  // field indices in the workspace: r = 6, p0 = 7, p1 = 8 (see pf_create)
  y.below_r = has_below ? field(y.rank - 1, 6, s.nyl) : nullptr;
  y.below_p0 = has_below ? field(y.rank - 1, 7, s.nyl) : nullptr;
  y.below_p1 = has_below ? field(y.rank - 1, 8, s.nyl) : nullptr;
  y.above_r = has_above ? field(y.rank + 1, 6, -1) : nullptr;
  y.above_p0 = has_above ? field(y.rank + 1, 7, -1) : nullptr;
  y.above_p1 = has_above ? field(y.rank + 1, 8, -1) : nullptr;
}

int reset_trace(pf_ctx* c, cudaStream_t st) {
  if (!(c->opt_sync & 4)) return PF_OK;   // barrier timeline: min slots start at ~0ULL, the rest at 0
  PF_CUDA(c, cudaMemsetAsync(c->cg_part + PF_TRACE_OFFSET, 0, 64 * sizeof(double), st));
  for (int k = 0; k < 8; ++k)
    PF_CUDA(c, cudaMemsetAsync(c->cg_part + PF_TRACE_OFFSET + 8 * k, 0xff, sizeof(double), st));
  return PF_OK;
}

// k_cg_small instantiation for `np` column pairs per thread (1, 2, 4 or 8)
const void* small_kernel(bool unit, int np) {
  switch (np) {
    case 1: return unit ? (const void*)k_cg_small<true, 1> : (const void*)k_cg_small<false, 1>;
    case 2: return unit ? (const void*)k_cg_small<true, 2> : (const void*)k_cg_small<false, 2>;
    case 4: return unit ? (const void*)k_cg_small<true, 4> : (const void*)k_cg_small<false, 4>;
    default: return unit ? (const void*)k_cg_small<true, 8> : (const void*)k_cg_small<false, 8>;
  }
}

template <bool UNIT>
int launch_cg(pf_ctx* c, CgArgs& a, cudaStream_t st) {
  void* args[] = {&a};
  // small grids have fewer tiles than co-resident CTAs: launch only as many CTAs as there are tiles, so
  // that the two grid barriers and partial folds per iteration involve no idle CTAs
  const int grid = std::max(1, std::min(c->cg_grid, a.ntiles));
  PF_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_cg_persistent<UNIT>, dim3(grid), dim3(CG_THREADS),
                                         args, 0, st));
  c->launches++;
  return PF_OK;
}

// rhs already in c->b (raw); finishes it, solves, leaves x in c->x and the true
// residual in S_RES_SUMSQ.
int solve_on_device(pf_ctx* c, double dx, double dy, int jacobi, double invd_interior, double tol, int maxiter,
                    int subtract_mean, int pin, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, bool with_resid = true) {
  const Slab& s = c->s;
  const Red red{c->red_part, c->red_counter, c->scal};
  const double invd = jacobi ? invd_interior : 1.0;
  const double ncells = (double)s.nx * (double)s.nyg;
  if (c->comm && subtract_mean) PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_RHS_SUM, 1, /*max=*/0, st, c->err));
  k_rhs2<<<c->sgrid, KT, 0, st>>>(s, c->b, c->x, c->r, ncells, subtract_mean, pin, invd, red);
  PF_LAUNCHED(c);
  if (e0) PF_CUDA(c, cudaEventRecord(e0, st));
  CgArgs a{};
  a.s = s; a.b = c->b; a.x = c->x; a.r = c->r; a.p0 = c->p0; a.p1 = c->p1;
  a.part = c->cg_part; a.scal = c->scal;
  a.dx = dx; a.dy = dy; a.invd = invd; a.tol = tol; a.maxiter = maxiter;
  cg_geometry(c, a);
  const bool unit = is_unit(dx, dy);
  const int small_rows = (s.nyl + c->sm_count - 1) / c->sm_count;
  const size_t small_smem = small_smem_bytes(s.nx, small_rows);
  const long long small_pairs = (long long)small_rows * (s.nx / 2);          // column pairs per CTA
  if (c->opt_kernel == 3 && !c->comm && c->small_ok && s.nx % 2 == 0 && s.nx <= 2048 && c->sm_count <= SM_MAX_GRID &&
      small_pairs <= (long long)SM_PAIRS * SM_THREADS && small_smem <= (size_t)SM_MAX_SMEM) {
    // on-chip solve for small grids (config 2): x, r, q, p in registers, p + halo rows in shared memory
    SmallArgs sa{};
    sa.s = s; sa.b = c->b; sa.x = c->x; sa.rhalo = c->small_halo; sa.part = c->small_part; sa.scal = c->scal;
    sa.epoch = (unsigned long long*)(c->ws_base + c->sync_off) + 7;
    sa.err = (unsigned long long*)(c->ws_base + c->sync_off) + 3;
    sa.dx = dx; sa.dy = dy; sa.invd = invd; sa.tol = tol; sa.maxiter = maxiter; sa.rows = small_rows;
    sa.trace = (c->opt_sync & 4) ? (unsigned long long*)c->cg_part : nullptr;      // [G][8] stamps of iteration 100
    PF_CUDA(c, cudaMemsetAsync(sa.err, 0, sizeof(unsigned long long), st));
    const int grid = (s.nyl + small_rows - 1) / small_rows;
    void* args[] = {&sa};
    const int np = small_pairs <= SM_THREADS ? 1 : small_pairs <= 2 * SM_THREADS ? 2 : small_pairs <= 4 * SM_THREADS ? 4 : 8;
    const void* fn = small_kernel(unit, np);
    PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(SM_THREADS), args, small_smem, st));
    c->launches++;
  } else if (c->opt_kernel == 3 && (!c->comm || c->peers) && c->tma_grid > 0 && s.nx >= TM_CW && s.nyl >= 2) {
    // TMA-fed sweeps, dynamically scheduled 1024-column tiles (the last column chunk narrower when nx is not a
    // multiple of 1024; any nx >= 1024, odd ones too), one CTA per SM; with several ranks: halo rows + dot-product
    // partials over NVLink peer memory inside the same kernel
    // tile geometry: 16-row tiles (taller on huge slabs, shorter on short ones) with a few bands of quarter-height
    // tiles at both ends of the slab, so that both sweeps end on short tiles (pf_cg_tma.cuh)
    tmad_geometry(a, c->tma_grid, c->opt_tile_rows, c->opt_small_rows, c->opt_nb_small);
    PF_CUDA(c, cudaMemsetAsync(c->cg_part + PF_GCOUNT_OFFSET, 0, sizeof(unsigned) * CG_MAX_GROUPS, st));
    if (c->comm) {
      PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_BB, 2, 0, st, c->err));
      PF_TRY(comm_halo(c->comm, c->r, 1, 1, st, c->err));
    }
    fill_sync(c, a.y);
    PF_CUDA(c, cudaMemsetAsync(a.y.err, 0, sizeof(unsigned long long), st));
    PF_TRY(reset_trace(c, st));
    void* args[] = {&a};
    const void* fn = c->comm ? (unit ? (const void*)k_cg_tmad<true, true> : (const void*)k_cg_tmad<false, true>)
                             : (unit ? (const void*)k_cg_tmad<true, false> : (const void*)k_cg_tmad<false, false>);
    PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(c->tma_grid), dim3(TM_THREADS), args, TM_SMEM_BYTES, st));
    c->launches++;
  } else if (c->opt_kernel == 2 && !c->comm && c->tma_grid > 0 && s.nx % TM_CW == 0 &&
      s.nyl >= std::max(1, c->tma_grid / (s.nx / TM_CW))) {
    // TMA-fed sweeps: 1024-column tiles, one CTA per SM, rows streamed through a shared-memory ring
    a.nchunks = s.nx / TM_CW;
    a.nbands = std::max(1, c->tma_grid / a.nchunks);
    a.ntiles = a.nchunks * a.nbands;
    a.wx = TM_WARPS; a.wy = 1;
    a.y.opt = c->opt_sync;
    void* args[] = {&a};
    const void* fn = unit ? (const void*)k_cg_tma<true> : (const void*)k_cg_tma<false>;
    PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(c->tma_grid), dim3(TM_THREADS), args, TM_SMEM_BYTES, st));
    c->launches++;
  } else if (c->comm && c->peers) {
    // static-tile persistent kernel per rank over NVLink peer memory (cross-check)
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_BB, 2, 0, st, c->err));
    PF_TRY(comm_halo(c->comm, c->r, 1, 1, st, c->err));
    fill_sync(c, a.y);
    PF_CUDA(c, cudaMemsetAsync(a.y.err, 0, sizeof(unsigned long long), st));
    PF_TRY(reset_trace(c, st));
    void* args[] = {&a};
    const void* fn = unit ? (const void*)k_cg_persistent_multi<true> : (const void*)k_cg_persistent_multi<false>;
    PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(c->cg_grid), dim3(CG_THREADS), args, 0, st));
    c->launches++;
  } else if (c->comm) {
    // split-phase CG: halo rows + scalar all-reduces between the phases (NCCL, same stream)
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_BB, 2, 0, st, c->err));
    PF_TRY(comm_halo(c->comm, c->r, 1, 1, st, c->err));
    k_cg_begin<<<1, 32, 0, st>>>(c->scal, tol);
    PF_LAUNCHED(c);
    unsigned* counter = c->red_counter + 1;
    const int poll = 50;
    for (int it = 0; it < maxiter; ++it) {
      double* pnext = (it & 1) ? c->p0 : c->p1;
      if (unit) k_cg_phaseA<true><<<c->cg_grid, CG_THREADS, 0, st>>>(a, it, counter);
      else k_cg_phaseA<false><<<c->cg_grid, CG_THREADS, 0, st>>>(a, it, counter);
      PF_LAUNCHED(c);
      PF_TRY(comm_halo(c->comm, pnext, 1, 1, st, c->err));
      PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_PQ, 1, 0, st, c->err));
      if (unit) k_cg_phaseB<true><<<c->cg_grid, CG_THREADS, 0, st>>>(a, it, counter);
      else k_cg_phaseB<false><<<c->cg_grid, CG_THREADS, 0, st>>>(a, it, counter);
      PF_LAUNCHED(c);
      PF_TRY(comm_halo(c->comm, c->r, 1, 1, st, c->err));
      PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_RR_NEW, 2, 0, st, c->err));
      k_cg_update<<<1, 32, 0, st>>>(c->scal, it, maxiter);
      PF_LAUNCHED(c);
      if ((it + 1) % poll == 0 && it + 1 < maxiter) {
        // all ranks hold identical scalars, so they all leave at the same iteration
        PF_CUDA(c, cudaMemcpyAsync(c->h_scal + S_DONE, c->scal + S_DONE, sizeof(double), cudaMemcpyDeviceToHost, st));
        PF_CUDA(c, cudaStreamSynchronize(st));
        if (c->h_scal[S_DONE] != 0.0) break;
      }
    }
    k_cg_finish<<<1, 32, 0, st>>>(c->scal, maxiter);
    PF_LAUNCHED(c);
  } else if (c->opt_multi_single) {
    fill_sync(c, a.y);
    PF_TRY(reset_trace(c, st));
    void* args[] = {&a};
    const void* fn = unit ? (const void*)k_cg_persistent_multi<true> : (const void*)k_cg_persistent_multi<false>;
    PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(c->cg_grid), dim3(CG_THREADS), args, 0, st));
    c->launches++;
  } else {
    if (unit) PF_TRY(launch_cg<true>(c, a, st));
    else PF_TRY(launch_cg<false>(c, a, st));
  }
  if (e1) PF_CUDA(c, cudaEventRecord(e1, st));
  if (!with_resid) return PF_OK;          // the caller's next pass forms ||b - A x|| itself (k_finish)
  if (c->comm) PF_TRY(comm_halo(c->comm, c->x, 1, 1, st, c->err));
  if (unit) k_resid<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->x, dx, dy, red);
  else k_resid<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->x, dx, dy, red);
  PF_LAUNCHED(c);
  if (c->comm) PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_RES_SUMSQ, 2, 0, st, c->err));
  return PF_OK;
}

int read_scalars(pf_ctx* c, cudaStream_t st, bool with_center) {
  PF_CUDA(c, cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
  if (with_center)
    PF_CUDA(c, cudaMemcpyAsync(c->h_center, c->center, sizeof(double) * (size_t)c->s.nyl, cudaMemcpyDeviceToHost, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

double nanmax(double mx, double nancount) { return nancount > 0.0 ? NAN : mx; }

// dense host array (row stride `stride`) <-> device field
int h2d_field(pf_ctx* c, double* dst, const double* src, int64_t stride, cudaStream_t st) {
  PF_CUDA(c, cudaMemcpy2DAsync(dst + cell(c->s, 0, 0), sizeof(double) * (size_t)c->s.pitch, src,
                               sizeof(double) * (size_t)stride, sizeof(double) * (size_t)c->s.nx, (size_t)c->s.nyl,
                               cudaMemcpyHostToDevice, st));
  return PF_OK;
}
int d2h_field(pf_ctx* c, double* dst, const double* src, int64_t stride, cudaStream_t st) {
  PF_CUDA(c, cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)stride, src + cell(c->s, 0, 0),
                               sizeof(double) * (size_t)c->s.pitch, sizeof(double) * (size_t)c->s.nx, (size_t)c->s.nyl,
                               cudaMemcpyDeviceToHost, st));
  return PF_OK;
}

// predictor + rhs pass 1 in one pass: (u, v) -> (uo, vo) buffers = u*, v*; b
template <bool UNIT>
int launch_front(pf_ctx* c, int scheme, double dt, double dt_safe, double dx, double dy, cudaStream_t st) {
  const Slab& s = c->s;
  const Red red{c->red_part, c->red_counter, c->scal};
  switch (scheme) {
    case PF_ADV_UPWIND:
      k_front<UNIT, PF_ADV_UPWIND><<<c->fgrid, c->fthreads, 0, st>>>(s, c->u, c->v, c->uo, c->vo, c->b, dt, dt_safe, dx, dy, c->fband, red);
      break;
    case PF_ADV_QUICK:
      k_front<UNIT, PF_ADV_QUICK><<<c->fgrid, c->fthreads, 0, st>>>(s, c->u, c->v, c->uo, c->vo, c->b, dt, dt_safe, dx, dy, c->fband, red);
      break;
    case PF_ADV_NONE:
      k_front<UNIT, PF_ADV_NONE><<<c->fgrid, c->fthreads, 0, st>>>(s, c->u, c->v, c->uo, c->vo, c->b, dt, dt_safe, dx, dy, c->fband, red);
      break;
    default:
      return fail(c, PF_ERR_ARG, "pf_step", "unknown advection scheme");
  }
  PF_LAUNCHED(c);
  return PF_OK;
}

template <bool UNIT>
int launch_predict(pf_ctx* c, int scheme, double dt, double dx, double dy, cudaStream_t st) {
  const Slab& s = c->s;
  switch (scheme) {
    case PF_ADV_UPWIND:
      k_predict<UNIT, PF_ADV_UPWIND><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->uo, c->vo, dt, dx, dy);
      break;
    case PF_ADV_QUICK:
      k_predict<UNIT, PF_ADV_QUICK><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->uo, c->vo, dt, dx, dy);
      break;
    case PF_ADV_NONE:
      k_predict<UNIT, PF_ADV_NONE><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->uo, c->vo, dt, dx, dy);
      break;
    default:
      return fail(c, PF_ERR_ARG, "pf_step", "unknown advection scheme");
  }
  PF_LAUNCHED(c);
  return PF_OK;
}

}  // namespace

extern "C" {

int pf_abi_version(void) { return PF_ABI_VERSION; }

const char* pf_strerror(int status) {
  switch (status) {
    case PF_OK: return "ok";
    case PF_ERR_ARG: return "invalid argument";
    case PF_ERR_CUDA: return "CUDA error";
    case PF_ERR_NCCL: return "NCCL error";
    case PF_ERR_STATE: return "invalid call order / state";
    case PF_ERR_UNSUPPORTED: return "unsupported device or configuration";
    default: return "unknown status";
  }
}

const char* pf_last_error(const pf_ctx* ctx) { return ctx ? ctx->err : ""; }

size_t pf_workspace_bytes(int64_t nx, int64_t ny_local) {
  if (nx < 2 || ny_local < 1) return 0;
  const size_t fb = align_up(field_elems_of(nx, ny_local) * sizeof(double), ALIGN);
  return NFIELDS * fb + small_layout(ny_local, nx).total + ALIGN;
}

int pf_create(pf_ctx** out, int64_t nx, int64_t ny_local, int64_t ny_global, int64_t row_offset, int device,
              void* workspace, size_t workspace_bytes) {
  if (!out) return PF_ERR_ARG;
  *out = nullptr;
  if (nx < 2 || ny_global < 2 || ny_local < 1 || row_offset < 0 || row_offset + ny_local > ny_global) return PF_ERR_ARG;
  if (nx > (1 << 20) || ny_global > (1 << 28) || field_elems_of(nx, ny_local) > (size_t)0x7fffffff) return PF_ERR_ARG;
  if (!workspace || ((uintptr_t)workspace % ALIGN) != 0 || workspace_bytes < pf_workspace_bytes(nx, ny_local))
    return PF_ERR_ARG;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return PF_ERR_CUDA;
  pf_ctx* c = new pf_ctx();
  c->device = device;
  int coop = 0, major = 0;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (!coop || major < 10) {
    delete c;
    return PF_ERR_UNSUPPORTED;
  }
  c->s.nx = (int)nx;
  c->s.nyl = (int)ny_local;
  c->s.nyg = (int)ny_global;
  c->s.row0 = (int)row_offset;
  c->s.pitch = pitch_of(nx);
  c->field_elems = field_elems_of(nx, ny_local);
  const size_t fb = align_up(c->field_elems * sizeof(double), ALIGN);
  char* base = (char*)workspace;
  double** fields[NFIELDS] = {&c->u, &c->v, &c->uo, &c->vo, &c->b, &c->x, &c->r, &c->p0, &c->p1};
  for (int k = 0; k < NFIELDS; ++k) *fields[k] = (double*)(base + (size_t)k * fb);
  c->ws_base = base;
  c->field_stride = fb;
  char* small = base + (size_t)NFIELDS * fb;
  const SmallLayout L = small_layout(ny_local, nx);
  c->red_part = (double*)(small + L.red_part);
  c->red_counter = (unsigned*)(small + L.red_counter);
  c->scal = (double*)(small + L.scal);
  c->cg_part = (double*)(small + L.cg_part);
  c->center = (double*)(small + L.center);
  c->small_halo = (double*)(small + L.small_halo);
  c->small_part = (double*)(small + L.small_part);
  c->sync_off = (size_t)NFIELDS * fb + L.sync;
  c->sgrid = simple_grid(c->s);
  fused_geometry(c);
#define PF_CREATE_CUDA(call)                                  \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) {                                 \
      fprintf(stderr, "pf_create: %s: %s\n", #call, cudaGetErrorString(e__)); \
      pf_destroy(c);                                          \
      return PF_ERR_CUDA;                                     \
    }                                                         \
  } while (0)
  PF_CREATE_CUDA(cudaMemset(workspace, 0, pf_workspace_bytes(nx, ny_local) - ALIGN));
  PF_CREATE_CUDA(cudaHostAlloc((void**)&c->h_scal, sizeof(double) * S_COUNT, cudaHostAllocDefault));
  PF_CREATE_CUDA(cudaHostAlloc((void**)&c->h_center, sizeof(double) * (size_t)ny_local, cudaHostAllocDefault));
  for (auto& ev : c->ev) PF_CREATE_CUDA(cudaEventCreate(&ev));
  int occ_u = 0, occ_g = 0;
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, k_cg_persistent<true>, CG_THREADS, 0));
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, k_cg_persistent<false>, CG_THREADS, 0));
  int occ_mu = 0, occ_mg = 0;
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_mu, k_cg_persistent_multi<true>, CG_THREADS, 0));
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_mg, k_cg_persistent_multi<false>, CG_THREADS, 0));
  const int occ = std::min(std::min(occ_u, occ_g), std::min(occ_mu, occ_mg));
  if (occ < 1) {
    pf_destroy(c);
    return PF_ERR_UNSUPPORTED;
  }
  c->cg_grid = std::min(c->sm_count * occ, CG_MAX_GRID);
  int occ_gen_u = 0, occ_gen_g = 0;
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_gen_u, k_cg_generic<true>, CG_THREADS, 0));
  PF_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_gen_g, k_cg_generic<false>, CG_THREADS, 0));
  {
    cudaError_t e1 = cudaFuncSetAttribute(k_cg_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(k_cg_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_cg_tmad<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_cg_tmad<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
    int o1 = 0, o2 = 0;
    if (e1 == cudaSuccess && e2 == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_cg_tma<true>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_cg_tma<false>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
        o1 >= 1 && o2 >= 1) {
      int o3 = 0, o4 = 0;
      cudaError_t e3 = cudaFuncSetAttribute(k_cg_tmad<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
      cudaError_t e4 = cudaFuncSetAttribute(k_cg_tmad<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
      int o5 = 0, o6 = 0;
      if (e3 == cudaSuccess && e4 == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_cg_tmad<true, false>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o4, k_cg_tmad<false, false>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o5, k_cg_tmad<true, true>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o6, k_cg_tmad<false, true>, TM_THREADS, TM_SMEM_BYTES) == cudaSuccess &&
          o3 >= 1 && o4 >= 1 && o5 >= 1 && o6 >= 1)
        c->tma_grid = c->sm_count;   // one CTA per SM
    }
    (void)cudaGetLastError();
  }
  c->small_ok = true;
  for (int np = 1; np <= SM_PAIRS; np *= 2)
    for (int unit = 0; unit < 2; ++unit)
      c->small_ok = c->small_ok && cudaFuncSetAttribute(small_kernel(unit != 0, np), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         SM_MAX_SMEM) == cudaSuccess;
  (void)cudaGetLastError();
  c->gen_grid = std::min(c->sm_count * std::max(1, std::min(occ_gen_u, occ_gen_g)), CG_MAX_GRID);
#undef PF_CREATE_CUDA
  *out = c;
  return PF_OK;
}

int pf_destroy(pf_ctx* c) {
  if (!c) return PF_OK;
  cudaSetDevice(c->device);
  if (c->comm) comm_destroy(c->comm);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->h_center) cudaFreeHost(c->h_center);
  for (auto& ev : c->ev)
    if (ev) cudaEventDestroy(ev);
  delete c;
  return PF_OK;
}

int pf_device_info(pf_ctx* c, int* sm_count, int* cg_grid_blocks, int* cg_block_threads) {
  if (!c) return PF_ERR_ARG;
  if (sm_count) *sm_count = c->sm_count;
  if (cg_grid_blocks) *cg_grid_blocks = c->cg_grid;
  if (cg_block_threads) *cg_block_threads = CG_THREADS;
  return PF_OK;
}

int pf_set_option(pf_ctx* c, const char* key, int value) {
  if (!c || !key) return PF_ERR_ARG;
  if (strcmp(key, "cg_multi_single") == 0) { c->opt_multi_single = value; return PF_OK; }
  if (strcmp(key, "cg_sync_opt") == 0) { c->opt_sync = value; return PF_OK; }
  if (strcmp(key, "cg_kernel") == 0) { c->opt_kernel = value; return PF_OK; }
  if (strcmp(key, "step_fused") == 0) { c->opt_step_fused = value ? 1 : 0; return PF_OK; }
  if (strcmp(key, "cg_rows_min") == 0 && value >= 1) { c->opt_rows_min = value; return PF_OK; }
  if (strcmp(key, "cg_tile_rows") == 0 && value >= 0) { c->opt_tile_rows = value; return PF_OK; }
  if (strcmp(key, "cg_small_rows") == 0 && value >= -1) { c->opt_small_rows = value; return PF_OK; }
  if (strcmp(key, "cg_nb_small") == 0 && value >= -1) { c->opt_nb_small = value; return PF_OK; }
  return fail(c, PF_ERR_ARG, "pf_set_option", key);
}

int pf_debug_read_partials(pf_ctx* c, double* out, int offset, int n) {
  if (!c || !out || offset < 0 || n < 0 || offset + n > PF_PART_DOUBLES) return PF_ERR_ARG;
  PF_CUDA(c, cudaMemcpy(out, c->cg_part + offset, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  return PF_OK;
}

int pf_launch_count(pf_ctx* c, int64_t* out) {
  if (!c || !out) return PF_ERR_ARG;
  *out = (int64_t)c->launches;
  return PF_OK;
}

int pf_upload(pf_ctx* c, const double* u, const double* v, const double* p, int64_t row_stride, void* stream) {
  if (!c || row_stride < c->s.nx) return PF_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (u) c->center_valid = false;
  if (u) PF_TRY(h2d_field(c, c->u, u, row_stride, st));
  if (v) PF_TRY(h2d_field(c, c->v, v, row_stride, st));
  if (p) PF_TRY(h2d_field(c, c->x, p, row_stride, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_download(pf_ctx* c, double* u, double* v, double* p, int64_t row_stride, void* stream) {
  if (!c || row_stride < c->s.nx) return PF_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (u) PF_TRY(d2h_field(c, u, c->u, row_stride, st));
  if (v) PF_TRY(d2h_field(c, v, c->v, row_stride, st));
  if (p) PF_TRY(d2h_field(c, p, c->x, row_stride, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_prestep_reduce(pf_ctx* c, double out[3], void* stream) {
  if (!c || !out) return PF_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const Red red{c->red_part, c->red_counter, c->scal};
  if (c->comm) PF_TRY(comm_halo(c->comm, c->v, 1, 1, st, c->err));
  k_prestep<<<c->sgrid, KT, 0, st>>>(c->s, c->u, c->v, red);
  PF_LAUNCHED(c);
  if (c->comm) {
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_UMAX, 2, 1, st, c->err));
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_DIVC_SUMSQ, 7, 0, st, c->err));
  }
  PF_TRY(read_scalars(c, st, false));
  out[0] = nanmax(c->h_scal[S_UMAX], c->h_scal[S_UNAN]);
  out[1] = nanmax(c->h_scal[S_VMAX], c->h_scal[S_VNAN]);
  out[2] = c->h_scal[S_DIVC_SUMSQ];
  return PF_OK;
}

int pf_step(pf_ctx* c, const pf_step_params* P, pf_step_out* out, void* stream) {
  if (!c || !P || !out) return PF_ERR_ARG;
  if (P->lin_maxiter < 0) return fail(c, PF_ERR_ARG, "pf_step", "lin_maxiter < 0");
  cudaStream_t st = (cudaStream_t)stream;
  const Slab& s = c->s;
  const Red red{c->red_part, c->red_counter, c->scal};
  const bool unit = is_unit(P->dx, P->dy);
  const double dt = P->dt, dx = P->dx, dy = P->dy;
  const double dt_safe = (1e-14 > dt) ? 1e-14 : dt;      // Python max(dt, 1e-14) incl. NaN behaviour
  PF_CUDA(c, cudaEventRecord(c->ev[0], st));

  const bool fused = c->opt_step_fused && !c->comm && c->fgrid.x > 0;
  if (fused) {
    // one GPU: two passes around the solve (pf_step_fused.cuh); u, v ping-pong between the two buffer pairs
    if (!P->projection_only) {
      if (unit) PF_TRY(launch_front<true>(c, P->scheme, dt, dt_safe, dx, dy, st));
      else PF_TRY(launch_front<false>(c, P->scheme, dt, dt_safe, dx, dy, st));
      std::swap(c->u, c->uo);
      std::swap(c->v, c->vo);
    } else {
      if (unit) k_rhs1<true><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->b, dt_safe, dx, dy, red);
      else k_rhs1<false><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->b, dt_safe, dx, dy, red);
      PF_LAUNCHED(c);
    }
    PF_TRY(solve_on_device(c, dx, dy, P->jacobi, P->inv_diag_interior, P->lin_tol, P->lin_maxiter, 1, 1, st, c->ev[1],
                           c->ev[2], /*with_resid=*/false));
    // residual + correction + divergence + lid row + post pass: (u*, v*) in c->u / c->v, the previous step's fields
    // in c->uo / c->vo, which receive the new fields
    const int lid_set = P->lid_set ? 1 : 0;
    if (P->projection_only) {
      if (unit) k_finish<true, false><<<c->fgrid, c->fthreads, 0, st>>>(s, c->b, c->x, c->u, c->v, c->uo, c->vo, dt, dx, dy, 0, 0.0, c->center, c->fband, red);
      else k_finish<false, false><<<c->fgrid, c->fthreads, 0, st>>>(s, c->b, c->x, c->u, c->v, c->uo, c->vo, dt, dx, dy, 0, 0.0, c->center, c->fband, red);
    } else {
      if (unit) k_finish<true, true><<<c->fgrid, c->fthreads, 0, st>>>(s, c->b, c->x, c->u, c->v, c->uo, c->vo, dt, dx, dy, lid_set, P->lid_velocity, c->center, c->fband, red);
      else k_finish<false, true><<<c->fgrid, c->fthreads, 0, st>>>(s, c->b, c->x, c->u, c->v, c->uo, c->vo, dt, dx, dy, lid_set, P->lid_velocity, c->center, c->fband, red);
    }
    PF_LAUNCHED(c);
    std::swap(c->u, c->uo);
    std::swap(c->v, c->vo);
  } else {
    if (!P->projection_only) {
      // predictor: (u, v) -> (uo, vo) buffers, then swap so that u/v hold u*, v*
      // and uo/vo keep the previous step's fields for the residuals.
      if (c->comm) {
        PF_TRY(comm_halo(c->comm, c->u, 2, 2, st, c->err));
        PF_TRY(comm_halo(c->comm, c->v, 2, 2, st, c->err));
      }
      if (unit) PF_TRY(launch_predict<true>(c, P->scheme, dt, dx, dy, st));
      else PF_TRY(launch_predict<false>(c, P->scheme, dt, dx, dy, st));
      std::swap(c->u, c->uo);
      std::swap(c->v, c->vo);
    }
    // rhs pass 1
    if (c->comm) PF_TRY(comm_halo(c->comm, c->v, 1, 0, st, c->err));
    if (unit) k_rhs1<true><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->b, dt_safe, dx, dy, red);
    else k_rhs1<false><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->b, dt_safe, dx, dy, red);
    PF_LAUNCHED(c);
    if (c->comm) PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_DIVB_SUMSQ, 1, 0, st, c->err));
    // rhs pass 2 + CG + true residual
    PF_TRY(solve_on_device(c, dx, dy, P->jacobi, P->inv_diag_interior, P->lin_tol, P->lin_maxiter, 1, 1, st, c->ev[1],
                           c->ev[2]));
    // correction
    if (unit) k_correct<true><<<c->sgrid, KT, 0, st>>>(s, c->x, c->x, c->u, c->v, dt, dx, dy, c->scal);
    else k_correct<false><<<c->sgrid, KT, 0, st>>>(s, c->x, c->x, c->u, c->v, dt, dx, dy, c->scal);
    PF_LAUNCHED(c);
    if (c->comm) PF_TRY(comm_halo(c->comm, c->v, 1, 0, st, c->err));
    if (unit) k_divnorm<true><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, dx, dy, red, S_DIVA_SUMSQ);
    else k_divnorm<false><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, dx, dy, red, S_DIVA_SUMSQ);
    PF_LAUNCHED(c);
    if (!P->projection_only) {
      if (P->lid_set) {
        k_lid<<<(s.nx + KT - 1) / KT, KT, 0, st>>>(s, c->u, P->lid_velocity);
        PF_LAUNCHED(c);
      }
      if (c->comm) PF_TRY(comm_halo(c->comm, c->v, 1, 1, st, c->err));
      k_post<<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, c->uo, c->vo, c->x, c->center, red);
      PF_LAUNCHED(c);
    }
  }
  if (c->comm) {
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_UMAX, 2, 1, st, c->err));
    PF_TRY(comm_allreduce_scalars(c->comm, c->scal + S_DIVC_SUMSQ, 7, 0, st, c->err));
  }
  PF_CUDA(c, cudaEventRecord(c->ev[3], st));
  PF_TRY(read_scalars(c, st, !P->projection_only));
  c->center_valid = !P->projection_only;

  const double* h = c->h_scal;
  memset(out, 0, sizeof(*out));
  const double ncells = (double)s.nx * (double)s.nyg;
  out->div_before_norm = sqrt(h[S_DIVB_SUMSQ]);
  out->rhs_mean = h[S_RHS_SUM] / ncells;
  out->bnorm = sqrt(h[S_BB]);
  out->rnorm = sqrt(h[S_RES_SUMSQ]);
  out->div_after_norm = sqrt(h[S_DIVA_SUMSQ]);
  out->p00 = h[S_X0] - h[S_X0];   /* p -= p.flat[0] (pressure_solver.py:148) */
  if (h[S_INFO] == -2.0) return fail(c, PF_ERR_STATE, "pf_step", "a bounded wait inside the CG kernel timed out (pipeline or cross-GPU exchange)");
  out->info = (int32_t)h[S_INFO];
  out->true_iters = (int32_t)h[S_ITERS];
  if (!P->projection_only) {
    out->ru_sumsq = h[S_RU_SUMSQ];
    out->rv_sumsq = h[S_RV_SUMSQ];
    out->cont_sumsq = h[S_DIVC_SUMSQ];
    out->umax_raw = nanmax(h[S_UMAX], h[S_UNAN]);
    out->vmax_raw = nanmax(h[S_VMAX], h[S_VNAN]);
    out->nonfinite = h[S_NONFINITE] > 0.0 ? 1 : 0;
  }
  float ms = 0.f;
  PF_CUDA(c, cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]));
  out->cg_ms = ms;
  PF_CUDA(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]));
  out->step_ms = ms;
  return PF_OK;
}

int pf_divergence_stats(pf_ctx* c, double dx, double dy, double out[2], void* stream) {
  if (!c || !out) return PF_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const Slab& s = c->s;
  const Red red{c->red_part, c->red_counter, c->scal};
  const double ncells = (double)s.nx * (double)s.nyg;
  if (c->comm) PF_TRY(comm_halo(c->comm, c->v, 1, 1, st, c->err));
  for (int pass = 0; pass < 2; ++pass) {
    if (is_unit(dx, dy)) k_divc_stats<true><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, dx, dy, pass, ncells, red);
    else k_divc_stats<false><<<c->sgrid, KT, 0, st>>>(s, c->u, c->v, dx, dy, pass, ncells, red);
    PF_LAUNCHED(c);
    if (c->comm) PF_TRY(comm_allreduce_scalars(c->comm, c->scal + (pass ? S_TMP1 : S_TMP0), 1, 0, st, c->err));
  }
  PF_TRY(read_scalars(c, st, false));
  out[0] = c->h_scal[S_TMP0] / ncells;
  out[1] = sqrt(c->h_scal[S_TMP1]);
  return PF_OK;
}

int pf_centerline(pf_ctx* c, double* host_out, void* stream) {
  if (!c || !host_out) return PF_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->center_valid) {           // came back with the last pf_step's scalar read-back
    memcpy(host_out, c->h_center, sizeof(double) * (size_t)c->s.nyl);
    return PF_OK;
  }
  k_centerline<<<(c->s.nyl + KT - 1) / KT, KT, 0, st>>>(c->s, c->u, c->center);
  PF_LAUNCHED(c);
  PF_CUDA(c, cudaMemcpyAsync(c->h_center, c->center, sizeof(double) * (size_t)c->s.nyl, cudaMemcpyDeviceToHost, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  memcpy(host_out, c->h_center, sizeof(double) * (size_t)c->s.nyl);
  return PF_OK;
}

// ---- operator-level seams ------------------------------------------------------
// Scratch use: inputs land in b / r / p0, outputs in p1 / uo-like scratch.  The
// State fields (u, v, x) are left untouched.
#define PF_OP_PROLOGUE()                                             \
  if (!c) return PF_ERR_ARG;                                         \
  if (c->comm) return fail(c, PF_ERR_STATE, "operator seam", "single-GPU contexts only"); \
  cudaStream_t st = (cudaStream_t)stream;                            \
  const Slab& s = c->s;                                              \
  const int64_t nxs = s.nx;                                          \
  (void)nxs

int pf_op_gradient_coherent(pf_ctx* c, const double* p, double* gx, double* gy, double dx, double dy, void* stream) {
  PF_OP_PROLOGUE();
  if (!p || !gx || !gy) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, p, nxs, st));
  if (is_unit(dx, dy)) k_op_grad<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dx, dy);
  else k_op_grad<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dx, dy);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, gx, c->p0, nxs, st));
  PF_TRY(d2h_field(c, gy, c->p1, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

static int op_div(pf_ctx* c, const double* u, const double* v, double* out, double dx, double dy, int central,
                  void* stream) {
  PF_OP_PROLOGUE();
  if (!u || !v || !out) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, u, nxs, st));
  PF_TRY(h2d_field(c, c->r, v, nxs, st));
  if (is_unit(dx, dy)) k_op_div<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, dx, dy, central);
  else k_op_div<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, dx, dy, central);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, out, c->p0, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_divergence_coherent(pf_ctx* c, const double* u, const double* v, double* out, double dx, double dy,
                              void* stream) {
  return op_div(c, u, v, out, dx, dy, 0, stream);
}
int pf_op_divergence_central(pf_ctx* c, const double* u, const double* v, double* out, double dx, double dy,
                             void* stream) {
  return op_div(c, u, v, out, dx, dy, 1, stream);
}

int pf_op_laplacian_coherent(pf_ctx* c, const double* p, double* out, double dx, double dy, void* stream) {
  PF_OP_PROLOGUE();
  if (!p || !out) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, p, nxs, st));
  if (is_unit(dx, dy)) k_op_lap<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, dx, dy);
  else k_op_lap<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, dx, dy);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, out, c->p0, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_gradient_central(pf_ctx* c, const double* p, double* gx, double* gy, double dx, double dy, void* stream) {
  PF_OP_PROLOGUE();
  if (!p || !gx || !gy) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, p, nxs, st));
  if (is_unit(dx, dy)) k_op_grad_central<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dx, dy);
  else k_op_grad_central<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dx, dy);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, gx, c->p0, nxs, st));
  PF_TRY(d2h_field(c, gy, c->p1, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_laplacian_central(pf_ctx* c, const double* p, double* out, double dx, double dy, void* stream) {
  PF_OP_PROLOGUE();
  if (!p || !out) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, p, nxs, st));
  if (is_unit(dx, dy)) k_op_lap_central<true><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, dx, dy);
  else k_op_lap_central<false><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, dx, dy);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, out, c->p0, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_pressure_correction_central(pf_ctx* c, const double* u, const double* v, const double* p, double dt,
                                      double dx, double dy, double* u_out, double* v_out, void* stream) {
  PF_OP_PROLOGUE();
  if (!u || !v || !p || !u_out || !v_out) return PF_ERR_ARG;
  const bool unit = is_unit(dx, dy);
  PF_TRY(h2d_field(c, c->b, p, nxs, st));
  PF_TRY(h2d_field(c, c->r, u, nxs, st));
  PF_TRY(h2d_field(c, c->p0, v, nxs, st));
  if (unit) k_op_correct_central<true, 0><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p1, dt, dx);
  else k_op_correct_central<false, 0><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p1, dt, dx);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, u_out, c->p1, nxs, st));
  // stream order: the copy of p1 above completes before the next kernel overwrites it
  if (unit) k_op_correct_central<true, 1><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dt, dy);
  else k_op_correct_central<false, 1><<<c->sgrid, KT, 0, st>>>(s, c->b, c->p0, c->p1, dt, dy);
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, v_out, c->p1, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_advect(pf_ctx* c, int scheme, const double* u, const double* v, const double* f, double* out, double dx,
                 double dy, void* stream) {
  PF_OP_PROLOGUE();
  if (!u || !v || !f || !out) return PF_ERR_ARG;
  if (scheme != PF_ADV_UPWIND && scheme != PF_ADV_QUICK) return fail(c, PF_ERR_ARG, "pf_op_advect", "scheme");
  PF_TRY(h2d_field(c, c->b, u, nxs, st));
  PF_TRY(h2d_field(c, c->r, v, nxs, st));
  PF_TRY(h2d_field(c, c->p0, f, nxs, st));
  const bool unit = is_unit(dx, dy);
  if (scheme == PF_ADV_UPWIND) {
    if (unit) k_op_advect<true, PF_ADV_UPWIND><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, c->p1, dx, dy);
    else k_op_advect<false, PF_ADV_UPWIND><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, c->p1, dx, dy);
  } else {
    if (unit) k_op_advect<true, PF_ADV_QUICK><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, c->p1, dx, dy);
    else k_op_advect<false, PF_ADV_QUICK><<<c->sgrid, KT, 0, st>>>(s, c->b, c->r, c->p0, c->p1, dx, dy);
  }
  PF_LAUNCHED(c);
  PF_TRY(d2h_field(c, out, c->p1, nxs, st));
  PF_CUDA(c, cudaStreamSynchronize(st));
  return PF_OK;
}

int pf_op_pcg(pf_ctx* c, const double* b, double* x, double dx, double dy, int jacobi, double inv_diag_interior,
              double tol, int maxiter, int* info, int* true_iters, double* rnorm, void* stream) {
  PF_OP_PROLOGUE();
  if (!b || !x || maxiter < 0) return PF_ERR_ARG;
  // this seam overwrites the p field (x) of the context
  PF_TRY(h2d_field(c, c->b, b, nxs, st));
  PF_TRY(solve_on_device(c, dx, dy, jacobi, inv_diag_interior, tol, maxiter, 0, 0, st, nullptr, nullptr));
  PF_TRY(d2h_field(c, x, c->x, nxs, st));
  PF_TRY(read_scalars(c, st, false));
  if (c->h_scal[S_INFO] == -2.0) return fail(c, PF_ERR_STATE, "pf_op_pcg", "a bounded wait inside the CG kernel timed out");
  if (info) *info = (int)c->h_scal[S_INFO];
  if (true_iters) *true_iters = (int)c->h_scal[S_ITERS];
  if (rnorm) *rnorm = sqrt(c->h_scal[S_RES_SUMSQ]);
  return PF_OK;
}

int pf_op_cg_generic(pf_ctx* c, int mode, const double* b, double* x, double dx, double dy, int jacobi,
                     double inv_diag_interior, double tol, int maxiter, int* info, int* true_iters, double* rnorm,
                     double* bnorm, void* stream) {
  PF_OP_PROLOGUE();
  if (!b || maxiter < 0 || (mode != MODE_PINNED && mode != MODE_DIRICHLET)) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, b, nxs, st));
  GenArgs a{};
  a.s = s; a.b = c->b; a.x = c->x; a.r = c->r; a.p = c->p0; a.q = c->p1;
  a.part = c->cg_part; a.scal = c->scal;
  a.dx = dx; a.dy = dy; a.invd = jacobi ? inv_diag_interior : 1.0; a.tol = tol; a.maxiter = maxiter; a.mode = mode;
  void* args[] = {&a};
  const void* fn = is_unit(dx, dy) ? (const void*)k_cg_generic<true> : (const void*)k_cg_generic<false>;
  const long long ncell = (long long)s.nx * s.nyl;
  const int grid = (int)std::max(1LL, std::min((long long)c->gen_grid, (ncell + CG_THREADS - 1) / CG_THREADS));
  PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(CG_THREADS), args, 0, st));
  c->launches++;
  c->center_valid = false;
  if (x) PF_TRY(d2h_field(c, x, c->x, nxs, st));
  PF_TRY(read_scalars(c, st, false));
  if (info) *info = (int)c->h_scal[S_INFO];
  if (true_iters) *true_iters = (int)c->h_scal[S_ITERS];
  if (rnorm) *rnorm = sqrt(c->h_scal[S_RES_SUMSQ]);
  if (bnorm) *bnorm = sqrt(c->h_scal[S_BB]);
  return PF_OK;
}

// BiCGSTAB (method 1) or CG with a caller-supplied diagonal (method 0, inv_diag != NULL) on the pinned operator.
// These seams overwrite the p field (x) of the context; t and the diagonal live in the previous-step velocity
// buffers, which hold nothing between time steps (k_predict rewrites them before anything reads them).
static int krylov_generic(pf_ctx* c, int method, const double* b, const double* inv_diag, double* x, double dx,
                          double dy, int jacobi, double inv_diag_interior, double tol, int maxiter, int* info,
                          int* true_iters, double* rnorm, void* stream, const double* ring5_coef = nullptr) {
  PF_OP_PROLOGUE();
  if (!b || !x || maxiter < 0 || (method != 0 && method != 1) || (method == 0 && !inv_diag)) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, b, nxs, st));
  if (inv_diag) PF_TRY(h2d_field(c, c->vo, inv_diag, nxs, st));
  BiArgs a{};
  a.s = s; a.b = c->b; a.x = c->x; a.r = c->r; a.p = c->p0; a.v = c->p1; a.t = c->uo;
  a.part = c->cg_part; a.scal = c->scal;
  a.minv = inv_diag ? c->vo : nullptr;
  a.dx = dx; a.dy = dy; a.invd = jacobi ? inv_diag_interior : 1.0; a.tol = tol; a.maxiter = maxiter;
  a.ring5 = ring5_coef ? 1 : 0;
  for (int k = 0; k < 5; ++k) a.coef[k] = ring5_coef ? ring5_coef[k] : 0.0;
  void* args[] = {&a};
  const bool unit = is_unit(dx, dy);
  const void* fn = method == 1 ? (unit ? (const void*)k_bicgstab<true> : (const void*)k_bicgstab<false>)
                               : (unit ? (const void*)k_cg_diag<true> : (const void*)k_cg_diag<false>);
  int occ = 0;
  PF_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, CG_THREADS, 0));
  if (occ < 1) return fail(c, PF_ERR_UNSUPPORTED, "krylov_generic", "kernel does not fit on an SM");
  const long long ncell = (long long)s.nx * s.nyl;
  const long long resident = std::min((long long)c->sm_count * occ, (long long)CG_MAX_GRID);
  const int grid = (int)std::max(1LL, std::min(resident, (ncell + CG_THREADS - 1) / CG_THREADS));
  PF_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(CG_THREADS), args, 0, st));
  c->launches++;
  c->center_valid = false;
  PF_TRY(d2h_field(c, x, c->x, nxs, st));
  PF_TRY(read_scalars(c, st, false));
  if (info) *info = (int)c->h_scal[S_INFO];
  if (true_iters) *true_iters = (int)c->h_scal[S_ITERS];
  if (rnorm) *rnorm = sqrt(c->h_scal[S_RES_SUMSQ]);
  return PF_OK;
}

int pf_op_bicgstab(pf_ctx* c, const double* b, double* x, double dx, double dy, int jacobi, double inv_diag_interior,
                   double tol, int maxiter, int* info, int* true_iters, double* rnorm, void* stream) {
  return krylov_generic(c, 1, b, nullptr, x, dx, dy, jacobi, inv_diag_interior, tol, maxiter, info, true_iters, rnorm,
                        stream);
}

int pf_op_krylov_diag(pf_ctx* c, int method, const double* b, const double* inv_diag, double* x, double dx, double dy,
                      double tol, int maxiter, int* info, int* true_iters, double* rnorm, void* stream) {
  if (!inv_diag) return PF_ERR_ARG;
  return krylov_generic(c, method, b, inv_diag, x, dx, dy, 0, 1.0, tol, maxiter, info, true_iters, rnorm, stream);
}

// Krylov solve on the matrix the reference assembles (identity ring, 5-point interior; see MODE_RING5).
int pf_op_krylov_stencil5(pf_ctx* c, int method, const double* b, double* x, const double coef[5], int jacobi,
                          double inv_diag_interior, double tol, int maxiter, int* info, int* true_iters,
                          double* rnorm, void* stream) {
  if (!c || !coef) return PF_ERR_ARG;
  if (method == 1)
    return krylov_generic(c, 1, b, nullptr, x, 1.0, 1.0, jacobi, inv_diag_interior, tol, maxiter, info, true_iters,
                          rnorm, stream, coef);
  PF_OP_PROLOGUE();
  if (!b || !x || maxiter < 0 || method != 0) return PF_ERR_ARG;
  PF_TRY(h2d_field(c, c->b, b, nxs, st));
  GenArgs a{};
  a.s = s; a.b = c->b; a.x = c->x; a.r = c->r; a.p = c->p0; a.q = c->p1;
  a.part = c->cg_part; a.scal = c->scal;
  a.dx = 1.0; a.dy = 1.0; a.invd = jacobi ? inv_diag_interior : 1.0; a.tol = tol; a.maxiter = maxiter; a.mode = MODE_RING5;
  for (int k = 0; k < 5; ++k) a.coef[k] = coef[k];
  void* args[] = {&a};
  const long long ncell = (long long)s.nx * s.nyl;
  const int grid = (int)std::max(1LL, std::min((long long)c->gen_grid, (ncell + CG_THREADS - 1) / CG_THREADS));
  PF_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_cg_generic<true>, dim3(grid), dim3(CG_THREADS), args, 0, st));
  c->launches++;
  c->center_valid = false;
  PF_TRY(d2h_field(c, x, c->x, nxs, st));
  PF_TRY(read_scalars(c, st, false));
  if (info) *info = (int)c->h_scal[S_INFO];
  if (true_iters) *true_iters = (int)c->h_scal[S_ITERS];
  if (rnorm) *rnorm = sqrt(c->h_scal[S_RES_SUMSQ]);
  return PF_OK;
}

int pf_comm_unique_id(void* out128) { return comm_unique_id(out128); }

int pf_comm_init(pf_ctx* c, const void* id128, int rank, int nranks) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PF_ERR_ARG;
  if (c->comm) return fail(c, PF_ERR_STATE, "pf_comm_init", "already initialised");
  if (nranks == 1) return PF_OK;
  return comm_create(&c->comm, id128, rank, nranks, c->s, c->err);
}

int pf_comm_set_peers(pf_ctx* c, const void* const* peer_workspaces, int n) {
  if (!c || !peer_workspaces) return PF_ERR_ARG;
  if (!c->comm) return fail(c, PF_ERR_STATE, "pf_comm_set_peers", "pf_comm_init first");
  if (n != comm_nranks(c->comm) || n > CG_MAX_RANKS) return fail(c, PF_ERR_ARG, "pf_comm_set_peers", "rank count");
  if (c->s.nyg % n != 0 || c->s.nyl != c->s.nyg / n)
    return fail(c, PF_ERR_ARG, "pf_comm_set_peers", "peer mode needs equal slabs (ny_global % nranks == 0)");
  for (int q = 0; q < n; ++q) {
    if (!peer_workspaces[q]) return fail(c, PF_ERR_ARG, "pf_comm_set_peers", "null peer pointer");
    c->peer_ws[q] = (char*)peer_workspaces[q];
  }
  if (c->peer_ws[comm_rank(c->comm)] != c->ws_base)
    return fail(c, PF_ERR_ARG, "pf_comm_set_peers", "own entry must be the workspace given to pf_create");
  c->peers = true;
  return PF_OK;
}

}  // extern "C"
