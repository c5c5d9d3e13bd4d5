// Host execution of the ELEMENT-WISE CUDA kernels of libpyflow_b200 (tests only).
//
// The .cuh sources are compiled by g++ against a minimal stand-in for <cuda_runtime.h> (this directory) and each
// kernel body is run once per emulated thread, block after block, on buffers in the device field layout
// ((ny + 2*HALO) rows x pitch doubles).  That is a faithful execution for kernels whose threads do not communicate
// (no __syncthreads, no shuffles, no atomics): the operator-level kernels, the predictor, the correction, the lid row,
// the centre line.  Compiled with -ffp-contract=off it evaluates the very expressions nvcc compiles with -fmad=false,
// so the results must equal the golden vectors of the reference bit for bit -- on a machine without a GPU.
// Kernels with grid reductions and the cooperative CG kernels are NOT emulated here; their parity is a GPU matter.
#include "cuda_runtime.h"

#include <string.h>

#include <thread>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
emu::Team emu::team;

#include "pf_kernels.cuh"
#include "pf_cg_generic.cuh"

using namespace pf;

namespace {

Slab slab_of(int nx, int ny) {
  Slab s;
  s.nx = nx; s.nyl = ny; s.nyg = ny; s.row0 = 0;
  s.pitch = (nx + PITCH_ALIGN - 1) / PITCH_ALIGN * PITCH_ALIGN;
  return s;
}

// same launch geometry as simple_grid() in pf_api.cu
template <class Body>
void launch(const Slab& s, Body body) {
  const unsigned gx = (unsigned)((s.nx + KT - 1) / KT);
  unsigned gy = 2368u / gx;
  if (gy < 1u) gy = 1u;
  if (gy > (unsigned)s.nyl) gy = (unsigned)s.nyl;
  gridDim = dim3(gx, gy, 1);
  blockDim = dim3(KT, 1, 1);
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx)
      for (unsigned tx = 0; tx < (unsigned)KT; ++tx) {
        blockIdx = {bx, by, 0};
        threadIdx = {tx, 0, 0};
        body();
      }
}

template <class Body>
void launch_1d(int n, Body body) {
  const unsigned gx = (unsigned)((n + KT - 1) / KT);
  gridDim = dim3(gx, 1, 1);
  blockDim = dim3(KT, 1, 1);
  for (unsigned bx = 0; bx < gx; ++bx)
    for (unsigned tx = 0; tx < (unsigned)KT; ++tx) {
      blockIdx = {bx, 0, 0};
      threadIdx = {tx, 0, 0};
      body();
    }
}

bool unit(double dx, double dy) { return dx == 1.0 && dy == 1.0; }

// One OS thread per CUDA thread of a block; the blocks of the grid one after the other.
template <class Body>
void launch_threaded(dim3 grid, int nthreads, Body body) {
  emu::Team& t = emu::team;
  pthread_barrier_init(&t.block, nullptr, (unsigned)nthreads);
  for (int w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&t.warp[w], nullptr, 32u);
  t.threaded = true;
  std::vector<std::thread> pool;
  for (int tx = 0; tx < nthreads; ++tx)
    pool.emplace_back([=] {
      threadIdx = {(unsigned)tx, 0, 0};
      gridDim = grid;
      blockDim = dim3((unsigned)nthreads, 1, 1);
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          blockIdx = {bx, by, 0};
          body();
          pthread_barrier_wait(&emu::team.block);      // the next block reuses the (static) shared memory
        }
    });
  for (auto& th : pool) th.join();
  t.threaded = false;
  pthread_barrier_destroy(&t.block);
  for (int w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&t.warp[w]);
}

dim3 simple_grid(const Slab& s) {
  const unsigned gx = (unsigned)((s.nx + KT - 1) / KT);
  unsigned gy = 2368u / gx;
  if (gy < 1u) gy = 1u;
  if (gy > (unsigned)s.nyl) gy = (unsigned)s.nyl;
  return dim3(gx, gy, 1);
}

}  // namespace

extern "C" {

// Can this machine run `n` OS threads at once?  (The threaded emulation needs one per CUDA thread of a block.)
int emu_thread_probe(int n) {
  std::vector<std::thread> pool;
  bool ok = true;
  try {
    for (int k = 0; k < n; ++k) pool.emplace_back([] {});
  } catch (...) {
    ok = false;
  }
  for (auto& th : pool) th.join();
  return ok ? 1 : 0;
}

int emu_pitch(int nx) { return slab_of(nx, 1).pitch; }
int emu_halo(void) { return HALO; }

void emu_op_grad(int nx, int ny, const double* p, double* gx, double* gy, double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  if (unit(dx, dy)) launch(s, [&] { k_op_grad<true>(s, p, gx, gy, dx, dy); });
  else launch(s, [&] { k_op_grad<false>(s, p, gx, gy, dx, dy); });
}

void emu_op_div(int nx, int ny, const double* u, const double* v, double* out, double dx, double dy, int central) {
  const Slab s = slab_of(nx, ny);
  if (unit(dx, dy)) launch(s, [&] { k_op_div<true>(s, u, v, out, dx, dy, central); });
  else launch(s, [&] { k_op_div<false>(s, u, v, out, dx, dy, central); });
}

void emu_op_lap(int nx, int ny, const double* p, double* out, double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  if (unit(dx, dy)) launch(s, [&] { k_op_lap<true>(s, p, out, dx, dy); });
  else launch(s, [&] { k_op_lap<false>(s, p, out, dx, dy); });
}

void emu_op_advect(int scheme, int nx, int ny, const double* u, const double* v, const double* f, double* out,
                   double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  const bool un = unit(dx, dy);
  if (scheme == PF_ADV_UPWIND) {
    if (un) launch(s, [&] { k_op_advect<true, PF_ADV_UPWIND>(s, u, v, f, out, dx, dy); });
    else launch(s, [&] { k_op_advect<false, PF_ADV_UPWIND>(s, u, v, f, out, dx, dy); });
  } else {
    if (un) launch(s, [&] { k_op_advect<true, PF_ADV_QUICK>(s, u, v, f, out, dx, dy); });
    else launch(s, [&] { k_op_advect<false, PF_ADV_QUICK>(s, u, v, f, out, dx, dy); });
  }
}

void emu_op_grad_central(int nx, int ny, const double* p, double* gx, double* gy, double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  if (unit(dx, dy)) launch(s, [&] { k_op_grad_central<true>(s, p, gx, gy, dx, dy); });
  else launch(s, [&] { k_op_grad_central<false>(s, p, gx, gy, dx, dy); });
}

void emu_op_lap_central(int nx, int ny, const double* p, double* out, double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  if (unit(dx, dy)) launch(s, [&] { k_op_lap_central<true>(s, p, out, dx, dy); });
  else launch(s, [&] { k_op_lap_central<false>(s, p, out, dx, dy); });
}

void emu_op_correct_central(int dir, int nx, int ny, const double* p, const double* f, double* out, double dt,
                            double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  const bool un = unit(dx, dy);
  const double h = dir == 0 ? dx : dy;
  if (dir == 0) {
    if (un) launch(s, [&] { k_op_correct_central<true, 0>(s, p, f, out, dt, h); });
    else launch(s, [&] { k_op_correct_central<false, 0>(s, p, f, out, dt, h); });
  } else {
    if (un) launch(s, [&] { k_op_correct_central<true, 1>(s, p, f, out, dt, h); });
    else launch(s, [&] { k_op_correct_central<false, 1>(s, p, f, out, dt, h); });
  }
}

// solver.py:71-95 -- (uo, vo) -> (un, vn)
void emu_predict(int scheme, int nx, int ny, const double* uo, const double* vo, double* un, double* vn, double dt,
                 double dx, double dy) {
  const Slab s = slab_of(nx, ny);
  const bool unt = unit(dx, dy);
#define PF_EMU_PREDICT(U, S) launch(s, [&] { k_predict<U, S>(s, uo, vo, un, vn, dt, dx, dy); })
  if (scheme == PF_ADV_UPWIND) { if (unt) PF_EMU_PREDICT(true, PF_ADV_UPWIND); else PF_EMU_PREDICT(false, PF_ADV_UPWIND); }
  else if (scheme == PF_ADV_QUICK) { if (unt) PF_EMU_PREDICT(true, PF_ADV_QUICK); else PF_EMU_PREDICT(false, PF_ADV_QUICK); }
  else { if (unt) PF_EMU_PREDICT(true, PF_ADV_NONE); else PF_EMU_PREDICT(false, PF_ADV_NONE); }
#undef PF_EMU_PREDICT
}

// pressure_solver.py:147-157 -- p = x - x[0]; u -= dt dp/dx; v -= dt dp/dy  (x0 passed through the scalar block)
void emu_correct(int nx, int ny, double* x, double* u, double* v, double dt, double dx, double dy, double x0) {
  const Slab s = slab_of(nx, ny);
  double scal[S_COUNT] = {0.0};
  scal[S_X0] = x0;
  if (unit(dx, dy)) launch(s, [&] { k_correct<true>(s, x, x, u, v, dt, dx, dy, scal); });
  else launch(s, [&] { k_correct<false>(s, x, x, u, v, dt, dx, dy, scal); });
}

void emu_lid(int nx, int ny, double* u, double lid) {
  const Slab s = slab_of(nx, ny);
  launch_1d(nx, [&] { k_lid(s, u, lid); });
}

void emu_centerline(int nx, int ny, const double* u, double* center) {
  const Slab s = slab_of(nx, ny);
  launch_1d(ny, [&] { k_centerline(s, u, center); });
}

// ---- kernels with grid reductions: one OS thread per CUDA thread -----------------------------------------------
// `scratch` holds the reduction workspace: MAX_RED_VALS*MAX_RED_BLOCKS partials, then 4 counters (as doubles' bytes),
// `scal` the S_COUNT-slot device scalar block.
static Red red_of(double* scratch, double* scal) {
  Red r;
  r.part = scratch;
  r.counter = (unsigned*)(scratch + (size_t)MAX_RED_VALS * MAX_RED_BLOCKS);
  r.scal = scal;
  return r;
}
int emu_scratch_doubles(void) { return MAX_RED_VALS * MAX_RED_BLOCKS + 8; }
int emu_scal_count(void) { return S_COUNT; }
int emu_slot(const char* name) {
#define PF_EMU_SLOT(n) if (!strcmp(name, #n)) return n
  PF_EMU_SLOT(S_UMAX); PF_EMU_SLOT(S_VMAX); PF_EMU_SLOT(S_DIVC_SUMSQ); PF_EMU_SLOT(S_DIVA_SUMSQ);
  PF_EMU_SLOT(S_RU_SUMSQ); PF_EMU_SLOT(S_RV_SUMSQ); PF_EMU_SLOT(S_NONFINITE); PF_EMU_SLOT(S_UNAN); PF_EMU_SLOT(S_VNAN);
  PF_EMU_SLOT(S_DIVB_SUMSQ); PF_EMU_SLOT(S_RHS_SUM); PF_EMU_SLOT(S_BB); PF_EMU_SLOT(S_BZ); PF_EMU_SLOT(S_RES_SUMSQ);
  PF_EMU_SLOT(S_X0); PF_EMU_SLOT(S_RR); PF_EMU_SLOT(S_INFO); PF_EMU_SLOT(S_ITERS); PF_EMU_SLOT(S_TMP0); PF_EMU_SLOT(S_TMP1);
#undef PF_EMU_SLOT
  return -1;
}

void emu_prestep(int nx, int ny, const double* u, const double* v, double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  launch_threaded(simple_grid(s), KT, [&] { k_prestep(s, u, v, red); });
}

void emu_rhs1(int nx, int ny, const double* u, const double* v, double* b, double dt_safe, double dx, double dy,
              double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  if (unit(dx, dy)) launch_threaded(simple_grid(s), KT, [&] { k_rhs1<true>(s, u, v, b, dt_safe, dx, dy, red); });
  else launch_threaded(simple_grid(s), KT, [&] { k_rhs1<false>(s, u, v, b, dt_safe, dx, dy, red); });
}

void emu_rhs2(int nx, int ny, double* b, double* x, double* r, int subtract_mean, int pin, double invd,
              double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  const double ncells = (double)nx * (double)ny;
  launch_threaded(simple_grid(s), KT, [&] { k_rhs2(s, b, x, r, ncells, subtract_mean, pin, invd, red); });
}

void emu_resid(int nx, int ny, const double* b, const double* x, double dx, double dy, double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  if (unit(dx, dy)) launch_threaded(simple_grid(s), KT, [&] { k_resid<true>(s, b, x, dx, dy, red); });
  else launch_threaded(simple_grid(s), KT, [&] { k_resid<false>(s, b, x, dx, dy, red); });
}

void emu_divnorm(int nx, int ny, const double* u, const double* v, double dx, double dy, double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  if (unit(dx, dy)) launch_threaded(simple_grid(s), KT, [&] { k_divnorm<true>(s, u, v, dx, dy, red, S_DIVA_SUMSQ); });
  else launch_threaded(simple_grid(s), KT, [&] { k_divnorm<false>(s, u, v, dx, dy, red, S_DIVA_SUMSQ); });
}

void emu_post(int nx, int ny, const double* u, const double* v, const double* uo, const double* vo, const double* p,
              double* center, double* scratch, double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  launch_threaded(simple_grid(s), KT, [&] { k_post(s, u, v, uo, vo, p, center, red); });
}

void emu_divc_stats(int nx, int ny, const double* u, const double* v, double dx, double dy, int pass, double* scratch,
                    double* scal) {
  const Slab s = slab_of(nx, ny);
  const Red red = red_of(scratch, scal);
  const double ncells = (double)nx * (double)ny;
  if (unit(dx, dy)) launch_threaded(simple_grid(s), KT, [&] { k_divc_stats<true>(s, u, v, dx, dy, pass, ncells, red); });
  else launch_threaded(simple_grid(s), KT, [&] { k_divc_stats<false>(s, u, v, dx, dy, pass, ncells, red); });
}

// The plain device CG (q stored, one thread per cell) on a cooperative grid of ONE block.
void emu_cg_generic(int mode, int nx, int ny, const double* b, double* x, double* r, double* p, double* q, double dx,
                    double dy, double invd, double tol, int maxiter, double* part, double* scal) {
  GenArgs a;
  a.s = slab_of(nx, ny);
  a.b = b; a.x = x; a.r = r; a.p = p; a.q = q; a.part = part; a.scal = scal;
  a.dx = dx; a.dy = dy; a.invd = invd; a.tol = tol; a.maxiter = maxiter; a.mode = mode;
  if (unit(dx, dy)) launch_threaded(dim3(1, 1, 1), CG_THREADS, [&] { k_cg_generic<true>(a); });
  else launch_threaded(dim3(1, 1, 1), CG_THREADS, [&] { k_cg_generic<false>(a); });
}
int emu_cg_part_doubles(void) { return PF_PART_DOUBLES; }

// The register-window persistent CG kernel (k_cg_persistent, the production kernel for widths that are not a multiple
// of 1024) on a cooperative grid of ONE block, with the tile geometry cg_geometry() of pf_api.cu computes for it.
// r = b, x = 0, S_BB and S_BZ must have been prepared (k_rhs2), as solve_on_device does.
void emu_cg_persistent(int nx, int ny, const double* b, double* x, double* r, double* p0, double* p1, double dx,
                       double dy, double invd, double tol, int maxiter, int rows_min_per_warp, double* part,
                       double* scal) {
  CgArgs a{};
  a.s = slab_of(nx, ny);
  a.b = b; a.x = x; a.r = r; a.p0 = p0; a.p1 = p1; a.part = part; a.scal = scal;
  a.dx = dx; a.dy = dy; a.invd = invd; a.tol = tol; a.maxiter = maxiter;
  const int G = 1;
  const int nstrips = (nx + STRIP - 1) / STRIP;
  int wx = 1;
  while (wx * 2 <= CG_WARPS && wx * 2 <= nstrips) wx *= 2;
  a.wx = wx;
  a.wy = CG_WARPS / wx;
  a.nchunks = (nstrips + wx - 1) / wx;
  const int rows_min = rows_min_per_warp * a.wy;
  const int max_bands = ny / rows_min > 1 ? ny / rows_min : 1;
  const int want = G / a.nchunks > 1 ? G / a.nchunks : 1;
  a.nbands = want < max_bands ? want : max_bands;
  a.ntiles = a.nchunks * a.nbands;
  if (unit(dx, dy)) launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_persistent<true>(a); });
  else launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_persistent<false>(a); });
}

// The split-phase CG kernels (k_cg_begin / k_cg_phaseA / k_cg_phaseB / k_cg_update / k_cg_finish): the sequence the
// library launches between NCCL calls on multi-GPU slabs without peer memory, here for ONE rank (no exchange needed)
// on a grid of G CTAs, so that several bands per chunk and several tiles per CTA are exercised.
void emu_cg_split(int nx, int ny, int G, const double* b, double* x, double* r, double* p0, double* p1, double dx,
                  double dy, double invd, double tol, int maxiter, int rows_min_per_warp, double* part, double* scal,
                  double* scratch) {
  CgArgs a{};
  a.s = slab_of(nx, ny);
  a.b = b; a.x = x; a.r = r; a.p0 = p0; a.p1 = p1; a.part = part; a.scal = scal;
  a.dx = dx; a.dy = dy; a.invd = invd; a.tol = tol; a.maxiter = maxiter;
  const int nstrips = (nx + STRIP - 1) / STRIP;
  int wx = 1;
  while (wx * 2 <= CG_WARPS && wx * 2 <= nstrips) wx *= 2;
  a.wx = wx;
  a.wy = CG_WARPS / wx;
  a.nchunks = (nstrips + wx - 1) / wx;
  const int rows_min = rows_min_per_warp * a.wy;
  const int max_bands = ny / rows_min > 1 ? ny / rows_min : 1;
  const int want = G / a.nchunks > 1 ? G / a.nchunks : 1;
  a.nbands = want < max_bands ? want : max_bands;
  a.ntiles = a.nchunks * a.nbands;
  unsigned* counter = (unsigned*)(scratch + (size_t)MAX_RED_VALS * MAX_RED_BLOCKS) + 1;
  const bool un = unit(dx, dy);
  launch_threaded(dim3(1, 1, 1), 32, [&] { k_cg_begin(scal, tol); });
  for (int it = 0; it < maxiter; ++it) {
    if (un) launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_phaseA<true>(a, it, counter); });
    else launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_phaseA<false>(a, it, counter); });
    if (un) launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_phaseB<true>(a, it, counter); });
    else launch_threaded(dim3(G, 1, 1), CG_THREADS, [&] { k_cg_phaseB<false>(a, it, counter); });
    launch_threaded(dim3(1, 1, 1), 32, [&] { k_cg_update(scal, it, maxiter); });
    if (scal[S_DONE] != 0.0) break;
  }
  launch_threaded(dim3(1, 1, 1), 32, [&] { k_cg_finish(scal, maxiter); });
}
int emu_cg_ntiles(int nx, int ny, int G, int rows_min_per_warp) {
  const int nstrips = (nx + STRIP - 1) / STRIP;
  int wx = 1;
  while (wx * 2 <= CG_WARPS && wx * 2 <= nstrips) wx *= 2;
  const int wy = CG_WARPS / wx, nchunks = (nstrips + wx - 1) / wx;
  const int rows_min = rows_min_per_warp * wy;
  const int max_bands = ny / rows_min > 1 ? ny / rows_min : 1;
  const int want = G / nchunks > 1 ? G / nchunks : 1;
  return nchunks * (want < max_bands ? want : max_bands);
}

}  // extern "C"
