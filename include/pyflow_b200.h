/*
 * pyflow_b200.h -- C-ABI of the B200-native PyFlow projection time step.
 *
 * Drop-in boundary for ONE hot path of beautifulplanet/PyFlow-Solver: the
 * projection-method step of the 2D lid-driven cavity.  The reference is pure
 * Python and has no FFI; the seams this library sits behind are the Python
 * callables listed per entry point (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every pointer is either a HOST pointer
 *     (caller-owned, never retained) or the device workspace handed to
 *     pf_create (caller-owned, e.g. a torch.empty() buffer; must outlive ctx).
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).
 *   - return 0 (PF_OK) or a negative pf_status; no exceptions, no aborts.
 *     Numerical conditions (NaN, non-convergence) are DATA in pf_step_out.
 *   - one context per GPU, not thread-safe; the caller serialises calls.
 *   - all fields are float64.  Host arrays use the reference's ghosted layout:
 *     pass the address of the first INTERIOR cell and the row stride in
 *     doubles (nx+2 for core/ghost_fields.py:47-57 arrays, nx for dense ones).
 *   - a context owns the row slab [row_offset, row_offset+ny_local) of a
 *     global ny_global x nx grid (single GPU: row_offset=0, ny_local=ny_global).
 */
#ifndef PYFLOW_B200_H
#define PYFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 1

typedef struct pf_ctx pf_ctx;

typedef enum pf_status {
  PF_OK = 0,
  PF_ERR_ARG = -1,          /* bad argument / size / null pointer            */
  PF_ERR_CUDA = -2,         /* CUDA runtime error (see pf_last_error)         */
  PF_ERR_NCCL = -3,         /* NCCL error                                     */
  PF_ERR_STATE = -4,        /* call order violated (e.g. comm not initialised)*/
  PF_ERR_UNSUPPORTED = -5   /* device is not sm_100 / cooperative launch n/a  */
} pf_status;

/* advection_scheme of solvers/solver.py:79-86 */
enum { PF_ADV_UPWIND = 0, PF_ADV_QUICK = 1, PF_ADV_NONE = 2 };

/* Inputs of one projection step that the HOST decides (solvers/solver.py:25-62
 * keeps the dt / safeguard branch logic in Python so those floats stay
 * bit-identical); everything else happens on the device.                     */
typedef struct pf_step_params {
  double dt;                 /* solver.py:32-57                                */
  double dx, dy;             /* solver.py:19-22                                */
  double lid_velocity;       /* solver.py:104-105 (used if lid_set)            */
  double lin_tol;            /* pressure_solver.py:134                         */
  double inv_diag_interior;  /* 1/(2/dx^2+2/dy^2), pressure_solver.py:35; host-computed */
  int32_t lid_set;           /* lid_velocity is not None                       */
  int32_t scheme;            /* PF_ADV_*  (PF_ADV_NONE = disable_advection)    */
  int32_t jacobi;            /* 1: ring-identity Jacobi (pressure_solver.py:139-143), 0: M=I */
  int32_t lin_maxiter;       /* pressure_solver.py:135                         */
  int32_t projection_only;   /* 1: solve_pressure_poisson seam only (no predictor, lid BC, residual pass) */
  int32_t reserved;
} pf_step_params;

/* Everything the host needs back from one step: one small D2H.               */
typedef struct pf_step_out {
  double div_before_norm;    /* ||div_coherent(u*,v*)||   pressure_solver.py:116-119 */
  double rhs_mean;           /* mean subtracted from rhs  pressure_solver.py:126      */
  double bnorm;              /* ||rhs||                                               */
  double rnorm;              /* ||b - A x||               interface.py:48-49          */
  double div_after_norm;     /* ||div_coherent(u,v)|| after correction, ps.py:165-169 */
  double ru_sumsq, rv_sumsq; /* sum (u-u_old)^2, sum (v-v_old)^2   solver.py:108-112  */
  double cont_sumsq;         /* sum div_central(u,v;1,1)^2 post-BC  solver.py:114     */
  double umax_raw, vmax_raw; /* max|u|, max|v| of the POST-step fields (NaN-propagating):
                                 next step's solver.py:25-26 inputs                    */
  double p00;                /* p at the pinned reference cell (solver.py:166-167 invariant) */
  double cg_ms;              /* device time of the CG solve (CUDA events)             */
  double step_ms;            /* device time of the whole step                         */
  int32_t info;              /* scipy cg info: 0 converged, lin_maxiter otherwise     */
  int32_t true_iters;        /* CG iterations actually executed                       */
  int32_t nonfinite;         /* any non-finite value in u, v, p  solver.py:116-121    */
  int32_t reserved;
} pf_step_out;

int         pf_abi_version(void);
const char* pf_strerror(int status);
/* detail of the last failure on this context ("" if none) */
const char* pf_last_error(const pf_ctx* ctx);

/* Bytes of device workspace pf_create needs for an nx x ny_local slab. */
size_t pf_workspace_bytes(int64_t nx, int64_t ny_local);

/* Replaces allocate_state (core/ghost_fields.py:47-57) for the device side.
 * `workspace` is device memory of >= pf_workspace_bytes(), 256-byte aligned. */
int pf_create(pf_ctx** out, int64_t nx, int64_t ny_local, int64_t ny_global,
              int64_t row_offset, int device, void* workspace, size_t workspace_bytes);
int pf_destroy(pf_ctx* ctx);

/* Host State <-> device fields (State.fields of core/ghost_fields.py:8-45).
 * Any of u/v/p may be NULL (skipped).  Synchronous on return.                */
int pf_upload(pf_ctx* ctx, const double* u, const double* v, const double* p,
              int64_t row_stride, void* stream);
int pf_download(pf_ctx* ctx, double* u, double* v, double* p,
                int64_t row_stride, void* stream);

/* solver.py:25-26,45: out = { max|u|, max|v| (NaN-propagating, without the
 * +1e-12), sum of div_central(u,v;1,1)^2 }.  Multi-GPU: already all-reduced. */
int pf_prestep_reduce(pf_ctx* ctx, double out[3], void* stream);

/* solver.py:71-121 + pressure_solver.py:97-183 + interface.py:35-52 + the CG
 * of scipy.sparse.linalg.cg, entirely on the device.  Synchronous on return. */
int pf_step(pf_ctx* ctx, const pf_step_params* params, pf_step_out* out, void* stream);

/* drivers/simulation_driver.py:89-96: u[:, nx//2] of the local rows
 * (ny_local doubles) as of the last pf_step / pf_upload.                     */
int pf_centerline(pf_ctx* ctx, double* host_out, void* stream);

/* driver/simulation_driver.py:126-130 (the alternate, dict-payload driver): mean and
 * mean-free 2-norm of fluid_ops.divergence(u, v, dx, dy) of the device fields:
 * out = { mean(div), ||div - mean(div)|| }.                                       */
int pf_divergence_stats(pf_ctx* ctx, double dx, double dy, double out[2], void* stream);

/* Operator-level seams on dense host arrays (row stride nx, ny_local rows):
 * numerics/core_ops_coherent.py:45,63,81; numerics/fluid_ops.py:30;
 * numerics/operators/advection.py:17,45; linear_solvers/interface.py:23 with
 * A = _NegLaplacianPinned (pressure_solver.py:97-105).  Single-GPU contexts. */
int pf_op_gradient_coherent(pf_ctx* ctx, const double* p, double* gx, double* gy,
                            double dx, double dy, void* stream);
int pf_op_divergence_coherent(pf_ctx* ctx, const double* u, const double* v, double* out,
                              double dx, double dy, void* stream);
int pf_op_laplacian_coherent(pf_ctx* ctx, const double* p, double* out,
                             double dx, double dy, void* stream);
int pf_op_divergence_central(pf_ctx* ctx, const double* u, const double* v, double* out,
                             double dx, double dy, void* stream);
/* Legacy central operators of the same module as pf_op_divergence_central, off the
 * time-step path but imported by the reference's own projection tests:
 * numerics/fluid_ops.py:43-67 (gradient), :69-91 (laplacian), :108-125
 * (apply_pressure_correction: u_out = u - dt*dp/dx, v_out = v - dt*dp/dy).      */
int pf_op_gradient_central(pf_ctx* ctx, const double* p, double* gx, double* gy,
                           double dx, double dy, void* stream);
int pf_op_laplacian_central(pf_ctx* ctx, const double* p, double* out,
                            double dx, double dy, void* stream);
int pf_op_pressure_correction_central(pf_ctx* ctx, const double* u, const double* v, const double* p,
                                      double dt, double dx, double dy,
                                      double* u_out, double* v_out, void* stream);
int pf_op_advect(pf_ctx* ctx, int scheme, const double* u, const double* v, const double* f,
                 double* out, double dx, double dy, void* stream);
int pf_op_pcg(pf_ctx* ctx, const double* b, double* x, double dx, double dy,
              int jacobi, double inv_diag_interior, double tol, int maxiter,
              int* info, int* true_iters, double* rnorm, void* stream);

/* Straightforward device CG (q stored, one thread per cell) for the operators off
 * the hot path.  mode 0: the pinned Neumann operator again (independent cross-check
 * of the tuned kernels); mode 1: +laplacian_coherent with identity rows on the
 * boundary ring = the manufactured / Dirichlet verification mode of
 * solve_pressure_poisson (pressure_solver.py:61-93).  b, x: dense host arrays
 * (x may be NULL); the solution also stays in the context's p field.            */
int pf_op_cg_generic(pf_ctx* ctx, int mode, const double* b, double* x, double dx, double dy, int jacobi,
                     double inv_diag_interior, double tol, int maxiter, int* info, int* true_iters,
                     double* rnorm, double* bnorm, void* stream);

/* BiCGSTAB on the pinned Neumann operator: the `method='bicgstab'` branch (and default) of the reference's
 * linear-solve seam, linear_solvers/interface.py:23,44-45 -> scipy.sparse.linalg.bicgstab with x0 = 0.  Same
 * conventions as pf_op_pcg; *info is scipy's exit code (0 converged, maxiter exhausted, -10 rho breakdown,
 * -11 omega / rtilde.v breakdown), *true_iters the completed iterations, *rnorm = ||b - A x||.  Off the time-step
 * path; overwrites the context's p field and the dead previous-step velocity buffer.  */
int pf_op_bicgstab(pf_ctx* ctx, const double* b, double* x, double dx, double dy,
                   int jacobi, double inv_diag_interior, double tol, int maxiter,
                   int* info, int* true_iters, double* rnorm, void* stream);

/* The same seam with a USER-SUPPLIED DIAGONAL preconditioner (`solve(A, b, M=...)`, interface.py:23,43-45, where M
 * is any operator that turns out to be diagonal): inv_diag is a dense host array (ny, nx) of the entries M applies
 * (z = inv_diag * r).  method 0: scipy's cg recurrence, method 1: scipy's bicgstab.  Outputs as pf_op_bicgstab.
 * Non-diagonal operators have no device form and are refused on the host side (no CPU fallback).            */
int pf_op_krylov_diag(pf_ctx* ctx, int method, const double* b, const double* inv_diag, double* x,
                      double dx, double dy, double tol, int maxiter,
                      int* info, int* true_iters, double* rnorm, void* stream);

/* The same two recurrences on the matrix the reference ASSEMBLES (assemble_negative_laplacian,
 * src/pyflow/solvers/pressure_solver.py:15-40): identity rows on the boundary ring, the plain 5-point stencil
 * coef = {south, west, centre, east, north} on interior rows, products summed in CSR column order like scipy's
 * csr_matvec.  This is what linear_solvers/interface.py:23-53 `solve(A, b)` runs when a caller hands it that sparse
 * matrix (recognised on the host from its CSR arrays) instead of the matrix-free operator.  method 0: cg, 1: bicgstab;
 * jacobi != 0: M = diag(1 on the ring, inv_diag_interior inside).  Outputs as pf_op_bicgstab.                      */
int pf_op_krylov_stencil5(pf_ctx* ctx, int method, const double* b, double* x, const double coef[5], int jacobi,
                          double inv_diag_interior, double tol, int maxiter,
                          int* info, int* true_iters, double* rnorm, void* stream);

/* Multi-GPU slabs (no reference counterpart: SURVEY.md section 8e).  One
 * process per GPU; rank 0 makes the id, the host side broadcasts the 128
 * bytes (torch.distributed), every rank calls pf_comm_init.                  */
int pf_comm_unique_id(void* out128);
int pf_comm_init(pf_ctx* ctx, const void* id128, int rank, int nranks);
/* Optional, after pf_comm_init: hand over the device workspaces of ALL ranks as
 * mapped into this process (symmetric memory / CUDA IPC peer pointers, entry
 * [rank] == own workspace, every rank's workspace zero-initialised by pf_create
 * and equal slabs).  The CG solve then runs as ONE persistent kernel per rank
 * that stores halo rows into the neighbours' HBM and exchanges the dot-product
 * partials through peer mailboxes (NVLink), instead of NCCL calls per phase.  */
int pf_comm_set_peers(pf_ctx* ctx, const void* const* peer_workspaces, int n);

/* Introspection used by bench.py / tests. */
int pf_device_info(pf_ctx* ctx, int* sm_count, int* cg_grid_blocks, int* cg_block_threads);
/* tuning / debugging switches: "cg_kernel" (3 automatic, 2 static TMA tiles,
 * 1 register-window kernels only), "step_fused" (1: one GPU runs the per-step
 * work around the solve as the two fused passes of pf_step_fused.cuh, the
 * default; 0: the seven separate kernels that slab runs on several GPUs use),
 * "cg_multi_single" (run the multi-GPU persistent kernel, own barrier included,
 * on a single-GPU context), "cg_tile_rows" / "cg_small_rows" / "cg_nb_small" /
 * "cg_rows_min" / "cg_sync_opt" (tile geometry and protocol experiments). */
int pf_set_option(pf_ctx* ctx, const char* key, int value);
/* debugging: copy n doubles of the CG partial-sum scratch (barrier timeline when
 * option "cg_sync_opt" bit 4 is set) */
int pf_debug_read_partials(pf_ctx* ctx, double* out, int offset, int n);
/* number of kernels of this library launched on the context so far */
int pf_launch_count(pf_ctx* ctx, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PYFLOW_B200_H */
