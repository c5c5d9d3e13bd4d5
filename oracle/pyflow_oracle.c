/* CPU oracle, C / OpenMP restatement (TEST INFRASTRUCTURE ONLY).
 *
 * The same algorithm as oracle/pyflow_oracle.py (which is pinned bitwise against
 * the reference by oracle/make_golden.py), written as plain C loops so that the
 * CPU baseline of bench.py can use every host core.  Only tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it; the product
 * package never does.
 *
 * Parity status: PINNED through the numpy oracle -- tests/test_oracle_golden.py
 * checks every element-wise routine below bit-for-bit against the golden vectors
 * the unmodified reference produced (tests/golden/operators.npz) and whole
 * trajectories against tests/golden/trajectories.npz within the documented noise
 * floors (only the summation order of the dot products differs).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off (no FMA contraction, no fast-math),
 * see oracle/Makefile.  Reference citations are relative to the reference root.
 *
 * Arrays are the INTERIOR fields, row-major (ny rows of nx doubles, row 0 = bottom).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double dt, dx, dy, lid_velocity, lin_tol, inv_diag_interior;
    int32_t lid_set;      /* 0: leave the lid row alone */
    int32_t scheme;       /* 0 none (disable_advection), 1 upwind, 2 quick */
    int32_t jacobi;       /* ring-identity Jacobi on/off */
    int32_t lin_maxiter;
} pfo_params;

typedef struct {
    double div_before_norm, rhs_mean, bnorm, rnorm, div_after_norm;
    double ru_sumsq, rv_sumsq, cont_sumsq;
    int32_t info, true_iters, nonfinite, reserved;
} pfo_out;

int pfo_abi_version(void) { return 1; }

int pfo_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py / the parity checker run the oracle on every host core even when the launcher exported
 * OMP_NUM_THREADS=1 (torchrun does); results do not depend on the thread count (per-row partial sums). */
void pfo_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* deterministic reduction: one partial per row (independent of the thread count), folded serially */
static double fold(const double *rowsum, int ny) {
    double s = 0.0;
    for (int j = 0; j < ny; ++j) s += rowsum[j];
    return s;
}

/* ---- numerics/core_ops_coherent.py:45-84 ------------------------------------------------ */
static inline double gx_at(const double *p, int nx, int i, double dx) {      /* forward difference, 0 on the east face */
    return (i < nx - 1) ? (p[i + 1] - p[i]) / dx : 0.0;
}

/* laplacian_coherent = divergence_coherent(gradient_coherent(p)) for one cell */
static inline double lap_cell(const double *p, int nx, int ny, int i, int j, double dx, double dy) {
    const double *row = p + (size_t)j * nx;
    double xt = (i == 0) ? gx_at(row, nx, 0, dx) / dx : (gx_at(row, nx, i, dx) - gx_at(row, nx, i - 1, dx)) / dx;
    double gy = (j < ny - 1) ? (row[i + nx] - row[i]) / dy : 0.0;
    double yt;
    if (j == 0) yt = gy / dy;
    else {
        double gys = (row[i] - row[i - nx]) / dy;
        yt = (gy - gys) / dy;
    }
    return (0.0 + xt) + yt;
}

void pfo_gradient_coherent(int nx, int ny, const double *p, double dx, double dy, double *gx, double *gy) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            gx[c] = (nx >= 2 && ny >= 2) ? gx_at(p + (size_t)j * nx, nx, i, dx) : 0.0;
            gy[c] = (nx >= 2 && ny >= 2 && j < ny - 1) ? (p[c + nx] - p[c]) / dy : 0.0;
        }
}

static inline double div_cell(const double *u, const double *v, int nx, int i, int j, double dx, double dy) {
    size_t c = (size_t)j * nx + i;
    double xt = (i == 0) ? u[c] / dx : (u[c] - u[c - 1]) / dx;
    double yt = (j == 0) ? v[c] / dy : (v[c] - v[c - nx]) / dy;
    return (0.0 + xt) + yt;
}

void pfo_divergence_coherent(int nx, int ny, const double *u, const double *v, double dx, double dy, double *out) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            out[(size_t)j * nx + i] = (nx >= 2 && ny >= 2) ? div_cell(u, v, nx, i, j, dx, dy) : 0.0;
}

void pfo_laplacian_coherent(int nx, int ny, const double *p, double dx, double dy, double *out) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            out[(size_t)j * nx + i] = (nx >= 2 && ny >= 2) ? lap_cell(p, nx, ny, i, j, dx, dy) : 0.0;
}

/* ---- numerics/fluid_ops.py:8-28 (right column without corners is never written) ----------- */
static inline double divc_cell(const double *u, const double *v, int nx, int ny, int i, int j, double dx, double dy) {
    size_t c = (size_t)j * nx + i;
    int top = (j == ny - 1), bot = (j == 0), left = (i == 0), right = (i == nx - 1);
    if (right && !top && !bot) return 0.0;
    double xt, yt;
    if (left) xt = (u[c + 1] - u[c]) / dx;
    else if (right) xt = (u[c] - u[c - 1]) / dx;
    else xt = (u[c + 1] - u[c - 1]) / (2.0 * dx);
    if (bot) yt = (v[c + nx] - v[c]) / dy;
    else if (top) yt = (v[c] - v[c - nx]) / dy;
    else yt = (v[c + nx] - v[c - nx]) / (2.0 * dy);
    return xt + yt;
}

void pfo_divergence_central(int nx, int ny, const double *u, const double *v, double dx, double dy, double *out) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) out[(size_t)j * nx + i] = divc_cell(u, v, nx, ny, i, j, dx, dy);
}

/* sum of squares of the central divergence (dx = dy = 1 in solver.py:60,139) */
double pfo_divc_sumsq(int nx, int ny, const double *u, const double *v, double dx, double dy, double *rowsum) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) {
            double d = divc_cell(u, v, nx, ny, i, j, dx, dy);
            s += d * d;
        }
        rowsum[j] = s;
    }
    return fold(rowsum, ny);
}

/* max |u|, max |v| (solver.py:31-32) */
void pfo_absmax(int nx, int ny, const double *u, const double *v, double out[2]) {
    double mu = 0.0, mv = 0.0;
    size_t n = (size_t)nx * ny;
#pragma omp parallel for schedule(static) reduction(max : mu, mv)
    for (size_t c = 0; c < n; ++c) {
        double a = fabs(u[c]), b = fabs(v[c]);
        if (a > mu || a != a) mu = a;
        if (b > mv || b != b) mv = b;
    }
    out[0] = mu;
    out[1] = mv;
}

/* ---- numerics/operators/advection.py:17-84 ------------------------------------------------ */
static inline int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

static inline double upwind_cell(const double *u, const double *v, const double *f, int nx, int ny, int i, int j,
                                 double dx, double dy) {
    size_t c = (size_t)j * nx + i;
    double P = f[c];
    double W = f[(size_t)j * nx + clampi(i - 1, 0, nx - 1)], E = f[(size_t)j * nx + clampi(i + 1, 0, nx - 1)];
    double S = f[(size_t)clampi(j - 1, 0, ny - 1) * nx + i], N = f[(size_t)clampi(j + 1, 0, ny - 1) * nx + i];
    double fx = (u[c] > 0) ? (P - W) / dx : (E - P) / dx;
    double fy = (v[c] > 0) ? (P - S) / dy : (N - P) / dy;
    return u[c] * fx + v[c] * fy;
}

static inline double quick_cell(const double *u, const double *v, const double *f, int nx, int ny, int i, int j,
                                double dx, double dy) {
    const double a = 3.0 / 8.0, b = 6.0 / 8.0, k = 1.0 / 8.0;
    size_t c = (size_t)j * nx + i;
    const double *row = f + (size_t)j * nx;
    double P = f[c];
    double E = row[clampi(i + 1, 0, nx - 1)], W = row[clampi(i - 1, 0, nx - 1)], WW = row[clampi(i - 2, 0, nx - 1)];
    double N = f[(size_t)clampi(j + 1, 0, ny - 1) * nx + i], S = f[(size_t)clampi(j - 1, 0, ny - 1) * nx + i];
    double SS = f[(size_t)clampi(j - 2, 0, ny - 1) * nx + i];
    double uc = u[c], vc = v[c], fe, fw, fn, fs;
    if (uc >= 0) { fe = a * E + b * P - k * W; fw = a * P + b * W - k * WW; }
    else         { fe = a * W + b * P - k * E; fw = a * WW + b * W - k * P; }
    if (vc >= 0) { fn = a * N + b * P - k * S; fs = a * P + b * S - k * SS; }
    else         { fn = a * S + b * P - k * N; fs = a * SS + b * S - k * P; }
    return (uc * fe - uc * fw) / dx + (vc * fn - vc * fs) / dy;
}

void pfo_advect(int nx, int ny, int scheme, const double *u, const double *v, const double *f, double dx, double dy,
                double *out) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            out[(size_t)j * nx + i] = (scheme == 2) ? quick_cell(u, v, f, nx, ny, i, j, dx, dy)
                                                    : upwind_cell(u, v, f, nx, ny, i, j, dx, dy);
}

/* ---- pressure_solver.py:97-105: x -> -(laplacian_coherent x), flat cell 0 replaced by the identity ---- */
static inline double negA_cell(const double *x, int nx, int ny, int i, int j, double dx, double dy) {
    if (i == 0 && j == 0) return x[0];
    return -lap_cell(x, nx, ny, i, j, dx, dy);
}

/* ring-identity Jacobi (pressure_solver.py:15-40, preconditioners.py:8-25) */
static inline double invd_cell(int nx, int ny, int i, int j, double inner) {
    return (i == 0 || j == 0 || i == nx - 1 || j == ny - 1) ? 1.0 : inner;
}

/* Row kernels of the CG sweeps.  Interior cells (1 <= i <= nx-2 of rows 1 <= j <= ny-2) take a branch-free
 * loop the compiler vectorises; the expression tree is the one of lap_cell / invd_cell, so results are identical. */
static double row_rho(const double *restrict r, int nx, int ny, int j, int jacobi, double inner) {
    const double *rr = r + (size_t)j * nx;
    double s = 0.0;
    if (!jacobi || j == 0 || j == ny - 1) {
        for (int i = 0; i < nx; ++i) s += rr[i] * rr[i];            /* z = 1.0 * r = r */
        return s;
    }
    s += rr[0] * rr[0];
#pragma omp simd reduction(+ : s)
    for (int i = 1; i < nx - 1; ++i) s += rr[i] * (inner * rr[i]);
    s += rr[nx - 1] * rr[nx - 1];
    return s;
}

static void row_update_p(double *restrict p, const double *restrict r, int nx, int ny, int j, int jacobi, double inner,
                         int first, double beta) {
    double *pp = p + (size_t)j * nx;
    const double *rr = r + (size_t)j * nx;
    int ring = (!jacobi || j == 0 || j == ny - 1);
    for (int i = 0; i < nx; ++i) {
        double z = (ring || i == 0 || i == nx - 1) ? rr[i] : inner * rr[i];
        pp[i] = first ? z : pp[i] * beta + z;
    }
}

static double row_q(const double *restrict p, double *restrict q, int nx, int ny, int j, double dx, double dy) {
    const double *pc = p + (size_t)j * nx;
    double *qq = q + (size_t)j * nx;
    double s = 0.0;
    if (j == 0 || j == ny - 1) {
        for (int i = 0; i < nx; ++i) {
            double v = negA_cell(p, nx, ny, i, j, dx, dy);
            qq[i] = v;
            s += pc[i] * v;
        }
        return s;
    }
    const double *pn = pc + nx, *ps = pc - nx;
    double v0 = negA_cell(p, nx, ny, 0, j, dx, dy);
    qq[0] = v0;
    s += pc[0] * v0;
#pragma omp simd reduction(+ : s)
    for (int i = 1; i < nx - 1; ++i) {
        double xt = ((pc[i + 1] - pc[i]) / dx - (pc[i] - pc[i - 1]) / dx) / dx;
        double yt = ((pn[i] - pc[i]) / dy - (pc[i] - ps[i]) / dy) / dy;
        double v = -((0.0 + xt) + yt);
        qq[i] = v;
        s += pc[i] * v;
    }
    double v1 = negA_cell(p, nx, ny, nx - 1, j, dx, dy);
    qq[nx - 1] = v1;
    s += pc[nx - 1] * v1;
    return s;
}

static double row_update_xr(double *restrict x, double *restrict r, const double *restrict p, const double *restrict q,
                            int nx, int j, double alpha) {
    size_t o = (size_t)j * nx;
    double s = 0.0;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < nx; ++i) {
        x[o + i] = x[o + i] + alpha * p[o + i];
        double rn = r[o + i] - alpha * q[o + i];
        r[o + i] = rn;
        s += rn * rn;
    }
    return s;
}

/* scipy 1.13.1 _isolve/iterative.py cg (x0 = 0), as called from linear_solvers/interface.py:35-52.
 * work: r, p, q (3 fields) + rowsum[2*ny].  Returns info; *true_iters = iterations performed. */
int pfo_pcg(int nx, int ny, double dx, double dy, const double *b, double *x, double tol, int maxiter, int jacobi,
            double inner, double *r, double *p, double *q, double *rowsum, int *true_iters, double *bnorm_out) {
    double *rs2 = rowsum + ny;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            x[c] = 0.0;
            r[c] = b[c];
            s += b[c] * b[c];
        }
        rowsum[j] = s;
    }
    double rr = fold(rowsum, ny);
    double bnorm = sqrt(rr);
    if (bnorm_out) *bnorm_out = bnorm;
    *true_iters = 0;
    if (bnorm == 0.0) return 0;
    double atol = tol * bnorm;
    if (atol < 0.0) atol = 0.0;
    double rho_prev = 0.0;
    for (int it = 0; it < maxiter; ++it) {
        if (sqrt(rr) < atol) return 0;                              /* strict test at the top of the iteration */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ny; ++j) rowsum[j] = row_rho(r, nx, ny, j, jacobi, inner);      /* rho = r . (M r) */
        double rho = fold(rowsum, ny);
        double beta = (it > 0) ? rho / rho_prev : 0.0;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ny; ++j) row_update_p(p, r, nx, ny, j, jacobi, inner, it == 0, beta);   /* p = p*beta + z */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ny; ++j) rowsum[j] = row_q(p, q, nx, ny, j, dx, dy);            /* q = A p ; p . q */
        double alpha = rho / fold(rowsum, ny);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ny; ++j) rs2[j] = row_update_xr(x, r, p, q, nx, j, alpha);      /* x, r, r . r */
        rr = fold(rs2, ny);
        rho_prev = rho;
        *true_iters = it + 1;
    }
    return maxiter;
}

/* ---- one projection step on the interior fields (solver.py:71-154 + pressure_solver.py:107-219) ----
 * dt and the dt safeguard are host scalars computed by the caller (solver.py:24-69).
 * work: 7 fields (u_old, v_old, rhs, x, r, pv, q) followed by 2*ny doubles of row partials. */
int pfo_step(int nx, int ny, double *u, double *v, double *p, const pfo_params *prm, pfo_out *out, double *work) {
    if (nx < 2 || ny < 2) return -1;
    size_t n = (size_t)nx * ny;
    double *u_old = work, *v_old = work + n, *rhs = work + 2 * n, *x = work + 3 * n, *r = work + 4 * n,
           *pv = work + 5 * n, *q = work + 6 * n, *rowsum = work + 7 * n, *rs2 = rowsum + ny;
    const double dt = prm->dt, dx = prm->dx, dy = prm->dy;
    memcpy(u_old, u, n * sizeof(double));
    memcpy(v_old, v, n * sizeof(double));
    /* predictor, solver.py:71-95 (no diffusion term in the reference) */
    if (prm->scheme) {
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                size_t c = (size_t)j * nx + i;
                double cu = (prm->scheme == 2) ? quick_cell(u_old, v_old, u_old, nx, ny, i, j, dx, dy)
                                               : upwind_cell(u_old, v_old, u_old, nx, ny, i, j, dx, dy);
                double cv = (prm->scheme == 2) ? quick_cell(u_old, v_old, v_old, nx, ny, i, j, dx, dy)
                                               : upwind_cell(u_old, v_old, v_old, nx, ny, i, j, dx, dy);
                u[c] = u_old[c] - dt * cu;
                v[c] = v_old[c] - dt * cv;
            }
    }
    /* rhs = -div/dt, mean shift, pin (pressure_solver.py:107-128) */
    double dts = dt > 1e-14 ? dt : 1e-14;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0, t = 0.0;
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            double d = div_cell(u, v, nx, i, j, dx, dy);
            s += d * d;
            double b = -d / dts;
            rhs[c] = b;
            t += b;
        }
        rowsum[j] = s;
        rs2[j] = t;
    }
    out->div_before_norm = sqrt(fold(rowsum, ny));
    double mean = fold(rs2, ny) / (double)n;
    out->rhs_mean = mean;
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) rhs[c] = rhs[c] - mean;
    rhs[0] = 0.0;
    /* CG + true residual (interface.py:35-52) */
    int iters = 0;
    double bnorm = 0.0;
    out->info = pfo_pcg(nx, ny, dx, dy, rhs, x, prm->lin_tol, prm->lin_maxiter, prm->jacobi, prm->inv_diag_interior,
                        r, pv, q, rowsum, &iters, &bnorm);
    out->true_iters = iters;
    out->bnorm = bnorm;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            double res = rhs[c] - negA_cell(x, nx, ny, i, j, dx, dy);
            s += res * res;
        }
        rowsum[j] = s;
    }
    out->rnorm = sqrt(fold(rowsum, ny));
    /* p = x - x[0]; u -= dt*gx; v -= dt*gy (pressure_solver.py:147-157) */
    double x0 = x[0];
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < n; ++c) p[c] = x[c] - x0;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            double gx = gx_at(p + (size_t)j * nx, nx, i, dx);
            double gy = (j < ny - 1) ? (p[c + nx] - p[c]) / dy : 0.0;
            u[c] = u[c] - dt * gx;
            v[c] = v[c] - dt * gy;
        }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0;
        for (int i = 0; i < nx; ++i) {
            double d = div_cell(u, v, nx, i, j, dx, dy);
            s += d * d;
        }
        rowsum[j] = s;
    }
    out->div_after_norm = sqrt(fold(rowsum, ny));
    /* lid BC after the projection (solver.py:96-105) */
    if (prm->lid_set)
        for (int i = 0; i < nx; ++i) u[(size_t)(ny - 1) * nx + i] = prm->lid_velocity;
    /* residuals, non-finite scan (solver.py:107-154) */
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int j = 0; j < ny; ++j) {
        double s = 0.0, t = 0.0;
        for (int i = 0; i < nx; ++i) {
            size_t c = (size_t)j * nx + i;
            double du = u[c] - u_old[c], dv = v[c] - v_old[c];
            s += du * du;
            t += dv * dv;
            if (!isfinite(u[c]) || !isfinite(v[c]) || !isfinite(p[c])) bad |= 1;
        }
        rowsum[j] = s;
        rs2[j] = t;
    }
    out->ru_sumsq = fold(rowsum, ny);
    out->rv_sumsq = fold(rs2, ny);
    out->nonfinite = bad;
    out->cont_sumsq = pfo_divc_sumsq(nx, ny, u, v, 1.0, 1.0, rowsum);
    out->reserved = 0;
    return 0;
}
