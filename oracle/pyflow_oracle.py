"""CPU oracle for the PyFlow projection time step (TEST INFRASTRUCTURE ONLY).

This file is a numpy float64 restatement of the reference hot path
(beautifulplanet/PyFlow-Solver).  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``pyflow-solver_b200/``) never does, and has no CPU fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference/src`` (possible only in the build container),
checks every function below bitwise against it (operators, advection, Jacobi
diagonal, CG recurrence, whole multi-step trajectories) and writes the golden
vectors under ``tests/golden/``.  The reference ships no known-answer vectors
of its own, so those generated fixtures plus the
survey's measured KATs are the pins.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).  Arithmetic order is kept identical to numpy's so
that all element-wise results are bit-exact; only the dot products / norms
(BLAS ``ddot``) depend on the machine's summation order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

__all__ = [
    "OracleState", "allocate_state", "interior",
    "gradient_coherent", "divergence_coherent", "laplacian_coherent",
    "divergence_central", "gradient_central", "laplacian_central", "pressure_correction_central",
    "advect_upwind", "advect_quick",
    "jacobi_inverse_diagonal", "neg_laplacian_pinned", "pcg",
    "solve_pressure_poisson", "solve_manufactured", "step", "run",
]


# --------------------------------------------------------------------------
# State container -- src/pyflow/core/ghost_fields.py:8-60
# --------------------------------------------------------------------------
@dataclass
class OracleState:
    nx: int
    ny: int
    fields: dict[str, np.ndarray]
    meta: dict[str, Any] = field(default_factory=dict)


def allocate_state(nx: int, ny: int) -> OracleState:
    """(ny+2, nx+2) float64 zero fields with an untouched ghost ring
    (core/ghost_fields.py:47-57)."""
    return OracleState(nx, ny, {k: np.zeros((ny + 2, nx + 2)) for k in ("u", "v", "p")})


def interior(a: np.ndarray) -> np.ndarray:
    """core/ghost_fields.py:59-60."""
    return a[1:-1, 1:-1]


# --------------------------------------------------------------------------
# Coherent operator trio -- src/pyflow/numerics/core_ops_coherent.py
# --------------------------------------------------------------------------
def gradient_coherent(p, dx, dy):
    """Forward differences, zero on the high faces (core_ops_coherent.py:45-61)."""
    ny, nx = p.shape
    gx = np.zeros_like(p)
    gy = np.zeros_like(p)
    if ny < 2 or nx < 2:
        return gx, gy
    np.divide(p[:, 1:] - p[:, :-1], dx, out=gx[:, :-1])
    np.divide(p[1:, :] - p[:-1, :], dy, out=gy[:-1, :])
    return gx, gy


def divergence_coherent(u, v, dx, dy):
    """Backward differences; the x term is added to zero first, then the y
    term (core_ops_coherent.py:63-79)."""
    ny, nx = u.shape
    out = np.zeros_like(u)
    if ny < 2 or nx < 2:
        return out
    xt = np.empty_like(u)
    xt[:, 0] = u[:, 0] / dx
    xt[:, 1:] = (u[:, 1:] - u[:, :-1]) / dx
    yt = np.empty_like(u)
    yt[0, :] = v[0, :] / dy
    yt[1:, :] = (v[1:, :] - v[:-1, :]) / dy
    out += xt          # 0 + xterm
    out += yt          # (0 + xterm) + yterm
    return out


def laplacian_coherent(p, dx, dy):
    """div_coherent(grad_coherent(p)) (core_ops_coherent.py:81-84)."""
    gx, gy = gradient_coherent(p, dx, dy)
    return divergence_coherent(gx, gy, dx, dy)


# --------------------------------------------------------------------------
# Central divergence used only for diagnostics / the dt safeguard
# src/pyflow/numerics/fluid_ops.py:8-28
# --------------------------------------------------------------------------
def divergence_central(u, v, dx, dy):
    ny, nx = u.shape
    d = np.zeros_like(u)
    tdx = 2.0 * dx
    tdy = 2.0 * dy
    # x part / y part per region, summed in the reference's order (x then y)
    d[1:-1, 1:-1] = (u[1:-1, 2:] - u[1:-1, :-2]) / tdx + (v[2:, 1:-1] - v[:-2, 1:-1]) / tdy
    d[1:-1, 0] = (u[1:-1, 1] - u[1:-1, 0]) / dx + (v[2:, 0] - v[:-2, 0]) / tdy
    d[0, 1:-1] = (u[0, 2:] - u[0, :-2]) / tdx + (v[1, 1:-1] - v[0, 1:-1]) / dy
    d[-1, 1:-1] = (u[-1, 2:] - u[-1, :-2]) / tdx + (v[-1, 1:-1] - v[-2, 1:-1]) / dy
    d[0, 0] = (u[0, 1] - u[0, 0]) / dx + (v[1, 0] - v[0, 0]) / dy
    d[0, -1] = (u[0, -1] - u[0, -2]) / dx + (v[1, -1] - v[0, -1]) / dy
    d[-1, 0] = (u[-1, 1] - u[-1, 0]) / dx + (v[-1, 0] - v[-2, 0]) / dy
    d[-1, -1] = (u[-1, -1] - u[-1, -2]) / dx + (v[-1, -1] - v[-2, -1]) / dy
    # NOTE (fluid_ops.py:16-27): the right column excluding corners is never
    # written by the reference and stays zero.  Reproduced here on purpose.
    return d


# --------------------------------------------------------------------------
# Legacy central gradient / Laplacian / correction of the same module
# src/pyflow/numerics/fluid_ops.py:43-67, 69-91, 108-125 -- off the time-step
# path, but imported by the reference's own projection tests
# (tests/test_projection_method_contract.py:304,337,409)
# --------------------------------------------------------------------------
def gradient_central(p, dx, dy):
    """fluid_ops.gradient (:43-67): difference of the two clamped neighbours over
    ``2*spacing`` inside and over ``spacing`` in the first / last column or row."""
    ny, nx = p.shape
    ii = np.arange(nx)
    jj = np.arange(ny)
    span_x = np.where((ii > 0) & (ii < nx - 1), 2.0 * dx, dx)
    span_y = np.where((jj > 0) & (jj < ny - 1), 2.0 * dy, dy)
    gx = (p[:, np.minimum(ii + 1, nx - 1)] - p[:, np.maximum(ii - 1, 0)]) / span_x[None, :]
    gy = (p[np.minimum(jj + 1, ny - 1), :] - p[np.maximum(jj - 1, 0), :]) / span_y[:, None]
    return gx, gy


def laplacian_central(p, dx, dy):
    """fluid_ops.laplacian (:69-91): 5-point second differences inside; every wall
    cell copies the value of its nearest interior cell (rows first, then columns,
    so corners take the diagonal neighbour); all zero unless nx > 2 and ny > 2."""
    ny, nx = p.shape
    if nx <= 2 or ny <= 2:
        return np.zeros_like(p)
    c = p[1:-1, 1:-1]
    core = (p[1:-1, 2:] - 2.0 * c + p[1:-1, :-2]) / (dx * dx) + (p[2:, 1:-1] - 2.0 * c + p[:-2, 1:-1]) / (dy * dy)
    jj = np.clip(np.arange(ny), 1, ny - 2) - 1
    ii = np.clip(np.arange(nx), 1, nx - 2) - 1
    return core[np.ix_(jj, ii)]


def pressure_correction_central(u, v, p, dt, dx, dy):
    """fluid_ops.apply_pressure_correction (:108-125) on interior arrays:
    returns (u - dt*dp/dx, v - dt*dp/dy) with the central gradient above."""
    gx, gy = gradient_central(p, dx, dy)
    return u - dt * gx, v - dt * gy


# --------------------------------------------------------------------------
# Advection -- src/pyflow/numerics/operators/advection.py
# (the reference runs Python double loops; these are vectorised forms with the
#  same per-cell expression tree, verified bitwise by make_golden.py)
# --------------------------------------------------------------------------
def advect_upwind(u, v, f, dx, dy):
    """First-order upwind, strict ``> 0`` test, clamped neighbour at the walls
    (advection.py:17-43)."""
    ny, nx = f.shape
    ii = np.arange(nx)
    jj = np.arange(ny)
    west = f[:, np.maximum(ii - 1, 0)]
    east = f[:, np.minimum(ii + 1, nx - 1)]
    south = f[np.maximum(jj - 1, 0), :]
    north = f[np.minimum(jj + 1, ny - 1), :]
    fx = np.where(u > 0, (f - west) / dx, (east - f) / dx)
    fy = np.where(v > 0, (f - south) / dy, (north - f) / dy)
    return u * fx + v * fy


def advect_quick(u, v, f, dx, dy):
    """The reference's QUICK variant: cell-centred velocity for both faces,
    ``>= 0`` test, edge padding by 2, literal negative-velocity branch
    (advection.py:45-84)."""
    ny, nx = f.shape
    ph = np.pad(f, 2, mode="edge")
    P = ph[2:-2, 2:-2]
    E = ph[2:-2, 3:nx + 3]
    W = ph[2:-2, 1:nx + 1]
    WW = ph[2:-2, 0:nx]
    N = ph[3:ny + 3, 2:-2]
    S = ph[1:ny + 1, 2:-2]
    SS = ph[0:ny, 2:-2]
    a, b, c = 3 / 8, 6 / 8, 1 / 8
    upos = u >= 0
    vpos = v >= 0
    fe = np.where(upos, a * E + b * P - c * W, a * W + b * P - c * E)
    fw = np.where(upos, a * P + b * W - c * WW, a * WW + b * W - c * P)
    fn = np.where(vpos, a * N + b * P - c * S, a * S + b * P - c * N)
    fs = np.where(vpos, a * P + b * S - c * SS, a * SS + b * S - c * P)
    return (u * fe - u * fw) / dx + (v * fn - v * fs) / dy


# --------------------------------------------------------------------------
# Jacobi preconditioner -- pressure_solver.py:15-40 + preconditioners.py:8-25
# The reference assembles a lil_matrix every step only to read its diagonal:
# identity rows on the whole boundary ring, 2/dx^2 + 2/dy^2 inside.
# --------------------------------------------------------------------------
def jacobi_inverse_diagonal(nx, ny, dx, dy):
    d = np.full((ny, nx), 2.0 / dx ** 2 + 2.0 / dy ** 2)
    d[0, :] = 1.0
    d[-1, :] = 1.0
    d[:, 0] = 1.0
    d[:, -1] = 1.0
    d = d.reshape(-1)
    return np.where(d != 0, 1.0 / d, 0.0)


def neg_laplacian_pinned(xvec, nx, ny, dx, dy):
    """_NegLaplacianPinned._matvec (pressure_solver.py:97-105)."""
    out = (-laplacian_coherent(xvec.reshape(ny, nx), dx, dy)).reshape(-1)
    out[0] = xvec[0]
    return out


# --------------------------------------------------------------------------
# Preconditioned CG -- scipy.sparse.linalg.cg (scipy==1.13.1 pinned by the
# reference, pyproject.toml:20; un-vendored third party; the recurrence below
# follows scipy/sparse/linalg/_isolve/iterative.py `cg`), called from
# linear_solvers/interface.py:35-52.
# --------------------------------------------------------------------------
def pcg(matvec, b, tol, maxiter, inv_diag=None, dot=np.dot):
    """Returns (x, info, true_iterations).  x0 = 0; stop test is the strict
    ``||r|| < max(0, tol*||b||)`` at the TOP of each iteration; no test after
    the last one (info = maxiter then)."""
    x = np.zeros_like(b)
    bnorm = float(np.sqrt(dot(b, b)))
    atol = max(0.0, float(tol) * bnorm)
    if bnorm == 0:
        return b, 0, 0
    r = b.copy()
    rho_prev = None
    p = None
    iters = 0
    for it in range(maxiter):
        if np.sqrt(dot(r, r)) < atol:
            return x, 0, iters
        z = r if inv_diag is None else inv_diag * r
        rho = dot(r, z)
        if it > 0:
            p *= rho / rho_prev
            p += z
        else:
            p = z.copy()
        q = matvec(p)
        alpha = rho / dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
        iters += 1
    return x, maxiter, iters


def bicgstab(matvec, b, tol, maxiter, inv_diag=None, dot=np.dot):
    """scipy.sparse.linalg.bicgstab (scipy/sparse/linalg/_isolve/iterative.py; third-party, pinned
    scipy==1.13.1 in the reference's pyproject.toml:20, same loop in the 1.18 installed here) with x0 = 0, the
    way ``linear_solvers/interface.py:33,44-45`` calls it for ``method='bicgstab'`` (the default of ``solve``).
    Restated statement for statement; pinned bitwise against scipy itself in tests/test_oracle_golden.py.
    Returns (x, info, completed_iterations); info: 0 converged, maxiter exhausted, -10 / -11 breakdown."""
    bnorm = float(np.sqrt(dot(b, b)))
    atol = max(0.0, float(tol) * bnorm)
    if bnorm == 0:
        return b, 0, 0
    x = np.zeros_like(b)
    rhotol = np.finfo(b.dtype.char).eps ** 2
    omegatol = rhotol
    rho_prev = omega = alpha = p = v = None
    r = b.copy()
    rtilde = r.copy()
    s = np.empty_like(r)
    iters = 0
    for it in range(maxiter):
        if np.sqrt(dot(r, r)) < atol:
            return x, 0, iters
        rho = dot(rtilde, r)
        if np.abs(rho) < rhotol:
            return x, -10, iters
        if it > 0:
            if np.abs(omega) < omegatol:
                return x, -11, iters
            beta = (rho / rho_prev) * (alpha / omega)
            p -= omega * v
            p *= beta
            p += r
        else:
            p = r.copy()
        phat = p if inv_diag is None else inv_diag * p
        v = matvec(phat)
        rv = dot(rtilde, v)
        if rv == 0:
            return x, -11, iters
        alpha = rho / rv
        r -= alpha * v
        s[:] = r[:]
        if np.sqrt(dot(s, s)) < atol:
            x += alpha * phat
            return x, 0, iters
        shat = s if inv_diag is None else inv_diag * s
        t = matvec(shat)
        omega = dot(t, s) / dot(t, t)
        x += alpha * phat
        x += omega * shat
        r -= omega * t
        rho_prev = rho
        iters += 1
    return x, maxiter, iters


def interface_solve(b, nx, ny, dx, dy, method="bicgstab", tol=1e-8, maxiter=500, jacobi=False, dot=np.dot):
    """``linear_solvers/interface.py:solve`` (:23-53) on the pinned operator: the Krylov solve, then
    ``rnorm = ||b - A x||``, ``converged = (info == 0 and rnorm <= tol*||b||)``, ``iterations = info if info > 0
    else 0``.  Returns (x, converged, iterations, rnorm, info)."""
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    A = lambda t: neg_laplacian_pinned(t, nx, ny, dx, dy)                        # noqa: E731
    M = jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    x, info, _ = (pcg if method == "cg" else bicgstab)(A, b, tol, maxiter, M, dot)
    rnorm = float(np.linalg.norm(b - A(x)))
    bnorm = float(np.linalg.norm(b)) or 1.0
    return x, bool(info == 0 and rnorm <= tol * bnorm), (info if info > 0 else 0), rnorm, info


# --------------------------------------------------------------------------
# Pressure projection -- src/pyflow/solvers/pressure_solver.py:42-219
# (normal projection mode only; manufactured mode is SURVEY 8f "next")
# --------------------------------------------------------------------------
def solve_pressure_poisson(state, dt, dx, dy, cfg, dot=np.dot):
    ui = interior(state.fields["u"])
    vi = interior(state.fields["v"])
    ny, nx = ui.shape
    div_u = divergence_coherent(ui, vi, dx, dy)
    div_before = float(np.sqrt(dot(div_u.ravel(), div_u.ravel())))
    rhs = (-div_u / max(dt, 1e-14)).reshape(-1)
    rhs -= float(np.mean(rhs))
    rhs[0] = 0.0
    tol = getattr(cfg, "lin_tol", 1e-10)
    maxiter = getattr(cfg, "lin_maxiter", 400)
    inv_diag = jacobi_inverse_diagonal(nx, ny, dx, dy) if getattr(cfg, "enable_jacobi_pc", True) else None

    def A(xv):
        return neg_laplacian_pinned(xv, nx, ny, dx, dy)

    x, info, true_iters = pcg(A, rhs, tol, maxiter, inv_diag, dot)
    # interface.py:48-52
    res = rhs - A(x)
    rnorm = float(np.sqrt(dot(res, res)))
    bnorm = float(np.sqrt(dot(rhs, rhs))) or 1.0
    converged = (info == 0) and (rnorm <= tol * bnorm)
    iterations = info if info > 0 else 0
    # pressure_solver.py:147-157
    p = x.reshape(ny, nx).copy()
    p -= p.flat[0]
    interior(state.fields["p"])[:] = p
    gx, gy = gradient_coherent(interior(state.fields["p"]), dx, dy)
    ui[:] -= dt * gx
    vi[:] -= dt * gy
    d_after = divergence_coherent(ui, vi, dx, dy)
    div_norm = float(np.sqrt(dot(d_after.ravel(), d_after.ravel())))
    # pressure_solver.py:185-194
    if iterations == 0 and rnorm > tol:
        iterations = 1
    diag = {
        "Rp_iterations": iterations, "Rp_residual": rnorm, "Rp_converged": converged,
        "Rp_method": "cg", "divergence_norm": div_norm,
        "divergence_norm_initial": div_before,
        "true_iterations": true_iters, "info": info,
    }
    return p, diag


def solve_manufactured(state, dx, dy, cfg, rhs_override, dot=np.dot):
    """Manufactured / Dirichlet verification mode (pressure_solver.py:61-93):
    A x = +laplacian_coherent(x) with identity rows on the boundary ring, no
    automatic Jacobi, velocities untouched."""
    ny, nx = interior(state.fields["u"]).shape
    mask = np.zeros((ny, nx), dtype=bool)
    mask[0, :] = mask[-1, :] = True
    mask[:, 0] = mask[:, -1] = True
    ring = np.flatnonzero(mask.reshape(-1))

    def A(xv):
        out = laplacian_coherent(xv.reshape(ny, nx), dx, dy).reshape(-1)
        out[ring] = xv[ring]
        return out

    rhs = np.copy(rhs_override).reshape(-1)
    tol = getattr(cfg, "lin_tol", 1e-10)
    maxiter = getattr(cfg, "lin_maxiter", 500)
    x, info, true_iters = pcg(A, rhs, tol, maxiter, None, dot)
    res = rhs - A(x)
    rnorm = float(np.sqrt(dot(res, res)))
    bnorm = float(np.sqrt(dot(rhs, rhs))) or 1.0
    p = x.reshape(ny, nx)
    interior(state.fields["p"])[:] = p
    iterations = info if info > 0 else 0
    if iterations == 0 and rnorm > tol:
        iterations = 1
    return p, {"Rp_iterations": iterations, "Rp_residual": rnorm, "Rp_converged": (info == 0) and (rnorm <= tol * bnorm),
               "Rp_method": "cg", "divergence_norm": rnorm, "true_iterations": true_iters, "info": info}


# --------------------------------------------------------------------------
# One projection time step -- src/pyflow/solvers/solver.py:15-193
# --------------------------------------------------------------------------
def step(cfg, state, iteration, dot=np.dot):
    u_i = interior(state.fields["u"])
    v_i = interior(state.fields["v"])
    ny, nx = u_i.shape
    lx = getattr(cfg, "lx", float(nx - 1))
    ly = getattr(cfg, "ly", float(ny - 1))
    dx = lx / max(nx - 1, 1)
    dy = ly / max(ny - 1, 1)
    umax = float(np.max(np.abs(u_i))) + 1e-12
    vmax = float(np.max(np.abs(v_i))) + 1e-12
    cfl = getattr(cfg, "cfl_target", 0.5)
    growth = getattr(cfg, "cfl_growth", 1.05)
    Re = float(getattr(cfg, "Re", 100.0))
    nu = 1.0 / Re
    dt_prev = float(state.meta.get("dt_prev", 1e-2))
    dt_cfl = min(cfl * dx / umax, cfl * dy / vmax)
    dt_diff = 0.25 * min(dx * dx, dy * dy) / max(nu, 1e-12)
    dt_cand = min(dt_cfl, dt_diff)
    dt = dt_prev * growth if dt_cand > dt_prev * growth else dt_cand
    dc = divergence_central(u_i, v_i, 1.0, 1.0).ravel()
    hist = state.meta.get("divergence_history", [])
    hist.append(float(np.sqrt(dot(dc, dc))))
    hist = hist[-5:]
    state.meta["divergence_history"] = hist
    if len(hist) >= 3:
        a, b, c = hist[-3], hist[-2], hist[-1]
        if a > 0 and b > a * 1.01 and c > b * 1.01:
            dt *= 0.5
    state.meta["dt_prev"] = dt
    state.meta["dx"] = dx
    state.meta["dy"] = dy
    state.meta["nu"] = nu
    cfl_now = max(umax * dt / dx, vmax * dt / dy)
    u_old = u_i.copy()
    v_old = v_i.copy()
    if not getattr(cfg, "disable_advection", False):
        adv = advect_quick if getattr(cfg, "advection_scheme", "upwind") == "quick" else advect_upwind
        cu = adv(u_old, v_old, u_old, dx, dy)
        cv = adv(u_old, v_old, v_old, dx, dy)
        u_i[:, :] = u_old - dt * cu
        v_i[:, :] = v_old - dt * cv
    _, pdiag = solve_pressure_poisson(state, dt, dx, dy, cfg, dot)
    state.meta["prev_div_norm"] = pdiag["divergence_norm"]
    lid = getattr(cfg, "lid_velocity", None)
    if lid is not None:
        state.fields["u"][-2, 1:-1] = lid
    du = (u_i - u_old).ravel()
    dv = (v_i - v_old).ravel()
    Ru = float(np.sqrt(np.sum(du * du)) / max(1, du.size))
    Rv = float(np.sqrt(np.sum(dv * dv)) / max(1, dv.size))
    dc = divergence_central(u_i, v_i, 1.0, 1.0).ravel()
    continuity = float(np.sqrt(dot(dc, dc)) / max(1.0, nx * ny))
    nan_detected = not all(np.isfinite(interior(state.fields[k])).all() for k in ("u", "v", "p"))
    residuals = {"Ru": Ru, "Rv": Rv, "Rp": 0.0, "continuity": continuity}
    diagnostics = {
        "iteration": iteration, "dt": dt, "CFL": cfl_now, "nu": nu,
        "Rp_iterations": pdiag["Rp_iterations"], "Rp_residual": pdiag["Rp_residual"],
        "Rp_converged": pdiag["Rp_converged"], "Rp_method": "cg",
        "divergence_norm": pdiag["divergence_norm"], "nan_detected": nan_detected,
        "true_iterations": pdiag["true_iterations"],
        "max_residual": max(residuals.values()),
        "residual_drop_Ru_orders": max(0.0, (0.0 - math.log10(Ru)) if Ru > 0 else 0.0),
    }
    return state, residuals, diagnostics


def run(cfg, steps, state=None, dot=np.dot):
    """Plain loop equivalent of SimulationDriver.run without the I/O
    (drivers/simulation_driver.py:58-110).  Returns (state, [diagnostics])."""
    if state is None:
        state = allocate_state(cfg.nx, cfg.ny)
    out = []
    for it in range(steps):
        _, res, diag = step(cfg, state, it, dot)
        diag = dict(diag)
        diag.update(res)
        out.append(diag)
    return state, out


def pairwise_dot(a, b):
    """Alternative summation order (numpy pairwise) used to measure the
    oracle's own noise floor against BLAS ddot."""
    return float(np.sum(a * b))
