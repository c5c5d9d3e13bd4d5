"""Pressure projection, device-backed (drop-in for the reference's
``src/pyflow/solvers/pressure_solver.py``).

``solve_pressure_poisson`` keeps the reference signature and diagnostics keys
(:42, :203-219).  Normal projection mode (:94-169) runs entirely on the GPU:
RHS ``-(div u*)/dt`` with mean shift and pin, matrix-free pinned Neumann
operator, ring-identity Jacobi PCG with scipy's recurrence, correction,
divergence norms, true residual.

The Jacobi preconditioner of the reference is the inverse diagonal of a matrix
it assembles with a Python double loop EVERY step (:15-40, :139-143; 60-67 % of
its wall time).  That diagonal is analytic -- 1 on the boundary ring,
``2/dx^2 + 2/dy^2`` inside -- so the device applies it in-kernel; only the one
interior scalar is computed here, with the reference's float expression.
"""
from __future__ import annotations

import numpy as np

from .._b200 import lib as _lib
from .._b200.engine import attach
from ..core.ghost_fields import interior_view


def assemble_negative_laplacian(nx: int, ny: int, dx: float, dy: float):
    """Host utility kept for API compatibility (reference :15-40): CSR matrix
    with identity rows on the whole boundary ring and the 5-point -Laplacian
    inside.  Vectorised; NOT used by the device path."""
    import scipy.sparse as sp
    N = nx * ny
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    idx = (jj * nx + ii).ravel()
    ring = ((ii == 0) | (ii == nx - 1) | (jj == 0) | (jj == ny - 1)).ravel()
    inner = idx[~ring]
    cx, cy = 1.0 / dx ** 2, 1.0 / dy ** 2
    rows = [idx[ring], inner, inner, inner, inner, inner]
    cols = [idx[ring], inner, inner - 1, inner + 1, inner - nx, inner + nx]
    vals = [np.ones(int(ring.sum())), np.full(inner.size, 2.0 / dx ** 2 + 2.0 / dy ** 2),
            np.full(inner.size, -cx), np.full(inner.size, -cx), np.full(inner.size, -cy), np.full(inner.size, -cy)]
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    return A.tocsr()


def _projection_params(cfg, dt, dx, dy, preconditioner=None, default_maxiter=400, grid=None) -> _lib.StepParams:
    inv_diag_interior = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)      # reference :35 and preconditioners.py:19
    if preconditioner is not None:
        from ..linear_solvers.preconditioners import resolve
        nx, ny = grid if grid is not None else (getattr(preconditioner, "nx", 0), getattr(preconditioner, "ny", 0))
        inv_diag_interior = resolve(preconditioner, nx, ny).inv_diag_interior
        jacobi = 1
    else:
        jacobi = 1 if getattr(cfg, "enable_jacobi_pc", True) else 0
    p = _lib.StepParams()
    p.dt, p.dx, p.dy = float(dt), float(dx), float(dy)
    p.lin_tol = float(getattr(cfg, "lin_tol", 1e-10))
    p.lin_maxiter = int(getattr(cfg, "lin_maxiter", default_maxiter))
    p.jacobi = jacobi
    p.inv_diag_interior = inv_diag_interior
    p.scheme = _lib.PF_ADV_NONE
    p.projection_only = 1
    return p


def _rp_diagnostics(cfg, out: _lib.StepOut) -> dict:
    """interface.py:48-52 + pressure_solver.py:185-194 on the device scalars."""
    lin_tol = getattr(cfg, "lin_tol", 1e-10)
    rnorm = float(out.rnorm)
    bnorm = float(out.bnorm) or 1.0
    converged = (out.info == 0) and (rnorm <= lin_tol * bnorm)
    iterations = int(out.info) if out.info > 0 else 0
    if iterations == 0 and rnorm > lin_tol:
        iterations = 1
    return {
        "Rp_iterations": iterations,
        "Rp_residual": rnorm,
        "Rp_converged": bool(converged),
        "Rp_method": "cg",
        "divergence_norm": float(out.div_after_norm),
    }


def _solve_manufactured(state, dx, dy, cfg, preconditioner, rhs_override):
    """Manufactured / Dirichlet verification mode (reference :61-93): solve
    ``laplacian_coherent(p) = rhs`` with identity rows on the boundary ring, no
    automatic Jacobi, velocities untouched; ``divergence_norm`` is the linear
    residual ``||A p - rhs||``.  Runs in the plain device CG (off the hot path)."""
    if preconditioner is not None:
        raise NotImplementedError("preconditioners in manufactured mode are not supported on the device")
    bind = attach(state)
    slab = bind.slab
    rhs = np.ascontiguousarray(np.copy(rhs_override), dtype=np.float64).reshape(slab.ny_local, slab.nx)
    lin_tol = getattr(cfg, "lin_tol", 1e-10)
    lin_maxiter = getattr(cfg, "lin_maxiter", 500)
    p, info, iters, rnorm, bnorm = slab.cg_generic("dirichlet", rhs, dx, dy, lin_tol, lin_maxiter)
    bind.finish(prestep_next=None)            # the device p field holds the solution; u, v unchanged
    converged = (info == 0) and (rnorm <= lin_tol * (bnorm or 1.0))
    rp_iters = info if info > 0 else 0
    if rp_iters == 0 and rnorm > lin_tol:
        rp_iters = 1
    rhs_norm = bnorm or 1.0
    state.meta["cg_true_iterations"] = int(iters)
    diag = {
        "Rp_iterations": rp_iters, "Rp_residual": float(rnorm), "Rp_converged": bool(converged), "Rp_method": "cg",
        "divergence_norm": float(rnorm), "lin_residual_norm": float(rnorm), "lin_residual_rel": float(rnorm) / rhs_norm,
        "divergence_reduction_factor": None, "divergence_norm_initial": None, "divergence_norm_after_raw": None,
        "divergence_norm_final": float(rnorm),
    }
    return p, diag


def solve_pressure_poisson(state, dt: float, dx: float, dy: float, cfg, preconditioner=None, rhs_override=None):
    """Project ``state``'s (u, v) to reduce their coherent divergence.

    Returns ``(p, diagnostics)`` like the reference (:42-219); mutates the
    state's u, v, p.  ``rhs_override`` selects the manufactured
    Dirichlet verification mode (:61-93).
    """
    if rhs_override is not None:
        return _solve_manufactured(state, dx, dy, cfg, preconditioner, rhs_override)
    verbose = getattr(cfg, "diagnostics", True) and not getattr(cfg, "force_quiet", False)
    bind = attach(state)
    params = _projection_params(cfg, dt, dx, dy, preconditioner, grid=(bind.slab.nx, bind.slab.ny_global))
    out = bind.slab.step(params)
    bind.finish(prestep_next=None)
    diag = _rp_diagnostics(cfg, out)
    div_before = float(out.div_before_norm)
    div_norm = diag["divergence_norm"]
    rhs_norm = float(out.bnorm) or 1.0
    # ||A p - rhs|| == ||rhs - A p|| term by term (reference :174-183 vs interface.py:48)
    diag["lin_residual_norm"] = float(out.rnorm)
    diag["lin_residual_rel"] = float(out.rnorm) / rhs_norm
    diag["divergence_reduction_factor"] = (div_before / div_norm) if div_norm not in (None, 0.0) else None
    diag["divergence_norm_initial"] = div_before
    diag["divergence_norm_after_raw"] = div_norm
    diag["divergence_norm_final"] = div_norm
    if verbose:
        print("=== PRESSURE PROJECTION DIAGNOSTICS ===")
        print(f"DIVERGENCE_NORM_BEFORE={div_before}")
        print(f"[pressure_solver] RHS mean subtracted (compatibility) = {out.rhs_mean:.6e}")
        print(f"DIVERGENCE_NORM_AFTER_RAW={div_norm}")
        print(f"DIVERGENCE_NORM_FINAL={div_norm}")
        print(f"LIN_RESIDUAL_NORM={diag['lin_residual_norm']}")
        print(f"LIN_RESIDUAL_REL={diag['lin_residual_rel']}")
    p = interior_view(state.fields["p"]).copy()
    state.meta["cg_true_iterations"] = int(out.true_iters)
    return p, diag
