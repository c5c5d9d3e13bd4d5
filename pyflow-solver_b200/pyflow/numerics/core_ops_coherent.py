"""Coherent gradient / divergence / Laplacian on the GPU (drop-in for the
reference's ``src/pyflow/numerics/core_ops_coherent.py`` :45, :63, :81).

Pure functions on ``(ny, nx)`` float64 arrays returning fresh arrays, as in the
reference; the arithmetic runs in the ``pf_op_*`` kernels of libpyflow_b200
(exact numpy operation order, bit-identical results).  Degenerate grids
(``nx < 2`` or ``ny < 2``) return zeros as the reference does (:52-53, :71-72).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .._b200.device import slab_for_shape

Array2D = np.ndarray

__all__ = ["coherence_diagnostics", "divergence_coherent", "gradient_coherent", "laplacian_coherent"]


def _degenerate(a: Array2D) -> bool:
    ny, nx = a.shape
    return ny < 2 or nx < 2


def gradient_coherent(p: Array2D, dx: float, dy: float) -> tuple[Array2D, Array2D]:
    """Forward differences, zero on the high (Neumann) faces."""
    if _degenerate(p):
        return np.zeros_like(p), np.zeros_like(p)
    return slab_for_shape(*p.shape).gradient_coherent(p, dx, dy)


def divergence_coherent(u: Array2D, v: Array2D, dx: float, dy: float) -> Array2D:
    """Backward differences paired with ``gradient_coherent``."""
    if _degenerate(u):
        return np.zeros_like(u)
    return slab_for_shape(*u.shape).divergence_coherent(u, v, dx, dy)


def laplacian_coherent(p: Array2D, dx: float, dy: float) -> Array2D:
    """``divergence_coherent(*gradient_coherent(p))`` fused in one kernel."""
    if _degenerate(p):
        return np.zeros_like(p)
    return slab_for_shape(*p.shape).laplacian_coherent(p, dx, dy)


@dataclass(slots=True)
class CoherenceDiagnostics:
    max_abs_diff: float
    l2_rel_diff: float
    lap_norm: float
    divgrad_norm: float
    constant_laplacian_max: float
    mean_laplacian: float


def coherence_diagnostics(nx: int, ny: int, dx: float, dy: float, seed: int = 0) -> CoherenceDiagnostics:
    """Fused Laplacian vs the explicit composition on a random field (the
    reference's version compares the Laplacian with itself, :96-119; here the
    two device code paths are really compared)."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((ny, nx))
    lap = laplacian_coherent(p, dx, dy)
    divgrad = divergence_coherent(*gradient_coherent(p, dx, dy), dx, dy)
    diff = lap - divgrad
    lap_norm = float(np.linalg.norm(lap)) or 1.0
    const = laplacian_coherent(np.full((ny, nx), 7.123), dx, dy)
    return CoherenceDiagnostics(
        max_abs_diff=float(np.max(np.abs(diff))),
        l2_rel_diff=float(np.linalg.norm(diff)) / lap_norm,
        lap_norm=lap_norm,
        divgrad_norm=float(np.linalg.norm(divgrad)),
        constant_laplacian_max=float(np.max(np.abs(const))),
        mean_laplacian=float(np.mean(lap)),
    )
