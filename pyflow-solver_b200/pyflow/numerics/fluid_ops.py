"""Central-difference operators (drop-in for the reference's
``src/pyflow/numerics/fluid_ops.py``), every one a device kernel behind the
C-ABI -- there is no numpy path.

On the time-step path: ``divergence`` (:30, ``_divergence_impl`` :8-28), used by
the dt safeguard and the ``continuity`` residual.  Off the path but part of the
module's surface (the reference's own projection tests import them,
tests/test_projection_method_contract.py:304,337,409): ``gradient`` :43-67,
``laplacian`` :69-91, ``legacy_divergence_state`` :93, ``legacy_grad_p_state``
:101, ``apply_pressure_correction`` :108-125.
"""
from __future__ import annotations

import numpy as np

from .._b200.device import slab_for_shape
from ..core.ghost_fields import State, interior_view


def _shape(a: np.ndarray, what: str) -> tuple[int, int]:
    ny, nx = a.shape
    if ny < 2 or nx < 2:
        raise ValueError(f"{what} needs at least a 2x2 grid")
    return ny, nx


def divergence(u: np.ndarray, v: np.ndarray, dx: float, dy: float) -> np.ndarray:
    """Central differences inside, one-sided on the edges; the right column
    (corners excluded) stays zero exactly as in the reference (:16-27)."""
    ny, nx = _shape(u, "divergence")
    return slab_for_shape(ny, nx).divergence_central(u, v, dx, dy)


_divergence_impl = divergence


def gradient(p: np.ndarray, dx: float, dy: float) -> tuple[np.ndarray, np.ndarray]:
    """Central differences inside, one-sided in the first / last column and row (:43-67)."""
    ny, nx = _shape(p, "gradient")
    return slab_for_shape(ny, nx).gradient_central(p, dx, dy)


def laplacian(p: np.ndarray, dx: float, dy: float) -> np.ndarray:
    """5-point second differences inside; wall cells copy the nearest interior value (:69-91)."""
    ny, nx = _shape(p, "laplacian")
    return slab_for_shape(ny, nx).laplacian_central(p, dx, dy)


def legacy_divergence_state(state: State, dx: float = 1.0, dy: float = 1.0) -> np.ndarray:
    return divergence(interior_view(state.fields["u"]), interior_view(state.fields["v"]), dx, dy)


def legacy_grad_p_state(state: State, dx: float = 1.0, dy: float = 1.0):
    return gradient(interior_view(state.fields["p"]), dx, dy)


def apply_pressure_correction(state: State, p: np.ndarray, dt: float, dx: float, dy: float) -> None:
    """``state.p <- p``; ``u -= dt*dp/dx``; ``v -= dt*dp/dy`` with the central gradient (:108-125).
    The subtraction runs on the device as well; the host arrays are updated in place."""
    ui = interior_view(state.fields["u"])
    vi = interior_view(state.fields["v"])
    ny, nx = _shape(ui, "apply_pressure_correction")
    interior_view(state.fields["p"])[:] = p
    un, vn = slab_for_shape(ny, nx).pressure_correction_central(ui, vi, p, dt, dx, dy)
    ui[:] = un
    vi[:] = vn
