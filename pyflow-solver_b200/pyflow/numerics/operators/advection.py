"""Convective derivative on the GPU (drop-in for the reference's
``src/pyflow/numerics/operators/advection.py``: ``advect_upwind`` :17-43,
``advect_quick`` :45-84, which are Python double loops).

Both keep the reference's exact per-cell expression: upwind uses a strict
``> 0`` test and a clamped neighbour; QUICK uses ``>= 0``, ``np.pad(..., 2,
'edge')`` and the reference's literal (non-standard) negative-velocity branch.
"""
from __future__ import annotations

import numpy as np

from ..._b200.device import slab_for_shape


def _run(scheme: str, u, v, field, dx, dy):
    ny, nx = field.shape
    if ny < 2 or nx < 2:
        raise ValueError("advection needs at least a 2x2 grid")
    return slab_for_shape(ny, nx).advect(scheme, u, v, field, dx, dy)


def advect_upwind(u, v, field, dx, dy) -> np.ndarray:
    """u * d(field)/dx + v * d(field)/dy, first-order upwind."""
    return _run("upwind", u, v, field, dx, dy)


def advect_quick(u, v, field, dx, dy) -> np.ndarray:
    """Flux-difference form with QUICK face values (cell-centred velocities)."""
    return _run("quick", u, v, field, dx, dy)
