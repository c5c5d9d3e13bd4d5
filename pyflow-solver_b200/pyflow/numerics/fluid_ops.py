"""Central-difference divergence used by the dt safeguard and the
``continuity`` residual (drop-in for the on-path part of the reference's
``src/pyflow/numerics/fluid_ops.py``: ``divergence`` :30, ``_divergence_impl``
:8-28).  The legacy ``gradient`` / ``laplacian`` / ``apply_pressure_correction``
helpers of that file are off the hot path and are
not provided.
"""
from __future__ import annotations

import numpy as np

from .._b200.device import slab_for_shape
from ..core.ghost_fields import State, interior_view


def divergence(u: np.ndarray, v: np.ndarray, dx: float, dy: float) -> np.ndarray:
    """Central differences inside, one-sided on the edges; the right column
    (corners excluded) stays zero exactly as in the reference (:16-27)."""
    ny, nx = u.shape
    if ny < 2 or nx < 2:
        raise ValueError("divergence needs at least a 2x2 grid")
    return slab_for_shape(ny, nx).divergence_central(u, v, dx, dy)


_divergence_impl = divergence


def legacy_divergence_state(state: State, dx: float = 1.0, dy: float = 1.0) -> np.ndarray:
    return divergence(interior_view(state.fields["u"]), interior_view(state.fields["v"]), dx, dy)
