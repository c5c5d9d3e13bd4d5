"""Binding between a host ``State`` and its GPU context.

Two kinds of state reach ``step`` / ``solve_pressure_poisson``:

* our own ``State`` (``fields`` is a ``SyncedFields``): device-resident between
  steps; only host-touched fields are uploaded, nothing is downloaded until
  somebody reads ``state.fields[...]``;
* a foreign object with plain ``fields`` / ``meta`` dicts (e.g. the reference's
  own dataclass), or our ``State`` with ``fields.eager`` set: *eager* mode --
  upload before, download after, every call, into the caller's arrays in place
  (exact reference semantics, slower).
"""
from __future__ import annotations

import numpy as np

from ..core.ghost_fields import SyncedFields
from .device import DeviceSlab


def _dims(arr: np.ndarray) -> tuple[int, int]:
    ny, nx = arr.shape
    return nx - 2, ny - 2


def _as_field(arr) -> np.ndarray:
    if not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or not arr.flags.c_contiguous:
        raise TypeError("state fields must be C-contiguous float64 numpy arrays (core/ghost_fields.py layout)")
    return arr


class Binding:
    """Result of ``attach``: the slab plus how to finish the call."""

    def __init__(self, state, slab: DeviceSlab, lazy: bool):
        self.state, self.slab, self.lazy = state, slab, lazy

    def prestep(self):
        f = self.state.fields
        if self.lazy and f._prestep is not None and not f.host_touched():
            return f._prestep
        return self.slab.prestep()

    def finish(self, prestep_next=None):
        """Device now holds the newest u, v, p."""
        f = self.state.fields
        if self.lazy:
            f.mark_device_updated()
            f._prestep = prestep_next
        else:
            if isinstance(f, SyncedFields):            # our State in eager mode: same arrays, nothing left stale
                self.slab.download(u=f.peek("u"), v=f.peek("v"), p=f.peek("p"))
                f._stale.clear()
                f._prestep = None
            else:
                self.slab.download(u=f["u"], v=f["v"], p=f["p"])


def attach(state) -> Binding:
    """Make the device copy of ``state`` current and return its binding."""
    f = state.fields
    ours = isinstance(f, SyncedFields)
    lazy = ours and not f.eager
    u = f.peek("u") if ours else f["u"]
    nx, ny = _dims(u)
    slab = f._slab if ours else getattr(state, "_pf_slab", None)
    if slab is None or (slab.nx, slab.ny_local) != (nx, ny):
        if getattr(state, "slab", None) is not None:      # row slab of a multi-GPU grid
            from .distributed import make_slab
            slab = make_slab(state)
        else:
            slab = DeviceSlab(nx, ny)
        if ours:
            f._slab = slab
            f._touched.update(("u", "v", "p"))
            f._stale.clear()
            f._prestep = None
        else:
            try:
                state._pf_slab = slab
            except Exception:
                pass
    if lazy:
        dirty = f.host_touched()
        if dirty:
            slab.upload(**{n: _as_field(f.peek(n)) for n in dirty})
            f._touched.clear()
            f._prestep = None
    else:
        # eager: the host arrays are authoritative before every call.  (For our own State, __getitem__ first pulls
        # back anything a previous lazy call left stale.)
        slab.upload(u=_as_field(f["u"]), v=_as_field(f["v"]), p=_as_field(f["p"]))
    return Binding(state, slab, lazy)
