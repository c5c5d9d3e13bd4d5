"""State container of the drop-in (mirrors ``src/pyflow/core/ghost_fields.py``
of the reference: ``State`` :8-45, ``allocate_state`` :47-57, ``interior_view``
:59-60).

Layout contract kept: every field is a float64 ``(ny+2, nx+2)`` array with a
ghost ring the solver never touches; ``state.fields[name]`` is a numpy array the
caller may read and write in place; ``state.meta`` carries ``dt_prev``, ``dx``,
``dy``, ``nu``, ``divergence_history``, ``prev_div_norm``.

What is new: the authoritative copy of ``u, v, p`` lives in HBM between steps.
``state.fields`` is a dict whose read accessors *materialise* a field from the
device on demand (one D2H of that field) and mark it host-touched, so the next
step re-uploads it.  A consumer that only looks at the yielded scalars never
pays a field transfer (at 4096^2 the three fields are 400 MB per step).
Limitation, by design: an array obtained from ``state.fields`` reflects the
device contents at the time of that access; re-read ``state.fields[name]`` after
the next step (every reference call site and every reference test does).  A
caller that keeps array views across steps WITHOUT going back through
``state.fields`` -- reading them after a step, or writing into them -- can ask
for the reference's exact live-array semantics with ``state.fields.eager = True``
(or ``PYFLOW_B200_EAGER=1`` in the environment): all three fields are then
uploaded before and downloaded after every step, into the same arrays, at the
cost of those transfers.
"""
from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Any

import numpy as np

_SOLVER_FIELDS = ("u", "v", "p")


class SyncedFields(dict):
    """``dict[str, ndarray]`` whose reads pull stale fields back from the GPU."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._slab = None              # DeviceSlab once the state has been stepped
        self._stale: set[str] = set()  # host copy older than the device copy
        self._touched: set[str] = set(_SOLVER_FIELDS)  # host copy possibly newer than the device copy
        self._prestep = None           # cached (max|u|, max|v|, sum div_central^2) of the device fields
        import os
        self.eager = os.environ.get("PYFLOW_B200_EAGER", "0") == "1"   # live-array semantics, see module docstring

    # -- device binding (used by solvers.solver) -----------------------------
    def peek(self, name: str) -> np.ndarray:
        """Host array without synchronisation (shape / identity queries)."""
        return dict.__getitem__(self, name)

    def _materialise(self, name: str):
        if name in self._stale and self._slab is not None:
            arr = dict.__getitem__(self, name)
            self._slab.download(**{name: arr})
            self._stale.discard(name)
        if name in _SOLVER_FIELDS:
            self._touched.add(name)

    def read(self, name: str) -> np.ndarray:
        """Current contents WITHOUT marking the field host-touched: the caller
        promises not to write (checkpoint writer, read-only consumers)."""
        if name in self._stale and self._slab is not None:
            self._slab.download(**{name: dict.__getitem__(self, name)})
            self._stale.discard(name)
        return dict.__getitem__(self, name)

    def mark_device_updated(self):
        self._stale = {n for n in _SOLVER_FIELDS if dict.__contains__(self, n)}
        self._touched.clear()

    def host_touched(self) -> set[str]:
        return set(self._touched)

    # -- dict protocol ---------------------------------------------------------
    def __getitem__(self, name):
        self._materialise(name)
        return dict.__getitem__(self, name)

    def get(self, name, default=None):
        if dict.__contains__(self, name):
            return self[name]
        return default

    def __setitem__(self, name, value):
        dict.__setitem__(self, name, value)
        self._stale.discard(name)
        if name in _SOLVER_FIELDS:
            self._touched.add(name)

    def items(self):
        return [(k, self[k]) for k in dict.keys(self)]

    def values(self):
        return [self[k] for k in dict.keys(self)]

    def copy(self):
        return {k: self[k] for k in dict.keys(self)}

    # every mutating entry point of dict goes through __setitem__ / the stale + touched bookkeeping:
    # the inherited C implementations would bypass it (a replaced field would be overwritten by the
    # next download and ignored by the next step)
    def update(self, *a, **k):
        for name, value in dict(*a, **k).items():
            self[name] = value

    def setdefault(self, name, default=None):
        if not dict.__contains__(self, name):
            self[name] = default
        return self[name]

    def pop(self, name, *default):
        if dict.__contains__(self, name):
            self._materialise(name)
            self._stale.discard(name)
            self._touched.discard(name)
        return dict.pop(self, name, *default)

    def popitem(self):
        name = next(reversed(dict.keys(self)))
        return name, self.pop(name)

    def __delitem__(self, name):
        dict.__delitem__(self, name)
        self._stale.discard(name)
        self._touched.discard(name)

    def clear(self):
        dict.clear(self)
        self._stale.clear()
        self._touched.clear()

    def __ior__(self, other):
        self.update(other)
        return self

    def __or__(self, other):
        out = self.copy()
        out.update(other)
        return out

    # copies are plain host snapshots: they never share the GPU context (one pf_destroy per context)
    def _snapshot(self):
        out = SyncedFields({k: np.array(self[k]) if isinstance(self[k], np.ndarray) else self[k]
                            for k in dict.keys(self)})
        out.eager = self.eager
        return out

    def __copy__(self):
        return self._snapshot()

    def __deepcopy__(self, memo):
        out = self._snapshot()
        memo[id(self)] = out
        return out

    def __reduce__(self):
        return (SyncedFields, ({k: np.array(self[k]) if isinstance(self[k], np.ndarray) else self[k]
                                for k in dict.keys(self)},))


@dataclass
class State:
    """Simulation fields + lightweight metadata (reference ghost_fields.py:8-45)."""
    nx: int
    ny: int
    fields: dict[str, np.ndarray]
    meta: dict[str, Any] = field(default_factory=dict)

    def __post_init__(self):
        if not isinstance(self.fields, SyncedFields):
            self.fields = SyncedFields(self.fields)

    def arr(self, name: str) -> np.ndarray:
        return self.fields[name]

    def __getitem__(self, key: str) -> np.ndarray:
        return self.fields[key]

    def __setitem__(self, key: str, value: np.ndarray) -> None:
        self.fields[key] = value

    def get(self, key: str, default=None):
        if key in self.fields:
            return self.fields[key]
        return self.meta.get(key, default)

    def set_meta(self, key: str, value: Any) -> None:
        self.meta[key] = value

    def keys(self):
        return self.fields.keys()

    def items(self):
        return self.fields.items()

    def values(self):
        return self.fields.values()


def _zeros(shape) -> np.ndarray:
    """Zeroed float64 host array; page-locked when a GPU is present so that the field
    uploads / downloads run at full PCIe speed (PYFLOW_B200_PINNED=0 disables)."""
    import os
    if os.environ.get("PYFLOW_B200_PINNED", "1") != "0" and shape[0] * shape[1] >= 1 << 16:
        try:
            import torch
            if torch.cuda.is_available():
                return torch.zeros(shape, dtype=torch.float64, pin_memory=True).numpy()
        except Exception:
            pass
    return np.zeros(shape, dtype=float)


def allocate_state(nx: int, ny: int, fields: tuple[str, ...] = ("u", "v", "p")) -> State:
    """Zero fields including the ghost ring (reference ghost_fields.py:47-57)."""
    shape = (ny + 2, nx + 2)
    return State(nx=nx, ny=ny, fields={k: _zeros(shape) for k in fields})


def interior_view(arr: np.ndarray) -> np.ndarray:
    return arr[1:-1, 1:-1]


def ghost_shapes(state: Mapping[str, np.ndarray] | State) -> dict[str, tuple[int, int]]:
    items = state.fields.items() if isinstance(state, State) else state.items()
    return {k: v.shape for k, v in items if isinstance(v, np.ndarray)}
