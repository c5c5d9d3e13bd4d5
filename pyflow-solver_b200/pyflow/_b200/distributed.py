"""Row-slab decomposition over the GPUs of one box (new design; the reference
has no distributed path -- SURVEY.md section 8e).

One process per GPU (torchrun).  Rank g of G owns the interior rows
``[g*ny/G, (g+1)*ny/G)`` of the global ``ny x nx`` grid, all columns; its host
``State`` holds only that slab (``(ny_local+2, nx+2)`` ghosted arrays, same
layout contract as the single-GPU state).  ``torch.distributed`` is the
plumbing (rendezvous, broadcasting the 128-byte NCCL id); halo rows and the
scalar all-reduces travel inside libpyflow_b200 over its own NCCL communicator
on the compute stream.

Only rank 0 of the bottom slab owns the pinned cell, only the top slab owns the
lid row; every rank gets identical step scalars (dt logic runs redundantly and
bit-identically on every rank).
"""
from __future__ import annotations

import numpy as np

from ..core.ghost_fields import State
from .device import DeviceSlab


def slab_rows(ny_global: int, rank: int, world: int) -> tuple[int, int]:
    """(row_offset, ny_local) of `rank`; the same integer split the device uses."""
    lo = rank * ny_global // world
    hi = (rank + 1) * ny_global // world
    return lo, hi - lo


def allocate_slab_state(nx: int, ny_global: int, rank: int, world: int) -> State:
    """Zero state for this rank's slab; the slab geometry rides on the State."""
    row0, nyl = slab_rows(ny_global, rank, world)
    if nyl < 2:
        raise ValueError(f"slab of rank {rank} has {nyl} rows; need at least 2 per GPU")
    st = State(nx=nx, ny=nyl, fields={k: np.zeros((nyl + 2, nx + 2)) for k in ("u", "v", "p")})
    st.slab = {"ny_global": int(ny_global), "row_offset": int(row0), "rank": int(rank), "world": int(world)}
    return st


def scatter_state(full: State | None, nx: int, ny_global: int, rank: int, world: int) -> State:
    """Cut this rank's slab out of a full-grid state (every rank passes the same
    full arrays, e.g. loaded from a checkpoint)."""
    st = allocate_slab_state(nx, ny_global, rank, world)
    if full is not None:
        row0, nyl = slab_rows(ny_global, rank, world)
        for k in ("u", "v", "p"):
            src = full.fields[k]
            st.fields.peek(k)[1:-1, :] = src[1 + row0:1 + row0 + nyl, :]
    return st


def gather_field(state: State, name: str):
    """Full (ny_global+2, nx+2) array on every rank (ghost ring zero)."""
    import torch.distributed as dist
    spec = state.slab
    local = np.ascontiguousarray(state.fields[name][1:-1, :])
    parts = [None] * spec["world"]
    dist.all_gather_object(parts, local)
    out = np.zeros((spec["ny_global"] + 2, state.nx + 2))
    out[1:-1, :] = np.concatenate(parts, axis=0)
    return out


def make_slab(state: State) -> DeviceSlab:
    """DeviceSlab for a slab state, with the NCCL communicator initialised."""
    import torch
    import torch.distributed as dist
    spec = state.slab
    slab = DeviceSlab(state.nx, state.ny, spec["ny_global"], spec["row_offset"], device=torch.cuda.current_device())
    if spec["world"] > 1:
        box = [DeviceSlab.make_unique_id() if spec["rank"] == 0 else None]
        dist.broadcast_object_list(box, src=0)
        slab.init_comm(box[0], spec["rank"], spec["world"])
    return slab
