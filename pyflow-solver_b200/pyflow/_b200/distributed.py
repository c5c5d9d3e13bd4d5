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
    from ..core.ghost_fields import _zeros
    st = State(nx=nx, ny=nyl, fields={k: _zeros((nyl + 2, nx + 2)) for k in ("u", "v", "p")})
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


def gather_centerline(segment: np.ndarray, state: State) -> np.ndarray:
    """Concatenation of every rank's column segment (u[:, nx//2] of its slab), on every rank: one tensor
    all-gather of ny_local doubles per rank (CUDA tensors under NCCL, host tensors under gloo)."""
    import torch
    import torch.distributed as dist
    spec = state.slab
    world = spec["world"]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    rows = [slab_rows(spec["ny_global"], r, world)[1] for r in range(world)]
    width = max(rows)                                    # ragged slabs: pad to the tallest one, cut after the gather
    mine = torch.zeros(width, dtype=torch.float64)
    mine[:len(segment)] = torch.from_numpy(np.ascontiguousarray(segment, dtype=np.float64))
    mine = mine.to(dev)
    parts = [torch.empty(width, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat([p[:n] for p, n in zip(parts, rows)]).cpu().numpy()


def gathered_state(state: State) -> State:
    """Full-grid host State assembled from the slabs (every rank gets it; used for checkpoints, which only
    rank 0 writes).  `meta` is shared by reference: the dt-controller memory is identical on every rank."""
    full = State(nx=state.nx, ny=state.slab["ny_global"], fields={k: gather_field(state, k) for k in ("u", "v", "p")})
    full.meta = state.meta
    return full


def _symmetric_allocator(nbytes: int, device: int):
    """Workspace in torch symmetric memory: every rank allocates the same size
    and the buffers are peer-mapped across the GPUs of the box (NVLink)."""
    import torch
    import torch.distributed._symmetric_memory as symm
    return symm.empty(nbytes, dtype=torch.uint8, device=torch.device(f"cuda:{device}"))


def make_slab(state: State) -> DeviceSlab:
    """DeviceSlab for a slab state with the NCCL communicator initialised and,
    when the slabs are equal, the peer-memory CG path enabled
    (PYFLOW_B200_CG_COMM=nccl forces NCCL calls per phase instead)."""
    import os
    import warnings

    import torch
    import torch.distributed as dist
    spec = state.slab
    world, rank = spec["world"], spec["rank"]
    want_peer = (world > 1 and spec["ny_global"] % world == 0
                 and os.environ.get("PYFLOW_B200_CG_COMM", "peer").lower() != "nccl")
    slab = None
    if want_peer:
        try:
            slab = DeviceSlab(state.nx, state.ny, spec["ny_global"], spec["row_offset"],
                              device=torch.cuda.current_device(), allocator=_symmetric_allocator)
        except Exception as exc:   # symmetric memory unavailable: stay on the GPU, use NCCL per phase
            warnings.warn(f"symmetric memory unavailable ({exc!r}); CG falls back to NCCL halo/all-reduce per phase")
            want_peer = False
    if slab is None:
        slab = DeviceSlab(state.nx, state.ny, spec["ny_global"], spec["row_offset"], device=torch.cuda.current_device())
    if world > 1:
        box = [DeviceSlab.make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        slab.init_comm(box[0], rank, world)
        # every rank must agree on the mode
        flags = [None] * world
        dist.all_gather_object(flags, bool(want_peer))
        if all(flags):
            import torch.distributed._symmetric_memory as symm
            hdl = symm.rendezvous(slab.workspace, dist.group.WORLD)
            torch.cuda.synchronize()
            dist.barrier()                     # all workspaces are zero-initialised before anyone writes remotely
            slab.set_peers(list(hdl.buffer_ptrs))
            slab._symm_handle = hdl
        elif any(flags):
            # symmetric memory came up on some ranks only: EVERY rank stays on the NCCL split-phase path (leaving
            # set_peers unset everywhere keeps the collectives matched; raising on a subset would hang the others)
            warnings.warn("symmetric memory is not available on every rank; all ranks use NCCL halo / all-reduce "
                          "calls per CG phase")
    return slab
