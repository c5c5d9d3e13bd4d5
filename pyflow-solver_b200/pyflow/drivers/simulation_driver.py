"""Generator driver (drop-in for the reference's
``src/pyflow/drivers/simulation_driver.py:SimulationDriver`` :38-110).

Protocol kept: ``run`` yields ``(state, residuals, diagnostics)`` after every
step, always the SAME ``State`` object; ``max_steps`` is an absolute iteration
bound (:79-80); the state is allocated lazily through the injectable
``allocate`` hook (:39-44, :81-86); ``u_centerline`` and ``wall_time`` are
added to the diagnostics (:89-99); checkpoints are written AFTER the yield when
``iteration % interval == 0`` (:103-109).

The centre-line column comes back with the step's scalar read-back (ny doubles)
instead of slicing a host array, so a yield costs no field transfer.

New, opt-in: ``driver.async_checkpoint = True``, a truthy
``config.async_checkpoint`` or ``PYFLOW_B200_ASYNC_CHECKPOINT=1`` hands due
checkpoints to ``io.checkpoint.BackgroundCheckpointWriter``: the fields are
snapshotted from HBM at the reference's point in the loop (after the yield) and
hashed / compressed / renamed on a worker thread while the next steps run.  Same
bytes on disk; the file of checkpoint k is complete before checkpoint k+1 starts
and before the generator finishes or is closed.  Default: synchronous, as in the
reference.
"""
from __future__ import annotations

import os
import time
from collections.abc import Callable, Generator
from typing import Any

from ..core.ghost_fields import State, SyncedFields, allocate_state, interior_view
from ..io.checkpoint import BackgroundCheckpointWriter, save_checkpoint
from ..residuals.manager import ResidualManager
from ..solvers.solver import step


class SimulationDriver:
    def __init__(self, config: Any, state: State | None = None, tracker: ResidualManager | None = None, *,
                 allocate: Callable[[int, int], State] = allocate_state):
        self.config = config
        self.state = state
        self.tracker = tracker or ResidualManager()
        self.allocate_fn = allocate
        self._t0 = None
        self.iteration = 0
        self.async_checkpoint = False
        self._writer: BackgroundCheckpointWriter | None = None

    def wait_checkpoints(self) -> None:
        """Block until a checkpoint handed to the background writer is on disk (no-op otherwise).  Callers that
        write the same path themselves (the CLI's Ctrl-C handler) call this first, so that the worker's rename
        cannot land after -- and replace -- their newer file."""
        if self._writer is not None:
            self._writer.wait()

    def ensure_state(self, nx: int, ny: int) -> State:
        if self.state is None:
            self.state = self.allocate_fn(nx, ny)
        return self.state

    def _centerline(self, state) -> list | None:
        f = state.fields
        if isinstance(f, SyncedFields) and f._slab is not None:
            u_center = f._slab.centerline()
            if f._slab.nranks > 1:                 # concatenate the slabs' column segments (collective)
                from .._b200.distributed import gather_centerline
                u_center = gather_centerline(u_center, state)
            n = u_center.shape[0]
        else:
            ui = interior_view(f["u"])
            n, nxi = ui.shape
            u_center = ui[:, nxi // 2]
        ly = getattr(self.config, "ly", float(n - 1))
        ys = [j * (ly / (n - 1)) for j in range(n)] if n > 1 else [0.0]
        return [ys, u_center.tolist()]

    def run(self, max_steps: int | None = None, *, start_iteration: int = 0, progress: bool = False,
            checkpoint_path: str | None = None, checkpoint_interval: int | None = None
            ) -> Generator[tuple[State, dict[str, float], dict[str, Any]], None, None]:
        self.iteration = start_iteration
        self._t0 = time.time()
        writer = None
        if checkpoint_path and (self.async_checkpoint or getattr(self.config, "async_checkpoint", False)
                                or os.environ.get("PYFLOW_B200_ASYNC_CHECKPOINT", "0") == "1"):
            writer = BackgroundCheckpointWriter()
        self._writer = writer
        try:
            yield from self._loop(max_steps, progress, checkpoint_path, checkpoint_interval, writer)
        finally:
            if writer is not None:
                writer.wait()                  # the last checkpoint is on disk when the generator ends or is closed
            self._writer = None

    def _loop(self, max_steps, progress, checkpoint_path, checkpoint_interval, writer):
        while True:
            if max_steps is not None and self.iteration >= max_steps:
                return
            if self.state is None:
                nx = getattr(self.config, "nx", None)
                ny = getattr(self.config, "ny", None)
                if nx is None or ny is None:
                    raise ValueError("SimulationDriver requires either an initial State or config.nx & config.ny for allocation")
                self.state = self.allocate_fn(nx, ny)
            state, residuals, diagnostics = step(self.config, self.state, self.tracker, self.iteration)
            if getattr(state, "slab", None) is not None:
                # row slabs: the gather is a collective -- a swallowed failure on one rank would hang the others
                diagnostics["u_centerline"] = self._centerline(state)
            else:
                try:
                    diagnostics["u_centerline"] = self._centerline(state)
                except Exception:  # pragma: no cover
                    pass
            diagnostics["wall_time"] = time.time() - self._t0
            if progress:
                print(f"iter={self.iteration} dt={diagnostics.get('dt'):.3g} CFL={diagnostics.get('CFL'):.3g} "
                      f"continuity={residuals.get('continuity'):.3e}")
            yield state, residuals, diagnostics
            if checkpoint_path and checkpoint_interval and checkpoint_interval > 0:
                if self.iteration % checkpoint_interval == 0:
                    spec = getattr(state, "slab", None)
                    if spec is not None and spec["world"] > 1:
                        # row slabs: ONE file of the whole grid, written by rank 0 (every rank takes part in the
                        # gather; per-rank files under the same path would overwrite each other)
                        from .._b200.distributed import gathered_state
                        full = gathered_state(state)
                        written = True
                        if spec["rank"] == 0:
                            try:
                                save_checkpoint(checkpoint_path, full, self.iteration,
                                                diagnostics.get("time", diagnostics.get("wall_time", 0.0)), self.config)
                            except Exception:  # pragma: no cover
                                written = False
                        diagnostics["checkpoint_written"] = written
                        self.iteration += 1
                        continue
                    try:
                        sim_time = diagnostics.get("time", diagnostics.get("wall_time", 0.0))
                        if writer is not None:
                            writer.submit(checkpoint_path, state, self.iteration, sim_time, self.config)
                        else:
                            save_checkpoint(checkpoint_path, state, self.iteration, sim_time, self.config)
                        diagnostics["checkpoint_written"] = True
                    except Exception:  # pragma: no cover
                        diagnostics["checkpoint_written"] = False
            self.iteration += 1


__all__ = ["SimulationDriver"]
