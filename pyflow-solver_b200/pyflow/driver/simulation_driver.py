"""Dict-payload driver with steady-state detection and pre/post hooks (drop-in for
the reference's alternate ``src/pyflow/driver/simulation_driver.py`` :46-191;
SURVEY.md 8f item 4).

Kept: ``PhysicsConfig`` / ``DriverConfig`` fields and defaults (:46-72), the
generator protocol of ``SimulationDriver.run`` (one dict per step with the keys of
``_build_payload`` :163-176, the live ``State`` by reference), steady detection by the
slope of log10(Ru) over the last ``steady_window`` samples (:132-146), the
``pre_step`` / ``post_step`` hooks, the periodic report line (:153-159).

Device note: the reference slices u and v on the host every step to get the mean and
the mean-free norm of the central divergence (:126-130); here those two scalars come
from two small reduction kernels on the device fields (``pf_divergence_stats``), so a
step costs no field transfer unless a hook touches ``state.fields``.
"""
from __future__ import annotations

import time
from collections.abc import Callable, Generator
from dataclasses import dataclass
from typing import Any

import numpy as np

from ..core.ghost_fields import State, SyncedFields, allocate_state, interior_view
from ..numerics.fluid_ops import divergence as div_op
from ..residuals.manager import ResidualManager
from ..solvers.solver import step as core_step


@dataclass
class PhysicsConfig:
    Re: float = 100.0
    lid_velocity: float = 1.0
    disable_advection: bool = False
    advection_scheme: str = "quick"
    cfl_target: float = 0.5
    cfl_growth: float = 1.05
    lin_tol: float = 1e-10
    lin_maxiter: int = 800
    diagnostics: bool = False
    lx: float = 1.0
    ly: float = 1.0


@dataclass
class DriverConfig:
    nx: int = 65
    ny: int = 65
    max_steps: int = 10_000
    report_interval: int = 100
    steady_window: int = 100
    steady_slope_tol: float = 2e-3
    stop_on_steady: bool = True
    save_interval: int = 0
    pre_step: Callable[["SimulationDriver", int, State], None] | None = None
    post_step: Callable[["SimulationDriver", int, State], None] | None = None


def mean_free_stats(field: np.ndarray) -> tuple[float, float]:
    mean = float(field.mean())
    return mean, float(np.linalg.norm(field - mean))


class SimulationDriver:
    def __init__(self, driver_cfg: DriverConfig, physics_cfg: PhysicsConfig | None = None):
        self.driver_cfg = driver_cfg
        self.physics_cfg = physics_cfg or PhysicsConfig()
        self.state: State = allocate_state(driver_cfg.nx, driver_cfg.ny)
        self.tracker = ResidualManager()
        self._start_wall: float | None = None
        self._steady_detected = False
        self._steady_iteration: int | None = None

    @property
    def steady_detected(self) -> bool:
        return self._steady_detected

    @property
    def steady_iteration(self) -> int | None:
        return self._steady_iteration

    @property
    def elapsed(self) -> float:
        return 0.0 if self._start_wall is None else time.time() - self._start_wall

    def _divergence_stats(self, state) -> tuple[float, float]:
        f = state.fields
        dx, dy = state.meta["dx"], state.meta["dy"]
        if isinstance(f, SyncedFields) and f._slab is not None and not f.host_touched():
            return f._slab.divergence_stats(dx, dy)          # device fields are current: two reductions, no transfer
        return mean_free_stats(div_op(interior_view(f["u"]), interior_view(f["v"]), dx, dy))

    def _steady_slope(self) -> float | None:
        cfg = self.driver_cfg
        series = self.tracker.series.get("Ru")
        if series is None or len(series.values) < cfg.steady_window:
            return None
        recent = np.array(series.values[-cfg.steady_window:])
        if not np.all(recent > 0):
            return None
        return float(np.polyfit(np.arange(recent.size), np.log10(recent), 1)[0])

    def run(self) -> Generator[dict[str, Any], None, None]:
        self._start_wall = time.time()
        cfg = self.driver_cfg
        for it in range(cfg.max_steps):
            if cfg.pre_step:
                cfg.pre_step(self, it, self.state)
            state, residuals, diag = core_step(self.physics_cfg, self.state, self.tracker, it)
            div_mean, div_mf = self._divergence_stats(state)
            slope = self._steady_slope()
            steady = slope is not None and abs(slope) < cfg.steady_slope_tol
            if steady:
                self._steady_detected = True
                self._steady_iteration = it
                if cfg.stop_on_steady:                       # final payload, then stop (hooks/report skipped, as in the reference)
                    yield self._build_payload(it, residuals, diag, div_mean, div_mf, True, slope)
                    break
            payload = self._build_payload(it, residuals, diag, div_mean, div_mf, steady, slope)
            if cfg.post_step:
                cfg.post_step(self, it, state)
            if it % cfg.report_interval == 0 or it == cfg.max_steps - 1:
                print(f"Iter {it:5d} | dt={diag['dt']:.3e} CFL={diag['CFL']:.2f} Ru={residuals['Ru']:.3e} "
                      f"Rv={residuals['Rv']:.3e} div_mf={div_mf:.3e} div_mean={div_mean:.3e}"
                      + (f" steady|slope={slope:.3e}" if steady else ""))
            yield payload

    def _build_payload(self, iteration, residuals, diag, div_mean, div_mf, steady, steady_slope) -> dict[str, Any]:
        return {
            "iteration": iteration,
            "state": self.state,
            "residuals": residuals,
            "diagnostics": diag,
            "divergence_mean": div_mean,
            "divergence_mean_free_norm": div_mf,
            "steady": steady,
            "steady_slope": steady_slope,
            "elapsed": self.elapsed,
        }


def create_default_driver() -> SimulationDriver:
    return SimulationDriver(DriverConfig(), PhysicsConfig())


__all__ = ["DriverConfig", "PhysicsConfig", "SimulationDriver", "create_default_driver"]
