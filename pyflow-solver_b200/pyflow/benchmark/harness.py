"""Grid-size sweep of the cavity time step on the device path (the reference's
``src/pyflow/benchmark/harness.py`` :35-103 re-pointed at the GPU).

    python -m pyflow.benchmark.harness --grids 512,1024,2048,4096 --steps 20 --json-out sweep.json

Same entry points and result layout as the reference (``make_config``, ``run_case``,
``benchmark``; one record per grid with ``nx, ny, steps, total_time_s,
avg_time_per_step_s, avg_cg_iterations`` under ``runs``), with two deliberate changes:

* the reference's harness configuration (lid at rest, advection off, :35-53) keeps every
  field at zero, so its pressure solve returns at once and it times nothing
  (``avg_cg_iterations = 0``).  Here the defaults drive the cavity -- lid velocity 1,
  QUICK advection, 300 CG iterations allowed, as ``pyflow run`` does -- and both can still
  be overridden through ``make_config``;
* each record also carries what the device reports: true CG iterations (not the 0 / 1 /
  maxiter flag), device time per step and per solve from CUDA events, MLUPS and CG
  iterations per second.  The first step of a run from rest has no pressure work and is
  reported but excluded from the steady-state rates.

``bench.py`` at the repository root is the contract benchmark (roofline, end-to-end, CPU
baseline); this module is the reference-shaped convenience sweep.
"""
from __future__ import annotations

import argparse
import json
import time
from types import SimpleNamespace
from typing import Any

from ..core.ghost_fields import allocate_state
from ..drivers.simulation_driver import SimulationDriver
from ..residuals.manager import ResidualManager


def make_config(nx: int, ny: int, *, Re: float = 100.0, lid_velocity: float = 1.0, disable_advection: bool = False,
                enable_jacobi_pc: bool = True, advection_scheme: str = "quick", lin_maxiter: int = 300):
    """Duck-typed configuration with every attribute ``step`` / ``solve_pressure_poisson`` read."""
    return SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=lid_velocity, cfl_target=0.5, cfl_growth=1.05,
                           advection_scheme=advection_scheme, disable_advection=disable_advection, lin_tol=1e-10,
                           lin_maxiter=lin_maxiter, diagnostics=False, enable_jacobi_pc=enable_jacobi_pc,
                           force_quiet=True, lx=nx - 1, ly=ny - 1)


def run_case(nx: int, ny: int, steps: int, config: Any | None = None) -> dict[str, Any]:
    cfg = config if config is not None else make_config(nx, ny)
    state = allocate_state(nx, ny)
    driver = SimulationDriver(cfg, state, ResidualManager())
    flagged, true_iters, step_ms, cg_ms = [], [], [], []
    started = time.time()
    for st, _residuals, diag in driver.run(max_steps=steps):
        flagged.append(int(diag.get("Rp_iterations", 0)))
        true_iters.append(int(st.meta.get("cg_true_iterations", 0)))
        step_ms.append(float(st.meta.get("step_ms", 0.0)))
        cg_ms.append(float(st.meta.get("cg_ms", 0.0)))
    elapsed = time.time() - started
    done = len(flagged)
    record: dict[str, Any] = {
        "nx": nx, "ny": ny, "steps": done,
        "total_time_s": elapsed,
        "avg_time_per_step_s": elapsed / max(done, 1),
        "avg_cg_iterations": sum(flagged) / max(done, 1),
    }
    steady = slice(1, None) if done > 1 else slice(0, None)        # step 0 from rest: b = 0, no CG work
    n_steady = len(true_iters[steady])
    dev_s = sum(step_ms[steady]) * 1e-3
    solve_s = sum(cg_ms[steady]) * 1e-3
    record.update({
        "avg_true_cg_iterations": sum(true_iters[steady]) / max(n_steady, 1),
        "avg_device_step_ms": sum(step_ms[steady]) / max(n_steady, 1),
        "avg_cg_solve_ms": sum(cg_ms[steady]) / max(n_steady, 1),
        "mlups": (nx * ny * n_steady / dev_s / 1e6) if dev_s > 0 else None,
        "cg_iterations_per_s": (sum(true_iters[steady]) / solve_s) if solve_s > 0 else None,
    })
    return record


def benchmark(grids: list[int], steps: int) -> dict[str, Any]:
    return {"benchmark": "lid_driven_cavity", "steps_requested": steps,
            "runs": [run_case(n, n, steps) for n in grids]}


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description="PyFlow benchmark sweep (B200 device path)")
    ap.add_argument("--grids", type=str, default="32,64,128,256", help="comma-separated square grid sizes")
    ap.add_argument("--steps", type=int, default=100, help="time steps per grid")
    ap.add_argument("--json-out", type=str, default=None, help="also write the JSON document to this file")
    args = ap.parse_args(argv)
    sizes = [int(tok) for tok in args.grids.split(",") if tok.strip()]
    text = json.dumps(benchmark(sizes, args.steps), indent=2)
    if args.json_out:
        with open(args.json_out, "w", encoding="utf-8") as fh:
            fh.write(text + "\n")
    print(text)
    return 0


if __name__ == "__main__":  # pragma: no cover
    raise SystemExit(main())

__all__ = ["benchmark", "make_config", "run_case"]
