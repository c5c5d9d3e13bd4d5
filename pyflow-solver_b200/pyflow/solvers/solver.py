"""One projection time step, device-backed (drop-in for the reference's
``src/pyflow/solvers/solver.py:step`` :15-193).

Split of work:
  * HOST (this file, plain Python floats so the values are bit-identical to the
    reference): grid spacing :17-22, the CFL / diffusion / growth-limited dt
    :25-41, the divergence-history safeguard :43-57, CFL number :62, the
    residual / diagnostics dictionaries :122-154, optional JSONL logging and
    the emergency NaN checkpoint :170-192.
  * DEVICE (``libpyflow_b200``, one ``pf_step`` call): the three pre-step
    reductions, advective predictor, RHS, Jacobi-PCG pressure solve, velocity
    correction, lid BC, residual sums, finite check, centre-line gather.

Quirks of the reference that are reproduced on purpose: no
diffusion term (``nu`` is metadata only), lid row written AFTER the projection,
diagnostics divergence = central operator with unit spacing, ``Rp`` is 0.0,
``Rp_iterations`` is 0 / 1 / lin_maxiter rather than a count, ``Ru`` divides by
N not sqrt(N).
"""
from __future__ import annotations

import math
from typing import Any

from .._b200 import lib as _lib
from .._b200.engine import attach
from ..logging.structured import JsonlLogger
from .pressure_solver import _projection_params, _rp_diagnostics

_SCHEMES = {"upwind": _lib.PF_ADV_UPWIND, "quick": _lib.PF_ADV_QUICK}


def step(config, state, tracker, iteration: int) -> tuple[Any, dict[str, float], dict[str, Any]]:
    verbose = getattr(config, "diagnostics", True)
    bind = attach(state)
    slab = bind.slab
    nx, ny = slab.nx, slab.ny_global
    lx = getattr(config, "lx", float(nx - 1))
    ly = getattr(config, "ly", float(ny - 1))
    dx = lx / max(nx - 1, 1)
    dy = ly / max(ny - 1, 1)
    Re = float(getattr(config, "Re", 100.0))
    if verbose:
        print(f"[Step {iteration}] Grid: nx={nx}, ny={ny}, dx={dx:.4g}, dy={dy:.4g}, Re={Re}")

    # ---- pre-step reductions (device, or cached from the previous step) ----
    umax_raw, vmax_raw, divc_sumsq = bind.prestep()
    umax = float(umax_raw) + 1e-12
    vmax = float(vmax_raw) + 1e-12

    # ---- dt control: reference solver.py:27-41, same float operations -------
    cfl_target = getattr(config, "cfl_target", 0.5)
    cfl_growth = getattr(config, "cfl_growth", 1.05)
    nu = 1.0 / Re
    dt_prev = float(state.meta.get("dt_prev", 1e-2))
    dt_cfl = min(cfl_target * dx / umax, cfl_target * dy / vmax)
    dt_diff = 0.25 * min(dx * dx, dy * dy) / max(nu, 1e-12)
    dt_candidate = min(dt_cfl, dt_diff)
    dt = dt_prev * cfl_growth if dt_candidate > dt_prev * cfl_growth else dt_candidate

    # ---- divergence-trend safeguard: solver.py:43-57 --------------------------
    div_baseline = math.sqrt(divc_sumsq) if divc_sumsq == divc_sumsq else float("nan")
    hist = state.meta.get("divergence_history", [])
    hist.append(div_baseline)
    hist = hist[-5:]
    state.meta["divergence_history"] = hist
    if len(hist) >= 3:
        a, b, c = hist[-3], hist[-2], hist[-1]
        if a > 0 and b > a * 1.01 and c > b * 1.01:
            dt *= 0.5
            if verbose:
                print(f"[Step {iteration}] dt safeguard triggered (divergence rising {a:.3g}->{b:.3g}->{c:.3g}); new dt={dt:.3g}")
    state.meta["dt_prev"] = dt
    state.meta["dx"] = dx
    state.meta["dy"] = dy
    state.meta["nu"] = nu
    current_cfl = max(umax * dt / dx, vmax * dt / dy)
    if verbose:
        print(f"[Step {iteration}] dt={dt:.3g}, umax={umax:.3g}, vmax={vmax:.3g}, CFL={current_cfl:.3g}")
        print(f"[Step {iteration}] Initial divergence norm: {div_baseline:.3g}")

    # ---- the device step -------------------------------------------------------
    if getattr(config, "disable_advection", False):
        scheme_name, scheme = "(disabled)", _lib.PF_ADV_NONE
    else:
        scheme_name = getattr(config, "advection_scheme", "upwind")
        scheme = _SCHEMES.get(scheme_name, _lib.PF_ADV_UPWIND)   # reference: anything but 'quick' is upwind
    lid_velocity = getattr(config, "lid_velocity", None)
    params = _projection_params(config, dt, dx, dy, preconditioner=None)
    params.scheme = scheme
    params.projection_only = 0
    params.lid_set = 0 if lid_velocity is None else 1
    params.lid_velocity = 0.0 if lid_velocity is None else float(lid_velocity)
    out = slab.step(params)
    bind.finish(prestep_next=(out.umax_raw, out.vmax_raw, out.cont_sumsq))
    pressure_diag = _rp_diagnostics(config, out)
    if verbose:
        print(f"[Step {iteration}] Advection scheme: {scheme_name}")
        print(f"[Step {iteration}] Pressure solve: iterations={pressure_diag['Rp_iterations']}, "
              f"residual={pressure_diag['Rp_residual']:.3g}, converged={pressure_diag['Rp_converged']}")
    state.meta["prev_div_norm"] = pressure_diag["divergence_norm"]
    # extras that are not part of the reference's dictionaries live in meta
    state.meta["cg_true_iterations"] = int(out.true_iters)
    state.meta["cg_ms"] = float(out.cg_ms)
    state.meta["step_ms"] = float(out.step_ms)

    # ---- residuals: solver.py:106-121 ------------------------------------------
    n_cells = max(1, nx * ny)
    Ru = float(math.sqrt(out.ru_sumsq) / n_cells) if out.ru_sumsq == out.ru_sumsq else float("nan")
    Rv = float(math.sqrt(out.rv_sumsq) / n_cells) if out.rv_sumsq == out.rv_sumsq else float("nan")
    Rp = 0.0
    cont_norm = math.sqrt(out.cont_sumsq) if out.cont_sumsq == out.cont_sumsq else float("nan")
    continuity = float(cont_norm / max(1.0, nx * ny))
    nan_detected = bool(out.nonfinite)
    if verbose:
        print(f"[Step {iteration}] Residuals: Ru={Ru:.3g}, Rv={Rv:.3g}, continuity={continuity:.3g}")
    tracker.add("Ru", Ru)
    tracker.add("Rv", Rv)
    tracker.add("Rp", Rp)
    tracker.add("continuity", continuity)
    tracker.add("Rp_iterations", pressure_diag["Rp_iterations"])
    tracker.add("Rp_residual", pressure_diag["Rp_residual"])
    residuals = {"Ru": Ru, "Rv": Rv, "Rp": Rp, "continuity": continuity}

    ru_series = tracker.series["Ru"].values if "Ru" in tracker.series else None
    if ru_series and Ru > 0 and ru_series[0] > 0:
        drop = math.log10(ru_series[0]) - math.log10(Ru)
    else:
        drop = 0.0
    diagnostics = {
        "iteration": iteration,
        "dt": dt,
        "CFL": current_cfl,
        "plateau_Ru": tracker.plateau_detect("Ru", window=10, threshold=-0.005),
        "max_residual": max(residuals.values()),
        "residual_drop_Ru_orders": max(0.0, drop),
        "nu": nu,
        "Rp_iterations": pressure_diag["Rp_iterations"],
        "Rp_residual": pressure_diag["Rp_residual"],
        "Rp_converged": pressure_diag["Rp_converged"],
        "Rp_method": pressure_diag["Rp_method"],
        "divergence_norm": pressure_diag["divergence_norm"],
        "nan_detected": nan_detected,
    }
    if getattr(config, "seed", None) is not None:
        diagnostics["seed"] = config.seed

    # ---- optional invariants: solver.py:158-167 ---------------------------------
    if getattr(config, "assert_invariants", False):
        prev_div = pressure_diag.get("divergence_norm")
        if prev_div is None:
            raise AssertionError("Invariant failed: divergence_norm missing from pressure diagnostics")
        assert continuity <= max(1e3, prev_div * 10)
        assert abs(out.p00) < 1e-10, "Invariant failed: reference pressure cell not ~0"

    # ---- structured logging + emergency checkpoint: solver.py:170-192 -----------
    log_path = getattr(config, "log_path", None)
    log_stream = getattr(config, "log_stream", None)
    if log_path or log_stream:
        logger = getattr(config, "_logger_instance", None)
        if logger is None:
            logger = JsonlLogger(path=log_path, stream=log_stream)
            config._logger_instance = logger
        logger.log({"type": "step", "it": iteration, "dt": dt, "CFL": current_cfl, "Ru": Ru, "Rv": Rv,
                    "continuity": continuity, "Rp_it": pressure_diag["Rp_iterations"], "nan": nan_detected})
        if nan_detected:
            logger.log({"type": "error", "it": iteration, "reason": "nan_detected"})
            try:
                from ..io.checkpoint import save_checkpoint
                ck_path = getattr(config, "emergency_checkpoint_path", None)
                if ck_path:
                    base, ext = (ck_path.rsplit(".", 1) + ["npz"])[:2]
                    fail_path = f"{base}_FAIL_NAN.{ext}"
                else:
                    fail_path = f"emergency_FAIL_NAN_it{iteration}.npz"
                save_checkpoint(fail_path, state, iteration,
                                diagnostics.get("time", diagnostics.get("wall_time", 0.0)), config)
                logger.log({"type": "emergency_checkpoint", "path": fail_path, "it": iteration})
            except Exception:
                logger.log({"type": "emergency_checkpoint", "status": "failed", "it": iteration})
    return state, residuals, diagnostics
