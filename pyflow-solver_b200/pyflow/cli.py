"""``pyflow`` command line (same sub-commands, flags, exit codes and JSON-line
schema as the reference's ``src/pyflow/cli.py`` :35-194), driving the
device-backed ``SimulationDriver``.

    python -m pyflow run --nx 4096 --ny 4096 --re 1000 --steps 100 --json-stream

One JSON object per step on stdout (:154-173):
  {type:"step", iteration, dt, CFL, continuity, wall_time,
   residuals:{Ru,Rv,Rp,continuity}, diagnostics:{...all other step keys...},
   [run_meta:{seed}]}
Exit codes: 0 ok, 2 invalid configuration, 3 restart hash mismatch,
130 interrupted (a final checkpoint is attempted).
"""
from __future__ import annotations

import argparse
import json
import sys

from pydantic import ValidationError

from .config.model import ConfigError, SimulationConfig
from .core.ghost_fields import allocate_state
from .drivers.simulation_driver import SimulationDriver
from .residuals.manager import ResidualManager

_STEP_LEVEL_KEYS = ("iteration", "dt", "CFL", "wall_time")


def _add_common(sp: argparse.ArgumentParser) -> None:
    sp.add_argument("--nx", type=int, default=64)
    sp.add_argument("--ny", type=int, default=64)
    sp.add_argument("--re", type=float, default=100.0)
    sp.add_argument("--lid-velocity", type=float, default=1.0)
    sp.add_argument("--scheme", choices=["quick", "upwind"], default="quick")
    sp.add_argument("--disable-advection", action="store_true")
    sp.add_argument("--cfl", type=float, default=0.5)
    sp.add_argument("--cfl-growth", type=float, default=1.05)
    sp.add_argument("--lin-tol", type=float, default=1e-10)
    sp.add_argument("--lin-maxiter", type=int, default=300)
    sp.add_argument("--checkpoint-interval", type=int, default=0)   # 0 = no periodic checkpoints
    sp.add_argument("--log-jsonl", type=str, default=None)
    sp.add_argument("--seed", type=int, default=None)
    sp.add_argument("--diagnostics", action="store_true")


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="PyFlow CLI (B200 device path)")
    sub = parser.add_subparsers(dest="command", required=True)
    run_p = sub.add_parser("run", help="Run a simulation")
    _add_common(run_p)
    run_p.add_argument("--steps", type=int, default=100)
    run_p.add_argument("--checkpoint", type=str, default=None)
    run_p.add_argument("--restart", type=str, default=None)
    run_p.add_argument("--allow-hash-mismatch", action="store_true")
    # new: restore the dt controller's memory saved in the checkpoint, so the resumed run
    # continues the interrupted one instead of restarting the dt ramp (the reference's, and the default, behaviour)
    run_p.add_argument("--exact-restart", action="store_true")
    run_p.add_argument("--json-stream", action="store_true")
    run_p.add_argument("--continuity-threshold", type=float, default=None)
    run_p.add_argument("--progress", action="store_true")
    run_p.add_argument("--assert-invariants", action="store_true")
    _add_common(sub.add_parser("validate", help="Validate and display normalized config"))
    _add_common(sub.add_parser("show-config", help="Show normalized config (hash only)"))
    return parser


def _config_from_args(args) -> SimulationConfig:
    interval = args.checkpoint_interval if args.checkpoint_interval not in (None, 0) else None
    return SimulationConfig(
        nx=args.nx, ny=args.ny, Re=args.re, lid_velocity=args.lid_velocity,
        cfl_target=args.cfl, cfl_growth=args.cfl_growth,
        advection_scheme=args.scheme, disable_advection=args.disable_advection,
        lin_tol=args.lin_tol, lin_maxiter=args.lin_maxiter,
        diagnostics=args.diagnostics, checkpoint_interval=interval,
        log_path=args.log_jsonl, seed=args.seed, turbulence_model=None,
    )


def _step_payload(cfg: SimulationConfig, residuals: dict, diag: dict) -> dict:
    payload = {
        "type": "step",
        "iteration": diag.get("iteration"),
        "dt": diag.get("dt"),
        "CFL": diag.get("CFL"),
        "continuity": residuals.get("continuity"),
        "wall_time": diag.get("wall_time"),
        "residuals": {k: residuals.get(k) for k in ("Ru", "Rv", "Rp", "continuity")},
        "diagnostics": {k: v for k, v in diag.items() if k not in _STEP_LEVEL_KEYS},
    }
    if cfg.seed is not None:
        payload["run_meta"] = {"seed": cfg.seed}
    return payload


def main(argv: list[str] | None = None) -> int:
    args = build_parser().parse_args(argv)
    try:
        cfg = _config_from_args(args)
    except (ValidationError, ConfigError) as exc:
        print("Configuration invalid:\n" + str(exc))
        return 2

    if args.command == "validate":
        out = cfg.model_dump()
        out["config_hash"] = cfg.config_hash
        print(json.dumps(out, indent=2, sort_keys=True))
        return 0
    if args.command == "show-config":
        print(json.dumps({"hash": cfg.config_hash, "brief": cfg.brief()}, indent=2))
        return 0

    start_it = 0
    state = None
    if args.restart:
        try:
            from .io.checkpoint import load_checkpoint
            state, meta = load_checkpoint(args.restart)
            saved = meta.get("structured_config_hash")
            if saved and saved != cfg.config_hash and not args.allow_hash_mismatch:
                print(f"Refusing restart: checkpoint hash {saved} != current {cfg.config_hash} "
                      "(use --allow-hash-mismatch to override)")
                return 3
            start_it = int(meta.get("iteration", 0)) + 1
            from .io.checkpoint import exact_restart_requested, restore_step_state
            if exact_restart_requested(args.exact_restart):
                restore_step_state(state, meta)
        except Exception as exc:
            print(f"Failed to load checkpoint {args.restart}: {exc}")
            state, start_it = None, 0
    if state is None:
        state = allocate_state(cfg.nx, cfg.ny)
    driver = SimulationDriver(cfg, state, ResidualManager())

    json_mode = bool(args.json_stream)
    if json_mode:
        cfg.force_quiet = True
    # --assert-invariants is parsed but never reaches the config in the reference
    # either (cli.py:69 vs :85-101); kept as an accepted no-op flag.
    if cfg.seed is not None:
        import numpy as _np
        _np.random.seed(int(cfg.seed))
    interval = None if args.checkpoint_interval in (None, 0) else args.checkpoint_interval
    last = None
    try:
        for st, residuals, diag in driver.run(max_steps=args.steps, start_iteration=start_it,
                                              progress=(args.progress and not json_mode),
                                              checkpoint_path=args.checkpoint, checkpoint_interval=interval):
            last = (st, diag)
            if json_mode:
                sys.stdout.write(json.dumps(_step_payload(cfg, residuals, diag), separators=(",", ":")) + "\n")
                sys.stdout.flush()
            thr = args.continuity_threshold
            if thr is not None and residuals.get("continuity", 1.0) < thr:
                if not json_mode:
                    print(f"Stopping early: continuity {residuals['continuity']:.3e} < threshold {thr:.3e}")
                break
    except KeyboardInterrupt:
        if not json_mode:
            print("KeyboardInterrupt received - attempting graceful checkpoint...")
        if last is not None:
            path = args.checkpoint or "interrupt_final_checkpoint.npz"
            try:
                from .io.checkpoint import save_checkpoint
                driver.wait_checkpoints()        # a periodic checkpoint still in flight must not land after this one
                save_checkpoint(path, last[0], last[1].get("iteration", 0), last[1].get("wall_time", 0.0), cfg)
                if not json_mode:
                    print(f"Final checkpoint written: {path}")
            except Exception as exc:  # pragma: no cover
                if not json_mode:
                    print(f"Failed to write final checkpoint: {exc}")
        return 130
    return 0


if __name__ == "__main__":  # pragma: no cover
    raise SystemExit(main())
