"""Simulation configuration of the drop-in.

Schema contract (the reference's ``src/pyflow/config/model.py`` :15-99 defines
it; checkpoints, ``--restart`` hash enforcement and ``pyflow validate`` depend on
it, so field names, defaults and ranges must match exactly):

  field                      default   accepted
  nx, ny                     required  >= 3
  Re                         100.0     > 0
  lid_velocity               1.0       any float
  cfl_target                 0.5       (0, 1]
  cfl_growth                 1.05      (1, 1.2]
  lin_tol                    1e-10     [1e-14, 1e-2]
  lin_maxiter                200       [1, 10000]
  advection_scheme           "upwind"  "upwind" | "quick"
  disable_advection          False
  diagnostics                True
  log_path, log_stream       None      mutually exclusive
  force_quiet                False     (set by the CLI under --json-stream)
  checkpoint_interval        None      >= 1
  emergency_checkpoint_path  None
  seed                       None      [0, 2^32 - 1]
  schema_version             1
  turbulence_model           None      reserved

``config_hash`` (:75-86): sha1 over the canonical JSON -- keys sorted, separators
"," and ":" -- of every field except the two runtime-only ones (``log_stream``,
``force_quiet``), first 12 hex digits.  Known answers:
CLI defaults at ``nx=ny=64`` -> ``8dddaac6b3f9``; ``nx=ny=4096, Re=1000`` ->
``4de9eeab2d1d`` (asserted in tests/test_host_logic.py against fixtures written
by the reference).
"""
from __future__ import annotations

import hashlib
import json
from typing import Annotated, Any, Literal, Optional

from pydantic import BaseModel, Field, model_validator

RUNTIME_ONLY_FIELDS = frozenset({"log_stream", "force_quiet"})
HASH_HEX_DIGITS = 12

GridCount = Annotated[int, Field(ge=3)]
UnitIntervalOpenLow = Annotated[float, Field(gt=0.0, le=1.0)]
GrowthFactor = Annotated[float, Field(gt=1.0, le=1.2)]
SolverTolerance = Annotated[float, Field(ge=1e-14, le=1e-2)]
IterationCap = Annotated[int, Field(ge=1, le=10000)]
Seed32 = Annotated[int, Field(ge=0, le=2 ** 32 - 1)]
PositiveInt = Annotated[int, Field(ge=1)]


class ConfigError(ValueError):
    """Cross-field inconsistency of a configuration."""


def canonical_json(payload: dict) -> bytes:
    return json.dumps(payload, sort_keys=True, separators=(",", ":")).encode()


def short_digest(payload: dict, digits: int = HASH_HEX_DIGITS) -> str:
    return hashlib.sha1(canonical_json(payload)).hexdigest()[:digits]


class SimulationConfig(BaseModel):
    # --- grid and flow
    nx: GridCount
    ny: GridCount
    Re: float = Field(default=100.0, gt=0.0)
    lid_velocity: float = 1.0
    # --- advective predictor
    advection_scheme: Literal["upwind", "quick"] = "upwind"
    disable_advection: bool = False
    # --- adaptive time step
    cfl_target: UnitIntervalOpenLow = 0.5
    cfl_growth: GrowthFactor = 1.05
    # --- pressure Poisson solve
    lin_tol: SolverTolerance = 1e-10
    lin_maxiter: IterationCap = 200
    # --- output
    diagnostics: bool = True
    log_path: Optional[str] = None
    log_stream: Optional[Any] = None
    force_quiet: bool = False
    checkpoint_interval: Optional[PositiveInt] = None
    emergency_checkpoint_path: Optional[str] = None
    # --- bookkeeping
    seed: Optional[Seed32] = None
    schema_version: int = 1
    turbulence_model: Optional[str] = None

    @model_validator(mode="after")
    def _consistency(self):
        if self.log_stream is not None and self.log_path:
            raise ConfigError("log_path and log_stream are mutually exclusive")
        ratio = self.nx / self.ny
        if not 0.1 <= ratio <= 10:
            # soft warning only, kept where the reference keeps it
            object.__setattr__(self, "_soft_warning", f"Extreme aspect ratio nx/ny={ratio:.2f}")
        return self

    # ---- hashing ---------------------------------------------------------------
    def hash_payload(self) -> dict:
        return {k: v for k, v in self.model_dump().items() if k not in RUNTIME_ONLY_FIELDS}

    @property
    def config_hash(self) -> str:
        return short_digest(self.hash_payload())

    # ---- short description printed by `pyflow show-config` -----------------------
    def brief(self) -> dict:
        scheme = None if self.disable_advection else self.advection_scheme
        out = {name: getattr(self, name) for name in ("nx", "ny", "Re")}
        out["scheme"] = scheme
        for name in ("lin_tol", "lin_maxiter", "cfl_target", "cfl_growth"):
            out[name] = getattr(self, name)
        out["hash"] = self.config_hash
        return out


__all__ = ["ConfigError", "SimulationConfig"]
