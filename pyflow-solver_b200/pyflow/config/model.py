"""Simulation configuration (same schema, validation ranges and ``config_hash``
as the reference's ``src/pyflow/config/model.py`` :15-99, so checkpoints and
restart hash enforcement stay interchangeable).

``config_hash`` = first 12 hex digits of sha1 over the canonical JSON
(sorted keys, compact separators) of every field except the runtime-only
``log_stream`` and ``force_quiet`` (:75-86).  Known answers (SURVEY.md
section 5): ``nx=ny=64`` CLI defaults -> ``8dddaac6b3f9``; ``nx=ny=4096,
Re=1000`` -> ``4de9eeab2d1d``.
"""
from __future__ import annotations

import hashlib
import json
from typing import Any, Literal

from pydantic import BaseModel, Field, model_validator


class ConfigError(ValueError):
    pass


AdvectionScheme = Literal["upwind", "quick"]

_NOT_HASHED = ("log_stream", "force_quiet")


class SimulationConfig(BaseModel):
    # grid / physics
    nx: int = Field(..., ge=3, description="interior cells in x")
    ny: int = Field(..., ge=3, description="interior cells in y")
    Re: float = Field(100.0, gt=0.0, description="Reynolds number")
    lid_velocity: float = Field(1.0, description="tangential velocity of the top lid")
    # time stepping
    cfl_target: float = Field(0.5, gt=0.0, le=1.0)
    cfl_growth: float = Field(1.05, gt=1.0, le=1.2)
    # pressure solve
    lin_tol: float = Field(1e-10, ge=1e-14, le=1e-2)
    lin_maxiter: int = Field(200, ge=1, le=10000)
    # advection
    advection_scheme: AdvectionScheme = "upwind"
    disable_advection: bool = False
    # diagnostics / logging
    diagnostics: bool = True
    log_path: str | None = None
    log_stream: Any | None = None     # in-memory sink for tests; not hashed
    force_quiet: bool = False         # set by the CLI under --json-stream; not hashed
    # checkpointing
    checkpoint_interval: int | None = Field(None, ge=1)
    emergency_checkpoint_path: str | None = None
    # reproducibility / versioning
    seed: int | None = Field(None, ge=0, le=2 ** 32 - 1)
    schema_version: int = 1
    turbulence_model: str | None = Field(None, description="reserved")

    @model_validator(mode="after")
    def _cross_checks(self):
        problems: list[str] = []
        if self.log_path and self.log_stream is not None:
            problems.append("log_path and log_stream are mutually exclusive")
        aspect = max(self.nx, 1) / max(self.ny, 1)
        if aspect > 10 or aspect < 0.1:
            object.__setattr__(self, "_soft_warning", f"Extreme aspect ratio nx/ny={aspect:.2f}")
        if problems:
            raise ConfigError("; ".join(problems))
        return self

    def hash_payload(self) -> dict:
        data = self.model_dump()
        for key in _NOT_HASHED:
            data.pop(key, None)
        return data

    @property
    def config_hash(self) -> str:
        blob = json.dumps(self.hash_payload(), sort_keys=True, separators=(",", ":")).encode()
        return hashlib.sha1(blob).hexdigest()[:12]

    def brief(self) -> dict:
        return {
            "nx": self.nx, "ny": self.ny, "Re": self.Re,
            "scheme": None if self.disable_advection else self.advection_scheme,
            "lin_tol": self.lin_tol, "lin_maxiter": self.lin_maxiter,
            "cfl_target": self.cfl_target, "cfl_growth": self.cfl_growth,
            "hash": self.config_hash,
        }


__all__ = ["ConfigError", "SimulationConfig"]
