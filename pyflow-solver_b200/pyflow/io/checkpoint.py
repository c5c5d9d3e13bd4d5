"""Checkpoint I/O with the reference's on-disk layout
(``src/pyflow/io/checkpoint.py`` :51-127): a compressed ``.npz`` holding

  meta          JSON string: iteration, time, nx, ny, config_hash (sha1[:10] of
                the config's public attributes), structured_config_hash
                (``SimulationConfig.config_hash`` or null), field_hashes
                (sha1[:16] of each array's raw bytes), seed, schema_version=1
  field_<name>  float64 (ny+2, nx+2) array INCLUDING the ghost ring

written atomically (temp file + fsync + ``os.replace``).  ``state.meta`` is not
persisted (the reference drops ``dt_prev`` / ``divergence_history`` too, so a
restart resets dt -- kept for interchangeability).

Device note: fields are pulled from HBM with the read-only accessor, so writing
a checkpoint does not force a re-upload on the next step.
"""
from __future__ import annotations

import hashlib
import io
import json
import os
import tempfile
from typing import Any

import numpy as np

from ..config.model import SimulationConfig
from ..core.ghost_fields import State, SyncedFields


def _hash_config(cfg: Any) -> str:
    try:
        if hasattr(cfg, "__dict__"):
            data = {k: v for k, v in vars(cfg).items() if not k.startswith("_")}
        else:
            data = {k: getattr(cfg, k) for k in dir(cfg) if not k.startswith("_")}
        return hashlib.sha1(json.dumps(data, sort_keys=True, default=str).encode()).hexdigest()[:10]
    except Exception:
        return "unknown"


def _hash_array(arr: np.ndarray) -> str:
    return hashlib.sha1(arr.tobytes()).hexdigest()[:16]


def _host_fields(state) -> dict[str, np.ndarray]:
    f = state.fields
    if isinstance(f, SyncedFields):
        return {k: f.read(k) for k in f.keys()}
    return dict(f.items())


def _meta(state, fields, iteration, sim_time, cfg, structured: bool) -> dict[str, Any]:
    meta: dict[str, Any] = {
        "iteration": iteration, "time": sim_time, "nx": state.nx, "ny": state.ny,
        "config_hash": _hash_config(cfg),
    }
    if structured:
        sh = None
        if isinstance(cfg, SimulationConfig):
            try:
                sh = cfg.config_hash
            except Exception:
                sh = None
        meta["structured_config_hash"] = sh
    meta["field_hashes"] = {k: _hash_array(v) for k, v in fields.items()}
    meta["seed"] = getattr(cfg, "seed", None)
    meta["schema_version"] = 1
    return meta


def save_checkpoint(path: str, state, iteration: int, sim_time: float, cfg: Any) -> str:
    folder = os.path.dirname(path) or "."
    os.makedirs(folder, exist_ok=True)
    fields = _host_fields(state)
    meta = _meta(state, fields, iteration, sim_time, cfg, structured=True)
    fd, tmp = tempfile.mkstemp(suffix=".tmp", dir=folder)
    try:
        with os.fdopen(fd, "wb") as fh:
            np.savez_compressed(fh, meta=json.dumps(meta), **{f"field_{k}": v for k, v in fields.items()})
            fh.flush()
            try:
                os.fsync(fh.fileno())
            except OSError:
                pass
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            try:
                os.remove(tmp)
            except Exception:
                pass
    return path


def _state_from_npz(data) -> tuple[State, dict[str, Any]]:
    meta = json.loads(str(data["meta"]))
    fields = {k.removeprefix("field_"): data[k].copy() for k in data.files if k.startswith("field_")}
    return State(nx=meta.get("nx"), ny=meta.get("ny"), fields=fields), meta


def load_checkpoint(path: str) -> tuple[State, dict[str, Any]]:
    with np.load(path) as data:
        return _state_from_npz(data)


def save_checkpoint_bytes(state, iteration: int, sim_time: float, cfg: Any) -> bytes:
    fields = _host_fields(state)
    meta = _meta(state, fields, iteration, sim_time, cfg, structured=False)
    bio = io.BytesIO()
    np.savez_compressed(bio, meta=json.dumps(meta), **{f"field_{k}": v for k, v in fields.items()})
    return bio.getvalue()


def load_checkpoint_bytes(buf: bytes) -> tuple[State, dict[str, Any]]:
    with np.load(io.BytesIO(buf)) as data:
        return _state_from_npz(data)


__all__ = ["save_checkpoint", "load_checkpoint", "save_checkpoint_bytes", "load_checkpoint_bytes"]
