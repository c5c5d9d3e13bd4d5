"""Checkpoint I/O with the reference's on-disk layout
(``src/pyflow/io/checkpoint.py`` :51-127): a compressed ``.npz`` holding

  meta          JSON string: iteration, time, nx, ny, config_hash (sha1[:10] of
                the config's public attributes), structured_config_hash
                (``SimulationConfig.config_hash`` or null), field_hashes
                (sha1[:16] of each array's raw bytes), seed, schema_version=1
  field_<name>  float64 (ny+2, nx+2) array INCLUDING the ghost ring

written atomically (temp file + fsync + ``os.replace``).

New: the time-step controller's memory -- ``dt_prev``,
``divergence_history``, ``prev_div_norm`` of ``state.meta`` -- rides along under
ONE extra meta key, ``step_state``.  The reference drops it, so its restarts
begin a fresh dt ramp; that stays the default here (``load_checkpoint`` returns a
State with empty ``meta`` exactly like the reference's).  ``restore_step_state``
(``pyflow run --restart ... --exact-restart`` or ``PYFLOW_B200_EXACT_RESTART=1``)
puts it back, and the resumed run continues the uninterrupted one: same dt
sequence, same fields.  Old readers, the reference's included, ignore unknown
meta keys, so files stay interchangeable both ways.

Device note: fields are pulled from HBM with the read-only accessor, so writing
a checkpoint does not force a re-upload on the next step.

``BackgroundCheckpointWriter`` (new) takes the expensive
part off the time loop: at 4096^2 the sha1 hashes and the deflate of 400 MB
cost seconds per checkpoint against a 53 ms step.  ``submit`` snapshots the three
fields straight from HBM into its own page-locked buffers (the host copies of
the State are not touched) and a worker thread hashes, compresses and renames
while the GPU keeps stepping; the bytes on disk are the same as the synchronous
writer's.  One write is in flight at most; ``wait`` joins it and re-raises its
error.
"""
from __future__ import annotations

import hashlib
import io
import json
import os
import tempfile
import threading
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


STEP_STATE_KEYS = ("dt_prev", "divergence_history", "prev_div_norm")


def _step_state(state) -> dict[str, Any] | None:
    """The dt controller's memory (solver.py:27-62 of the reference keeps it in ``state.meta``) as plain JSON
    values; floats survive ``json`` exactly (shortest round-trip repr, NaN / Infinity literals)."""
    src = getattr(state, "meta", None)
    if not isinstance(src, dict):
        return None
    out: dict[str, Any] = {}
    for key in STEP_STATE_KEYS:
        if key in src:
            v = src[key]
            out[key] = [float(t) for t in v] if isinstance(v, (list, tuple)) else float(v)
    return out or None


def restore_step_state(state, meta: dict[str, Any]) -> bool:
    """Copy a checkpoint's ``step_state`` into ``state.meta`` (opt-in exact restart).  False when the file has
    none (written by the reference or by an older version): the run then restarts the dt ramp as before."""
    saved = meta.get("step_state") if isinstance(meta, dict) else None
    if not isinstance(saved, dict):
        return False
    for key in STEP_STATE_KEYS:
        if key in saved:
            v = saved[key]
            state.meta[key] = [float(t) for t in v] if isinstance(v, list) else float(v)
    return True


def exact_restart_requested(flag: bool = False) -> bool:
    return bool(flag) or os.environ.get("PYFLOW_B200_EXACT_RESTART", "0") == "1"


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
    step_state = _step_state(state)
    if step_state is not None:                      # new key, after the reference's own (see the module docstring)
        meta["step_state"] = step_state
    return meta


def save_checkpoint(path: str, state, iteration: int, sim_time: float, cfg: Any) -> str:
    fields = _host_fields(state)
    meta = _meta(state, fields, iteration, sim_time, cfg, structured=True)
    return _write_npz(path, meta, fields)


def _write_npz(path: str, meta: dict[str, Any], fields: dict[str, np.ndarray]) -> str:
    folder = os.path.dirname(path) or "."
    os.makedirs(folder, exist_ok=True)
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


class BackgroundCheckpointWriter:
    """Snapshot now, hash / compress / rename on a worker thread (see the module docstring)."""

    def __init__(self):
        self._thread: threading.Thread | None = None
        self._error: BaseException | None = None
        self._buffers: dict[str, np.ndarray] = {}
        self.written = 0

    def _buffer(self, name: str, like: np.ndarray) -> np.ndarray:
        buf = self._buffers.get(name)
        if buf is None or buf.shape != like.shape:
            from ..core.ghost_fields import _zeros
            buf = self._buffers[name] = _zeros(like.shape)
        return buf

    def _snapshot(self, state) -> dict[str, np.ndarray]:
        f = state.fields
        lazy = isinstance(f, SyncedFields)
        snap = {}
        for k in list(f.keys()):
            host = f.peek(k) if lazy else f[k]
            buf = self._buffer(k, host)
            if lazy and k in f._stale and f._slab is not None:
                # newest copy is in HBM: D2H straight into the snapshot; the ghost ring comes from the host array
                buf[0, :], buf[-1, :], buf[:, 0], buf[:, -1] = host[0, :], host[-1, :], host[:, 0], host[:, -1]
                f._slab.download(**{k: buf})
            else:
                np.copyto(buf, host)
            snap[k] = buf
        return snap

    def submit(self, path: str, state, iteration: int, sim_time: float, cfg: Any) -> None:
        """Returns as soon as the fields are copied; the file appears when the worker is done."""
        self.wait()                                    # one write in flight; its buffers are free again
        fields = self._snapshot(state)
        from types import SimpleNamespace
        shape = SimpleNamespace(nx=state.nx, ny=state.ny, meta=dict(getattr(state, "meta", None) or {}))
        head = _meta(shape, {}, iteration, sim_time, cfg, structured=True)     # config hashes, step state: taken now

        def work():
            try:
                hashes = {k: _hash_array(v) for k, v in fields.items()}
                meta = {k: (hashes if k == "field_hashes" else v) for k, v in head.items()}   # key order kept
                _write_npz(path, meta, fields)
                self.written += 1
            except BaseException as exc:           # surfaced by wait()
                self._error = exc

        self._thread = threading.Thread(target=work, name="pyflow-checkpoint", daemon=False)
        self._thread.start()

    def wait(self) -> None:
        t, self._thread = self._thread, None
        if t is not None:
            t.join()
        err, self._error = self._error, None
        if err is not None:
            raise err


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


__all__ = ["save_checkpoint", "load_checkpoint", "save_checkpoint_bytes", "load_checkpoint_bytes",
           "BackgroundCheckpointWriter", "restore_step_state", "exact_restart_requested", "STEP_STATE_KEYS"]
