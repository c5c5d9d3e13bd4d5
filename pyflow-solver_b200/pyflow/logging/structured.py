"""JSON-lines event log of the drop-in (same surface and record format as the
reference's ``src/pyflow/logging/structured.py`` :14-54).

Contract kept for the callers on the path (``solvers/solver.py`` writes one
``step`` record per time step, plus ``error`` / ``emergency_checkpoint`` records
when a NaN shows up):

* ``JsonlLogger(path=...)`` appends to a file, ``JsonlLogger(stream=...)``
  writes to any text stream (tests pass a ``StringIO``); one of the two is
  required, the stream wins when both are given;
* ``log(record)`` stamps ``record["ts"]`` (wall clock) unless the caller set it,
  and emits the record as ONE compact JSON line, flushed immediately so that a
  tailing consumer sees complete lines only;
* writers in different threads never interleave lines (one process-wide mutex);
* ``close_all_jsonl_logs()`` closes every file the process still has open
  (test fixtures call it between tests); closing twice is harmless.
"""
from __future__ import annotations

import json
import threading
import time
import weakref
from pathlib import Path
from typing import IO, Any

_ENCODER = json.JSONEncoder(separators=(",", ":"))
_MUTEX = threading.Lock()
_OPEN_FILES: dict[int, "weakref.ref[JsonlLogger]"] = {}


class JsonlLogger:
    """Line-oriented JSON writer bound to a file it owns or a stream it borrows."""

    def __init__(self, path: str | None = None, stream: IO[str] | None = None):
        if stream is None and path is None:
            raise ValueError("JsonlLogger requires either path or stream")
        self._borrowed = stream
        self._owned = None
        if path is not None:
            target = Path(path)
            target.parent.mkdir(parents=True, exist_ok=True)
            self._owned = target.open("a", encoding="utf-8", buffering=1)      # line buffered
            key = id(self)
            _OPEN_FILES[key] = weakref.ref(self, lambda _r, key=key: _OPEN_FILES.pop(key, None))

    @property
    def _out(self):
        return self._borrowed if self._borrowed is not None else self._owned

    def log(self, record: dict[str, Any]) -> None:
        out = self._out
        if out is None or getattr(out, "closed", False):
            return
        if "ts" not in record:
            record["ts"] = time.time()
        line = _ENCODER.encode(record)
        with _MUTEX:
            out.write(line)
            out.write("\n")
            flush = getattr(out, "flush", None)
            if flush is not None:
                try:
                    flush()
                except (OSError, ValueError):
                    pass

    def close(self) -> None:
        fh, self._owned = self._owned, None
        if fh is not None:
            _OPEN_FILES.pop(id(self), None)
            try:
                fh.close()
            except OSError:
                pass

    def __enter__(self) -> "JsonlLogger":
        return self

    def __exit__(self, *_exc) -> None:
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def close_all_jsonl_logs() -> None:
    """Close every file-backed logger that is still alive."""
    for ref in list(_OPEN_FILES.values()):
        logger = ref()
        if logger is not None:
            logger.close()


__all__ = ["JsonlLogger", "close_all_jsonl_logs"]
