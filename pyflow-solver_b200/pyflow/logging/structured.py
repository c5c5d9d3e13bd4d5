"""JSONL logger (reference ``src/pyflow/logging/structured.py`` :14-54): one
compact JSON object per line to a file or an in-memory stream, a ``ts`` stamp
added when missing, a module lock, a weak registry for ``close_all_jsonl_logs``."""
from __future__ import annotations

import json
import os
import threading
import time
import weakref
from typing import IO, Any

_write_lock = threading.Lock()
_registry: "weakref.WeakSet[JsonlLogger]" = weakref.WeakSet()


class JsonlLogger:
    def __init__(self, path: str | None = None, stream: IO[str] | None = None):
        if path is None and stream is None:
            raise ValueError("JsonlLogger requires either path or stream")
        self._stream = stream
        self._file = None
        if path is not None:
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            self._file = open(path, "a", encoding="utf-8", buffering=1)
        _registry.add(self)

    def log(self, record: dict[str, Any]):
        sink = self._stream if self._stream is not None else self._file
        if sink is None:
            return
        with _write_lock:
            record.setdefault("ts", time.time())
            sink.write(json.dumps(record, separators=(",", ":")) + "\n")
            try:
                sink.flush()
            except Exception:
                pass

    def close(self):
        if self._file is not None:
            try:
                self._file.close()
            except Exception:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def close_all_jsonl_logs():
    for lg in list(_registry):
        try:
            lg.close()
        except Exception:
            pass


__all__ = ["JsonlLogger", "close_all_jsonl_logs"]
