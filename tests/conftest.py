"""Test configuration.

* ``-m "not gpu"`` (runs anywhere): oracle vs the golden vectors generated from
  the unmodified reference, host-side logic, C-ABI symbol checks.
* ``-m gpu`` (needs a B200): parity of the CUDA path against the oracle and the
  golden vectors, through the C-ABI / the drop-in Python surface.
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pyflow-solver_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_library_built():
    """Build libpyflow_b200.so if it is missing or stale (nvcc cross-compiles on CPU)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pf_build", os.path.join(ROOT, "pyflow-solver_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build_native()
    except Exception as exc:  # the symbol tests will then fail loudly
        print(f"[conftest] native build failed: {exc}")
    yield


@pytest.fixture(autouse=True)
def _close_jsonl_logs():
    yield
    from pyflow.logging.structured import close_all_jsonl_logs
    close_all_jsonl_logs()


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    out = {}
    for name in ("operators", "legacy_ops", "pressure", "trajectories"):
        with np.load(os.path.join(GOLDEN, f"{name}.npz")) as z:
            out[name] = {k: z[k] for k in z.files}
    with open(os.path.join(GOLDEN, "trajectories.json")) as fh:
        out["trajectories_meta"] = json.load(fh)
    with open(os.path.join(GOLDEN, "config_hashes.json")) as fh:
        out["config"] = json.load(fh)
    return out
