"""Time the UNMODIFIED reference (bytecode in oracle/_ref, see build_ref.py) on THIS machine's host cores.

MEASUREMENT INFRASTRUCTURE: executed only by ``bench.py``'s CPU legs, as a subprocess (the reference package and the
drop-in package are both called ``pyflow``).  Drives the reference's own public API --
``pyflow.drivers.simulation_driver.SimulationDriver(config).run(max_steps=...)`` (reference
src/pyflow/drivers/simulation_driver.py:58-110) -- on a cavity from rest and prints ONE JSON line:

    {"nx":..,"ny":..,"Re":..,"lin_maxiter":..,"steps_timed":K,"s_per_step":..,"MLUPS":..,"threads":..,
     "final":{"u_norm":..,"v_norm":..,"p_norm":..,"dt":..,"continuity":..}}

Step 0 from rest has b = 0 (no CG work) and is never timed.  ``final`` lets the caller relate the run to the GPU
path on the same configuration (norms of the fields after the last step).

Usage:  python oracle/time_unmodified.py NX NY RE STEPS [LIN_MAXITER]
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import sys
import time
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
sys.path.insert(0, HERE)


def main(argv):
    nx, ny, Re, steps = int(argv[1]), int(argv[2]), float(argv[3]), int(argv[4])
    maxit = int(argv[5]) if len(argv) > 5 else 300
    stamp = json.load(open(os.path.join(REF, "stamp.json")))
    if stamp.get("magic") != importlib.util.MAGIC_NUMBER.hex():
        raise SystemExit(f"oracle/_ref was compiled by Python {stamp.get('python')}, this is {sys.version.split()[0]}")
    import build_ref
    build_ref.install_importer()
    import numpy as np
    from pyflow.drivers.simulation_driver import SimulationDriver
    assert os.path.dirname(os.path.dirname(os.path.abspath(sys.modules["pyflow"].__file__))) == REF, "wrong pyflow imported"
    cfg = SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=maxit, advection_scheme="quick", disable_advection=False, diagnostics=False)
    gen = SimulationDriver(cfg).run(max_steps=steps + 1)
    last = None
    with contextlib.redirect_stdout(io.StringIO()):
        next(gen)
        t0 = time.perf_counter()
        for last in gen:
            pass
        dt = time.perf_counter() - t0
    state, res, diag = last
    f = {k: np.asarray(state.fields[k])[1:-1, 1:-1] for k in ("u", "v", "p")}
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = None
    print(json.dumps({
        "nx": nx, "ny": ny, "Re": Re, "lin_maxiter": maxit, "steps_timed": steps, "s_per_step": dt / steps,
        "MLUPS": nx * ny * steps / dt / 1e6, "threads": threads,
        "final": {"u_norm": float(np.linalg.norm(f["u"])), "v_norm": float(np.linalg.norm(f["v"])),
                  "p_norm": float(np.linalg.norm(f["p"])), "dt": float(diag["dt"]), "continuity": float(res["continuity"])}}))


if __name__ == "__main__":
    main(sys.argv)
