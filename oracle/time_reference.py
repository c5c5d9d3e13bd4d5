"""Time the UNMODIFIED reference next to its two restatements on THIS machine (TEST INFRASTRUCTURE).

Runs only where ``/root/reference`` exists (the build container; the GPU box never sees the reference, which
is why ``bench.py``'s CPU legs use the restatements).  Writes one JSON document with, per configuration, the
seconds per step and MLUPS of

  reference   ``pyflow.drivers.simulation_driver.SimulationDriver.run`` of /root/reference/src, as shipped
  numpy_port  ``oracle/pyflow_oracle.py``  (bitwise equal to the reference, vectorised advection, analytic Jacobi)
  c_port      ``oracle/pyflow_oracle.c``   (OpenMP, all host threads; what bench.py reports as cpu_baseline)

so that the ``cpu_baseline`` of a bench line (kind "port") can be related to the reference itself.

Usage:  python oracle/time_reference.py [out.json]
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
sys.path.insert(0, HERE)

CONFIGS = [
    # name, nx, ny, Re, timed steps (after step 0, which has b = 0 and no CG work)
    ("config 1: 64x64 Re=100 (all 50 steps of `pyflow run`; step 0 untimed)", 64, 64, 100.0, 49),
    ("config 2: 512x512 Re=1000", 512, 512, 1000.0, 3),
]


def cfg_of(nx, ny, Re):
    return SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                           lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)


def time_reference(nx, ny, Re, steps):
    sys.path.insert(0, REF_SRC)
    try:
        from pyflow.drivers.simulation_driver import SimulationDriver
        gen = SimulationDriver(cfg_of(nx, ny, Re)).run(max_steps=steps + 1)
        with contextlib.redirect_stdout(io.StringIO()):
            next(gen)
            t0 = time.perf_counter()
            for _ in gen:
                pass
        return (time.perf_counter() - t0) / steps
    finally:
        sys.path.remove(REF_SRC)
        for k in [k for k in sys.modules if k == "pyflow" or k.startswith("pyflow.")]:
            del sys.modules[k]


def time_port(nx, ny, Re, steps, impl):
    import pyflow_oracle as orc
    cfg = cfg_of(nx, ny, Re)
    st = orc.allocate_state(nx, ny)
    if impl == "c":
        import pyflow_oracle_c as orc_c
        ws = orc_c.Workspace(nx, ny)
        stepper = lambda it: orc_c.step(cfg, st, it, ws)       # noqa: E731
    else:
        stepper = lambda it: orc.step(cfg, st, it)             # noqa: E731
    stepper(0)
    t0 = time.perf_counter()
    for it in range(1, steps + 1):
        stepper(it)
    return (time.perf_counter() - t0) / steps


def main():
    if not os.path.isdir(REF_SRC):
        print("the reference is not present on this machine")
        return 1
    import pyflow_oracle_c as orc_c
    out = {"host_threads": orc_c.threads(), "cpu_count": os.cpu_count(), "lin_maxiter": 300,
           "note": "seconds per step over steps >= 1 from rest (300 CG iterations each); MLUPS = nx*ny/s/1e6",
           "configs": {}}
    for name, nx, ny, Re, steps in CONFIGS:
        row = {}
        for label, fn in (("reference", lambda: time_reference(nx, ny, Re, steps)),
                          ("numpy_port", lambda: time_port(nx, ny, Re, steps, "numpy")),
                          ("c_port", lambda: time_port(nx, ny, Re, steps, "c"))):
            sps = fn()
            row[label] = {"s_per_step": sps, "MLUPS": nx * ny / sps / 1e6, "steps_timed": steps}
            print(f"{name:28s} {label:10s} {sps * 1e3:10.2f} ms/step  {row[label]['MLUPS']:8.4f} MLUPS", flush=True)
        row["reference_over_c_port"] = row["reference"]["s_per_step"] / row["c_port"]["s_per_step"]
        out["configs"][name] = row
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "profiles",
                                                             "r1_reference_cpu_timings.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("written", path)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
