"""A/B of the per-step work around the CG solve: the two fused passes (step_fused=1, pf_step_fused.cuh) against the
seven separate kernels (step_fused=0, pf_kernels.cuh) on one GPU.

    python profiles/experiments/step_passes_ab.py [NX NY [LIN_MAXITER [STEPS]]] ...

For each grid and each setting: `steps` device-resident steps after 3 warm-up steps, CUDA events of the library
(pf_step: whole step / CG launch).  Prints the mean of (step - CG) = everything that is not the solve, the step, MLUPS.
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel times (profiles/capture.sh steppasses).
"""
from __future__ import annotations

import os
import statistics
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pyflow-solver_b200"))


def run(nx, ny, maxit, steps, fused):
    import torch
    os.environ["PYFLOW_B200_OPTIONS"] = f"step_fused={fused}"
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = SimpleNamespace(nx=nx, ny=ny, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=maxit, advection_scheme="quick", disable_advection=False, diagnostics=False)
    st, tr = allocate_state(nx, ny), ResidualManager()
    it = 0
    for _ in range(3):
        step(cfg, st, tr, it)
        it += 1
    torch.cuda.synchronize()
    n0 = st.fields._slab.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    rest, whole = [], []
    for _ in range(steps):
        step(cfg, st, tr, it)
        it += 1
        rest.append(st.meta["step_ms"] - st.meta["cg_ms"])
        whole.append(st.meta["step_ms"])
    ev1.record()
    torch.cuda.synchronize()
    wall = ev0.elapsed_time(ev1) / steps
    launches = (st.fields._slab.launch_count() - n0) / steps
    st.fields._slab.close()
    return statistics.mean(rest), min(rest), statistics.mean(whole), wall, launches


def main(argv):
    grids = []
    a = [int(x) for x in argv[1:]]
    while a:
        nx, ny = a[0], a[1]
        maxit = a[2] if len(a) > 2 else 300
        steps = a[3] if len(a) > 3 else 10
        grids.append((nx, ny, maxit, steps))
        a = a[4:]
    if not grids:
        grids = [(4096, 4096, 300, 10), (512, 512, 300, 50), (4000, 4000, 300, 6), (16384, 4096, 20, 4)]
    for nx, ny, maxit, steps in grids:
        for fused in (1, 0, 1, 0):
            rest, rest_min, whole, wall, launches = run(nx, ny, maxit, steps, fused)
            print(f"STEPPASS {nx}x{ny} maxit={maxit} step_fused={fused}: non-CG part of a step mean {rest * 1e3:8.1f} us "
                  f"(min {rest_min * 1e3:8.1f}), device step {whole:8.3f} ms, step incl. host {wall:8.3f} ms, "
                  f"{nx * ny / wall / 1e3:7.1f} MLUPS, {launches:.1f} launches/step", flush=True)


if __name__ == "__main__":
    main(sys.argv)
