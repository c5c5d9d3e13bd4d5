"""General grid spacing (lx, ly given: dx, dy != 1) on one GPU: parity with the oracle, then the timing.

    python profiles/experiments/general_spacing_check.py

The general-spacing instantiations evaluate the reference's divisions (`(p[i+1] - p[i]) / dx` ..., twelve per cell and
CG iteration) with `div_by_invariant` (csrc/pf_common.cuh): reciprocal + two FMA corrections, bit for bit the IEEE
quotient (tests/test_exact_division.py).  Here: (1) whole steps against the C oracle on grids that take each CG kernel
(on-chip, register-window, TMA-fed), fused and separate per-step passes; (2) the element-wise operators bit for bit
against the numpy oracle; (3) step / CG-launch times of the 4096^2 and 512^2 cavities with lx = ly = 1 next to the
unit-spacing ones.  Prints GSPACE lines; exit code 1 if a check fails.
"""
from __future__ import annotations

import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "pyflow-solver_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import parity                   # noqa: E402
import pyflow_oracle as orc     # noqa: E402


def timing(nx, ny, steps, **spacing):
    import torch
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = parity.make_cfg(nx, ny, 1000.0, **spacing)
    st, tr = allocate_state(nx, ny), ResidualManager()
    for it in range(3):
        step(cfg, st, tr, it)
    torch.cuda.synchronize()
    cg, whole = [], []
    for it in range(3, 3 + steps):
        step(cfg, st, tr, it)
        cg.append(st.meta["cg_ms"])
        whole.append(st.meta["step_ms"])
    st.fields._slab.close()
    cg_ms, step_ms = statistics.mean(cg), statistics.mean(whole)
    gbs = nx * ny * 64.0 * 300 / (cg_ms * 1e-3) / 1e9
    print(f"GSPACE time {nx}x{ny} spacing={spacing or 'unit'}: step {step_ms:.3f} ms, CG launch {cg_ms:.3f} ms = "
          f"{gbs:.0f} GB/s algorithmic, {nx * ny / step_ms / 1e3:.1f} MLUPS", flush=True)


def main():
    ok = True
    parity.use_all_host_threads()
    cases = [("59x47 upwind (register-window kernel)", parity.make_cfg(59, 47, 100.0, "upwind", 30, lx=3.0, ly=2.0), 4, None),
             ("59x47 upwind, separate per-step kernels", parity.make_cfg(59, 47, 100.0, "upwind", 30, lx=3.0, ly=2.0), 4, "step_fused=0"),
             ("200x150 quick (on-chip kernel)", parity.make_cfg(200, 150, 1000.0, "quick", 40, lx=7.3, ly=2.9), 4, None),
             ("1100x130 quick (TMA-fed kernel, clipped chunk)", parity.make_cfg(1100, 130, 1000.0, "quick", 120, lx=1.0, ly=0.37), 3, None),
             ("2048x600 quick (TMA-fed kernel), separate per-step kernels", parity.make_cfg(2048, 600, 1000.0, "quick", 60, lx=1.0, ly=1.0), 3, "step_fused=0")]
    for name, cfg, steps, options in cases:
        r = parity.compare(name, parity.gpu_run(cfg, steps, options=options), parity.oracle_run(cfg, steps))
        print(f"GSPACE parity {name}: rel_l2 {r['max_rel_l2']:.2e} dt_equal {r['dt_equal']} iters {r['iters']} "
              f"oracle {r['oracle_iters']} {'OK' if r['ok'] else 'FAIL'}", flush=True)
        ok = ok and r["ok"]
    # element-wise operators, bit for bit (NaN included)
    from pyflow.numerics import core_ops_coherent as coh
    from pyflow.numerics.operators import advection as adv
    rng = np.random.default_rng(5)
    for shape, (dx, dy) in (((257, 513), (0.37, 1.9)), ((300, 1030), (1.0 / 1029, 1.0 / 299)), ((64, 3), (3.0, 0.1))):
        p, u, v = (rng.standard_normal(shape) for _ in range(3))
        p[shape[0] // 2, shape[1] // 2] = np.nan
        u[rng.random(shape) < 0.1] = 0.0
        with np.errstate(invalid="ignore"):
            gx, gy = coh.gradient_coherent(p, dx, dy)
            ogx, ogy = orc.gradient_coherent(p, dx, dy)
            same = (np.array_equal(gx, ogx, equal_nan=True) and np.array_equal(gy, ogy, equal_nan=True)
                    and np.array_equal(coh.laplacian_coherent(p, dx, dy), orc.laplacian_coherent(p, dx, dy), equal_nan=True)
                    and np.array_equal(coh.divergence_coherent(u, v, dx, dy), orc.divergence_coherent(u, v, dx, dy))
                    and np.array_equal(adv.advect_quick(u, v, u, dx, dy), orc.advect_quick(u, v, u, dx, dy))
                    and np.array_equal(adv.advect_upwind(u, v, v, dx, dy), orc.advect_upwind(u, v, v, dx, dy)))
        print(f"GSPACE operators {shape} dx={dx:.4g} dy={dy:.4g}: {'bit-identical' if same else 'DIFFER'}", flush=True)
        ok = ok and same
    timing(4096, 4096, 5, lx=1.0, ly=1.0)
    timing(4096, 4096, 5)
    timing(512, 512, 20, lx=1.0, ly=1.0)
    timing(512, 512, 20)
    print("GSPACE_ALL_OK" if ok else "GSPACE_FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
