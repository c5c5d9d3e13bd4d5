"""Small runs of every CG kernel variant for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tests/sanitize_cases.py

Few iterations on small grids -- the tools slow a kernel down by two or three orders of magnitude -- but every
code path that synchronises: the on-chip kernel with its LL-word exchanges (k_cg_small), the TMA ring with mbarriers,
tickets, lean and generic row loops, the fence-free group fold and the prefetch across the barrier (k_cg_tmad), the
static TMA kernel, the register-window kernel, the plain and the BiCGSTAB kernels, and whole projection steps.
Every result is checked against the numpy oracle, so a tool that perturbs the schedule cannot hide a wrong answer.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pyflow-solver_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import pyflow_oracle as orc                                    # noqa: E402
from pyflow._b200.device import DeviceSlab                     # noqa: E402


def rhs(ny, nx, seed):
    b = np.random.default_rng(seed).standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    return b


def check(tag, x, ny, nx, b, maxiter, dx=1.0, dy=1.0):
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)          # noqa: E731
    xo, _info, _it = orc.pcg(A, b.ravel().copy(), 1e-10, maxiter, orc.jacobi_inverse_diagonal(nx, ny, dx, dy))
    err = float(np.linalg.norm(x.ravel() - xo) / np.linalg.norm(xo))
    print(f"SANITIZE {tag}: rel err vs oracle {err:.2e} {'OK' if err < 1e-9 else 'FAIL'}", flush=True)
    return err < 1e-9


def steps_only():
    """`python tests/sanitize_cases.py steps`: whole projection steps only -- the two fused per-step passes
    (pf_step_fused.cuh: shuffles between the lanes of a warp, clamped loads, out-of-place u / v) and the seven
    separate kernels (step_fused=0), on an even, an odd and a TMA-kernel width, both advection schemes."""
    from types import SimpleNamespace
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    ok = True
    for opts in (None, "step_fused=0"):
        for nx, ny, maxiter, scheme in ((64, 48, 20, "quick"), (59, 37, 15, "upwind"), (117, 9, 10, "quick"),
                                        (1024, 1100, 5, "quick")):
            if opts:
                os.environ["PYFLOW_B200_OPTIONS"] = opts
            else:
                os.environ.pop("PYFLOW_B200_OPTIONS", None)
            cfg = SimpleNamespace(nx=nx, ny=ny, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05,
                                  lin_tol=1e-10, lin_maxiter=maxiter, advection_scheme=scheme, disable_advection=False,
                                  diagnostics=False)
            st, tr = allocate_state(nx, ny), ResidualManager()
            for it in range(3):
                step(cfg, st, tr, it)
            ost, _od = orc.run(cfg, 3)
            err = max(float(np.linalg.norm(st.fields[f][1:-1, 1:-1] - orc.interior(ost.fields[f])) /
                            (np.linalg.norm(orc.interior(ost.fields[f])) or 1.0)) for f in ("u", "v", "p"))
            print(f"SANITIZE step {nx}x{ny} {scheme} options={opts}: max rel err vs oracle {err:.2e} "
                  f"{'OK' if err < 1e-9 else 'FAIL'}", flush=True)
            ok = ok and err < 1e-9
            st.fields._slab.close()
    os.environ.pop("PYFLOW_B200_OPTIONS", None)
    print("SANITIZE_ALL_OK" if ok else "SANITIZE_FAILED", flush=True)
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "steps":
        return steps_only()
    ok = True
    cases = [
        # tag, ny, nx, options, maxiter
        ("k_cg_small 96x80 (one-row bands)", 80, 96, None, 6),
        ("k_cg_small 512x300 (NP=2)", 300, 512, None, 4),
        ("k_cg_tmad 1024x1200 (lean + generic tiles, group fold, prefetch)", 1200, 1024, None, 4),
        ("k_cg_tmad 2048x700 (2 chunks)", 700, 2048, None, 3),
        ("k_cg_tma static 1024x1200", 1200, 1024, "cg_kernel=2", 3),
        ("k_cg_persistent 1100x40", 40, 1100, None, 4),
        ("k_cg_persistent forced 1024x600", 600, 1024, "cg_kernel=1", 3),
    ]
    for tag, ny, nx, opts, maxiter in cases:
        if opts:
            os.environ["PYFLOW_B200_OPTIONS"] = opts
        else:
            os.environ.pop("PYFLOW_B200_OPTIONS", None)
        slab = DeviceSlab(nx, ny)
        b = rhs(ny, nx, nx + ny)
        x, _info, _iters, _rn = slab.pcg(b, 1.0, 1.0, 1e-10, maxiter)
        ok = check(tag, x, ny, nx, b, maxiter) and ok
        slab.close()
    os.environ.pop("PYFLOW_B200_OPTIONS", None)
    # plain CG (q stored) and BiCGSTAB kernels
    slab = DeviceSlab(70, 66)
    b = rhs(66, 70, 5)
    x, *_ = slab.cg_generic("pinned", b, 1.0, 1.0, 1e-10, 5, jacobi=True, inv_diag_interior=0.25)
    ok = check("k_cg_generic 70x66", x, 66, 70, b, 5) and ok
    xb, info, iters, _rn = slab.bicgstab(b, 1.0, 1.0, 1e-10, 4)
    print(f"SANITIZE k_bicgstab 70x66: info {info} iters {iters} finite {bool(np.isfinite(xb).all())}", flush=True)
    ok = ok and bool(np.isfinite(xb).all())
    slab.close()
    # whole projection steps (every per-step kernel), small-kernel and TMA-kernel grids
    from types import SimpleNamespace
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    for nx, ny, maxiter in ((64, 48, 20), (1024, 1100, 5)):
        cfg = SimpleNamespace(nx=nx, ny=ny, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                              lin_maxiter=maxiter, advection_scheme="quick", disable_advection=False, diagnostics=False)
        st, tr = allocate_state(nx, ny), ResidualManager()
        for it in range(3):
            step(cfg, st, tr, it)
        ost, _od = orc.run(cfg, 3)
        err = max(float(np.linalg.norm(st.fields[f][1:-1, 1:-1] - orc.interior(ost.fields[f])) /
                        (np.linalg.norm(orc.interior(ost.fields[f])) or 1.0)) for f in ("u", "v", "p"))
        print(f"SANITIZE step {nx}x{ny}: max rel err vs oracle {err:.2e} {'OK' if err < 1e-9 else 'FAIL'}", flush=True)
        ok = ok and err < 1e-9
    print("SANITIZE_ALL_OK" if ok else "SANITIZE_FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
