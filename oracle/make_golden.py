"""Pin the oracle against the UNMODIFIED reference and emit golden vectors.

Runs only where ``/root/reference`` exists (the build container).  It
  1. imports the reference package from ``/root/reference/src``;
  2. checks ``oracle/pyflow_oracle.py`` BITWISE against it: operator trio,
     central divergence, both advection schemes, the Jacobi diagonal, the CG
     recurrence (vs scipy.sparse.linalg.cg) and whole multi-step trajectories;
  3. writes small fixtures to ``tests/golden/`` (inputs + reference outputs),
     which travel to the GPU box where the reference itself cannot.

Usage:  python oracle/make_golden.py [part ...]   (exit code != 0 on any mismatch; parts: operators,
        legacy_ops, pressure, trajectories, config, config2, tiny, bicgstab -- default all)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
GOLD = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, REF_SRC)
sys.path.insert(0, HERE)

import scipy.sparse.linalg as spla  # noqa: E402

import pyflow_oracle as orc  # noqa: E402
from pyflow.core.ghost_fields import allocate_state as ref_allocate, interior_view  # noqa: E402
from pyflow.linear_solvers.preconditioners import jacobi_preconditioner  # noqa: E402
from pyflow.numerics import core_ops_coherent as ref_coh  # noqa: E402
from pyflow.numerics import fluid_ops as ref_fluid  # noqa: E402
from pyflow.numerics.operators import advection as ref_adv  # noqa: E402
from pyflow.residuals.manager import ResidualManager  # noqa: E402
from pyflow.solvers import pressure_solver as ref_ps  # noqa: E402
from pyflow.solvers.solver import step as ref_step  # noqa: E402

FAILS: list[str] = []


def check(name: str, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    ok = a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    print(f"  [{'ok' if ok else 'MISMATCH'}] {name}")
    if not ok:
        FAILS.append(name)


# ---- true CG iteration counting on the unmodified reference -----------------
_TRUE_ITERS: list[int] = []
_orig_cg = spla.cg


def _counting_cg(A, b, **kw):
    n = [0]

    def cb(_xk):
        n[0] += 1
    x, info = _orig_cg(A, b, callback=cb, **kw)
    _TRUE_ITERS.append(n[0])
    return x, info


spla.cg = _counting_cg  # interface.py looks spla.cg up at call time


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def operator_vectors():
    print("operators")
    cases = {}
    for k, (nx, ny, unit) in enumerate([(8, 8, False), (16, 12, False), (5, 9, False), (11, 7, True),
                                        (33, 17, True), (70, 66, False)]):
        rng = np.random.default_rng(42 + k)
        dx, dy = (1.0, 1.0) if unit else (1.0 / (nx - 1), 1.0 / (ny - 1))
        p = rng.standard_normal((ny, nx))
        u = rng.standard_normal((ny, nx))
        v = rng.standard_normal((ny, nx))
        u[rng.random((ny, nx)) < 0.1] = 0.0   # exercise the >= / > asymmetry
        v[rng.random((ny, nx)) < 0.1] = 0.0
        f = rng.standard_normal((ny, nx))
        gx, gy = ref_coh.gradient_coherent(p, dx, dy)
        dv = ref_coh.divergence_coherent(u, v, dx, dy)
        lp = ref_coh.laplacian_coherent(p, dx, dy)
        dc = ref_fluid.divergence(u, v, dx, dy)
        dc1 = ref_fluid.divergence(u, v, 1.0, 1.0)
        au = ref_adv.advect_upwind(u, v, f, dx, dy)
        aq = ref_adv.advect_quick(u, v, f, dx, dy)
        tag = f"{nx}x{ny}"
        ogx, ogy = orc.gradient_coherent(p, dx, dy)
        check(f"grad_x {tag}", ogx, gx)
        check(f"grad_y {tag}", ogy, gy)
        check(f"div_coherent {tag}", orc.divergence_coherent(u, v, dx, dy), dv)
        check(f"lap_coherent {tag}", orc.laplacian_coherent(p, dx, dy), lp)
        check(f"div_central {tag}", orc.divergence_central(u, v, dx, dy), dc)
        check(f"div_central unit {tag}", orc.divergence_central(u, v, 1.0, 1.0), dc1)
        check(f"advect_upwind {tag}", orc.advect_upwind(u, v, f, dx, dy), au)
        check(f"advect_quick {tag}", orc.advect_quick(u, v, f, dx, dy), aq)
        # Jacobi diagonal: assembled matrix vs analytic
        A = ref_ps.assemble_negative_laplacian(nx, ny, dx, dy)
        M = jacobi_preconditioner(A)
        ones = np.ones(nx * ny)
        check(f"jacobi invD {tag}", orc.jacobi_inverse_diagonal(nx, ny, dx, dy), M.matvec(ones))
        for nm, arr in dict(p=p, u=u, v=v, f=f, gx=gx, gy=gy, div=dv, lap=lp, divc=dc, divc1=dc1,
                            upwind=au, quick=aq).items():
            cases[f"{tag}/{nm}"] = arr
        cases[f"{tag}/dxdy"] = np.array([dx, dy])
    np.savez_compressed(os.path.join(GOLD, "operators.npz"), **cases)


def legacy_ops_vectors():
    """fluid_ops.gradient / laplacian / apply_pressure_correction (numerics/fluid_ops.py:43-125): off the
    time-step path but imported by the reference's own projection tests."""
    print("legacy central operators")
    cases = {}
    for k, (nx, ny, unit) in enumerate([(8, 8, False), (16, 12, False), (5, 9, True), (2, 2, False), (3, 3, True),
                                        (2, 7, False), (9, 2, True), (33, 17, True), (70, 66, False)]):
        rng = np.random.default_rng(100 + k)
        dx, dy = (1.0, 1.0) if unit else (1.3 / max(nx - 1, 1), 0.7 / max(ny - 1, 1))
        dt = 0.0137
        p = rng.standard_normal((ny, nx))
        u = rng.standard_normal((ny, nx))
        v = rng.standard_normal((ny, nx))
        gx, gy = ref_fluid.gradient(p, dx, dy)
        lp = ref_fluid.laplacian(p, dx, dy)
        st = ref_allocate(nx, ny)
        interior_view(st.fields["u"])[:] = u
        interior_view(st.fields["v"])[:] = v
        ref_fluid.apply_pressure_correction(st, p, dt, dx, dy)
        uc, vc = interior_view(st.fields["u"]).copy(), interior_view(st.fields["v"]).copy()
        tag = f"{nx}x{ny}"
        ogx, ogy = orc.gradient_central(p, dx, dy)
        ouc, ovc = orc.pressure_correction_central(u, v, p, dt, dx, dy)
        check(f"gradient_central x {tag}", ogx, gx)
        check(f"gradient_central y {tag}", ogy, gy)
        check(f"laplacian_central {tag}", orc.laplacian_central(p, dx, dy), lp)
        check(f"pressure_correction u {tag}", ouc, uc)
        check(f"pressure_correction v {tag}", ovc, vc)
        check(f"pressure_correction p {tag}", interior_view(st.fields["p"]), p)
        for nm, arr in dict(p=p, u=u, v=v, gx=gx, gy=gy, lap=lp, ucorr=uc, vcorr=vc).items():
            cases[f"{tag}/{nm}"] = arr
        cases[f"{tag}/dxdydt"] = np.array([dx, dy, dt])
    np.savez_compressed(os.path.join(GOLD, "legacy_ops.npz"), **cases)


def pressure_vectors():
    print("pressure solve")
    out = {}
    # tests/test_preconditioner_effect.py:12-23 (24x20 bump), with and without Jacobi
    for pc in (True, False):
        nx, ny = 24, 20
        cfg = SimpleNamespace(nx=nx, ny=ny, lin_tol=1e-10, lin_maxiter=300, diagnostics=False, enable_jacobi_pc=pc)
        st = ref_allocate(nx, ny)
        interior_view(st.fields["u"])[ny // 3, nx // 3] = 1.0
        _TRUE_ITERS.clear()
        p, diag = quiet(ref_ps.solve_pressure_poisson, st, 0.01, 1.0, 1.0, cfg)
        ost = orc.allocate_state(nx, ny)
        orc.interior(ost.fields["u"])[ny // 3, nx // 3] = 1.0
        op, odiag = orc.solve_pressure_poisson(ost, 0.01, 1.0, 1.0, cfg)
        tag = f"bump_pc{int(pc)}"
        check(f"{tag} p", op, p)
        check(f"{tag} u", ost.fields["u"], st.fields["u"])
        check(f"{tag} v", ost.fields["v"], st.fields["v"])
        check(f"{tag} true iters", odiag["true_iterations"], _TRUE_ITERS[0])
        check(f"{tag} Rp_iterations", odiag["Rp_iterations"], diag["Rp_iterations"])
        check(f"{tag} Rp_residual", odiag["Rp_residual"], diag["Rp_residual"])
        check(f"{tag} divergence_norm", odiag["divergence_norm"], diag["divergence_norm"])
        out[f"{tag}/p"] = p
        out[f"{tag}/u"] = st.fields["u"].copy()
        out[f"{tag}/v"] = st.fields["v"].copy()
        out[f"{tag}/scalars"] = np.array([_TRUE_ITERS[0], diag["Rp_iterations"], diag["Rp_residual"],
                                          diag["divergence_norm"], float(diag["Rp_converged"]),
                                          diag["divergence_norm_initial"]])
    # tests/test_projection_coherent.py:11-45 (manufactured cos*cos, non-unit spacing)
    for nx, ny in [(8, 8), (12, 10)]:
        dx, dy = 1.0 / (nx - 1), 1.0 / (ny - 1)
        X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
        dt = 0.01
        us = dt * (-np.pi * np.sin(np.pi * X) * np.cos(np.pi * Y))
        vs = dt * (-np.pi * np.cos(np.pi * X) * np.sin(np.pi * Y))
        cfg = SimpleNamespace(lin_tol=1e-12, lin_maxiter=500, diagnostics=False)
        st = ref_allocate(nx, ny)
        interior_view(st.fields["u"])[:] = us
        interior_view(st.fields["v"])[:] = vs
        _TRUE_ITERS.clear()
        p, diag = quiet(ref_ps.solve_pressure_poisson, st, dt, dx, dy, cfg)
        ost = orc.allocate_state(nx, ny)
        orc.interior(ost.fields["u"])[:] = us
        orc.interior(ost.fields["v"])[:] = vs
        op, odiag = orc.solve_pressure_poisson(ost, dt, dx, dy, cfg)
        tag = f"coscos_{nx}x{ny}"
        check(f"{tag} p", op, p)
        check(f"{tag} u", ost.fields["u"], st.fields["u"])
        check(f"{tag} true iters", odiag["true_iterations"], _TRUE_ITERS[0])
        out[f"{tag}/u_star"] = us
        out[f"{tag}/v_star"] = vs
        out[f"{tag}/p"] = p
        out[f"{tag}/u"] = st.fields["u"].copy()
        out[f"{tag}/v"] = st.fields["v"].copy()
        out[f"{tag}/scalars"] = np.array([_TRUE_ITERS[0], diag["Rp_iterations"], diag["Rp_residual"],
                                          diag["divergence_norm"], float(diag["Rp_converged"]),
                                          diag["divergence_norm_initial"]])
    # tests/test_projection_method_contract.py:408-466 (manufactured Dirichlet mode, rhs_override)
    for nx, ny in [(8, 8), (14, 11)]:
        dx, dy = 1.0 / (nx - 1), 1.0 / (ny - 1)
        X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
        pa = np.sin(np.pi * X) * np.sin(np.pi * Y)
        rhs_grid = np.zeros((ny, nx))
        rhs_grid[1:-1, 1:-1] = -2.0 * np.pi ** 2 * pa[1:-1, 1:-1]
        rhs = rhs_grid.reshape(-1)
        cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=500, diagnostics=False)
        st = ref_allocate(nx, ny)
        _TRUE_ITERS.clear()
        p, diag = quiet(ref_ps.solve_pressure_poisson, st, 1.0, dx, dy, cfg, rhs_override=rhs.copy())
        ost = orc.allocate_state(nx, ny)
        op, odiag = orc.solve_manufactured(ost, dx, dy, cfg, rhs.copy())
        tag = f"manufactured_{nx}x{ny}"
        check(f"{tag} p", op, p)
        check(f"{tag} true iters", odiag["true_iterations"], _TRUE_ITERS[0])
        check(f"{tag} divergence_norm", odiag["divergence_norm"], diag["divergence_norm"])
        check(f"{tag} Rp_iterations", odiag["Rp_iterations"], diag["Rp_iterations"])
        out[f"{tag}/rhs"] = rhs_grid
        out[f"{tag}/p"] = p.copy()
        out[f"{tag}/scalars"] = np.array([_TRUE_ITERS[0], diag["Rp_iterations"], diag["Rp_residual"],
                                          diag["divergence_norm"], float(diag["Rp_converged"])])
    np.savez_compressed(os.path.join(GOLD, "pressure.npz"), **out)


TRAJECTORIES = {
    # name: (cfg kwargs, steps)
    "quick32_conv": (dict(nx=32, ny=32, Re=100.0, advection_scheme="quick", lin_maxiter=10000), 5),
    "quick128_trunc": (dict(nx=128, ny=128, Re=100.0, advection_scheme="quick", lin_maxiter=300), 10),
    "upwind96x64_trunc": (dict(nx=96, ny=64, Re=1000.0, advection_scheme="upwind", lin_maxiter=300), 6),
    "quick40x24_spacing": (dict(nx=40, ny=24, Re=400.0, advection_scheme="quick", lin_maxiter=5000,
                                lx=2.0, ly=1.0), 6),
    "upwind17x33_nopc": (dict(nx=17, ny=33, Re=50.0, advection_scheme="upwind", lin_maxiter=4000,
                              enable_jacobi_pc=False), 5),
    "quick64_default": (dict(nx=64, ny=64, Re=100.0, advection_scheme="quick", lin_maxiter=300), 4),
}


def make_cfg(kw):
    base = dict(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                disable_advection=False, diagnostics=False)
    base.update(kw)
    return SimpleNamespace(**base)


def trajectory_vectors():
    print("trajectories")
    out = {}
    meta = {}
    for name, (kw, steps) in TRAJECTORIES.items():
        cfg = make_cfg(kw)
        st = ref_allocate(cfg.nx, cfg.ny)
        tr = ResidualManager()
        _TRUE_ITERS.clear()
        rows = []
        for it in range(steps):
            _, res, diag = quiet(ref_step, cfg, st, tr, it)
            rows.append([diag["dt"], diag["CFL"], res["Ru"], res["Rv"], res["continuity"],
                         diag["Rp_iterations"], diag["Rp_residual"], float(diag["Rp_converged"]),
                         diag["divergence_norm"], float(diag["nan_detected"])])
        true_iters = list(_TRUE_ITERS)
        ost, odiags = orc.run(make_cfg(kw), steps)
        for f in ("u", "v", "p"):
            check(f"{name} field {f}", ost.fields[f], st.fields[f])
        check(f"{name} dt", [d["dt"] for d in odiags], [r[0] for r in rows])
        check(f"{name} Ru", [d["Ru"] for d in odiags], [r[2] for r in rows])
        check(f"{name} continuity", [d["continuity"] for d in odiags], [r[4] for r in rows])
        check(f"{name} Rp_residual", [d["Rp_residual"] for d in odiags], [r[6] for r in rows])
        check(f"{name} true iters", [d["true_iterations"] for d in odiags], true_iters)
        # noise floor: same oracle, pairwise dots instead of BLAS ddot
        nst, ndiags = orc.run(make_cfg(kw), steps, dot=orc.pairwise_dot)
        noise = {}
        for f in ("u", "v", "p"):
            a = orc.interior(st.fields[f])
            b = orc.interior(nst.fields[f])
            noise[f] = float(np.linalg.norm(a - b) / (np.linalg.norm(a) or 1.0))
        noise["iters_maxdiff"] = int(max(abs(a["true_iterations"] - b) for a, b in zip(ndiags, true_iters)))
        print(f"    noise floor {name}: {noise}")
        for f in ("u", "v", "p"):
            out[f"{name}/{f}"] = st.fields[f].copy()
        out[f"{name}/scalars"] = np.array(rows)
        out[f"{name}/true_iters"] = np.array(true_iters)
        meta[name] = {"cfg": kw, "steps": steps, "noise_floor": noise,
                      "columns": ["dt", "CFL", "Ru", "Rv", "continuity", "Rp_iterations", "Rp_residual",
                                  "Rp_converged", "divergence_norm", "nan_detected"]}
    np.savez_compressed(os.path.join(GOLD, "trajectories.npz"), **out)
    with open(os.path.join(GOLD, "trajectories.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)


def config2_vectors(steps=4):
    """BASELINE config 2 (512x512, Re=1000, QUICK, lin_maxiter=300) stepped by the UNMODIFIED reference
    (about 16 s per step: the lil_matrix assembly and the Python advection loops).  The fields are 2 MiB
    each, so the fixture keeps per-step scalars, norms, the centre lines and an every-16th-cell sample."""
    print(f"config 2 (512^2), {steps} reference steps")
    kw = dict(nx=512, ny=512, Re=1000.0, advection_scheme="quick", lin_maxiter=300)
    cfg = make_cfg(kw)
    st = ref_allocate(cfg.nx, cfg.ny)
    tr = ResidualManager()
    _TRUE_ITERS.clear()
    rows = []
    for it in range(steps):
        _, res, diag = quiet(ref_step, cfg, st, tr, it)
        rows.append([diag["dt"], diag["CFL"], res["Ru"], res["Rv"], res["continuity"], diag["Rp_iterations"],
                     diag["Rp_residual"], float(diag["Rp_converged"]), diag["divergence_norm"]])
        print(f"    step {it}: dt={diag['dt']:.6g} continuity={res['continuity']:.6e}", flush=True)
    ost, odiags = orc.run(make_cfg(kw), steps)
    for f in ("u", "v", "p"):
        check(f"config2 field {f}", ost.fields[f], st.fields[f])
    check("config2 dt", [d["dt"] for d in odiags], [r[0] for r in rows])
    check("config2 true iters", [d["true_iterations"] for d in odiags], list(_TRUE_ITERS))
    nst, _ = orc.run(make_cfg(kw), steps, dot=orc.pairwise_dot)
    noise = {f: float(np.linalg.norm(interior_view(st.fields[f]) - orc.interior(nst.fields[f]))
                      / (np.linalg.norm(interior_view(st.fields[f])) or 1.0)) for f in ("u", "v", "p")}
    print(f"    noise floor config2: {noise}")
    out = {"scalars": np.array(rows), "true_iters": np.array(_TRUE_ITERS)}
    for f in ("u", "v", "p"):
        a = interior_view(st.fields[f])
        out[f"{f}/norm"] = np.array([np.linalg.norm(a), np.abs(a).max(), a.sum()])
        out[f"{f}/sample16"] = a[::16, ::16].copy()
        out[f"{f}/center_col"] = a[:, cfg.nx // 2].copy()
        out[f"{f}/center_row"] = a[cfg.ny // 2, :].copy()
        out[f"{f}/top_rows"] = a[-3:, :].copy()
    np.savez_compressed(os.path.join(GOLD, "config2_512.npz"), **out)
    with open(os.path.join(GOLD, "config2_512.json"), "w") as fh:
        json.dump({"cfg": kw, "steps": steps, "noise_floor": noise,
                   "columns": ["dt", "CFL", "Ru", "Rv", "continuity", "Rp_iterations", "Rp_residual",
                               "Rp_converged", "divergence_norm"]}, fh, indent=1, sort_keys=True)


TINY = [(3, 3, "quick"), (3, 3, "upwind"), (4, 3, "quick"), (3, 5, "upwind"), (5, 4, "quick"), (2, 2, "quick"),
        (2, 5, "upwind"), (6, 2, "quick")]


def tiny_vectors(steps=6):
    """The smallest grids the reference steps (down to 2x2; SimulationConfig itself asks for >= 3): every cell is
    a wall cell or next to one, QUICK's edge padding reaches across the whole grid, CG converges in a handful of
    iterations."""
    print("tiny grids")
    out = {}
    for nx, ny, scheme in TINY:
        kw = dict(nx=nx, ny=ny, Re=100.0, advection_scheme=scheme, lin_maxiter=300)
        cfg = make_cfg(kw)
        st = ref_allocate(nx, ny)
        tr = ResidualManager()
        _TRUE_ITERS.clear()
        rows = []
        for it in range(steps):
            _, res, diag = quiet(ref_step, cfg, st, tr, it)
            rows.append([diag["dt"], diag["CFL"], res["Ru"], res["Rv"], res["continuity"], diag["Rp_iterations"],
                         diag["Rp_residual"], float(diag["Rp_converged"]), diag["divergence_norm"]])
        ost, odiags = orc.run(make_cfg(kw), steps)
        tag = f"{nx}x{ny}_{scheme}"
        for f in ("u", "v", "p"):
            check(f"tiny {tag} field {f}", ost.fields[f], st.fields[f])
            out[f"{tag}/{f}"] = st.fields[f].copy()
        check(f"tiny {tag} dt", [d["dt"] for d in odiags], [r[0] for r in rows])
        check(f"tiny {tag} true iters", [d["true_iterations"] for d in odiags], list(_TRUE_ITERS))
        out[f"{tag}/scalars"] = np.array(rows)
        out[f"{tag}/true_iters"] = np.array(_TRUE_ITERS)
    np.savez_compressed(os.path.join(GOLD, "tiny.npz"), **out)


def config_vectors():
    print("config / checkpoint")
    from pyflow.config.model import SimulationConfig
    from pyflow.io.checkpoint import load_checkpoint
    kats = {}
    for kw in (dict(nx=64, ny=64, advection_scheme="quick", lin_maxiter=300),
               dict(nx=4096, ny=4096, Re=1000.0, advection_scheme="quick", lin_maxiter=300),
               dict(nx=512, ny=512, Re=1000.0, advection_scheme="quick", lin_maxiter=300),
               dict(nx=128, ny=128, advection_scheme="quick", lin_maxiter=300),
               dict(nx=96, ny=64, Re=1000.0, advection_scheme="upwind", lin_maxiter=300, seed=7)):
        kats[json.dumps(kw, sort_keys=True)] = SimulationConfig(**kw).config_hash
    # the shipped checkpoint artefact is the layout exemplar
    st, meta = load_checkpoint("/root/reference/emergency_FAIL_NAN_it0.npz")
    with open(os.path.join(GOLD, "config_hashes.json"), "w") as fh:
        json.dump({"config_hash": kats,
                   "checkpoint_exemplar": {"meta_keys": sorted(meta.keys()),
                                           "fields": {k: list(v.shape) for k, v in st.fields.items()},
                                           "meta": meta}}, fh, indent=1, sort_keys=True)


BICG_CASES = [   # tag, nx, ny, dx, dy, jacobi, tol, maxiter
    ("24x20_plain", 24, 20, 1.0, 1.0, False, 1e-10, 2000),
    ("24x20_jacobi", 24, 20, 1.0, 1.0, True, 1e-10, 2000),
    ("33x17_jacobi_nonunit", 33, 17, 0.5, 0.25, True, 1e-10, 4000),
    ("40x28_truncated7", 40, 28, 1.0, 1.0, True, 1e-10, 7),
    ("16x16_default_tol", 16, 16, 1.0, 1.0, False, 1e-8, 500),
    ("12x10_zero_rhs", 12, 10, 1.0, 1.0, True, 1e-10, 50),
]


def bicgstab_rhs(tag, nx, ny):
    """Projection-like right-hand side: mean-free, pinned entry zero (pressure_solver.py:124-128)."""
    if "zero_rhs" in tag:
        return np.zeros(nx * ny)
    b = np.random.default_rng(nx * 100 + ny).standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    return b


def bicgstab_vectors():
    """`method='bicgstab'` -- the DEFAULT of the reference's linear_solvers.interface.solve (interface.py:23,44-45)
    -- on the pinned Neumann operator (pressure_solver.py:97-105, rebuilt here around the reference's own
    laplacian_coherent because the class is local to solve_pressure_poisson), with the reference's own Jacobi
    operator.  The oracle restatement must equal it bitwise; inputs and outputs become tests/golden/bicgstab.npz."""
    import scipy.sparse as sp
    from pyflow.linear_solvers.interface import solve as ref_solve
    print("bicgstab through the reference's interface.solve")
    out = {}
    for tag, nx, ny, dx, dy, jac, tol, maxiter in BICG_CASES:
        class Pinned(spla.LinearOperator):
            def __init__(self):
                super().__init__(dtype=np.float64, shape=(nx * ny, nx * ny))

            def _matvec(self, x):
                o = (-ref_coh.laplacian_coherent(x.reshape(ny, nx), dx, dy)).reshape(-1)
                o[0] = x[0]
                return o
        b = bicgstab_rhs(tag, nx, ny)
        kw = {}
        if jac:
            kw["M"] = jacobi_preconditioner(sp.csr_matrix(ref_ps.assemble_negative_laplacian(nx, ny, dx, dy)))
        n = [0]
        res = ref_solve(Pinned(), b.copy(), tol=tol, maxiter=maxiter, callback=lambda _x: n.__setitem__(0, n[0] + 1), **kw)
        assert res.method == "bicgstab"
        x, conv, its, rnorm, info = orc.interface_solve(b.copy(), nx, ny, dx, dy, "bicgstab", tol, maxiter, jac)
        _, _, true_its = orc.bicgstab(lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy), b.copy(), tol, maxiter,
                                      orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jac else None)
        check(f"bicgstab {tag} x", x, res.x)
        check(f"bicgstab {tag} bookkeeping", [conv, its, rnorm], [res.converged, res.iterations, res.residual_norm])
        check(f"bicgstab {tag} completed iterations", true_its, n[0])
        print(f"    {tag}: converged={res.converged} iterations={res.iterations} callbacks={n[0]} rnorm={res.residual_norm:.3e}")
        out[f"{tag}/b"] = b
        out[f"{tag}/x"] = np.asarray(res.x)
        out[f"{tag}/result"] = np.array([float(res.converged), float(res.iterations), res.residual_norm, float(n[0])])
        out[f"{tag}/params"] = np.array([nx, ny, dx, dy, float(jac), tol, maxiter])
    np.savez_compressed(os.path.join(GOLD, "bicgstab.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    parts = {"operators": operator_vectors, "legacy_ops": legacy_ops_vectors, "pressure": pressure_vectors,
             "trajectories": trajectory_vectors, "config": config_vectors, "config2": config2_vectors, "tiny": tiny_vectors,
             "bicgstab": bicgstab_vectors}
    only = sys.argv[1:]                      # e.g. `make_golden.py legacy_ops` regenerates one fixture file
    for name, fn in parts.items():
        if not only or name in only:
            fn()
    if FAILS:
        print(f"\n{len(FAILS)} MISMATCHES: {FAILS}")
        return 1
    print("\noracle == reference (bitwise) on every check; golden vectors written to tests/golden/")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
