"""The two fused per-step passes (csrc/pf_step_fused.cuh: k_front = predictor + rhs pass 1, k_finish = everything
after the solve) against the seven separate kernels of csrc/pf_kernels.cuh ON THE GPU.

One GPU steps with the fused passes by default, so every other single-GPU test of this suite already holds them
against the golden vectors of the reference and against the oracle.  Here:

* the SEPARATE kernels -- what slab runs on several GPUs step with -- stay covered on a one-GPU box
  (``step_fused=0``) against the same golden trajectories;
* both paths on the same cavities: identical dt / CFL sequences, identical iteration counts, fields equal to the
  last few bits (every field value is formed by the same expression in both; only the summation order of the rhs
  mean and of the dot products differs), five kernel launches fewer per step.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest

import pyflow_oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb else 1.0))


def _cfg(**kw):
    return SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, disable_advection=False,
                           diagnostics=False, **kw)


def _run(cfg, steps, options, monkeypatch):
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    if options:
        monkeypatch.setenv("PYFLOW_B200_OPTIONS", options)
    else:
        monkeypatch.delenv("PYFLOW_B200_OPTIONS", raising=False)
    st = allocate_state(cfg.nx, cfg.ny)
    tr = ResidualManager()
    rows, launches = [], []
    for it in range(steps):
        n0 = st.fields._slab.launch_count() if getattr(st.fields, "_slab", None) is not None else None
        _, res, diag = step(cfg, st, tr, it)
        n1 = st.fields._slab.launch_count()
        launches.append(None if n0 is None else n1 - n0)
        rows.append((diag["dt"], diag["CFL"], st.meta["cg_true_iterations"], res["Ru"], res["Rv"], res["continuity"],
                     diag["divergence_norm"], diag["Rp_residual"]))
    fields = {f: st.fields[f][1:-1, 1:-1].copy() for f in ("u", "v", "p")}
    st.fields._slab.close()
    return rows, fields, launches


@pytest.mark.parametrize("name", ["quick32_conv", "quick128_trunc", "upwind96x64_trunc", "quick40x24_spacing",
                                  "upwind17x33_nopc"])
def test_separate_step_kernels_vs_golden(golden, monkeypatch, name):
    """step_fused=0: the per-step kernels of the multi-GPU path, on one GPU, against the reference's trajectories."""
    meta = golden["trajectories_meta"][name]
    g = golden["trajectories"]
    cfg = _cfg(**meta["cfg"])
    rows, fields, _l = _run(cfg, meta["steps"], "step_fused=0", monkeypatch)
    sc = g[f"{name}/scalars"]
    assert [r[0] for r in rows] == list(sc[:, 0]), "dt sequence must be bit-identical"
    assert [r[1] for r in rows] == list(sc[:, 1]), "CFL sequence must be bit-identical"
    for f in ("u", "v", "p"):
        tol = max(1e-9, 100 * meta["noise_floor"][f])
        assert rel(fields[f], g[f"{name}/{f}"][1:-1, 1:-1]) < tol, f


@pytest.mark.parametrize("nx,ny,scheme,spacing,maxit,steps", [
    (200, 150, "quick", None, 300, 5),            # k_cg_small; four warps per row, bands of 2 rows
    (59, 47, "upwind", (3.0, 2.0), 30, 5),        # odd width just past one warp's 58 columns, general spacing
                                                  # (30 iterations: a solve truncated MIDWAY, e.g. 400 here, amplifies the
                                                  # last-bit difference of the rhs mean to 1e-2 -- between the numpy and
                                                  # the C oracle just the same, SURVEY 8c on config 1)
    (1100, 130, "quick", None, 120, 4),           # k_cg_tmad with a clipped chunk; three column blocks
    (1537, 300, "quick", None, 60, 3),            # odd width, four column blocks
    (2048, 2048, "quick", None, 40, 3),           # 14-row bands, 5 column blocks
])
def test_fused_and_separate_step_kernels_agree(monkeypatch, nx, ny, scheme, spacing, maxit, steps):
    cfg = _cfg(nx=nx, ny=ny, Re=1000.0, advection_scheme=scheme, lin_maxiter=maxit)
    if spacing:
        cfg.lx, cfg.ly = spacing
    rf, ff, lf = _run(cfg, steps, None, monkeypatch)
    rs, fs, ls = _run(cfg, steps, "step_fused=0", monkeypatch)
    assert [r[0] for r in rf] == [r[0] for r in rs], "dt"
    assert [r[1] for r in rf] == [r[1] for r in rs], "CFL"
    assert [r[2] for r in rf] == [r[2] for r in rs], "true CG iterations"
    for k in range(3, 8):
        np.testing.assert_allclose([r[k] for r in rf], [r[k] for r in rs], rtol=1e-9, atol=1e-300)
    for f in ("u", "v", "p"):
        assert rel(ff[f], fs[f]) < 1e-10, (f, rel(ff[f], fs[f]))
    # steps >= 1 (step 0 creates the context): predictor + rhs1 -> one launch, five kernels after the solve -> one
    if os.environ.get("PYFLOW_GPU_TESTS_ON_STAND_IN") == "1":
        return                                    # the CPU stand-in of the developer dry run counts no kernels
    assert all(a is not None and b is not None and b - a == 5 for a, b in zip(lf[1:], ls[1:])), (lf, ls)


def test_fused_passes_vs_live_oracle_with_poisoned_velocity(monkeypatch):
    """A NaN seeded into u before the run: both paths report it (nan_detected, CG runs to maxiter) and agree with the
    oracle on which cells are poisoned after two steps."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = _cfg(nx=130, ny=70, Re=100.0, advection_scheme="quick", lin_maxiter=25)
    masks = []
    for options in (None, "step_fused=0"):
        if options:
            monkeypatch.setenv("PYFLOW_B200_OPTIONS", options)
        else:
            monkeypatch.delenv("PYFLOW_B200_OPTIONS", raising=False)
        st = allocate_state(cfg.nx, cfg.ny)
        st.fields["u"][20, 30] = np.nan
        tr = ResidualManager()
        flags = []
        with np.errstate(invalid="ignore"):
            for it in range(2):
                _, _res, diag = step(cfg, st, tr, it)
                flags.append(bool(diag["nan_detected"]))
        assert flags == [True, True]
        masks.append(np.isnan(st.fields["u"][1:-1, 1:-1]))
        st.fields._slab.close()
    ost = orc.allocate_state(cfg.nx, cfg.ny)
    ost.fields["u"][20, 30] = np.nan
    with np.errstate(invalid="ignore"):
        for it in range(2):
            orc.step(cfg, ost, it)
    want = np.isnan(orc.interior(ost.fields["u"]))
    assert np.array_equal(masks[0], masks[1]) and np.array_equal(masks[0], want)
