"""CPU: the oracle against the golden vectors produced from the unmodified
reference by oracle/make_golden.py (element-wise operators bit-exact; CG-driven
quantities within the reference's own summation-order noise)."""
from types import SimpleNamespace

import numpy as np
import pytest

import pyflow_oracle as orc
import pyflow_oracle_c as orc_c          # C / OpenMP restatement (bench.py's CPU baseline); built on first use

OP_TAGS = ["8x8", "16x12", "5x9", "11x7", "33x17", "70x66"]


@pytest.mark.parametrize("tag", OP_TAGS)
def test_operators_bit_exact(golden, tag):
    g = {k.split("/", 1)[1]: v for k, v in golden["operators"].items() if k.startswith(tag + "/")}
    dx, dy = g["dxdy"]
    gx, gy = orc.gradient_coherent(g["p"], dx, dy)
    assert np.array_equal(gx, g["gx"]) and np.array_equal(gy, g["gy"])
    assert np.array_equal(orc.divergence_coherent(g["u"], g["v"], dx, dy), g["div"])
    assert np.array_equal(orc.laplacian_coherent(g["p"], dx, dy), g["lap"])
    assert np.array_equal(orc.divergence_central(g["u"], g["v"], dx, dy), g["divc"])
    assert np.array_equal(orc.divergence_central(g["u"], g["v"], 1.0, 1.0), g["divc1"])
    assert np.array_equal(orc.advect_upwind(g["u"], g["v"], g["f"], dx, dy), g["upwind"])
    assert np.array_equal(orc.advect_quick(g["u"], g["v"], g["f"], dx, dy), g["quick"])


def test_central_divergence_right_column_quirk(golden):
    g = golden["operators"]
    d = g["16x12/divc"]
    assert np.all(d[1:-1, -1] == 0.0) and np.any(d[1:-1, 0] != 0.0)


LEGACY_TAGS = ["8x8", "16x12", "5x9", "2x2", "3x3", "2x7", "9x2", "33x17", "70x66"]


@pytest.mark.parametrize("tag", LEGACY_TAGS)
def test_legacy_central_operators_bit_exact(golden, tag):
    """fluid_ops.gradient / laplacian / apply_pressure_correction (numerics/fluid_ops.py:43-125)."""
    g = {k.split("/", 1)[1]: v for k, v in golden["legacy_ops"].items() if k.startswith(tag + "/")}
    dx, dy, dt = (float(t) for t in g["dxdydt"])
    gx, gy = orc.gradient_central(g["p"], dx, dy)
    assert np.array_equal(gx, g["gx"]) and np.array_equal(gy, g["gy"])
    assert np.array_equal(orc.laplacian_central(g["p"], dx, dy), g["lap"])
    uc, vc = orc.pressure_correction_central(g["u"], g["v"], g["p"], dt, dx, dy)
    assert np.array_equal(uc, g["ucorr"]) and np.array_equal(vc, g["vcorr"])


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) or 1.0))


@pytest.mark.parametrize("pc", [True, False])
def test_pressure_bump(golden, pc):
    g = golden["pressure"]
    nx, ny = 24, 20
    cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=300, enable_jacobi_pc=pc)
    st = orc.allocate_state(nx, ny)
    orc.interior(st.fields["u"])[ny // 3, nx // 3] = 1.0
    p, diag = orc.solve_pressure_poisson(st, 0.01, 1.0, 1.0, cfg)
    tag = f"bump_pc{int(pc)}"
    assert rel(p, g[f"{tag}/p"]) < 1e-9
    assert rel(st.fields["u"], g[f"{tag}/u"]) < 1e-9
    assert abs(diag["true_iterations"] - g[f"{tag}/scalars"][0]) <= 1


@pytest.mark.parametrize("nx,ny", [(8, 8), (14, 11)])
def test_manufactured_mode(golden, nx, ny):
    g = golden["pressure"]
    tag = f"manufactured_{nx}x{ny}"
    st = orc.allocate_state(nx, ny)
    cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=500)
    p, diag = orc.solve_manufactured(st, 1.0 / (nx - 1), 1.0 / (ny - 1), cfg, g[f"{tag}/rhs"].reshape(-1).copy())
    assert rel(p, g[f"{tag}/p"]) < 1e-9 and abs(diag["true_iterations"] - g[f"{tag}/scalars"][0]) <= 1


@pytest.mark.parametrize("name", ["quick32_conv", "quick128_trunc", "upwind96x64_trunc", "quick40x24_spacing",
                                  "upwind17x33_nopc"])
def test_trajectories(golden, name):
    meta = golden["trajectories_meta"][name]
    cfg = SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          disable_advection=False, diagnostics=False, **meta["cfg"])
    st, diags = orc.run(cfg, meta["steps"])
    g = golden["trajectories"]
    floor = meta["noise_floor"]
    for f in ("u", "v", "p"):
        tol = max(1e-9, 100 * floor[f])
        assert rel(orc.interior(st.fields[f]), g[f"{name}/{f}"][1:-1, 1:-1]) < tol
    sc = g[f"{name}/scalars"]
    assert np.array_equal([d["dt"] for d in diags], sc[:, 0])     # dt is summation-order independent
    it_tol = max(1, 2 * floor["iters_maxdiff"])
    assert max(abs(a["true_iterations"] - b) for a, b in zip(diags, g[f"{name}/true_iters"])) <= it_tol


def test_survey_kats():
    """Known answers measured on the reference."""
    cfg = SimpleNamespace(nx=32, ny=32, Re=100.0, advection_scheme="quick", lin_maxiter=10000, lid_velocity=1.0,
                          cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10)
    st, diags = orc.run(cfg, 5)
    assert [d["dt"] for d in diags][:4] == [0.0105, 0.011025000000000002, 0.011576250000000003, 0.012155062500000004]
    assert diags[-1]["dt"] == 0.0063814078125000026
    assert diags[0]["true_iterations"] == 0
    assert all(abs(a - b) <= 1 for a, b in zip([d["true_iterations"] for d in diags], [0, 279, 301, 298, 291]))
    assert abs(np.linalg.norm(orc.interior(st.fields["u"])) - 5.738896948423326) < 1e-9
    assert abs(np.linalg.norm(orc.interior(st.fields["p"])) - 1.970969104493264e+01) < 1e-7
    assert diags[-1]["Rp_iterations"] == 1


# ---------------------------------------------------------------- the C / OpenMP restatement (oracle/pyflow_oracle.c)
@pytest.mark.parametrize("tag", OP_TAGS)
def test_c_oracle_operators_bit_exact(golden, tag):
    g = {k.split("/", 1)[1]: v for k, v in golden["operators"].items() if k.startswith(tag + "/")}
    dx, dy = (float(t) for t in g["dxdy"])
    gx, gy = orc_c.gradient_coherent(g["p"], dx, dy)
    assert np.array_equal(gx, g["gx"]) and np.array_equal(gy, g["gy"])
    assert np.array_equal(orc_c.divergence_coherent(g["u"], g["v"], dx, dy), g["div"])
    assert np.array_equal(orc_c.laplacian_coherent(g["p"], dx, dy), g["lap"])
    assert np.array_equal(orc_c.divergence_central(g["u"], g["v"], dx, dy), g["divc"])
    assert np.array_equal(orc_c.divergence_central(g["u"], g["v"], 1.0, 1.0), g["divc1"])
    assert np.array_equal(orc_c.advect_upwind(g["u"], g["v"], g["f"], dx, dy), g["upwind"])
    assert np.array_equal(orc_c.advect_quick(g["u"], g["v"], g["f"], dx, dy), g["quick"])


@pytest.mark.parametrize("name", ["quick32_conv", "quick128_trunc", "upwind96x64_trunc", "quick40x24_spacing",
                                  "upwind17x33_nopc"])
def test_c_oracle_trajectories(golden, name):
    meta = golden["trajectories_meta"][name]
    cfg = SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          disable_advection=False, diagnostics=False, **meta["cfg"])
    st, diags = orc_c.run(cfg, meta["steps"])
    g = golden["trajectories"]
    floor = meta["noise_floor"]
    for f in ("u", "v", "p"):
        tol = max(1e-9, 100 * floor[f])
        assert rel(orc.interior(st.fields[f]), g[f"{name}/{f}"][1:-1, 1:-1]) < tol
    assert np.array_equal([d["dt"] for d in diags], g[f"{name}/scalars"][:, 0])
    it_tol = max(1, 2 * floor["iters_maxdiff"])
    assert max(abs(a["true_iterations"] - b) for a, b in zip(diags, g[f"{name}/true_iters"])) <= it_tol


def test_c_oracle_cg_matches_numpy_oracle():
    rng = np.random.default_rng(5)
    nx, ny, dx, dy = 37, 29, 0.7, 1.3
    b = rng.standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    for pc in (True, False):
        inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if pc else None
        x0, info0, it0 = orc.pcg(lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy), b, 1e-10, 5000, inv)
        x1, info1, it1 = orc_c.pcg_pinned(b, nx, ny, dx, dy, 1e-10, 5000, jacobi=pc)
        assert info0 == info1 == 0 and abs(it0 - it1) <= 1 and rel(x1, x0) < 1e-9
    # truncated solve: same iterate after exactly 7 iterations (up to summation order)
    x0, info0, _ = orc.pcg(lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy), b, 1e-10, 7,
                           orc.jacobi_inverse_diagonal(nx, ny, dx, dy))
    x1, info1, it1 = orc_c.pcg_pinned(b, nx, ny, dx, dy, 1e-10, 7)
    assert info0 == info1 == 7 and it1 == 7 and rel(x1, x0) < 1e-12


# ---------------------------------------------------------------- BASELINE config 2 against the reference itself
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_config2_512_vs_reference_summary(golden, impl):
    """512x512, Re=1000, QUICK, 300 CG iterations per step, 4 steps from rest: both oracles against the summary
    fixture the unmodified reference wrote (tests/golden/config2_512.*)."""
    from conftest import check_config2_summary
    meta = golden["config2_512_meta"]
    cfg = SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, disable_advection=False,
                          diagnostics=False, **meta["cfg"])
    st = orc.allocate_state(cfg.nx, cfg.ny)
    ws = orc_c.Workspace(cfg.nx, cfg.ny) if impl == "c" else None
    rows, iters = [], []
    for it in range(meta["steps"]):
        _, res, d = orc_c.step(cfg, st, it, ws) if impl == "c" else orc.step(cfg, st, it)
        rows.append([d["dt"], d["CFL"], res["Ru"], res["Rv"], res["continuity"], d["Rp_iterations"], d["Rp_residual"],
                     float(d["Rp_converged"]), d["divergence_norm"]])
        iters.append(d["true_iterations"])
    check_config2_summary(golden, {f: orc.interior(st.fields[f]) for f in ("u", "v", "p")}, rows, iters)


# ---------------------------------------------------------------- the smallest grids the reference steps
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_tiny_grids_vs_reference(golden, impl):
    from conftest import TINY_TAGS, tiny_cfg
    for tag in TINY_TAGS:
        cfg = tiny_cfg(tag)
        g = {k.split("/", 1)[1]: v for k, v in golden["tiny"].items() if k.startswith(tag + "/")}
        st = orc.allocate_state(cfg.nx, cfg.ny)
        ws = orc_c.Workspace(cfg.nx, cfg.ny) if impl == "c" else None
        for it in range(len(g["true_iters"])):
            _, res, d = orc_c.step(cfg, st, it, ws) if impl == "c" else orc.step(cfg, st, it)
            assert d["dt"] == g["scalars"][it, 0] and d["CFL"] == g["scalars"][it, 1], (tag, it)
            assert abs(d["true_iterations"] - g["true_iters"][it]) <= (0 if impl == "numpy" else 1), (tag, it)
        for f in ("u", "v", "p"):
            if impl == "numpy":
                assert np.array_equal(st.fields[f], g[f]), (tag, f)
            else:
                assert rel(st.fields[f], g[f]) < 1e-9, (tag, f)


# ---------------------------------------------------------------- method='bicgstab' (default of interface.solve)
BICG_TAGS = ["24x20_plain", "24x20_jacobi", "33x17_jacobi_nonunit", "40x28_truncated7", "16x16_default_tol",
             "12x10_zero_rhs"]


@pytest.mark.parametrize("tag", BICG_TAGS)
def test_bicgstab_oracle_vs_reference_golden(tag):
    """oracle.interface_solve(method='bicgstab') against what the UNMODIFIED reference's interface.solve returned
    (tests/golden/bicgstab.npz, oracle/make_golden.py bicgstab): same machine arithmetic -> bitwise."""
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, "bicgstab.npz")) as z:
        g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/")}
    nx, ny, dx, dy, jac, tol, maxiter = g["params"]
    x, conv, its, rnorm, _info = orc.interface_solve(g["b"].copy(), int(nx), int(ny), dx, dy, "bicgstab", tol,
                                                     int(maxiter), bool(jac))
    assert np.array_equal(x, g["x"])
    assert [float(conv), float(its), rnorm] == list(g["result"][:3])


def test_bicgstab_oracle_is_scipys_recurrence():
    """The restatement against scipy.sparse.linalg.bicgstab itself (present on every box), bit for bit, incl. the
    preconditioned and the truncated case."""
    import scipy.sparse.linalg as spla
    nx, ny, dx, dy = 21, 14, 0.7, 1.3
    b = np.random.default_rng(5).standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    mv = lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy)                      # noqa: E731
    A = spla.LinearOperator((nx * ny, nx * ny), matvec=mv, dtype=np.float64)
    invd = orc.jacobi_inverse_diagonal(nx, ny, dx, dy)
    M = spla.LinearOperator((nx * ny, nx * ny), matvec=lambda t: invd * t, dtype=np.float64)
    for kw, inv, maxiter in (({}, None, 3000), ({"M": M}, invd, 3000), ({"M": M}, invd, 5)):
        xs, infos = spla.bicgstab(A, b, x0=np.zeros_like(b), rtol=1e-10, maxiter=maxiter, **kw)
        xo, infoo, _ = orc.bicgstab(mv, b.copy(), 1e-10, maxiter, inv)
        assert infos == infoo and np.array_equal(xs, xo)
