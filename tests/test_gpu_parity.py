"""GPU parity: the CUDA path (through the C-ABI / drop-in Python surface)
against the oracle and the golden vectors of the unmodified reference.

Bars (written here as the north star asks):
  * element-wise operators: BIT-EXACT (np.array_equal);
  * anything downstream of a CG dot product: relative L2 <= max(1e-9, 100 x the
    reference's own summation-order noise floor recorded in
    tests/golden/trajectories.json), true CG iterations within
    max(1, 2 x the reference's own iteration noise).
"""
from types import SimpleNamespace

import numpy as np
import pytest

import pyflow_oracle as orc

pytestmark = pytest.mark.gpu

OP_TAGS = ["8x8", "16x12", "5x9", "11x7", "33x17", "70x66"]


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) or 1.0))


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


# ---------------------------------------------------------------- operators
@pytest.mark.parametrize("tag", OP_TAGS)
def test_operators_bit_exact_vs_golden(golden, tag):
    from pyflow.numerics.core_ops_coherent import divergence_coherent, gradient_coherent, laplacian_coherent
    from pyflow.numerics.fluid_ops import divergence
    from pyflow.numerics.operators.advection import advect_quick, advect_upwind
    g = {k.split("/", 1)[1]: v for k, v in golden["operators"].items() if k.startswith(tag + "/")}
    dx, dy = g["dxdy"]
    gx, gy = gradient_coherent(g["p"], dx, dy)
    assert bits_equal(gx, g["gx"]) and bits_equal(gy, g["gy"])
    assert bits_equal(divergence_coherent(g["u"], g["v"], dx, dy), g["div"])
    assert bits_equal(laplacian_coherent(g["p"], dx, dy), g["lap"])
    assert bits_equal(divergence(g["u"], g["v"], dx, dy), g["divc"])
    assert bits_equal(divergence(g["u"], g["v"], 1.0, 1.0), g["divc1"])
    assert bits_equal(advect_upwind(g["u"], g["v"], g["f"], dx, dy), g["upwind"])
    assert bits_equal(advect_quick(g["u"], g["v"], g["f"], dx, dy), g["quick"])


@pytest.mark.parametrize("shape,unit", [((257, 513), True), ((300, 1030), False), ((64, 2), True), ((2, 64), False),
                                        ((1024, 1024), True)])
def test_operators_bit_exact_vs_oracle(shape, unit):
    from pyflow.numerics.core_ops_coherent import divergence_coherent, gradient_coherent, laplacian_coherent
    from pyflow.numerics.fluid_ops import divergence
    from pyflow.numerics.operators.advection import advect_quick, advect_upwind
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    p, u, v, f = (rng.standard_normal(shape) for _ in range(4))
    u[rng.random(shape) < 0.05] = 0.0
    gx, gy = gradient_coherent(p, dx, dy)
    ogx, ogy = orc.gradient_coherent(p, dx, dy)
    assert bits_equal(gx, ogx) and bits_equal(gy, ogy)
    assert bits_equal(divergence_coherent(u, v, dx, dy), orc.divergence_coherent(u, v, dx, dy))
    assert bits_equal(laplacian_coherent(p, dx, dy), orc.laplacian_coherent(p, dx, dy))
    assert bits_equal(divergence(u, v, dx, dy), orc.divergence_central(u, v, dx, dy))
    assert bits_equal(advect_upwind(u, v, f, dx, dy), orc.advect_upwind(u, v, f, dx, dy))
    assert bits_equal(advect_quick(u, v, f, dx, dy), orc.advect_quick(u, v, f, dx, dy))


def test_operator_nan_inf_propagation():
    from pyflow.numerics.core_ops_coherent import laplacian_coherent
    from pyflow.numerics.operators.advection import advect_quick
    rng = np.random.default_rng(5)
    p = rng.standard_normal((20, 24))
    p[3, 4] = np.nan
    p[10, 0] = np.inf
    with np.errstate(invalid="ignore"):
        assert bits_equal(laplacian_coherent(p, 1.0, 1.0), orc.laplacian_coherent(p, 1.0, 1.0))
        u = rng.standard_normal((20, 24))
        u[5, 5] = np.nan
        assert bits_equal(advect_quick(u, p, p, 1.0, 1.0), orc.advect_quick(u, p, p, 1.0, 1.0))


def test_degenerate_shapes_return_zeros():
    from pyflow.numerics.core_ops_coherent import gradient_coherent, laplacian_coherent
    p = np.ones((1, 7))
    assert np.all(laplacian_coherent(p, 1.0, 1.0) == 0.0)
    gx, gy = gradient_coherent(p, 1.0, 1.0)
    assert gx.shape == p.shape and not gx.any() and not gy.any()


# ---------------------------------------------------------------- CG on the pinned operator
@pytest.mark.parametrize("ny,nx,unit,jac", [(20, 24, True, True), (33, 17, False, True), (64, 64, True, False),
                                            (130, 70, True, True)])
def test_pcg_matches_scipy_recurrence(ny, nx, unit, jac):
    from pyflow.linear_solvers.interface import NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    rng = np.random.default_rng(nx + ny)
    dx, dy = (1.0, 1.0) if unit else (1.0 / (nx - 1), 1.0 / (ny - 1))
    b = rng.standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    A = NegLaplacianPinned(nx, ny, dx, dy)
    kw = {"M": jacobi_preconditioner(A)} if jac else {}
    res = solve(A, b, method="cg", tol=1e-10, maxiter=5000, **kw)
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jac else None
    xo, info, iters = orc.pcg(lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy), b, 1e-10, 5000, inv)
    assert info == 0 and res.converged
    assert rel(res.x, xo) < 1e-8
    # operator seam: A @ x on the device equals the oracle's matvec bit for bit
    assert bits_equal(A.matvec(xo), orc.neg_laplacian_pinned(xo, nx, ny, dx, dy))


@pytest.mark.parametrize("ny,nx", [(70, 65), (257, 513), (40, 1100), (33, 4096), (48, 577), (21, 129),
                                   (36, 2049), (50, 1030), (24, 1089), (200, 3000), (19, 1025)])
def test_pcg_strip_and_chunk_layouts_vs_oracle(ny, nx):
    """Column layouts of the tuned kernel: one column beyond a strip (nx % 64 == 1, the
    neighbour strip's halo column is then the wall), several column chunks, partial chunks."""
    from pyflow._b200.device import slab_for_shape
    slab = slab_for_shape(ny, nx)
    rng = np.random.default_rng(ny * 7 + nx)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    its = 25
    x, info, iters, rnorm = slab.pcg(b, 1.0, 1.0, 1e-10, its, jacobi=True)
    inv = orc.jacobi_inverse_diagonal(nx, ny, 1.0, 1.0)
    xo, oinfo, oiters = orc.pcg(lambda z: orc.neg_laplacian_pinned(z, nx, ny, 1.0, 1.0), b.ravel().copy(), 1e-10, its, inv)
    assert info == oinfo == its and iters == oiters
    assert rel(x.ravel(), xo) < 1e-12, rel(x.ravel(), xo)


def test_pcg_truncated_and_zero_rhs():
    from pyflow._b200.device import slab_for_shape
    ny, nx = 96, 80
    slab = slab_for_shape(ny, nx)
    rng = np.random.default_rng(1)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    x, info, iters, rnorm = slab.pcg(b, 1.0, 1.0, 1e-10, 25, jacobi=True)
    assert info == 25 and iters == 25
    inv = orc.jacobi_inverse_diagonal(nx, ny, 1.0, 1.0)
    xo, oinfo, oiters = orc.pcg(lambda z: orc.neg_laplacian_pinned(z, nx, ny, 1.0, 1.0), b.ravel(), 1e-10, 25, inv)
    assert oinfo == 25 and rel(x.ravel(), xo) < 1e-12
    ro = b.ravel() - orc.neg_laplacian_pinned(xo, nx, ny, 1.0, 1.0)
    assert abs(rnorm - np.linalg.norm(ro)) <= 1e-10 * np.linalg.norm(ro)
    x, info, iters, rnorm = slab.pcg(np.zeros((ny, nx)), 1.0, 1.0, 1e-10, 25, jacobi=True)
    assert info == 0 and iters == 0 and not x.any() and rnorm == 0.0


# ---------------------------------------------------------------- pressure projection seam
@pytest.mark.parametrize("pc", [True, False])
def test_projection_bump_vs_golden(golden, pc):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.solvers.pressure_solver import solve_pressure_poisson
    g = golden["pressure"]
    nx, ny = 24, 20
    cfg = SimpleNamespace(nx=nx, ny=ny, lin_tol=1e-10, lin_maxiter=300, diagnostics=False, enable_jacobi_pc=pc)
    st = allocate_state(nx, ny)
    interior_view(st.fields["u"])[ny // 3, nx // 3] = 1.0
    p, diag = solve_pressure_poisson(st, 0.01, 1.0, 1.0, cfg)
    tag = f"bump_pc{int(pc)}"
    sc = g[f"{tag}/scalars"]   # true_iters, Rp_iterations, Rp_residual, divergence_norm, converged, div_initial
    assert rel(p, g[f"{tag}/p"]) < 1e-9
    assert rel(st.fields["u"], g[f"{tag}/u"]) < 1e-9 and rel(st.fields["v"], g[f"{tag}/v"]) < 1e-9
    assert abs(st.meta["cg_true_iterations"] - sc[0]) <= 1
    assert diag["Rp_iterations"] == int(sc[1]) and diag["Rp_converged"] == bool(sc[4])
    assert abs(diag["divergence_norm_initial"] - sc[5]) <= 1e-12 * sc[5]
    assert diag["divergence_norm"] < 1e-6 * sc[5]


@pytest.mark.parametrize("nx,ny", [(8, 8), (12, 10)])
def test_projection_coscos_vs_golden(golden, nx, ny):
    """tests/test_projection_coherent.py of the reference, non-unit spacing."""
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.solvers.pressure_solver import solve_pressure_poisson
    g = golden["pressure"]
    tag = f"coscos_{nx}x{ny}"
    dx, dy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    cfg = SimpleNamespace(lin_tol=1e-12, lin_maxiter=500, diagnostics=False)
    st = allocate_state(nx, ny)
    interior_view(st.fields["u"])[:] = g[f"{tag}/u_star"]
    interior_view(st.fields["v"])[:] = g[f"{tag}/v_star"]
    p, diag = solve_pressure_poisson(st, 0.01, dx, dy, cfg)
    assert rel(p, g[f"{tag}/p"]) < 1e-9
    assert rel(st.fields["u"], g[f"{tag}/u"]) < 1e-8
    assert diag["divergence_norm"] < max(1e-10, diag["divergence_norm_initial"] * 5e-6)


@pytest.mark.parametrize("nx,ny", [(8, 8), (14, 11)])
def test_manufactured_dirichlet_mode_vs_golden(golden, nx, ny):
    """rhs_override mode (reference pressure_solver.py:61-93; tests/test_projection_method_contract.py:408-466)."""
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.solvers.pressure_solver import solve_pressure_poisson
    g = golden["pressure"]
    tag = f"manufactured_{nx}x{ny}"
    dx, dy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=500, diagnostics=False)
    st = allocate_state(nx, ny)
    interior_view(st.fields["u"])[:] = 0.25          # velocities must come back untouched
    p, diag = solve_pressure_poisson(st, 1.0, dx, dy, cfg, rhs_override=g[f"{tag}/rhs"].reshape(-1).copy())
    sc = g[f"{tag}/scalars"]            # true_iters, Rp_iterations, Rp_residual, divergence_norm, converged
    assert rel(p, g[f"{tag}/p"]) < 1e-9
    assert np.array_equal(interior_view(st.fields["p"]), p)
    assert np.all(interior_view(st.fields["u"]) == 0.25) and not st.fields["v"].any()
    assert abs(st.meta["cg_true_iterations"] - sc[0]) <= 1
    assert diag["Rp_iterations"] == int(sc[1]) and diag["Rp_converged"] == bool(sc[4])
    assert diag["divergence_norm"] < 1e-8 and diag["divergence_norm_initial"] is None
    X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    assert np.allclose(p[1:-1, 1:-1], (np.sin(np.pi * X) * np.sin(np.pi * Y))[1:-1, 1:-1], atol=5e-2)


@pytest.mark.parametrize("ny,nx,unit", [(96, 130, True), (50, 64, False), (257, 513, True)])
def test_tuned_cg_kernel_agrees_with_plain_cg_kernel(ny, nx, unit):
    """Two independent device implementations of the same recurrence (register-window
    persistent kernel vs one-thread-per-cell kernel with q stored)."""
    from pyflow._b200.device import slab_for_shape
    slab = slab_for_shape(ny, nx)
    rng = np.random.default_rng(ny + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
    x1, info1, it1, rn1 = slab.pcg(b, dx, dy, 1e-10, 60, jacobi=True, inv_diag_interior=invd)
    x2, info2, it2, rn2, bn2 = slab.cg_generic("pinned", b, dx, dy, 1e-10, 60, jacobi=True, inv_diag_interior=invd)
    assert info1 == info2 == 60 and it1 == it2 == 60
    assert rel(x1, x2) < 1e-10, rel(x1, x2)      # summation order of the dot products differs between the two
    assert abs(rn1 - rn2) <= 1e-9 * rn2, (rn1, rn2)
    assert abs(bn2 - np.linalg.norm(b)) <= 1e-12 * bn2


# ---------------------------------------------------------------- full trajectories
def _cfg(kw):
    return SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                           disable_advection=False, diagnostics=False, **kw)


@pytest.mark.parametrize("name", ["quick32_conv", "quick128_trunc", "upwind96x64_trunc", "quick40x24_spacing",
                                  "upwind17x33_nopc"])
def test_trajectory_vs_golden(golden, name):
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    meta = golden["trajectories_meta"][name]
    g = golden["trajectories"]
    cfg = _cfg(meta["cfg"])
    st = allocate_state(cfg.nx, cfg.ny)
    tr = ResidualManager()
    rows, iters = [], []
    for it in range(meta["steps"]):
        _, res, diag = step(cfg, st, tr, it)
        rows.append([diag["dt"], diag["CFL"], res["Ru"], res["Rv"], res["continuity"], diag["Rp_iterations"],
                     diag["Rp_residual"], float(diag["Rp_converged"]), diag["divergence_norm"],
                     float(diag["nan_detected"])])
        iters.append(st.meta["cg_true_iterations"])
    rows = np.array(rows)
    sc = g[f"{name}/scalars"]
    floor = meta["noise_floor"]
    errs = {}
    for f in ("u", "v", "p"):
        tol = max(1e-9, 100 * floor[f])
        errs[f] = rel(st.fields[f][1:-1, 1:-1], g[f"{name}/{f}"][1:-1, 1:-1])
        assert errs[f] < tol, (f, errs, floor)
    print(f"{name}: relL2 {errs}  reference self-noise {floor}")
    assert np.array_equal(rows[:, 0], sc[:, 0]), "dt sequence must be bit-identical"
    assert np.array_equal(rows[:, 1], sc[:, 1]), "CFL sequence must be bit-identical"
    it_tol = max(1, 2 * floor["iters_maxdiff"])
    assert np.max(np.abs(np.array(iters) - g[f"{name}/true_iters"])) <= it_tol
    assert np.array_equal(rows[:, 5], sc[:, 5])          # reported Rp_iterations (0/1/maxiter)
    np.testing.assert_allclose(rows[:, 2:5], sc[:, 2:5], rtol=1e-7)
    assert not rows[:, 9].any()


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("name", ["quick128_trunc", "quick40x24_spacing", "upwind17x33_nopc"])
def test_trajectory_with_forced_cg_kernel(golden, monkeypatch, name, kernel):
    """Small grids default to the shared-memory-resident kernel; the register-window kernel
    (cg_kernel=1) and the static TMA path selector (cg_kernel=2, falls back to register-window for
    these widths) must give the same physics."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    monkeypatch.setenv("PYFLOW_B200_OPTIONS", f"cg_kernel={kernel}")
    meta = golden["trajectories_meta"][name]
    g = golden["trajectories"]
    cfg = _cfg(meta["cfg"])
    st = allocate_state(cfg.nx, cfg.ny)
    tr = ResidualManager()
    for it in range(meta["steps"]):
        step(cfg, st, tr, it)
    for f in ("u", "v", "p"):
        tol = max(1e-9, 100 * meta["noise_floor"][f])
        assert rel(st.fields[f][1:-1, 1:-1], g[f"{name}/{f}"][1:-1, 1:-1]) < tol


def test_default_config_trajectory_within_reference_noise(golden):
    """Config 1 (64^2, 300 its) is truncated midway: even reference-vs-reference
    differs by 1e-5..1e-3; the GPU must stay inside that band."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    name = "quick64_default"
    meta = golden["trajectories_meta"][name]
    cfg = _cfg(meta["cfg"])
    st = allocate_state(cfg.nx, cfg.ny)
    tr = ResidualManager()
    for it in range(meta["steps"]):
        step(cfg, st, tr, it)
        assert st.meta["cg_true_iterations"] == golden["trajectories"][f"{name}/true_iters"][it]
    for f in ("u", "v", "p"):
        assert rel(st.fields[f], golden["trajectories"][f"{name}/{f}"]) < 100 * meta["noise_floor"][f]


def test_trajectory_tma_kernel_vs_live_oracle():
    """A 1024-wide grid takes the TMA-fed, dynamically scheduled CG kernel (nx % 1024 == 0):
    whole steps against the oracle, QUICK, truncated solves."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = _cfg(dict(nx=1024, ny=96, Re=1000.0, advection_scheme="quick", lin_maxiter=150))
    steps = 4
    ost, odiags = orc.run(cfg, steps)
    st = allocate_state(1024, 96)
    tr = ResidualManager()
    for it in range(steps):
        _, res, diag = step(cfg, st, tr, it)
        assert diag["dt"] == odiags[it]["dt"]
        assert st.meta["cg_true_iterations"] == odiags[it]["true_iterations"]
    for f in ("u", "v", "p"):
        e = rel(st.fields[f][1:-1, 1:-1], orc.interior(ost.fields[f]))
        assert e < 1e-9, (f, e)


def test_trajectory_vs_live_oracle_512():
    """Config 2 grid (512^2, Re=1000, quick, 300 its), first steps, against the
    oracle run on the host in this very test."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = _cfg(dict(nx=512, ny=512, Re=1000.0, advection_scheme="quick", lin_maxiter=300))
    steps = 3
    ost, odiags = orc.run(cfg, steps)
    st = allocate_state(512, 512)
    tr = ResidualManager()
    for it in range(steps):
        _, res, diag = step(cfg, st, tr, it)
        assert diag["dt"] == odiags[it]["dt"]
        assert st.meta["cg_true_iterations"] == odiags[it]["true_iterations"]
    for f in ("u", "v", "p"):
        e = rel(st.fields[f][1:-1, 1:-1], orc.interior(ost.fields[f]))
        assert e < 1e-9, (f, e)


# ---------------------------------------------------------------- driver-level behaviour
def test_driver_yields_same_state_and_centerline():
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.drivers.simulation_driver import SimulationDriver
    cfg = _cfg(dict(nx=48, ny=40, Re=100.0, advection_scheme="quick", lin_maxiter=200))
    state = allocate_state(48, 40)
    drv = SimulationDriver(cfg, state)
    n = 0
    for st, res, diag in drv.run(max_steps=4):
        assert st is state
        ys, uc = diag["u_centerline"]
        assert len(ys) == 40 and len(uc) == 40
        assert np.array_equal(np.array(uc), interior_view(st.fields["u"])[:, 24])
        assert ys[-1] == 39.0
        n += 1
    assert n == 4
    assert np.allclose(state.fields["u"][-2, 1:-1], 1.0)      # lid row enforced after the projection


def test_host_mutation_between_steps_is_uploaded():
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = _cfg(dict(nx=24, ny=24, Re=100.0, advection_scheme="upwind", lin_maxiter=2000))   # converged regime
    a, b = allocate_state(24, 24), orc.allocate_state(24, 24)
    tr = ResidualManager()
    step(cfg, a, tr, 0)
    orc.step(cfg, b, 0)
    interior_view(a.fields["v"])[5:9, 5:9] = 0.25
    orc.interior(b.fields["v"])[5:9, 5:9] = 0.25
    step(cfg, a, tr, 1)
    orc.step(cfg, b, 1)
    for f in ("u", "v", "p"):
        assert rel(a.fields[f], b.fields[f]) < 1e-9


def test_foreign_state_object_is_synchronised_eagerly():
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = _cfg(dict(nx=20, ny=16, Re=100.0, advection_scheme="quick", lin_maxiter=2000))   # converged regime
    foreign = orc.allocate_state(20, 16)          # plain dataclass with dict fields
    ref = orc.allocate_state(20, 16)
    held = foreign.fields["u"]
    tr = ResidualManager()
    for it in range(3):
        step(cfg, foreign, tr, it)
        orc.step(cfg, ref, it)
    assert held is foreign.fields["u"] and rel(held, ref.fields["u"]) < 1e-9


def test_nan_is_reported_not_raised():
    import io
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.drivers.simulation_driver import SimulationDriver
    buf = io.StringIO()
    cfg = SimpleNamespace(disable_advection=True, advection_scheme="upwind", cfl_target=0.5, cfl_growth=1.05,
                          Re=100.0, lid_velocity=0.0, lin_tol=1e-10, lin_maxiter=30, nx=10, ny=10, lx=9, ly=9,
                          diagnostics=False, log_stream=buf)
    state = allocate_state(10, 10)
    interior_view(state.fields["u"])[2, 2] = np.nan
    import os
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            for _st, _res, diag in SimulationDriver(cfg, state).run(max_steps=1):
                assert diag["nan_detected"] is True
                assert diag["Rp_iterations"] == 30          # NaN never satisfies the stop test
        finally:
            os.chdir(cwd)
    assert any('"reason":"nan_detected"' in line for line in buf.getvalue().splitlines())


# ---------------------------------------------------------------- size-independent properties at scale
def test_large_grid_properties_4096():
    """BASELINE config 3 grid: properties that need no CPU reference.
    (1) fused Laplacian == div(grad) bit for bit; (2) constants are in the null
    space; (3) one projection of a random field drives the coherent divergence
    down and leaves p[0,0] == 0; (4) the step is deterministic run to run."""
    from pyflow._b200.device import DeviceSlab
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    n = 4096
    slab = DeviceSlab(n, n)
    rng = np.random.default_rng(0)
    p = rng.standard_normal((n, n))
    lap = slab.laplacian_coherent(p, 1.0, 1.0)
    gx, gy = slab.gradient_coherent(p, 1.0, 1.0)
    assert np.array_equal(lap, slab.divergence_coherent(gx, gy, 1.0, 1.0))
    assert not slab.laplacian_coherent(np.full((n, n), 3.25), 1.0, 1.0).any()
    assert abs(lap.sum()) < 1e-6 * np.abs(lap).sum()
    slab.close()
    cfg = _cfg(dict(nx=n, ny=n, Re=1000.0, advection_scheme="quick", lin_maxiter=300))
    finals = []
    for _ in range(2):
        st = allocate_state(n, n)
        tr = ResidualManager()
        for it in range(3):
            _, res, diag = step(cfg, st, tr, it)
        assert st.meta["cg_true_iterations"] == 300 and not diag["nan_detected"]
        finals.append(interior_view(st.fields["p"]).copy())
        assert finals[-1][0, 0] == 0.0
        assert np.allclose(st.fields["u"][-2, 1:-1], 1.0)
    assert np.array_equal(finals[0], finals[1])


def test_alternate_driver_divergence_stats_on_device():
    """driver/simulation_driver.py (singular): mean / mean-free norm of the central divergence come from
    device reductions; compare with numpy on the downloaded fields and with the oracle trajectory."""
    from pyflow.driver.simulation_driver import DriverConfig, PhysicsConfig, SimulationDriver
    phys = PhysicsConfig(lin_maxiter=3000, lx=32.0, ly=20.0)          # non-unit spacing: dx = 32/39, dy = 20/23
    drv = SimulationDriver(DriverConfig(nx=40, ny=24, max_steps=4, report_interval=1000), phys)
    payloads = list(drv.run())
    assert len(payloads) == 4 and payloads[-1]["state"] is drv.state
    ost, od = orc.run(SimpleNamespace(nx=40, ny=24, **phys.__dict__), 4)
    st = drv.state
    assert rel(st.fields["u"][1:-1, 1:-1], orc.interior(ost.fields["u"])) < 1e-9
    d = orc.divergence_central(orc.interior(ost.fields["u"]), orc.interior(ost.fields["v"]), st.meta["dx"], st.meta["dy"])
    mean, mf = float(d.mean()), float(np.linalg.norm(d - d.mean()))
    assert abs(payloads[-1]["divergence_mean"] - mean) <= 1e-9 * max(abs(mean), 1e-12) + 1e-15
    assert abs(payloads[-1]["divergence_mean_free_norm"] - mf) <= 1e-9 * mf
    assert payloads[-1]["diagnostics"]["dt"] == od[-1]["dt"]


# ---------------------------------------------------------------- CLI / checkpoint / restart on the device path
def _run_cli(*argv, cwd=None):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "pyflow-solver_b200"))
    return subprocess.run([sys.executable, "-m", "pyflow", *argv], capture_output=True, text=True, env=env, cwd=cwd)


def test_cli_json_stream_checkpoint_and_restart(tmp_path):
    """`pyflow run --json-stream` (reference cli.py:154-173 schema), periodic checkpoint with the reference's
    npz layout, restart with hash enforcement -- all through the device-backed driver."""
    import json
    ck = str(tmp_path / "ck.npz")
    common = ["--nx", "48", "--ny", "40", "--re", "100", "--lin-maxiter", "2000"]
    r = _run_cli("run", *common, "--steps", "4", "--json-stream", "--checkpoint", ck, "--checkpoint-interval", "3",
                 cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    assert [l["iteration"] for l in lines] == [0, 1, 2, 3] and all(l["type"] == "step" for l in lines)
    for l in lines:
        assert set(l) == {"type", "iteration", "dt", "CFL", "continuity", "wall_time", "residuals", "diagnostics"}
        assert set(l["residuals"]) == {"Ru", "Rv", "Rp", "continuity"}
        assert {"plateau_Ru", "max_residual", "residual_drop_Ru_orders", "nu", "Rp_iterations", "Rp_residual",
                "Rp_converged", "Rp_method", "divergence_norm", "nan_detected", "u_centerline"} <= set(l["diagnostics"])
        assert len(l["diagnostics"]["u_centerline"][1]) == 40
    assert lines[0]["dt"] == 0.0105 and lines[1]["dt"] == 0.011025000000000002      # reference dt sequence from rest
    # oracle agreement of the streamed scalars
    cfg = _cfg(dict(nx=48, ny=40, Re=100.0, advection_scheme="quick", lin_maxiter=2000))
    ost, od = orc.run(cfg, 4)
    for l, d in zip(lines, od):
        assert l["dt"] == d["dt"] and abs(l["residuals"]["Ru"] - d["Ru"]) <= 1e-9 * d["Ru"]
    # checkpoint written after the yield of iteration 3 (3 % 3 == 0) with the reference layout
    with np.load(ck) as z:
        meta = json.loads(str(z["meta"]))
        assert meta["iteration"] == 3 and z["field_u"].shape == (42, 50)
        assert rel(z["field_u"][1:-1, 1:-1], orc.interior(ost.fields["u"])) < 1e-9
    # restart: hash mismatch refused (exit 3), matching config resumes at iteration 4
    bad = _run_cli("run", "--nx", "48", "--ny", "40", "--re", "200", "--steps", "6", "--restart", ck, cwd=str(tmp_path))
    assert bad.returncode == 3
    # (checkpoint_interval is part of the hashed config, as in the reference: it must match too)
    ok = _run_cli("run", *common, "--steps", "6", "--json-stream", "--restart", ck, "--checkpoint-interval", "3",
                  cwd=str(tmp_path))
    assert ok.returncode == 0, ok.stdout[-500:] + ok.stderr[-500:]
    assert [json.loads(l)["iteration"] for l in ok.stdout.splitlines() if l.strip()] == [4, 5]
