"""More GPU parity (collected last):

* the legacy central operators of ``fluid_ops`` (gradient, laplacian, apply_pressure_correction:
  reference numerics/fluid_ops.py:43-125) -- BIT-EXACT against the golden vectors of the unmodified
  reference and against the oracle on larger / odd shapes (off the time-step path; the reference's
  own projection tests import them);
* whole steps on the TMA-fed CG kernel against the live oracle;
* BASELINE config 2 (512^2) against the summary fixture written by the unmodified reference.
"""
import numpy as np
import pytest

import pyflow_oracle as orc

pytestmark = pytest.mark.gpu

LEGACY_TAGS = ["8x8", "16x12", "5x9", "2x2", "3x3", "2x7", "9x2", "33x17", "70x66"]


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("tag", LEGACY_TAGS)
def test_legacy_central_operators_bit_exact_vs_golden(golden, tag):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.numerics import fluid_ops
    g = {k.split("/", 1)[1]: v for k, v in golden["legacy_ops"].items() if k.startswith(tag + "/")}
    dx, dy, dt = (float(t) for t in g["dxdydt"])
    gx, gy = fluid_ops.gradient(g["p"], dx, dy)
    assert bits_equal(gx, g["gx"]) and bits_equal(gy, g["gy"])
    assert bits_equal(fluid_ops.laplacian(g["p"], dx, dy), g["lap"])
    ny, nx = g["p"].shape
    st = allocate_state(nx, ny)
    interior_view(st.fields["u"])[:] = g["u"]
    interior_view(st.fields["v"])[:] = g["v"]
    fluid_ops.apply_pressure_correction(st, g["p"], dt, dx, dy)
    assert bits_equal(interior_view(st.fields["u"]), g["ucorr"]) and bits_equal(interior_view(st.fields["v"]), g["vcorr"])
    assert bits_equal(interior_view(st.fields["p"]), g["p"])
    lgx, lgy = fluid_ops.legacy_grad_p_state(st, dx, dy)
    assert bits_equal(lgx, g["gx"]) and bits_equal(lgy, g["gy"])


@pytest.mark.parametrize("shape,unit", [((257, 513), True), ((300, 1030), False), ((64, 3), True), ((3, 64), False)])
def test_legacy_central_operators_bit_exact_vs_oracle(shape, unit):
    from pyflow.numerics import fluid_ops
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx + 1)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    p = rng.standard_normal(shape)
    p[ny // 2, nx // 2] = np.nan
    with np.errstate(invalid="ignore"):
        gx, gy = fluid_ops.gradient(p, dx, dy)
        ogx, ogy = orc.gradient_central(p, dx, dy)
        assert bits_equal(gx, ogx) and bits_equal(gy, ogy)
        assert bits_equal(fluid_ops.laplacian(p, dx, dy), orc.laplacian_central(p, dx, dy))


def test_trajectory_on_the_tma_fed_cg_kernel_vs_live_oracle():
    """Whole steps on a grid that is too wide for the shared-memory-resident CG kernel (nx > 2048) and a
    multiple of 1024 columns, i.e. on the TMA-fed, dynamically scheduled kernel (k_cg_tmad) -- the same slab
    shape as the (33, 4096) case of test_pcg_strip_and_chunk_layouts_vs_oracle -- against the oracle."""
    from types import SimpleNamespace
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = SimpleNamespace(nx=4096, ny=33, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=150, advection_scheme="quick", disable_advection=False, diagnostics=False)
    steps = 4
    ost, odiags = orc.run(cfg, steps)
    st, tr = allocate_state(cfg.nx, cfg.ny), ResidualManager()
    for it in range(steps):
        _, _res, diag = step(cfg, st, tr, it)
        assert diag["dt"] == odiags[it]["dt"]
        assert st.meta["cg_true_iterations"] == odiags[it]["true_iterations"]
    for f in ("u", "v", "p"):
        a, b = st.fields[f][1:-1, 1:-1], orc.interior(ost.fields[f])
        e = float(np.linalg.norm(a - b) / (np.linalg.norm(b) or 1.0))
        assert e < 1e-9, (f, e)


def test_config2_512_on_the_gpu_vs_reference_summary(golden):
    """BASELINE config 2 (512x512, Re=1000, QUICK, 300 CG iterations per step), 4 steps from rest through the
    drop-in step(): against the summary fixture written by the UNMODIFIED reference (not the oracle)."""
    from types import SimpleNamespace
    from conftest import check_config2_summary
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    meta = golden["config2_512_meta"]
    cfg = SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, disable_advection=False,
                          diagnostics=False, **meta["cfg"])
    st, tr = allocate_state(cfg.nx, cfg.ny), ResidualManager()
    rows, iters = [], []
    for it in range(meta["steps"]):
        _, res, d = step(cfg, st, tr, it)
        rows.append([d["dt"], d["CFL"], res["Ru"], res["Rv"], res["continuity"], d["Rp_iterations"], d["Rp_residual"],
                     float(d["Rp_converged"]), d["divergence_norm"]])
        iters.append(st.meta["cg_true_iterations"])
    # scalars that sit downstream of a truncated CG solve get the bars of test_trajectory_vs_golden (1e-7) and a
    # looser one for the residual norms; the field samples carry the 1e-9 parity bar
    check_config2_summary(golden, {f: st.fields[f][1:-1, 1:-1] for f in ("u", "v", "p")}, rows, iters,
                          scalar_rtol=1e-7, residual_rtol=1e-5)


def test_background_checkpoint_writer_on_the_device_path(tmp_path):
    """Opt-in asynchronous checkpoints snapshot the fields straight from HBM: same arrays and meta on disk as the
    synchronous writer, and the State's host copies stay untouched (still stale, nothing re-uploaded)."""
    from pyflow.config.model import SimulationConfig
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.io.checkpoint import load_checkpoint
    cfg = SimulationConfig(nx=96, ny=80, Re=100.0, advection_scheme="quick", lin_maxiter=300, diagnostics=False)
    paths = {}
    for mode in ("sync", "async"):
        drv = SimulationDriver(cfg)
        drv.async_checkpoint = mode == "async"
        paths[mode] = str(tmp_path / f"{mode}.npz")
        for _ in drv.run(max_steps=5, checkpoint_path=paths[mode], checkpoint_interval=2):
            pass
        if mode == "async":
            assert drv.state.fields._stale == {"u", "v", "p"}
    (a, ma), (b, mb) = load_checkpoint(paths["async"]), load_checkpoint(paths["sync"])
    assert ma["iteration"] == mb["iteration"] == 4 and ma["field_hashes"] == mb["field_hashes"]
    for f in ("u", "v", "p"):
        assert bits_equal(a.fields[f], b.fields[f])


def test_tiny_grids_vs_reference_golden(golden):
    """The smallest grids the reference steps (2x2 ... 6x2, both schemes), 6 steps each, against golden
    trajectories written by the unmodified reference: dt / CFL bit-identical, CG iterations within 1, fields to
    1e-9 relative L2 (every solve converges in a handful of iterations)."""
    from conftest import TINY_TAGS, tiny_cfg
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    for tag in TINY_TAGS:
        cfg = tiny_cfg(tag)
        g = {k.split("/", 1)[1]: v for k, v in golden["tiny"].items() if k.startswith(tag + "/")}
        st, tr = allocate_state(cfg.nx, cfg.ny), ResidualManager()
        for it in range(len(g["true_iters"])):
            _, _res, d = step(cfg, st, tr, it)
            assert d["dt"] == g["scalars"][it, 0] and d["CFL"] == g["scalars"][it, 1], (tag, it)
            assert abs(st.meta["cg_true_iterations"] - g["true_iters"][it]) <= 1, (tag, it)
            assert d["Rp_iterations"] == g["scalars"][it, 5] and not d["nan_detected"], (tag, it)
        for f in ("u", "v", "p"):
            a, b = st.fields[f], g[f]
            e = float(np.linalg.norm(a - b) / (np.linalg.norm(b) or 1.0))
            assert e < 1e-9, (tag, f, e)


def test_assembled_matrix_jacobi_runs_on_the_device():
    """The reference's own recipe, M = jacobi_preconditioner(csr(assemble_negative_laplacian(...)))
    (pressure_solver.py:139-143), is recognised as the ring-identity diagonal: same pressure, bit for bit, as the
    automatic Jacobi path."""
    from types import SimpleNamespace
    import scipy.sparse as sp
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian, solve_pressure_poisson
    nx, ny, dx, dy = 24, 20, 0.5, 0.25
    M = jacobi_preconditioner(sp.csr_matrix(assemble_negative_laplacian(nx, ny, dx, dy)))
    out = []
    for pre, auto in ((M, False), (None, True)):
        cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=300, diagnostics=False, enable_jacobi_pc=auto)
        st = allocate_state(nx, ny)
        interior_view(st.fields["u"])[ny // 3, nx // 3] = 1.0
        p, diag = solve_pressure_poisson(st, 0.01, dx, dy, cfg, preconditioner=pre)
        out.append((p, diag["Rp_residual"]))
    assert bits_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


@pytest.mark.parametrize("method,nx,ny,dx,dy,maxiter,pc", [
    ("cg", 24, 20, 1.0, 1.0, 40, True), ("cg", 70, 33, 0.5, 0.25, 25, False), ("cg", 300, 200, 1.0, 1.0, 30, True),
    ("bicgstab", 24, 20, 1.0, 1.0, 400, True), ("bicgstab", 70, 33, 0.5, 0.25, 12, False),
    ("bicgstab", 40, 48, 1.0, 1.0, 2000, True)])
def test_solve_accepts_the_assembled_matrix_of_the_reference(method, nx, ny, dx, dy, maxiter, pc):
    """linear_solvers.solve(A, b) with A = assemble_negative_laplacian(...) as a scipy CSR matrix -- the one matrix the
    reference itself builds (pressure_solver.py:15-40) -- is recognised from its CSR arrays and solved on the device
    (k_cg_generic MODE_RING5 / k_bicgstab, products summed in CSR column order).  Checker: scipy itself on the same
    matrix, called the way the reference's interface.py:17-45 calls it.  The matrix is not symmetric (identity ring
    rows), so CG -- which the reference runs on it when asked -- is compared truncated, BiCGSTAB also converged."""
    import scipy.sparse.linalg as spla
    from pyflow.linear_solvers.interface import recognise_ring5, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian
    A = assemble_negative_laplacian(nx, ny, dx, dy)
    assert recognise_ring5(A) is not None
    rng = np.random.default_rng(nx * 131 + ny)
    b = rng.standard_normal(nx * ny)
    tol = 1e-10
    M_dev = jacobi_preconditioner(A) if pc else None
    got = solve(A, b, method=method, tol=tol, maxiter=maxiter, **({"M": M_dev} if pc else {}))
    D = A.diagonal()
    M_ref = spla.LinearOperator(A.shape, matvec=lambda v: v / D) if pc else None
    x_ref, info = (spla.cg if method == "cg" else spla.bicgstab)(A, b, rtol=tol, maxiter=maxiter, M=M_ref)
    rn = float(np.linalg.norm(b - A @ x_ref))
    assert got.iterations == (info if info > 0 else 0)
    assert got.converged == ((info == 0) and rn <= tol * np.linalg.norm(b))
    err = float(np.linalg.norm(got.x - x_ref) / np.linalg.norm(x_ref))
    assert err < (1e-9 if method == "cg" or info == 0 else 1e-7), (method, err, info)


def test_solve_refuses_matrices_it_cannot_map_onto_the_device():
    import scipy.sparse as sp
    from pyflow.linear_solvers.interface import solve
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian
    A = assemble_negative_laplacian(9, 7, 1.0, 1.0).tolil()
    A[10, 11] = -0.75                                   # one interior row differs: not the reference's matrix any more
    with pytest.raises(NotImplementedError):
        solve(A.tocsr(), np.ones(63), method="cg")
    with pytest.raises(NotImplementedError):
        solve(sp.identity(63).tocsr(), np.ones(63), method="cg")


def test_exact_restart_on_the_device_path(tmp_path):
    """SURVEY 8f item 3 (optional part): checkpoint after 5 steps, load, restore the saved dt-controller memory
    (``step_state``), run 4 more: the same dt sequence as the uninterrupted 9 steps (through the safeguard halving)
    and the same fields, to the bar of the reference's own restart test (tests/test_checkpoint_restart.py:49-51,
    atol 1e-12)."""
    from pyflow.config.model import SimulationConfig
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.io.checkpoint import load_checkpoint, restore_step_state
    cfg = SimulationConfig(nx=96, ny=80, Re=100.0, advection_scheme="quick", lin_maxiter=300, diagnostics=False)
    total, cut = 9, 5
    full = SimulationDriver(cfg)
    full_dt = [d["dt"] for _s, _r, d in full.run(max_steps=total)]
    ck = str(tmp_path / "ck.npz")
    for _ in SimulationDriver(cfg).run(max_steps=cut, checkpoint_path=ck, checkpoint_interval=1):
        pass
    st, meta = load_checkpoint(ck)
    assert meta["iteration"] == cut - 1 and restore_step_state(st, meta)
    dts = [d["dt"] for _s, _r, d in SimulationDriver(cfg, st).run(max_steps=total, start_iteration=cut)]
    assert dts == full_dt[cut:]
    for f in ("u", "v", "p"):
        assert np.allclose(st.fields[f], full.state.fields[f], rtol=0.0, atol=1e-12), f


BICG_TAGS = ["24x20_plain", "24x20_jacobi", "33x17_jacobi_nonunit", "40x28_truncated7", "16x16_default_tol",
             "12x10_zero_rhs"]


@pytest.mark.parametrize("tag", BICG_TAGS)
def test_bicgstab_seam_vs_reference_golden(tag):
    """``solve(A, b)`` with the signature's default ``method='bicgstab'`` (interface.py:23,44-45) on the device
    (k_bicgstab) against what the UNMODIFIED reference returned for the same right-hand sides
    (tests/golden/bicgstab.npz): the truncated solve is the same iterate up to summation order (1e-10), converged
    solves meet the reference's convergence criterion and agree with its solution to what the tolerance allows
    (2 tol ||b|| / sigma_min, sigma_min > 1e-3 on these grids), ``converged`` / ``iterations`` bookkeeping equal."""
    import os
    from conftest import GOLDEN
    from pyflow.linear_solvers.interface import NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    with np.load(os.path.join(GOLDEN, "bicgstab.npz")) as z:
        g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/")}
    nx, ny, dx, dy, jac, tol, maxiter = g["params"]
    A = NegLaplacianPinned(int(nx), int(ny), float(dx), float(dy))
    kw = {"M": jacobi_preconditioner(A)} if jac else {}
    res = solve(A, g["b"].copy(), tol=float(tol), maxiter=int(maxiter), **kw)
    conv_ref, its_ref, _rnorm_ref, _done_ref = g["result"]
    bnorm = float(np.linalg.norm(g["b"])) or 1.0
    assert res.method == "bicgstab" and res.x.shape == g["b"].shape
    assert res.iterations == int(its_ref)
    # `converged` also asks the TRUE residual to be <= tol ||b|| (interface.py:50); the recurrence stops on the
    # recursive one, so a solve that ends within rounding drift of the threshold may legitimately say False
    assert res.converged == (res.iterations == 0 and res.residual_norm <= tol * bnorm)
    true_res = float(np.linalg.norm(g["b"] - orc.neg_laplacian_pinned(res.x, int(nx), int(ny), dx, dy)))
    assert abs(res.residual_norm - true_res) <= 1e-6 * max(true_res, 1e-300) or true_res == 0.0
    if "truncated" in tag:
        assert np.linalg.norm(res.x - g["x"]) <= 1e-10 * np.linalg.norm(g["x"])
    elif "zero_rhs" in tag:
        assert not res.x.any()
    else:
        assert bool(conv_ref) and res.residual_norm <= 1.05 * tol * bnorm
        assert np.linalg.norm(res.x - g["x"]) <= 2.05 * tol * bnorm / 1e-3


def test_bicgstab_on_a_grid_of_many_ctas_vs_oracle():
    """A grid large enough for several hundred CTAs (257 x 513): a truncated BiCGSTAB solve equals the oracle's
    iterate up to summation order; a second call on the same context gives the same bits (partial-slot reuse)."""
    from pyflow.linear_solvers.interface import NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    nx, ny = 513, 257
    b = np.random.default_rng(11).standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    A = NegLaplacianPinned(nx, ny, 1.0, 1.0)
    r1 = solve(A, b.copy(), tol=1e-10, maxiter=12, M=jacobi_preconditioner(A))
    r2 = solve(A, b.copy(), tol=1e-10, maxiter=12, M=jacobi_preconditioner(A))
    xo, _conv, its, _rn, info = orc.interface_solve(b.copy(), nx, ny, 1.0, 1.0, "bicgstab", 1e-10, 12, True)
    assert info == 12 and r1.iterations == its == 12 and not r1.converged
    assert np.array_equal(r1.x, r2.x)
    assert np.linalg.norm(r1.x - xo) <= 1e-10 * np.linalg.norm(xo)


@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_user_supplied_diagonal_preconditioner_on_the_device(method):
    """``solve(A, b, M=<foreign operator>)`` (the seam accepts an external M): a scipy LinearOperator
    that proves to be diagonal runs on the device from an array of its entries (k_cg_diag / k_bicgstab) and gives
    scipy's own iterate for the same operator and preconditioner (truncated solve: summation order only, 1e-10);
    the reference-style ring-identity operator takes the tuned CG kernels bit for bit; an operator that is not
    diagonal is refused."""
    import scipy.sparse.linalg as spla
    from pyflow.linear_solvers.interface import NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    nx, ny, dx, dy = 45, 31, 0.5, 0.25
    n = nx * ny
    rng = np.random.default_rng(23)
    b = rng.standard_normal(n)
    b -= b.mean()
    b[0] = 0.0
    d = rng.uniform(0.05, 1.0, n)
    mv = lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy)                       # noqa: E731
    A_sp = spla.LinearOperator((n, n), matvec=mv, dtype=np.float64)
    M_sp = spla.LinearOperator((n, n), matvec=lambda t: d * t, dtype=np.float64)
    A = NegLaplacianPinned(nx, ny, dx, dy)
    fn = spla.cg if method == "cg" else spla.bicgstab
    xs, info = fn(A_sp, b, x0=np.zeros(n), rtol=1e-10, maxiter=10, M=M_sp)
    res = solve(A, b.copy(), method=method, tol=1e-10, maxiter=10, M=M_sp)
    assert info == 10 and res.iterations == 10 and not res.converged
    assert np.linalg.norm(res.x - xs) <= 1e-10 * np.linalg.norm(xs)
    full = solve(A, b.copy(), method=method, tol=1e-10, maxiter=20000, M=M_sp)
    assert full.iterations == 0 and full.residual_norm <= 1.05e-10 * np.linalg.norm(b)
    if method == "cg":
        invd = orc.jacobi_inverse_diagonal(nx, ny, dx, dy)
        ring_op = spla.LinearOperator((n, n), matvec=lambda t: invd * t, dtype=np.float64)
        r1 = solve(A, b.copy(), method="cg", tol=1e-10, maxiter=50, M=ring_op)
        r2 = solve(A, b.copy(), method="cg", tol=1e-10, maxiter=50, M=jacobi_preconditioner(A))
        assert bits_equal(r1.x, r2.x)
        with pytest.raises(NotImplementedError):
            solve(A, b.copy(), method="cg", M=spla.LinearOperator((n, n), matvec=lambda t: np.roll(t, 1)))
