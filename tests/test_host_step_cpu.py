"""Host side of the drop-in exercised END TO END on a machine without a GPU.

The GPU context is replaced by the oracle-backed stand-in (``tests/cpu_slab.py``), so what
these tests check is everything the product does on the HOST around the ``pf_step`` call:
the dt / safeguard floats, the residual and diagnostics dictionaries, the lazy State
synchronisation protocol (which fields are uploaded / downloaded and when), the generator
driver, checkpoints of device-resident fields, the CLI run loop and its JSON stream, the
alternate dict-payload driver, the array-level operator seams.  The kernels themselves are
checked by the ``-m gpu`` suite.
"""
import io
import json
from types import SimpleNamespace

import numpy as np
import pytest

import cpu_slab
import pyflow_oracle as orc


@pytest.fixture()
def slab_cls(monkeypatch):
    return cpu_slab.install(monkeypatch)


def make_cfg(**kw):
    base = dict(nx=24, ny=20, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
    base.update(kw)
    return SimpleNamespace(**base)


# ------------------------------------------------------------------ whole trajectories through step()
@pytest.mark.parametrize("kw,steps", [
    (dict(nx=32, ny=32, lin_maxiter=2000), 6),
    (dict(nx=40, ny=24, Re=400.0, lx=2.0, ly=1.0, lin_maxiter=3000), 6),
    (dict(nx=17, ny=33, Re=50.0, advection_scheme="upwind", enable_jacobi_pc=False, lin_maxiter=2000), 5),
    (dict(nx=24, ny=20, disable_advection=True), 4),
    (dict(nx=24, ny=20, lid_velocity=None, advection_scheme="anything-else-is-upwind"), 3),
])
def test_step_matches_the_oracle_step_for_step(slab_cls, kw, steps):
    """Same floats as the reference: dt, CFL, residuals, Rp_* and the fields, every step."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = make_cfg(**kw)
    st, tr = allocate_state(cfg.nx, cfg.ny), ResidualManager()
    ost = orc.allocate_state(cfg.nx, cfg.ny)
    if cfg.lid_velocity is None:                           # nothing drives the flow: seed a vortex
        yy, xx = np.mgrid[0:cfg.ny, 0:cfg.nx]
        for s in (st, ost):
            s.fields["u"][1:-1, 1:-1] = np.sin(0.3 * xx) * np.cos(0.2 * yy)
            s.fields["v"][1:-1, 1:-1] = -np.cos(0.3 * xx) * np.sin(0.2 * yy)
    for it in range(steps):
        same, res, diag = step(cfg, st, tr, it)
        _, ores, odiag = orc.step(cfg, ost, it)
        assert same is st
        assert diag["dt"] == odiag["dt"] and diag["CFL"] == odiag["CFL"] and diag["iteration"] == it
        for k in ("Ru", "Rv", "Rp", "continuity"):
            assert res[k] == ores[k], (it, k)
        for k in ("Rp_iterations", "Rp_residual", "Rp_converged", "Rp_method", "divergence_norm", "nan_detected",
                  "max_residual", "nu"):
            assert diag[k] == odiag[k], (it, k)
        assert st.meta["cg_true_iterations"] == odiag["true_iterations"]
        assert st.meta["dt_prev"] == ost.meta["dt_prev"]
        assert st.meta["divergence_history"] == ost.meta["divergence_history"]
        assert st.meta["prev_div_norm"] == ost.meta["prev_div_norm"]
    for f in ("u", "v", "p"):
        assert np.array_equal(st.fields[f], ost.fields[f]), f
    assert tr.series["Ru"].values[0] == 1.0 and len(tr.series["Ru"].values) == steps + 1
    assert set(tr.series) == {"Ru"}                        # the reference's tracker is a stub that keeps Ru only


def test_dt_sequence_known_answers_from_rest(slab_cls):
    """SURVEY 8c KAT: growth-limited dt from rest, then the divergence safeguard halves it."""
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = make_cfg(nx=32, ny=32, lin_maxiter=10000)
    st, tr = allocate_state(32, 32), ResidualManager()
    dts = [step(cfg, st, tr, it)[2]["dt"] for it in range(5)]
    assert dts == [0.0105, 0.011025000000000002, 0.011576250000000003, 0.012155062500000004, 0.006381407812500003]


# ------------------------------------------------------------------ lazy State protocol around step()
def test_fields_stay_on_the_device_between_steps(slab_cls):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = make_cfg()
    st, tr = allocate_state(cfg.nx, cfg.ny), ResidualManager()
    for it in range(3):
        step(cfg, st, tr, it)
    slab = st.fields._slab
    assert isinstance(slab, slab_cls) and slab.steps == 3
    assert slab.uploads == [("u", "v", "p")] and slab.downloads == []      # one upload, nothing pulled back
    # the pre-step reductions of step k+1 come from step k's post pass: no extra device call is needed,
    # and a host write invalidates that cache
    assert st.fields._prestep is not None
    u = st.fields["u"]                                                     # materialise: one D2H of u only
    assert slab.downloads == [("u",)] and st.fields.host_touched() == {"u"}
    interior_view(u)[3, 4] += 0.25                                         # host mutation between steps ...
    step(cfg, st, tr, 3)
    assert slab.uploads[-1] == ("u",)                                      # ... is uploaded (u only) before the step
    ost = orc.allocate_state(cfg.nx, cfg.ny)
    for it in range(3):
        orc.step(cfg, ost, it)
    orc.interior(ost.fields["u"])[3, 4] += 0.25
    orc.step(cfg, ost, 3)
    for f in ("u", "v", "p"):
        assert np.array_equal(st.fields[f], ost.fields[f])


def test_eager_mode_gives_live_array_semantics(slab_cls, monkeypatch):
    """``state.fields.eager = True``: views taken ONCE stay live across steps, both ways, like the reference's
    plain numpy State (upload before / download after every call)."""
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.drivers.simulation_driver import SimulationDriver
    cfg = make_cfg(nx=16, ny=12)
    st = allocate_state(16, 12)
    st.fields.eager = True
    ui = interior_view(st.fields["u"])                     # taken once, never re-fetched
    ost = orc.allocate_state(16, 12)
    for k, (s_, _r, d) in enumerate(SimulationDriver(cfg, st).run(max_steps=4)):
        orc.step(cfg, ost, k)
        assert np.array_equal(ui, orc.interior(ost.fields["u"]))               # reads see the new step
        assert d["u_centerline"][1] == ui[:, 8].tolist()
        if k == 1:                                                              # writes through the old view count
            ui[5, 5] += 0.125
            orc.interior(ost.fields["u"])[5, 5] += 0.125
    slab = st.fields._slab
    assert len(slab.uploads) == 4 and len(slab.downloads) == 4 and all(len(t) == 3 for t in slab.uploads)
    # default (lazy) mode for contrast: the old view is stale until state.fields is consulted again
    st2 = allocate_state(16, 12)
    u2 = interior_view(st2.fields["u"])
    list(SimulationDriver(cfg, st2).run(max_steps=2))
    assert not u2.any() and st2.fields._slab.downloads == []
    assert interior_view(st2.fields["u"]) is not None and u2.any()              # same array, refreshed in place
    monkeypatch.setenv("PYFLOW_B200_EAGER", "1")
    assert allocate_state(4, 4).fields.eager is True


def test_every_mutating_dict_method_keeps_the_device_copy_coherent(slab_cls):
    """``update``, ``setdefault``, ``pop``, ``del``, ``|=`` and copies of ``state.fields`` go through the same
    stale / touched bookkeeping as ``__setitem__`` (the inherited dict methods would bypass it: a field replaced
    with ``update`` would be overwritten by the next download and ignored by the next step); copies of a State are
    host snapshots that never share the GPU context."""
    import copy
    import pickle
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = make_cfg(nx=12, ny=9, lin_maxiter=8)
    st, tr = allocate_state(12, 9), ResidualManager()
    model = orc.allocate_state(12, 9)
    for it in range(2):
        step(cfg, st, tr, it)
        orc.step(cfg, model, it)
    new_u = model.fields["u"] * 0.5
    st.fields.update(u=new_u.copy())                       # replaced through update(): must be uploaded, not lost
    model.fields["u"] = new_u
    assert "u" in st.fields.host_touched() and np.array_equal(st.fields["u"], new_u)
    st.fields |= {"v": model.fields["v"] * 2.0}
    model.fields["v"] = model.fields["v"] * 2.0
    assert st.fields.setdefault("p") is st.fields.peek("p")
    assert st.fields.setdefault("extra", 7) == 7 and st.fields.pop("extra") == 7 and "extra" not in st.fields
    step(cfg, st, tr, 2)
    orc.step(cfg, model, 2)
    snap = copy.deepcopy(st)                               # device copy is newer than the host arrays here
    assert snap.fields._slab is None and snap.fields is not st.fields
    clone = pickle.loads(pickle.dumps(st.fields))
    for f in ("u", "v", "p"):
        assert np.array_equal(snap.fields[f], model.fields[f]) and np.array_equal(clone[f], model.fields[f])
        assert np.array_equal(st.fields[f], model.fields[f])
    popped = st.fields.pop("p")                            # pop of a stale field returns the current contents
    assert np.array_equal(popped, model.fields["p"])
    st.fields["p"] = popped
    step(cfg, st, tr, 3)
    orc.step(cfg, model, 3)
    for f in ("u", "v", "p"):
        assert np.array_equal(st.fields[f], model.fields[f])
    del st.fields["p"]
    assert "p" not in st.fields and "p" not in st.fields.host_touched()


@pytest.mark.parametrize("seed", range(12))
def test_lazy_state_protocol_random_walk(slab_cls, seed, tmp_path):
    """Random interleavings of steps, reads, in-place writes through freshly fetched arrays, whole-array
    replacement, read-only checkpoint reads and projection-only calls: the device-resident State must stay
    indistinguishable from plain numpy arrays stepped by the oracle, and must move only what was touched."""
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.io.checkpoint import load_checkpoint, save_checkpoint
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.pressure_solver import solve_pressure_poisson
    from pyflow.solvers.solver import step
    rng = np.random.default_rng(1000 + seed)
    cfg = make_cfg(nx=10, ny=7, lin_maxiter=6, advection_scheme=("quick", "upwind")[seed % 2])
    st, tr = allocate_state(10, 7), ResidualManager()
    model = orc.allocate_state(10, 7)
    it = 0
    for _ in range(40):
        op = rng.choice(["step", "step", "read", "write", "replace", "checkpoint", "project"])
        name = str(rng.choice(["u", "v", "p"]))
        if op == "step":
            up0 = len(st.fields._slab.uploads) if st.fields._slab else 0
            dirty = st.fields.host_touched()
            step(cfg, st, tr, it)
            orc.step(cfg, model, it)
            it += 1
            ups = st.fields._slab.uploads[up0:]
            assert [set(t) for t in ups] == ([dirty] if dirty else [])          # only touched fields go up
        elif op == "read":
            assert np.array_equal(st.fields[name], model.fields[name]), (op, name)
        elif op == "write":                                                    # fetch, then modify in place
            j, i = int(rng.integers(0, 7)), int(rng.integers(0, 10))
            delta = float(rng.standard_normal()) * 1e-3
            interior_view(st.fields[name])[j, i] += delta
            orc.interior(model.fields[name])[j, i] += delta
        elif op == "replace":                                                  # state.fields[name] = new array
            new = model.fields[name] * 0.5
            st.fields[name] = new.copy()
            model.fields[name] = new
        elif op == "checkpoint":                                               # read-only: nothing becomes dirty
            dirty = st.fields.host_touched()
            path = str(tmp_path / "rw.npz")
            save_checkpoint(path, st, it, 0.0, cfg)
            assert st.fields.host_touched() == dirty
            back, _meta = load_checkpoint(path)
            for f in ("u", "v", "p"):
                assert np.array_equal(back.fields[f], model.fields[f]), (op, f)
        else:                                                                  # projection-only seam on the same state
            p, _d = solve_pressure_poisson(st, 0.01, 1.0, 1.0, cfg)
            op_, _od = orc.solve_pressure_poisson(model, 0.01, 1.0, 1.0, cfg)
            assert np.array_equal(p, op_)
    for f in ("u", "v", "p"):
        assert np.array_equal(st.fields[f], model.fields[f]), f


def test_foreign_state_objects_are_synchronised_eagerly(slab_cls):
    """A state that is not ours (plain dict fields, e.g. the reference's dataclass): upload before and
    download after every call, into the caller's own arrays."""
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    cfg = make_cfg(nx=16, ny=12)
    foreign = SimpleNamespace(nx=16, ny=12, fields={k: np.zeros((14, 18)) for k in "uvp"}, meta={})
    u_arr = foreign.fields["u"]
    ost = orc.allocate_state(16, 12)
    for it in range(3):
        step(cfg, foreign, ResidualManager(), it)
        orc.step(cfg, ost, it)
        assert foreign.fields["u"] is u_arr and np.array_equal(u_arr, ost.fields["u"])
    assert len(foreign._pf_slab.uploads) == 3 and len(foreign._pf_slab.downloads) == 3
    with pytest.raises(TypeError):
        bad = SimpleNamespace(nx=16, ny=12, fields={k: np.zeros((14, 18), dtype=np.float32) for k in "uvp"}, meta={})
        step(cfg, bad, ResidualManager(), 0)


# ------------------------------------------------------------------ solve_pressure_poisson seam
def test_solve_pressure_poisson_diagnostics_and_modes(slab_cls):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.linear_solvers.interface import NegLaplacianPinned
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    from pyflow.solvers.pressure_solver import solve_pressure_poisson
    nx, ny = 24, 20
    cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=300, diagnostics=False)
    results = {}
    for tag, kw, cfgkw in (("auto", {}, {}), ("none", {}, dict(enable_jacobi_pc=False)),
                           ("explicit", dict(preconditioner=jacobi_preconditioner(NegLaplacianPinned(nx, ny, 1.0, 1.0))),
                            dict(enable_jacobi_pc=False))):
        st = allocate_state(nx, ny)
        interior_view(st.fields["u"])[ny // 3, nx // 3] = 1.0
        c = SimpleNamespace(**{**vars(cfg), **cfgkw})
        p, diag = solve_pressure_poisson(st, 0.01, 1.0, 1.0, c, **kw)
        ost = orc.allocate_state(nx, ny)
        orc.interior(ost.fields["u"])[ny // 3, nx // 3] = 1.0
        oc = SimpleNamespace(**{**vars(cfg), "enable_jacobi_pc": tag != "none"})
        op, odiag = orc.solve_pressure_poisson(ost, 0.01, 1.0, 1.0, oc)
        assert np.array_equal(p, op) and p.shape == (ny, nx)
        for k in ("Rp_iterations", "Rp_residual", "Rp_converged", "Rp_method", "divergence_norm",
                  "divergence_norm_initial"):
            assert diag[k] == odiag[k], (tag, k)
        assert set(diag) >= {"lin_residual_norm", "lin_residual_rel", "divergence_reduction_factor",
                             "divergence_norm_after_raw", "divergence_norm_final"}
        assert np.array_equal(st.fields["u"], ost.fields["u"]) and np.array_equal(st.fields["p"], ost.fields["p"])
        results[tag] = p
    assert np.array_equal(results["auto"], results["explicit"])
    # reference tests/test_preconditioner_effect.py:23: preconditioned and plain solves agree to 1e-9
    assert np.max(np.abs(results["auto"] - results["none"])) <= 1e-9 * max(1.0, np.max(np.abs(results["none"])))
    with pytest.raises(NotImplementedError):
        solve_pressure_poisson(allocate_state(nx, ny), 0.01, 1.0, 1.0, cfg, preconditioner=object())
    # manufactured / Dirichlet mode (rhs_override): velocities untouched, p solves lap(p) = rhs
    X, Y = np.meshgrid(np.linspace(0, 1, 14), np.linspace(0, 1, 11))
    rhs = np.zeros((11, 14))
    rhs[1:-1, 1:-1] = (-2.0 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y))[1:-1, 1:-1]
    st = allocate_state(14, 11)
    interior_view(st.fields["u"])[:] = 3.0
    mcfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=500, diagnostics=False)
    p, diag = solve_pressure_poisson(st, 1.0, 1 / 13, 1 / 10, mcfg, rhs_override=rhs.reshape(-1))
    op, odiag = orc.solve_manufactured(orc.allocate_state(14, 11), 1 / 13, 1 / 10, mcfg, rhs.reshape(-1))
    assert np.array_equal(p, op) and diag["Rp_iterations"] == odiag["Rp_iterations"]
    assert diag["divergence_norm"] == odiag["divergence_norm"] and diag["divergence_norm_initial"] is None
    assert np.all(interior_view(st.fields["u"]) == 3.0) and np.array_equal(interior_view(st.fields["p"]), op)


# ------------------------------------------------------------------ driver, checkpoint, restart
def test_driver_protocol_checkpoint_and_restart(slab_cls, tmp_path):
    from pyflow.config.model import SimulationConfig
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.io.checkpoint import load_checkpoint
    cfg = SimulationConfig(nx=20, ny=16, Re=100.0, advection_scheme="quick", lin_maxiter=300, diagnostics=False)
    ck = str(tmp_path / "ck.npz")
    drv = SimulationDriver(cfg)
    seen, states = [], []
    for st, res, diag in drv.run(max_steps=5, checkpoint_path=ck, checkpoint_interval=2):
        seen.append((diag["iteration"], diag["dt"], res["continuity"], diag.get("checkpoint_written")))
        states.append(st)
        ys, uc = diag["u_centerline"]
        assert len(ys) == 16 and ys[-1] == 15.0 and uc == st.fields._slab.u[:, 10].tolist()
        assert diag["wall_time"] >= 0.0
    assert [s[0] for s in seen] == [0, 1, 2, 3, 4] and all(s is states[0] for s in states)
    assert drv.iteration == 5 and all(s[3] is None for s in seen)          # written AFTER the yield
    slab = drv.state.fields._slab
    # checkpoints at it = 0, 2, 4 read the fields through the read-only accessor: nothing is re-uploaded
    assert slab.uploads == [("u", "v", "p")] and len(slab.downloads) == 9
    st2, meta = load_checkpoint(ck)
    assert meta["iteration"] == 4 and meta["structured_config_hash"] == cfg.config_hash
    ost, odiags = orc.run(cfg, 5)
    for f in ("u", "v", "p"):
        assert np.array_equal(st2.fields[f], ost.fields[f])
    assert [s[1] for s in seen] == [d["dt"] for d in odiags]
    # restart: meta (dt_prev, history) is not checkpointed, exactly as in the reference, so the resumed
    # run restarts the dt ramp; compare with the oracle doing the same
    drv2 = SimulationDriver(cfg, st2)
    out = list(drv2.run(max_steps=7, start_iteration=meta["iteration"] + 1))
    assert [d["iteration"] for _s, _r, d in out] == [5, 6]
    ost.meta.clear()
    for it in (5, 6):
        _, _, od = orc.step(cfg, ost, it)
    assert out[-1][2]["dt"] == od["dt"] and np.array_equal(st2.fields["u"], ost.fields["u"])


def test_background_checkpoint_writer(slab_cls, tmp_path, monkeypatch):
    """Opt-in asynchronous checkpoints: snapshot at the reference's point in the loop, same bytes on disk as the
    synchronous writer, written while later steps run, complete when the generator ends or is closed."""
    import threading

    import pyflow.io.checkpoint as ckpt
    from pyflow.config.model import SimulationConfig
    from pyflow.drivers.simulation_driver import SimulationDriver
    cfg = SimulationConfig(nx=20, ny=16, Re=100.0, advection_scheme="quick", lin_maxiter=300, diagnostics=False)
    sync_path, async_path = str(tmp_path / "sync.npz"), str(tmp_path / "async.npz")
    for _ in SimulationDriver(cfg).run(max_steps=4, checkpoint_path=sync_path, checkpoint_interval=3):
        pass
    # slow the worker down: the write of checkpoint 0 is released only after two more steps have run
    gate, started = threading.Event(), threading.Event()
    real_write = ckpt._write_npz

    def slow_write(path, meta, fields):
        started.set()
        assert gate.wait(timeout=60)
        return real_write(path, meta, fields)

    monkeypatch.setattr(ckpt, "_write_npz", slow_write)
    drv = SimulationDriver(cfg)
    drv.async_checkpoint = True
    gen = drv.run(max_steps=4, checkpoint_path=async_path, checkpoint_interval=3)
    seen = []
    for st, _res, diag in gen:
        seen.append(diag["iteration"])
        if diag["iteration"] == 2:
            assert started.wait(timeout=60)                     # checkpoint 0 was handed over after the first yield ...
            assert not (tmp_path / "async.npz").exists()        # ... and is still being written two steps later
            gate.set()
    assert seen == [0, 1, 2, 3] and (tmp_path / "async.npz").exists()        # joined when the generator finished
    slab = drv.state.fields._slab
    assert slab.uploads == [("u", "v", "p")]                    # snapshots do not dirty the State's host copies
    assert drv.state.fields._stale == {"u", "v", "p"}           # ... nor refresh them
    a, ma = ckpt.load_checkpoint(async_path)
    b, mb = ckpt.load_checkpoint(sync_path)
    assert ma["iteration"] == mb["iteration"] == 3
    assert {k: v for k, v in ma.items() if k != "time"} == {k: v for k, v in mb.items() if k != "time"}
    assert list(ma) == list(mb)                                 # same key order in the meta JSON
    for f in ("u", "v", "p"):
        assert np.array_equal(a.fields[f], b.fields[f])
    # the snapshot is isolated from later steps: checkpoint 0 (first pass, written late) held step-0 fields
    monkeypatch.setattr(ckpt, "_write_npz", real_write)
    w = ckpt.BackgroundCheckpointWriter()
    drv2 = SimulationDriver(cfg)
    g2 = drv2.run(max_steps=3)
    st, _r, _d = next(g2)
    w.submit(str(tmp_path / "c0.npz"), st, 0, 0.0, cfg)
    for _ in g2:
        pass
    w.wait()
    ost, _ = orc.run(cfg, 1)
    c0, _m = ckpt.load_checkpoint(str(tmp_path / "c0.npz"))
    assert np.array_equal(c0.fields["u"], ost.fields["u"]) and w.written == 1
    # errors of the worker surface in wait(); closing the generator early still joins the writer
    w.submit(str(tmp_path / "no" / "\0bad" / "x.npz"), st, 0, 0.0, cfg)
    with pytest.raises(Exception):
        w.wait()
    drv3 = SimulationDriver(cfg)
    drv3.async_checkpoint = True
    g3 = drv3.run(max_steps=10, checkpoint_path=str(tmp_path / "early.npz"), checkpoint_interval=1)
    next(g3); next(g3)
    g3.close()
    assert ckpt.load_checkpoint(str(tmp_path / "early.npz"))[1]["iteration"] == 0


def test_nan_is_reported_and_logged_not_raised(slab_cls, tmp_path, monkeypatch):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    monkeypatch.chdir(tmp_path)
    buf = io.StringIO()
    cfg = make_cfg(nx=10, ny=10, log_stream=buf, emergency_checkpoint_path=str(tmp_path / "em.npz"))
    st = allocate_state(10, 10)
    interior_view(st.fields["u"])[4, 4] = np.nan
    _, res, diag = step(cfg, st, ResidualManager(), 0)
    assert diag["nan_detected"] is True and diag["Rp_converged"] is False and diag["Rp_iterations"] == 300
    assert np.isnan(res["Ru"]) or np.isnan(res["continuity"])
    recs = [json.loads(line) for line in buf.getvalue().splitlines()]
    assert [r["type"] for r in recs] == ["step", "error", "emergency_checkpoint"]
    assert recs[1]["reason"] == "nan_detected" and recs[0]["nan"] is True
    assert (tmp_path / "em_FAIL_NAN.npz").exists()


# ------------------------------------------------------------------ CLI run loop, in process
def test_cli_run_json_stream_checkpoint_restart(slab_cls, tmp_path, capsys):
    from pyflow import cli
    ck = str(tmp_path / "ck.npz")
    rc = cli.main(["run", "--nx", "16", "--ny", "12", "--steps", "4", "--json-stream", "--checkpoint", ck,
                   "--checkpoint-interval", "1", "--seed", "3"])
    assert rc == 0
    lines = [json.loads(line) for line in capsys.readouterr().out.splitlines() if line.startswith("{")]
    assert [p["iteration"] for p in lines] == [0, 1, 2, 3] and all(p["type"] == "step" for p in lines)
    first = lines[1]
    assert set(first) == {"type", "iteration", "dt", "CFL", "continuity", "wall_time", "residuals", "diagnostics",
                          "run_meta"}
    assert set(first["residuals"]) == {"Ru", "Rv", "Rp", "continuity"} and first["run_meta"] == {"seed": 3}
    assert {"Rp_iterations", "Rp_residual", "Rp_converged", "Rp_method", "divergence_norm", "nan_detected",
            "u_centerline", "plateau_Ru", "max_residual", "residual_drop_Ru_orders", "nu",
            "seed"} <= set(first["diagnostics"])
    ocfg = make_cfg(nx=16, ny=12, Re=100.0)
    _, odiags = orc.run(ocfg, 4)
    assert [p["dt"] for p in lines] == [d["dt"] for d in odiags]
    assert [p["continuity"] for p in lines] == [d["continuity"] for d in odiags]
    # restart from the checkpoint of iteration 3 with the same configuration: continues at 4
    same = ["--checkpoint", ck, "--checkpoint-interval", "1", "--seed", "3"]     # part of the hashed configuration
    rc = cli.main(["run", "--nx", "16", "--ny", "12", "--steps", "6", "--json-stream", "--restart", ck, *same])
    assert rc == 0
    lines = [json.loads(line) for line in capsys.readouterr().out.splitlines() if line.startswith("{")]
    assert [p["iteration"] for p in lines] == [4, 5]
    # a different configuration is refused with exit code 3
    rc = cli.main(["run", "--nx", "16", "--ny", "12", "--re", "55", "--steps", "6", "--restart", ck, *same])
    assert rc == 3


def test_exact_restart_continues_the_uninterrupted_run(slab_cls, tmp_path):
    """SURVEY 8f item 3 (optional part): the dt controller's memory (dt_prev, divergence_history, prev_div_norm)
    is saved under the extra meta key ``step_state``; restored on request, the resumed run continues the
    uninterrupted one -- same dt sequence through the safeguard halving, same fields bit for bit.  Without the
    request a restart begins a fresh dt ramp, as in the reference."""
    import os
    import subprocess
    import sys
    from pyflow.config.model import SimulationConfig
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.io.checkpoint import STEP_STATE_KEYS, load_checkpoint, restore_step_state
    cfg = SimulationConfig(nx=20, ny=16, Re=100.0, advection_scheme="quick", lin_maxiter=300, diagnostics=False)
    total, cut = 9, 5
    full = SimulationDriver(cfg)
    full_dt = [d["dt"] for _s, _r, d in full.run(max_steps=total)]
    assert any(b < a for a, b in zip(full_dt, full_dt[1:]))          # the ramp is not monotone: history matters
    for mode in ("sync", "async"):
        ck = str(tmp_path / f"{mode}.npz")
        part = SimulationDriver(cfg)
        part.async_checkpoint = mode == "async"
        for _ in part.run(max_steps=cut, checkpoint_path=ck, checkpoint_interval=1):
            pass
        st, meta = load_checkpoint(ck)
        assert meta["iteration"] == cut - 1 and st.meta == {}        # loading alone restores nothing (reference)
        saved = meta["step_state"]
        assert set(saved) == set(STEP_STATE_KEYS) and list(meta)[-1] == "step_state"
        assert saved["dt_prev"] == part.state.meta["dt_prev"] == full_dt[cut - 1]
        assert saved["divergence_history"] == part.state.meta["divergence_history"]
        assert restore_step_state(st, meta) and st.meta["dt_prev"] == full_dt[cut - 1]
        resumed = SimulationDriver(cfg, st)
        dts = [d["dt"] for _s, _r, d in resumed.run(max_steps=total, start_iteration=cut)]
        assert dts == full_dt[cut:]
        for f in ("u", "v", "p"):
            assert np.array_equal(st.fields[f], full.state.fields[f])
    # default restart: fresh ramp, different dt
    st, meta = load_checkpoint(ck)
    plain = [d["dt"] for _s, _r, d in SimulationDriver(cfg, st).run(max_steps=total, start_iteration=cut)]
    assert plain != full_dt[cut:] and plain[0] <= 1e-2 * cfg.cfl_growth
    # files written before this key existed / by the reference: nothing to restore, not an error
    assert restore_step_state(st, {k: v for k, v in meta.items() if k != "step_state"}) is False
    # the reference's own loader reads the file (unknown meta keys are ignored) -- build container only
    if os.path.isdir("/root/reference/src/pyflow"):
        code = ("import sys; from pyflow.io.checkpoint import load_checkpoint; st, m = load_checkpoint(sys.argv[1]); "
                "print(m['iteration'], m['schema_version'], sorted(st.fields), st.fields['u'].shape)")
        r = subprocess.run([sys.executable, "-c", code, ck], capture_output=True, text=True,
                           env=dict(os.environ, PYTHONPATH="/root/reference/src"), timeout=300)
        assert r.returncode == 0 and r.stdout.split()[:2] == [str(cut - 1), "1"], r.stdout + r.stderr


def test_cli_exact_restart_flag(slab_cls, tmp_path, capsys, monkeypatch):
    from pyflow import cli
    ck = str(tmp_path / "ck.npz")
    common = ["--nx", "16", "--ny", "12", "--json-stream", "--checkpoint", ck, "--checkpoint-interval", "1"]

    def dts(argv):
        assert cli.main(argv) == 0
        return [json.loads(line)["dt"] for line in capsys.readouterr().out.splitlines() if line.startswith("{")]

    full = dts(["run", "--steps", "8", *common])
    assert len(dts(["run", "--steps", "5", *common])) == 5
    assert dts(["run", "--steps", "8", "--restart", ck, "--exact-restart", *common]) == full[5:]
    dts(["run", "--steps", "5", *common])                            # checkpoint of iteration 4 again
    assert dts(["run", "--steps", "8", "--restart", ck, *common]) != full[5:]       # reference behaviour: fresh ramp
    dts(["run", "--steps", "5", *common])
    monkeypatch.setenv("PYFLOW_B200_EXACT_RESTART", "1")
    assert dts(["run", "--steps", "8", "--restart", ck, *common]) == full[5:]


# ------------------------------------------------------------------ alternate dict-payload driver
def test_alternate_driver_runs_on_the_device_statistics(slab_cls):
    import pyflow.driver.simulation_driver as alt
    cfg = alt.DriverConfig(nx=16, ny=16, max_steps=4, report_interval=1000)
    drv = alt.SimulationDriver(cfg, alt.PhysicsConfig(lin_maxiter=200))
    payloads = list(drv.run())
    assert len(payloads) == 4 and all(p["state"] is drv.state for p in payloads)
    slab = drv.state.fields._slab
    m, n = slab.divergence_stats(drv.state.meta["dx"], drv.state.meta["dy"])
    assert payloads[-1]["divergence_mean"] == m and payloads[-1]["divergence_mean_free_norm"] == n
    assert slab.downloads == []                            # the statistics are device reductions, no field D2H


# ------------------------------------------------------------------ array-level operator seams
def test_operator_seams_signatures_and_shapes(slab_cls):
    from pyflow.linear_solvers.interface import LinearSolveResult, NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    from pyflow.numerics import fluid_ops
    from pyflow.numerics.core_ops_coherent import divergence_coherent, gradient_coherent, laplacian_coherent
    from pyflow.numerics.operators.advection import advect_quick, advect_upwind
    rng = np.random.default_rng(5)
    p, u, v = (rng.standard_normal((9, 13)) for _ in range(3))
    dx, dy = 0.5, 0.25
    gx, gy = gradient_coherent(p, dx, dy)
    assert np.array_equal(laplacian_coherent(p, dx, dy), divergence_coherent(gx, gy, dx, dy))
    assert np.array_equal(fluid_ops.divergence(u, v, dx, dy), orc.divergence_central(u, v, dx, dy))
    cgx, cgy = fluid_ops.gradient(p, dx, dy)
    ogx, ogy = orc.gradient_central(p, dx, dy)
    assert np.array_equal(cgx, ogx) and np.array_equal(cgy, ogy)
    assert np.array_equal(fluid_ops.laplacian(p, dx, dy), orc.laplacian_central(p, dx, dy))
    assert np.array_equal(advect_quick(u, v, p, dx, dy), orc.advect_quick(u, v, p, dx, dy))
    assert np.array_equal(advect_upwind(u, v, p, dx, dy), orc.advect_upwind(u, v, p, dx, dy))
    from pyflow.core.ghost_fields import allocate_state, interior_view
    st = allocate_state(13, 9)
    interior_view(st.fields["u"])[:] = u
    interior_view(st.fields["v"])[:] = v
    fluid_ops.apply_pressure_correction(st, p, 0.01, dx, dy)
    ou, ov = orc.pressure_correction_central(u, v, p, 0.01, dx, dy)
    assert np.array_equal(interior_view(st.fields["u"]), ou) and np.array_equal(interior_view(st.fields["v"]), ov)
    assert np.array_equal(interior_view(st.fields["p"]), p)
    lg = fluid_ops.legacy_grad_p_state(st, dx, dy)
    assert np.array_equal(lg[0], ogx) and np.array_equal(fluid_ops.legacy_divergence_state(st, dx, dy),
                                                         orc.divergence_central(ou, ov, dx, dy))
    with pytest.raises(ValueError):
        fluid_ops.gradient(np.zeros((1, 5)), 1.0, 1.0)
    # linear-solve seam: x0 = 0, scipy semantics of converged / iterations
    A = NegLaplacianPinned(13, 9, dx, dy)
    b = rng.standard_normal(13 * 9)
    b -= b.mean()
    b[0] = 0.0
    r = solve(A, b, method="cg", tol=1e-10, maxiter=2000, M=jacobi_preconditioner(A))
    assert isinstance(r, LinearSolveResult) and r.converged and r.iterations == 0 and r.method == "cg"
    assert np.linalg.norm(A.matvec(r.x) - b) <= 1e-9 * np.linalg.norm(b) and r.x.shape == b.shape
    r2 = solve(A, b, method="cg", tol=1e-10, maxiter=3)
    assert (not r2.converged) and r2.iterations == 3
    r3 = solve(A, b, tol=1e-10, maxiter=2000)                   # the signature's default method (interface.py:23)
    assert r3.method == "bicgstab" and r3.converged and r3.iterations == 0
    assert np.linalg.norm(A.matvec(r3.x) - b) <= 1e-10 * np.linalg.norm(b)
    r4 = solve(A, b, method="bicgstab", tol=1e-10, maxiter=4, M=jacobi_preconditioner(A))
    assert (not r4.converged) and r4.iterations == 4
    with pytest.raises(NotImplementedError):
        solve(np.eye(4), np.ones(4), method="cg")


# ------------------------------------------------------------------ benchmark sweep (reference-shaped harness)
def test_benchmark_harness_records(slab_cls):
    from pyflow.benchmark.harness import benchmark, make_config, run_case
    res = benchmark([8, 12], steps=5)
    assert res["benchmark"] == "lid_driven_cavity" and res["steps_requested"] == 5 and len(res["runs"]) == 2
    for r in res["runs"]:
        assert {"nx", "ny", "steps", "total_time_s", "avg_time_per_step_s", "avg_cg_iterations"} <= set(r)
        assert r["steps"] == 5 and r["avg_time_per_step_s"] > 0.0
        assert r["avg_true_cg_iterations"] > 0 and r["mlups"] > 0 and r["cg_iterations_per_s"] > 0   # a driven cavity
    # the reference's own harness configuration is still reachable: nothing moves, no CG work
    rest = run_case(8, 8, 3, make_config(8, 8, lid_velocity=0.0, disable_advection=True))
    assert rest["avg_cg_iterations"] == 0 and rest["avg_true_cg_iterations"] == 0


def test_cli_keyboard_interrupt_writes_final_checkpoint_and_returns_130(slab_cls, tmp_path, monkeypatch, capsys):
    """Reference cli.py:179-193: Ctrl-C during the run attempts a final checkpoint of the last yielded state."""
    from pyflow import cli
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.io.checkpoint import load_checkpoint
    real_run = SimulationDriver.run

    def interrupted_run(self, *a, **k):
        for n, item in enumerate(real_run(self, *a, **k)):
            yield item
            if n == 2:
                raise KeyboardInterrupt

    monkeypatch.setattr(SimulationDriver, "run", interrupted_run)
    ck = str(tmp_path / "final.npz")
    rc = cli.main(["run", "--nx", "12", "--ny", "10", "--steps", "50", "--checkpoint", ck])
    assert rc == 130
    out = capsys.readouterr().out
    assert "KeyboardInterrupt received" in out and "Final checkpoint written" in out
    st, meta = load_checkpoint(ck)
    assert meta["iteration"] == 2
    ost, _ = orc.run(make_cfg(nx=12, ny=10, Re=100.0), 3)
    assert np.array_equal(st.fields["u"], ost.fields["u"]) and np.array_equal(st.fields["p"], ost.fields["p"])
    # --json-stream keeps stdout machine-readable even then
    rc = cli.main(["run", "--nx", "12", "--ny", "10", "--steps", "50", "--checkpoint", ck, "--json-stream"])
    lines = capsys.readouterr().out.splitlines()
    assert rc == 130 and len(lines) == 3 and all(line.startswith("{") for line in lines)


def test_reference_recipe_for_the_jacobi_preconditioner_is_recognised(slab_cls):
    """pressure_solver.py:139-143 builds M as jacobi_preconditioner(csr(assemble_negative_laplacian(...))): that
    object (a diagonal read off the assembled matrix) is accepted when it IS the ring-identity diagonal of the grid
    being solved.  Foreign operator objects are probed for diagonality; general diagonals run through the array-form
    kernels of the linear-solve seam; operators that are not diagonal are refused -- there is no host path."""
    import scipy.sparse as sp
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.linear_solvers.interface import NegLaplacianPinned, solve
    from pyflow.linear_solvers.preconditioners import DiagonalJacobi, RingJacobi, jacobi_preconditioner
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian, solve_pressure_poisson
    nx, ny, dx, dy = 18, 14, 0.5, 0.25
    A = assemble_negative_laplacian(nx, ny, dx, dy)
    M = jacobi_preconditioner(sp.csr_matrix(A))
    assert isinstance(M, DiagonalJacobi) and M.shape == (nx * ny, nx * ny)
    ring = M.as_ring(nx, ny)
    assert isinstance(ring, RingJacobi) and ring.inv_diag_interior == RingJacobi(nx, ny, dx, dy).inv_diag_interior
    assert M.as_ring(ny, nx) is None and M.as_ring(nx, ny + 1) is None
    assert np.array_equal(M.inv_diag, orc.jacobi_inverse_diagonal(nx, ny, dx, dy))
    cfg = SimpleNamespace(lin_tol=1e-10, lin_maxiter=300, diagnostics=False, enable_jacobi_pc=False)

    def _seeded(nx_, ny_):
        st_ = allocate_state(nx_, ny_)
        interior_view(st_.fields["u"])[ny_ // 3, nx_ // 3] = 1.0
        return st_

    results = []
    for pre in (M, None):
        st = _seeded(nx, ny)
        c = cfg if pre is not None else SimpleNamespace(**{**vars(cfg), "enable_jacobi_pc": True})
        results.append(solve_pressure_poisson(st, 0.01, dx, dy, c, preconditioner=pre)[0])
    assert np.array_equal(results[0], results[1])
    # the linear-solve seam takes it too
    op = NegLaplacianPinned(nx, ny, dx, dy)
    b = np.random.default_rng(2).standard_normal(nx * ny)
    b -= b.mean()
    b[0] = 0.0
    r1 = solve(op, b, method="cg", tol=1e-10, maxiter=3000, M=M)
    r2 = solve(op, b, method="cg", tol=1e-10, maxiter=3000, M=jacobi_preconditioner(op))
    assert r1.converged and np.array_equal(r1.x, r2.x)
    # anything else is refused
    other = DiagonalJacobi(np.linspace(0.1, 1.0, nx * ny))
    with pytest.raises(NotImplementedError):
        solve_pressure_poisson(allocate_state(nx, ny), 0.01, dx, dy, cfg, preconditioner=other)
    with pytest.raises(NotImplementedError):
        jacobi_preconditioner(lambda x: x)
    # ---- user-supplied operators ("accepts external M LinearOperator") are probed for diagonality
    import scipy.sparse.linalg as spla
    from pyflow.linear_solvers.preconditioners import RingJacobi as Ring, classify, probe_diagonal
    n = nx * ny
    invd = orc.jacobi_inverse_diagonal(nx, ny, dx, dy)
    ref_style = spla.LinearOperator((n, n), matvec=lambda x: invd * x)        # reference preconditioners.py:20-25
    assert isinstance(classify(ref_style, nx, ny), Ring)
    r3 = solve(op, b, method="cg", tol=1e-10, maxiter=3000, M=ref_style)
    assert np.array_equal(r3.x, r2.x)                                         # ring diagonal -> the tuned CG kernels
    p3 = solve_pressure_poisson(_seeded(nx, ny), 0.01, dx, dy, cfg, preconditioner=ref_style)[0]
    assert np.array_equal(p3, results[0])
    # a general diagonal runs through the array-form kernels of the seam (both methods)
    gen = spla.LinearOperator((n, n), matvec=lambda x: other.inv_diag * x)
    for method, krylov in (("cg", orc.pcg), ("bicgstab", orc.bicgstab)):
        for Mx in (other, gen, jacobi_preconditioner(sp.diags(1.0 / other.inv_diag, format="csr"))):
            rg = solve(op, b, method=method, tol=1e-10, maxiter=9, M=Mx)
            xo, info, _ = krylov(lambda t: orc.neg_laplacian_pinned(t, nx, ny, dx, dy), b.copy(), 1e-10, 9,
                                 probe_diagonal(gen, n).inv_diag)
            assert rg.iterations == info == 9 and np.linalg.norm(rg.x - xo) <= 1e-12 * np.linalg.norm(xo)
    # operators that are not diagonal have no device form
    coupled = sp.diags([np.full(n - 1, 0.1), np.ones(n)], [1, 0], format="csr")
    for bad in (spla.aslinearoperator(coupled), spla.LinearOperator((n, n), matvec=lambda x: np.roll(x, 1)),
                spla.LinearOperator((n - 1, n - 1), matvec=lambda x: x)):
        assert probe_diagonal(bad, n) is None
        with pytest.raises(NotImplementedError):
            solve(op, b, method="cg", M=bad)
        with pytest.raises(NotImplementedError):
            solve_pressure_poisson(allocate_state(nx, ny), 0.01, dx, dy, cfg, preconditioner=bad)
    assert jacobi_preconditioner(sp.identity(4, format="csr")).as_ring(2, 2).inv_diag_interior == 0.25


# ------------------------------------------------------------------ linear_solvers.solve on the assembled matrix
def test_solve_recognises_the_assembled_matrix_of_the_reference(slab_cls):
    """``solve(A, b)`` with the sparse matrix ``assemble_negative_laplacian`` returns (reference
    pressure_solver.py:15-40): recognised from its CSR arrays (identity ring rows + one 5-point interior row), routed
    to the device entry point with the stencil coefficients; anything else is refused -- no CPU fallback."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from pyflow.linear_solvers.interface import Ring5Matrix, recognise_ring5, solve
    from pyflow.linear_solvers.preconditioners import jacobi_preconditioner
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian
    nx, ny, dx, dy = 12, 9, 0.5, 0.25
    A = assemble_negative_laplacian(nx, ny, dx, dy)
    found = recognise_ring5(A)
    assert found == Ring5Matrix(nx, ny, (-1.0 / dy ** 2, -1.0 / dx ** 2, 2.0 / dx ** 2 + 2.0 / dy ** 2, -1.0 / dx ** 2, -1.0 / dy ** 2))
    assert recognise_ring5(A.tolil()) == found and recognise_ring5(sp.csc_matrix(A)) == found      # any sparse format
    b = np.random.default_rng(5).standard_normal(nx * ny)
    # the matrix is NOT symmetric (interior rows couple to ring cells, ring rows are identity rows): CG on it is what
    # the reference does when asked, so it is compared truncated; BiCGSTAB converges
    for method, ref, maxiter in (("cg", spla.cg, 25), ("bicgstab", spla.bicgstab, 400)):
        got = solve(A, b, method=method, tol=1e-10, maxiter=maxiter, M=jacobi_preconditioner(A))
        D = A.diagonal()
        x_ref, info = ref(A, b, rtol=1e-10, maxiter=maxiter, M=spla.LinearOperator(A.shape, matvec=lambda v: v / D))
        assert got.iterations == (info if info > 0 else 0) and (method == "cg" or (info == 0 and got.converged))
        assert np.linalg.norm(got.x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    bad = A.tolil()
    bad[nx + 1, nx + 2] = -3.0
    for M in (bad.tocsr(), sp.identity(nx * ny).tocsr(), sp.random(nx * ny, nx * ny, 0.05, random_state=1).tocsr()):
        assert recognise_ring5(M) is None
        with pytest.raises(NotImplementedError):
            solve(M, b, method="cg")
