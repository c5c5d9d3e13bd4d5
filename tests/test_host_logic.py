"""CPU-only tests: the C-ABI library loads and exports everything the header
declares (no compute calls), and the host-side mirror of the reference's
interface (config / hash, checkpoint layout, State synchronisation logic, CLI,
logger, tracker) behaves like the reference's."""
import io
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------ C-ABI
def test_library_exports_every_symbol_of_the_header():
    from pyflow._b200 import lib
    hdr = open(os.path.join(ROOT, "include", "pyflow_b200.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pf_status"}
    assert declared, "no declarations parsed"
    handle = lib.load()
    missing = [name for name in sorted(declared) if not hasattr(handle, name)]
    assert not missing, f"header declares symbols the library lacks: {missing}"
    assert set(lib.EXPORTS) == declared, (set(lib.EXPORTS) ^ declared)
    assert handle.pf_abi_version() == lib.ABI_VERSION
    assert handle.pf_strerror(0) == b"ok" and b"argument" in handle.pf_strerror(-1)


def test_struct_layouts_match_the_header():
    import ctypes as C
    from pyflow._b200 import lib
    assert C.sizeof(lib.StepParams) == 6 * 8 + 6 * 4
    assert C.sizeof(lib.StepOut) == 13 * 8 + 4 * 4
    hdr = open(os.path.join(ROOT, "include", "pyflow_b200.h")).read()
    body = hdr[hdr.index("typedef struct pf_step_out"):hdr.index("} pf_step_out;")]
    names = re.findall(r"\b(?:double|int32_t)\s+([a-z0-9_, ]+);", body)
    flat = [n.strip() for grp in names for n in grp.split(",")]
    assert flat == [f[0] for f in lib.StepOut._fields_]
    body = hdr[hdr.index("typedef struct pf_step_params"):hdr.index("} pf_step_params;")]
    names = re.findall(r"\b(?:double|int32_t)\s+([a-z0-9_, ]+);", body)
    flat = [n.strip() for grp in names for n in grp.split(",")]
    assert flat == [f[0] for f in lib.StepParams._fields_]


def test_workspace_size_accounts_for_all_fields():
    from pyflow._b200 import lib
    h = lib.load()
    n = h.pf_workspace_bytes(4096, 4096)
    assert n >= 9 * 4100 * 4096 * 8 and n < 10 * 4100 * 4096 * 8
    assert h.pf_workspace_bytes(1, 5) == 0            # degenerate grids are refused
    assert h.pf_workspace_bytes(16384, 16384) < 25 * 2 ** 30   # config 5 slab fits a 180 GB B200 many times


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_fails_loudly():
    from pyflow._b200.lib import NativeLibraryError
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.numerics.core_ops_coherent import laplacian_coherent
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    from types import SimpleNamespace
    with pytest.raises(NativeLibraryError):
        step(SimpleNamespace(nx=8, ny=8, diagnostics=False), allocate_state(8, 8), ResidualManager(), 0)
    with pytest.raises(NativeLibraryError):
        laplacian_coherent(np.ones((8, 8)), 1.0, 1.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyflow-solver_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyflow_oracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f
                # ... nor the test doubles (oracle-backed stand-in slab, host emulation of the kernels)
                assert "cpu_slab" not in src and "emu_kernels" not in src, f
                # (the only hooks for tests are the #ifdef PF_HOST_EMULATION ... #endif blocks around PTX helpers in
                # pf_common.cuh; never defined in the product build)
                # four #ifdef blocks: loads / stores, LL words, fences, the out-of-line copy of the IEEE division
                # (a __noinline__ device function there, a plain inline one on the host)
                assert src.count("PF_HOST_EMULATION") == (6 if f == "pf_common.cuh" else 0), f


# ------------------------------------------------------------------ config
def test_config_hash_known_answers(golden):
    from pyflow.config.model import SimulationConfig
    for key, want in golden["config"]["config_hash"].items():
        assert SimulationConfig(**json.loads(key)).config_hash == want
    # SURVEY section 5 KATs (CLI defaults => diagnostics False)
    assert SimulationConfig(nx=64, ny=64, advection_scheme="quick", lin_maxiter=300,
                            diagnostics=False).config_hash == "8dddaac6b3f9"
    assert SimulationConfig(nx=4096, ny=4096, Re=1000.0, advection_scheme="quick", lin_maxiter=300,
                            diagnostics=False).config_hash == "4de9eeab2d1d"


def test_config_hash_ignores_runtime_fields_and_validates_ranges():
    from pydantic import ValidationError
    from pyflow.config.model import ConfigError, SimulationConfig
    a = SimulationConfig(nx=16, ny=16)
    b = SimulationConfig(nx=16, ny=16, log_stream=io.StringIO(), force_quiet=True)
    assert a.config_hash == b.config_hash and len(a.config_hash) == 12
    assert a.config_hash != SimulationConfig(nx=16, ny=16, Re=200.0).config_hash
    for bad in (dict(nx=2, ny=16), dict(nx=16, ny=16, Re=0.0), dict(nx=16, ny=16, cfl_target=1.5),
                dict(nx=16, ny=16, cfl_growth=1.0), dict(nx=16, ny=16, lin_maxiter=10001),
                dict(nx=16, ny=16, lin_tol=1e-15), dict(nx=16, ny=16, advection_scheme="weno")):
        with pytest.raises(ValidationError):
            SimulationConfig(**bad)
    with pytest.raises((ConfigError, ValidationError)):
        SimulationConfig(nx=16, ny=16, log_path="x.jsonl", log_stream=io.StringIO())
    assert SimulationConfig(nx=16, ny=16).brief()["scheme"] == "upwind"
    assert SimulationConfig(nx=16, ny=16, disable_advection=True).brief()["scheme"] is None


# ------------------------------------------------------------------ checkpoint layout
def test_checkpoint_layout_and_roundtrip(tmp_path, golden):
    from pyflow.config.model import SimulationConfig
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.io.checkpoint import (load_checkpoint, load_checkpoint_bytes, save_checkpoint,
                                      save_checkpoint_bytes, _hash_array)
    cfg = SimulationConfig(nx=10, ny=12, seed=3)
    st = allocate_state(10, 12)
    rng = np.random.default_rng(0)
    for k in ("u", "v", "p"):
        st.fields[k][:] = rng.standard_normal((14, 12))
    path = str(tmp_path / "sub" / "ck.npz")
    assert save_checkpoint(path, st, 7, 1.25, cfg) == path
    with np.load(path) as z:
        assert sorted(z.files) == ["field_p", "field_u", "field_v", "meta"]
        assert z["field_u"].dtype == np.float64 and z["field_u"].shape == (14, 12)
        meta = json.loads(str(z["meta"]))
    exemplar = golden["config"]["checkpoint_exemplar"]
    assert sorted(meta.keys()) == exemplar["meta_keys"]            # same keys as the reference's shipped .npz
    assert meta["iteration"] == 7 and meta["time"] == 1.25 and meta["nx"] == 10 and meta["ny"] == 12
    assert meta["structured_config_hash"] == cfg.config_hash and meta["seed"] == 3 and meta["schema_version"] == 1
    assert len(meta["config_hash"]) == 10
    st2, meta2 = load_checkpoint(path)
    for k in ("u", "v", "p"):
        assert np.array_equal(st2.fields[k], st.fields[k])
        assert meta2["field_hashes"][k] == _hash_array(st.fields[k])
    assert st2.meta == {}                                           # meta is not persisted (as in the reference)
    st3, meta3 = load_checkpoint_bytes(save_checkpoint_bytes(st, 7, 1.25, cfg))
    assert np.array_equal(st3.fields["p"], st.fields["p"]) and "structured_config_hash" not in meta3
    assert not [f for f in os.listdir(tmp_path / "sub") if f.endswith(".tmp")]


# ------------------------------------------------------------------ State <-> device synchronisation logic
class FakeSlab:
    """Stands in for DeviceSlab: remembers uploads, serves downloads."""

    def __init__(self, nx, ny):
        self.nx, self.ny_local = nx, ny
        self.dev = {k: np.zeros((ny + 2, nx + 2)) for k in ("u", "v", "p")}
        self.uploads, self.downloads = [], []

    def upload(self, **kw):
        for k, a in kw.items():
            self.dev[k][1:-1, 1:-1] = a[1:-1, 1:-1]
            self.uploads.append(k)

    def download(self, **kw):
        for k, a in kw.items():
            a[1:-1, 1:-1] = self.dev[k][1:-1, 1:-1]
            self.downloads.append(k)


def test_synced_fields_lazy_protocol():
    from pyflow.core.ghost_fields import State, SyncedFields, allocate_state, interior_view
    st = allocate_state(6, 5)
    f = st.fields
    assert isinstance(f, SyncedFields) and f.host_touched() == {"u", "v", "p"}
    slab = FakeSlab(6, 5)
    f._slab = slab
    u_arr = f.peek("u")
    # "device step": device copy changes, host copy goes stale
    slab.dev["u"][1:-1, 1:-1] = 7.0
    f.mark_device_updated()
    assert f.host_touched() == set() and not slab.downloads
    assert f.read("u") is u_arr and np.all(interior_view(u_arr) == 7.0)      # read-only accessor: one D2H
    assert slab.downloads == ["u"] and f.host_touched() == set()
    f.read("u")
    assert slab.downloads == ["u"]                                            # cached until the next step
    v = st.fields["v"]                                                        # normal accessor marks host-touched
    assert slab.downloads == ["u", "v"] and f.host_touched() == {"v"}
    assert st["v"] is v and st.get("v") is v and st.get("nope", 5) == 5
    st.fields["p"] = np.ones((7, 8))
    assert "p" in f.host_touched() and "p" not in f._stale
    assert sorted(st.keys()) == ["p", "u", "v"] and len(st.items()) == 3
    plain = State(nx=2, ny=2, fields={"u": np.zeros((4, 4))})
    assert isinstance(plain.fields, SyncedFields)


def test_driver_requires_state_or_grid():
    from types import SimpleNamespace
    from pyflow.drivers.simulation_driver import SimulationDriver
    with pytest.raises(ValueError):
        next(SimulationDriver(SimpleNamespace()).run(max_steps=1))
    assert list(SimulationDriver(SimpleNamespace(nx=8, ny=8)).run(max_steps=0)) == []      # absolute bound
    assert list(SimulationDriver(SimpleNamespace(nx=8, ny=8)).run(max_steps=3, start_iteration=5)) == []


def test_rp_diagnostics_formula():
    """interface.py:52 + pressure_solver.py:190-194: 0 / 1 / maxiter, never a count."""
    from types import SimpleNamespace
    from pyflow._b200.lib import StepOut
    from pyflow.solvers.pressure_solver import _rp_diagnostics
    cfg = SimpleNamespace(lin_tol=1e-10)
    o = StepOut(); o.info = 0; o.rnorm = 5e-9; o.bnorm = 100.0; o.true_iters = 280
    d = _rp_diagnostics(cfg, o)
    assert d["Rp_iterations"] == 1 and d["Rp_converged"] is True and d["Rp_method"] == "cg"
    o.info = 300; o.rnorm = 3.0
    d = _rp_diagnostics(cfg, o)
    assert d["Rp_iterations"] == 300 and d["Rp_converged"] is False
    o.info = 0; o.rnorm = 0.0; o.bnorm = 0.0
    d = _rp_diagnostics(cfg, o)
    assert d["Rp_iterations"] == 0 and d["Rp_converged"] is True


def test_assemble_negative_laplacian_rows():
    """Reference contract (tests/test_projection_method_contract.py:1-22): interior rows
    sum to zero, boundary-ring rows are identity."""
    from pyflow.solvers.pressure_solver import assemble_negative_laplacian
    A = assemble_negative_laplacian(7, 6, 0.5, 0.25).toarray()
    d = np.diag(A).reshape(6, 7)
    assert np.all(d[0] == 1) and np.all(d[-1] == 1) and np.all(d[:, 0] == 1) and np.all(d[:, -1] == 1)
    assert np.allclose(d[1:-1, 1:-1], 2 / 0.25 + 2 / 0.0625)
    sums = A.sum(axis=1).reshape(6, 7)
    assert np.allclose(sums[1:-1, 1:-1], 0.0) and np.allclose(sums[0], 1.0)


# ------------------------------------------------------------------ logger / tracker
def test_jsonl_logger_and_tracker():
    from pyflow.logging.structured import JsonlLogger
    from pyflow.residuals.manager import ResidualManager
    buf = io.StringIO()
    lg = JsonlLogger(stream=buf)
    lg.log({"type": "step", "it": 1})
    rec = json.loads(buf.getvalue())
    assert rec["type"] == "step" and "ts" in rec and ", " not in buf.getvalue()
    with pytest.raises(ValueError):
        JsonlLogger()
    tr = ResidualManager()
    tr.add("Ru", 0.5); tr.add("Rv", 0.1)
    assert tr.series["Ru"].values == [1.0, 0.5] and tr.plateau_detect("Ru") is False


# ------------------------------------------------------------------ CLI (no compute)
def _cli(*argv):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "pyflow-solver_b200"))
    return subprocess.run([sys.executable, "-m", "pyflow", *argv], capture_output=True, text=True, env=env)


def test_cli_validate_show_config_and_exit_codes(tmp_path):
    r = _cli("validate", "--nx", "64", "--ny", "64")
    assert r.returncode == 0
    out = json.loads(r.stdout)
    assert out["config_hash"] == "8dddaac6b3f9" and out["lin_maxiter"] == 300 and out["advection_scheme"] == "quick"
    r = _cli("show-config", "--nx", "4096", "--ny", "4096", "--re", "1000")
    assert r.returncode == 0 and json.loads(r.stdout)["hash"] == "4de9eeab2d1d"
    r = _cli("run", "--nx", "2", "--steps", "1")
    assert r.returncode == 2 and "Configuration invalid" in r.stdout
    assert _cli().returncode != 0                                   # sub-command is required
    # restart hash enforcement happens before any compute: exit code 3
    from pyflow.config.model import SimulationConfig
    from pyflow.core.ghost_fields import allocate_state
    from pyflow.io.checkpoint import save_checkpoint
    ck = str(tmp_path / "ck.npz")
    save_checkpoint(ck, allocate_state(16, 16), 4, 0.0, SimulationConfig(nx=16, ny=16, Re=50.0))
    r = _cli("run", "--nx", "16", "--ny", "16", "--steps", "0", "--restart", ck)
    assert r.returncode == 3 and "Refusing restart" in r.stdout
    r = _cli("run", "--nx", "16", "--ny", "16", "--steps", "0", "--restart", ck, "--allow-hash-mismatch", "--json-stream")
    assert r.returncode == 0 and r.stdout == ""


def test_cli_step_payload_schema():
    from pyflow.cli import _step_payload
    from pyflow.config.model import SimulationConfig
    diag = {"iteration": 3, "dt": 0.01, "CFL": 0.4, "wall_time": 1.0, "plateau_Ru": False, "max_residual": 0.1,
            "residual_drop_Ru_orders": 0.0, "nu": 0.01, "Rp_iterations": 300, "Rp_residual": 1.0,
            "Rp_converged": False, "Rp_method": "cg", "divergence_norm": 0.1, "nan_detected": False,
            "u_centerline": [[0.0], [0.0]]}
    res = {"Ru": 0.1, "Rv": 0.05, "Rp": 0.0, "continuity": 1e-3}
    p = _step_payload(SimulationConfig(nx=8, ny=8, seed=5), res, diag)
    assert list(p.keys()) == ["type", "iteration", "dt", "CFL", "continuity", "wall_time", "residuals",
                              "diagnostics", "run_meta"]
    assert p["type"] == "step" and p["run_meta"] == {"seed": 5} and p["continuity"] == 1e-3
    assert set(p["diagnostics"]) == set(diag) - {"iteration", "dt", "CFL", "wall_time"}
    assert "run_meta" not in _step_payload(SimulationConfig(nx=8, ny=8), res, diag)


def test_bench_reference_arm_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--ref-grid", "48"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["nx"] == 4096 and line["higher_is_better"] is True


# ------------------------------------------------------------------ alternate (dict-payload) driver, host logic
def test_alternate_driver_steady_detection_and_hooks(monkeypatch):
    """Reference driver/simulation_driver.py:113-176 semantics with the solver step mocked out."""
    import pyflow.driver.simulation_driver as alt
    calls = {"pre": 0, "post": 0}

    def fake_step(cfg, state, tracker, it):
        ru = 1e-3 * (0.5 ** it) if it < 6 else 1e-5          # decays, then flat => slope ~ 0
        tracker.add("Ru", ru)
        state.meta.update(dx=1.0, dy=1.0)
        return state, {"Ru": ru, "Rv": ru, "Rp": 0.0, "continuity": 0.0}, {"dt": 0.01, "CFL": 0.5}

    monkeypatch.setattr(alt, "core_step", fake_step)
    monkeypatch.setattr(alt.SimulationDriver, "_divergence_stats", lambda self, st: (0.25, 0.5))
    cfg = alt.DriverConfig(nx=8, ny=8, max_steps=40, steady_window=5, steady_slope_tol=1e-3, report_interval=1000,
                           pre_step=lambda d, it, st: calls.__setitem__("pre", calls["pre"] + 1),
                           post_step=lambda d, it, st: calls.__setitem__("post", calls["post"] + 1))
    drv = alt.SimulationDriver(cfg, alt.PhysicsConfig())
    payloads = list(drv.run())
    last = payloads[-1]
    assert set(last) == {"iteration", "state", "residuals", "diagnostics", "divergence_mean",
                         "divergence_mean_free_norm", "steady", "steady_slope", "elapsed"}
    assert last["steady"] is True and drv.steady_detected and drv.steady_iteration == last["iteration"] < 39
    assert all(p["state"] is drv.state for p in payloads) and not any(p["steady"] for p in payloads[:-1])
    assert calls["pre"] == len(payloads) and calls["post"] == len(payloads) - 1     # no post hook on the final steady yield
    assert last["divergence_mean"] == 0.25 and last["divergence_mean_free_norm"] == 0.5
    # without stop_on_steady the loop runs to max_steps
    cfg2 = alt.DriverConfig(nx=8, ny=8, max_steps=12, steady_window=5, steady_slope_tol=1e-3, stop_on_steady=False,
                            report_interval=1000)
    assert len(list(alt.SimulationDriver(cfg2).run())) == 12
    assert alt.create_default_driver().driver_cfg.nx == 65 and alt.PhysicsConfig().lin_maxiter == 800


# ------------------------------------------------------------------ C-ABI conventions without a GPU
def test_c_abi_argument_validation_and_clean_failure_without_a_gpu():
    """0 / negative status codes, nothing throws or aborts across the ABI (include/pyflow_b200.h conventions):
    argument errors are refused before any CUDA call; without a device pf_create reports PF_ERR_CUDA."""
    import ctypes as C
    from pyflow._b200 import lib
    h = lib.load()
    ctx = C.c_void_p()
    fake = C.c_void_p(1 << 20)                                   # aligned, never dereferenced on these paths
    big = h.pf_workspace_bytes(64, 64)
    assert big > 9 * 68 * 64 * 8
    assert h.pf_create(None, 64, 64, 64, 0, 0, fake, big) == -1                       # no out pointer
    assert h.pf_create(C.byref(ctx), 1, 64, 64, 0, 0, fake, big) == -1                # nx < 2
    assert h.pf_create(C.byref(ctx), 64, 64, 1, 0, 0, fake, big) == -1                # ny_global < 2
    assert h.pf_create(C.byref(ctx), 64, 40, 64, 30, 0, fake, big) == -1              # slab sticks out of the grid
    assert h.pf_create(C.byref(ctx), 64, 64, 64, 0, 0, None, big) == -1               # no workspace
    assert h.pf_create(C.byref(ctx), 64, 64, 64, 0, 0, C.c_void_p((1 << 20) + 8), big) == -1      # misaligned
    assert h.pf_create(C.byref(ctx), 64, 64, 64, 0, 0, fake, big - 1) == -1           # too small
    assert ctx.value is None
    if not _has_cuda():
        assert h.pf_create(C.byref(ctx), 64, 64, 64, 0, 0, fake, big) == -2 and ctx.value is None
    assert h.pf_destroy(None) == 0
    for call in (lambda: h.pf_step(None, None, None, None), lambda: h.pf_upload(None, None, None, None, 66, None),
                 lambda: h.pf_prestep_reduce(None, None, None), lambda: h.pf_centerline(None, None, None),
                 lambda: h.pf_launch_count(None, None), lambda: h.pf_set_option(None, b"cg_kernel", 1),
                 lambda: h.pf_comm_init(None, None, 0, 1), lambda: h.pf_comm_unique_id(None),
                 lambda: h.pf_op_laplacian_central(None, None, None, 1.0, 1.0, None)):
        assert call() == -1
    names = {code: h.pf_strerror(code).decode() for code in (0, -1, -2, -3, -4, -5, -99)}
    assert names[0] == "ok" and len(set(names.values())) == 7 and names[-99] == "unknown status"


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/pyflow_b200.h against the argtypes declared in pyflow/_b200/lib.py: same number
    of parameters, scalars of the same kind and width, pointers where the header has pointers."""
    import ctypes as C
    from pyflow._b200 import lib
    h = lib.load()
    hdr = open(os.path.join(ROOT, "include", "pyflow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char\*)\s+(pf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(lib.EXPORTS)
    scalar = {"double": C.c_double, "int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "int32_t": C.c_int32}
    pointer_like = (C.c_void_p, C.c_char_p)
    for name, params in protos:
        fn = getattr(h, name)
        plist = [p.strip() for p in params.replace("\n", " ").split(",") if p.strip() and p.strip() != "void"]
        if fn.argtypes is None:
            assert not plist, f"{name}: header has parameters, binding declares none"
            continue
        assert len(fn.argtypes) == len(plist), (name, plist, fn.argtypes)
        for decl, at in zip(plist, fn.argtypes):
            is_ptr = "*" in decl or "[" in decl
            if is_ptr:
                assert at in pointer_like or hasattr(at, "contents") or issubclass(at, C._Pointer), (name, decl, at)
            else:
                ctype = decl.replace("const ", "").split()[0]
                assert at is scalar[ctype], (name, decl, at)
