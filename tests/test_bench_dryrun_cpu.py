"""bench.py's GPU arm executed on the CPU with stand-ins (no number produced here means anything):
the oracle-backed slab replaces the GPU context and ``torch.cuda`` timing calls are stubbed, so that
every line of the measurement code runs once before it meets a real B200 -- contract keys, the e2e legs
(host arrays in / out, and the every-step-copies variant), launch counting, roofline bookkeeping."""
import importlib.util
import json
import os
import time
from types import SimpleNamespace

import pytest

import cpu_slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


def test_gpu_arm_runs_end_to_end_on_stand_ins(monkeypatch, capsys):
    import torch
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cpu_slab.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_a, **_k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_a, **_k: None)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setitem(bench.WORKLOADS, "tiny", dict(nx=24, ny=20, Re=100.0, scaling="weak", per_gpu=True))
    args = SimpleNamespace(gpus=1, steps=3, warmup=3, impl="b200", workload="tiny", no_cpu_baseline=True, ref_grid=0)
    bench.run_gpu_arm(args, bench.WORKLOADS["tiny"], 0, 1)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == "MLUPS per projection step" and line["unit"] == "MLUPS" and line["dtype"] == "f64"
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("lid-driven cavity 24x20") and "l2" in line["config"]
    assert line["gpu_launches"] == 3 * 9                                   # the stand-in counts 9 launches per step
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > e2e["h2d_bytes_per_step"] - 97
    every = e2e["every_step_copies"]
    assert "error" not in every, every
    assert every["h2d_bytes_per_step"] == 3 * 24 * 20 * 8 and every["steps"] == 3
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] == 24 * 20 * 64.0 * line["cg_iters_per_step"]
    assert r["kernel"].startswith("k_cg_small")
    assert r["traffic"] is None and "traffic_source" in r and len(r["kernel_source_hash"]) == 16
    par = line["parity"]                                   # untimed block: stand-in (numpy oracle) vs the C oracle
    assert par["ok"] and par["cases"] == 2 and par["dt_equal"] and par["max_iter_diff"] <= 1 and par["max_rel_l2"] < 1e-9
    assert [d["case"] for d in par["details"]] == ["tiny: 24x256 slab (256 rows per GPU), 300 its", "tiny: full 24x20 grid, 300 its"]
    assert bench.cg_kernel_name(4096, 4096, 1).startswith("k_cg_tmad") and bench.cg_kernel_name(513, 513, 1).startswith("k_cg_persistent")


@pytest.mark.parametrize("rank", [0, 1])
def test_reference_arm_prints_on_rank_zero_only(rank, capsys):
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    args = SimpleNamespace(gpus=2, steps=1, warmup=1, impl="reference", workload="cavity4096", no_cpu_baseline=False,
                           ref_grid=32, unmodified_steps=2)
    bench.run_reference_arm(args, bench.WORKLOADS["cavity4096"], rank)
    out = capsys.readouterr().out.strip()
    if rank:
        assert out == ""
    else:
        line = json.loads(out.splitlines()[-1])
        assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["e2e"]["value"] == line["value"]
        assert line["cpu_baseline"]["value"] == line["value"] and line["config"]["ny"] == 8192
        assert line["cpu_baseline"]["sample_grid"] == {"nx": 32, "ny": 32} and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
        # same `config` object as the GPU arm prints for this workload and GPU count
        assert line["config"] == bench.workload_config("cavity4096", bench.WORKLOADS["cavity4096"], 2, 300)
        unmod = line["cpu_baseline"]["unmodified_reference"]
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pyflow", "__init__.pfbc")):
            case = unmod["cases"]["64x64 Re=100"]
            assert case["steps_timed"] == 2 and case["MLUPS"] > 0 and "error" not in case
        else:
            assert "unavailable" in unmod


def test_reference_arm_uses_every_host_core_even_under_torchrun(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must not inherit that."""
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    import pyflow_oracle_c as orc_c
    orc_c.set_threads(1)
    assert bench.use_all_host_threads() == (os.cpu_count() or 1)


def test_reference_sample_keeps_the_row_length_and_fits_the_budget():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.sized_oracle_rows(256, 1024, 100.0, 20, 3, 1e9) == 1024          # fits: the whole slab
    assert bench.sized_oracle_rows(256, 1024, 100.0, 20, 3, 0.0) == 256           # never below 256 rows


def test_unmodified_reference_bytecode_agrees_with_the_oracle():
    """oracle/_ref (CPython bytecode of the unmodified reference, oracle/build_ref.py) stepped through its own
    SimulationDriver.run in a subprocess ends on the fields the numpy oracle ends on -- the pin of the oracle
    (oracle/make_golden.py) re-checked on whatever machine runs this suite, the GPU box included."""
    import subprocess
    import sys

    import numpy as np
    import build_ref
    import pyflow_oracle as orc
    if build_ref.build() is None:
        pytest.skip("no /root/reference and no prebuilt oracle/_ref")
    assert not [f for _d, _s, fs in os.walk(os.path.join(ROOT, "oracle", "_ref")) for f in fs
                if not (f.endswith(".pfbc") or f == "stamp.json")], "oracle/_ref must hold bytecode only, never reference sources"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_unmodified.py"), "40", "24", "100", "5"],
                       capture_output=True, text=True, cwd="/tmp", timeout=120,
                       env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    cfg = SimpleNamespace(nx=40, ny=24, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
    st, diags = orc.run(cfg, 6)
    assert got["final"]["dt"] == diags[-1]["dt"]
    for name in ("u", "v", "p"):
        want = float(np.linalg.norm(orc.interior(st.fields[name])))
        assert abs(got["final"][f"{name}_norm"] - want) <= 1e-12 * max(1.0, want), (name, got["final"], want)
