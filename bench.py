"""bench.py -- throughput of the PyFlow projection time step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): MLUPS per projection step = nx * ny_global * steps / seconds / 1e6.
A "step" is one full projection time step (predictor, RHS, <= lin_maxiter Jacobi-PCG
iterations, correction, lid BC, residual pass) of the lid-driven cavity started
from rest -- no RNG anywhere on this path, so the data are "synthetic" by
construction.

Workloads (BASELINE.json configs):
  cavity4096       4096 x 4096 rows PER GPU, Re=1000, QUICK, lin_maxiter=300  (config 3; default;
                   with N GPUs the global grid is 4096 x 4096N row slabs: weak scaling)
  cavity512        512 x 512, Re=1000 (config 2; shared-memory resident, latency-bound)
  strong8192       8192 x 8192 total, Re=5000, rows split over N GPUs (config 4; strong)
  weak16384        16384 x 16384 rows per GPU, Re=1000, lin_maxiter=300 (config 5 geometry; weak)
  weak16384_tight  the same slabs with the spec's tight-tolerance regime: lin_tol=1e-10 and
                   lin_maxiter=10000 (the model cap, reference config/model.py:28) -- 20 000
                   exchanges per step

One JSON line on stdout from rank 0.  `value` is device-resident stepping
through the public `step()`; `e2e` is the same metric through
`SimulationDriver.run` starting from HOST arrays and ending with the fields
back on the host, copies inside the timed region.  `roofline` is the CG kernel:
algorithmic bytes (64 B per cell per CG iteration) / its CUDA-event
duration, against MEASURED_PEAKS.json.  `cpu_baseline` (and `--impl reference`)
is the C / OpenMP restatement of the reference algorithm (oracle/pyflow_oracle.c)
on every host core of this box.  `parity` is an UNTIMED block after the timed
regions: the CUDA path against that oracle on the workload's own grid (one GPU)
and on slab cases at every GPU count; a failed parity check makes the exit code 1.
`other_configs` carries BASELINE configs 2, 4 and 5 (device-resident, few steps)
at the same GPU count, so that a 1/2/4/8 sweep of the default command records the
strong- and weak-scaling lines as well.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "pyflow-solver_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {
    # name: nx, ny (per GPU or total), Re, scaling
    "cavity4096": dict(nx=4096, ny=4096, Re=1000.0, scaling="weak", per_gpu=True),
    "cavity512": dict(nx=512, ny=512, Re=1000.0, scaling="weak", per_gpu=True),
    # config 1, the reference's own CPU-runnable default (`pyflow run`): the one the UNMODIFIED reference is timed on
    "cavity64": dict(nx=64, ny=64, Re=100.0, scaling="weak", per_gpu=True),
    "strong8192": dict(nx=8192, ny=8192, Re=5000.0, scaling="strong", per_gpu=False),
    "weak16384": dict(nx=16384, ny=16384, Re=1000.0, scaling="weak", per_gpu=True),
    "weak16384_tight": dict(nx=16384, ny=16384, Re=1000.0, scaling="weak", per_gpu=True, lin_maxiter=10000),
    # widths that are not a multiple of the TMA kernel's 1024-column chunk (clipped last chunk)
    "cavity4000": dict(nx=4000, ny=4000, Re=1000.0, scaling="weak", per_gpu=True),
    "cavity5000x3000": dict(nx=5000, ny=3000, Re=1000.0, scaling="weak", per_gpu=True),
    # general grid spacing (lx = ly = 1: dx = dy = 1/4095, the <false, ...> instantiations with the reference's IEEE
    # divides; no BASELINE config uses it, the reference's config model and CLI have no lx / ly: VERDICT r1 item 13)
    "cavity4096_lx1": dict(nx=4096, ny=4096, Re=1000.0, scaling="weak", per_gpu=True, lx=1.0, ly=1.0),
    # the slab ONE GPU owns in the 8-GPU strong-scaling run of config 4 (tile-geometry experiments on one GPU)
    "slab8192x1024": dict(nx=8192, ny=1024, Re=5000.0, scaling="weak", per_gpu=True),
}
CG_BYTES_PER_CELL_ITER = 64.0     # SURVEY.md 8d contract figure
STEP_BYTES_PER_CELL = 200.0


def make_cfg(nx, ny, Re, lin_maxiter=300, lin_tol=1e-10, lx=None, ly=None):
    cfg = SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05,
                          lin_tol=lin_tol, lin_maxiter=lin_maxiter, advection_scheme="quick",
                          disable_advection=False, diagnostics=False)
    if lx is not None:          # absent attributes = the reference's default lx = nx - 1, ly = ny - 1 (solver.py:19-22)
        cfg.lx, cfg.ly = float(lx), float(ly if ly is not None else lx)
    return cfg


def lin_maxiter_of(args, wl):
    return int(getattr(args, "lin_maxiter", 0) or wl.get("lin_maxiter", 300))


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/pf_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.fh.close()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash() -> str:
    """sha256 over the CUDA / C-ABI sources the library is built from (csrc/*, include/*.h, build.py).  Recorded
    next to every ncu traffic figure in profiles/cg_traffic.json, so that bench.py only quotes a `traffic` number
    that was captured from the kernels it is timing."""
    h = hashlib.sha256()
    files = [os.path.join(ROOT, "include", "pyflow_b200.h"), os.path.join(ROOT, "pyflow-solver_b200", "build.py")]
    csrc = os.path.join(ROOT, "pyflow-solver_b200", "csrc")
    files += [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def recorded_traffic(workload, kernel):
    """(bytes per launch or None, provenance note) from profiles/cg_traffic.json."""
    tpath = os.path.join(ROOT, "profiles", "cg_traffic.json")
    try:
        rec = json.load(open(tpath)).get(workload)
    except Exception:
        return None, "no profiles/cg_traffic.json"
    if not isinstance(rec, dict):
        return None, "no ncu capture recorded for this workload"
    now = kernel_source_hash()
    if rec.get("source_hash") != now:
        return None, (f"ncu capture in profiles/cg_traffic.json is of source hash {rec.get('source_hash')}, the running "
                      f"library is built from {now}: not quoted")
    if rec.get("kernel", "").split("<")[0] != kernel.split("<")[0]:
        return None, f"ncu capture is of {rec.get('kernel')}, the timed kernel is {kernel}: not quoted"
    return float(rec["dram_bytes_per_launch"]), (f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one "
                                                 f"{rec.get('kernel')} launch ({rec.get('report')}), source hash {now}")


# ------------------------------------------------------------------ CPU legs (oracle)
def use_all_host_threads():
    """The C oracle on every host core, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    import pyflow_oracle_c as orc_c
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    orc_c.set_threads(n)
    return orc_c.threads()


def oracle_sample(nx: int, ny: int, Re: float, lin_maxiter: int, steps: int, warmup: int, impl: str = "c"):
    """Times `steps` oracle steps (after `warmup`) of an nx x ny cavity from rest.
    impl "c": oracle/pyflow_oracle.c on every host core (OpenMP); "numpy": oracle/pyflow_oracle.py.
    Returns (mlups, seconds_per_step, true CG iterations per step)."""
    import pyflow_oracle as orc
    stepper = orc.step
    if impl == "c":
        import pyflow_oracle_c as orc_c
        ws = orc_c.Workspace(nx, ny)
        stepper = lambda cfg, st, it: orc_c.step(cfg, st, it, ws)          # noqa: E731
    cfg = make_cfg(nx, ny, Re, lin_maxiter)
    st = orc.allocate_state(nx, ny)
    it = 0
    for _ in range(max(1, warmup)):          # step 0 from rest has b = 0 (no CG work): never timed
        stepper(cfg, st, it)
        it += 1
    t0 = time.perf_counter()
    iters = []
    for _ in range(steps):
        _, _, d = stepper(cfg, st, it)
        iters.append(d["true_iterations"])
        it += 1
    dt = time.perf_counter() - t0
    return nx * ny * steps / dt / 1e6, dt / steps, iters


def sized_oracle_rows(nx: int, ny_max: int, Re: float, lin_maxiter: int, total_steps: int, budget_s: float):
    """Rows of the CPU sample: the workload's own row length always; as many rows (ny_max, ny_max/2, ... >= 256)
    as `total_steps` C-oracle steps fit into the time budget, from a calibration step on nx x 256 (cost scales
    with cells x CG iterations; x1.5 because the larger grid leaves the last-level cache)."""
    cal_rows = min(256, ny_max)
    cal_it = min(lin_maxiter, 300)
    _m, sps, _it = oracle_sample(nx, cal_rows, Re, cal_it, 1, 1)
    per_cell = sps / (nx * cal_rows) * 1.5 * (lin_maxiter / cal_it)
    rows = ny_max
    while rows > 256 and nx * rows * per_cell * total_steps > budget_s:
        rows //= 2
    return rows


def unmodified_reference_sample(cases=((64, 64, 100.0, 49), (512, 512, 1000.0, 1)), timeout_s=240.0):
    """The UNMODIFIED reference -- CPython bytecode of /root/reference/src/pyflow compiled by oracle/build_ref.py
    into oracle/_ref, nothing of this repo on its path -- driven through its own SimulationDriver.run on this box's
    host cores (BASELINE configs 1 and 2; SURVEY 8d: "time the unmodified reference through SimulationDriver.run").
    A subprocess per case: the reference package and the drop-in package are both called `pyflow`."""
    script = os.path.join(ROOT, "oracle", "time_unmodified.py")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pyflow", "__init__.pfbc")):
        return {"unavailable": "oracle/_ref not built (python oracle/build_ref.py needs /root/reference)"}
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    env["OMP_NUM_THREADS"] = env["OPENBLAS_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun exports 1
    out = {"what": "unmodified reference (bytecode in oracle/_ref), SimulationDriver.run from rest, step 0 untimed; "
                   "single-threaded Python / numpy / scipy, BLAS dots on all host threads", "cases": {}}
    for nx, ny, Re, steps in cases:
        key = f"{nx}x{ny} Re={Re:g}"
        try:
            r = subprocess.run([sys.executable, script, str(nx), str(ny), str(Re), str(steps)], env=env, cwd="/tmp",
                               capture_output=True, text=True, timeout=timeout_s)
            if r.returncode != 0:
                out["cases"][key] = {"error": (r.stderr or r.stdout).strip().splitlines()[-1:]}
            else:
                out["cases"][key] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as exc:
            out["cases"][key] = {"error": repr(exc)}
    return out


def run_reference_arm(args, wl, rank):
    """`--impl reference`: the reference's CPU algorithm for the path on this box's host cores.  The reference
    itself is Python and cannot travel to the box; its restatement oracle/pyflow_oracle.c (pinned through the
    numpy oracle and the golden vectors) runs on ALL host threads, same metric / config as the GPU arm.  Each step
    is a bounded sample of the workload: the workload's own row length, Re, scheme and solver settings on the rows of
    one GPU's slab (the whole grid at N = 1), halved only if K + W steps would not fit a few minutes."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    maxit = lin_maxiter_of(args, wl)
    total_steps = args.steps + max(1, args.warmup)
    ny_global = wl["ny"] * args.gpus if wl["per_gpu"] else wl["ny"]
    rows_max = wl["ny"] if wl["per_gpu"] else max(256, wl["ny"] // args.gpus)
    if args.ref_grid:
        nx, rows = args.ref_grid, args.ref_grid
    else:
        nx = wl["nx"]
        rows = sized_oracle_rows(nx, rows_max, wl["Re"], maxit, total_steps, getattr(args, 'ref_budget', 200.0))
    mlups, sps, iters = oracle_sample(nx, rows, wl["Re"], maxit, args.steps, args.warmup)
    whole = (nx == wl["nx"] and rows == ny_global)
    sample = (f"{args.steps} steps (after {max(1, args.warmup)} warm-up) of a {nx}x{rows} cavity from rest "
              f"({'the whole workload grid' if whole else f'a bounded sample of the {wl_nx_ny(wl, args.gpus)} workload: same row length, Re, scheme and solver settings'}), "
              f"Re={wl['Re']:g}, QUICK, lin_maxiter={maxit} ({int(statistics.mean(iters))} true CG iterations/step); C/OpenMP "
              f"restatement of the reference algorithm on {cores} threads (the unmodified reference is single-threaded "
              "Python/numpy/scipy and 20-50x slower, profiles/r1_reference_cpu_timings.json)")
    line = {
        "impl": "reference", "metric": "MLUPS per projection step", "value": mlups, "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, wl, args.gpus, maxit),
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample,
                         "sample_grid": {"nx": nx, "ny": rows},
                         "unmodified_reference": unmodified_reference_sample(
                             cases=((64, 64, 100.0, int(getattr(args, "unmodified_steps", 49))),))},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def wl_nx_ny(wl, n_gpus):
    return f"{wl['nx']}x{wl['ny'] * n_gpus if wl['per_gpu'] else wl['ny']}"


def workload_config(name, wl, n_gpus, lin_maxiter=300):
    """`config` of the JSON line: identical in the GPU arm and in the reference arm."""
    ny_global = wl["ny"] * n_gpus if wl["per_gpu"] else wl["ny"]
    return {
        "workload": f"lid-driven cavity {wl['nx']}x{ny_global} Re={wl['Re']:g} QUICK Jacobi-PCG lin_maxiter={lin_maxiter} "
                    f"lin_tol=1e-10 from rest ({name})",
        "nx": wl["nx"], "ny": ny_global, "rows_per_gpu": ny_global // n_gpus, "Re": wl["Re"],
        "advection_scheme": "quick", "lin_maxiter": lin_maxiter, "lin_tol": 1e-10,
        "parallelism": "single GPU" if n_gpus == 1 else
                       f"row slabs x{n_gpus}: one persistent CG kernel per rank, halo rows + dot-product partials through "
                       "NVLink peer memory inside the kernel; NCCL only for the per-step halo rows and step scalars",
        "l2": "no flush needed: every CG sweep streams >= 3 vectors of nx*ny*8 B each, far larger than the 126 MB L2"
              if wl["nx"] * (ny_global // n_gpus) * 8 * 3 > 3 * 126e6 else
              "working set is shared-memory / L2-resident by nature of this workload (latency-bound; HBM roofline informational)",
    }


def cg_kernel_name(nx, ny_local, world, sm_count=148):
    """Which CG kernel pf_step dispatches to for this slab (mirror of solve_on_device in csrc/pf_api.cu)."""
    rows = -(-ny_local // sm_count)
    if (world == 1 and nx % 2 == 0 and nx <= 2048 and rows * (nx // 2) <= 8 * 512
            and 8 * nx * ((rows + 2) * 2 + rows) <= 200 * 1024):
        return "k_cg_small<unit> (shared-memory-resident CG; L2/latency-bound workload, HBM roofline informational)"
    if nx >= 1024 and ny_local >= 2:
        return "k_cg_tmad<unit> (TMA-fed persistent CG)"
    return "k_cg_persistent<unit> (register-window persistent CG)"


# ------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, wl, rank, world):
    import torch
    import torch.distributed as dist

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from pyflow.core.ghost_fields import allocate_state
    from pyflow.drivers.simulation_driver import SimulationDriver
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    def new_state(nx, ny_global):
        if world > 1:
            from pyflow._b200.distributed import allocate_slab_state
            return allocate_slab_state(nx, ny_global, rank, world)
        return allocate_state(nx, ny_global)

    def timed_steps(cfg, state, tracker, it, warm, steps):
        """`warm` untimed + `steps` timed device-resident steps; CUDA events, barrier + synchronize on both sides,
        max over ranks."""
        for _ in range(warm):
            step(cfg, state, tracker, it)
            it += 1
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        cg_ms, cg_iters, step_ms = [], [], []
        for _ in range(steps):
            step(cfg, state, tracker, it)
            it += 1
            cg_ms.append(state.meta["cg_ms"])
            cg_iters.append(state.meta["cg_true_iterations"])
            step_ms.append(state.meta["step_ms"])
        ev1.record()
        barrier()
        return it, max_over_ranks(ev0.elapsed_time(ev1)), cg_ms, cg_iters, step_ms

    nx = wl["nx"]
    ny_global = wl["ny"] * world if wl["per_gpu"] else wl["ny"]
    maxit = lin_maxiter_of(args, wl)
    cfg = make_cfg(nx, ny_global, wl["Re"], maxit, lx=wl.get("lx"),
                   ly=(wl["ly"] * (world if wl["per_gpu"] else 1)) if wl.get("ly") else None)
    state = new_state(nx, ny_global)
    tracker = ResidualManager()
    warm = max(3, args.warmup)
    it = 0
    step(cfg, state, tracker, it)               # creates the device context (counted in the warm-up)
    it += 1
    slab = state.fields._slab
    sampler = ClockSampler(local_rank)
    for _ in range(warm - 1):
        step(cfg, state, tracker, it)
        it += 1
    launches0 = slab.launch_count()
    if rank == 0:
        sampler.start()
    # ---- timed region: K device-resident steps ---------------------------------
    it, elapsed_ms, cg_ms, cg_iters, step_ms = timed_steps(cfg, state, tracker, it, 0, args.steps)
    launches = slab.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    cells = nx * ny_global
    value = cells * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- e2e: host arrays -> driver.run K steps -> host arrays -------------------
    # (with several GPUs every rank moves its own slab; time = max over ranks)
    for name in ("u", "v", "p"):
        state.fields[name]                      # D2H now (untimed); marks the fields host-owned
    drv = SimulationDriver(cfg, state, tracker)
    barrier()
    t0 = time.perf_counter()
    last = None
    for st, res, diag in drv.run(max_steps=it + args.steps, start_iteration=it):
        last = (res["continuity"], diag["u_centerline"][1][0])
    fields_back = [state.fields[name] for name in ("u", "v", "p")]       # D2H inside the timed region
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    it += args.steps
    field_bytes = sum(a[1:-1, 1:-1].nbytes for a in fields_back) * world
    per_step_scalars = (32 * 8) * world + ny_global * 8
    e2e = {"value": cells * args.steps / e2e_s / 1e6, "unit": "MLUPS",
           "h2d_bytes_per_step": int(field_bytes / args.steps + 96 * world),
           "d2h_bytes_per_step": int(field_bytes / args.steps + per_step_scalars),
           "note": "SimulationDriver.run from host State arrays (page-locked numpy from allocate_state): 3 fields H2D "
                   "before the first step, per-step scalar block + centre-line D2H, 3 fields D2H after the last step, "
                   "all timed; bytes are totals over all ranks"}

    # ---- strictest reading of e2e, N = 1 only: the reference's live-array State semantics (state.fields.eager),
    # i.e. ALL three fields go host -> device before and device -> host after EVERY step (page-locked arrays)
    if world == 1:
        try:
            n_eager = max(1, min(args.steps, 5))
            state.fields.eager = True
            drv = SimulationDriver(cfg, state, tracker)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _st, _res, _diag in drv.run(max_steps=it + n_eager, start_iteration=it):
                pass
            torch.cuda.synchronize()
            eager_s = time.perf_counter() - t0
            it += n_eager
            per_step = sum(a[1:-1, 1:-1].nbytes for a in fields_back)
            e2e["every_step_copies"] = {
                "value": cells * n_eager / eager_s / 1e6, "unit": "MLUPS", "steps": n_eager,
                "h2d_bytes_per_step": int(per_step), "d2h_bytes_per_step": int(per_step + per_step_scalars),
                "note": "state.fields.eager = True: u, v, p uploaded before and downloaded after every step"}
        except Exception as exc:          # never lose the bench line over the extra figure
            e2e["every_step_copies"] = {"error": repr(exc)}
        finally:
            state.fields.eager = False

    peer_mode = bool(getattr(slab, "peer_mode", False))
    if os.environ.get("PF_BENCH_VERBOSE"):
        print(f"[rank {rank}] cg_ms mean {statistics.mean(cg_ms):.3f} min {min(cg_ms):.3f} max {max(cg_ms):.3f} "
              f"step_ms mean {statistics.mean(step_ms):.3f}", file=sys.stderr, flush=True)
    del drv, fields_back, state, slab
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (persistent CG), rank 0's launches -----------------
    peak, peak_src = measured_peak()

    def cg_roofline(nx_, nyl_, cg_ms_, cg_iters_):
        local_cells = nx_ * nyl_
        cg_bytes = sum(local_cells * CG_BYTES_PER_CELL_ITER * k for k in cg_iters_)
        cg_s = sum(cg_ms_) * 1e-3
        achieved = cg_bytes / cg_s / 1e9 if cg_s > 0 else 0.0
        return achieved, cg_bytes, cg_s

    achieved, cg_bytes, cg_s = cg_roofline(nx, ny_global // world, cg_ms, cg_iters)
    kernel = cg_kernel_name(nx, ny_global // world, world)
    traffic, traffic_src = recorded_traffic(args.workload, kernel) if world == 1 else (None, "single-GPU captures only")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": kernel, "kernel_source_hash": kernel_source_hash(),
                "algorithmic_bytes_per_launch": cg_bytes / max(1, len(cg_iters)),
                "avg_launch_ms": statistics.mean(cg_ms) if cg_ms else None,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "cg_share_of_step": sum(cg_ms) / sum(step_ms) if step_ms else None}
    local_cells = nx * (ny_global // world)
    step_bytes = sum(local_cells * (CG_BYTES_PER_CELL_ITER * k + STEP_BYTES_PER_CELL) for k in cg_iters)

    # ---- the other BASELINE configs at this GPU count (device-resident, few steps, all ranks) ----------
    other = None
    if not getattr(args, 'no_extra', False) and args.workload == "cavity4096":
        other = {}
        extra = ([("cavity64", 3, 49), ("cavity512", 3, 50), ("cavity4000", 2, 6), ("cavity5000x3000", 2, 6),
                  ("cavity4096_lx1", 2, 4)] if world == 1 else [])
        extra += [("strong8192", 2, 4), ("weak16384", 2, 2), ("weak16384_tight", 1, 1)]
        for name, w2, n2 in extra:
            wl2 = WORKLOADS[name]
            ny2 = wl2["ny"] * world if wl2["per_gpu"] else wl2["ny"]
            if ny2 % world or ny2 // world < 2:
                continue
            try:
                cfg2 = make_cfg(wl2["nx"], ny2, wl2["Re"], wl2.get("lin_maxiter", 300), lx=wl2.get("lx"),
                                ly=(wl2["ly"] * (world if wl2["per_gpu"] else 1)) if wl2.get("ly") else None)
                st2, tr2 = new_state(wl2["nx"], ny2), ResidualManager()
                _it, ms2, cg2, its2, _sm = timed_steps(cfg2, st2, tr2, 0, w2, n2)
                ach2, _b, cgs2 = cg_roofline(wl2["nx"], ny2 // world, cg2, its2)
                other[name] = {
                    "config": workload_config(name, wl2, world, wl2.get("lin_maxiter", 300))["workload"],
                    "scaling": wl2["scaling"], "steps": n2, "warmup": w2, "ms_per_step": ms2 / n2,
                    "MLUPS": wl2["nx"] * ny2 * n2 / (ms2 * 1e-3) / 1e6,
                    "cg_iters_per_step": statistics.mean(its2), "cg_iters_per_s": sum(its2) / cgs2 if cgs2 > 0 else None,
                    "cg_kernel": cg_kernel_name(wl2["nx"], ny2 // world, world).split(" ")[0].replace(
                        "<unit>", "<general spacing>" if wl2.get("lx") else "<unit>"),
                    "cg_hbm_gbs_algorithmic_per_gpu": ach2, "cg_frac_of_measured_peak_per_gpu": ach2 / peak,
                    "peer_kernel": bool(getattr(st2.fields._slab, "peer_mode", False))}
                st2.fields._slab.close()
                del st2
                torch.cuda.empty_cache()
            except Exception as exc:          # an extra figure must never cost the headline line
                other[name] = {"error": repr(exc)}

    # ---- parity (UNTIMED): the CUDA path against the C oracle, at this GPU count -------------------------
    parity = None
    if not getattr(args, 'no_parity', False):
        import parity as parity_mod
        try:
            parity = parity_mod.run_bench_block(args.workload, nx, ny_global // world, wl["Re"], rank, world)
        except Exception as exc:
            parity = {"ok": False, "error": repr(exc)}

    if rank != 0:
        if world > 1:
            ok = torch.tensor([1], device="cuda")
            dist.broadcast(ok, src=0)
            dist.destroy_process_group()
            return 0 if int(ok.item()) == 1 else 1
        return 0
    # ---- CPU baseline: bounded oracle sample on this box ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = use_all_host_threads()
        rows = sized_oracle_rows(nx, wl["ny"], wl["Re"], maxit, 3, 25.0)
        mlups, sps, iters = oracle_sample(nx, rows, wl["Re"], maxit, 2, 1)
        npn = min(nx, 512)
        np_mlups, np_sps, _ = oracle_sample(npn, npn, wl["Re"], 300, 1, 1, impl="numpy")
        cpu = {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
               "sample": f"2 steps (#1-#2 from rest, {int(statistics.mean(iters))} true CG iterations each) of a {nx}x{rows} "
                         f"cavity, same Re/scheme/solver settings, {sps:.2f} s/step; C/OpenMP restatement of the "
                         "reference algorithm (oracle/pyflow_oracle.c) on all host threads",
               "sample_grid": {"nx": nx, "ny": rows},
               "numpy_port_mlups": np_mlups,
               "numpy_port_sample": f"1 step of a {npn}x{npn} cavity with the single-threaded numpy "
                                    "oracle (the unmodified reference is slower still: Python loops for advection)",
               # BASELINE configs 1 and 2 with the reference itself; the GPU figures for the same two configurations
               # are other_configs.cavity64 / other_configs.cavity512 of this line
               "unmodified_reference": unmodified_reference_sample()}
    line = {
        "metric": "MLUPS per projection step", "value": value, "unit": "MLUPS", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, wl, world, maxit),
        "cg_comm": ("single GPU" if world == 1 else
                    "one persistent kernel per rank, halo rows + dot-product partials over NVLink peer memory"
                    if peer_mode else "NCCL send/recv + all-reduce between phase kernels"),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "cg_iters_per_s": sum(cg_iters) / cg_s if cg_s > 0 else None,
        "cg_iters_per_step": statistics.mean(cg_iters) if cg_iters else None,
        "step_hbm_gbs_algorithmic": step_bytes / (sum(step_ms) * 1e-3) / 1e9 if step_ms else None,
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)
    good = parity is None or bool(parity.get("ok"))
    if world > 1:
        ok = torch.tensor([1 if good else 0], device="cuda")
        dist.broadcast(ok, src=0)
        dist.destroy_process_group()
    if not good:
        print("bench.py: PARITY FAILED (see the `parity` object of the JSON line)", file=sys.stderr, flush=True)
    return 0 if good else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cavity4096", choices=sorted(WORKLOADS))
    ap.add_argument("--lin-maxiter", type=int, default=0, help="override the workload's CG iteration cap")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity block")
    ap.add_argument("--no-extra", action="store_true", help="skip `other_configs` (BASELINE configs 2, 4, 5)")
    ap.add_argument("--ref-grid", type=int, default=0, help="reference arm: force the sample grid size (tests)")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="reference arm: seconds for K + W oracle steps")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return 0
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    return run_gpu_arm(args, wl, rank, world)


if __name__ == "__main__":
    raise SystemExit(main())
