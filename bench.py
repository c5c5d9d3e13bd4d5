"""bench.py -- throughput of the PyFlow projection time step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): MLUPS per projection step = nx * ny_global * steps / seconds / 1e6.
A "step" is one full projection time step (predictor, RHS, <=300 Jacobi-PCG
iterations, correction, lid BC, residual pass) of the lid-driven cavity started
from rest -- no RNG anywhere on this path, so the data are "synthetic" by
construction.

Workloads (BASELINE.json configs):
  cavity4096  4096 x 4096 rows PER GPU, Re=1000, QUICK, lin_maxiter=300  (config 3; default;
              with N GPUs the global grid is 4096 x 4096N row slabs: weak scaling)
  cavity512   512 x 512, Re=1000 (config 2; L2-resident, latency-bound)
  strong8192  8192 x 8192 total, Re=5000, rows split over N GPUs (config 4; strong)
  weak16384   16384 x 16384 rows per GPU, Re=1000 (config 5; weak)

One JSON line on stdout from rank 0.  `value` is device-resident stepping
through the public `step()`; `e2e` is the same metric through
`SimulationDriver.run` starting from HOST arrays and ending with the fields
back on the host, copies inside the timed region.  `roofline` is the CG kernel:
algorithmic bytes (64 B per cell per CG iteration) / its CUDA-event
duration, against MEASURED_PEAKS.json.  `cpu_baseline` (and `--impl reference`)
is the C / OpenMP restatement of the reference algorithm (oracle/pyflow_oracle.c)
on a bounded sample on all of this box's host cores; the single-threaded numpy
oracle is reported beside it.
"""
from __future__ import annotations

import argparse
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
    # name: (nx, ny_per_gpu or total, Re, scaling, description)
    "cavity4096": dict(nx=4096, ny=4096, Re=1000.0, scaling="weak", per_gpu=True),
    "cavity512": dict(nx=512, ny=512, Re=1000.0, scaling="weak", per_gpu=True),
    "strong8192": dict(nx=8192, ny=8192, Re=5000.0, scaling="strong", per_gpu=False),
    "weak16384": dict(nx=16384, ny=16384, Re=1000.0, scaling="weak", per_gpu=True),
}
CG_BYTES_PER_CELL_ITER = 64.0     # SURVEY.md 8d contract figure
STEP_BYTES_PER_CELL = 200.0


def make_cfg(nx, ny, Re, lin_maxiter=300, lin_tol=1e-10):
    return SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05,
                           lin_tol=lin_tol, lin_maxiter=lin_maxiter, advection_scheme="quick",
                           disable_advection=False, diagnostics=False)


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


# ------------------------------------------------------------------ CPU legs (oracle)
def oracle_sample(n: int, Re: float, lin_maxiter: int, steps: int, warmup: int, impl: str = "c"):
    """Times `steps` oracle steps (after `warmup`) of an n x n cavity from rest.
    impl "c": oracle/pyflow_oracle.c on every host core (OpenMP); "numpy": oracle/pyflow_oracle.py.
    Returns (mlups, seconds_per_step, true CG iterations per step)."""
    import pyflow_oracle as orc
    stepper = orc.step
    if impl == "c":
        import pyflow_oracle_c as orc_c
        ws = orc_c.Workspace(n, n)
        stepper = lambda cfg, st, it: orc_c.step(cfg, st, it, ws)          # noqa: E731
    cfg = make_cfg(n, n, Re, lin_maxiter)
    st = orc.allocate_state(n, n)
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
    return n * n * steps / dt / 1e6, dt / steps, iters


def cpu_threads():
    """Host threads the C / OpenMP oracle runs on."""
    import pyflow_oracle_c as orc_c
    return orc_c.threads()


def sized_oracle_grid(n_max: int, total_steps: int, budget_s: float):
    """Largest power-of-two-scaled grid <= n_max whose `total_steps` C-oracle steps fit the time budget, from a
    calibration step at 512^2 (cost scales with cells x CG iterations; large grids leave the caches, hence x2)."""
    _m, sps, _it = oracle_sample(512, 1000.0, 300, 1, 1)
    per_cell = sps / (512 * 512) * 2.0
    n = n_max
    while n > 512 and n * n * per_cell * total_steps > budget_s:
        n //= 2
    return n


def run_reference_arm(args, wl, rank):
    """`--impl reference`: the reference's CPU algorithm for the path on this box's host cores.  The reference
    itself is Python and cannot travel to the box; its restatement oracle/pyflow_oracle.c (pinned through the
    numpy oracle and the golden vectors) runs on all host threads, same metric / config keys."""
    if rank != 0:
        return
    total_steps = args.steps + max(1, args.warmup)
    n = args.ref_grid or sized_oracle_grid(min(wl["nx"], 4096), total_steps, 150.0)
    mlups, sps, iters = oracle_sample(n, wl["Re"], 300, args.steps, args.warmup)
    cores = cpu_threads()
    sample = (f"{args.steps} steps (after {max(1, args.warmup)} warm-up) of a {n}x{n} cavity from rest, Re={wl['Re']}, "
              f"QUICK, lin_maxiter=300 ({int(statistics.mean(iters))} true CG iterations/step); C/OpenMP restatement "
              f"of the reference algorithm on {cores} threads (the unmodified reference is single-threaded "
              "Python/numpy/scipy and far slower)")
    line = {
        "impl": "reference", "metric": "MLUPS per projection step", "value": mlups, "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, wl, args.gpus),
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, wl, n_gpus):
    ny_global = wl["ny"] * n_gpus if wl["per_gpu"] else wl["ny"]
    return {
        "workload": f"lid-driven cavity {wl['nx']}x{ny_global} Re={wl['Re']:g} QUICK Jacobi-PCG lin_maxiter=300 "
                    f"lin_tol=1e-10 from rest ({name})",
        "nx": wl["nx"], "ny": ny_global, "rows_per_gpu": ny_global // n_gpus, "Re": wl["Re"],
        "advection_scheme": "quick", "lin_maxiter": 300, "lin_tol": 1e-10,
        "parallelism": "single GPU" if n_gpus == 1 else f"row slabs x{n_gpus} (NCCL halo rows + scalar all-reduce)",
        "l2": "no flush needed: every CG sweep streams >= 3 vectors of nx*ny*8 B each, far larger than the 126 MB L2"
              if wl["nx"] * (ny_global // n_gpus) * 8 * 3 > 3 * 126e6 else
              "working set is L2-resident by nature of this workload (latency-bound; HBM roofline informational)",
    }


def cg_kernel_name(nx, ny_local, world, sm_count=148):
    """Which CG kernel pf_step dispatches to for this slab (mirror of solve_on_device in csrc/pf_api.cu)."""
    rows = -(-ny_local // sm_count)
    if (world == 1 and nx % 2 == 0 and nx <= 2048 and rows * (nx // 2) <= 8 * 512
            and 8 * nx * ((rows + 2) * 2 + rows) <= 200 * 1024):
        return "k_cg_small<unit> (shared-memory-resident CG; L2/latency-bound workload, HBM roofline informational)"
    if nx % 1024 == 0:
        return "k_cg_tmad<unit> (TMA-fed persistent CG)"
    return "k_cg_persistent<unit> (register-window persistent CG)"


# ------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, wl, rank, world):
    import numpy as np
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

    nx = wl["nx"]
    ny_global = wl["ny"] * world if wl["per_gpu"] else wl["ny"]
    cfg = make_cfg(nx, ny_global, wl["Re"])
    if world > 1:
        from pyflow._b200.distributed import allocate_slab_state
        state = allocate_slab_state(nx, ny_global, rank, world)
    else:
        state = allocate_state(nx, ny_global)
    tracker = ResidualManager()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    it = 0
    for _ in range(max(3, args.warmup)):
        step(cfg, state, tracker, it)
        it += 1
    slab = state.fields._slab
    launches0 = slab.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region: K device-resident steps ---------------------------------
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    cg_ms, cg_iters, step_ms = [], [], []
    for _ in range(args.steps):
        step(cfg, state, tracker, it)
        it += 1
        cg_ms.append(state.meta["cg_ms"])
        cg_iters.append(state.meta["cg_true_iterations"])
        step_ms.append(state.meta["step_ms"])
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = slab.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
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
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
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
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (persistent CG) -------------------------
    peak, peak_src = measured_peak()
    local_cells = nx * (ny_global // world)
    cg_bytes = sum(local_cells * CG_BYTES_PER_CELL_ITER * k for k in cg_iters)
    cg_s = sum(cg_ms) * 1e-3
    achieved = cg_bytes / cg_s / 1e9 if cg_s > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "cg_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": cg_kernel_name(nx, ny_global // world, world),
                "algorithmic_bytes_per_launch": cg_bytes / max(1, len(cg_iters)),
                "avg_launch_ms": statistics.mean(cg_ms) if cg_ms else None,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "cg_share_of_step": sum(cg_ms) / sum(step_ms) if step_ms else None}
    step_bytes = sum(local_cells * (CG_BYTES_PER_CELL_ITER * k + STEP_BYTES_PER_CELL) for k in cg_iters)
    # ---- CPU baseline: bounded oracle sample on this box ---------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n = sized_oracle_grid(min(wl["nx"], 4096), 3, 25.0)
        mlups, sps, iters = oracle_sample(n, wl["Re"], 300, 2, 1)
        np_mlups, np_sps, _ = oracle_sample(min(n, 512), wl["Re"], 300, 1, 1, impl="numpy")
        cpu = {"value": mlups, "unit": "MLUPS", "cores": cpu_threads(), "kind": "port",
               "sample": f"2 steps (#1-#2 from rest, {int(statistics.mean(iters))} true CG iterations each) of a {n}x{n} "
                         f"cavity, same Re/scheme/solver settings, {sps:.2f} s/step; C/OpenMP restatement of the "
                         "reference algorithm (oracle/pyflow_oracle.c) on all host threads",
               "numpy_port_mlups": np_mlups,
               "numpy_port_sample": f"1 step of a {min(n, 512)}x{min(n, 512)} cavity with the single-threaded numpy "
                                    "oracle (the unmodified reference is slower still: Python loops for advection)"}
    # ---- BASELINE config 2 (512^2, L2-resident, latency-bound) beside the headline, N = 1 only -------
    other = None
    if world == 1 and args.workload == "cavity4096":
        cfg2 = make_cfg(512, 512, 1000.0)
        st2, tr2 = allocate_state(512, 512), ResidualManager()
        for k in range(3):
            step(cfg2, st2, tr2, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n2, its2, cg2 = 50, 0, 0.0
        for k in range(3, 3 + n2):
            step(cfg2, st2, tr2, k)
            its2 += st2.meta["cg_true_iterations"]
            cg2 += st2.meta["cg_ms"]
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n2
        other = {"cavity512 (BASELINE config 2: 512x512 Re=1000, 300 CG its/step, L2-resident)": {
            "ms_per_step": ms2, "MLUPS": 512 * 512 / (ms2 * 1e-3) / 1e6, "cg_iters_per_s": its2 / (cg2 * 1e-3),
            "steps": n2}}
    line = {
        "metric": "MLUPS per projection step", "value": value, "unit": "MLUPS", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(workload_config(args.workload, wl, world),
                                            cg_comm=("single GPU" if world == 1 else
                                                     "one persistent kernel per rank, halo rows + dot-product partials over NVLink peer memory"
                                                     if peer_mode else "NCCL send/recv + all-reduce between phase kernels")),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "cg_iters_per_s": sum(cg_iters) / cg_s if cg_s > 0 else None,
        "cg_iters_per_step": statistics.mean(cg_iters) if cg_iters else None,
        "step_hbm_gbs_algorithmic": step_bytes / (sum(step_ms) * 1e-3) / 1e9 if step_ms else None,
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cavity4096", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-grid", type=int, default=0, help="reference arm: force the sample grid size (tests)")
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
    run_gpu_arm(args, wl, rank, world)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
