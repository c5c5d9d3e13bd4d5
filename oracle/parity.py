"""Parity checker: the CUDA path against the C / OpenMP oracle (TEST INFRASTRUCTURE ONLY).

Used by ``tests/`` (the `-m gpu` suite), ``__graft_entry__.smoke()`` and the
UNTIMED ``parity`` block that ``bench.py`` appends to its JSON line at every
GPU count.  It drives the product through its public surface
(``allocate_state`` / ``allocate_slab_state``, ``solvers.solver.step``,
``gather_field``) and the oracle through ``pyflow_oracle_c.run`` -- the C
restatement of the reference algorithm (reference solver.py:15-193,
pressure_solver.py:97-219; pinned through the numpy oracle and the golden
vectors written by the unmodified reference, tests/test_oracle_golden.py).
Nothing here is timed, shipped, or imported by the product package.

Bars: dt sequence bit-identical, true CG iteration counts within 1,
u / v / p relative L2 <= 1e-9 after N steps (truncated-early and converged
regimes both sit orders of magnitude below that, see the noise table of
SURVEY 8c).
"""
from __future__ import annotations

import os
import time
from types import SimpleNamespace

import numpy as np

REL_BAR = 1e-9
ITER_BAR = 1


def iter_bar(oracle_iters):
    """True CG iteration counts agree within 1 (north star) for solves of up to 2500 iterations -- the range SURVEY 8c's
    noise table covers (the reference against itself with another summation order: +-1 up to 2604 iterations).  Longer
    converged solves get +-1 per further 2500 iterations: after several thousand iterations the recursive residual
    of two runs that differ only in summation order no longer crosses the tolerance in the same iteration (measured:
    1024 x 512 cavity, 7857 / 7322 iterations, 8 GPUs vs oracle: 0 / 3; fields 3e-10)."""
    longest = max(oracle_iters, default=0)
    return ITER_BAR + longest // 2500


def make_cfg(nx, ny, Re=1000.0, scheme="quick", lin_maxiter=300, lin_tol=1e-10, **extra):
    return SimpleNamespace(nx=nx, ny=ny, Re=Re, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05,
                           lin_tol=lin_tol, lin_maxiter=lin_maxiter, advection_scheme=scheme,
                           disable_advection=False, diagnostics=False, **extra)


def rel_l2(a, b):
    d = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (d or 1.0)


def use_all_host_threads():
    """The C oracle runs on every host core, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    import pyflow_oracle_c as orc_c
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    orc_c.set_threads(n)
    return orc_c.threads()


def gpu_run(cfg, steps, rank=0, world=1, options=None):
    """`steps` projection steps from rest through the product's public step(); returns the full-grid interior
    fields (every rank when world > 1), the dt sequence, the true CG iteration counts and the CG kernel's
    communication mode."""
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    old = os.environ.get("PYFLOW_B200_OPTIONS")
    if options:
        os.environ["PYFLOW_B200_OPTIONS"] = options
    try:
        if world > 1:
            from pyflow._b200.distributed import allocate_slab_state, gather_field
            st = allocate_slab_state(cfg.nx, cfg.ny, rank, world)
        else:
            from pyflow.core.ghost_fields import allocate_state
            st = allocate_state(cfg.nx, cfg.ny)
        tr = ResidualManager()
        dts, iters = [], []
        for it in range(steps):
            _, _res, diag = step(cfg, st, tr, it)
            dts.append(diag["dt"])
            iters.append(int(st.meta["cg_true_iterations"]))
        peer = bool(getattr(st.fields._slab, "peer_mode", False))
        if world > 1:
            fields = {k: gather_field(st, k)[1:-1, 1:-1] for k in ("u", "v", "p")}
        else:
            fields = {k: np.array(st.fields[k][1:-1, 1:-1]) for k in ("u", "v", "p")}
        st.fields._slab.close()
    finally:
        if options:
            if old is None:
                os.environ.pop("PYFLOW_B200_OPTIONS", None)
            else:
                os.environ["PYFLOW_B200_OPTIONS"] = old
    return fields, dts, iters, peer


def oracle_run(cfg, steps):
    import pyflow_oracle as orc
    import pyflow_oracle_c as orc_c
    ost, od = orc_c.run(cfg, steps)
    return ({k: orc.interior(ost.fields[k]) for k in ("u", "v", "p")}, [d["dt"] for d in od],
            [int(d["true_iterations"]) for d in od])


def compare(name, got, want, rel_bar=REL_BAR):
    gf, gdt, git, peer = got
    wf, wdt, wit = want
    errs = {k: rel_l2(gf[k], wf[k]) for k in ("u", "v", "p")}
    dit = max((abs(a - b) for a, b in zip(git, wit)), default=0)
    finite = all(np.isfinite(gf[k]).all() for k in gf)
    bar = iter_bar(wit)
    ok = bool(gdt == wdt and dit <= bar and finite and all(e <= rel_bar for e in errs.values()))
    return {"case": name, "rel_l2": errs, "max_rel_l2": max(errs.values()), "dt_equal": gdt == wdt,
            "iters": git, "oracle_iters": wit, "max_iter_diff": int(dit), "iter_bar": bar, "peer_kernel": peer, "ok": ok}


_GLOO = None


def _host_barrier():
    """Barrier over a gloo group: the ranks that wait for rank 0's oracle SLEEP in a socket poll.  (Waiting in the next
    NCCL collective instead spins one host core per rank, and the OpenMP barriers of the oracle collapse when its
    threads have to share cores with spinning processes: measured 16 s instead of 0.5 s for a 4096 x 512 case.)"""
    global _GLOO
    import torch.distributed as dist
    if _GLOO is None:
        _GLOO = dist.new_group(backend="gloo")
    dist.barrier(group=_GLOO)


def check_case(name, cfg, steps, rank=0, world=1, options=None):
    """One case on `world` GPUs; the comparison happens on rank 0 (other ranks return None).  Collective."""
    got = gpu_run(cfg, steps, rank, world, options)
    out = None
    if rank == 0:
        t0 = time.perf_counter()
        want = oracle_run(cfg, steps)
        out = compare(name, got, want)
        out["oracle_s"] = round(time.perf_counter() - t0, 2)
    if world > 1:
        _host_barrier()
    return out


def bench_cases(workload, nx, ny_local, Re, world):
    """Parity cases of bench.py's untimed block for a workload on `world` GPUs (all C-oracle backed):
    the workload's own grid when the oracle finishes it in seconds, a 256-rows-per-GPU slab of the same
    width / Re / solver settings, and (several GPUs) the slab cases of tests/mp_slab_check.py."""
    cases = []
    rows = 256
    cases.append((f"{workload}: {nx}x{rows * world} slab (256 rows per GPU), 300 its", make_cfg(nx, rows * world, Re), 3))
    if world == 1 and nx * ny_local <= 4096 * 4096:
        cases.append((f"{workload}: full {nx}x{ny_local} grid, 300 its", make_cfg(nx, ny_local, Re), 3))
    if world > 1:
        cases += [
            ("96x80 converged (lin_maxiter 10000)", make_cfg(96, 80, 100.0, lin_maxiter=10000), 4),
            ("128x128 truncated", make_cfg(128, 128, 100.0), 5),
            ("70x61 ragged non-unit upwind", make_cfg(70, 61, 400.0, "upwind", 100, lx=2.0, ly=1.0), 4),
            (f"200x{48 * world} register-window peer kernel", make_cfg(200, 48 * world, 1000.0), 4),
            (f"1024x{36 * world} TMA kernel", make_cfg(1024, 36 * world, 1000.0), 3),
            (f"2048x{20 * world} TMA kernel, 2 chunks, upwind", make_cfg(2048, 20 * world, 1000.0, "upwind", 200), 3),
            (f"1024x{64 * world} converged (lin_maxiter 10000)", make_cfg(1024, 64 * world, 100.0, lin_maxiter=10000), 3),
            (f"1100x{24 * world} TMA kernel, clipped last chunk", make_cfg(1100, 24 * world, 1000.0), 3),
        ]
    return cases


def run_bench_block(workload, nx, ny_local, Re, rank, world):
    """The `parity` object of bench.py's JSON line.  Collective: every rank calls it."""
    threads = use_all_host_threads() if rank == 0 else None
    details = []
    for name, cfg, steps in bench_cases(workload, nx, ny_local, Re, world):
        r = check_case(name, cfg, steps, rank, world)
        if r is not None:
            details.append(r)
    if rank != 0:
        return None
    return {"cases": len(details), "ok": all(d["ok"] for d in details),
            "max_rel_l2": max(d["max_rel_l2"] for d in details),
            "max_iter_diff": max(d["max_iter_diff"] for d in details),
            "dt_equal": all(d["dt_equal"] for d in details),
            "bar": {"rel_l2": REL_BAR, "iters": "1 (+1 per 2500 iterations of a converged solve)", "dt": "bit-identical"},
            "oracle": f"oracle/pyflow_oracle.c (C/OpenMP restatement of the reference, {threads} threads), untimed",
            "details": [{k: d[k] for k in ("case", "max_rel_l2", "max_iter_diff", "iter_bar", "dt_equal", "iters",
                                           "peer_kernel", "ok", "oracle_s")} for d in details]}
