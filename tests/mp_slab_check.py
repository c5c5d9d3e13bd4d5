"""Multi-process checks of the row-slab path; launched by the tests through
``python -m torch.distributed.run --nproc-per-node N tests/mp_slab_check.py MODE``.

MODE gloo : host-side slab logic only (CPU, no CUDA): row split, scatter from a
            full state, gather back, checkpoint of a gathered state.
MODE nccl : N GPUs; a slab-decomposed cavity run is compared with the oracle on
            the full grid (rank 0 prints PARITY lines; exit code 1 on failure).
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "pyflow-solver_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

from pyflow._b200.distributed import allocate_slab_state, gather_field, scatter_state, slab_rows
from pyflow.core.ghost_fields import allocate_state


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) or 1.0))


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gloo":
        dist.init_process_group("gloo")
        nx, ny = 13, 11                       # ragged split on purpose
        covered = [slab_rows(ny, r, world) for r in range(world)]
        assert covered[0][0] == 0 and sum(n for _, n in covered) == ny
        assert all(covered[i][0] + covered[i][1] == covered[i + 1][0] for i in range(world - 1))
        full = allocate_state(nx, ny)
        rng = np.random.default_rng(3)
        for k in ("u", "v", "p"):
            full.fields[k][1:-1, 1:-1] = rng.standard_normal((ny, nx))
        st = scatter_state(full, nx, ny, rank, world)
        row0, nyl = slab_rows(ny, rank, world)
        assert st.ny == nyl and st.slab["row_offset"] == row0
        assert np.array_equal(st.fields["u"][1:-1], full.fields["u"][1 + row0:1 + row0 + nyl])
        for k in ("u", "v", "p"):
            back = gather_field(st, k)
            assert np.array_equal(back[1:-1, 1:-1], full.fields[k][1:-1, 1:-1]), k
        try:
            allocate_slab_state(8, 2, rank, 4)
            raise SystemExit("thin slab accepted")
        except ValueError:
            pass
        # the drop-in driver on slab states with the GPU contexts replaced by the gathering stand-in
        # (tests/cpu_slab.py): every rank steps the same dt sequence, the centre line is the concatenation
        # of the slabs' column segments, the gathered fields equal the oracle on the full grid
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cpu_slab
        import pyflow_oracle as orc
        from pyflow.drivers.simulation_driver import SimulationDriver
        cpu_slab.install_slabs(rank, world)
        cfg = SimpleNamespace(nx=14, ny=11, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05,
                              lin_tol=1e-10, lin_maxiter=40, advection_scheme="quick", disable_advection=False,
                              diagnostics=False)
        st = allocate_slab_state(cfg.nx, cfg.ny, rank, world)
        diags = [dict(d) for _s, _r, d in SimulationDriver(cfg, st).run(max_steps=4)]
        assert st.fields._slab.nranks == world and st.fields._slab.downloads == []
        ost, od = orc.run(cfg, 4)
        assert [d["dt"] for d in diags] == [d["dt"] for d in od]                 # identical on EVERY rank
        assert [d["Rp_residual"] for d in diags] == [d["Rp_residual"] for d in od]
        line = diags[-1]["u_centerline"]
        assert len(line[0]) == cfg.ny and line[1] == orc.interior(ost.fields["u"])[:, cfg.nx // 2].tolist()
        for k in ("u", "v", "p"):
            assert np.array_equal(gather_field(st, k), ost.fields[k]), k
        if rank == 0:
            print("GLOO_OK")
        dist.destroy_process_group()
        return 0

    # ---- nccl / peer: slab run vs oracle -----------------------------------------
    # MODE nccl : CG with NCCL halo + all-reduce calls per phase
    # MODE peer : CG as one persistent kernel per rank over NVLink peer memory
    os.environ["PYFLOW_B200_CG_COMM"] = "nccl" if mode == "nccl" else "peer"
    import pyflow_oracle as orc
    from pyflow.residuals.manager import ResidualManager
    from pyflow.solvers.solver import step
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ok = True
    cases = [
        dict(nx=96, ny=80, Re=100.0, advection_scheme="quick", lin_maxiter=10000, steps=4),   # converged regime
        dict(nx=128, ny=128, Re=100.0, advection_scheme="quick", lin_maxiter=300, steps=5),   # truncated early
        dict(nx=70, ny=61, Re=400.0, advection_scheme="upwind", lin_maxiter=100, steps=4, lx=2.0, ly=1.0),  # ragged, non-unit, truncated early
        dict(nx=200, ny=48 * world, Re=1000.0, advection_scheme="quick", lin_maxiter=300, steps=4),           # equal slabs, several strips
        dict(nx=1024, ny=36 * world, Re=1000.0, advection_scheme="quick", lin_maxiter=300, steps=3),          # TMA-fed kernel (nx % 1024 == 0)
        dict(nx=2048, ny=20 * world, Re=1000.0, advection_scheme="upwind", lin_maxiter=200, steps=3),         # TMA-fed kernel, 2 column chunks
    ]
    used_peer = []
    for case in cases:
        steps = case.pop("steps")
        cfg = SimpleNamespace(lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                              disable_advection=False, diagnostics=False, **case)
        st = allocate_slab_state(cfg.nx, cfg.ny, rank, world)
        tr = ResidualManager()
        dts, iters = [], []
        for it in range(steps):
            _, res, diag = step(cfg, st, tr, it)
            dts.append(diag["dt"])
            iters.append(st.meta["cg_true_iterations"])
        used_peer.append(bool(st.fields._slab.peer_mode))
        fields = {k: gather_field(st, k) for k in ("u", "v", "p")}
        if rank == 0:
            ost, od = orc.run(cfg, steps)
            errs = {k: rel(fields[k][1:-1, 1:-1], orc.interior(ost.fields[k])) for k in ("u", "v", "p")}
            same_dt = dts == [d["dt"] for d in od]
            dit = max(abs(a - d["true_iterations"]) for a, d in zip(iters, od))
            good = same_dt and all(e < 1e-8 for e in errs.values()) and dit <= 2
            print(f"PARITY mode={mode} peer_kernel={used_peer[-1]} world={world} {cfg.nx}x{cfg.ny} errs={errs} dt_equal={same_dt} iters={iters} max_iter_diff={dit} "
                  f"{'OK' if good else 'FAIL'}", flush=True)
            ok = ok and good
    # the drop-in driver on slab states: every rank yields, the centre line is the concatenation of the slabs
    from pyflow.drivers.simulation_driver import SimulationDriver
    cfg = SimpleNamespace(nx=64, ny=16 * world, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=200, advection_scheme="quick", disable_advection=False, diagnostics=False)
    st = allocate_slab_state(cfg.nx, cfg.ny, rank, world)
    diags = [d for _s, _r, d in SimulationDriver(cfg, st).run(max_steps=3)]
    full_u = gather_field(st, "u")
    if rank == 0:
        ost, od = orc.run(cfg, 3)
        line = np.array(diags[-1]["u_centerline"][1])
        good = (len(line) == cfg.ny and np.array_equal(line, full_u[1:-1, 1 + cfg.nx // 2])
                and rel(full_u[1:-1, 1:-1], orc.interior(ost.fields["u"])) < 1e-9
                and [d["dt"] for d in diags] == [d["dt"] for d in od])
        print(f"PARITY driver world={world} centerline_len={len(line)} {'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    raise SystemExit(main())
