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

from pyflow._b200.distributed import allocate_slab_state, gather_field, scatter_state, slab_rows  # noqa: E402
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
        import tempfile
        box = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ckpt = os.path.join(box[0], "slab.npz")
        diags = [dict(d) for _s, _r, d in SimulationDriver(cfg, st).run(max_steps=4, checkpoint_path=ckpt,
                                                                         checkpoint_interval=3)]
        assert st.fields._slab.nranks == world
        # ONE checkpoint of the WHOLE grid, written by rank 0 after step 3 (per-rank files would overwrite each other)
        dist.barrier()
        from pyflow.io.checkpoint import load_checkpoint
        back, meta = load_checkpoint(ckpt)
        assert (back.nx, back.ny) == (cfg.nx, cfg.ny) and meta["iteration"] == 3 and sorted(os.listdir(box[0])) == ["slab.npz"]
        ost, od = orc.run(cfg, 4)
        assert [d["dt"] for d in diags] == [d["dt"] for d in od]                 # identical on EVERY rank
        assert [d["Rp_residual"] for d in diags] == [d["Rp_residual"] for d in od]
        line = diags[-1]["u_centerline"]
        assert len(line[0]) == cfg.ny and line[1] == orc.interior(ost.fields["u"])[:, cfg.nx // 2].tolist()
        for k in ("u", "v", "p"):
            assert np.array_equal(gather_field(st, k), ost.fields[k]), k
            assert np.array_equal(back.fields[k], ost.fields[k]), k
        if rank == 0:
            print("GLOO_OK")
        dist.destroy_process_group()
        return 0

    # ---- nccl / peer: slab run vs oracle -----------------------------------------
    # MODE nccl : CG with NCCL halo + all-reduce calls per phase
    # MODE peer : CG as one persistent kernel per rank over NVLink peer memory
    os.environ["PYFLOW_B200_CG_COMM"] = "nccl" if mode == "nccl" else "peer"
    import parity
    import pyflow_oracle as orc
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if rank == 0:
        parity.use_all_host_threads()
    ok = True
    # the slab cases of bench.py's untimed parity block (ragged, non-unit, converged and truncated solves, the
    # register-window and the TMA-fed peer kernels) plus a 4096-wide slab: config 3's row length, 256 rows per GPU
    cases = parity.bench_cases("cavity4096", 4096, 4096, 1000.0, world)
    if mode == "nccl":                  # split-phase path: launch-bound (5 launches + 4 NCCL calls per iteration); skip the long solves
        cases = [c for c in cases if "converged (lin_maxiter 10000)" not in c[0] or c[1].nx < 1024]
    used_peer = []
    for name, cfg, steps in cases:
        r = parity.check_case(name, cfg, steps, rank, world)
        if rank == 0:
            used_peer.append(r["peer_kernel"])
            print(f"PARITY mode={mode} peer_kernel={r['peer_kernel']} world={world} {cfg.nx}x{cfg.ny} [{name}] errs={r['rel_l2']} "
                  f"dt_equal={r['dt_equal']} iters={r['iters']} max_iter_diff={r['max_iter_diff']} "
                  f"{'OK' if r['ok'] else 'FAIL'}", flush=True)
            ok = ok and r["ok"]
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
