import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200'))
import numpy as np, torch, torch.distributed as dist
from types import SimpleNamespace
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{torch.cuda.current_device()}"))
os.environ["PYFLOW_B200_OPTIONS"] = "cg_sync_opt=4"
from pyflow._b200.distributed import allocate_slab_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
NX = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
MAXIT = int(sys.argv[3]) if len(sys.argv) > 3 else 300
cfg = SimpleNamespace(nx=NX, ny=ROWS*world, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=MAXIT, advection_scheme="quick", disable_advection=False, diagnostics=False)
st = allocate_slab_state(NX, ROWS*world, rank, world); tr = ResidualManager()
for it in range(4): step(cfg, st, tr, it)
slab = st.fields._slab
buf = np.zeros(64); slab._lib.pf_debug_read_partials(slab.ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), 4 * 16384, 64)
t = buf.view(np.uint64).reshape(8, 8).astype(np.int64)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"rank {rank} grid {NX}x{ROWS}/gpu cg_ms {st.meta['cg_ms']:.2f}  ({st.meta['cg_ms']*1e3/MAXIT:.1f} us per iteration)")
        prev_exit = None
        for k in range(8):
            first, last, fold, pub, seen0, seen1, b0 = t[k][:7]
            print(f"  barrier {k}: spread first->last arrival {(last-first)/1e3:7.2f} us | fold {(fold-last)/1e3:6.2f} | publish {(pub-fold)/1e3:6.2f} | blk0 saw rank0 pkt {(seen0-pub)/1e3:6.2f} rank1 pkt {(seen1-pub)/1e3:6.2f} | blk0 out {(b0-max(seen0,seen1))/1e3:6.2f} | total last-arrival->out {(b0-last)/1e3:6.2f} | phase since prev exit {((first-prev_exit)/1e3 if prev_exit else 0):7.2f}", flush=True)
            prev_exit = b0
dist.destroy_process_group()
