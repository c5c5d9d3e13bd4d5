import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200')); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from types import SimpleNamespace
os.environ.setdefault("MASTER_ADDR","127.0.0.1"); os.environ.setdefault("MASTER_PORT","29811"); os.environ.setdefault("RANK","0"); os.environ.setdefault("WORLD_SIZE","1")
mode = sys.argv[1]
torch.cuda.set_device(0)
from pyflow._b200.device import DeviceSlab
from pyflow._b200 import device as D
if mode == "symm":
    dist.init_process_group("nccl", device_id=torch.device("cuda:0"))
    from pyflow._b200.distributed import _symmetric_allocator
    orig = DeviceSlab.__init__
    def patched(self, *a, **k):
        k.setdefault("allocator", _symmetric_allocator); orig(self, *a, **k)
    DeviceSlab.__init__ = patched
from pyflow.core.ghost_fields import allocate_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
cfg = SimpleNamespace(nx=4096, ny=4096, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
st = allocate_state(4096,4096); tr = ResidualManager(); ms=[]
for it in range(8):
    step(cfg, st, tr, it)
    if it>=3: ms.append(st.meta["cg_ms"])
print(mode, os.environ.get("PYFLOW_B200_OPTIONS"), "cg_ms", statistics.mean(ms))
