import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200'))
import numpy as np, torch
from types import SimpleNamespace
os.environ["PYFLOW_B200_OPTIONS"] = "cg_sync_opt=4"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
from pyflow.core.ghost_fields import allocate_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
cfg = SimpleNamespace(nx=n, ny=n, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
st = allocate_state(n, n); tr = ResidualManager()
for it in range(5): step(cfg, st, tr, it)
slab = st.fields._slab
buf = np.zeros(64); slab._lib.pf_debug_read_partials(slab.ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), 4 * 16384, 64)
t = buf.view(np.uint64).reshape(8, 8).astype(np.int64)
print(f"n={n} cg_ms {st.meta['cg_ms']:.2f}")
prev_exit = None
for k in range(8):
    first, last, fold, pub, seen, rel, b0 = t[k][:7]
    print(f"  barrier {k}: spread {(last-first)/1e3:7.2f} us | fold {(fold-last)/1e3:6.2f} | publish {(pub-fold)/1e3:6.2f} | seen {(seen-pub)/1e3:6.2f} | release {(rel-seen)/1e3:6.2f} | blk0 out {(b0-rel)/1e3:6.2f} | phase to first arrival {((first-prev_exit)/1e3 if prev_exit else 0):7.2f}", flush=True)
    prev_exit = b0
