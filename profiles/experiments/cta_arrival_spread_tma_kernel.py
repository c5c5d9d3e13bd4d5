import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200'))
import numpy as np, torch
from types import SimpleNamespace
os.environ["PYFLOW_B200_OPTIONS"] = "cg_kernel=2,cg_sync_opt=4"
n = 4096
from pyflow.core.ghost_fields import allocate_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
cfg = SimpleNamespace(nx=n, ny=n, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
st = allocate_state(n, n); tr = ResidualManager()
for it in range(5): step(cfg, st, tr, it)
slab = st.fields._slab
G = 148
buf = np.zeros(1040); slab._lib.pf_debug_read_partials(slab.ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), 4096, 1040)
raw = buf.view(np.uint64)
print("TMA kernel cg_ms", round(st.meta['cg_ms'], 2))
exitA, exitB, exitPrev = int(raw[1024]), int(raw[1025]), int(raw[1026])
for kb, name, start, ex in ((0, "A", exitPrev, exitA), (1, "B", exitA, exitB)):
    v = raw[kb*512: kb*512+G]
    t = (v >> np.uint64(10)).astype(np.int64); sm = (v & np.uint64(1023)).astype(int)
    print(f" phase {name}: start->first arrival {(t.min()-start)/1e3:.1f} us, spread {(t.max()-t.min())/1e3:.1f} us, last arrival->exit {(ex-t.max())/1e3:.1f} us; total {(ex-start)/1e3:.1f}")
    tt = (t - t.min())/1e3
    bid = np.arange(G)
    print("   by chunk (bid%4):", np.round([tt[bid%4==c].mean() for c in range(4)],1), " pct10/50/90:", np.round(np.percentile(tt,[10,50,90]),1))
    order = np.argsort(tt); print("   slowest bids:", order[-8:], "sm:", sm[order[-8:]], np.round(tt[order[-8:]],1))
    print("   by die (sm<74 vs >=74):", round(tt[sm<74].mean(),1), round(tt[sm>=74].mean(),1), " by band third:", np.round([tt[(bid//4)//13==b].mean() for b in range(3)],1))
