"""Timeline of one iteration (#100) of the on-chip CG kernel k_cg_small at 512^2: globaltimer stamps of every CTA
(pf_set_option cg_sync_opt = 4): iteration start, compute A done, exchange 1 {posted, all packets seen, done},
phase B done, exchange 2 {posted, all seen}."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pyflow-solver_b200"))
os.environ["PYFLOW_B200_OPTIONS"] = "cg_sync_opt=4"
from types import SimpleNamespace
from pyflow.core.ghost_fields import allocate_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = SimpleNamespace(nx=n, ny=n, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=300,
                      advection_scheme="quick", disable_advection=False, diagnostics=False)
st, tr = allocate_state(n, n), ResidualManager()
for it in range(4):
    step(cfg, st, tr, it)
slab = st.fields._slab
G = (n + (-(-n // 148)) - 1) // (-(-n // 148))
buf = np.zeros(G * 8)
slab._lib.pf_debug_read_partials(slab.ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), 0, G * 8)
t = buf.view(np.uint64).reshape(G, 8).astype(np.int64)
t0 = t[:, 0].min()
names = ["start", "computeA done", "x1 posted", "x1 all seen", "x1 done", "phaseB done", "x2 posted", "x2 all seen"]
print(f"k_cg_small {n}x{n}: {G} CTAs, cg_ms {st.meta['cg_ms']:.3f} ({st.meta['cg_ms'] * 1e3 / 300:.2f} us per iteration)")
for k, name in enumerate(names):
    c = (t[:, k] - t0) / 1e3
    print(f"  {name:14s}: min {c.min():6.2f}  median {np.median(c):6.2f}  max {c.max():6.2f} us")
seg = np.diff(t, axis=1) / 1e3
for k in range(7):
    print(f"  {names[k]:14s} -> {names[k + 1]:14s}: median {np.median(seg[:, k]):5.2f}  max {seg[:, k].max():5.2f} us")
