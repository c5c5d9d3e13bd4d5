import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200'))
import numpy as np, torch
from types import SimpleNamespace
os.environ["PYFLOW_B200_OPTIONS"] = sys.argv[1]
n = 4096
from pyflow.core.ghost_fields import allocate_state
from pyflow.residuals.manager import ResidualManager
from pyflow.solvers.solver import step
cfg = SimpleNamespace(nx=n, ny=n, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10, lin_maxiter=300, advection_scheme="quick", disable_advection=False, diagnostics=False)
st = allocate_state(n, n); tr = ResidualManager()
for it in range(5): step(cfg, st, tr, it)
slab = st.fields._slab
G = slab.info()["cg_grid_blocks"]
buf = np.zeros(1024); slab._lib.pf_debug_read_partials(slab.ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), 4096, 1024)
raw = buf.view(np.uint64).reshape(2, 512)[:, :G]
print(sys.argv[1], "cg_ms", round(st.meta['cg_ms'], 2), "G", G)
for kb, name in enumerate(("A", "B")):
    t = (raw[kb] >> np.uint64(10)).astype(np.int64); sm = (raw[kb] & np.uint64(1023)).astype(int)
    t = (t - t.min()) / 1e3
    print(f" phase {name}: arrival spread {t.max():.1f} us; mean {t.mean():.1f}; pct10 {np.percentile(t,10):.1f} pct50 {np.percentile(t,50):.1f} pct90 {np.percentile(t,90):.1f}")
    bid = np.arange(G)
    # static kernel: tile = bid: chunk = bid % 8, band = bid // 8
    by_chunk = [t[bid % 8 == c].mean() for c in range(8)]
    print("   mean arrival by chunk (bid%8):", np.round(by_chunk, 1))
    by_band = [t[(bid // 8) // 5 == b].mean() for b in range(8)]
    print("   mean arrival by band group:", np.round(by_band, 1))
    order = np.argsort(t)
    print("   slowest 12 bids:", order[-12:], "their sm:", sm[order[-12:]], "t:", np.round(t[order[-12:]], 1))
    print("   fastest 8 bids:", order[:8], "sm:", sm[order[:8]])
    # by SM: pair mates
    tsm = {}
    for b in range(G): tsm.setdefault(sm[b], []).append(t[b])
    diffs = [abs(v[0]-v[1]) for v in tsm.values() if len(v) == 2]
    print(f"   CTAs sharing an SM: mean |dt| {np.mean(diffs):.1f} us; corr by smid parity: even {t[sm%2==0].mean():.1f} odd {t[sm%2==1].mean():.1f}; sm<74 {t[sm<74].mean():.1f} sm>=74 {t[sm>=74].mean():.1f}")
