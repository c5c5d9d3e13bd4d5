import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pyflow_oracle as orc
from pyflow._b200.device import slab_for_shape
def rel(a,b): return float(np.linalg.norm(a-b)/(np.linalg.norm(b) or 1))
for (ny,nx,its) in [(257,513,60),(257,513,5),(257,512,60),(257,577,60),(64,513,60),(300,513,1),(300,513,2)]:
    slab = slab_for_shape(ny, nx)
    rng = np.random.default_rng(ny+nx); b = rng.standard_normal((ny,nx)); b -= b.mean(); b[0,0]=0
    x1,i1,t1,r1 = slab.pcg(b,1.0,1.0,1e-10,its,jacobi=True,inv_diag_interior=0.25)
    x2,i2,t2,r2,bn = slab.cg_generic("pinned",b,1.0,1.0,1e-10,its,jacobi=True,inv_diag_interior=0.25)
    inv = orc.jacobi_inverse_diagonal(nx,ny,1.0,1.0)
    xo,io,to = orc.pcg(lambda z: orc.neg_laplacian_pinned(z,nx,ny,1.0,1.0), b.ravel().copy(), 1e-10, its, inv)
    xo = xo.reshape(ny,nx)
    d = np.abs(x1-xo); j,i = np.unravel_index(np.argmax(d), d.shape)
    print(f"{ny}x{nx} its={its}: tuned-vs-oracle {rel(x1,xo):.2e} plain-vs-oracle {rel(x2,xo):.2e} rn {r1:.6e} {r2:.6e}; worst cell (j={j}, i={i}) {d[j,i]:.2e}; bad cols:", np.unique(np.where(d > 1e-9*np.abs(xo).max())[1])[:10])
