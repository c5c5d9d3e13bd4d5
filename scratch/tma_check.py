import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'pyflow-solver_b200')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy as np
import pyflow_oracle as orc
from pyflow._b200.device import DeviceSlab
def rel(a,b): return float(np.linalg.norm(a-b)/(np.linalg.norm(b) or 1))
for (ny,nx,its,unit) in [(160,1024,3,True),(300,2048,25,True),(151,4096,25,True),(200,1024,25,False),(17,1024,40,True),(33,3072,10,True)]:
    slab = DeviceSlab(nx, ny)
    slab.set_option("cg_kernel", int(os.environ.get("CGK", "2")))
    rng = np.random.default_rng(ny+nx); b = rng.standard_normal((ny,nx)); b -= b.mean(); b[0,0]=0
    dx, dy = (1.0,1.0) if unit else (0.5,0.25)
    invd = 1.0/(2.0/dx**2+2.0/dy**2)
    x1,i1,t1,r1 = slab.pcg(b,dx,dy,1e-10,its,jacobi=True,inv_diag_interior=invd)
    inv = orc.jacobi_inverse_diagonal(nx,ny,dx,dy)
    xo,io,to = orc.pcg(lambda z: orc.neg_laplacian_pinned(z,nx,ny,dx,dy), b.ravel().copy(), 1e-10, its, inv)
    print(f"{ny}x{nx} its={its} unit={unit}: info {i1} iters {t1}; tuned(TMA)-vs-oracle {rel(x1.ravel(),xo):.2e}", flush=True)
    slab.close()
