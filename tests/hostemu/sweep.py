"""Exploratory sweep of the CG kernels on the host emulation (developer tool, not collected by pytest).

    python tests/hostemu/sweep.py

Builds the emulation library like tests/test_kernels_on_host.py does and runs every emulated CG kernel over layouts
picked to hit classification edges: widths around multiples of 64 and 512 for the register-window and split-phase
kernels, band / width combinations for the shared-memory-resident kernel, tile heights above / below / not dividing the
slab height for the TMA-fed kernel, two-row slabs and up to 8 ranks for the multi-GPU kernels.  Every case is compared
with scipy's recurrence on the whole grid (relative error < 1e-10 after 3 iterations).  Round 1: 203 cases, 0 failures.
"""
import sys, os, ctypes as C, numpy as np, time, subprocess
sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import pyflow_oracle as orc
import test_kernels_on_host as T
import tempfile
build=tempfile.mkdtemp(prefix='pf_hostemu_')
T.generate_small_kernel_copy(build); T.generate_tma_kernel_copy(build); T.generate_multi_kernel_copy(build)
ROOT='/root/repo'
out=os.path.join(build,'emu.so')
cmd=["g++","-std=c++17","-O1","-ffp-contract=off","-DPF_HOST_EMULATION","-DPF_EMU_WITH_SMALL","-DPF_EMU_WITH_TMA","-DPF_EMU_WITH_MULTI","-pthread","-fPIC","-shared","-I",build,"-I",ROOT+"/tests/hostemu","-I",ROOT+"/pyflow-solver_b200/csrc","-I",ROOT+"/include",ROOT+"/tests/hostemu/emu_kernels.cpp","-o",out]
subprocess.run(cmd,check=True)
lib=C.CDLL(out)
dp,d,i=C.POINTER(C.c_double),C.c_double,C.c_int
lib.emu_slot.argtypes=[C.c_char_p]
lib.emu_rhs2.argtypes=[i,i,dp,dp,dp,i,i,d,dp,dp]
lib.emu_cg_persistent.argtypes=[i,i,dp,dp,dp,dp,dp,d,d,d,d,i,i,dp,dp]
lib.emu_cg_split.argtypes=[i,i,i,dp,dp,dp,dp,dp,d,d,d,d,i,i,dp,dp,dp]
ptr=T.ptr
bad=[]
def check(kind,ny,nx,G,unit,maxiter):
    rng=np.random.default_rng(ny*1009+nx)
    dx,dy=(1.0,1.0) if unit else (0.5,0.25); invd=1.0/(2.0/dx**2+2.0/dy**2)
    b=rng.standard_normal((ny,nx)); b-=b.mean(); b[0,0]=0
    dev=T.Dev(lib,ny,nx); scal=np.zeros(lib.emu_scal_count()); scratch=np.zeros(lib.emu_scratch_doubles())
    bb,x,r,p0,p1=dev.put(b),dev.put(),dev.put(),dev.put(),dev.put()
    lib.emu_rhs2(nx,ny,ptr(bb),ptr(x),ptr(r),0,0,invd,ptr(scratch),ptr(scal))
    part=np.zeros(lib.emu_cg_part_doubles())
    if kind=="pers": lib.emu_cg_persistent(nx,ny,ptr(bb),ptr(x),ptr(r),ptr(p0),ptr(p1),dx,dy,invd,1e-10,maxiter,4,ptr(part),ptr(scal))
    else: lib.emu_cg_split(nx,ny,G,ptr(bb),ptr(x),ptr(r),ptr(p0),ptr(p1),dx,dy,invd,1e-10,maxiter,4,ptr(part),ptr(scal),ptr(scratch))
    xo,io,to=orc.pcg(lambda z: orc.neg_laplacian_pinned(z,nx,ny,dx,dy), b.ravel().copy(),1e-10,maxiter,orc.jacobi_inverse_diagonal(nx,ny,dx,dy))
    try:
        got=dev.get(x).ravel(); err=np.linalg.norm(got-xo)/np.linalg.norm(xo)
    except AssertionError: err=float('inf')
    ok = err<1e-10 and int(scal[lib.emu_slot(b"S_INFO")])==io
    if not ok: bad.append((kind,ny,nx,G,unit,err)); print("BAD",kind,ny,nx,G,unit,err,flush=True)
t=time.time()
for nx in [2,3,4,5,31,32,33,62,63,64,65,66,67,95,96,97,126,127,128,129,130,131,190,191,192,193,194,255,256,257,300,511,512,513,514,575,576,577,578,640,641,1023,1025]:
    for ny in (2,3,5):
        check("pers",ny,nx,1,(nx+ny)%2==0,3)
print("persistent sweep done",f"{time.time()-t:.0f}s","bad:",len(bad),flush=True)
t=time.time()
for nx in [17,64,65,129,130,193,257,513,577,641]:
    for (ny,G) in ((70,2),(40,3),(9,2)):
        check("split",ny,nx,G,(nx+ny)%2==1,3)
print("split sweep done",f"{time.time()-t:.0f}s","bad:",len(bad),flush=True)

ppd=C.POINTER(dp)
lib.emu_cg_small.argtypes=[i,i,i,dp,dp,dp,dp,dp,C.POINTER(C.c_ulonglong),d,d,d,d,i]
lib.emu_cg_tmad.argtypes=[i,i,i,i,i,dp,dp,dp,dp,dp,d,d,d,d,i,dp,dp,C.POINTER(C.c_ulonglong)]
lib.emu_cg_tmad_multi.argtypes=[i,i,i,i,i,i,ppd,ppd,ppd,ppd,ppd,d,d,d,d,i,ppd,ppd,C.POINTER(C.POINTER(C.c_ubyte))]
lib.emu_cg_persistent_multi.argtypes=lib.emu_cg_tmad_multi.argtypes
ptr=T.ptr
bad=[]
def ref(b,dx,dy,maxiter):
    ny,nx=b.shape
    return orc.pcg(lambda z: orc.neg_laplacian_pinned(z,nx,ny,dx,dy), b.ravel().copy(),1e-10,maxiter,orc.jacobi_inverse_diagonal(nx,ny,dx,dy))
def mk(ny,nx):
    rng=np.random.default_rng(ny*1009+nx); b=rng.standard_normal((ny,nx)); b-=b.mean(); b[0,0]=0; return b
def small(ny,nx,rows,unit,maxiter=3):
    dx,dy=(1.0,1.0) if unit else (0.5,0.25); invd=1.0/(2.0/dx**2+2.0/dy**2); b=mk(ny,nx)
    dev=T.Dev(lib,ny,nx); S=T.Scalars(lib); sol=T.SmallSolver(lib,ny,nx,rows)
    bb,x,r=dev.put(b),dev.put(),dev.put()
    lib.emu_rhs2(nx,ny,ptr(bb),ptr(x),ptr(r),0,0,invd,ptr(S.scratch),ptr(S.scal))
    sol.solve(bb,x,S.scal,dx,dy,invd,1e-10,maxiter)
    xo,io,to=ref(b,dx,dy,maxiter); err=np.linalg.norm(dev.get(x).ravel()-xo)/np.linalg.norm(xo)
    if not err<1e-10: bad.append(("small",ny,nx,rows,err)); print("BAD small",ny,nx,rows,err,flush=True)
def tmad(ny,nx,tr,unit,maxiter=3):
    dx,dy=(1.0,1.0) if unit else (0.5,0.25); invd=1.0/(2.0/dx**2+2.0/dy**2); b=mk(ny,nx)
    dev=T.Dev(lib,ny,nx); S=T.Scalars(lib); sync=np.zeros(256,dtype=np.uint64)
    x=T._tmad_solve(lib,dev,S,sync,b,dx,dy,invd,1e-10,maxiter,tr)
    xo,io,to=ref(b,dx,dy,maxiter); err=np.linalg.norm(x.ravel()-xo)/np.linalg.norm(xo)
    if not err<1e-10: bad.append(("tmad",ny,nx,tr,err)); print("BAD tmad",ny,nx,tr,err,flush=True)
def multi(kind,R,ny,nx,tr,unit,maxiter=3):
    dx,dy=(1.0,1.0) if unit else (0.5,0.25); b=mk(ny,nx)
    ranks=T.Ranks(lib,R,ny,nx)
    x,info,iters=ranks.solve(b,dx,dy,True,1e-10,maxiter,tr,kernel=kind)
    xo,io,to=ref(b,dx,dy,maxiter); err=np.linalg.norm(x.ravel()-xo)/np.linalg.norm(xo)
    if not err<1e-10: bad.append((kind,R,ny,nx,tr,err)); print("BAD",kind,R,ny,nx,tr,err,flush=True)
t=time.time()
for (ny,nx,rows) in [(2,2,1),(2,2,2),(3,2,1),(3,8,2),(4,16,3),(5,64,2),(3,66,1),(2,128,1),(4,130,2),(3,256,1),(2,510,1),(2,512,2),(3,1024,1),(2,2048,1),(6,4,4),(7,12,7),(5,30,1)]:
    small(ny,nx,rows,(ny+nx)%3!=0)
print("small sweep done",f"{time.time()-t:.0f}s bad:",len(bad),flush=True)
t=time.time()
for (ny,nx,tr) in [(2,1024,4),(2,1024,2),(3,1024,2),(4,1024,4),(5,1024,4),(7,1024,2),(17,1024,16),(16,1024,16),(2,3072,4),(3,2048,2),(9,2048,8),(6,1024,8),(13,1024,(4,1,2)),(20,2048,(4,2,2)),(40,1024,(8,2,3)),(9,1024,(4,1,1))]:
    tmad(ny,nx,tr,(ny+(tr if isinstance(tr,int) else tr[0]))%2==0)
print("tmad sweep done",f"{time.time()-t:.0f}s bad:",len(bad),flush=True)
t=time.time()
for (R,ny,nx,tr) in [(2,4,1024,4),(2,4,1024,2),(3,6,1024,4),(2,10,1024,4),(4,8,1024,2),(2,6,2048,2),(8,16,1024,2),(2,28,1024,(4,1,2)),(3,36,1024,(4,1,1))]:
    multi("tmad",R,ny,nx,tr,(R+ny)%2==0)
for (R,ny,nx) in [(2,4,2),(2,4,3),(3,6,64),(2,8,65),(4,8,129),(2,6,577),(8,16,33),(5,10,70)]:
    multi("register-window",R,ny,nx,4,(R+nx)%2==0)
print("multi sweep done",f"{time.time()-t:.0f}s bad:",len(bad),flush=True)
print("ALL BAD:",bad)
