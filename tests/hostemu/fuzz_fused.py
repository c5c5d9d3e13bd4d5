"""Randomised comparison of the two fused per-step passes with the seven separate kernels ON THE HOST EMULATION
(developer tool, not part of the suite: the suite's fixed layouts are tests/test_kernels_on_host.py::test_fused_*).

    g++ -std=c++17 -O1 -ffp-contract=off -DPF_HOST_EMULATION -pthread -fPIC -shared -I tests/hostemu \
        -I pyflow-solver_b200/csrc -I include tests/hostemu/emu_kernels.cpp -o /tmp/fuzz/emu.so
    python tests/hostemu/fuzz_fused.py SEED SECONDS

Random widths around the warp / column-block boundaries (58, 116, 174, 464 columns), heights 4 .. 40, launch geometries
for 1 .. 148 SMs, unit and general spacing, the three advection settings, NaN / Inf poisons: the fields the fused
passes write must equal the separate kernels' bit for bit (NaN halos and padding untouched), the reductions to 1e-11.
Run of the final tree: seed 1, 420 s, 329 cases, all equal.
"""
import ctypes as C, numpy as np, sys, time
lib=C.CDLL('/tmp/fuzz/emu.so')
dp,d,i=C.POINTER(C.c_double),C.c_double,C.c_int
lib.emu_predict.argtypes=[i,i,i,dp,dp,dp,dp,d,d,d]
lib.emu_rhs1.argtypes=[i,i,dp,dp,dp,d,d,d,dp,dp]
lib.emu_resid.argtypes=[i,i,dp,dp,d,d,dp,dp]
lib.emu_correct.argtypes=[i,i,dp,dp,dp,d,d,d,d]
lib.emu_divnorm.argtypes=[i,i,dp,dp,d,d,dp,dp]
lib.emu_lid.argtypes=[i,i,dp,d]
lib.emu_post.argtypes=[i,i,dp,dp,dp,dp,dp,dp,dp,dp]
lib.emu_front.argtypes=[i,i,i,dp,dp,dp,dp,dp,d,d,d,d,i,dp,dp]
lib.emu_finish.argtypes=[i,i,i,dp,dp,dp,dp,dp,dp,d,d,d,i,d,dp,i,dp,dp]
lib.emu_slot.argtypes=[C.c_char_p]
lib.emu_fused_geometry.argtypes=[i,i,i,C.POINTER(C.c_int)]
def ptr(a): return a.ctypes.data_as(dp)
pitch=lambda nx: lib.emu_pitch(nx); H=lib.emu_halo()
def put(ny,nx,a=None):
    b=np.full((ny+2*H,pitch(nx)),np.nan); b[H:H+ny,:nx]=0.0 if a is None else a; return b
def get(b,ny,nx):
    assert np.isnan(b[:H]).all() and np.isnan(b[H+ny:]).all() and np.isnan(b[:,nx:]).all()
    return b[H:H+ny,:nx].copy()
same=lambda a,b: np.array_equal(a,b,equal_nan=True)
NS=lib.emu_scal_count(); NSC=lib.emu_scratch_doubles()
slot=lambda n: lib.emu_slot(n.encode())
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); n=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 300):
    nx=int(rng.choice([rng.integers(4,70), rng.integers(55,62), rng.integers(110,122), rng.integers(170,180), rng.integers(460,470), rng.integers(4,300)]))
    ny=int(rng.choice([rng.integers(4,9), rng.integers(4,40)]))
    sms=int(rng.choice([1,2,3,8,148]))
    unit=bool(rng.integers(0,2)); dx,dy=(1.0,1.0) if unit else (float(rng.uniform(0.1,3)),float(rng.uniform(0.1,3)))
    dt=float(rng.uniform(1e-3,5e-2)); scheme=int(rng.integers(0,3)); lid=1.25; post=int(rng.integers(0,2))
    geo=(C.c_int*4)(); assert lib.emu_fused_geometry(nx,ny,sms,geo)==1
    shp=(ny,nx)
    u,v,b,x,uo,vo=(rng.standard_normal(shp) for _ in range(6)); x[0,0]=0.0
    u[rng.random(shp)<0.1]=0.0; v[rng.random(shp)<0.1]=0.0
    if rng.random()<0.15: u[rng.integers(0,ny),rng.integers(0,nx)]=np.nan
    if rng.random()<0.1: v[rng.integers(0,ny),rng.integers(0,nx)]=np.inf
    S=np.zeros(NS); Ssc=np.zeros(NSC); F=np.zeros(NS); Fsc=np.zeros(NSC)
    with np.errstate(all='ignore'):
        du,dv_=put(ny,nx,u),put(ny,nx,v); us,vs,bb=put(ny,nx),put(ny,nx),put(ny,nx)
        lib.emu_predict(scheme,nx,ny,ptr(du),ptr(dv_),ptr(us),ptr(vs),dt,dx,dy)
        lib.emu_rhs1(nx,ny,ptr(us),ptr(vs),ptr(bb),dt,dx,dy,ptr(Ssc),ptr(S))
        fus,fvs,fb=put(ny,nx),put(ny,nx),put(ny,nx)
        assert lib.emu_front(scheme,nx,ny,ptr(du),ptr(dv_),ptr(fus),ptr(fvs),ptr(fb),dt,dt,dx,dy,sms,ptr(Fsc),ptr(F))==1
        ok=same(get(fus,ny,nx),get(us,ny,nx)) and same(get(fvs,ny,nx),get(vs,ny,nx)) and same(get(fb,ny,nx),get(bb,ny,nx))
        assert ok, ('front',nx,ny,sms,unit,scheme)
        # finish
        db,dxx,du2,dv2,duo,dvo=put(ny,nx,b),put(ny,nx,x),put(ny,nx,u),put(ny,nx,v),put(ny,nx,uo),put(ny,nx,vo)
        cen=np.full(ny,np.nan)
        lib.emu_resid(nx,ny,ptr(db),ptr(dxx),dx,dy,ptr(Ssc),ptr(S))
        lib.emu_correct(nx,ny,ptr(dxx),ptr(du2),ptr(dv2),dt,dx,dy,float(S[slot('S_X0')]))
        lib.emu_divnorm(nx,ny,ptr(du2),ptr(dv2),dx,dy,ptr(Ssc),ptr(S))
        if post:
            lib.emu_lid(nx,ny,ptr(du2),lid)
            lib.emu_post(nx,ny,ptr(du2),ptr(dv2),ptr(duo),ptr(dvo),ptr(dxx),ptr(cen),ptr(Ssc),ptr(S))
        fb2,fx,fu,fv,fuo,fvo=put(ny,nx,b),put(ny,nx,x),put(ny,nx,u),put(ny,nx,v),put(ny,nx,uo),put(ny,nx,vo)
        fcen=np.full(ny,np.nan)
        assert lib.emu_finish(post,nx,ny,ptr(fb2),ptr(fx),ptr(fu),ptr(fv),ptr(fuo),ptr(fvo),dt,dx,dy,1,lid,ptr(fcen),sms,ptr(Fsc),ptr(F))==1
        ok=same(get(fuo,ny,nx),get(du2,ny,nx)) and same(get(fvo,ny,nx),get(dv2,ny,nx)) and same(get(fx,ny,nx),get(dxx,ny,nx)) and (not post or same(fcen,cen))
        assert ok, ('finish',nx,ny,sms,unit,post)
        names=['S_RES_SUMSQ','S_DIVA_SUMSQ']+(['S_RU_SUMSQ','S_RV_SUMSQ','S_DIVC_SUMSQ','S_NONFINITE','S_UNAN','S_VNAN','S_UMAX','S_VMAX'] if post else [])
        for nm in names:
            a_,w_=F[slot(nm)],S[slot(nm)]
            assert (np.isnan(a_) and np.isnan(w_)) or a_==w_ or abs(a_-w_)<=1e-11*max(abs(a_),abs(w_)), (nm,a_,w_,nx,ny,sms)
    n+=1
print('fuzz cases', n, 'all equal')
