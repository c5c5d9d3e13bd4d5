"""Oracle-backed stand-in for ``pyflow._b200.device.DeviceSlab`` (TEST DOUBLE).

The product has no CPU path: without the CUDA library and a GPU every entry
point raises.  To exercise the HOST side of the drop-in on a machine without a
GPU -- the dt / safeguard floats of ``solvers/solver.py``, the diagnostics
dictionaries, the lazy ``State`` synchronisation protocol, the driver, the CLI
payload, checkpoints of device-resident fields -- the CPU test-suite swaps the
GPU context for this class, which answers every ``DeviceSlab`` call with the
numpy oracle (``oracle/pyflow_oracle.py``).  It lives under ``tests/`` and is
never imported by the package; nothing measured or shipped goes through it.

The "device" copies of u, v, p are private arrays: host arrays handed to
``upload`` are copied, ``download`` copies back, exactly like the HBM fields.
"""
from __future__ import annotations

import numpy as np
import pyflow_oracle as orc

from pyflow._b200 import lib as _lib


def _norm2(a) -> float:
    a = np.asarray(a).ravel()
    return float(np.dot(a, a))


class CpuSlab:
    instances = 0

    def __init__(self, nx, ny_local, ny_global=None, row_offset=0, device=None, allocator=None):
        if nx < 2 or ny_local < 2:
            raise ValueError(f"unsupported grid {nx}x{ny_local} (need nx >= 2, ny >= 2)")
        self.nx, self.ny_local = int(nx), int(ny_local)
        self.ny_global = int(ny_global if ny_global is not None else ny_local)
        assert self.ny_global == self.ny_local and row_offset == 0, "the stand-in models one GPU"
        self.row_offset, self.nranks, self.rank, self.peer_mode = 0, 1, 0, False
        self.u = np.zeros((self.ny_local, self.nx))
        self.v = np.zeros_like(self.u)
        self.p = np.zeros_like(self.u)
        self.uploads: list[tuple[str, ...]] = []       # which fields each upload() carried
        self.downloads: list[tuple[str, ...]] = []
        self.steps = 0
        self._launches = 0
        CpuSlab.instances += 1

    # ------------------------------------------------------------------ plumbing
    def close(self):
        pass

    def info(self):
        return {"sm_count": 0, "cg_grid_blocks": 0, "cg_block_threads": 0}

    def set_option(self, key, value):
        pass

    def launch_count(self):
        return self._launches

    def _check(self, a):
        if a is not None and (a.dtype != np.float64 or not a.flags.c_contiguous
                              or a.shape != (self.ny_local + 2, self.nx + 2)):
            raise ValueError(f"field must be C-contiguous float64 of shape {(self.ny_local + 2, self.nx + 2)}")

    def upload(self, u=None, v=None, p=None):
        got = []
        for name, a in (("u", u), ("v", v), ("p", p)):
            self._check(a)
            if a is not None:
                getattr(self, name)[:] = a[1:-1, 1:-1]
                got.append(name)
        self.uploads.append(tuple(got))

    def download(self, u=None, v=None, p=None):
        got = []
        for name, a in (("u", u), ("v", v), ("p", p)):
            self._check(a)
            if a is not None:
                a[1:-1, 1:-1] = getattr(self, name)
                got.append(name)
        self.downloads.append(tuple(got))

    # ------------------------------------------------------------------ step
    def prestep(self):
        return (float(np.max(np.abs(self.u))), float(np.max(np.abs(self.v))),
                _norm2(orc.divergence_central(self.u, self.v, 1.0, 1.0)))

    def step(self, params) -> _lib.StepOut:
        """pf_step: predictor (unless projection_only), projection, lid row, residual sums."""
        out = _lib.StepOut()
        dt, dx, dy = params.dt, params.dx, params.dy
        ny, nx = self.u.shape
        u_old, v_old = self.u.copy(), self.v.copy()
        if not params.projection_only and params.scheme != _lib.PF_ADV_NONE:
            adv = orc.advect_quick if params.scheme == _lib.PF_ADV_QUICK else orc.advect_upwind
            cu = adv(u_old, v_old, u_old, dx, dy)
            cv = adv(u_old, v_old, v_old, dx, dy)
            self.u[:] = u_old - dt * cu
            self.v[:] = v_old - dt * cv
        div_u = orc.divergence_coherent(self.u, self.v, dx, dy)
        out.div_before_norm = float(np.sqrt(_norm2(div_u)))
        rhs = (-div_u / max(dt, 1e-14)).reshape(-1)
        mean = float(np.mean(rhs))
        rhs -= mean
        rhs[0] = 0.0
        out.rhs_mean = mean
        inv_diag = None
        if params.jacobi:
            inv_diag = np.full((ny, nx), params.inv_diag_interior)
            inv_diag[0, :] = inv_diag[-1, :] = 1.0
            inv_diag[:, 0] = inv_diag[:, -1] = 1.0
            inv_diag = inv_diag.reshape(-1)

        def A(xv):
            return orc.neg_laplacian_pinned(xv, nx, ny, dx, dy)

        x, info, iters = orc.pcg(A, rhs, params.lin_tol, params.lin_maxiter, inv_diag)
        res = rhs - A(x)
        out.bnorm = float(np.sqrt(_norm2(rhs)))
        out.rnorm = float(np.sqrt(_norm2(res)))
        out.info, out.true_iters = int(info), int(iters)
        pn = x.reshape(ny, nx).copy()
        pn -= pn.flat[0]
        self.p[:] = pn
        gx, gy = orc.gradient_coherent(self.p, dx, dy)
        self.u -= dt * gx
        self.v -= dt * gy
        out.div_after_norm = float(np.sqrt(_norm2(orc.divergence_coherent(self.u, self.v, dx, dy))))
        if not params.projection_only:
            if params.lid_set:
                self.u[-1, :] = params.lid_velocity
            du, dv = (self.u - u_old).ravel(), (self.v - v_old).ravel()
            out.ru_sumsq = float(np.sum(du * du))
            out.rv_sumsq = float(np.sum(dv * dv))
        out.cont_sumsq = _norm2(orc.divergence_central(self.u, self.v, 1.0, 1.0))
        out.umax_raw = float(np.max(np.abs(self.u)))
        out.vmax_raw = float(np.max(np.abs(self.v)))
        out.p00 = float(self.p[0, 0])
        out.nonfinite = int(not all(np.isfinite(a).all() for a in (self.u, self.v, self.p)))
        out.cg_ms, out.step_ms = 1.0, 1.25                 # placeholders (nothing is timed here)
        self.steps += 1
        self._launches += 9
        return out

    def divergence_stats(self, dx, dy):
        d = orc.divergence_central(self.u, self.v, dx, dy)
        m = float(d.mean())
        return m, float(np.linalg.norm(d - m))

    def centerline(self):
        return self.u[:, self.nx // 2].copy()

    # ------------------------------------------------------------------ operator seams
    def _dense(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (self.ny_local, self.nx):
            raise ValueError(f"expected shape {(self.ny_local, self.nx)}, got {a.shape}")
        return a

    def gradient_coherent(self, p, dx, dy):
        return orc.gradient_coherent(self._dense(p), dx, dy)

    def divergence_coherent(self, u, v, dx, dy):
        return orc.divergence_coherent(self._dense(u), self._dense(v), dx, dy)

    def divergence_central(self, u, v, dx, dy):
        return orc.divergence_central(self._dense(u), self._dense(v), dx, dy)

    def laplacian_coherent(self, p, dx, dy):
        return orc.laplacian_coherent(self._dense(p), dx, dy)

    def gradient_central(self, p, dx, dy):
        return orc.gradient_central(self._dense(p), dx, dy)

    def laplacian_central(self, p, dx, dy):
        return orc.laplacian_central(self._dense(p), dx, dy)

    def pressure_correction_central(self, u, v, p, dt, dx, dy):
        return orc.pressure_correction_central(self._dense(u), self._dense(v), self._dense(p), dt, dx, dy)

    def advect(self, scheme, u, v, f, dx, dy):
        fn = {"upwind": orc.advect_upwind, "quick": orc.advect_quick}[scheme]
        return fn(self._dense(u), self._dense(v), self._dense(f), dx, dy)

    def pcg(self, b, dx, dy, tol, maxiter, jacobi=True, inv_diag_interior=None, _krylov=None):
        b = self._dense(b)
        ny, nx = b.shape
        inv_diag = None
        if jacobi:
            if inv_diag_interior is None:
                inv_diag_interior = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
            inv_diag = np.full((ny, nx), inv_diag_interior)
            inv_diag[0, :] = inv_diag[-1, :] = 1.0
            inv_diag[:, 0] = inv_diag[:, -1] = 1.0
            inv_diag = inv_diag.reshape(-1)

        def A(xv):
            return orc.neg_laplacian_pinned(xv, nx, ny, dx, dy)

        rhs = b.reshape(-1).copy()
        x, info, iters = (_krylov or orc.pcg)(A, rhs, tol, maxiter, inv_diag)
        x = np.array(x, dtype=np.float64)
        rnorm = float(np.sqrt(_norm2(rhs - A(x))))
        self.p[:] = x.reshape(ny, nx)
        return x.reshape(ny, nx), int(info), int(iters), rnorm

    def krylov_diag(self, method, b, inv_diag, dx, dy, tol, maxiter):
        b, d = self._dense(b), self._dense(inv_diag)
        ny, nx = b.shape

        def A(xv):
            return orc.neg_laplacian_pinned(xv, nx, ny, dx, dy)

        rhs = b.reshape(-1).copy()
        x, info, iters = (orc.pcg if method == "cg" else orc.bicgstab)(A, rhs, tol, maxiter, d.reshape(-1).copy())
        x = np.array(x, dtype=np.float64)
        self.p[:] = x.reshape(ny, nx)
        return x.reshape(ny, nx), int(info), int(iters), float(np.sqrt(_norm2(rhs - A(x))))

    def krylov_stencil5(self, method, b, coef, tol, maxiter, jacobi=False, inv_diag_interior=1.0):
        """The assembled ring-identity 5-point matrix (reference pressure_solver.py:15-40) through scipy's own CSR
        matvec, so the stand-in has the arithmetic of the reference's `solve(A, b)` on that matrix."""
        import scipy.sparse as sp
        b = self._dense(b)
        ny, nx = b.shape
        n = nx * ny
        jj, ii = np.divmod(np.arange(n), nx)
        ring = (ii == 0) | (ii == nx - 1) | (jj == 0) | (jj == ny - 1)
        inner = np.flatnonzero(~ring)
        rows = np.concatenate([np.flatnonzero(ring), np.repeat(inner, 5)])
        cols = np.concatenate([np.flatnonzero(ring), (inner[:, None] + np.array([-nx, -1, 0, 1, nx])[None, :]).reshape(-1)])
        vals = np.concatenate([np.ones(int(ring.sum())), np.tile(np.asarray(coef, dtype=np.float64), inner.size)])
        Amat = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
        Amat.sort_indices()
        inv_diag = np.where(ring, 1.0, inv_diag_interior) if jacobi else None
        rhs = b.reshape(-1).copy()
        x, info, iters = (orc.pcg if method == "cg" else orc.bicgstab)(lambda v: Amat @ v, rhs, tol, maxiter, inv_diag)
        x = np.array(x, dtype=np.float64)
        self.p[:] = x.reshape(ny, nx)
        return x.reshape(ny, nx), int(info), int(iters), float(np.sqrt(_norm2(rhs - Amat @ x)))

    def bicgstab(self, b, dx, dy, tol, maxiter, jacobi=False, inv_diag_interior=None):
        return self.pcg(b, dx, dy, tol, maxiter, jacobi, inv_diag_interior, _krylov=orc.bicgstab)

    def cg_generic(self, mode, b, dx, dy, tol, maxiter, jacobi=False, inv_diag_interior=1.0, want_x=True):
        b = self._dense(b)
        ny, nx = b.shape
        if mode == "pinned":
            x, info, iters, rnorm = self.pcg(b, dx, dy, tol, maxiter, jacobi, inv_diag_interior)
            return (x if want_x else None), info, iters, rnorm, float(np.sqrt(_norm2(b)))
        mask = np.zeros((ny, nx), dtype=bool)
        mask[0, :] = mask[-1, :] = True
        mask[:, 0] = mask[:, -1] = True
        ring = np.flatnonzero(mask.reshape(-1))

        def A(xv):
            o = orc.laplacian_coherent(xv.reshape(ny, nx), dx, dy).reshape(-1)
            o[ring] = xv[ring]
            return o

        rhs = b.reshape(-1).copy()
        x, info, iters = orc.pcg(A, rhs, tol, maxiter, None)
        x = np.array(x, dtype=np.float64)
        rnorm = float(np.sqrt(_norm2(rhs - A(x))))
        self.p[:] = x.reshape(ny, nx)
        return (x.reshape(ny, nx).copy() if want_x else None), int(info), int(iters), rnorm, float(np.sqrt(_norm2(rhs)))


class GatherSlab(CpuSlab):
    """Stand-in for one rank's slab context of a multi-GPU run (host-logic tests under gloo).

    Every call that the real library answers with halo exchanges and all-reduces is answered here by gathering
    the whole grid on every rank (``torch.distributed.all_gather_object``), running the single-grid stand-in and
    keeping this rank's rows -- so all ranks hold identical step scalars, exactly as with the real library."""

    def __init__(self, nx, ny_local, ny_global, row_offset, rank, world):
        CpuSlab.__init__(self, nx, ny_local)
        self.ny_global, self.row_offset, self.rank, self.nranks = int(ny_global), int(row_offset), int(rank), int(world)

    def _gathered(self):
        import torch.distributed as dist
        parts = [None] * self.nranks
        dist.all_gather_object(parts, (self.u, self.v, self.p))
        full = CpuSlab(self.nx, self.ny_global)
        for k, name in enumerate(("u", "v", "p")):
            getattr(full, name)[:] = np.concatenate([t[k] for t in parts], axis=0)
        return full

    def _keep_mine(self, full):
        rows = slice(self.row_offset, self.row_offset + self.ny_local)
        for name in ("u", "v", "p"):
            getattr(self, name)[:] = getattr(full, name)[rows]

    def prestep(self):
        return self._gathered().prestep()

    def step(self, params):
        full = self._gathered()
        out = full.step(params)
        self._keep_mine(full)
        self.steps += 1
        return out

    def divergence_stats(self, dx, dy):
        return self._gathered().divergence_stats(dx, dy)


def install_slabs(rank, world):
    """Permanent swap (one process per rank under torchrun): slab states get a ``GatherSlab``."""
    import pyflow._b200.distributed as distributed

    def make_slab(state):
        spec = state.slab
        return GatherSlab(state.nx, state.ny, spec["ny_global"], spec["row_offset"], rank, world)

    distributed.make_slab = make_slab
    install(None)


def install(monkeypatch=None):
    """Swap the GPU context for the stand-in in every module that binds it.  With a pytest
    ``monkeypatch`` the swap is undone at the end of the test; with ``None`` it is permanent
    (used from a ``sitecustomize`` for CLI subprocesses of host-logic tests)."""
    import importlib

    import pyflow._b200.device as device
    import pyflow._b200.engine as engine
    cache: dict = {}

    def slab_for_shape(ny, nx):
        key = (int(ny), int(nx))
        if key not in cache:
            cache[key] = CpuSlab(nx, ny)
        return cache[key]

    def put(mod, name, value):
        if monkeypatch is not None:
            monkeypatch.setattr(mod, name, value)
        else:
            setattr(mod, name, value)

    # import the modules that bind `slab_for_shape` by name BEFORE anything is swapped, so that what
    # monkeypatch restores afterwards is the real function and not the stand-in
    users = [importlib.import_module(name) for name in (
        "pyflow.numerics.core_ops_coherent", "pyflow.numerics.fluid_ops", "pyflow.numerics.operators.advection",
        "pyflow.linear_solvers.interface")]
    put(device, "DeviceSlab", CpuSlab)
    put(engine, "DeviceSlab", CpuSlab)
    put(device, "slab_for_shape", slab_for_shape)
    for mod in users:
        if hasattr(mod, "slab_for_shape"):
            put(mod, "slab_for_shape", slab_for_shape)
    return CpuSlab
