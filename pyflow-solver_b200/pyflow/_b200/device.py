"""Device side of the drop-in: one ``DeviceSlab`` = one GPU context of
``libpyflow_b200`` owning the row slab of a cavity grid.

PyTorch is used for exactly two things here, as the north star asks: the device
workspace (``torch.empty`` on the GPU) and the CUDA stream handle.  All compute
is in the hand-written kernels behind the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import lib as _lib


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise _lib.NativeLibraryError(
            "no CUDA device visible: the PyFlow B200 path has no CPU fallback (torch.cuda.is_available() is False)")
    return torch


def _ptr(a: np.ndarray | None):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _interior_ptr(ghosted: np.ndarray | None):
    """Address of the first interior cell of a (ny+2, nx+2) C-contiguous array."""
    if ghosted is None:
        return None
    stride = ghosted.shape[1]
    return C.c_void_p(ghosted.ctypes.data + (stride + 1) * 8)


class DeviceSlab:
    """GPU context for an ``nx x ny_local`` slab of a ``nx x ny_global`` grid."""

    def __init__(self, nx: int, ny_local: int, ny_global: int | None = None, row_offset: int = 0,
                 device: int | None = None, allocator=None):
        torch = _require_cuda()
        self._lib = _lib.load()
        self.nx, self.ny_local = int(nx), int(ny_local)
        self.ny_global = int(ny_global if ny_global is not None else ny_local)
        self.row_offset = int(row_offset)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._torch = torch
        nbytes = int(self._lib.pf_workspace_bytes(self.nx, self.ny_local))
        if nbytes == 0:
            raise ValueError(f"unsupported grid {nx}x{ny_local} (need nx >= 2, ny >= 2)")
        if allocator is not None:           # e.g. torch symmetric memory for peer-mapped slabs
            self.workspace = allocator(nbytes, self.device)
        else:
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.device}")
        self.workspace_bytes = nbytes
        base = self.workspace.data_ptr()
        if base % 256:
            raise RuntimeError("torch returned a workspace that is not 256-byte aligned")
        ctx = C.c_void_p()
        rc = self._lib.pf_create(C.byref(ctx), self.nx, self.ny_local, self.ny_global, self.row_offset,
                                 self.device, C.c_void_p(base), nbytes)
        _lib.check(rc, None, "pf_create")
        self.ctx = ctx
        self._finalizer = weakref.finalize(self, DeviceSlab._destroy, self._lib, ctx)
        self._center = np.empty(self.ny_local, dtype=np.float64)
        import os
        for item in filter(None, os.environ.get("PYFLOW_B200_OPTIONS", "").split(",")):   # e.g. "cg_multi_single=1"
            k, _, v = item.partition("=")
            self.set_option(k.strip(), int(v or 1))
        self.nranks = 1
        self.rank = 0
        self.peer_mode = False

    @staticmethod
    def _destroy(lib, ctx):
        try:
            lib.pf_destroy(ctx)
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def close(self):
        self._finalizer()

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def info(self) -> dict:
        sm, grid, thr = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._lib.pf_device_info(self.ctx, C.byref(sm), C.byref(grid), C.byref(thr)), self.ctx)
        return {"sm_count": sm.value, "cg_grid_blocks": grid.value, "cg_block_threads": thr.value}

    def set_option(self, key: str, value: int):
        _lib.check(self._lib.pf_set_option(self.ctx, key.encode(), int(value)), self.ctx, "pf_set_option")

    def launch_count(self) -> int:
        n = C.c_int64()
        _lib.check(self._lib.pf_launch_count(self.ctx, C.byref(n)), self.ctx)
        return int(n.value)

    def init_comm(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _lib.check(self._lib.pf_comm_init(self.ctx, buf, rank, nranks), self.ctx, "pf_comm_init")
        self.rank, self.nranks = rank, nranks

    def set_peers(self, peer_ptrs):
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _lib.check(self._lib.pf_comm_set_peers(self.ctx, arr, len(peer_ptrs)), self.ctx, "pf_comm_set_peers")
        self.peer_mode = True

    @staticmethod
    def make_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().pf_comm_unique_id(buf), None, "pf_comm_unique_id")
        return buf.raw

    # ------------------------------------------------------------------ fields
    @staticmethod
    def _check_ghosted(a, nx, ny):
        if a is None:
            return
        if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != (ny + 2, nx + 2):
            raise ValueError(f"field must be C-contiguous float64 of shape {(ny + 2, nx + 2)}, got {a.dtype} {a.shape}")

    def upload(self, u=None, v=None, p=None):
        """Ghosted (ny_local+2, nx+2) host arrays -> device interior."""
        for a in (u, v, p):
            self._check_ghosted(a, self.nx, self.ny_local)
        rc = self._lib.pf_upload(self.ctx, _interior_ptr(u), _interior_ptr(v), _interior_ptr(p),
                                 self.nx + 2, self._stream())
        _lib.check(rc, self.ctx, "pf_upload")

    def download(self, u=None, v=None, p=None):
        for a in (u, v, p):
            self._check_ghosted(a, self.nx, self.ny_local)
        rc = self._lib.pf_download(self.ctx, _interior_ptr(u), _interior_ptr(v), _interior_ptr(p),
                                   self.nx + 2, self._stream())
        _lib.check(rc, self.ctx, "pf_download")

    # ------------------------------------------------------------------ step
    def prestep(self):
        out = (C.c_double * 3)()
        _lib.check(self._lib.pf_prestep_reduce(self.ctx, out, self._stream()), self.ctx, "pf_prestep_reduce")
        return float(out[0]), float(out[1]), float(out[2])

    def step(self, params: _lib.StepParams) -> _lib.StepOut:
        out = _lib.StepOut()
        _lib.check(self._lib.pf_step(self.ctx, C.byref(params), C.byref(out), self._stream()), self.ctx, "pf_step")
        return out

    def divergence_stats(self, dx: float, dy: float) -> tuple[float, float]:
        """(mean, mean-free 2-norm) of the central divergence of the device u, v."""
        out = (C.c_double * 2)()
        _lib.check(self._lib.pf_divergence_stats(self.ctx, float(dx), float(dy), out, self._stream()), self.ctx,
                   "pf_divergence_stats")
        return float(out[0]), float(out[1])

    def centerline(self) -> np.ndarray:
        _lib.check(self._lib.pf_centerline(self.ctx, _ptr(self._center), self._stream()), self.ctx, "pf_centerline")
        return self._center.copy()

    # ------------------------------------------------------------------ operator seams (dense (ny, nx) arrays)
    def _dense(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (self.ny_local, self.nx):
            raise ValueError(f"expected shape {(self.ny_local, self.nx)}, got {a.shape}")
        return a

    def gradient_coherent(self, p, dx, dy):
        p = self._dense(p)
        gx, gy = np.empty_like(p), np.empty_like(p)
        _lib.check(self._lib.pf_op_gradient_coherent(self.ctx, _ptr(p), _ptr(gx), _ptr(gy), float(dx), float(dy),
                                                     self._stream()), self.ctx, "pf_op_gradient_coherent")
        return gx, gy

    def divergence_coherent(self, u, v, dx, dy):
        u, v = self._dense(u), self._dense(v)
        out = np.empty_like(u)
        _lib.check(self._lib.pf_op_divergence_coherent(self.ctx, _ptr(u), _ptr(v), _ptr(out), float(dx), float(dy),
                                                       self._stream()), self.ctx, "pf_op_divergence_coherent")
        return out

    def divergence_central(self, u, v, dx, dy):
        u, v = self._dense(u), self._dense(v)
        out = np.empty_like(u)
        _lib.check(self._lib.pf_op_divergence_central(self.ctx, _ptr(u), _ptr(v), _ptr(out), float(dx), float(dy),
                                                      self._stream()), self.ctx, "pf_op_divergence_central")
        return out

    def laplacian_coherent(self, p, dx, dy):
        p = self._dense(p)
        out = np.empty_like(p)
        _lib.check(self._lib.pf_op_laplacian_coherent(self.ctx, _ptr(p), _ptr(out), float(dx), float(dy),
                                                      self._stream()), self.ctx, "pf_op_laplacian_coherent")
        return out

    def gradient_central(self, p, dx, dy):
        """fluid_ops.gradient (reference numerics/fluid_ops.py:43-67)."""
        p = self._dense(p)
        gx, gy = np.empty_like(p), np.empty_like(p)
        _lib.check(self._lib.pf_op_gradient_central(self.ctx, _ptr(p), _ptr(gx), _ptr(gy), float(dx), float(dy),
                                                    self._stream()), self.ctx, "pf_op_gradient_central")
        return gx, gy

    def laplacian_central(self, p, dx, dy):
        """fluid_ops.laplacian (reference numerics/fluid_ops.py:69-91)."""
        p = self._dense(p)
        out = np.empty_like(p)
        _lib.check(self._lib.pf_op_laplacian_central(self.ctx, _ptr(p), _ptr(out), float(dx), float(dy),
                                                     self._stream()), self.ctx, "pf_op_laplacian_central")
        return out

    def pressure_correction_central(self, u, v, p, dt, dx, dy):
        """(u - dt*dp/dx, v - dt*dp/dy) with the central gradient (reference numerics/fluid_ops.py:108-125)."""
        u, v, p = self._dense(u), self._dense(v), self._dense(p)
        uo, vo = np.empty_like(u), np.empty_like(v)
        _lib.check(self._lib.pf_op_pressure_correction_central(self.ctx, _ptr(u), _ptr(v), _ptr(p), float(dt),
                                                               float(dx), float(dy), _ptr(uo), _ptr(vo),
                                                               self._stream()),
                   self.ctx, "pf_op_pressure_correction_central")
        return uo, vo

    def advect(self, scheme: str, u, v, f, dx, dy):
        code = {"upwind": _lib.PF_ADV_UPWIND, "quick": _lib.PF_ADV_QUICK}[scheme]
        u, v, f = self._dense(u), self._dense(v), self._dense(f)
        out = np.empty_like(f)
        _lib.check(self._lib.pf_op_advect(self.ctx, code, _ptr(u), _ptr(v), _ptr(f), _ptr(out), float(dx), float(dy),
                                          self._stream()), self.ctx, "pf_op_advect")
        return out

    def pcg(self, b, dx, dy, tol, maxiter, jacobi=True, inv_diag_interior=None):
        """Solve (_NegLaplacianPinned) x = b with scipy's CG recurrence on the device.
        Returns (x (ny,nx), info, true_iterations, ||b - A x||)."""
        b = self._dense(b)
        x = np.empty_like(b)
        if inv_diag_interior is None:
            inv_diag_interior = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
        info, iters, rnorm = C.c_int(), C.c_int(), C.c_double()
        _lib.check(self._lib.pf_op_pcg(self.ctx, _ptr(b), _ptr(x), float(dx), float(dy), int(bool(jacobi)),
                                       float(inv_diag_interior), float(tol), int(maxiter), C.byref(info),
                                       C.byref(iters), C.byref(rnorm), self._stream()), self.ctx, "pf_op_pcg")
        return x, info.value, iters.value, rnorm.value

    def bicgstab(self, b, dx, dy, tol, maxiter, jacobi=False, inv_diag_interior=None):
        """Solve (_NegLaplacianPinned) x = b with scipy's BiCGSTAB recurrence on the device (k_bicgstab).
        Returns (x (ny,nx), info, completed_iterations, ||b - A x||); info is scipy's exit code."""
        b = self._dense(b)
        x = np.empty_like(b)
        if inv_diag_interior is None:
            inv_diag_interior = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
        info, iters, rnorm = C.c_int(), C.c_int(), C.c_double()
        _lib.check(self._lib.pf_op_bicgstab(self.ctx, _ptr(b), _ptr(x), float(dx), float(dy), int(bool(jacobi)),
                                            float(inv_diag_interior), float(tol), int(maxiter), C.byref(info),
                                            C.byref(iters), C.byref(rnorm), self._stream()), self.ctx, "pf_op_bicgstab")
        return x, info.value, iters.value, rnorm.value

    def krylov_diag(self, method: str, b, inv_diag, dx, dy, tol, maxiter):
        """CG ('cg') or BiCGSTAB (anything else, as in the reference) on the pinned operator with a caller-supplied
        diagonal preconditioner: ``inv_diag`` (ny, nx) are the entries M applies.  Returns like ``pcg``."""
        b = self._dense(b)
        d = self._dense(inv_diag)
        x = np.empty_like(b)
        info, iters, rnorm = C.c_int(), C.c_int(), C.c_double()
        _lib.check(self._lib.pf_op_krylov_diag(self.ctx, 0 if method == "cg" else 1, _ptr(b), _ptr(d), _ptr(x),
                                               float(dx), float(dy), float(tol), int(maxiter), C.byref(info),
                                               C.byref(iters), C.byref(rnorm), self._stream()),
                   self.ctx, "pf_op_krylov_diag")
        return x, info.value, iters.value, rnorm.value

    def krylov_stencil5(self, method: str, b, coef, tol, maxiter, jacobi=False, inv_diag_interior=1.0):
        """CG ('cg') or BiCGSTAB (anything else, as in the reference) on the ASSEMBLED ring-identity 5-point matrix
        (reference pressure_solver.py:15-40); coef = (south, west, centre, east, north) of an interior row.
        Returns like ``pcg``."""
        b = self._dense(b)
        x = np.empty_like(b)
        cf = (C.c_double * 5)(*[float(v) for v in coef])
        info, iters, rnorm = C.c_int(), C.c_int(), C.c_double()
        _lib.check(self._lib.pf_op_krylov_stencil5(self.ctx, 0 if method == "cg" else 1, _ptr(b), _ptr(x), cf,
                                                   int(bool(jacobi)), float(inv_diag_interior), float(tol), int(maxiter),
                                                   C.byref(info), C.byref(iters), C.byref(rnorm), self._stream()),
                   self.ctx, "pf_op_krylov_stencil5")
        return x, info.value, iters.value, rnorm.value

    def cg_generic(self, mode: str, b, dx, dy, tol, maxiter, jacobi=False, inv_diag_interior=1.0, want_x=True):
        """Reference-order CG with the plain (q stored) kernel.  mode 'pinned' | 'dirichlet'.
        The solution stays in the context's p field.  Returns (x|None, info, iters, ||b-Ax||, ||b||)."""
        code = {"pinned": 0, "dirichlet": 1}[mode]
        b = self._dense(b)
        x = np.empty_like(b) if want_x else None
        info, iters, rnorm, bnorm = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        _lib.check(self._lib.pf_op_cg_generic(self.ctx, code, _ptr(b), _ptr(x), float(dx), float(dy), int(bool(jacobi)),
                                              float(inv_diag_interior), float(tol), int(maxiter), C.byref(info),
                                              C.byref(iters), C.byref(rnorm), C.byref(bnorm), self._stream()),
                   self.ctx, "pf_op_cg_generic")
        return x, info.value, iters.value, rnorm.value, bnorm.value


_SLAB_CACHE: dict[tuple[int, int], DeviceSlab] = {}


def slab_for_shape(ny: int, nx: int) -> DeviceSlab:
    """Small cache of single-GPU contexts for the array-level operator seams."""
    key = (int(ny), int(nx))
    slab = _SLAB_CACHE.get(key)
    if slab is None:
        if len(_SLAB_CACHE) >= 8:
            _SLAB_CACHE.pop(next(iter(_SLAB_CACHE))).close()
        slab = _SLAB_CACHE[key] = DeviceSlab(nx, ny)
    return slab
