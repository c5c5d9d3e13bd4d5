"""ctypes face of the C / OpenMP oracle (TEST INFRASTRUCTURE ONLY).

``oracle/pyflow_oracle.c`` restates the same algorithm as ``pyflow_oracle.py``
with plain C loops so that the CPU baseline can use every host core.  The host
side of the step (adaptive dt, divergence-history safeguard, diagnostics
dictionary; solver.py:24-69,107-193 in the reference) is shared with the numpy
oracle; only the array work goes through the library.  Parity status: PINNED
through the numpy oracle and the golden vectors (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

import pyflow_oracle as _np_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libpyflow_oracle.so")
_dp = C.POINTER(C.c_double)


class Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("lid_velocity", C.c_double),
                ("lin_tol", C.c_double), ("inv_diag_interior", C.c_double),
                ("lid_set", C.c_int32), ("scheme", C.c_int32), ("jacobi", C.c_int32), ("lin_maxiter", C.c_int32)]


class Out(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("div_before_norm", "rhs_mean", "bnorm", "rnorm", "div_after_norm",
                                          "ru_sumsq", "rv_sumsq", "cont_sumsq")] + \
               [(k, C.c_int32) for k in ("info", "true_iters", "nonfinite", "reserved")]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pyflow_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.pfo_threads.restype = C.c_int
        lib.pfo_set_threads.argtypes = [C.c_int]
        lib.pfo_set_threads.restype = None
        lib.pfo_divc_sumsq.restype = C.c_double
        lib.pfo_divc_sumsq.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp]
        lib.pfo_absmax.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        lib.pfo_step.restype = C.c_int
        lib.pfo_step.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.POINTER(Params), C.POINTER(Out), _dp]
        lib.pfo_advect.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        for name in ("pfo_divergence_coherent", "pfo_divergence_central"):
            getattr(lib, name).argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp]
        lib.pfo_laplacian_coherent.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_double, _dp]
        lib.pfo_gradient_coherent.argtypes = [C.c_int, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp]
        lib.pfo_pcg.restype = C.c_int
        lib.pfo_pcg.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                C.c_double, _dp, _dp, _dp, _dp, C.POINTER(C.c_int), _dp]
        _lib = lib
    return _lib


def threads() -> int:
    return int(load().pfo_threads())


def set_threads(n: int) -> None:
    """OpenMP thread count of the C oracle (results are independent of it: per-row partials, serial fold)."""
    load().pfo_set_threads(int(n))


def _p(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ------------------------------------------------------------------ operator-level entry points (for the pin tests)
def gradient_coherent(p, dx, dy):
    p = _c(p); ny, nx = p.shape
    gx = np.empty_like(p); gy = np.empty_like(p)
    load().pfo_gradient_coherent(nx, ny, _p(p), dx, dy, _p(gx), _p(gy))
    return gx, gy


def divergence_coherent(u, v, dx, dy):
    u = _c(u); v = _c(v); ny, nx = u.shape
    out = np.empty_like(u)
    load().pfo_divergence_coherent(nx, ny, _p(u), _p(v), dx, dy, _p(out))
    return out


def laplacian_coherent(p, dx, dy):
    p = _c(p); ny, nx = p.shape
    out = np.empty_like(p)
    load().pfo_laplacian_coherent(nx, ny, _p(p), dx, dy, _p(out))
    return out


def divergence_central(u, v, dx, dy):
    u = _c(u); v = _c(v); ny, nx = u.shape
    out = np.empty_like(u)
    load().pfo_divergence_central(nx, ny, _p(u), _p(v), dx, dy, _p(out))
    return out


def _advect(scheme, u, v, f, dx, dy):
    u = _c(u); v = _c(v); f = _c(f); ny, nx = f.shape
    out = np.empty_like(f)
    load().pfo_advect(nx, ny, scheme, _p(u), _p(v), _p(f), dx, dy, _p(out))
    return out


def advect_upwind(u, v, f, dx, dy):
    return _advect(1, u, v, f, dx, dy)


def advect_quick(u, v, f, dx, dy):
    return _advect(2, u, v, f, dx, dy)


def pcg_pinned(b, nx, ny, dx, dy, tol, maxiter, jacobi=True):
    """CG on the pinned Neumann operator; returns (x, info, true_iterations)."""
    b = _c(b).reshape(-1)
    n = nx * ny
    x = np.empty(n); work = np.empty(3 * n + 2 * ny)
    it = C.c_int(0); bn = C.c_double(0.0)
    inner = float(_np_oracle.jacobi_inverse_diagonal(3, 3, dx, dy)[4])
    info = load().pfo_pcg(nx, ny, dx, dy, _p(b), _p(x), tol, maxiter, int(jacobi), inner,
                          _p(work[:n]), _p(work[n:2 * n]), _p(work[2 * n:3 * n]), _p(work[3 * n:]),
                          C.byref(it), C.byref(bn))
    return x, int(info), int(it.value)


# ------------------------------------------------------------------ the step (solver.py:15-193)
class Workspace:
    """Contiguous interior copies of u, v, p plus the scratch fields of one grid."""

    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny
        n = nx * ny
        self.u = np.zeros((ny, nx)); self.v = np.zeros((ny, nx)); self.p = np.zeros((ny, nx))
        self.work = np.empty(7 * n + 2 * ny)
        self.rows = np.empty(2 * ny)


def step(cfg, state, iteration, ws: Workspace | None = None):
    """Same contract as ``pyflow_oracle.step``; the array work runs in C / OpenMP."""
    lib = load()
    ny, nx = state.ny, state.nx
    ws = ws or state.meta.setdefault("_c_workspace", Workspace(nx, ny))
    for k, a in (("u", ws.u), ("v", ws.v), ("p", ws.p)):
        a[:, :] = state.fields[k][1:-1, 1:-1]
    lx = getattr(cfg, "lx", float(nx - 1)); ly = getattr(cfg, "ly", float(ny - 1))
    dx = lx / max(nx - 1, 1); dy = ly / max(ny - 1, 1)
    mx = (C.c_double * 2)()
    lib.pfo_absmax(nx, ny, _p(ws.u), _p(ws.v), mx)
    umax = float(mx[0]) + 1e-12; vmax = float(mx[1]) + 1e-12
    cfl = getattr(cfg, "cfl_target", 0.5); growth = getattr(cfg, "cfl_growth", 1.05)
    nu = 1.0 / float(getattr(cfg, "Re", 100.0))
    dt_prev = float(state.meta.get("dt_prev", 1e-2))
    dt_cand = min(min(cfl * dx / umax, cfl * dy / vmax), 0.25 * min(dx * dx, dy * dy) / max(nu, 1e-12))
    dt = dt_prev * growth if dt_cand > dt_prev * growth else dt_cand
    hist = state.meta.get("divergence_history", [])
    hist.append(math.sqrt(lib.pfo_divc_sumsq(nx, ny, _p(ws.u), _p(ws.v), 1.0, 1.0, _p(ws.rows))))
    hist = hist[-5:]
    state.meta["divergence_history"] = hist
    if len(hist) >= 3:
        a, b, c = hist[-3], hist[-2], hist[-1]
        if a > 0 and b > a * 1.01 and c > b * 1.01:
            dt *= 0.5
    state.meta.update(dt_prev=dt, dx=dx, dy=dy, nu=nu)
    cfl_now = max(umax * dt / dx, vmax * dt / dy)
    scheme = 0 if getattr(cfg, "disable_advection", False) else \
        (2 if getattr(cfg, "advection_scheme", "upwind") == "quick" else 1)
    lid = getattr(cfg, "lid_velocity", None)
    tol = getattr(cfg, "lin_tol", 1e-10)
    prm = Params(dt, dx, dy, 0.0 if lid is None else float(lid), tol,
                 float(_np_oracle.jacobi_inverse_diagonal(3, 3, dx, dy)[4]),
                 int(lid is not None), scheme, int(getattr(cfg, "enable_jacobi_pc", True)),
                 int(getattr(cfg, "lin_maxiter", 400)))
    out = Out()
    rc = lib.pfo_step(nx, ny, _p(ws.u), _p(ws.v), _p(ws.p), C.byref(prm), C.byref(out), _p(ws.work))
    if rc != 0:
        raise RuntimeError(f"pfo_step failed ({rc})")
    for k, a in (("u", ws.u), ("v", ws.v), ("p", ws.p)):
        state.fields[k][1:-1, 1:-1] = a
    n = nx * ny
    Ru = math.sqrt(out.ru_sumsq) / max(1, n); Rv = math.sqrt(out.rv_sumsq) / max(1, n)
    continuity = math.sqrt(out.cont_sumsq) / max(1.0, n)
    bnorm = out.bnorm or 1.0
    iterations = out.info if out.info > 0 else 0
    if iterations == 0 and out.rnorm > tol:
        iterations = 1
    state.meta["prev_div_norm"] = out.div_after_norm
    residuals = {"Ru": Ru, "Rv": Rv, "Rp": 0.0, "continuity": continuity}
    diagnostics = {
        "iteration": iteration, "dt": dt, "CFL": cfl_now, "nu": nu,
        "Rp_iterations": iterations, "Rp_residual": out.rnorm,
        "Rp_converged": (out.info == 0) and (out.rnorm <= tol * bnorm), "Rp_method": "cg",
        "divergence_norm": out.div_after_norm, "nan_detected": bool(out.nonfinite),
        "true_iterations": int(out.true_iters), "max_residual": max(residuals.values()),
        "residual_drop_Ru_orders": max(0.0, (0.0 - math.log10(Ru)) if Ru > 0 else 0.0),
    }
    return state, residuals, diagnostics


def run(cfg, steps, state=None):
    if state is None:
        state = _np_oracle.allocate_state(cfg.nx, cfg.ny)
    ws = Workspace(cfg.nx, cfg.ny)
    out = []
    for it in range(steps):
        _, res, diag = step(cfg, state, it, ws)
        diag = dict(diag); diag.update(res)
        out.append(diag)
    return state, out
