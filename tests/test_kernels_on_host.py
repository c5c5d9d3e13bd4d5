"""The element-wise CUDA kernels EXECUTED ON THE HOST, bit for bit against the reference's golden vectors.

``tests/hostemu/emu_kernels.cpp`` includes the real kernel sources (``pyflow-solver_b200/csrc/pf_kernels.cuh``)
under a minimal stand-in for ``<cuda_runtime.h>`` and runs each kernel body once per emulated thread, in the
launch geometry of ``pf_api.cu``, on buffers in the device field layout.  Built with ``g++ -ffp-contract=off``
this evaluates the same expressions ``nvcc -fmad=false`` compiles, so every result must equal the golden vectors
written by the unmodified reference -- which lets the arithmetic of these kernels be checked (and changed) on a
machine without a GPU.  Halo rows and padding columns are filled with NaN: a kernel that read outside the domain
would poison its output.

Scope: kernels whose threads do not communicate (operators, legacy central operators, predictor, correction, lid
row, centre line).  Kernels with grid reductions and the cooperative CG kernels are covered by ``-m gpu`` only.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pyflow_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OP_TAGS = ["8x8", "16x12", "5x9", "11x7", "33x17", "70x66"]
LEGACY_TAGS = ["8x8", "16x12", "5x9", "2x2", "3x3", "2x7", "9x2", "33x17", "70x66"]
UPWIND, QUICK, NONE = 0, 1, 2


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostemu") / "emu_kernels.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "tests", "hostemu"), "-I", os.path.join(ROOT, "pyflow-solver_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "hostemu", "emu_kernels.cpp"), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(out)
    dp, d, i = C.POINTER(C.c_double), C.c_double, C.c_int
    lib.emu_op_grad.argtypes = [i, i, dp, dp, dp, d, d]
    lib.emu_op_div.argtypes = [i, i, dp, dp, dp, d, d, i]
    lib.emu_op_lap.argtypes = [i, i, dp, dp, d, d]
    lib.emu_op_advect.argtypes = [i, i, i, dp, dp, dp, dp, d, d]
    lib.emu_op_grad_central.argtypes = [i, i, dp, dp, dp, d, d]
    lib.emu_op_lap_central.argtypes = [i, i, dp, dp, d, d]
    lib.emu_op_correct_central.argtypes = [i, i, i, dp, dp, dp, d, d, d]
    lib.emu_predict.argtypes = [i, i, i, dp, dp, dp, dp, d, d, d]
    lib.emu_correct.argtypes = [i, i, dp, dp, dp, d, d, d, d]
    lib.emu_lid.argtypes = [i, i, dp, d]
    lib.emu_centerline.argtypes = [i, i, dp, dp]
    return lib


class Dev:
    """Buffers in the device field layout: (ny + 2*HALO) rows x pitch doubles, NaN outside the interior."""

    def __init__(self, lib, ny, nx):
        self.lib, self.ny, self.nx = lib, ny, nx
        self.pitch, self.halo = lib.emu_pitch(nx), lib.emu_halo()

    def put(self, a=None):
        buf = np.full((self.ny + 2 * self.halo, self.pitch), np.nan)
        buf[self.halo:self.halo + self.ny, :self.nx] = 0.0 if a is None else a
        return buf

    def get(self, buf):
        assert np.isnan(buf[:self.halo]).all() and np.isnan(buf[self.halo + self.ny:]).all()      # halo rows untouched
        assert np.isnan(buf[:, self.nx:]).all()                                                  # padding never stored to
        return buf[self.halo:self.halo + self.ny, :self.nx].copy()


def ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("tag", OP_TAGS)
def test_operator_kernels_bit_exact_vs_golden(emu, golden, tag):
    g = {k.split("/", 1)[1]: v for k, v in golden["operators"].items() if k.startswith(tag + "/")}
    dx, dy = (float(t) for t in g["dxdy"])
    ny, nx = g["p"].shape
    dev = Dev(emu, ny, nx)
    p, u, v, f = (dev.put(g[k]) for k in ("p", "u", "v", "f"))
    gx, gy = dev.put(), dev.put()
    emu.emu_op_grad(nx, ny, ptr(p), ptr(gx), ptr(gy), dx, dy)
    assert same(dev.get(gx), g["gx"]) and same(dev.get(gy), g["gy"])
    out = dev.put()
    emu.emu_op_div(nx, ny, ptr(u), ptr(v), ptr(out), dx, dy, 0)
    assert same(dev.get(out), g["div"])
    emu.emu_op_div(nx, ny, ptr(u), ptr(v), ptr(out), dx, dy, 1)
    assert same(dev.get(out), g["divc"])
    emu.emu_op_div(nx, ny, ptr(u), ptr(v), ptr(out), 1.0, 1.0, 1)
    assert same(dev.get(out), g["divc1"])
    emu.emu_op_lap(nx, ny, ptr(p), ptr(out), dx, dy)
    assert same(dev.get(out), g["lap"])
    emu.emu_op_advect(UPWIND, nx, ny, ptr(u), ptr(v), ptr(f), ptr(out), dx, dy)
    assert same(dev.get(out), g["upwind"])
    emu.emu_op_advect(QUICK, nx, ny, ptr(u), ptr(v), ptr(f), ptr(out), dx, dy)
    assert same(dev.get(out), g["quick"])


@pytest.mark.parametrize("tag", LEGACY_TAGS)
def test_legacy_central_kernels_bit_exact_vs_golden(emu, golden, tag):
    g = {k.split("/", 1)[1]: v for k, v in golden["legacy_ops"].items() if k.startswith(tag + "/")}
    dx, dy, dt = (float(t) for t in g["dxdydt"])
    ny, nx = g["p"].shape
    dev = Dev(emu, ny, nx)
    p, u, v = (dev.put(g[k]) for k in ("p", "u", "v"))
    gx, gy, out = dev.put(), dev.put(), dev.put()
    emu.emu_op_grad_central(nx, ny, ptr(p), ptr(gx), ptr(gy), dx, dy)
    assert same(dev.get(gx), g["gx"]) and same(dev.get(gy), g["gy"])
    emu.emu_op_lap_central(nx, ny, ptr(p), ptr(out), dx, dy)
    assert same(dev.get(out), g["lap"])
    emu.emu_op_correct_central(0, nx, ny, ptr(p), ptr(u), ptr(out), dt, dx, dy)
    assert same(dev.get(out), g["ucorr"])
    emu.emu_op_correct_central(1, nx, ny, ptr(p), ptr(v), ptr(out), dt, dx, dy)
    assert same(dev.get(out), g["vcorr"])


@pytest.mark.parametrize("shape,unit", [((9, 13), True), ((40, 300), False), ((2, 2), True), ((3, 2), False),
                                        ((64, 513), True)])
def test_predictor_correction_lid_centerline_bit_exact_vs_oracle(emu, shape, unit):
    ny, nx = shape
    rng = np.random.default_rng(ny * 977 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    dt = 0.0123
    u, v, x = (rng.standard_normal(shape) for _ in range(3))
    u[rng.random(shape) < 0.1] = 0.0                       # exercises QUICK's >= against upwind's >
    dev = Dev(emu, ny, nx)
    for scheme, adv in ((QUICK, orc.advect_quick), (UPWIND, orc.advect_upwind), (NONE, None)):
        du, dv_, un, vn = dev.put(u), dev.put(v), dev.put(), dev.put()
        emu.emu_predict(scheme, nx, ny, ptr(du), ptr(dv_), ptr(un), ptr(vn), dt, dx, dy)
        eu = u if adv is None else u - dt * adv(u, v, u, dx, dy)
        ev = v if adv is None else v - dt * adv(u, v, v, dx, dy)
        assert same(dev.get(un), eu) and same(dev.get(vn), ev), scheme
    # correction with and without the pin shift p = x - x[0]
    for x0 in (0.0, float(x[0, 0])):
        dx_, du, dv_ = dev.put(x), dev.put(u), dev.put(v)
        emu.emu_correct(nx, ny, ptr(dx_), ptr(du), ptr(dv_), dt, dx, dy, x0)
        pe = x - x0
        gx, gy = orc.gradient_coherent(pe, dx, dy)
        assert same(dev.get(du), u - dt * gx) and same(dev.get(dv_), v - dt * gy)
        assert same(dev.get(dx_), pe)
    du = dev.put(u)
    emu.emu_lid(nx, ny, ptr(du), 1.5)
    lidded = u.copy()
    lidded[-1, :] = 1.5                                    # last interior row, every interior column (solver.py:104-105)
    assert same(dev.get(du), lidded)
    center = np.full(ny, np.nan)
    emu.emu_centerline(nx, ny, ptr(dev.put(u)), ptr(center))
    assert same(center, u[:, nx // 2])


def test_nan_and_inf_propagate_like_numpy(emu):
    rng = np.random.default_rng(11)
    p = rng.standard_normal((12, 20))
    p[3, 4], p[7, 0] = np.nan, np.inf
    dev = Dev(emu, 12, 20)
    out = dev.put()
    with np.errstate(invalid="ignore"):
        emu.emu_op_lap(20, 12, ptr(dev.put(p)), ptr(out), 1.0, 1.0)
        assert same(dev.get(out), orc.laplacian_coherent(p, 1.0, 1.0))
        emu.emu_op_lap_central(20, 12, ptr(dev.put(p)), ptr(out), 0.5, 0.25)
        assert same(dev.get(out), orc.laplacian_central(p, 0.5, 0.25))
