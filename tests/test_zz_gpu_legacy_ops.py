"""GPU parity of the legacy central operators of ``fluid_ops`` (gradient, laplacian,
apply_pressure_correction: reference numerics/fluid_ops.py:43-125) -- BIT-EXACT against the
golden vectors of the unmodified reference and against the oracle on larger / odd shapes.
These kernels are off the time-step path; the reference's own projection tests import them.
"""
import numpy as np
import pytest

import pyflow_oracle as orc

pytestmark = pytest.mark.gpu

LEGACY_TAGS = ["8x8", "16x12", "5x9", "2x2", "3x3", "2x7", "9x2", "33x17", "70x66"]


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("tag", LEGACY_TAGS)
def test_legacy_central_operators_bit_exact_vs_golden(golden, tag):
    from pyflow.core.ghost_fields import allocate_state, interior_view
    from pyflow.numerics import fluid_ops
    g = {k.split("/", 1)[1]: v for k, v in golden["legacy_ops"].items() if k.startswith(tag + "/")}
    dx, dy, dt = (float(t) for t in g["dxdydt"])
    gx, gy = fluid_ops.gradient(g["p"], dx, dy)
    assert bits_equal(gx, g["gx"]) and bits_equal(gy, g["gy"])
    assert bits_equal(fluid_ops.laplacian(g["p"], dx, dy), g["lap"])
    ny, nx = g["p"].shape
    st = allocate_state(nx, ny)
    interior_view(st.fields["u"])[:] = g["u"]
    interior_view(st.fields["v"])[:] = g["v"]
    fluid_ops.apply_pressure_correction(st, g["p"], dt, dx, dy)
    assert bits_equal(interior_view(st.fields["u"]), g["ucorr"]) and bits_equal(interior_view(st.fields["v"]), g["vcorr"])
    assert bits_equal(interior_view(st.fields["p"]), g["p"])
    lgx, lgy = fluid_ops.legacy_grad_p_state(st, dx, dy)
    assert bits_equal(lgx, g["gx"]) and bits_equal(lgy, g["gy"])


@pytest.mark.parametrize("shape,unit", [((257, 513), True), ((300, 1030), False), ((64, 3), True), ((3, 64), False)])
def test_legacy_central_operators_bit_exact_vs_oracle(shape, unit):
    from pyflow.numerics import fluid_ops
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx + 1)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    p = rng.standard_normal(shape)
    p[ny // 2, nx // 2] = np.nan
    with np.errstate(invalid="ignore"):
        gx, gy = fluid_ops.gradient(p, dx, dy)
        ogx, ogy = orc.gradient_central(p, dx, dy)
        assert bits_equal(gx, ogx) and bits_equal(gy, ogy)
        assert bits_equal(fluid_ops.laplacian(p, dx, dy), orc.laplacian_central(p, dx, dy))
