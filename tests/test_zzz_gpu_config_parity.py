"""Parity of the CUDA path against the C oracle ON THE GRIDS THE BENCH NUMBERS ARE QUOTED ON
(BASELINE configs 3, 4 and 5; reference solver.py:15-193 through oracle/pyflow_oracle.c).

The small-grid suites compare every kernel with the oracle where there are fewer tiles than CTAs; here the
TMA-fed kernel runs in the regime bench.py times -- 1024 tiles drawn by 148 CTAs, the ring running on across
tile boundaries for 600 sweeps per step, 16 column chunks and taller tiles at 16384 columns.
Bars: dt bit-identical, true CG iterations equal within 1, u / v / p relative L2 <= 1e-9.
"""
import os

import pytest

import parity

pytestmark = pytest.mark.gpu


def _host_gib():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def _check(name, cfg, steps, options=None):
    parity.use_all_host_threads()
    r = parity.check_case(name, cfg, steps, options=options)
    print(f"PARITY {name} options={options} rel_l2={r['rel_l2']} dt_equal={r['dt_equal']} iters={r['iters']} "
          f"oracle_iters={r['oracle_iters']} oracle_s={r['oracle_s']} {'OK' if r['ok'] else 'FAIL'}")
    assert r["dt_equal"], r
    assert r["max_iter_diff"] <= r["iter_bar"], r
    assert r["max_rel_l2"] <= parity.REL_BAR, r
    return r


@pytest.mark.parametrize("options", [None, "cg_kernel=1", "cg_kernel=2", "step_fused=0"])
def test_config3_4096_three_steps_vs_c_oracle(options):
    """BASELINE config 3 exactly as bench.py runs it (4096^2, Re=1000, QUICK, 300 iterations per step), 3 steps
    from rest: the dynamically scheduled TMA kernel (default), the register-window kernel (cg_kernel=1), the
    static TMA kernel (cg_kernel=2), and the default CG kernel between the seven separate per-step kernels instead of
    the two fused passes (step_fused=0)."""
    r = _check("config3 4096x4096", parity.make_cfg(4096, 4096, 1000.0), 3, options)
    assert r["iters"] == [0, 300, 300]


def test_config4_8192_vs_c_oracle():
    """BASELINE config 4 on one GPU (8192^2, Re=5000): 8 column chunks, 4096 tiles."""
    if _host_gib() < 24:
        pytest.skip("needs ~16 GiB of host memory for the oracle")
    _check("config4 8192x8192", parity.make_cfg(8192, 8192, 5000.0), 3)


def test_config5_16384_wide_slab_vs_c_oracle():
    """BASELINE config 5 geometry: 16384 columns (16 chunks), a 1024-row slab, 300 iterations."""
    _check("config5 16384x1024", parity.make_cfg(16384, 1024, 1000.0), 3)


def test_config5_16384_full_slab_vs_c_oracle():
    """The full 16384 x 16384 per-GPU slab of config 5 (2 GiB per vector, tiles grow to 32 rows), iteration cap
    lowered to 30 so that the oracle finishes in seconds."""
    if _host_gib() < 80:
        pytest.skip("needs ~60 GiB of host memory for the oracle and both states")
    _check("config5 16384x16384 (lin_maxiter 30)", parity.make_cfg(16384, 16384, 1000.0, lin_maxiter=30), 2)


@pytest.mark.parametrize("nx,ny", [(4000, 4000), (5000, 3000), (4097, 1500)])
def test_general_widths_take_the_tma_kernel_and_match_the_c_oracle(nx, ny):
    """Widths that are not a multiple of 1024 (clipped last column chunk, strips cut by the east wall, an odd width
    whose last chunk is a single column): the TMA-fed kernel, whole steps against the C oracle."""
    _check(f"{nx}x{ny}", parity.make_cfg(nx, ny, 1000.0), 3)
