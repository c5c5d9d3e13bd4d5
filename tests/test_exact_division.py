"""``div_by_invariant`` (csrc/pf_common.cuh) -- the reciprocal + two FMA corrections every general-spacing kernel and
the rhs pass use in place of a double division -- returns the IEEE quotient BIT FOR BIT.

The product source is compiled for the host (``tests/hostemu/divcheck.cpp``) and compared with ``a / h`` on 2e8 operand
pairs per run: random magnitudes over the whole double range, quotients constructed next to a midpoint between two
doubles, divisors with all-ones / power-of-two / few-bit significands, zeros of both signs, infinities, NaN,
subnormals, negative and out-of-range divisors (those take the real division).  2e10 pairs were run once when the
routine was written (0 mismatches; a multiply by the reciprocal alone mismatches on 25 %).
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_division_by_an_invariant_divisor_is_bit_exact(tmp_path):
    exe = str(tmp_path / "divcheck")
    flags = ["-mfma"] if " fma " in open("/proc/cpuinfo").read() else []      # hardware FMA: 10x faster, same results
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", *flags, "-DPF_HOST_EMULATION", "-pthread",
           "-I", os.path.join(ROOT, "tests", "hostemu"), "-I", os.path.join(ROOT, "pyflow-solver_b200", "csrc"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "hostemu", "divcheck.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    pairs = 200_000_000 if flags else 20_000_000
    r = subprocess.run([exe, str(pairs), "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    words = r.stdout.split()
    assert words[-1] == "0" and int(words[3]) > pairs // 2, r.stdout          # 0 mismatches, most pairs on the fast path
