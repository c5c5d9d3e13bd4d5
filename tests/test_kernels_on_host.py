"""The CUDA kernel sources EXECUTED ON THE HOST: element-wise kernels bit for bit against the reference's golden
vectors, reduction kernels and the plain CG kernel against numpy / the oracle, whole projection steps chained from them.

``tests/hostemu/emu_kernels.cpp`` includes the real kernel sources (``pyflow-solver_b200/csrc/pf_kernels.cuh``)
under a minimal stand-in for ``<cuda_runtime.h>`` and runs each kernel body once per emulated thread, in the
launch geometry of ``pf_api.cu``, on buffers in the device field layout.  Built with ``g++ -ffp-contract=off``
this evaluates the same expressions ``nvcc -fmad=false`` compiles, so every result must equal the golden vectors
written by the unmodified reference -- which lets the arithmetic of these kernels be checked (and changed) on a
machine without a GPU.  Halo rows and padding columns are filled with NaN: a kernel that read outside the domain
would poison its output.

Two execution styles: kernels whose threads do not communicate (operators, legacy central operators, predictor,
correction, lid row, centre line) run one emulated thread after the other; kernels with grid reductions (pre-step,
rhs, residual, divergence norms, post-step pass) and the plain cooperative CG kernel (``k_cg_generic``, on a grid of
one block) run with ONE OS THREAD PER CUDA THREAD -- ``__syncthreads`` is a barrier, warp shuffles go through a
per-warp exchange buffer, atomics are atomics -- so the "last block folds the partials" machinery and the whole CG
recurrence execute as written.  Chained in the order of ``pf_step`` they make a complete projection step from the
CUDA sources on the host, compared with the oracle.  The register-window production CG kernel (``k_cg_persistent``)
runs the same way on one CTA, the split-phase chain of the NCCL fallback on several CTAs one after the other, and the
shared-memory-resident kernel (``k_cg_small``) with ALL its CTAs at once, exchanging their tagged packets for real.
The TMA-fed production kernel (``k_cg_tmad``, single-GPU instantiation) runs on one CTA of 16 compute warps + the
producer warp, its mbarriers and bulk copies replaced by host stand-ins in a generated copy.  Its MULTI-GPU instantiation runs too: several
ranks, one CTA each, all at once in one process, storing their slab-boundary rows into the neighbours' halo rows and
all-reducing through the mailboxes -- the code that runs over NVLink, with "peer memory" being the other ranks'
buffers.  Nothing here says anything about timing or about the hardware memory model.
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


# k_cg_small: a GENERATED copy of pf_cg_small.cuh gets per-block storage for its shared memory, so that several of its
# CTAs can run at once in one process.  Its exchanges need no stand-ins: the LL-word helpers of pf_common.cuh have
# host versions (PF_HOST_EMULATION) that store a packet word by word, highest address first, yielding in between --
# so the pollers of the emulation regularly see TORN packets, which the encoding must (and does) reject.  The product
# source is not touched; if one of the patterns below no longer occurs exactly once the generation fails loudly.


def generate_small_kernel_copy(dest_dir):
    src = open(os.path.join(ROOT, "pyflow-solver_b200", "csrc", "pf_cg_small.cuh")).read()
    for old, new in (
            ('#include "pf_cg.cuh"', '#include "pf_cg.cuh"\nnamespace pf { inline unsigned long long emu_small_gtime() '
                                     '{ return (unsigned long long)clock64(); } }\n#define gtime emu_small_gtime'),
            ("extern __shared__ __align__(16) double smem_d[];", "double* smem_d = emu::dyn_shared();"),
            ("__shared__ double sm[3 * 32];", "double* sm = emu::blk_shared(0);"),
            ("__shared__ double s_res[8];", "double* s_res = emu::blk_shared(1);")):
        assert src.count(old) == 1, f"pf_cg_small.cuh changed: {old!r} occurs {src.count(old)} times"
        src = src.replace(old, new)
    with open(os.path.join(dest_dir, "pf_cg_small_emu.cuh"), "w") as fh:
        fh.write(src)


# The TMA-fed kernels: the block of mbarrier / bulk-copy PTX helpers at the top of pf_cg_tma.cuh is replaced, in a
# GENERATED copy, by host stand-ins -- an mbarrier is its 64-bit word of "shared memory" holding {phase, pending
# arrivals, expected count, outstanding bytes}; arrive / expect_tx / complete_tx update it under a mutex and flip the
# phase when arrivals and bytes are both down to zero; try_wait(parity) is true once the phase of that parity has
# completed; a bulk copy is a memcpy followed by complete_tx.  Everything else (ring, tickets, tile partials) is the
# product source, untouched.
TMA_HOST_MBAR = '''// ---- mbarrier / bulk-copy PTX: HOST STAND-INS (generated copy, tests only) --------
struct EmuMbar { uint32_t phase : 1, pending : 15, count : 15; int32_t tx; };
static_assert(sizeof(EmuMbar) == 8, "an mbarrier is one 64-bit word of shared memory");
static pthread_mutex_t emu_mbar_mutex = PTHREAD_MUTEX_INITIALIZER;
inline void emu_mbar_update(uint64_t* bar, int arrive, int tx_delta) {
  pthread_mutex_lock(&emu_mbar_mutex);
  EmuMbar s;
  memcpy(&s, bar, 8);
  s.tx += tx_delta;
  s.pending -= (uint32_t)arrive;
  if (s.pending == 0 && s.tx == 0) { s.phase ^= 1u; s.pending = s.count; }      // the phase completes
  memcpy(bar, &s, 8);
  pthread_mutex_unlock(&emu_mbar_mutex);
}
inline uint32_t smem_u32(const void* p) { return (uint32_t)(size_t)p; }
inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuMbar s;
  s.phase = 0; s.pending = count; s.count = count; s.tx = 0;
  memcpy(bar, &s, 8);
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_update(bar, 1, (int)bytes); }
inline void mbar_arrive(uint64_t* bar) { emu_mbar_update(bar, 1, 0); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  pthread_mutex_lock(&emu_mbar_mutex);
  EmuMbar s;
  memcpy(&s, bar, 8);
  pthread_mutex_unlock(&emu_mbar_mutex);
  const bool done = s.phase != parity;              // the phase with this parity has completed
  if (!done) sched_yield();
  return done;
}
inline bool mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 40000000000LL) return false;
  }
  return true;
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  memcpy(dst_smem, src_gmem, bytes);                // the copy engine: data lands, then complete_tx
  emu_mbar_update(bar, 0, -(int)bytes);
}
inline void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t) {
  bulk_g2s(dst_smem, src_gmem, bytes, bar);
}
// address-based variants of the lean loops: on the host an "address" is the pointer itself
typedef uintptr_t smem_addr_t;
inline smem_addr_t smem_addr_pin(const void* p) { return (smem_addr_t)p; }
inline bool mbar_wait_at(smem_addr_t bar, uint32_t parity) { return mbar_wait((uint64_t*)bar, parity); }
inline void mbar_arrive_at(smem_addr_t bar) { mbar_arrive((uint64_t*)bar); }
inline double2 lds2_at(smem_addr_t p) { return *(const double2*)p; }
inline double lds1_at(smem_addr_t p) { return *(const double*)p; }

'''

TMA_HOST_PACKETS = '''
// ---- host stand-ins for the acquire / timer PTX of pf_cg.cuh (generated copy, tests only); the LL packet words have
// host versions of their own in pf_common.cuh ----
namespace pf {
inline unsigned long long emu_tma_ld_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline unsigned long long emu_tma_gtime() { return (unsigned long long)clock64(); }
}  // namespace pf
#define ld_acquire_sys emu_tma_ld_u64
#define ld_acquire_gpu emu_tma_ld_u64
#define gtime emu_tma_gtime
'''


def generate_tma_kernel_copy(dest_dir):
    src = open(os.path.join(ROOT, "pyflow-solver_b200", "csrc", "pf_cg_tma.cuh")).read()
    begin, end = "// ---- mbarrier / bulk-copy PTX", "struct TmaTile {"
    assert src.count(begin) == 1 and src.count(end) == 1, "pf_cg_tma.cuh changed: PTX helper block not found"
    src = src[:src.index(begin)] + TMA_HOST_MBAR + src[src.index(end):]
    for old, new, times in (
            ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', "__threadfence();", 2),
            ("extern __shared__ __align__(128) unsigned char smem_raw[];",
             "unsigned char* smem_raw = (unsigned char*)emu::dyn_shared();", 2),
            ('asm volatile("mov.u32 %0, %smid;" : "=r"(smid));', "smid = 0;", 2),
            ('#include "pf_cg.cuh"', '#include "pf_cg.cuh"\n' + TMA_HOST_PACKETS, 1)):
        assert src.count(old) == times, f"pf_cg_tma.cuh changed: {old!r} occurs {src.count(old)} times"
        src = src.replace(old, new)
    # every `__shared__ T name[dims];` becomes a reference to the block's own zero-initialised storage, so that the
    # grids of several RANKS can run at once in one process (multi-GPU instantiation below)
    import re
    counter = [0]

    def per_block(m):
        counter[0] += 1
        indent, ctype, name, dims, rest = m.groups()
        return (f"{indent}auto& {name} = *reinterpret_cast<{ctype}(*){dims}>"
                f"(emu::shared_slot({counter[0]}, sizeof({ctype}{dims})));{rest}")

    src = re.sub(r"^(\s*)__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)((?:\[[^\]]+\])*);(.*)$", per_block, src, flags=re.M)
    assert "__shared__" not in src and 5 <= counter[0] <= 40, counter[0]
    with open(os.path.join(dest_dir, "pf_cg_tma_emu.cuh"), "w") as fh:
        fh.write(src)


# The multi-GPU register-window kernel lives in pf_cg.cuh, which every other source includes: its section (from the
# "multi-GPU persistent variant" banner to the "split-phase variant" banner) is copied into namespace pf::emu_multi,
# where the PTX helper functions are dropped in favour of host stand-ins of the same names (they shadow the originals),
# `__shared__` declarations get per-block storage, and the two templates that argument-dependent lookup would otherwise
# find twice are renamed.
MULTI_PTX_HELPERS = ["ld_acquire_sys", "st_release_sys", "ld_acquire_gpu", "st_release_gpu", "st_sys_f64", "ld_sys_f64",
                     "ld_relaxed_sys_u64", "gtime"]
MULTI_HEAD = '''// GENERATED (tests only) from the multi-GPU section of pf_cg.cuh.
#pragma once
#include "pf_cg.cuh"
#include <sched.h>
#include <string.h>
namespace pf {
namespace emu_multi {
inline unsigned long long ld_acquire_sys(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long ld_acquire_gpu(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline void st_release_gpu(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
inline void st_sys_f64(double* p, double v) { *(volatile double*)p = v; }
inline double ld_sys_f64(const double* p) { return *(const volatile double*)p; }
inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline unsigned long long gtime() { return (unsigned long long)clock64(); }
'''
def generate_multi_kernel_copy(dest_dir):
    import re
    src = open(os.path.join(ROOT, "pyflow-solver_b200", "csrc", "pf_cg.cuh")).read()
    begin, end = "// ---- multi-GPU persistent variant", "// ---- split-phase variant"
    assert src.count(begin) == 1 and src.count(end) == 1, "pf_cg.cuh changed: section banners not found"
    sec = src[src.index(begin):src.index(end)]
    for name in MULTI_PTX_HELPERS:                       # drop each PTX helper (brace-matched), keep everything else
        m = re.search(r"__device__ __forceinline__ [^\n(]*\b%s\(" % name, sec)
        assert m, f"pf_cg.cuh changed: helper {name} not found"
        j, depth = sec.index("{", m.end()), 0
        while True:
            depth += {"{": 1, "}": -1}.get(sec[j], 0)
            if depth == 0:
                break
            j += 1
        assert "asm volatile" in sec[m.start():j + 1], name
        sec = sec[:m.start()] + sec[j + 1:]
    smid = 'asm volatile("mov.u32 %0, %smid;" : "=r"(smid));'
    assert sec.count(smid) == 1
    sec = sec.replace(smid, "smid = 0;")
    assert "asm volatile" not in sec, "pf_cg.cuh changed: PTX left in the multi-GPU section"
    counter = [19]                                       # ids 1.. are taken by the TMA copy

    def per_block(mm):
        counter[0] += 1
        indent, ctype, name, dims, rest = mm.groups()
        return (f"{indent}auto& {name} = *reinterpret_cast<{ctype}(*){dims}>"
                f"(emu::shared_slot({counter[0]}, sizeof({ctype}{dims})));{rest}")

    sec = re.sub(r"^(\s*)__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)((?:\[[^\]]+\])*);(.*)$", per_block, sec, flags=re.M)
    assert "__shared__" not in sec and counter[0] < 48
    sec = (sec.replace("trace_slot(", "trace_slot_emu(").replace("grid_exchange<", "grid_exchange_emu<")
           .replace("bool grid_exchange(", "bool grid_exchange_emu("))
    with open(os.path.join(dest_dir, "pf_cg_multi_emu.cuh"), "w") as fh:
        fh.write(MULTI_HEAD + sec + "\n}  // namespace emu_multi\n}  // namespace pf\n")


def generate_bicgstab_kernel_copy(dest_dir):
    """pf_bicgstab.cuh again, in namespace pf::emu_bi with per-block storage for its `__shared__` arrays, so that
    several CTAs of k_bicgstab / k_cg_diag can run at once in one process (emu_krylov_grid)."""
    import re
    src = open(os.path.join(ROOT, "pyflow-solver_b200", "csrc", "pf_bicgstab.cuh")).read()
    for old, new in (("#pragma once", ""), ("namespace pf {", "namespace pf {\nnamespace emu_bi {"),
                     ("}  // namespace pf", "}  // namespace emu_bi\n}  // namespace pf")):
        assert src.count(old) == 1, f"pf_bicgstab.cuh changed: {old!r} occurs {src.count(old)} times"
        src = src.replace(old, new)
    counter = [0]

    def per_block(m):
        counter[0] += 1
        indent, ctype, name, dims, rest = m.groups()
        return (f"{indent}auto& {name} = *reinterpret_cast<{ctype}(*){dims}>"
                f"(emu::shared_slot({counter[0]}, sizeof({ctype}{dims})));{rest}")

    src = re.sub(r"^(\s*)__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)((?:\[[^\]]+\])*);(.*)$", per_block, src, flags=re.M)
    assert "__shared__" not in src and counter[0] == 2, counter[0]
    with open(os.path.join(dest_dir, "pf_bicgstab_emu.cuh"), "w") as fh:
        fh.write(src)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    build = str(tmp_path_factory.mktemp("hostemu"))
    out = os.path.join(build, "emu_kernels.so")
    generate_small_kernel_copy(build)
    generate_tma_kernel_copy(build)
    generate_multi_kernel_copy(build)
    generate_bicgstab_kernel_copy(build)
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DPF_HOST_EMULATION", "-DPF_EMU_WITH_SMALL",
           "-DPF_EMU_WITH_TMA", "-DPF_EMU_WITH_MULTI", "-DPF_EMU_WITH_BICG", "-pthread",
           "-fPIC", "-shared", "-I", build,
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
    lib.emu_slot.argtypes = [C.c_char_p]
    lib.emu_prestep.argtypes = [i, i, dp, dp, dp, dp]
    lib.emu_rhs1.argtypes = [i, i, dp, dp, dp, d, d, d, dp, dp]
    lib.emu_rhs2.argtypes = [i, i, dp, dp, dp, i, i, d, dp, dp]
    lib.emu_resid.argtypes = [i, i, dp, dp, d, d, dp, dp]
    lib.emu_divnorm.argtypes = [i, i, dp, dp, d, d, dp, dp]
    lib.emu_post.argtypes = [i, i, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.emu_divc_stats.argtypes = [i, i, dp, dp, d, d, i, dp, dp]
    lib.emu_fused_geometry.argtypes = [i, i, i, C.POINTER(C.c_int)]
    lib.emu_front.argtypes = [i, i, i, dp, dp, dp, dp, dp, d, d, d, d, i, dp, dp]
    lib.emu_finish.argtypes = [i, i, i, dp, dp, dp, dp, dp, dp, d, d, d, i, d, dp, i, dp, dp]
    lib.emu_cg_generic.argtypes = [i, i, i, dp, dp, dp, dp, dp, d, d, d, d, i, dp, dp]
    lib.emu_cg_persistent.argtypes = [i, i, dp, dp, dp, dp, dp, d, d, d, d, i, i, dp, dp]
    lib.emu_krylov.argtypes = [i, i, i, dp, dp, dp, dp, dp, dp, dp, d, d, d, d, i, dp, dp]
    lib.emu_krylov_grid.argtypes = [i] + lib.emu_krylov.argtypes
    lib.emu_cg_split.argtypes = [i, i, i, dp, dp, dp, dp, dp, d, d, d, d, i, i, dp, dp, dp]
    lib.emu_cg_small.argtypes = [i, i, i, dp, dp, dp, dp, dp, C.POINTER(C.c_ulonglong), d, d, d, d, i]
    lib.emu_cg_tmad.argtypes = [i, i, i, i, i, dp, dp, dp, dp, dp, d, d, d, d, i, dp, dp, C.POINTER(C.c_ulonglong)]
    ppd = C.POINTER(dp)
    lib.emu_cg_tmad_multi.argtypes = [i, i, i, i, i, i, ppd, ppd, ppd, ppd, ppd, d, d, d, d, i, ppd, ppd,
                                      C.POINTER(C.POINTER(C.c_ubyte))]
    lib.emu_cg_persistent_multi.argtypes = lib.emu_cg_tmad_multi.argtypes
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


def test_general_spacing_operators_bit_exact_over_random_spacings(emu):
    """Every operator kernel with 600 random spacing pairs -- arbitrary doubles, 1/n (the lx = ly = 1 grids), powers of
    two, the neighbours of 1.0 -- and field magnitudes 1e-8 .. 1e8, zeros of both signs included: the kernels' divisions
    (div_by_invariant: reciprocal + two FMA corrections) against numpy's IEEE divisions, bit for bit.  (79 095 pairs in
    a two-minute run of the same loop when the routine was written: all identical.)"""
    rng = np.random.default_rng(3)
    for _ in range(600):
        ny, nx = int(rng.integers(2, 12)), int(rng.integers(2, 40))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            dx, dy = float(rng.uniform(1e-4, 10)), float(rng.uniform(1e-4, 10))
        elif kind == 1:
            dx, dy = 1.0 / int(rng.integers(2, 5000)), 1.0 / int(rng.integers(2, 5000))
        elif kind == 2:
            dx, dy = float(2.0 ** rng.integers(-20, 20)), float(rng.uniform(1e-4, 10))
        else:
            dx, dy = float(np.nextafter(1.0, 2.0)), float(np.nextafter(1.0, 0.0))
        scale = float(10.0 ** rng.integers(-8, 8))
        p, u, v = (scale * rng.standard_normal((ny, nx)) for _ in range(3))
        u[rng.random((ny, nx)) < 0.2] = 0.0
        p[rng.random((ny, nx)) < 0.05] = -0.0
        dev = Dev(emu, ny, nx)
        gx, gy, o = dev.put(), dev.put(), dev.put()
        emu.emu_op_grad(nx, ny, ptr(dev.put(p)), ptr(gx), ptr(gy), dx, dy)
        ogx, ogy = orc.gradient_coherent(p, dx, dy)
        assert same(dev.get(gx), ogx) and same(dev.get(gy), ogy), (dx, dy)
        emu.emu_op_lap(nx, ny, ptr(dev.put(p)), ptr(o), dx, dy)
        assert same(dev.get(o), orc.laplacian_coherent(p, dx, dy)), (dx, dy)
        for central, fn in ((0, orc.divergence_coherent), (1, orc.divergence_central)):
            emu.emu_op_div(nx, ny, ptr(dev.put(u)), ptr(dev.put(v)), ptr(o), dx, dy, central)
            assert same(dev.get(o), fn(u, v, dx, dy)), (central, dx, dy)
        for scheme, fn in ((UPWIND, orc.advect_upwind), (QUICK, orc.advect_quick)):
            emu.emu_op_advect(scheme, nx, ny, ptr(dev.put(u)), ptr(dev.put(v)), ptr(dev.put(p)), ptr(o), dx, dy)
            assert same(dev.get(o), fn(u, v, p, dx, dy)), (scheme, dx, dy)
        emu.emu_op_grad_central(nx, ny, ptr(dev.put(p)), ptr(gx), ptr(gy), dx, dy)
        ogx, ogy = orc.gradient_central(p, dx, dy)
        assert same(dev.get(gx), ogx) and same(dev.get(gy), ogy), (dx, dy)
        emu.emu_op_lap_central(nx, ny, ptr(dev.put(p)), ptr(o), dx, dy)
        assert same(dev.get(o), orc.laplacian_central(p, dx, dy)), (dx, dy)


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


# ===================================================================================================================
# kernels with grid reductions and the plain CG kernel: one OS thread per CUDA thread
# ===================================================================================================================
class Scalars:
    def __init__(self, lib):
        if not lib.emu_thread_probe(256):
            pytest.skip("this machine cannot run 256 OS threads at once (one per CUDA thread of a block)")
        self.lib = lib
        self.scal = np.zeros(lib.emu_scal_count())
        self.scratch = np.zeros(lib.emu_scratch_doubles())

    def __getitem__(self, name):
        k = self.lib.emu_slot(name.encode())
        assert k >= 0, name
        return float(self.scal[k])


def close(a, b, rtol=1e-13):
    return abs(a - b) <= rtol * max(abs(a), abs(b), 1e-300)


@pytest.mark.parametrize("shape,unit", [((10, 24), True), ((5, 300), False), ((2, 2), True)])
def test_reduction_kernels_vs_numpy(emu, shape, unit):
    ny, nx = shape
    rng = np.random.default_rng(ny * 31 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    dt = 0.0123
    u, v, uo, vo, p = (rng.standard_normal(shape) for _ in range(5))
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    du, dv_ = dev.put(u), dev.put(v)
    # k_prestep: solver.py:25-26,45
    emu.emu_prestep(nx, ny, ptr(du), ptr(dv_), ptr(S.scratch), ptr(S.scal))
    assert S["S_UMAX"] == np.abs(u).max() and S["S_VMAX"] == np.abs(v).max() and S["S_UNAN"] == 0.0
    dc = orc.divergence_central(u, v, 1.0, 1.0)
    assert close(S["S_DIVC_SUMSQ"], float(np.sum(dc * dc)))
    # k_rhs1: pressure_solver.py:116-123 -- b bit-exact, sums in another order
    b = dev.put()
    emu.emu_rhs1(nx, ny, ptr(du), ptr(dv_), ptr(b), dt, dx, dy, ptr(S.scratch), ptr(S.scal))
    div = orc.divergence_coherent(u, v, dx, dy)
    rhs = -div / dt
    assert same(dev.get(b), rhs)
    assert close(S["S_DIVB_SUMSQ"], float(np.sum(div * div))) and close(S["S_RHS_SUM"], float(rhs.sum()), 1e-11)
    # k_rhs2: mean shift with the DEVICE's sum, pin, r = b, x = 0, b.b, b.Mb
    x, r = dev.put(np.ones(shape)), dev.put()
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
    emu.emu_rhs2(nx, ny, ptr(b), ptr(x), ptr(r), 1, 1, invd, ptr(S.scratch), ptr(S.scal))
    want = rhs - S["S_RHS_SUM"] / (nx * ny)
    want[0, 0] = 0.0
    assert same(dev.get(b), want) and same(dev.get(r), want) and not dev.get(x).any()
    M = orc.jacobi_inverse_diagonal(nx, ny, dx, dy).reshape(shape)
    assert close(S["S_BB"], float(np.sum(want * want))) and close(S["S_BZ"], float(np.sum(want * (M * want))))
    # k_resid: ||b - A x||^2 with the pinned operator, x[0] recorded
    xr = rng.standard_normal(shape)
    emu.emu_resid(nx, ny, ptr(b), ptr(dev.put(xr)), dx, dy, ptr(S.scratch), ptr(S.scal))
    res = want.ravel() - orc.neg_laplacian_pinned(xr.ravel(), nx, ny, dx, dy)
    assert close(S["S_RES_SUMSQ"], float(res @ res)) and S["S_X0"] == xr[0, 0]
    # k_divnorm
    emu.emu_divnorm(nx, ny, ptr(du), ptr(dv_), dx, dy, ptr(S.scratch), ptr(S.scal))
    assert close(S["S_DIVA_SUMSQ"], float(np.sum(div * div)))
    # k_post: solver.py:108-121, next step's maxima, the driver's centre line
    center = np.full(ny, np.nan)
    emu.emu_post(nx, ny, ptr(du), ptr(dv_), ptr(dev.put(uo)), ptr(dev.put(vo)), ptr(dev.put(p)), ptr(center),
                 ptr(S.scratch), ptr(S.scal))
    assert close(S["S_RU_SUMSQ"], float(np.sum((u - uo) ** 2))) and close(S["S_RV_SUMSQ"], float(np.sum((v - vo) ** 2)))
    assert close(S["S_DIVC_SUMSQ"], float(np.sum(dc * dc))) and S["S_NONFINITE"] == 0.0
    assert S["S_UMAX"] == np.abs(u).max() and same(center, u[:, nx // 2])
    # ... and a NaN is counted, not lost
    un = u.copy()
    un[ny // 2, nx // 2] = np.nan
    emu.emu_post(nx, ny, ptr(dev.put(un)), ptr(dv_), ptr(dev.put(uo)), ptr(dev.put(vo)), ptr(dev.put(p)), ptr(center),
                 ptr(S.scratch), ptr(S.scal))
    assert S["S_NONFINITE"] > 0.0 and S["S_UNAN"] == 1.0
    # k_divc_stats (alternate driver): mean, then mean-free norm
    emu.emu_divc_stats(nx, ny, ptr(du), ptr(dv_), dx, dy, 0, ptr(S.scratch), ptr(S.scal))
    emu.emu_divc_stats(nx, ny, ptr(du), ptr(dv_), dx, dy, 1, ptr(S.scratch), ptr(S.scal))
    dcs = orc.divergence_central(u, v, dx, dy)
    assert close(S["S_TMP0"] / (nx * ny), float(dcs.mean()), 1e-10)
    assert close(S["S_TMP1"], float(np.sum((dcs - S["S_TMP0"] / (nx * ny)) ** 2)), 1e-12)


def _cg_generic(emu, mode, b, dx, dy, invd, tol, maxiter):
    ny, nx = b.shape
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    x, r, p, q = dev.put(), dev.put(), dev.put(), dev.put()
    part = np.zeros(emu.emu_cg_part_doubles())
    emu.emu_cg_generic(mode, nx, ny, ptr(dev.put(b)), ptr(x), ptr(r), ptr(p), ptr(q), dx, dy, invd, tol, maxiter,
                       ptr(part), ptr(S.scal))
    return dev.get(x), int(S["S_INFO"]), int(S["S_ITERS"]), S["S_RES_SUMSQ"] ** 0.5


@pytest.mark.parametrize("shape,unit,jacobi,maxiter", [((8, 10), True, True, 2000), ((7, 5), False, True, 2000),
                                                       ((6, 6), True, False, 2000), ((8, 10), True, True, 7)])
def test_plain_cg_kernel_follows_scipys_recurrence(emu, shape, unit, jacobi, maxiter):
    """k_cg_generic, pinned operator: x0 = 0, strict test at the top of the loop, info = maxiter when exhausted."""
    ny, nx = shape
    rng = np.random.default_rng(5 * ny + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    b = rng.standard_normal(shape)
    b -= b.mean()
    b[0, 0] = 0.0
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jacobi else 1.0
    x, info, iters, rnorm = _cg_generic(emu, 0, b, dx, dy, invd, 1e-10, maxiter)
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    xo, io, to = orc.pcg(A, b.ravel().copy(), 1e-10, maxiter, inv)
    assert info == io and abs(iters - to) <= 1
    assert np.linalg.norm(x.ravel() - xo) <= 1e-9 * np.linalg.norm(xo)
    assert close(rnorm, float(np.linalg.norm(b.ravel() - A(x.ravel()))), 1e-6)
    zero = _cg_generic(emu, 0, np.zeros(shape), dx, dy, invd, 1e-10, 50)
    assert zero[1] == 0 and zero[2] == 0 and not zero[0].any()       # b = 0  =>  x = 0, no iterations


@pytest.mark.parametrize("tag", ["24x20_plain", "24x20_jacobi", "33x17_jacobi_nonunit", "40x28_truncated7",
                                 "16x16_default_tol", "12x10_zero_rhs"])
def test_bicgstab_kernel_vs_reference_golden(emu, tag):
    """k_bicgstab executed on the host against what the UNMODIFIED reference's interface.solve (default method,
    scipy's bicgstab) returned for the same right-hand side (tests/golden/bicgstab.npz): the truncated solve is the
    same iterate up to summation order; converged solves meet the reference's own convergence criterion and agree
    with its solution to the accuracy the tolerance allows; exit bookkeeping (iterations = info if info > 0) equal."""
    import os
    from conftest import GOLDEN
    if tag == "33x17_jacobi_nonunit" and os.environ.get("PYFLOW_SLOW_HOSTEMU") != "1":
        pytest.skip("688 emulated iterations (45 s): PYFLOW_SLOW_HOSTEMU=1 runs it; the GPU and oracle suites keep the case")
    with np.load(os.path.join(GOLDEN, "bicgstab.npz")) as z:
        g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/")}
    nx, ny, dx, dy, jac, tol, maxiter = g["params"]
    nx, ny, maxiter = int(nx), int(ny), int(maxiter)
    b = g["b"].reshape(ny, nx)
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    x, r, p, v, t = (dev.put() for _ in range(5))
    part = np.zeros(emu.emu_cg_part_doubles())
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jac else 1.0
    emu.emu_krylov(1, nx, ny, ptr(dev.put(b)), None, ptr(x), ptr(r), ptr(p), ptr(v), ptr(t), dx, dy, invd, tol,
                   maxiter, ptr(part), ptr(S.scal))
    xs, info, iters, rnorm = dev.get(x).ravel(), int(S["S_INFO"]), int(S["S_ITERS"]), S["S_RES_SUMSQ"] ** 0.5
    conv_ref, its_ref, rnorm_ref, done_ref = g["result"]
    bnorm = float(np.linalg.norm(g["b"])) or 1.0
    A = lambda z_: orc.neg_laplacian_pinned(z_, nx, ny, dx, dy)          # noqa: E731
    assert close(rnorm, float(np.linalg.norm(g["b"] - A(xs))), 1e-6) or rnorm == 0.0
    assert (info if info > 0 else 0) == int(its_ref)
    assert bool(info == 0 and rnorm <= tol * bnorm) == bool(conv_ref)
    if "truncated" in tag:
        assert iters == int(done_ref) == maxiter
        assert np.linalg.norm(xs - g["x"]) <= 1e-11 * np.linalg.norm(g["x"])
    elif "zero_rhs" in tag:
        assert iters == 0 and not xs.any()
    else:
        # both are solutions of the same system to ||b - A x|| <= tol ||b||: they differ by at most
        # 2 tol ||b|| / sigma_min(A)  (sigma_min of the pinned operator on these grids > 1e-3)
        assert np.linalg.norm(xs - g["x"]) <= 2.0 * tol * bnorm / 1e-3
        assert abs(iters - int(done_ref)) <= max(3, int(0.25 * done_ref))


@pytest.mark.parametrize("method,shape,unit,maxiter", [(0, (9, 13), True, 9), (1, (9, 13), True, 9),
                                                       (0, (12, 7), False, 3000), (1, (12, 7), False, 3000)])
def test_krylov_kernels_with_a_user_supplied_diagonal(emu, method, shape, unit, maxiter):
    """k_cg_diag / k_bicgstab reading the preconditioner's inverse entries from an array (a user-supplied diagonal
    M of the linear-solve seam) against the oracle recurrences with the same diagonal."""
    ny, nx = shape
    rng = np.random.default_rng(17 * ny + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    b = rng.standard_normal(shape)
    b -= b.mean()
    b[0, 0] = 0.0
    minv = rng.uniform(0.05, 1.0, shape)
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    x, r, p, v, t = (dev.put() for _ in range(5))
    part = np.zeros(emu.emu_cg_part_doubles())
    emu.emu_krylov(method, nx, ny, ptr(dev.put(b)), ptr(dev.put(minv)), ptr(x), ptr(r), ptr(p), ptr(v), ptr(t), dx, dy,
                   1.0, 1e-10, maxiter, ptr(part), ptr(S.scal))
    xs, info, iters, rnorm = dev.get(x).ravel(), int(S["S_INFO"]), int(S["S_ITERS"]), S["S_RES_SUMSQ"] ** 0.5
    A = lambda z_: orc.neg_laplacian_pinned(z_, nx, ny, dx, dy)          # noqa: E731
    xo, io, to = (orc.bicgstab if method else orc.pcg)(A, b.ravel().copy(), 1e-10, maxiter, minv.ravel().copy())
    assert info == io and close(rnorm, float(np.linalg.norm(b.ravel() - A(xs))), 1e-6)
    if maxiter < 100:
        assert iters == to == maxiter and np.linalg.norm(xs - xo) <= 1e-11 * np.linalg.norm(xo)
    else:
        assert info == 0 and rnorm <= 1e-10 * np.linalg.norm(b)
        assert np.linalg.norm(xs - xo) <= 1e-6 * np.linalg.norm(xo) and abs(iters - to) <= max(3, to // 4)


@pytest.mark.parametrize("method,jacobi", [(1, True), (0, False), (1, False)])
def test_krylov_kernels_on_three_ctas_at_once(emu, method, jacobi):
    """k_bicgstab / k_cg_diag on a cooperative grid of THREE blocks executing concurrently (768 OS threads, a real
    grid-wide barrier): per-CTA partials folded in CTA order by every CTA, single-use partial slots, identical
    exits in all CTAs.  Truncated solves against the oracle recurrences (summation order only); the analytic ring
    diagonal (minv = NULL) for BiCGSTAB, a user-supplied diagonal array otherwise."""
    ny, nx = 24, 37                                        # 888 cells on 768 threads: ragged last stride
    rng = np.random.default_rng(91 + method)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    dx, dy = 0.5, 0.25
    if jacobi:
        minv, invd = None, 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
        diag = orc.jacobi_inverse_diagonal(nx, ny, dx, dy)
    else:
        minv, invd = rng.uniform(0.05, 1.0, (ny, nx)), 1.0
        diag = minv.ravel().copy()
    if not emu.emu_thread_probe(3 * 256):
        pytest.skip("this machine cannot run 768 OS threads at once")
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    x, r, p, v, t = (dev.put() for _ in range(5))
    part = np.full(emu.emu_cg_part_doubles(), np.nan)      # a slot read before it was written would poison the sums
    emu.emu_krylov_grid(3, method, nx, ny, ptr(dev.put(b)), None if minv is None else ptr(dev.put(minv)), ptr(x),
                        ptr(r), ptr(p), ptr(v), ptr(t), dx, dy, invd, 1e-10, 8, ptr(part), ptr(S.scal))
    xs, info, iters = dev.get(x).ravel(), int(S["S_INFO"]), int(S["S_ITERS"])
    A = lambda z_: orc.neg_laplacian_pinned(z_, nx, ny, dx, dy)          # noqa: E731
    xo, io, to = (orc.bicgstab if method else orc.pcg)(A, b.ravel().copy(), 1e-10, 8, diag)
    assert info == io == 8 and iters == to == 8
    assert np.linalg.norm(xs - xo) <= 1e-11 * np.linalg.norm(xo)
    assert close(S["S_RES_SUMSQ"] ** 0.5, float(np.linalg.norm(b.ravel() - A(xs))), 1e-6)


def test_plain_cg_kernel_dirichlet_mode_vs_oracle(emu):
    """Manufactured / Dirichlet mode (pressure_solver.py:61-93): +laplacian with identity rows on the ring."""
    nx, ny = 14, 11
    dx, dy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    X, Y = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    rhs = np.zeros((ny, nx))
    rhs[1:-1, 1:-1] = (-2.0 * np.pi ** 2 * np.sin(np.pi * X) * np.sin(np.pi * Y))[1:-1, 1:-1]
    x, info, iters, _rn = _cg_generic(emu, 1, rhs, dx, dy, 1.0, 1e-10, 500)
    from types import SimpleNamespace
    po, od = orc.solve_manufactured(orc.allocate_state(nx, ny), dx, dy,
                                    SimpleNamespace(lin_tol=1e-10, lin_maxiter=500), rhs.reshape(-1))
    assert info == od["info"] and abs(iters - od["true_iterations"]) <= 1
    assert np.linalg.norm(x - po) <= 1e-9 * np.linalg.norm(po)


@pytest.mark.parametrize("chain", ["separate", "fused"])
@pytest.mark.parametrize("scheme_name,nx,ny,spacing,solver", [("quick", 8, 6, None, "generic"),
                                                              ("upwind", 7, 5, (2.0, 1.0), "register-window")])
def test_whole_projection_steps_from_the_cuda_sources_on_the_host(emu, scheme_name, nx, ny, spacing, solver, chain):
    """The kernels chained in the order of pf_step (pf_api.cu) -- the solve by the plain CG kernel or by the
    register-window kernel the library itself picks for an odd width --, host dt logic from the oracle: three steps
    of a driven cavity against the oracle.  chain "separate": the seven per-step kernels of pf_kernels.cuh (several
    GPUs, tiny grids); "fused": the two passes of pf_step_fused.cuh one GPU steps with."""
    from types import SimpleNamespace
    cfg = SimpleNamespace(nx=nx, ny=ny, Re=100.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=3000, advection_scheme=scheme_name, disable_advection=False, diagnostics=False)
    if spacing:
        cfg.lx, cfg.ly = spacing
    scheme = QUICK if scheme_name == "quick" else UPWIND
    ost = orc.allocate_state(nx, ny)
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    u, v, uo, vo = dev.put(), dev.put(), dev.put(), dev.put()
    b, x, r, p, q = dev.put(), dev.put(), dev.put(), dev.put(), dev.put()
    part = np.zeros(emu.emu_cg_part_doubles())
    center = np.zeros(ny)
    for it in range(3):
        _, ores, od = orc.step(cfg, ost, it)
        dt, dx, dy = od["dt"], ost.meta["dx"], ost.meta["dy"]
        invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
        if chain == "fused":
            assert emu.emu_front(scheme, nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), ptr(b), dt, max(dt, 1e-14), dx, dy,
                                 148, ptr(S.scratch), ptr(S.scal))
            u, uo, v, vo = uo, u, vo, v
        else:
            emu.emu_predict(scheme, nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), dt, dx, dy)
            u, uo, v, vo = uo, u, vo, v                               # u, v hold u*, v*; uo, vo the previous step
            emu.emu_rhs1(nx, ny, ptr(u), ptr(v), ptr(b), max(dt, 1e-14), dx, dy, ptr(S.scratch), ptr(S.scal))
        emu.emu_rhs2(nx, ny, ptr(b), ptr(x), ptr(r), 1, 1, invd, ptr(S.scratch), ptr(S.scal))
        if solver == "generic":
            emu.emu_cg_generic(0, nx, ny, ptr(b), ptr(x), ptr(r), ptr(p), ptr(q), dx, dy, invd, 1e-10, 3000, ptr(part),
                               ptr(S.scal))
        else:
            emu.emu_cg_persistent(nx, ny, ptr(b), ptr(x), ptr(r), ptr(p), ptr(q), dx, dy, invd, 1e-10, 3000, 4,
                                  ptr(part), ptr(S.scal))
        assert abs(int(S["S_ITERS"]) - od["true_iterations"]) <= 1
        if chain == "fused":
            assert emu.emu_finish(1, nx, ny, ptr(b), ptr(x), ptr(u), ptr(v), ptr(uo), ptr(vo), dt, dx, dy, 1, 1.0,
                                  ptr(center), 148, ptr(S.scratch), ptr(S.scal))
            u, uo, v, vo = uo, u, vo, v                               # the new fields landed in the other buffer pair
            assert close(S["S_DIVA_SUMSQ"] ** 0.5, od["divergence_norm"], 1e-6)
            # a converged solve stops at the noise level of its own summation order: same magnitude, not same digits
            assert 0.3 * od["Rp_residual"] <= S["S_RES_SUMSQ"] ** 0.5 <= 3.0 * od["Rp_residual"] + 1e-300
        else:
            emu.emu_resid(nx, ny, ptr(b), ptr(x), dx, dy, ptr(S.scratch), ptr(S.scal))
            emu.emu_correct(nx, ny, ptr(x), ptr(u), ptr(v), dt, dx, dy, S["S_X0"])
            emu.emu_divnorm(nx, ny, ptr(u), ptr(v), dx, dy, ptr(S.scratch), ptr(S.scal))
            assert close(S["S_DIVA_SUMSQ"] ** 0.5, od["divergence_norm"], 1e-6)
            emu.emu_lid(nx, ny, ptr(u), 1.0)
            emu.emu_post(nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), ptr(x), ptr(center), ptr(S.scratch), ptr(S.scal))
        assert close(S["S_RU_SUMSQ"] ** 0.5 / (nx * ny), ores["Ru"], 1e-8) and S["S_NONFINITE"] == 0.0
        assert close(S["S_DIVC_SUMSQ"] ** 0.5 / (nx * ny), ores["continuity"], 1e-8)
    for name, buf in (("u", u), ("v", v), ("p", x)):
        a, want = dev.get(buf), orc.interior(ost.fields[name])
        assert np.linalg.norm(a - want) <= 1e-9 * (np.linalg.norm(want) or 1.0), name
    assert same(center, dev.get(u)[:, nx // 2])


# ===================================================================================================================
# the two fused per-step passes (pf_step_fused.cuh) against the seven separate kernels they replace on one GPU
# ===================================================================================================================
# shape (ny, nx), unit spacing?, SMs the launch geometry is computed for (fewer SMs: taller bands).  Widths chosen around
# the 58 columns a warp owns (one warp, warp boundaries, odd widths, a second column block) and bands of 2 .. ny rows.
FUSED_CASES = [((9, 13), True, 148), ((14, 300), False, 148), ((11, 513), True, 148), ((7, 59), True, 2),
               ((12, 117), False, 3), ((13, 500), True, 148), ((6, 4), True, 148), ((4, 5), False, 1),
               ((5, 58), True, 1), ((17, 57), False, 148), ((6, 930), True, 148)]
FUSED_SLOW = {(14, 300), (11, 513), (17, 57), (6, 930)}      # run with PYFLOW_SLOW_HOSTEMU=1 (the CPU suite stays short)


def _skip_slow_fused(shape):
    if shape in FUSED_SLOW and os.environ.get("PYFLOW_SLOW_HOSTEMU") != "1":
        pytest.skip("wide / tall layout (several seconds of emulated threads): PYFLOW_SLOW_HOSTEMU=1 runs it")


@pytest.mark.parametrize("shape,unit,sms", FUSED_CASES)
def test_fused_front_pass_equals_predictor_plus_rhs_pass_1(emu, shape, unit, sms):
    _skip_slow_fused(shape)
    ny, nx = shape
    rng = np.random.default_rng(ny * 131 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    dt = 0.0123
    u, v = rng.standard_normal(shape), rng.standard_normal(shape)
    u[rng.random(shape) < 0.1] = 0.0                       # QUICK's >= against upwind's >
    v[rng.random(shape) < 0.1] = 0.0
    dev, S, F = Dev(emu, ny, nx), Scalars(emu), Scalars(emu)
    geo = (C.c_int * 4)()
    assert emu.emu_fused_geometry(nx, ny, sms, geo) == 1 and geo[0] * geo[1] >= 1 and geo[3] >= 2
    for scheme in (QUICK, UPWIND, NONE):
        du, dv_ = dev.put(u), dev.put(v)
        us, vs, b = dev.put(), dev.put(), dev.put()
        emu.emu_predict(scheme, nx, ny, ptr(du), ptr(dv_), ptr(us), ptr(vs), dt, dx, dy)
        emu.emu_rhs1(nx, ny, ptr(us), ptr(vs), ptr(b), dt, dx, dy, ptr(S.scratch), ptr(S.scal))
        fus, fvs, fb = dev.put(), dev.put(), dev.put()
        assert emu.emu_front(scheme, nx, ny, ptr(du), ptr(dv_), ptr(fus), ptr(fvs), ptr(fb), dt, dt, dx, dy, sms,
                             ptr(F.scratch), ptr(F.scal)) == 1
        # dev.get also asserts that halo rows and padding columns were never stored to
        assert same(dev.get(fus), dev.get(us)) and same(dev.get(fvs), dev.get(vs)) and same(dev.get(fb), dev.get(b)), scheme
        assert close(F["S_DIVB_SUMSQ"], S["S_DIVB_SUMSQ"]) and close(F["S_RHS_SUM"], S["S_RHS_SUM"], 1e-10), scheme


@pytest.mark.parametrize("poison", [None, "nan_u", "nan_x0", "inf_v"])
@pytest.mark.parametrize("shape,unit,sms", FUSED_CASES)
def test_fused_finish_pass_equals_the_five_kernels_after_the_solve(emu, shape, unit, sms, poison):
    _skip_slow_fused(shape)
    ny, nx = shape
    if poison and (ny, nx) not in ((9, 13), (12, 117)):
        pytest.skip("poisoned inputs on two shapes only")
    rng = np.random.default_rng(ny * 173 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.37, 1.9)
    dt, lid = 0.0123, 1.25
    b, x, us, vs, uo, vo = (rng.standard_normal(shape) for _ in range(6))
    x[0, 0] = 0.0                                          # the pin keeps x[0] at exactly 0 (or NaN)
    if poison == "nan_u":
        us[ny // 2, nx // 2] = np.nan
    elif poison == "nan_x0":
        x[0, 0] = np.nan
    elif poison == "inf_v":
        vs[ny - 2, 1] = np.inf
    dev, S, F = Dev(emu, ny, nx), Scalars(emu), Scalars(emu)
    with np.errstate(invalid="ignore"):
        for post in (1, 0):
            # the separate kernels in the order of pf_step
            db, dxx, du, dv_, duo, dvo = dev.put(b), dev.put(x), dev.put(us), dev.put(vs), dev.put(uo), dev.put(vo)
            center = np.full(ny, np.nan)
            emu.emu_resid(nx, ny, ptr(db), ptr(dxx), dx, dy, ptr(S.scratch), ptr(S.scal))
            emu.emu_correct(nx, ny, ptr(dxx), ptr(du), ptr(dv_), dt, dx, dy, S["S_X0"])
            emu.emu_divnorm(nx, ny, ptr(du), ptr(dv_), dx, dy, ptr(S.scratch), ptr(S.scal))
            if post:
                emu.emu_lid(nx, ny, ptr(du), lid)
                emu.emu_post(nx, ny, ptr(du), ptr(dv_), ptr(duo), ptr(dvo), ptr(dxx), ptr(center), ptr(S.scratch),
                             ptr(S.scal))
            # the fused pass: new fields land in the buffers that held the previous step's
            fb, fx, fu, fv, fuo, fvo = dev.put(b), dev.put(x), dev.put(us), dev.put(vs), dev.put(uo), dev.put(vo)
            fcenter = np.full(ny, np.nan)
            assert emu.emu_finish(post, nx, ny, ptr(fb), ptr(fx), ptr(fu), ptr(fv), ptr(fuo), ptr(fvo), dt, dx, dy, 1,
                                  lid, ptr(fcenter), sms, ptr(F.scratch), ptr(F.scal)) == 1
            assert same(dev.get(fuo), dev.get(du)) and same(dev.get(fvo), dev.get(dv_)), (post, poison)
            assert same(dev.get(fx), dev.get(dxx)), "p = x - x[0]"
            assert same(dev.get(fu), us) and same(dev.get(fv), vs) and same(dev.get(fb), b), "inputs are read-only"
            names = ["S_RES_SUMSQ", "S_DIVA_SUMSQ"]
            if post:
                names += ["S_RU_SUMSQ", "S_RV_SUMSQ", "S_DIVC_SUMSQ", "S_NONFINITE", "S_UNAN", "S_VNAN", "S_UMAX", "S_VMAX"]
                assert same(fcenter, center)
            for name in names:
                a, w = F[name], S[name]
                assert (np.isnan(a) and np.isnan(w)) or a == w or close(a, w, 1e-12), (name, a, w, post, poison)
            xs, fs = S["S_X0"], F["S_X0"]
            assert xs == fs or (np.isnan(xs) and np.isnan(fs))
            if poison == "nan_u" and post:
                assert F["S_UNAN"] == 1.0 and F["S_NONFINITE"] >= 1.0
            if poison == "nan_x0":
                assert np.isnan(dev.get(fx)).all() and np.isnan(dev.get(fvo)[:-1]).all()     # p = x - NaN everywhere


def test_fused_geometry_covers_every_row_and_column_once(emu):
    """Every cell belongs to exactly one (column block, warp, owning lane) and one band -- for the grids of the
    BASELINE configs and a few awkward ones."""
    geo = (C.c_int * 4)()
    for nx, ny, sms in ((64, 64, 148), (512, 512, 148), (4096, 4096, 148), (8192, 8192, 148), (16384, 16384, 148),
                        (4000, 4000, 148), (5000, 3000, 148), (4097, 1500, 132), (59, 7, 2), (4, 4, 148), (100000, 8, 148)):
        assert emu.emu_fused_geometry(nx, ny, sms, geo) == 1, (nx, ny)
        gx, gy, threads, rows = geo[0], geo[1], geo[2], geo[3]
        assert threads % 32 == 0 and 32 <= threads <= 256 and gx * gy <= 4096
        assert gx * (threads // 32) * 58 >= nx > (gx - 1) * (threads // 32) * 58          # columns covered, no empty block
        assert gy * rows >= ny > (gy - 1) * rows                                           # rows covered, no empty band
    for nx, ny in ((3, 100), (100, 3), (2, 2)):
        assert emu.emu_fused_geometry(nx, ny, 148, geo) == 0        # tiny grids keep the separate kernels


# ===================================================================================================================
# the register-window persistent CG kernel (production kernel for widths that are not a multiple of 1024)
# ===================================================================================================================
def _cg_persistent(emu, b, dx, dy, jacobi, tol, maxiter):
    ny, nx = b.shape
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jacobi else 1.0
    bb, x, r, p0, p1 = dev.put(b), dev.put(), dev.put(), dev.put(), dev.put()
    emu.emu_rhs2(nx, ny, ptr(bb), ptr(x), ptr(r), 0, 0, invd, ptr(S.scratch), ptr(S.scal))     # r = b, x = 0, b.b, b.Mb
    part = np.zeros(emu.emu_cg_part_doubles())
    emu.emu_cg_persistent(nx, ny, ptr(bb), ptr(x), ptr(r), ptr(p0), ptr(p1), dx, dy, invd, tol, maxiter, 4, ptr(part),
                          ptr(S.scal))
    return dev.get(x), dev.get(r), int(S["S_INFO"]), int(S["S_ITERS"])


@pytest.mark.parametrize("ny,nx,unit,jacobi,maxiter", [
    (3, 3, True, True, 300), (5, 3, True, True, 300), (4, 5, False, True, 300),       # the smallest odd-width grids
    (33, 17, False, True, 10), (12, 65, True, True, 12), (9, 129, True, False, 10),    # one partial strip; nx % 64 == 1
    (10, 200, False, True, 8), (40, 70, True, True, 10),                               # wall + partial strips, 5 rows per warp
    (7, 577, True, True, 8),                                                           # two chunks of eight strips
])
def test_register_window_cg_kernel_vs_oracle(emu, ny, nx, unit, jacobi, maxiter):
    """k_cg_persistent on one CTA against scipy's recurrence (oracle.pcg): same stopping decisions, x and the
    recursive residual to 1e-10, on the column layouts that select its strip kinds (inside / wall / partial)."""
    rng = np.random.default_rng(ny * 1009 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    x, r, info, iters = _cg_persistent(emu, b, dx, dy, jacobi, 1e-10, maxiter)
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    xo, io, to = orc.pcg(A, b.ravel().copy(), 1e-10, maxiter, inv)
    assert info == io and abs(iters - to) <= 1, (info, io, iters, to)
    assert np.isfinite(x).all() and np.isfinite(r).all()
    if iters == to:
        assert np.linalg.norm(x.ravel() - xo) <= 1e-10 * np.linalg.norm(xo)
    else:
        assert np.linalg.norm(x.ravel() - xo) <= 1e-8 * np.linalg.norm(xo)


@pytest.mark.parametrize("ny,nx,G,unit,maxiter,tiles", [
    (70, 17, 2, True, 6, 2),          # one partial strip, two bands of 35 rows
    (40, 200, 3, False, 6, 3),        # wall + partial strips, three bands
    (16, 577, 4, True, 5, 4),         # two column chunks x two bands, four CTAs
    (34, 66, 2, True, 200, 2),        # runs to convergence: the device-side done flag stops the sequence
    (12, 20, 3, True, 4, 1),          # more CTAs than tiles: idle CTAs take part in the folds
])
def test_split_phase_cg_kernels_vs_oracle(emu, ny, nx, G, unit, maxiter, tiles):
    """k_cg_begin / phaseA / phaseB / update / finish -- the sequence used between NCCL calls on multi-GPU slabs
    without peer memory -- for one rank on several CTAs (several row bands per chunk) against scipy's recurrence."""
    assert emu.emu_cg_ntiles(nx, ny, G, 4) == tiles
    rng = np.random.default_rng(ny * 1013 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    b = rng.standard_normal((ny, nx))
    b -= b.mean()
    b[0, 0] = 0.0
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2)
    bb, x, r, p0, p1 = dev.put(b), dev.put(), dev.put(), dev.put(), dev.put()
    emu.emu_rhs2(nx, ny, ptr(bb), ptr(x), ptr(r), 0, 0, invd, ptr(S.scratch), ptr(S.scal))
    part = np.zeros(emu.emu_cg_part_doubles())
    tol = 1e-10 if maxiter < 100 else 0.3
    emu.emu_cg_split(nx, ny, G, ptr(bb), ptr(x), ptr(r), ptr(p0), ptr(p1), dx, dy, invd, tol, maxiter, 4, ptr(part),
                     ptr(S.scal), ptr(S.scratch))
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    xo, io, to = orc.pcg(A, b.ravel().copy(), tol, maxiter, orc.jacobi_inverse_diagonal(nx, ny, dx, dy))
    assert int(S["S_INFO"]) == io and abs(int(S["S_ITERS"]) - to) <= 1
    got = dev.get(x).ravel()
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - xo) <= (1e-10 if int(S["S_ITERS"]) == to else 1e-6) * np.linalg.norm(xo)


@pytest.mark.parametrize("tag", ["3x3_quick", "3x5_upwind"])
def test_smallest_odd_width_cavities_from_the_cuda_sources_vs_reference_golden(emu, golden, tag):
    """The odd-width cases of tests/golden/tiny.npz (trajectories written by the unmodified reference), stepped by
    the kernels in the order of pf_step with the register-window CG kernel -- the one the library picks for them."""
    from conftest import tiny_cfg
    cfg = tiny_cfg(tag)
    nx, ny = cfg.nx, cfg.ny
    g = {k.split("/", 1)[1]: v for k, v in golden["tiny"].items() if k.startswith(tag + "/")}
    scheme = QUICK if cfg.advection_scheme == "quick" else UPWIND
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    u, v, uo, vo = dev.put(), dev.put(), dev.put(), dev.put()
    b, x, r, p, q = dev.put(), dev.put(), dev.put(), dev.put(), dev.put()
    part = np.zeros(emu.emu_cg_part_doubles())
    center = np.zeros(ny)
    for it in range(len(g["true_iters"])):
        dt = float(g["scalars"][it, 0])                     # the reference's own dt sequence (host logic, not kernels)
        emu.emu_predict(scheme, nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), dt, 1.0, 1.0)
        u, uo, v, vo = uo, u, vo, v
        emu.emu_rhs1(nx, ny, ptr(u), ptr(v), ptr(b), max(dt, 1e-14), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_rhs2(nx, ny, ptr(b), ptr(x), ptr(r), 1, 1, 0.25, ptr(S.scratch), ptr(S.scal))
        emu.emu_cg_persistent(nx, ny, ptr(b), ptr(x), ptr(r), ptr(p), ptr(q), 1.0, 1.0, 0.25, 1e-10, 300, 4, ptr(part),
                              ptr(S.scal))
        assert abs(int(S["S_ITERS"]) - int(g["true_iters"][it])) <= 1 and int(S["S_INFO"]) == 0, (tag, it)
        emu.emu_resid(nx, ny, ptr(b), ptr(x), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_correct(nx, ny, ptr(x), ptr(u), ptr(v), dt, 1.0, 1.0, S["S_X0"])
        emu.emu_lid(nx, ny, ptr(u), 1.0)
        emu.emu_post(nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), ptr(x), ptr(center), ptr(S.scratch), ptr(S.scal))
        assert close(S["S_DIVC_SUMSQ"] ** 0.5 / (nx * ny), float(g["scalars"][it, 4]), 1e-7), (tag, it)      # continuity
    for name, buf in (("u", u), ("v", v), ("p", x)):
        a, want = dev.get(buf), g[name][1:-1, 1:-1]
        assert np.linalg.norm(a - want) <= 1e-9 * (np.linalg.norm(want) or 1.0), (tag, name)


# ===================================================================================================================
# the shared-memory-resident CG kernel (k_cg_small): ALL its CTAs at once, talking through tagged packets
# ===================================================================================================================
class SmallSolver:
    """Persistent state of k_cg_small between solves on one grid: packet slots and the barrier epoch."""

    def __init__(self, emu, ny, nx, rows):
        if not emu.emu_thread_probe(512 * ((ny + rows - 1) // rows)):
            pytest.skip("this machine cannot run one OS thread per CUDA thread of every CTA at once")
        self.emu, self.ny, self.nx, self.rows = emu, ny, nx, rows
        G = (ny + rows - 1) // rows
        self.rhalo = np.zeros(G * 2 * emu.emu_small_row_packets(nx) * 4)
        self.part = np.zeros(2 * 160 * 4)                     # [2][SM_MAX_GRID] slots, SM_SLOT_DOUBLES apart
        self.sync = np.zeros(2, dtype=np.uint64)             # epoch (persisted), error flag

    def solve(self, b_dev, x_dev, scal, dx, dy, invd, tol, maxiter):
        self.sync[1] = 0
        self.emu.emu_cg_small(self.nx, self.ny, self.rows, ptr(b_dev), ptr(x_dev), ptr(self.rhalo), ptr(self.part),
                              ptr(scal), self.sync.ctypes.data_as(C.POINTER(C.c_ulonglong)), dx, dy, invd, tol, maxiter)
        assert self.sync[1] == 0, "a bounded wait inside k_cg_small timed out"


@pytest.mark.parametrize("ny,nx,rows,unit,jacobi,maxiter", [
    (2, 2, 1, True, True, 50), (3, 4, 1, True, True, 50), (5, 2, 1, True, True, 50), (2, 6, 1, True, True, 50),
    (4, 6, 1, False, True, 300),          # four CTAs of one row, non-unit spacing, runs to convergence
    (7, 10, 3, True, False, 12),          # bands of 3, 3 and 1 rows, no preconditioner, truncated early
])
def test_shared_memory_cg_kernel_with_concurrent_ctas_vs_oracle(emu, ny, nx, rows, unit, jacobi, maxiter):
    """The real k_cg_small source, one OS thread per CUDA thread of EVERY CTA at once, in the configuration the
    library uses for small grids (bands of ceil(ny / #SMs) rows): halo rows of p recomputed from exchanged rows of r,
    both reductions and the rows travelling as tagged packets, two solves in a row on the same packet slots."""
    rng = np.random.default_rng(ny * 1019 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jacobi else 1.0
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    solver = SmallSolver(emu, ny, nx, rows)
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    for _repeat in range(2):                                          # the epoch carries over to the second solve
        b = rng.standard_normal((ny, nx))
        b -= b.mean()
        b[0, 0] = 0.0
        bb, x, r = dev.put(b), dev.put(), dev.put()
        emu.emu_rhs2(nx, ny, ptr(bb), ptr(x), ptr(r), 0, 0, invd, ptr(S.scratch), ptr(S.scal))
        solver.solve(bb, x, S.scal, dx, dy, invd, 1e-10, maxiter)
        xo, io, to = orc.pcg(A, b.ravel().copy(), 1e-10, maxiter, inv)
        assert int(S["S_INFO"]) == io and abs(int(S["S_ITERS"]) - to) <= 1
        got = dev.get(x).ravel()
        assert np.linalg.norm(got - xo) <= 1e-9 * np.linalg.norm(xo)      # the parity bar of the project
    assert int(solver.sync[0]) > 0


@pytest.mark.parametrize("tag", ["2x2_quick", "2x5_upwind"])
def test_smallest_even_width_cavities_from_the_cuda_sources_vs_reference_golden(emu, golden, tag):
    """The even-width cases of tests/golden/tiny.npz, stepped by the kernels in the order of pf_step with k_cg_small
    on one-row bands (one CTA per grid row, as on the GPU)."""
    from conftest import tiny_cfg
    cfg = tiny_cfg(tag)
    nx, ny = cfg.nx, cfg.ny
    g = {k.split("/", 1)[1]: v for k, v in golden["tiny"].items() if k.startswith(tag + "/")}
    scheme = QUICK if cfg.advection_scheme == "quick" else UPWIND
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    solver = SmallSolver(emu, ny, nx, 1)
    u, v, uo, vo = dev.put(), dev.put(), dev.put(), dev.put()
    b, x, r = dev.put(), dev.put(), dev.put()
    center = np.zeros(ny)
    for it in range(len(g["true_iters"])):
        dt = float(g["scalars"][it, 0])
        emu.emu_predict(scheme, nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), dt, 1.0, 1.0)
        u, uo, v, vo = uo, u, vo, v
        emu.emu_rhs1(nx, ny, ptr(u), ptr(v), ptr(b), max(dt, 1e-14), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_rhs2(nx, ny, ptr(b), ptr(x), ptr(r), 1, 1, 0.25, ptr(S.scratch), ptr(S.scal))
        solver.solve(b, x, S.scal, 1.0, 1.0, 0.25, 1e-10, 300)
        assert abs(int(S["S_ITERS"]) - int(g["true_iters"][it])) <= 1 and int(S["S_INFO"]) == 0, (tag, it)
        emu.emu_resid(nx, ny, ptr(b), ptr(x), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_correct(nx, ny, ptr(x), ptr(u), ptr(v), dt, 1.0, 1.0, S["S_X0"])
        emu.emu_lid(nx, ny, ptr(u), 1.0)
        emu.emu_post(nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), ptr(x), ptr(center), ptr(S.scratch), ptr(S.scal))
    for name, buf in (("u", u), ("v", v), ("p", x)):
        a, want = dev.get(buf), g[name][1:-1, 1:-1]
        assert np.linalg.norm(a - want) <= 1e-9 * (np.linalg.norm(want) or 1.0), (tag, name)


# ===================================================================================================================
# the TMA-fed, dynamically scheduled CG kernel (k_cg_tmad): the production kernel of the 4096^2 ... 16384^2 configs
# ===================================================================================================================
def _tiles(tile_rows):
    """tile_rows: an int (uniform tiles of that height) or (big height, small height, small bands at each end)."""
    return (tile_rows, 0, 0) if isinstance(tile_rows, int) else tuple(tile_rows)


def _tmad_solve(emu, dev, S, sync, b, dx, dy, invd, tol, maxiter, tile_rows):
    ny, nx = b.shape
    bb, x, r, p0, p1 = dev.put(b), dev.put(), dev.put(), dev.put(), dev.put()
    emu.emu_rhs2(nx, ny, ptr(bb), ptr(x), ptr(r), 0, 0, invd, ptr(S.scratch), ptr(S.scal))
    part = np.zeros(emu.emu_cg_part_doubles())
    emu.emu_cg_tmad(nx, ny, *_tiles(tile_rows), ptr(bb), ptr(x), ptr(r), ptr(p0), ptr(p1), dx, dy, invd, tol, maxiter, ptr(part),
                    ptr(S.scal), sync.ctypes.data_as(C.POINTER(C.c_ulonglong)))
    assert sync[3] == 0, "a bounded wait inside k_cg_tmad timed out"
    return dev.get(x)


@pytest.mark.parametrize("ny,nx,tile_rows,unit,jacobi,maxiter", [
    (6, 1024, 4, True, True, 5),           # two tiles of 4 and 2 rows, one column chunk
    (9, 1024, 4, False, True, 5),          # three tiles, non-unit spacing
    (5, 2048, 4, True, False, 4),          # two column chunks x two bands = four tiles, no preconditioner
    (33, 1024, 16, True, True, 2),         # the library's default 16-row tiles: 16 + 16 + 1 rows
    (3, 1024, 4, True, True, 300),         # runs to convergence at a loose tolerance
    (13, 1024, (4, 1, 2), True, True, 4),  # two 1-row tiles at each end, 4 + 4 + 1 rows between them (both sweeps end on short tiles)
    (40, 2048, (16, 4, 2), False, True, 2),   # 2 chunks: 2 x 4-row bands, 16 + 8 rows, 2 x 4-row bands: lean interior tiles
    (36, 4096, (2, 1, 2), True, True, 2),  # 4 chunks x (2 + 16 + 2) bands = 80 tiles: two fold groups (64 + 16)
    (9, 1100, 4, True, True, 3),           # clipped last chunk: 76 columns = one full strip + a strip cut at 12 columns, 14 idle warps
    (11, 1030, 4, False, True, 3),         # last chunk of 6 columns (one cut strip), non-unit spacing
    (22, 2049, 4, True, True, 2),          # ODD width: the last chunk is ONE column (copied together with the padding column)
    (10, 1089, 4, True, True, 3),          # 1024 + 65: the strip whose east halo column is the wall, then a one-column strip
])
def test_tma_fed_cg_kernel_vs_oracle(emu, ny, nx, tile_rows, unit, jacobi, maxiter):
    """k_cg_tmad<unit, single GPU> on one CTA: producer thread, 8-stage ring, ticket-drawn tiles (ascending in phase
    A, descending in phase B), per-tile partials -- against scipy's recurrence; two solves in a row, so the ticket
    base persisted by the first launch is used by the second."""
    if not emu.emu_thread_probe(544):
        pytest.skip("this machine cannot run 544 OS threads at once")
    rng = np.random.default_rng(ny * 1021 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jacobi else 1.0
    tol = 1e-10 if maxiter < 100 else 0.9
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    sync = np.zeros(256, dtype=np.uint64)          # 2 KiB sync block (fill_sync layout)
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    for _repeat in range(2):
        b = rng.standard_normal((ny, nx))
        b -= b.mean()
        b[0, 0] = 0.0
        x = _tmad_solve(emu, dev, S, sync, b, dx, dy, invd, tol, maxiter, tile_rows)
        xo, io, to = orc.pcg(A, b.ravel().copy(), tol, maxiter, inv)
        assert int(S["S_INFO"]) == io and abs(int(S["S_ITERS"]) - to) <= 1
        assert np.linalg.norm(x.ravel() - xo) <= (1e-10 if int(S["S_ITERS"]) == to else 1e-6) * np.linalg.norm(xo)
    assert int(sync[5]) == int(sync[4]) > 0                          # every ticket drawn was accounted for
    zero = _tmad_solve(emu, dev, S, sync, np.zeros((ny, nx)), dx, dy, invd, tol, 10, tile_rows)
    assert int(S["S_INFO"]) == 0 and int(S["S_ITERS"]) == 0 and not zero.any()


def test_whole_projection_steps_with_the_tma_fed_kernel_vs_oracle(emu):
    """A 1024 x 6 driven cavity, three steps in the order of pf_step with k_cg_tmad for the pressure solve."""
    if not emu.emu_thread_probe(544):
        pytest.skip("this machine cannot run 544 OS threads at once")
    from types import SimpleNamespace
    nx, ny = 1024, 6
    cfg = SimpleNamespace(nx=nx, ny=ny, Re=1000.0, lid_velocity=1.0, cfl_target=0.5, cfl_growth=1.05, lin_tol=1e-10,
                          lin_maxiter=6, advection_scheme="quick", disable_advection=False, diagnostics=False)
    ost = orc.allocate_state(nx, ny)
    dev, S = Dev(emu, ny, nx), Scalars(emu)
    sync = np.zeros(256, dtype=np.uint64)          # 2 KiB sync block (fill_sync layout)
    u, v, uo, vo = dev.put(), dev.put(), dev.put(), dev.put()
    b, x, r, p0, p1 = dev.put(), dev.put(), dev.put(), dev.put(), dev.put()
    part = np.zeros(emu.emu_cg_part_doubles())
    center = np.zeros(ny)
    for it in range(3):
        _, ores, od = orc.step(cfg, ost, it)
        dt = od["dt"]
        emu.emu_predict(QUICK, nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), dt, 1.0, 1.0)
        u, uo, v, vo = uo, u, vo, v
        emu.emu_rhs1(nx, ny, ptr(u), ptr(v), ptr(b), max(dt, 1e-14), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_rhs2(nx, ny, ptr(b), ptr(x), ptr(r), 1, 1, 0.25, ptr(S.scratch), ptr(S.scal))
        emu.emu_cg_tmad(nx, ny, 4, 0, 0, ptr(b), ptr(x), ptr(r), ptr(p0), ptr(p1), 1.0, 1.0, 0.25, 1e-10, 6, ptr(part),
                        ptr(S.scal), sync.ctypes.data_as(C.POINTER(C.c_ulonglong)))
        assert int(S["S_ITERS"]) == od["true_iterations"] and sync[3] == 0
        emu.emu_resid(nx, ny, ptr(b), ptr(x), 1.0, 1.0, ptr(S.scratch), ptr(S.scal))
        emu.emu_correct(nx, ny, ptr(x), ptr(u), ptr(v), dt, 1.0, 1.0, S["S_X0"])
        emu.emu_lid(nx, ny, ptr(u), 1.0)
        emu.emu_post(nx, ny, ptr(u), ptr(v), ptr(uo), ptr(vo), ptr(x), ptr(center), ptr(S.scratch), ptr(S.scal))
        assert close(S["S_DIVC_SUMSQ"] ** 0.5 / (nx * ny), ores["continuity"], 1e-8)
    for name, buf in (("u", u), ("v", v), ("p", x)):
        a, want = dev.get(buf), orc.interior(ost.fields[name])
        assert np.linalg.norm(a - want) <= 1e-9 * (np.linalg.norm(want) or 1.0), name


# ===================================================================================================================
# the MULTI-GPU instantiation of the TMA-fed kernel: several ranks at once, peer halo stores + mailbox all-reduce
# ===================================================================================================================
class Ranks:
    """Row slabs of a global grid, one per rank, in the device layout, plus the per-rank scalar / partial / sync blocks."""

    def __init__(self, emu, R, ny, nx):
        self.emu, self.R, self.ny, self.nx, self.nyl = emu, R, ny, nx, ny // R
        self.pitch, self.H = emu.emu_pitch(nx), emu.emu_halo()
        self.parts = [np.zeros(emu.emu_cg_part_doubles()) for _ in range(R)]
        self.scals = [np.zeros(emu.emu_scal_count()) for _ in range(R)]
        self.syncs = [np.zeros(2048, dtype=np.uint8) for _ in range(R)]          # persist between solves (epochs, tickets)

    def slabs(self, full=None, halo=None):
        out = []
        for k in range(self.R):
            buf = np.full((self.nyl + 2 * self.H, self.pitch), np.nan)
            buf[self.H:self.H + self.nyl, :self.nx] = 0.0 if full is None else full[k * self.nyl:(k + 1) * self.nyl]
            if halo is not None:                       # one halo row from each existing neighbour
                if k > 0:
                    buf[self.H - 1, :self.nx] = halo[k * self.nyl - 1]
                if k < self.R - 1:
                    buf[self.H + self.nyl, :self.nx] = halo[(k + 1) * self.nyl]
            out.append(buf)
        return out

    def gather(self, bufs):
        for k, buf in enumerate(bufs):                 # only neighbours' halo rows may have been written
            lo = self.H - 1 if k > 0 else self.H
            hi = self.H + self.nyl + (1 if k < self.R - 1 else 0)
            assert np.isnan(buf[:lo]).all() and np.isnan(buf[hi:]).all() and np.isnan(buf[:, self.nx:]).all()
        return np.concatenate([buf[self.H:self.H + self.nyl, :self.nx] for buf in bufs])

    def arr(self, bufs):
        return (C.POINTER(C.c_double) * self.R)(*[ptr(a) for a in bufs])

    def solve(self, b, dx, dy, jacobi, tol, maxiter, tile_rows, kernel="tmad"):
        emu, R = self.emu, self.R
        invd = 1.0 / (2.0 / dx ** 2 + 2.0 / dy ** 2) if jacobi else 1.0
        M = orc.jacobi_inverse_diagonal(self.nx, self.ny, dx, dy).reshape(b.shape) if jacobi else np.ones_like(b)
        for sc in self.scals:                          # the host sequence all-reduces these before the launch
            sc[emu.emu_slot(b"S_BB")] = float(np.sum(b * b))
            sc[emu.emu_slot(b"S_BZ")] = float(np.sum(b * (M * b)))
        bs, xs, rs = self.slabs(b), self.slabs(), self.slabs(b, halo=b)
        zeros = np.zeros_like(b)
        p0, p1 = self.slabs(zeros, halo=zeros), self.slabs(zeros, halo=zeros)
        sy = (C.POINTER(C.c_ubyte) * R)(*[s_.ctypes.data_as(C.POINTER(C.c_ubyte)) for s_ in self.syncs])
        launch = emu.emu_cg_tmad_multi if kernel == "tmad" else emu.emu_cg_persistent_multi
        launch(R, self.nx, self.ny, *_tiles(tile_rows), self.arr(bs), self.arr(xs), self.arr(rs), self.arr(p0),
               self.arr(p1), dx, dy, invd, tol, maxiter, self.arr(self.parts), self.arr(self.scals), sy)
        assert all(int(s_.view(np.uint64)[3]) == 0 for s_ in self.syncs), "a bounded wait timed out"
        info = {int(sc[emu.emu_slot(b"S_INFO")]) for sc in self.scals}
        iters = {int(sc[emu.emu_slot(b"S_ITERS")]) for sc in self.scals}
        assert len(info) == 1 and len(iters) == 1, "ranks disagree on the stopping decision"
        return self.gather(xs), info.pop(), iters.pop()


@pytest.mark.parametrize("R,ny,nx,tile_rows,unit,jacobi,maxiter", [
    (2, 8, 1024, 4, True, True, 4), (2, 12, 1024, 4, False, True, 5), (3, 9, 1024, 4, True, False, 4),
    (4, 8, 2048, 2, True, True, 3),          # four ranks, two column chunks, two-row slabs
    (2, 4, 1024, 4, True, True, 300),        # runs to convergence: every rank takes the same stop decision
    (2, 28, 1024, (4, 1, 2), True, True, 3),   # per rank: two 1-row tiles at each end of the 14-row slab, 4 + 4 + 2 rows between
    (2, 12, 1100, 4, True, True, 3),           # clipped last chunk across two ranks (peer rows of a cut strip)
    (3, 12, 1025, 2, False, True, 3),          # odd width, three ranks
])
def test_multi_gpu_tma_kernel_with_concurrent_ranks_vs_single_domain_oracle(emu, R, ny, nx, tile_rows, unit, jacobi,
                                                                            maxiter):
    """k_cg_tmad<unit, MULTI> -- compute and collective in one kernel -- with R ranks running at once: the result on
    the slab-decomposed grid equals scipy's recurrence on the whole grid; all ranks hold the same iteration count;
    a second solve reuses the mailboxes with the persisted epochs."""
    if not emu.emu_thread_probe(544 * R):
        pytest.skip("this machine cannot run 544 OS threads per rank at once")
    rng = np.random.default_rng(R * 7919 + ny * 31 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    tol = 1e-10 if maxiter < 100 else 0.9
    ranks = Ranks(emu, R, ny, nx)
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    for _repeat in range(2):
        b = rng.standard_normal((ny, nx))
        b -= b.mean()
        b[0, 0] = 0.0
        x, info, iters = ranks.solve(b, dx, dy, jacobi, tol, maxiter, tile_rows)
        xo, io, to = orc.pcg(A, b.ravel().copy(), tol, maxiter, inv)
        assert info == io and abs(iters - to) <= 1
        assert np.linalg.norm(x.ravel() - xo) <= (1e-10 if iters == to else 1e-6) * np.linalg.norm(xo)
    epochs = {int(s_.view(np.uint64)[2]) for s_ in ranks.syncs}
    assert len(epochs) == 1 and epochs.pop() > 0                       # same number of exchanges on every rank


@pytest.mark.parametrize("R,ny,nx,unit,jacobi,maxiter", [
    (2, 16, 70, True, True, 5), (2, 10, 17, False, True, 6), (3, 12, 130, True, False, 5), (4, 8, 200, True, True, 4),
    (2, 4, 6, True, True, 300),              # runs to convergence
])
def test_multi_gpu_register_window_kernel_with_concurrent_ranks_vs_single_domain_oracle(emu, R, ny, nx, unit, jacobi,
                                                                                        maxiter):
    """k_cg_persistent_multi -- the peer-memory path for widths that are not a multiple of 1024 -- with R ranks running
    at once: same checks as for the TMA-fed kernel above."""
    if not emu.emu_thread_probe(256 * R):
        pytest.skip("this machine cannot run 256 OS threads per rank at once")
    rng = np.random.default_rng(R * 7927 + ny * 31 + nx)
    dx, dy = (1.0, 1.0) if unit else (0.5, 0.25)
    ranks = Ranks(emu, R, ny, nx)
    A = lambda z: orc.neg_laplacian_pinned(z, nx, ny, dx, dy)        # noqa: E731
    inv = orc.jacobi_inverse_diagonal(nx, ny, dx, dy) if jacobi else None
    for _repeat in range(2):
        b = rng.standard_normal((ny, nx))
        b -= b.mean()
        b[0, 0] = 0.0
        x, info, iters = ranks.solve(b, dx, dy, jacobi, 1e-10, maxiter, 4, kernel="register-window")
        xo, io, to = orc.pcg(A, b.ravel().copy(), 1e-10, maxiter, inv)
        assert info == io and abs(iters - to) <= 1
        assert np.linalg.norm(x.ravel() - xo) <= (1e-10 if info else 1e-9) * np.linalg.norm(xo)
    epochs = {int(s_.view(np.uint64)[2]) for s_ in ranks.syncs}
    assert len(epochs) == 1 and epochs.pop() > 0
