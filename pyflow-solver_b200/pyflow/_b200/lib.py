"""ctypes binding of ``libpyflow_b200.so`` (C-ABI: ``include/pyflow_b200.h``).

There is NO CPU fallback: if the shared library is missing or cannot be
loaded, every entry point raises ``NativeLibraryError`` telling the user to
build it (``python pyflow-solver_b200/build.py``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYFLOW_B200_LIB") or os.path.join(_HERE, "libpyflow_b200.so")   # override: tuning builds

PF_ADV_UPWIND, PF_ADV_QUICK, PF_ADV_NONE = 0, 1, 2
ABI_VERSION = 1

# every symbol include/pyflow_b200.h declares (checked by the CPU test-suite)
EXPORTS = (
    "pf_abi_version", "pf_strerror", "pf_last_error", "pf_workspace_bytes", "pf_create", "pf_destroy",
    "pf_upload", "pf_download", "pf_prestep_reduce", "pf_step", "pf_centerline", "pf_divergence_stats",
    "pf_op_gradient_coherent", "pf_op_divergence_coherent", "pf_op_laplacian_coherent",
    "pf_op_divergence_central", "pf_op_gradient_central", "pf_op_laplacian_central",
    "pf_op_pressure_correction_central", "pf_op_advect", "pf_op_pcg", "pf_op_cg_generic", "pf_op_bicgstab", "pf_op_krylov_diag", "pf_op_krylov_stencil5",
    "pf_comm_unique_id", "pf_comm_init", "pf_comm_set_peers", "pf_device_info", "pf_launch_count", "pf_set_option", "pf_debug_read_partials",
)


class NativeLibraryError(RuntimeError):
    pass


class StepParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("lid_velocity", C.c_double), ("lin_tol", C.c_double), ("inv_diag_interior", C.c_double),
        ("lid_set", C.c_int32), ("scheme", C.c_int32), ("jacobi", C.c_int32),
        ("lin_maxiter", C.c_int32), ("projection_only", C.c_int32), ("reserved", C.c_int32),
    ]


class StepOut(C.Structure):
    _fields_ = [
        ("div_before_norm", C.c_double), ("rhs_mean", C.c_double), ("bnorm", C.c_double),
        ("rnorm", C.c_double), ("div_after_norm", C.c_double),
        ("ru_sumsq", C.c_double), ("rv_sumsq", C.c_double), ("cont_sumsq", C.c_double),
        ("umax_raw", C.c_double), ("vmax_raw", C.c_double), ("p00", C.c_double),
        ("cg_ms", C.c_double), ("step_ms", C.c_double),
        ("info", C.c_int32), ("true_iters", C.c_int32), ("nonfinite", C.c_int32), ("reserved", C.c_int32),
    ]


_lib = None


def _declare(lib):
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.pf_abi_version.restype = C.c_int
    lib.pf_strerror.restype = C.c_char_p
    lib.pf_strerror.argtypes = [C.c_int]
    lib.pf_last_error.restype = C.c_char_p
    lib.pf_last_error.argtypes = [vp]
    lib.pf_workspace_bytes.restype = C.c_size_t
    lib.pf_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.pf_create.argtypes = [C.POINTER(vp), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, vp, C.c_size_t]
    lib.pf_destroy.argtypes = [vp]
    lib.pf_upload.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
    lib.pf_download.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
    lib.pf_prestep_reduce.argtypes = [vp, dp, vp]
    lib.pf_step.argtypes = [vp, C.POINTER(StepParams), C.POINTER(StepOut), vp]
    lib.pf_centerline.argtypes = [vp, vp, vp]
    lib.pf_divergence_stats.argtypes = [vp, C.c_double, C.c_double, dp, vp]
    lib.pf_op_gradient_coherent.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_divergence_coherent.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_laplacian_coherent.argtypes = [vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_divergence_central.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_gradient_central.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_laplacian_central.argtypes = [vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_pressure_correction_central.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, vp, vp]
    lib.pf_op_advect.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_double, C.c_double, vp]
    lib.pf_op_pcg.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                              C.POINTER(C.c_int), C.POINTER(C.c_int), dp, vp]
    lib.pf_op_bicgstab.argtypes = lib.pf_op_pcg.argtypes
    lib.pf_op_krylov_diag.argtypes = [vp, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), dp, vp]
    lib.pf_op_krylov_stencil5.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, vp]
    lib.pf_op_cg_generic.argtypes = [vp, C.c_int, vp, vp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp, vp]
    lib.pf_comm_unique_id.argtypes = [vp]
    lib.pf_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.pf_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pf_comm_set_peers.argtypes = [vp, C.POINTER(vp), C.c_int]
    lib.pf_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    lib.pf_debug_read_partials.argtypes = [vp, dp, C.c_int, C.c_int]
    lib.pf_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    for name in EXPORTS:
        getattr(lib, name)       # AttributeError here == header/library drift
    return lib


def load():
    """Load (once) and return the CDLL.  Raises NativeLibraryError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run "
            "`python pyflow-solver_b200/build.py` (needs nvcc). There is no CPU fallback.")
    try:
        # torch ships the NCCL the library links against; importing it first
        # guarantees libnccl.so.2 / libcudart are resolvable in this process.
        import torch  # noqa: F401
    except Exception:  # pragma: no cover - torch is part of the image
        pass
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    _declare(lib)
    if lib.pf_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI mismatch: library {lib.pf_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, ctx=None, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.pf_strerror(rc).decode()
    detail = lib.pf_last_error(ctx).decode() if ctx else ""
    raise RuntimeError(f"libpyflow_b200 {what} failed ({rc}: {msg}) {detail}".strip())
