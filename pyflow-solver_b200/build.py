"""Build libpyflow_b200.so in-tree with nvcc for sm_100a.

    python pyflow-solver_b200/build.py [--force] [--verbose]

The shared library is written next to the Python binding
(``pyflow-solver_b200/pyflow/_b200/libpyflow_b200.so``); it is git-ignored but
travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a
GPU.  ``-fmad=false`` is part of the parity contract (numpy never fuses a
multiply-add), ``-lineinfo`` lets ncu map stalls to source lines.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "pyflow", "_b200")
LIB_PATH = os.path.join(OUT_DIR, "libpyflow_b200.so")
SOURCES = ["pf_api.cu", "pf_comm.cu"]


def _nccl_dirs():
    """Headers + libnccl.so.2 bundled with torch's nvidia-nccl wheel."""
    site = sysconfig.get_paths()["purelib"]
    inc = os.path.join(site, "nvidia", "nccl", "include")
    lib = os.path.join(site, "nvidia", "nccl", "lib")
    if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
        return inc, lib
    return None, None


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _deps():
    files = [os.path.join(ROOT, "include", "pyflow_b200.h"), os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        files.append(os.path.join(CSRC, f))
    return files


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def build_native(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    if out is None and not force and not needs_build():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    inc, lib = _nccl_dirs()
    if inc is None:
        raise RuntimeError("NCCL headers/lib of the torch-bundled nvidia-nccl wheel not found")
    cmd = [
        _nvcc(), "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-fmad=false",
        "-Xcompiler", "-fPIC", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", inc,
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("PF_NVCC_FLAGS", "").split()      # tuning experiments, e.g. -DPF_CG_UA=6
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", out or LIB_PATH, "-L", lib, "-l:libnccl.so.2", "-Xlinker", f"-rpath={lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libpyflow_b200.so")
    return out or LIB_PATH


if __name__ == "__main__":
    out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), None)
    path = build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out)
    print(path)
