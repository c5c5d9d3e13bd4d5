"""Byte-compile the UNMODIFIED reference into oracle/_ref/ (TEST / MEASUREMENT INFRASTRUCTURE).

The reference is pure Python (src/pyflow/**), so "building" it means compiling every module to CPython bytecode:
``oracle/_ref/pyflow/<module>.pfbc`` -- the content of a sourceless ``.pyc`` under another extension, because
``*.pyc`` files are stripped from the snapshot that travels to the GPU box (first B200 run of this leg: the directory
arrived empty).  ``install_importer()`` registers a path hook that loads that extension with CPython's own
``SourcelessFileLoader``.  No reference SOURCE is copied anywhere; ``oracle/_ref/`` is git-ignored (it travels to the
GPU box with the snapshot, like the repo's own built ``.so``), and only ``bench.py``'s CPU legs execute it -- in a
subprocess, because the reference package and the drop-in package are both called ``pyflow``.

Runs only where ``/root/reference`` exists (the build container).  Same interpreter on the GPU box (same image), so
the bytecode's magic number matches there; ``stamp.json`` records it and ``time_unmodified.py`` refuses a mismatch.

Usage:  python oracle/build_ref.py            (called by __graft_entry__.build())
"""
from __future__ import annotations

import importlib.util
import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
SUFFIX = ".pfbc"            # CPython bytecode of one module (a .pyc by content)


def install_importer() -> None:
    """Make ``oracle/_ref`` importable: bytecode files with the SUFFIX extension, found like ordinary modules /
    packages (``__init__`` + SUFFIX), loaded by importlib's SourcelessFileLoader."""
    import importlib.machinery as mach
    hook = mach.FileFinder.path_hook((mach.SourcelessFileLoader, [SUFFIX]))

    def ref_only(path):
        ap = os.path.abspath(path)
        if ap == OUT or ap.startswith(OUT + os.sep):
            return hook(path)
        raise ImportError
    sys.path_hooks.insert(0, ref_only)
    sys.path_importer_cache.clear()
    sys.path.insert(0, OUT)


def build(force: bool = False) -> str | None:
    """Returns the output directory, or None where the reference is absent (the prebuilt directory is used)."""
    src_pkg = os.path.join(REF_SRC, "pyflow")
    if not os.path.isdir(src_pkg):
        return OUT if os.path.exists(os.path.join(OUT, "pyflow", "__init__" + SUFFIX)) else None
    stamp_path = os.path.join(OUT, "stamp.json")
    sources = []
    for d, _dirs, files in os.walk(src_pkg):
        for f in sorted(files):
            if f.endswith(".py"):
                sources.append(os.path.join(d, f))
    sources.sort()
    newest = max(os.path.getmtime(s) for s in sources)
    magic = importlib.util.MAGIC_NUMBER.hex()
    if not force and os.path.exists(stamp_path):
        try:
            st = json.load(open(stamp_path))
            if (st.get("magic") == magic and st.get("newest_source_mtime") == newest and st.get("modules") == len(sources)
                    and st.get("suffix") == SUFFIX):
                return OUT
        except Exception:
            pass
    shutil.rmtree(os.path.join(OUT, "pyflow"), ignore_errors=True)
    for s in sources:
        rel = os.path.relpath(s, REF_SRC)
        dst = os.path.join(OUT, rel[:-3] + SUFFIX)               # module.py -> module.pfbc where the module would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path tracebacks name; unchecked-hash pyc so that no source is looked for on the box
        py_compile.compile(s, cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(stamp_path, "w") as fh:
        json.dump({"magic": magic, "python": sys.version.split()[0], "modules": len(sources), "suffix": SUFFIX,
                   "newest_source_mtime": newest, "from": src_pkg,
                   "what": "CPython bytecode of the unmodified reference package, compiled by oracle/build_ref.py"}, fh, indent=1)
    return OUT


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(f"[build_ref] {out or 'reference absent and no prebuilt oracle/_ref'}")
