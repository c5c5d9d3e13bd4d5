#!/usr/bin/env bash
# profiles/sass_summary.sh -- instruction census of the CG kernels in the built library (cuobjdump -sass):
# proves the TMA bulk copies (UBLKCP), mbarriers (SYNCS), 128-bit shared loads / global stores, the absence of fused
# multiply-adds in the unit-spacing kernels (-fmad=false; DFMA only from the scalar divide / sqrt sequences), and
# which fences the exchanges use.  Usage: bash profiles/sass_summary.sh > profiles/r2b_sass_summary.txt
LIB=pyflow-solver_b200/pyflow/_b200/libpyflow_b200.so
echo "library: $LIB  sha256 $(sha256sum $LIB | cut -c1-16)  source hash $(python -c 'import bench; print(bench.kernel_source_hash())')"
cuobjdump -sass $LIB 2>/dev/null | awk '
  /Function : / { fn=$3; next }
  fn ~ /k_cg_tmad|k_cg_small|k_cg_persistent|k_cg_tma[^d]|k_cg_generic|k_bicgstab|k_front|k_finish/ {
    n[fn]++
    if ($0 ~ /UBLKCP/) ublkcp[fn]++
    if ($0 ~ /UTMALDG/) utmaldg[fn]++
    if ($0 ~ /SYNCS/) syncs[fn]++
    if ($0 ~ /LDS\.128/) lds128[fn]++
    if ($0 ~ /STG\.E\.128/) stg128[fn]++
    if ($0 ~ /LDG\.E\.128/) ldg128[fn]++
    if ($0 ~ /DFMA/) dfma[fn]++
    if ($0 ~ /DADD/) dadd[fn]++
    if ($0 ~ /DMUL/) dmul[fn]++
    if ($0 ~ /MEMBAR\.SC/) msc[fn]++
    if ($0 ~ /MEMBAR\.ALL/) mall[fn]++
    if ($0 ~ /ATOM|RED\./) atom[fn]++
    if ($0 ~ /STL|LDL/) local[fn]++
  }
  END {
    printf "%-64s %7s %6s %7s %5s %7s %7s %7s %5s %5s %5s %6s %6s %5s %5s\n", "kernel", "lines", "UBLKCP", "UTMALDG", "SYNCS", "LDS.128", "STG.128", "LDG.128", "DFMA", "DADD", "DMUL", "MB.SC", "MB.ALL", "ATOM", "LOCAL"
    for (f in n) printf "%-64s %7d %6d %7d %5d %7d %7d %7d %5d %5d %5d %6d %6d %5d %5d\n", substr(f,1,64), n[f], ublkcp[f], utmaldg[f], syncs[f], lds128[f], stg128[f], ldg128[f], dfma[f], dadd[f], dmul[f], msc[f], mall[f], atom[f], local[f]
  }' | sort
