# A/B of build variants of the fused per-step passes on one box (run under gpurun from the repo root).  The variant
# libraries are built beforehand (CPU, nvcc cross-compiles):
#   PF_NVCC_FLAGS="-DPF_FUSED_UNROLL=2" python pyflow-solver_b200/build.py --out=pyflow-solver_b200/pyflow/_b200/variants/libpf_u2.so
# and selected through PYFLOW_B200_LIB.  Per variant: the non-CG part of a step (CUDA events of pf_step) on three grids
# and the per-kernel times of the two passes (ncu launch list, 4096^2).  Then one `ncu --set full` capture of each pass
# of the default library.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_step_passes.py -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2b_step_passes_tests_v2.log
V=pyflow-solver_b200/pyflow/_b200/variants
for lib in default $(ls $V/*.so 2>/dev/null); do
  echo "== variant $lib" | tee -a gpurun_out/step_passes_variants2.txt
  if [ "$lib" = default ]; then unset PYFLOW_B200_LIB; else export PYFLOW_B200_LIB=$PWD/$lib; fi
  python profiles/experiments/step_passes_ab.py 4096 4096 300 8 512 512 300 30 2>&1 | grep "STEPPASS.*fused=1" | tee -a gpurun_out/step_passes_variants2.txt
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/variant2_launches_$(basename $lib .so).csv \
      -k 'regex:k_front|k_finish' python profiles/experiments/step_passes_ab.py 4096 4096 5 2 > /dev/null 2>&1
done
unset PYFLOW_B200_LIB
for k in k_front k_finish; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/${k}_full -f \
      python profiles/experiments/step_passes_ab.py 4096 4096 5 2 > gpurun_out/${k}_full.log 2>&1
  python profiles/ncu_summary.py gpurun_out/${k}_full.ncu-rep > gpurun_out/${k}_full_summary.txt 2>&1
done
