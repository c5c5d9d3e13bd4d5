set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_step_passes.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2b_step_passes_tests_v1.log
V=pyflow-solver_b200/pyflow/_b200/variants
for lib in default $V/libpf_v0.so $V/libpf_v2.so; do
  echo "== variant $lib" | tee -a gpurun_out/step_passes_variants.txt
  if [ "$lib" = default ]; then unset PYFLOW_B200_LIB; else export PYFLOW_B200_LIB=$PWD/$lib; fi
  python profiles/experiments/step_passes_ab.py 4096 4096 300 8 512 512 300 30 16384 4096 20 3 2>&1 | grep STEPPASS | tee -a gpurun_out/step_passes_variants.txt
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/variant_launches_$(basename $lib .so).csv \
      -k 'regex:k_front|k_finish' python profiles/experiments/step_passes_ab.py 4096 4096 5 2 > /dev/null 2>&1
done
unset PYFLOW_B200_LIB
