#!/usr/bin/env bash
# profiles/capture.sh -- the commands the files in profiles/ come from (run on a B200 box from the
# repo root, e.g. `gpurun --timeout 900 -- 'bash profiles/capture.sh launches full4096'`).
# Outputs go to gpurun_out/; copy the summaries you want kept into profiles/.
# Numbers printed by a run under ncu / compute-sanitizer are never bench values.
set -u
mkdir -p gpurun_out
for what in "$@"; do
  case "$what" in
    tests)        # parity suite
      python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/gpu_tests.log | tail -5 ;;
    newtests)     # the -m gpu tests written after round 1's GPU minutes were spent: first run on a B200, verbose
      python -m pytest tests/test_zz_gpu_more_parity.py -m gpu -v -k "not legacy" 2>&1 | tee gpurun_out/gpu_new_tests.log | tail -40 ;;
    bench)        # headline line + config 2 alone
      python bench.py > gpurun_out/bench_cavity4096.json 2> gpurun_out/bench_cavity4096.err
      python bench.py --workload cavity512 --steps 50 --no-cpu-baseline > gpurun_out/bench_cavity512.json ;;
    launches)     # launch list of the bench command (kernel shares of a step)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/launches_bench.log 2>&1 ;;
    full4096)     # ncu --set full of the TMA-fed CG kernel (4096^2): DRAM bytes -> profiles/cg_traffic.json
      ncu --set full --clock-control none --import-source on -k regex:k_cg_tmad -s 2 -c 1 \
          -o gpurun_out/cg_tmad_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
          > gpurun_out/cg_tmad_full.log 2>&1
      python profiles/ncu_summary.py gpurun_out/cg_tmad_full.ncu-rep > gpurun_out/cg_tmad_full_summary.txt
      # DRAM bytes of the launch + kernel name + source hash -> profiles/cg_traffic.json (a copy travels back)
      python profiles/record_traffic.py cavity4096 gpurun_out/cg_tmad_full.ncu-rep && cp profiles/cg_traffic.json gpurun_out/ ;;
    full512)      # the shared-memory-resident CG kernel (512^2): what bounds 5.6 us per iteration
      ncu --set full --clock-control none --import-source on -k regex:k_cg_small -s 2 -c 1 \
          -o gpurun_out/cg_small_full -f python bench.py --workload cavity512 --steps 1 --warmup 3 --no-cpu-baseline \
          > gpurun_out/cg_small_full.log 2>&1
      python profiles/ncu_summary.py gpurun_out/cg_small_full.ncu-rep > gpurun_out/cg_small_full_summary.txt
      python profiles/record_traffic.py cavity512 gpurun_out/cg_small_full.ncu-rep && cp profiles/cg_traffic.json gpurun_out/ ;;
    sanitize)     # memcheck + racecheck + synccheck + initcheck of every CG kernel variant on small grids
      for tool in memcheck racecheck synccheck initcheck; do
        timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tests/sanitize_cases.py \
          > gpurun_out/sanitizer_$tool.log 2>&1
        echo "$tool rc=$? $(grep -c SANITIZE gpurun_out/sanitizer_$tool.log) lines, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)" | tee -a gpurun_out/sanitizer_summary.txt
      done ;;
    sanitize_steps)   # the per-step passes (fused and separate) under memcheck / synccheck / initcheck
      for tool in memcheck synccheck initcheck; do
        timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tests/sanitize_cases.py steps \
          > gpurun_out/sanitizer_steps_$tool.log 2>&1
        echo "steps $tool rc=$? $(grep -c SANITIZE gpurun_out/sanitizer_steps_$tool.log) lines, $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_steps_$tool.log | tail -1)" | tee -a gpurun_out/sanitizer_steps_summary.txt
      done ;;
    scaling)      # configs 4 and 5 at every GPU count the box has (`gpurun --gpus 8 -- 'bash profiles/capture.sh scaling'`)
      ngpu=$(nvidia-smi -L | wc -l)
      for wl in cavity4096 strong8192 weak16384; do
        for n in 1 2 4 8; do
          [ "$n" -le "$ngpu" ] || continue
          steps=10; [ "$wl" = weak16384 ] && steps=5
          timeout 900 python bench.py --gpus $n --workload $wl --steps $steps --warmup 3 --no-cpu-baseline \
            > gpurun_out/bench_${wl}_${n}gpu.json 2> gpurun_out/bench_${wl}_${n}gpu.err
          echo "$wl x$n rc=$?" | tee -a gpurun_out/scaling_summary.txt
        done
      done ;;
    reference)    # the CPU arm on the box's host cores
      python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json ;;
    timeline)     # where the time of one k_cg_small iteration goes; the phase-end exchange across 2 GPUs
      python profiles/experiments/small_kernel_timeline.py 512 | tee gpurun_out/small_kernel_timeline.txt
      [ "$(nvidia-smi -L | wc -l)" -ge 2 ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29701 profiles/experiments/barrier_timeline_multigpu.py 4096 4096 2>&1 | grep -E "rank|barrier" | tee gpurun_out/barrier_timeline_2gpu.txt ;;
    steppasses)   # the two fused per-step passes against the seven separate kernels: A/B of the non-CG part of a step, then
                  # per-kernel times of both chains (ncu launch list; short solves so that the list stays small)
      python profiles/experiments/step_passes_ab.py 2>&1 | grep STEPPASS | tee gpurun_out/step_passes_ab.txt
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/step_passes_launches.csv \
          -k 'regex:k_front|k_finish|k_rhs|k_predict|k_resid|k_correct|k_divnorm|k_lid|k_post' \
          python profiles/experiments/step_passes_ab.py 4096 4096 5 2 512 512 5 2 > gpurun_out/step_passes_ncu.log 2>&1 ;;
    *) echo "unknown target $what (tests newtests bench launches full4096 full512 sanitize sanitize_steps scaling reference timeline steppasses)"; exit 2 ;;
  esac
done
