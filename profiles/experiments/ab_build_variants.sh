# A/B of build variants on one box (run under gpurun): prints ms/step, CG ms/launch, roofline fraction, SM clock
run() { python bench.py --no-extra --no-cpu-baseline --no-parity --steps 10 $1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['ms_per_step'], d['roofline']['avg_launch_ms'], round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; }
for flags in "" "-DPF_TMAD_MAXNREG" $EXTRA_VARIANTS; do
  PF_NVCC_FLAGS="$flags" python pyflow-solver_b200/build.py --force > /dev/null
  run "" "[$flags] 4096"; run "--workload weak16384 --steps 3" "[$flags] 16384"
done
python pyflow-solver_b200/build.py --force > /dev/null
