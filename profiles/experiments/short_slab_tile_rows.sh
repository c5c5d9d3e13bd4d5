# tile height on the short slab of the 8-GPU strong-scaling run (8192 x 1024 per GPU), measured on ONE GPU
for tr in 4 6 8 12 16 24; do
  echo -n "slab8192x1024 tile_rows=$tr: "
  PYFLOW_B200_OPTIONS=cg_tile_rows=$tr python bench.py --workload slab8192x1024 --no-extra --no-cpu-baseline --no-parity --steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
