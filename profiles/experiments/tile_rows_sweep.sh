for wl in cavity4096 strong8192; do
for tr in 12 16 20 24 28 32 40 48; do
  echo -n "$wl tile_rows=$tr: "
  PYFLOW_B200_OPTIONS=cg_tile_rows=$tr python bench.py --workload $wl --no-extra --no-cpu-baseline --no-parity --steps 12 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done; done
for tr in 16 32 48 64; do
  echo -n "weak16384 tile_rows=$tr: "
  PYFLOW_B200_OPTIONS=cg_tile_rows=$tr python bench.py --workload weak16384 --no-extra --no-cpu-baseline --no-parity --steps 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['avg_launch_ms'], round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
