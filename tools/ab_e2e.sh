for depth in 2 1; do
  echo -n "depth=$depth  "
  REGK_E2E_DEPTH=$depth timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['e2e']['value']/1e6,1), 'M rec/s', round(d['e2e']['ms_per_step'],3),'ms')"
done
