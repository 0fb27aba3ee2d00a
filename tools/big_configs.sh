#!/bin/bash
# BASELINE configs[2] (10 M records of config 3) and the per-GPU share of configs[4] (12.5 M of config 5) on one GPU:
# bench line + one full ncu capture each.
mkdir -p gpurun_out
for spec in "config3 10000000" "config5 12500000"; do
  set -- $spec
  timeout 600 python bench.py --config $1 --records $2 --steps 30 --warmup 3 --e2e-steps 4 --no-cpu-baseline 2>gpurun_out/bench_$1.err | tail -1 > gpurun_out/bench_$1.json
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:regk_ -s 6 -c 2 -f -o gpurun_out/prof_$1 \
      python bench.py --config $1 --records $2 --steps 6 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
  python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); print('$1', d['value']/1e9, d['ms_per_step'], d['roofline_kernels']['path']['frac'], d['roofline_kernels']['json']['frac'], d['e2e']['value']/1e6)"
done
