#!/bin/bash
# Round-2 evidence, run on the GPU box (gpurun), ONE GPU:
#   1. ncu launch list of a short headline bench run (kernel shares of the step)
#   2. one `ncu --set full` capture of both kernels for each single-GPU configuration
#   3. compute-sanitizer over every kernel path (tests/sanitize_run.py)
# Read here with tools/profile_r2_read.sh, which writes the summaries under profiles/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 60 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --only-headline --steps 8 --warmup 3 --no-verify --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_launches.log 2>&1
for spec in "config3 10000000" "config2 1000000" "config5 12500000"; do
  set -- $spec
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:regk_ -s 4 -c 2 -f -o gpurun_out/prof_r2_$1 \
      python tools/prof_one.py $1 $2 4 > gpurun_out/ncu_r2_$1.log 2>&1
done
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool python tests/sanitize_run.py 2>&1 | tail -4
done > gpurun_out/r2_sanitizer.txt 2>&1
ls -la gpurun_out/prof_r2_* gpurun_out/r2_*
