#!/bin/bash
# Run on the GPU box (gpurun): launch list of a short bench run + one full ncu capture of both kernels.
# usage: tools/profile_run.sh TAG
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:regk_ -s 6 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
