#!/bin/bash
# ncu full capture of selected kernels of the bench workload: bash tools/gpu_prof.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-p}; KRE=${2:-k_azimuth}; SKIP=${3:-2}; CNT=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
