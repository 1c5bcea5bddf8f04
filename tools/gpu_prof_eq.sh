#!/bin/bash
# ncu --set full of the cfg 2 kernels + launch list.  usage (under gpurun): bash tools/gpu_prof_eq.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
python tools/perf_sweep.py > gpurun_out/${TAG}_sweep.log 2>&1; cat gpurun_out/${TAG}_sweep.log
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth_fast" -s 10 -c 2 -f -o gpurun_out/${TAG}_prof_eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_prof_eqplane.ncu-rep
