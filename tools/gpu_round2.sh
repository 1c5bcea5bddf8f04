#!/bin/bash
# One 1-GPU session: all GPU tests, bench (both arms), ncu launch list + full capture of the cfg 2 kernels and of the stepwise kernel.
# usage (under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg5.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth_fast" -s 12 -c 3 -f -o gpurun_out/${TAG}_prof_eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes" -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_step \
    python bench.py --config 4 --size 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_step.log 2>&1
python tools/multi_bench.py --config 2 > gpurun_out/${TAG}_multi_bench.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench.json; cut -c1-300 gpurun_out/${TAG}_bench_cfg4.json; tail -n 5 gpurun_out/${TAG}_bench.err
