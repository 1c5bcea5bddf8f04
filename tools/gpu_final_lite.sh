#!/bin/bash
# short final check of a build: all GPU tests, smoke, bench cfg 2 / 3, ncu summary of the cfg 2 kernels (condensed on the box)
TAG=${1:-r06}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_cfg3.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
R=/tmp/${TAG}_rep; mkdir -p $R
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth_fast" -s 12 -c 3 -f -o $R/eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
python tools/ncu_summary.py $R/eqplane.ncu-rep gpurun_out/${TAG}_ncu_summary.csv > /dev/null 2>&1
cut -c1-260 gpurun_out/${TAG}_bench.json; cut -c1-200 gpurun_out/${TAG}_bench_cfg3.json
grep -E "^metric|duration|dram__bytes|registers" gpurun_out/${TAG}_ncu_summary.csv | cut -c1-140
tail -n 3 gpurun_out/${TAG}_bench.err
