#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list + one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace_eqplane|k_azimuth" -s 12 -c 6 -f -o gpurun_out/${TAG}_prof_eqplane \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
python tools/micro_carlson.py > gpurun_out/${TAG}_micro.log 2>&1
timeout 300 python tools/surface_bench.py > gpurun_out/${TAG}_surface.json 2> gpurun_out/${TAG}_surface.err
timeout 300 python tools/spectrum_bench.py > gpurun_out/${TAG}_spectrum.json 2> gpurun_out/${TAG}_spectrum.err
cat gpurun_out/${TAG}_bench.json
