#!/bin/bash
# quick GPU iteration: parity tests + micro benchmarks + perf sweep of the default build
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python tools/micro_carlson.py > gpurun_out/${TAG}_micro.log 2>&1; cat gpurun_out/${TAG}_micro.log
python tools/perf_sweep.py > gpurun_out/${TAG}_sweep.log 2>&1; cat gpurun_out/${TAG}_sweep.log
