#!/bin/bash
# GPU session: parity tests of the default build + A/B timing of the kernel variants under sim5_b200/variants/
TAG=${1:-s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
SWEEP_EXACT=1 python tools/perf_sweep.py > gpurun_out/${TAG}_sweep.log 2>&1; cat gpurun_out/${TAG}_sweep.log
