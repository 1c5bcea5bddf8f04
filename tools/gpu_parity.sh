#!/bin/bash
# GPU session: parity tests, full-size parity report against the unmodified reference, bench (our arm).
# usage (under gpurun): bash tools/gpu_parity.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python tools/parity_report.py --out gpurun_out/${TAG}_parity.json > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/${TAG}_parity.log
cat gpurun_out/${TAG}_parity.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
