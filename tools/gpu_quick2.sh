#!/bin/bash
# GPU iteration: parity tests + the mode benches of the "next" rows
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python tools/surface_bench.py > gpurun_out/${TAG}_surface.json 2> gpurun_out/${TAG}_surface.err; cat gpurun_out/${TAG}_surface.json; tail -3 gpurun_out/${TAG}_surface.err
timeout 300 python tools/spectrum_bench.py > gpurun_out/${TAG}_spectrum.json 2> gpurun_out/${TAG}_spectrum.err; cat gpurun_out/${TAG}_spectrum.json
