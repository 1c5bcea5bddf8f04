#!/bin/bash
# N-GPU box (run with gpurun --gpus 4): the shared ray queue (SIM5_FLAG_SHARED_QUEUE) -- its tests, then static split vs shared queue on cfg 4 and the SURFACE preset.
# No torch anywhere in this script (ctypes + numpy): the box's first `import torch` alone would cost a GPU-minute.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r07b_gpus.txt 2>&1
timeout 120 python -m pytest tests/test_multi_device.py -x -q -m gpu -k "shared_queue or device_planes" -p no:cacheprovider > gpurun_out/r07b_pytest_queue.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r07b_pytest_queue.log
tail -3 gpurun_out/r07b_pytest_queue.log
timeout 100 python tools/queue_bench.py --config 4 --reps 4 > gpurun_out/r07b_queue_cfg4.json 2> gpurun_out/r07b_queue_cfg4.err
cat gpurun_out/r07b_queue_cfg4.json; tail -2 gpurun_out/r07b_queue_cfg4.err
timeout 60 python tools/queue_bench.py --config 7 --reps 4 > gpurun_out/r07b_queue_cfg7.json 2> gpurun_out/r07b_queue_cfg7.err
cat gpurun_out/r07b_queue_cfg7.json; tail -2 gpurun_out/r07b_queue_cfg7.err
