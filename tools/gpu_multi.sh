#!/bin/bash
# multi-GPU session (gpurun --gpus N): parity tests on GPU 0, then the bench at N GPUs with both image-assembly paths
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
for G in peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --gather $G \
     > gpurun_out/${TAG}_bench_n${N}_${G}.json 2> gpurun_out/${TAG}_bench_n${N}_${G}.err
  echo "bench $G rc=$?"; tail -2 gpurun_out/${TAG}_bench_n${N}_${G}.err; cat gpurun_out/${TAG}_bench_n${N}_${G}.json
done
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cat gpurun_out/${TAG}_bench_n1.json
