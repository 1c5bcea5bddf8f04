#!/bin/bash
# scaling session (gpurun --gpus 8): the bench at N = 8, 4, 2 with the peer-memory assembly
TAG=${1:-sc}
mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
     > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
  echo "N=$N rc=$?"; tail -1 gpurun_out/${TAG}_bench_n${N}.json | cut -c1-200
done
