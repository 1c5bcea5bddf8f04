#!/bin/bash
# A/B at N GPUs (default 2): deferred redo passes against --no-defer-redo
TAG=${1:-sc}; N=${2:-2}
mkdir -p gpurun_out
for V in "" "--no-defer-redo" "" "--no-defer-redo"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu $V \
     > gpurun_out/${TAG}_bench_n${N}${V}.json 2> gpurun_out/${TAG}_bench_n${N}${V}.err
  echo "N=$N $V rc=$?"; tail -1 gpurun_out/${TAG}_bench_n${N}${V}.json | cut -c1-170
done
