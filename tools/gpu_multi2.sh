#!/bin/bash
# N-GPU validation: the multi-device C entry, the torchrun bench for cfg 2 / 4 / 5.  usage (under gpurun --gpus N): bash tools/gpu_multi2.sh <tag> <N>
TAG=${1:-r05}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -m pytest tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log
tail -4 gpurun_out/${TAG}_pytest_multi.log
for c in 2 4 5; do
  K=10; [ $c = 5 ] && K=3; [ $c = 4 ] && K=3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config $c --gpus $N --steps $K --warmup 3 \
      > gpurun_out/${TAG}_bench_cfg${c}_n$N.json 2> gpurun_out/${TAG}_bench_cfg${c}_n$N.err; echo "cfg$c N=$N rc=$?"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_cfg${c}_n$N.json").read().strip().split("\n")[-1])
    print("cfg$c N=$N value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"].get("image_check"), d["roofline"].get("kernels_ms"))
except Exception as e:
    print("unreadable:", e)
P
done
tail -q -n 4 gpurun_out/${TAG}_bench_cfg*_n$N.err
