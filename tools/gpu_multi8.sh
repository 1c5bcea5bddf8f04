#!/bin/bash
# 8-GPU session: multi-device C entry, scaling of cfg 2 / 4 / 5, D2H floors, single-process multi bench.  usage (under gpurun --gpus 8): bash tools/gpu_multi8.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -m pytest tests/test_multi_device.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log
tail -3 gpurun_out/${TAG}_pytest_multi.log
run() {  # cfg N K
  if [ $2 = 1 ]; then
    python bench.py --config $1 --steps $3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_cfg$1_n$2.json 2> gpurun_out/${TAG}_bench_cfg$1_n$2.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --config $1 --gpus $2 --steps $3 --warmup 3 \
      > gpurun_out/${TAG}_bench_cfg$1_n$2.json 2> gpurun_out/${TAG}_bench_cfg$1_n$2.err
  fi
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_cfg$1_n$2.json").read().strip().split("\n")[-1])
    print("cfg$1 N=$2 value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), (d["e2e"].get("image_check") or {}).get("host_image_equals_device_image"))
except Exception as e:
    print("cfg$1 N=$2 unreadable:", e)
P
}
run 2 8 20; run 2 4 20; run 2 2 20; run 2 1 20
run 5 8 3; run 4 8 3; run 5 4 3; run 4 4 3
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 tools/d2h_roofline.py > gpurun_out/${TAG}_d2h_n$n.json 2> gpurun_out/${TAG}_d2h_n$n.err; tail -n 1 gpurun_out/${TAG}_d2h_n$n.json
done
python tools/multi_bench.py --config 2 > gpurun_out/${TAG}_multi_bench_cfg2.json 2> gpurun_out/${TAG}_multi_bench.err; cat gpurun_out/${TAG}_multi_bench_cfg2.json
python tools/multi_bench.py --config 5 --reps 2 > gpurun_out/${TAG}_multi_bench_cfg5.json 2>> gpurun_out/${TAG}_multi_bench.err; cat gpurun_out/${TAG}_multi_bench_cfg5.json
tail -q -n 3 gpurun_out/${TAG}_*.err | grep -v "^$" | tail -20
