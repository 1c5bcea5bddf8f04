#!/bin/bash
# 8-GPU sanity of the final build through the driver's own launch lines.  usage (under gpurun --gpus 8): bash tools/gpu_multi8c.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
python -m pytest tests/test_multi_device.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log; tail -2 gpurun_out/${TAG}_pytest_multi.log
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 20 --warmup 3 \
      > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  python - <<P
import json
txt = open("gpurun_out/${TAG}_bench_n$n.json").read().strip().split("\n")
d = json.loads(txt[-1])
print("N=$n lines on stdout:", len(txt), "value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), (d["e2e"].get("image_check") or {}).get("host_image_equals_device_image"),
      [r["ms_per_step"] for r in d["roofline"]["per_rank"]])
P
done
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_n1.json').read().strip().split('\n')[-1]); print('N=1 value %.4e ms/step %.4f' % (d['value'], d['ms_per_step']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 \
      > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err; echo "reference arm under torchrun rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --config 5 --gpus 8 --steps 3 --warmup 3 \
      > gpurun_out/${TAG}_bench_cfg5_n8.json 2> gpurun_out/${TAG}_bench_cfg5_n8.err
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_cfg5_n8.json').read().strip().split('\n')[-1]); print('cfg5 N=8 value %.4e ms/step %.4f' % (d['value'], d['ms_per_step']), d['e2e']['image_check'])"
tail -q -n 3 gpurun_out/${TAG}_bench_*.err | grep -iE "error|Traceback|assert" | head
