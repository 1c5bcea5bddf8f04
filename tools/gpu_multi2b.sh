#!/bin/bash
TAG=${1:-r05}; N=${2:-2}
mkdir -p gpurun_out
python -m pytest tests/test_multi_device.py -m gpu -x -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log; tail -3 gpurun_out/${TAG}_pytest_multi.log
for mode in dma stores; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --config 2 --gpus $N --steps 20 --warmup 3 --peer-copy $mode \
      > gpurun_out/${TAG}_bench_cfg2_n${N}_$mode.json 2> gpurun_out/${TAG}_bench_cfg2_n${N}_$mode.err
  python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg2_n${N}_$mode.json").read().strip().split("\n")[-1])
print("cfg2 N=$N $mode value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), (d["e2e"].get("image_check") or {}).get("host_image_equals_device_image"), [ (r["ms_per_step"], r["phase_a_ms"], r["azimuth_ms"]) for r in d["roofline"]["per_rank"]])
P
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --config 4 --gpus $N --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg4_n$N.json 2> gpurun_out/${TAG}_bench_cfg4_n$N.err
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_cfg4_n1.json 2> gpurun_out/${TAG}_bench_cfg4_n1.err
python - <<P
import json
for f in ("gpurun_out/${TAG}_bench_cfg4_n$N.json", "gpurun_out/${TAG}_bench_cfg4_n1.json"):
    d = json.loads(open(f).read().strip().split("\n")[-1])
    print(f, "value %.4e ms/step %.3f" % (d["value"], d["ms_per_step"]), [r["ms_per_step"] for r in d["roofline"]["per_rank"]], d["roofline"]["frac"])
P
tail -q -n 3 gpurun_out/${TAG}_*.err | grep -iE "error|Traceback|assert" | head
