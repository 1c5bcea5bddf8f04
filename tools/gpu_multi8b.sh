#!/bin/bash
# 8-GPU A/B session: image assembly by DMA vs peer stores (cfg 2), row order of the stepwise kernel (cfg 4).  usage (under gpurun --gpus 8): bash tools/gpu_multi8b.sh <tag>
TAG=${1:-r05}
mkdir -p gpurun_out
run() {  # name cfg N K extra...
  name=$1; cfg=$2; n=$3; k=$4; shift 4
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --config $cfg --gpus $n --steps $k --warmup 3 "$@" \
      > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read().strip().split("\n")[-1])
    print("${name}: value %.4e ms/step %.4f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), (d["e2e"].get("image_check") or {}).get("host_image_equals_device_image"),
          [(r["ms_per_step"], r["phase_a_ms"], r["azimuth_ms"]) for r in d["roofline"]["per_rank"]])
except Exception as e:
    print("${name} unreadable:", e)
P
}
run cfg2_n8_dma 2 8 20 --peer-copy dma
run cfg2_n8_stores 2 8 20 --peer-copy stores
run cfg2_n4_dma 2 4 20 --peer-copy dma
run cfg2_n4_stores 2 4 20 --peer-copy stores
run cfg4_n8_centerout 4 8 3
run cfg4_n8_rowmajor 4 8 3 --row-major
run cfg4_n4_centerout 4 4 3
run cfg3_n8 3 8 20
run cfg1_n8 1 8 20
tail -q -n 3 gpurun_out/${TAG}_bench_*.err | grep -iE "error|Traceback|assert" | head
