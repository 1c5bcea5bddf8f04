#!/bin/bash
# 1-GPU A/B of the alternating launch streams of a deferred train, at the full image and at the per-rank slice of an 8-GPU run (1448^2 ~ 2.1 M rays)
TAG=${1:-r05}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_pytest.log
for size in 4096 1448; do
  for mode in "" "--alt-streams"; do
    for rep in 1 2; do
      python bench.py --size $size --defer-redo on $mode --steps 40 --warmup 5 --no-cpu > gpurun_out/${TAG}_ab.json 2> gpurun_out/${TAG}_ab.err
      python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_ab.json").read().strip().split("\n")[-1])
print("size $size mode '$mode' rep $rep: ms/step %.4f value %.4e" % (d["ms_per_step"], d["value"]), d["roofline"]["kernels_ms"])
P
    done
  done
done
tail -n 3 gpurun_out/${TAG}_ab.err
