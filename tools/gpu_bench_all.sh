#!/bin/bash
# all five BASELINE configurations through bench.py on one GPU, our arm and the reference arm.  usage (under gpurun): bash tools/gpu_bench_all.sh <tag> [steps]
TAG=${1:-r05}
K=${2:-5}
mkdir -p gpurun_out
for c in 1 2 3 4 5; do
  python bench.py --config $c --steps $K --warmup 3 > gpurun_out/${TAG}_bench_cfg$c.json 2> gpurun_out/${TAG}_bench_cfg$c.err; echo "cfg$c ours rc=$?"
  python bench.py --impl reference --config $c --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_cfg$c.json 2>> gpurun_out/${TAG}_bench_cfg$c.err; echo "cfg$c ref rc=$?"
  python - <<P
import json
for f in ("gpurun_out/${TAG}_bench_cfg$c.json", "gpurun_out/${TAG}_bench_ref_cfg$c.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        r = d.get("roofline") or {}
        print(f, "value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "roofline", r.get("kernel"), r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
P
done
tail -q -n 3 gpurun_out/${TAG}_bench_cfg*.err
