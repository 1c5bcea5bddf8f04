#!/bin/bash
# ncu full capture of the lane kernels (modes STEPWISE and SURFACE) on small images: bash tools/gpu_prof_modes.sh <tag>
TAG=${1:-pm}
mkdir -p gpurun_out
cat > /tmp/run_modes.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from sim5_b200 import abi, api
api.init(0)
for cfg, n in ((4, 512), (7, 1024), (5, 512), (6, 2048)):
    p = abi.default_params(cfg, n)
    if cfg == 5:
        p.n_spin, p.n_incl = 8, 4
    api.trace_image(p, api.HostPlanes(p, pinned=False))
PY
ncu --set full --clock-control none --import-source on -k regex:"k_trace_lanes|k_trace_histogram|k_trace_spectrum" -f -o gpurun_out/${TAG}_prof_modes python /tmp/run_modes.py > gpurun_out/${TAG}_ncu_modes.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_modes.log
