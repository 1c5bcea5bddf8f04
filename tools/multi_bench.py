#!/usr/bin/env python
"""The C caller's view of a multi-GPU box: ONE process, ONE call -- sim5_trace_image_multi over devices 0..N-1 with pinned HOST planes
(every GPU copies its own rows home over its own PCIe link) -- timed by wall clock, next to the single-GPU sim5_trace_image call.
  python tools/multi_bench.py [--config 2] [--reps 5]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sim5_b200 import abi, api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
api.init(0)
ndev = api.lib().sim5_gpu_device_count()
p = abi.default_params(args.config)
rays = p.nx * p.ny * (p.n_spin * p.n_incl if p.mode == abi.MODE_HISTOGRAM else 1)
hp = api.HostPlanes(p, pinned=True)
out = {"config": args.config, "rays": rays, "devices_on_box": ndev, "runs": []}
ref = None
for n in [d for d in (1, 2, 4, 8) if d <= ndev]:
    devs = list(range(n))
    for a in hp.arrays.values():
        a[...] = 0
    api.trace_image_multi(p, devs, hp)
    best = 1e30
    for _ in range(args.reps):
        t0 = time.perf_counter()
        _, st = api.trace_image_multi(p, devs, hp)
        best = min(best, time.perf_counter() - t0)
    key = "hist" if p.mode == abi.MODE_HISTOGRAM else ("g" if "g" in hp.arrays else "intensity")
    chk = float(np.nansum(hp[key]))
    if ref is None:
        ref = {k: v.copy() for k, v in hp.arrays.items()}
        same = True
    elif p.mode == abi.MODE_HISTOGRAM:
        same = bool(np.allclose(hp["hist"], ref["hist"], rtol=1e-7, atol=0))
    else:
        same = all(np.array_equal(ref[k], hp[k], equal_nan=True) for k in ref)
    out["runs"].append({"gpus": n, "wall_ms": round(best * 1e3, 3), "rays_per_s": round(rays / best, 0), "slowest_device_kernel_ms": round(st.kernel_ms, 3),
                        "checksum": chk, "equals_one_gpu_result": same})
print(json.dumps(out))
