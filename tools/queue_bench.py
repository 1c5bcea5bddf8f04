#!/usr/bin/env python
"""Step-wise image (BASELINE configs[3]) on the GPUs of one box, device planes on GPU 0, ONE process, ONE call
(sim5_trace_image_multi + SIM5_FLAG_DEVICE_PTRS): the static interleaved row split against SIM5_FLAG_SHARED_QUEUE (all GPUs pull rays
of the one image from one counter on GPU 0, results stored into GPU 0's planes over NVLink).  Wall clock around the call, best of
--reps; every result is compared bit for bit with the one-GPU image.  No torch: ctypes + numpy only.
  python tools/queue_bench.py [--config 4] [--size 1024] [--reps 5]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from sim5_b200 import abi, api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--size", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
api.init(0)
L = api.lib()
ndev = L.sim5_gpu_device_count()
p = abi.default_params(args.config, args.size or None)
assert p.mode in (abi.MODE_STEPWISE, abi.MODE_SURFACE), "a lane mode (config 4 or 7)"
rays = p.nx * p.ny
names = tuple(n for n, _, _ in abi.PLANES)
img = api.DevicePlanes(p, names=names)
out = {"config": args.config, "nx": p.nx, "ny": p.ny, "rays": rays, "devices_on_box": ndev, "runs": []}
ref = None


def wipe():
    for name, ptr in img.ptrs.items():
        api.check(L.sim5_device_memset(C.c_void_p(ptr), 0xFF, img.n * np.dtype(img.dtypes[name]).itemsize), "sim5_device_memset")


for n in [d for d in (1, 2, 4, 8) if d <= ndev]:
    dl = (C.c_int * n)(*range(n))
    for queue in ("static", "shared"):
        q = abi.ImageParams.from_buffer_copy(p)
        q.flags |= abi.FLAG_DEVICE_PTRS | (abi.FLAG_SHARED_QUEUE if queue == "shared" else 0)
        st = abi.TraceStats()
        wipe()
        api.check(L.sim5_trace_image_multi(C.byref(q), C.byref(img.out), C.byref(st), dl, n), "sim5_trace_image_multi")
        got = {k: img.to_host(k) for k in img.ptrs}
        if ref is None:
            ref = got
        same = all(np.array_equal(ref[k], got[k], equal_nan=True) for k in ref)
        best = 1e30
        for _ in range(args.reps):
            t0 = time.perf_counter()
            api.check(L.sim5_trace_image_multi(C.byref(q), C.byref(img.out), C.byref(st), dl, n), "sim5_trace_image_multi")
            best = min(best, time.perf_counter() - t0)
        out["runs"].append({"gpus": n, "queue": queue, "wall_ms": round(best * 1e3, 3), "rays_per_s": round(rays / best, 0),
                            "slowest_device_kernel_ms": round(st.kernel_ms, 3), "rays_counted": int(st.rays), "steps": int(st.total_steps),
                            "equals_one_gpu_static_result": same})
img.close()
print(json.dumps(out))
