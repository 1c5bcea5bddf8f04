#!/usr/bin/env python
"""CPU time to ENQUEUE one asynchronous sim5_trace_image call (no sync) against the device time of the step, on an eighth of the bench
image (what one of 8 GPUs traces): if the enqueue time approaches the step time, the train is launch-bound."""
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from sim5_b200 import abi, api  # noqa: E402

api.init(0)
L = api.lib()
res = {}
for tag, ny in (("eighth", 512), ("full", 4096)):
    p = abi.default_params(2, 4096, ny)
    p.rmax = abi.r_ms(p.bh_spin) + 20.0
    img = api.DevicePlanes(p)
    p.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC | abi.FLAG_DEFER_REDO
    st = abi.TraceStats()
    for _ in range(5):
        api.check(L.sim5_trace_image(C.byref(p), C.byref(img.out), C.byref(st)), "trace")
    api.check(L.sim5_synchronize(), "sync")
    K = 50
    t0 = time.perf_counter()
    for _ in range(K):
        api.check(L.sim5_trace_image(C.byref(p), C.byref(img.out), C.byref(st)), "trace")
    t1 = time.perf_counter()
    api.check(L.sim5_synchronize(), "sync")
    t2 = time.perf_counter()
    res[tag] = {"enqueue_us_per_call": round((t1 - t0) / K * 1e6, 1), "wall_us_per_call": round((t2 - t0) / K * 1e6, 1)}
    img.close()
print(json.dumps(res))
