#!/usr/bin/env python
"""PCIe roofline of the e2e number: the plain device->host copy of one image's planes (553 MB at 4096^2) from device memory into the
same pinned host planes the e2e call fills, timed alone -- the floor of any host-plane call."""
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from sim5_b200 import abi, api  # noqa: E402

api.init(0)
L = api.lib()
p = abi.default_params(2)
hp = api.HostPlanes(p, pinned=True)
dp = api.DevicePlanes(p, names=("r", "phi", "g", "flux", "status"))
nbytes = sum(a.nbytes for a in hp.arrays.values())
best = 1e30
for _ in range(5):
    t0 = time.perf_counter()
    for k, a in hp.arrays.items():
        api.check(L.sim5_device_to_host(a.ctypes.data, C.c_void_p(dp.ptrs[k]), a.nbytes), "d2h")
    best = min(best, time.perf_counter() - t0)
planes, st = api.trace_image(p, hp)
t_e2e = 1e30
for _ in range(5):
    t0 = time.perf_counter(); api.trace_image(p, hp); t_e2e = min(t_e2e, time.perf_counter() - t0)
print(json.dumps({"bytes": nbytes, "d2h_ms": round(best * 1e3, 3), "d2h_gb_s": round(nbytes / best / 1e9, 2),
                  "e2e_ms": round(t_e2e * 1e3, 3), "e2e_over_d2h": round(t_e2e / best, 3)}))
dp.close()
