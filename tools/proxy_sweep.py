#!/usr/bin/env python
"""A/B of build variants (sim5_b200/variants/*.so) on the 4096^2 image and on the per-rank slice of an 8-GPU run (1448^2 ~ 2.1 M rays):
device planes, a train of 30 async calls, per-kernel times from the library's events."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import harness as H
from sim5_b200 import abi, api
api.init(0)
L = api.lib()
res = {}
p = abi.default_params(2, 64); got, _ = api.trace_image(p)
try:
    H.assert_image_parity(got.arrays, H.golden("image_cfg2_64.npz"), "golden"); res["golden"] = "ok"
except AssertionError as e:
    res["golden"] = str(e)
for n in (4096, 1448):
    p = abi.default_params(2, n)
    img = api.DevicePlanes(p)
    q = abi.ImageParams.from_buffer_copy(p)
    q.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC | abi.FLAG_DEFER_REDO
    st = abi.TraceStats()
    for rep in range(2):
        for _ in range(30):
            api.check(L.sim5_trace_image(C.byref(q), C.byref(img.out), C.byref(st)), "trace")
        api.check(L.sim5_synchronize(), "sync")
    a = sum(api.phase_history(b)[0][0] for b in range(20)) / 20
    z = sum(api.phase_history(b)[0][1] for b in range(20)) / 20
    res["n%%d_phaseA_ms" %% n] = round(a, 4); res["n%%d_az_ms" %% n] = round(z, 4)
    img.close()
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SIM5_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
