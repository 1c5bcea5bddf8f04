#!/usr/bin/env python
"""A/B of stepwise-kernel build variants (sim5_b200/variants/*.so): cfg 4 at 512^2 and 1024^2, golden check, best of 3."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import harness as H
from sim5_b200 import abi, api
api.init(0)
res = {}
p = H.golden_params(4, 16, 16, abi.OUT_QERR); got, _ = api.trace_image(p)
try:
    H.assert_image_parity(got.arrays, H.golden("image_cfg4_16.npz"), "golden"); res["golden"] = "ok"
except AssertionError as e:
    res["golden"] = str(e)
for n in (512, 1024):
    p = abi.default_params(4, n); p.flags |= abi.FLAG_NO_OVERLAP
    planes = api.HostPlanes(p)
    best = 1e30
    for _ in range(3):
        _, st = api.trace_image(p, planes)
        best = min(best, st.kernel_ms)
    res["cfg4_%%d_ms" %% n] = round(best, 2); res["cfg4_%%d_steps_s" %% n] = "%%.3e" %% (st.total_steps / best * 1e3)
    res["sum_I_%%d" %% n] = float(planes["intensity"].sum()); res["steps_%%d" %% n] = int(st.total_steps)
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SIM5_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
