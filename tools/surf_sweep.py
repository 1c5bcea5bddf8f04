#!/usr/bin/env python
"""A/B of SURFACE kernel variants (sim5_b200/variants/*.so): kernel ms of the 1024^2 preset + golden check."""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import harness as H
from sim5_b200 import abi, api
api.init(0)
res = {}
p = abi.default_params(7, 48); got, _ = api.trace_image(p)
try:
    H.assert_image_parity(got.arrays, H.golden("image_cfg7_48.npz"), "golden"); res["golden"] = "ok"
except AssertionError as e:
    res["golden"] = str(e)[:100]
p = abi.default_params(7); planes = api.HostPlanes(p)
best = 1e30
for _ in range(3):
    _, st = api.trace_image(p, planes); best = min(best, st.kernel_ms)
res["cfg7_ms"] = round(best, 2); res["grid"] = [st.grid_ctas, st.cta_threads]
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))) or [os.path.join(ROOT, "sim5_b200", "libsim5b200.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SIM5_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
