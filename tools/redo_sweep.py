#!/usr/bin/env python
"""A/B of the redo-pass CTA size (sim5_b200/variants/*.so): per-kernel times of the bench camera on a full image and on an eighth of
its rows (what one of 8 GPUs traces)."""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import harness as H
from sim5_b200 import abi, api
api.init(0)
res = {}
for tag, ny in (("full", 4096), ("eighth", 512)):
    p = abi.default_params(2, 4096, ny); p.flags |= abi.FLAG_NO_OVERLAP
    p.rmax = abi.r_ms(p.bh_spin) + 20.0
    planes = api.HostPlanes(p)
    best = None
    for _ in range(4):
        _, st = api.trace_image(p, planes)
        ph = api.last_phase_ms()[0]
        if best is None or sum(ph) < sum(best): best = ph
    res[tag] = [round(v, 3) for v in best]
print(json.dumps(res))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))) or [os.path.join(ROOT, "sim5_b200", "libsim5b200.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SIM5_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
