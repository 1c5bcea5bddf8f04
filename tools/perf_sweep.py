#!/usr/bin/env python
"""A/B timing of kernel build variants (sim5_b200/variants/*.so): kernel ms of the main workloads + a golden check."""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import harness as H
from sim5_b200 import abi, api
api.init(0)
res = {}
p = abi.default_params(2, 64); got, _ = api.trace_image(p)
try:
    H.assert_image_parity(got.arrays, H.golden("image_cfg2_64.npz"), "golden"); res["golden"] = "ok"
except AssertionError as e:
    res["golden"] = str(e)
def t(p, reps=3):
    p.flags |= abi.FLAG_NO_OVERLAP            # one chunk: kernel_ms is the kernels alone and the per-kernel times are defined
    planes = api.HostPlanes(p)
    best = 1e30
    for _ in range(reps):
        _, st = api.trace_image(p, planes)
        if st.kernel_ms < best:
            best = st.kernel_ms
            try:
                res["_phases"] = [round(v, 3) for v in api.last_phase_ms()[0]]
            except api.Sim5Error:
                res["_phases"] = []
    return best, st
p = abi.default_params(2); ms, st = t(p); res["cfg2_phi_ms"] = round(ms,3); res["cfg2_phases_ms"] = res.pop("_phases"); res["cfg2_phi_rays_s"] = "%%.3e" %% (st.rays/ms*1e3)
if os.environ.get("SWEEP_EXACT"):
    p = abi.default_params(2); p.flags = abi.FLAG_EXACT_AZIMUTH; ms, st = t(p); res["cfg2_exact_ms"] = round(ms,3); res["cfg2_exact_phases_ms"] = res.pop("_phases")
p = abi.default_params(2); p.outputs = abi.OUT_R|abi.OUT_G|abi.OUT_FLUX|abi.OUT_STATUS; ms, st = t(p); res["cfg2_nophi_ms"] = round(ms,3); res["cfg2_nophi_rays_s"] = "%%.3e" %% (st.rays/ms*1e3)
p = abi.default_params(2); p.flags = abi.FLAG_SINGLE_PASS; ms, st = t(p); res["cfg2_single_pass_ms"] = round(ms,3); res.pop("_phases", None)
p = abi.default_params(3); ms, st = t(p); res["cfg3_ms"] = round(ms,3); res["cfg3_rays_s"] = "%%.3e" %% (st.rays/ms*1e3)
p = abi.default_params(4, 512); ms, st = t(p, 3); res["cfg4_512_ms"] = round(ms,2); res["cfg4_steps_s"] = "%%.3e" %% (st.total_steps/ms*1e3)
if os.environ.get("SWEEP_NOREFILL"):
    p.flags = abi.FLAG_NO_REFILL; ms, st = t(p, 1); res["cfg4_512_norefill_ms"] = round(ms,2)
p = abi.default_params(5, 512); p.n_spin, p.n_incl = 8, 4; ms, st = t(p, 2); res["cfg5_8x4x512_ms"] = round(ms,2); res["cfg5_rays_s"] = "%%.3e" %% (st.rays/ms*1e3)
p = abi.default_params(6); ms, st = t(p, 2); res["cfg6_ms"] = round(ms,3)
p = abi.default_params(7); ms, st = t(p, 2); res["cfg7_ms"] = round(ms,2)
p.flags |= abi.FLAG_NO_REFILL; ms, st = t(p, 1); res["cfg7_norefill_ms"] = round(ms,2)
if os.environ.get("SWEEP_CHUNKS"):
    import time
    L = api.lib()
    p = abi.default_params(2)
    planes = api.HostPlanes(p, pinned=True)
    for lg in (18, 19, 20, 21, 22, 23):
        L.sim5_set_chunk_rays(1 << lg)
        best = 1e30
        for _ in range(4):
            t0 = time.perf_counter(); api.trace_image(p, planes); best = min(best, time.perf_counter() - t0)
        res["e2e_chunk_2^%%d_ms" %% lg] = round(best * 1e3, 3)
    L.sim5_set_chunk_rays(0)
res.pop("_phases", None)
print(json.dumps(res))
''' % (ROOT, ROOT)
libs = sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))) or [os.path.join(ROOT, "sim5_b200", "libsim5b200.so")]
for lib in libs:
    env = dict(os.environ, SIM5_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-400:]
    print(os.path.basename(lib), line, flush=True)
