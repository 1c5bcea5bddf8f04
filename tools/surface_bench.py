#!/usr/bin/env python
"""SURFACE mode on the GPU: the preset (1024^2, thick disk H/R -> 0.2) timed on the device, then checked
against the unmodified reference (oracle/_ref, all host threads) on a 256^2 sample of the same camera."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from sim5_b200 import abi, api  # noqa: E402

api.init(0)
res = {}
p = abi.default_params(7)
hp = api.HostPlanes(p, pinned=True)
best, st = 1e30, None
for _ in range(3):
    _, st = api.trace_image(p, hp)
    best = min(best, st.kernel_ms)
res["preset"] = "%dx%d, a=%.3g, i=60deg, H/R->%.2g" % (p.nx, p.ny, p.bh_spin, p.surf_hr)
res["kernel_ms"] = round(best, 3)
res["rays_per_s"] = "%.3e" % (st.rays / best * 1e3)
res["follow_calls"] = int(st.total_steps)
res["follow_calls_per_s"] = "%.3e" % (st.total_steps / best * 1e3)
res["classes"] = {str(i): int(c) for i, c in enumerate(st.class_count) if c}
res["grid"] = [st.grid_ctas, st.cta_threads]
q = abi.default_params(7); q.flags |= abi.FLAG_NO_REFILL
_, st3 = api.trace_image(q, api.HostPlanes(q, pinned=True))
res["kernel_ms_no_refill"] = round(st3.kernel_ms, 3)
if H.have_ref():
    q = abi.default_params(7, 256)
    ref, rst, dt = H.run_ref(q)
    got, st2 = api.trace_image(q, api.HostPlanes(q, pinned=True))
    rep = H.assert_image_parity(got.arrays, ref.arrays, label="surface 256^2")
    res["ref_256_s"] = round(dt, 3)
    res["ref_rays_per_s"] = "%.3e" % (256 * 256 / dt)
    res["ref_threads"] = H.load_ref().ref_max_threads()
    res["parity_256"] = {k: "max %.2e exact %.4f" % (v["max"], v["exact"]) for k, v in rep.items()}
    res["status_and_steps_identical"] = bool(list(st2.class_count) == list(rst.class_count) and st2.total_steps == rst.total_steps)
print(json.dumps(res))
