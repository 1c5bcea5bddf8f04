#!/usr/bin/env python
"""SPECTRUM mode (N3) on the GPU: the preset (2048^2 image, 128 energies) timed, checked against the unmodified reference on
a 512^2 sample of the same camera (the spectrum scales with the pixel area, so the two agree to discretisation ~1e-3)."""
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
p = abi.default_params(6)
hp = api.HostPlanes(p, pinned=False)
best = 1e30
for _ in range(3):
    _, st = api.trace_image(p, hp)
    best = min(best, st.kernel_ms)
res["preset"] = "2048x2048, %d energies" % p.n_energy
res["kernel_ms"] = round(best, 3)
res["rays_per_s"] = "%.3e" % (st.rays / best * 1e3)
res["terms_per_s"] = "%.3e" % (sum(st.class_count[i] for i in (0, 1, 5)) * p.n_energy / best * 1e3)
full = hp["spectrum"].copy()
if H.have_ref():
    q = abi.default_params(6, 512)
    ref, dt = H.run_spectrum("ref", q)
    got, st2 = api.trace_image(q, api.HostPlanes(q, pinned=False))
    res["ref_512_s"] = round(dt, 3)
    res["ref_rays_per_s"] = "%.3e" % (512 * 512 / dt)
    res["gpu_vs_ref_512_max_rel_to_peak"] = float(np.max(np.abs(got["spectrum"] - ref)) / ref.max())
    res["coarse_vs_fine_grid_rel"] = float(np.max(np.abs(full - ref)) / ref.max())
print(json.dumps(res))
