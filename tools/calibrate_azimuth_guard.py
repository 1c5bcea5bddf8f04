#!/usr/bin/env python
"""Calibration of the conditioning guard of the tolerance-mode azimuth (pixel.cuh: azimuth_well_conditioned).

Builds nothing itself: expects a -DS5_AZ_DIAG host build of tests/hostsim (see the g++ line in DESIGN.md section 3) at
/tmp/libhostsim_diag.so and the unmodified reference (oracle/_ref).  For a set of cameras it prints, per decade of the
indicator kappa = mag / max(|phi|, 1), the number of disk hits and the worst deviation of the UNGUARDED tolerance-mode
phi from the reference -- the data behind S5_AZ_COND_LIMIT.
"""
import ctypes as C
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from sim5_b200 import abi  # noqa: E402

lib = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else "/tmp/libhostsim_diag.so")
dp = C.POINTER(C.c_double)
edges = [0, 1e2, 1e3, 1e4, 3e4, 1e5, 3e5, 1e6, 1e7, 1e9, 1e300]
worst = np.zeros(len(edges) - 1)
count = np.zeros(len(edges) - 1, dtype=np.int64)
cams = []
for spin in (0.998, 0.9, 0.5, 0.0):
    for inc in (5.0, 30.0, 60.0, 75.0, 88.0):
        cams.append((spin, inc))
n = int(os.environ.get("CAL_N", "640"))
for spin, inc in cams:
    p = abi.default_params(2, n)
    p.bh_spin = spin
    p.incl = abi.deg2rad(inc)
    p.rmax = abi.r_ms(max(spin, 1e-4)) + 20.0
    ref, _, _ = H.run_ref(p)
    kap = np.zeros(n * n)
    phi = np.zeros(n * n)
    lib.hs_fast_azimuth_kappa(C.byref(p), kap.ctypes.data_as(dp), phi.ctypes.data_as(dp))
    hit = (ref["status"] & 31) <= 1
    e = H.rel_err(phi, ref["phi"], 1.0)
    e[~hit] = 0.0
    e[np.isnan(phi) & hit] = 0.0          # domain fallbacks (NaN): handled by the redo pass, not by the guard
    for b in range(len(edges) - 1):
        m = hit & (kap >= edges[b]) & (kap < edges[b + 1])
        count[b] += m.sum()
        if m.any():
            worst[b] = max(worst[b], e[m].max())
    print("a=%.3f i=%4.1f: hits %d, worst unguarded deviation %.2e (kappa there %.1e)" % (spin, inc, hit.sum(), e.max(), kap[np.argmax(e)]), flush=True)
print("\nkappa range            hits      worst |dphi|/max(|phi|,1)")
for b in range(len(edges) - 1):
    print("[%8.0e, %8.0e)  %9d   %.2e" % (edges[b], edges[b + 1], count[b], worst[b]))
