#!/usr/bin/env python
"""Ad-hoc GPU check used during development: parity of the CUDA path vs the unmodified reference on
small grids of configs 1-4, then timing of the full config-2 image and the FP64 DFMA peak."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
from sim5_b200 import abi, api

def compare(p, label):
    t0 = time.time(); a, st = api.trace_image(p); t1 = time.time()
    b, _, dtb = H.run_ref(p)
    print("== %s: gpu call %.3fs (kernel %.3f ms, total %.3f ms), ref %.2fs" % (label, t1-t0, st.kernel_ms, st.total_ms, dtb))
    print("   classes", list(st.class_count)[:13], "gtypes", list(st.gtype_count)[:6], "steps", st.total_steps)
    print("   status mismatches: %d / %d" % (np.sum(a['status'] != b['status']), a['status'].size))
    for k in a.arrays:
        if k == 'status': continue
        if k == 'steps':
            print("   steps mismatches:", int(np.sum(a[k] != b[k]))); continue
        floor = 1.0 if k in ('phi', 'chi') else 0.0
        s = H.err_summary(a[k], b[k], floor)
        print("   %-9s max %.3e  p99.9 %.3e  exact %.6f" % (k, s['max'], s['p999'], s['exact']))

api.init(0)
print("fp64 peak TFLOP/s:", api.fp64_peak_tflops(0, 8192))
for cfg, n in ((1, 512), (2, 512), (3, 512), (4, 96)):
    p = abi.default_params(cfg, n)
    compare(p, "cfg%d %dx%d" % (cfg, p.nx, p.ny))
# timing of the big ones
for cfg, outs in ((2, None), (2, abi.OUT_R | abi.OUT_G | abi.OUT_FLUX | abi.OUT_STATUS), (3, None), (1, None)):
    p = abi.default_params(cfg)
    if cfg == 1: p.nx = p.ny = 4096
    if outs is not None: p.outputs = outs
    planes = api.HostPlanes(p)
    for it in range(3):
        planes, st = api.trace_image(p, planes)
    print("cfg%d %dx%d outputs=0x%x: kernel %.3f ms (%.3e rays/s), total %.3f ms (%.3e rays/s e2e) grid %d"
          % (cfg, p.nx, p.ny, p.outputs, st.kernel_ms, st.rays/st.kernel_ms*1e3, st.total_ms, st.rays/st.total_ms*1e3, st.grid_ctas))
p = abi.default_params(4, 256)
for flags in (0, abi.FLAG_NO_REFILL):
    p.flags = flags
    planes, st = api.trace_image(p)
    print("cfg4 256x256 flags=%d: kernel %.3f ms, %.3e rays/s, %.3e steps/s, mean steps %.1f" % (flags, st.kernel_ms, st.rays/st.kernel_ms*1e3, st.total_steps/st.kernel_ms*1e3, st.total_steps/st.rays))
