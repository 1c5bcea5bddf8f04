import glob, os, subprocess, sys, json
ROOT="/root/repo"
CHILD = r'''
import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import harness as H
from sim5_b200 import abi, api
api.init(0)
res = {}
p = abi.default_params(4, 16); p.outputs |= abi.OUT_QERR; got, _ = api.trace_image(p)
try:
    H.assert_image_parity(got.arrays, H.golden("image_cfg4_16.npz"), "golden"); res["golden"] = "ok"
except AssertionError as e:
    res["golden"] = str(e)
for flags, tag in ((0, "refill"), (abi.FLAG_NO_REFILL, "norefill")):
    p = abi.default_params(4, 512); p.flags |= flags | abi.FLAG_NO_OVERLAP
    planes = api.HostPlanes(p)
    best = 1e30
    for _ in range(2):
        _, st = api.trace_image(p, planes); best = min(best, st.kernel_ms)
    res["cfg4_512_%s_ms" % tag] = round(best, 2); res["steps_s_%s" % tag] = "%.3e" % (st.total_steps / best * 1e3)
print(json.dumps(res))
'''
for lib in sorted(glob.glob(os.path.join(ROOT, "sim5_b200", "variants", "*.so"))):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, SIM5_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
