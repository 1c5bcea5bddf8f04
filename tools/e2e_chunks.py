import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from sim5_b200 import abi, api
import numpy as np
api.init(0)
L = api.lib()
res = {}
p = abi.default_params(2)
planes = api.HostPlanes(p, pinned=True)
ref = None
for lg in (19, 20, 21, 22):
    L.sim5_set_chunk_rays(1 << lg)
    best = 1e30
    for _ in range(5):
        t0 = time.perf_counter(); api.trace_image(p, planes); best = min(best, time.perf_counter() - t0)
    res["e2e_chunk_2^%d_ms" % lg] = round(best * 1e3, 3)
    cs = float(np.nansum(planes["g"])) + float(np.nansum(planes["phi"]))
    ref = cs if ref is None else ref
    assert cs == ref, (cs, ref)
print(json.dumps(res))
