#!/usr/bin/env python
"""Full-size parity report (run on the GPU box): every BASELINE config at its full size, GPU (through the C-ABI) against
the unmodified reference (oracle/_ref) on the box's host cores.  Prints / writes, per output plane, the max and the
99.9th-percentile relative error, the share of bit-identical doubles and the number of status-byte mismatches --
the artefact BASELINE.json's north_star asks for ("reported as max and 99.9th-percentile error").

  python tools/parity_report.py [--out gpurun_out/parity.json] [--quick]

cfg 5 (2.1e9 rays) is sampled: a 4x2 sub-lattice of full 1024^2 images (the corners and the middle of the spin /
inclination lattice) -- the CPU reference needs ~0.35 s per image on 16 threads, the full lattice ~12 min.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from sim5_b200 import abi, api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    ap.add_argument("--quick", action="store_true", help="quarter-size images (smoke run of the tool)")
    args = ap.parse_args()
    assert H.have_ref(), "oracle/_ref/libsim5ref.so did not travel"
    api.init(0)
    sizes = {1: 512, 2: 4096, 3: 2048, 4: 1024, 7: 1024, "2+delay": 1024}
    if args.quick:
        sizes = {k: max(v // 4, 64) for k, v in sizes.items()}
    report = {"checker": "oracle/_ref (unmodified reference, %d host threads)" % H.load_ref().ref_max_threads(), "configs": {}}
    for cfg, n in sizes.items():
        if cfg == "2+delay":             # N2: the travel-time plane on the camera of config 2
            p = abi.default_params(2, n)
            p.outputs = abi.OUT_R | abi.OUT_DELAY | abi.OUT_STATUS
        else:
            p = abi.default_params(cfg, n)
        if cfg == 3:
            p.outputs |= abi.OUT_MUE
        if cfg == 4:
            p.outputs |= abi.OUT_QERR
        t0 = time.time()
        got, st = api.trace_image(p, api.HostPlanes(p, pinned=True))
        t1 = time.time()
        ref, rst, dt = H.run_ref(p)
        entry = {"size": "%dx%d" % (n, n), "rays": n * n, "gpu_call_s": t1 - t0, "gpu_kernel_ms": st.kernel_ms, "cpu_ref_s": dt,
                 "status_mismatches": int(np.sum(got["status"] != ref["status"])),
                 "class_count_equal": list(st.class_count) == list(rst.class_count),
                 "gtype_count_equal": list(st.gtype_count) == list(rst.gtype_count), "planes": {}}
        if cfg in (4, 7):
            entry["step_count_mismatches"] = int(np.sum(got["steps"] != ref["steps"]))
            entry["total_steps"] = int(st.total_steps)
        for k in got.arrays:
            if k in ("status", "steps"):
                continue
            s = H.err_summary(got[k], ref[k], H.FLOOR.get(k, 0.0))
            s["tol"] = H.TOL.get(k)
            s["within_tol"] = (s["tol"] is None) or (s["max"] <= s["tol"])
            entry["planes"][k] = s
        report["configs"]["cfg%s" % cfg] = entry
        print("cfg%s %s: status mismatches %d | %s" % (cfg, entry["size"], entry["status_mismatches"],
              " ".join("%s max %.2e p99.9 %.2e exact %.4f" % (k, v["max"], v["p999"], v["exact"]) for k, v in entry["planes"].items())), flush=True)

    # cfg 5: sub-lattice of full-size images
    p = abi.default_params(5)
    if args.quick:
        p.nx = p.ny = 256
    picks = [(js, ki) for js in (0, 21, 42, 63) for ki in (0, 31)]
    nb = p.n_bins
    worst, exact = 0.0, []
    t_gpu = t_cpu = 0.0
    for js, ki in picks:
        img = js * p.n_incl + ki
        p.lattice_begin, p.lattice_end = img, img + 1
        hp = api.HostPlanes(p, pinned=False)
        t0 = time.time()
        api.trace_image(p, hp)
        t_gpu += time.time() - t0
        hist = np.zeros(p.n_spin * p.n_incl * nb)
        t_cpu += H.load_ref().ref_trace_histogram(C.byref(p), hist.ctypes.data_as(C.POINTER(C.c_double)), 0, 1)
        a = hp["hist"].reshape(-1, nb)[img]
        b = hist.reshape(-1, nb)[img]
        s = H.err_summary(a, b)
        # bins are sums of ~1e3..1e5 positive terms accumulated in a different order: compare relative to the image's peak bin too
        worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)))
        exact.append(s["exact"])
        print("cfg5 image (spin %d, incl %d): max rel err per bin %.2e, max |d|/peak %.2e" % (js, ki, s["max"], worst), flush=True)
    report["configs"]["cfg5"] = {"size": "%dx%d x %d lattice images of 2048" % (p.nx, p.ny, len(picks)), "max_abs_over_peak": worst,
                                 "tol": 1e-7, "within_tol": worst <= 1e-7, "gpu_call_s": t_gpu, "cpu_ref_s": t_cpu}
    # N3: the thermal spectrum of the preset (2048^2 x 128 energies)
    p = abi.default_params(6)
    if args.quick:
        p.nx = p.ny = 512
    hp = api.HostPlanes(p, pinned=False)
    t0 = time.time(); api.trace_image(p, hp); t_gpu = time.time() - t0
    ref, t_cpu = H.run_spectrum("ref", p)
    s = H.err_summary(hp["spectrum"], ref)
    print("spectrum %dx%d x %d energies: max rel err %.2e p99.9 %.2e" % (p.nx, p.ny, p.n_energy, s["max"], s["p999"]), flush=True)
    report["configs"]["spectrum"] = {"size": "%dx%d x %d energies" % (p.nx, p.ny, p.n_energy), "planes": {"spectrum": dict(s, tol=1e-7, within_tol=s["max"] <= 1e-7)},
                                     "gpu_call_s": t_gpu, "cpu_ref_s": t_cpu}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)
    ok = all(c.get("status_mismatches", 0) == 0 and all(v["within_tol"] for v in c.get("planes", {}).values()) and c.get("within_tol", True)
             for c in report["configs"].values())
    print("PARITY", "OK" if ok else "FAILED", "->", args.out)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
