"""Test-side helpers: load the checkers (oracle/_ref = unmodified reference, oracle/ = our C port),
allocate SoA planes, run images through any of the implementations, and compare results.

Only tests/, bench.py's cpu_baseline / reference arm and __graft_entry__.smoke() import this.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sim5_b200 import abi  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsim5ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libsim5oracle.so")

_NP = {C.c_double: np.float64, C.c_int32: np.int32, C.c_uint8: np.uint8}


def have_ref():
    return os.path.exists(REF_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)


_libs = {}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def load_ref():
    if "ref" not in _libs:
        lib = C.CDLL(REF_SO)
        lib.ref_trace_image.restype = C.c_double
        lib.ref_trace_image.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(abi.ImageOut), C.c_int, C.c_int,
                                        C.POINTER(abi.TraceStats)]
        lib.ref_trace_histogram.restype = C.c_double
        lib.ref_trace_histogram.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(C.c_double), C.c_int, C.c_int]
        lib.ref_trace_spectrum.restype = C.c_double
        lib.ref_trace_spectrum.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(C.c_double), C.c_int, C.c_int]
        lib.ref_r_ms.restype = C.c_double
        lib.ref_r_ms.argtypes = [C.c_double]
        lib.ref_r_bh.restype = C.c_double
        lib.ref_r_bh.argtypes = [C.c_double]
        lib.ref_max_threads.restype = C.c_int
        _libs["ref"] = lib
    return _libs["ref"]


def load_oracle():
    if "oracle" not in _libs:
        lib = C.CDLL(ORACLE_SO)
        lib.orc_trace_image.restype = C.c_double
        lib.orc_trace_image.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(abi.ImageOut), C.c_int,
                                        C.POINTER(abi.TraceStats)]
        lib.orc_trace_histogram.restype = C.c_double
        lib.orc_trace_histogram.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(C.c_double), C.c_int]
        lib.orc_trace_spectrum.restype = C.c_double
        lib.orc_trace_spectrum.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(C.c_double), C.c_int]
        _libs["oracle"] = lib
    return _libs["oracle"]


class Planes:
    """Host SoA output planes for one image + the ImageOut struct pointing at them."""

    def __init__(self, p, fill=None):
        n = p.nx * p.ny
        self.arrays = {}
        self.out = abi.ImageOut()
        for name, bit, ct in abi.PLANES:
            if p.outputs & bit:
                a = np.zeros(n, dtype=_NP[ct])
                if fill is not None:
                    a[...] = fill
                self.arrays[name] = a
                setattr(self.out, name, a.ctypes.data)
        self.shape = (p.ny, p.nx)

    def __getitem__(self, k):
        return self.arrays[k]

    def image(self, k):
        return self.arrays[k].reshape(self.shape)


def run_ref(p, nthreads=0, quiet=True):
    lib = load_ref()
    pl = Planes(p)
    st = abi.TraceStats()
    dt = lib.ref_trace_image(C.byref(p), C.byref(pl.out), nthreads, 1 if quiet else 0, C.byref(st))
    assert dt >= 0, "ref_trace_image failed (%r)" % dt
    return pl, st, dt


def run_oracle(p, nthreads=0):
    lib = load_oracle()
    pl = Planes(p)
    st = abi.TraceStats()
    dt = lib.orc_trace_image(C.byref(p), C.byref(pl.out), nthreads, C.byref(st))
    assert dt >= 0, "orc_trace_image failed (%r)" % dt
    return pl, st, dt


def batch_call(lib, name, ins, nout=1, extra=()):
    """Call an n-element SoA function `name(n?, in..., out...)` of the ref / oracle library."""
    ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
    n = ins[0].size
    outs = [np.empty(n, dtype=np.float64) for _ in range(nout)]
    fn = getattr(lib, name)
    fn.restype = None
    args = list(extra) + [C.c_long(n)] + [_dp(a) for a in ins] + [_dp(o) for o in outs]
    fn(*args)
    return outs[0] if nout == 1 else outs


def rel_err(x, ref, floor=0.0):
    """|x-ref| / max(|ref|, floor) with exact-equal (incl. both-NaN) counted as 0."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.maximum(np.abs(ref), floor if floor > 0 else np.finfo(np.float64).tiny)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(x - ref) / den
    same = (x == ref) | (np.isnan(x) & np.isnan(ref))
    e = np.where(same, 0.0, e)
    e = np.where(np.isnan(e), np.inf, e)
    return e


def err_summary(x, ref, floor=0.0):
    e = rel_err(x, ref, floor)
    if e.size == 0:
        return {"max": 0.0, "p999": 0.0, "exact": 1.0, "n": 0}
    return {"max": float(e.max()), "p999": float(np.quantile(e, 0.999)), "exact": float(np.mean(e == 0.0)),
            "n": int(e.size)}


# tolerances of BASELINE.json's north_star: bit-exact classification / termination flags;
# relative error <= 1e-9 on r, phi, g and <= 1e-7 on the polarization angle and flux.
TOL = {"r": 1e-9, "phi": 1e-9, "g": 1e-9, "flux": 1e-7, "chi": 1e-7, "delta": 1e-7, "mue": 1e-9,
       "intensity": 1e-7, "tau": 1e-7, "qerr": 1e-9, "height": 1e-9, "delay": 1e-9}
FLOOR = {"phi": 1.0, "chi": 1.0, "height": 1.0, "qerr": 1.0}
# |dphi| / max(|phi|, 1): phi ~ 1e-5 on the alpha ~ 0 column; the height r cos(theta) of a SURFACE hit passes through
# zero at the inner edge of the disk; qerr = raytrace_error() IS a relative error (the drift |Q - Q0| / Q0 of Carter's constant, 1e-10
# ... 1e-3: a difference of two nearly equal numbers), so it is held to an ABSOLUTE 1e-9


def assert_image_parity(got, ref, label="", tol=TOL):
    """got/ref: dicts of flat arrays (Planes.arrays or npz).  Returns the per-plane error summaries."""
    report = {}
    if "status" in ref:
        bad = int(np.sum(np.asarray(got["status"]) != np.asarray(ref["status"])))
        assert bad == 0, "%s: status byte differs on %d pixels" % (label, bad)
    if "steps" in ref:
        bad = int(np.sum(np.asarray(got["steps"]) != np.asarray(ref["steps"])))
        assert bad == 0, "%s: step count differs on %d rays" % (label, bad)
    for k in ref:
        if k in ("status", "steps", "class_count", "gtype_count") or k not in got:
            continue
        t = tol.get(k)
        s = err_summary(got[k], ref[k], FLOOR.get(k, 0.0))
        report[k] = s
        if t is not None:
            assert s["max"] <= t, "%s: %s max rel err %.3e > %.1e (p99.9 %.3e)" % (label, k, s["max"], t, s["p999"])
    return report


def load_hostsim():
    if "hostsim" not in _libs:
        lib = C.CDLL(os.path.join(ROOT, "tests", "_build", "libhostsim.so"))
        lib.hs_trace_image.restype = C.c_double
        lib.hs_trace_image.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(abi.ImageOut), C.c_int]
        lib.hs_trace_spectrum.restype = C.c_double
        lib.hs_trace_spectrum.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(C.c_double), C.c_int]
        _libs["hostsim"] = lib
    return _libs["hostsim"]


def run_hostsim(p, nthreads=0):
    pl = Planes(p)
    dt = load_hostsim().hs_trace_image(C.byref(p), C.byref(pl.out), nthreads)
    return pl, None, dt


def run_spectrum(which, p, nthreads=0):
    """SPECTRUM mode through one of the CPU implementations: 'ref' (unmodified reference + driver), 'oracle' (C restatement),
    'hostsim' (host instantiation of the device headers).  Returns (spectrum[n_energy], seconds)."""
    spec = np.zeros(p.n_energy)
    ptr = spec.ctypes.data_as(C.POINTER(C.c_double))
    if which == "ref":
        dt = load_ref().ref_trace_spectrum(C.byref(p), ptr, nthreads, 1)
    elif which == "oracle":
        dt = load_oracle().orc_trace_spectrum(C.byref(p), ptr, nthreads)
    else:
        dt = load_hostsim().hs_trace_spectrum(C.byref(p), ptr, nthreads)
    assert dt >= 0, "%s spectrum failed (%r)" % (which, dt)
    return spec, dt


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


GOLDEN_IMAGES = (  # (file, cfg, nx, ny, extra output bits)
    ("image_cfg1_64.npz", 1, 64, 64, 0), ("image_cfg2_64.npz", 2, 64, 64, 0), ("image_cfg2_50x37.npz", 2, 50, 37, 0),
    ("image_cfg3_64.npz", 3, 64, 64, abi.OUT_MUE), ("image_cfg4_16.npz", 4, 16, 16, abi.OUT_QERR),
    ("image_cfg7_48.npz", 7, 48, 48, 0),
)


def golden_params(cfg, nx, ny, extra):
    p = abi.default_params(cfg, nx, ny)
    p.outputs |= extra
    return p
