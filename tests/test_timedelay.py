"""geodesic_timedelay (sim5kerr-geod.c:559-731) and the Byrd & Friedman integrals behind it
(sim5elliptic.c:645-1139), plus the per-pixel DELAY plane of the eq-plane image.  Checkers: the unmodified reference
(oracle/_ref) and the committed fixture tests/golden/timedelay.npz generated from it."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi

TOL = 1e-9
OPS = {"C1": 0, "C2": 1, "C2_cos": 2, "Z2": 3, "Rm1": 4, "Rm2": 5, "R2": 6, "R_r0_re": 7, "R_r0_re_inf": 8, "R_r1_re": 9,
       "R_r2_re": 10, "T_m0": 11, "T_m2": 12, "R_r0_cc": 13, "R_r0_cc_inf": 14, "R_r1_cc": 15, "R_r2_cc": 16, "R_rp_cc2": 17}


def integral_args(rng, n):
    """Argument sets in the ranges geodesic_timedelay produces: four ordered real roots a>b>c>d with X>a, or two real roots
    a>b and a complex pair (u, v) with a<X1<X2; Jacobi arguments 0<u<K, 0<m<1; |alpha|>1 for the R-type integrals."""
    z = np.zeros(n)
    d = rng.uniform(-12, -1, n); c = d + rng.uniform(0.05, 2, n); b = c + rng.uniform(0.05, 3, n); a = b + rng.uniform(0.05, 4, n)
    X = a + np.exp(rng.uniform(-3, 6, n))
    u = rng.uniform(0.02, 1.4, n); m = rng.uniform(0.02, 0.98, n)
    al = rng.uniform(1.15, 6, n) * np.where(rng.uniform(size=n) < 0.5, -1, 1)
    za = rng.uniform(1.05, 4, n); zb = rng.uniform(0.1, 3, n)
    cu = rng.uniform(-3, 3, n); cv = rng.uniform(0.1, 4, n)
    a2 = np.maximum(cu, 0) + rng.uniform(1.5, 6, n); b2 = a2 - rng.uniform(0.1, 4, n)
    X1 = a2 + np.exp(rng.uniform(-3, 3, n)); X2 = X1 + np.exp(rng.uniform(-3, 5, n))
    p = rng.uniform(0.05, 1.9, n)
    ta2 = rng.uniform(0.05, 30, n); tb2 = rng.uniform(0.05, 1.0, n); tX = np.sqrt(tb2) * rng.uniform(0, 0.999, n)
    return {
        "C1": [u, m, z, z, z, z, z], "C2": [u, m, z, z, z, z, z], "C2_cos": [rng.uniform(-0.99, 0.99, n), m, z, z, z, z, z],
        "Z2": [za, zb, u, m, z, z, z], "Rm1": [al, u, m, z, z, z, z], "Rm2": [al, u, m, z, z, z, z], "R2": [al, u, m, z, z, z, z],
        "R_r0_re": [a, b, c, d, X, z, z], "R_r0_re_inf": [a, b, c, d, z, z, z], "R_r1_re": [a, b, c, d, X, z, z],
        "R_r2_re": [a, b, c, d, X, z, z], "T_m0": [ta2, tb2, tX, z, z, z, z], "T_m2": [ta2, tb2, tX, z, z, z, z],
        "R_r0_cc": [a2, b2, cu, cv, X1, z, z], "R_r0_cc_inf": [a2, b2, cu, cv, z, z, z], "R_r1_cc": [a2, b2, cu, cv, X1, X2, z],
        "R_r2_cc": [a2, b2, cu, cv, X1, X2, z], "R_rp_cc2": [a2, b2, cu, cv, X1, X2, p],
    }


def delay_args(rng, n):
    alpha = rng.uniform(-14, 14, n); beta = rng.uniform(-14, 14, n)
    ra = rng.uniform(30, 2000, n); rb = rng.uniform(8, 25, n)
    return alpha, beta, ra, rb


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def cpu_integral(lib, fn, op, v):
    n = v[0].size
    out = np.empty(n)
    f = getattr(lib, fn)
    f.restype = None
    f(C.c_int(op), C.c_long(n), *[_dp(np.ascontiguousarray(x)) for x in v], _dp(out))
    return out


def cpu_timedelay(lib, fn, incl, a, alpha, beta, ra, rb):
    out = np.empty(alpha.size)
    f = getattr(lib, fn)
    f.restype = None
    f(C.c_long(alpha.size), C.c_double(incl), C.c_double(a), _dp(alpha), _dp(beta), _dp(ra), _dp(rb), _dp(out))
    return out


def delay_params(n):
    p = abi.default_params(2, n)
    p.outputs = abi.OUT_R | abi.OUT_DELAY | abi.OUT_STATUS
    return p


def _check(got, ref, label, tol=TOL):
    s = H.err_summary(got, ref)
    assert s["max"] <= tol, "%s: max rel err %.3e (exact %.4f)" % (label, s["max"], s["exact"])
    return s


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
def test_hostsim_integrals_and_timedelay_against_reference():
    ref, hs = H.load_ref(), H.load_hostsim()
    rng = np.random.default_rng(7)
    for name, v in integral_args(rng, 1500).items():
        want = cpu_integral(ref, "ref_batch_integral", OPS[name], v)
        assert np.isfinite(want).mean() > 0.98, name
        _check(cpu_integral(hs, "hs_batch_integral", OPS[name], v), want, name)
    for a, ideg in ((0.9, 60.0), (0.998, 75.0), (0.3, 20.0)):
        al, be, ra, rb = delay_args(rng, 1500)
        want = cpu_timedelay(ref, "ref_batch_timedelay", abi.deg2rad(ideg), a, al, be, ra, rb)
        got = cpu_timedelay(hs, "hs_batch_timedelay", abi.deg2rad(ideg), a, al, be, ra, rb)
        assert np.isfinite(want).sum() > 1000
        _check(got, want, "timedelay a=%g" % a)


def test_hostsim_against_golden():
    g = H.golden("timedelay.npz")
    hs = H.load_hostsim()
    for name in OPS:
        v = [g["%s_v%d" % (name, k)] for k in range(7)]
        _check(cpu_integral(hs, "hs_batch_integral", OPS[name], v), g[name], name)
    got = cpu_timedelay(hs, "hs_batch_timedelay", float(g["incl"]), float(g["spin"]), g["alpha"], g["beta"], g["ra"], g["rb"])
    _check(got, g["delay"], "timedelay golden")
    p = delay_params(48)
    img, _, _ = H.run_hostsim(p)
    assert np.array_equal(img["status"], g["img_status"])
    _check(img["delay"], g["img_delay"], "delay plane golden")


def test_timedelay_physics():
    """Size-independent properties: the delay is positive and additive along the ray, and far from the hole it approaches
    the flat-space light travel time (plus the Shapiro logarithm)."""
    hs = H.load_hostsim()
    rng = np.random.default_rng(11)
    n = 400
    al = rng.uniform(-6, 6, n); be = rng.uniform(2, 8, n)
    r1 = np.full(n, 4000.0); r2 = np.full(n, 900.0); r3 = np.full(n, 300.0)
    incl, a = abi.deg2rad(40.0), 0.7
    t12 = cpu_timedelay(hs, "hs_batch_timedelay", incl, a, al, be, r1, r2)
    t23 = cpu_timedelay(hs, "hs_batch_timedelay", incl, a, al, be, r2, r3)
    t13 = cpu_timedelay(hs, "hs_batch_timedelay", incl, a, al, be, r1, r3)
    ok = np.isfinite(t12) & np.isfinite(t23) & np.isfinite(t13)
    assert ok.sum() > 300 and np.all(t12[ok] > 0)
    assert np.max(np.abs(t12[ok] + t23[ok] - t13[ok]) / t13[ok]) < 1e-9
    flat = (4000.0 - 900.0) + 2.0 * np.log(4000.0 / 900.0)
    assert np.max(np.abs(t12[ok] - flat)) < 0.2


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
def test_hostsim_delay_plane_against_reference():
    p = delay_params(96)
    got, _, _ = H.run_hostsim(p)
    ref, st, _ = H.run_ref(p)
    rep = H.assert_image_parity(got.arrays, ref.arrays, label="delay plane")
    assert rep["delay"]["n"] == 96 * 96 and np.isfinite(ref["delay"]).all() is not None
    hit = (ref["status"] & 31) <= 1
    assert np.nanmin(ref["delay"][hit]) > 900.0          # every hit is ~1000 M from the reference sphere


@pytest.mark.gpu
def test_gpu_against_golden(gpu_api):
    L = gpu_api.lib()
    L.sim5_batch_integral.restype = C.c_int
    L.sim5_batch_timedelay.restype = C.c_int
    g = H.golden("timedelay.npz")
    for name in OPS:
        v = [np.ascontiguousarray(g["%s_v%d" % (name, k)]) for k in range(7)]
        n = v[0].size
        tab = (C.POINTER(C.c_double) * 7)(*[_dp(x) for x in v])
        out = np.empty(n)
        gpu_api.check(L.sim5_batch_integral(C.c_int(OPS[name]), C.c_int64(n), tab, _dp(out)), "sim5_batch_integral")
        _check(out, g[name], name)
    al, be, ra, rb = (np.ascontiguousarray(g[k]) for k in ("alpha", "beta", "ra", "rb"))
    out = np.empty(al.size)
    gpu_api.check(L.sim5_batch_timedelay(C.c_int64(al.size), C.c_double(float(g["incl"])), C.c_double(float(g["spin"])),
                                         _dp(al), _dp(be), _dp(ra), _dp(rb), _dp(out)), "sim5_batch_timedelay")
    _check(out, g["delay"], "timedelay golden")
    p = delay_params(48)
    img, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    assert np.array_equal(img["status"], g["img_status"])
    _check(img["delay"], g["img_delay"], "delay plane golden")


@pytest.mark.gpu
@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref did not travel")
def test_gpu_delay_plane_against_reference(gpu_api):
    for n, with_phi in ((256, False), (128, True)):
        p = delay_params(n)
        if with_phi:
            p.outputs |= abi.OUT_PHI | abi.OUT_G | abi.OUT_FLUX
        got, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
        ref, rst, _ = H.run_ref(p)
        rep = H.assert_image_parity(got.arrays, ref.arrays, label="delay plane %d" % n)
        print("delay plane %d^2:" % n, {k: "max %.2e exact %.4f" % (v["max"], v["exact"]) for k, v in rep.items()})


@pytest.mark.gpu
def test_scalar_timedelay_entry(gpu_api):
    """geodesic_timedelay through the scalar sim5lib.h API == the batched entry."""
    L = gpu_api.lib()
    g = H.golden("timedelay.npz")
    L.geodesic_init_inf.restype = C.c_int
    L.geodesic_init_inf.argtypes = [C.c_double] * 4 + [C.c_void_p, C.POINTER(C.c_int)]
    L.geodesic_P_int.restype = C.c_double
    L.geodesic_P_int.argtypes = [C.c_void_p, C.c_double, C.c_int]
    L.geodesic_timedelay.restype = C.c_double
    L.geodesic_timedelay.argtypes = [C.c_void_p] + [C.c_double] * 6
    for i in range(12):
        buf = (C.c_char * 240)()
        e = C.c_int(0)
        ok = L.geodesic_init_inf(float(g["incl"]), float(g["spin"]), float(g["alpha"][i]), float(g["beta"][i]), buf, C.byref(e))
        if not ok:
            assert np.isnan(g["delay"][i])
            continue
        Pa = L.geodesic_P_int(buf, float(g["ra"][i]), 0)
        Pb = L.geodesic_P_int(buf, float(g["rb"][i]), 0)
        t = L.geodesic_timedelay(buf, Pa, 0.0, 0.0, Pb, 0.0, 0.0)
        assert H.rel_err(np.array([t]), np.array([g["delay"][i]]))[0] <= TOL
