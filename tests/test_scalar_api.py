"""The scalar sim5lib.h API of libsim5b200.so (one-thread device launches) against the unmodified reference (oracle/_ref) for the
entry points no image mode exercises: geodesic_init_src (sim5kerr-geod.c:105-173; the reference's own round-trip check is
sim5unittests.c:171-255), polarization_vector / polarization_constant / polarization_constant_infinity (sim5polarization.c:13-105,
144-168, 248-268; invariants of sim5unittests.c:137-154: k.f <= 1e-9, f.f = 1, kappa conserved), tetrad_zamo,
kerr_metric_contravariant, flat_metric, flat_connection, kerr_connection (sim5kerr.c:30-48, 105-131, 198-316, 677-710) and the
RTOPT_FLAT stepper (sim5raytrace.c:43-245).  Both libraries are called with IDENTICAL inputs through identical ctypes prototypes."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi
from tools_golden import geodesic_struct_dtype

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref did not travel")

D4 = C.c_double * 4


class Cplx(C.Structure):                       # `double complex` and struct {double, double} share the x86-64 SysV calling convention
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class Metric(C.Structure):                     # sim5kerr.h:18-25
    _fields_ = [(k, C.c_double) for k in ("a", "r", "m", "g00", "g11", "g22", "g33", "g03")]


class Tetrad(C.Structure):                     # sim5kerr.h:27-30
    _fields_ = [("e", (C.c_double * 4) * 4), ("metric", Metric)]


class RayData(C.Structure):                    # sim5raytrace.h:21-43
    _fields_ = [("opt_gr", C.c_int), ("opt_pol", C.c_int), ("step_epsilon", C.c_double), ("bh_spin", C.c_double), ("E", C.c_double),
                ("Q", C.c_double), ("WP", Cplx), ("pass_", C.c_int), ("refines", C.c_int), ("dk", D4), ("df", D4), ("kt", C.c_double),
                ("error", C.c_float)]


def proto(L):
    """the reference's prototypes (the same in both libraries)"""
    d, i, vp = C.c_double, C.c_int, C.c_void_p
    sig = {
        "geodesic_init_inf": (i, [d, d, d, d, vp, C.POINTER(i)]),
        "geodesic_init_src": (i, [d, d, d, D4, i, vp, C.POINTER(i)]),
        "geodesic_find_midplane_crossing": (d, [vp, i]),
        "geodesic_position_rad": (d, [vp, d]),
        "geodesic_position_pol": (d, [vp, d]),
        "geodesic_P_int": (d, [vp, d, i]),
        "geodesic_momentum": (None, [vp, d, d, d, D4]),
        "photon_momentum": (None, [d, d, d, d, d, d, d, D4]),
        "photon_carter_const": (d, [D4, C.POINTER(Metric)]),
        "kerr_metric": (None, [d, d, d, C.POINTER(Metric)]),
        "kerr_metric_contravariant": (None, [d, d, d, C.POINTER(Metric)]),
        "flat_metric": (None, [d, d, C.POINTER(Metric)]),
        "kerr_connection": (None, [d, d, d, vp]),
        "flat_connection": (None, [d, d, vp]),
        "dotprod": (d, [D4, D4, C.POINTER(Metric)]),
        "vector_norm_to": (None, [D4, d, C.POINTER(Metric)]),
        "tetrad_zamo": (None, [C.POINTER(Metric), C.POINTER(Tetrad)]),
        "tetrad_azimuthal": (None, [C.POINTER(Metric), d, C.POINTER(Tetrad)]),
        "on2bl": (None, [D4, D4, C.POINTER(Tetrad)]),
        "bl2on": (None, [D4, D4, C.POINTER(Tetrad)]),
        "OmegaK": (d, [d, d]),
        "polarization_vector": (None, [D4, Cplx, C.POINTER(Metric), D4]),
        "polarization_constant": (Cplx, [D4, D4, C.POINTER(Metric)]),
        "polarization_constant_infinity": (Cplx, [d, d, d, d]),
        "polarization_angle_rotation": (d, [d, d, d, d, Cplx]),
        "raytrace_prepare": (None, [d, D4, D4, d, i, C.POINTER(RayData)]),
        "raytrace": (None, [D4, D4, C.POINTER(d), C.POINTER(RayData)]),
        "raytrace_error": (d, [D4, D4, C.POINTER(RayData)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


@pytest.fixture(scope="module")
def libs(gpu_api):
    mine = proto(gpu_api.lib())
    ref = proto(C.CDLL(H.REF_SO)) if H.have_ref() else None
    return mine, ref


def close(a, b, tol=1e-9, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(H.rel_err(a, b, floor) <= tol))


def geod(buf):
    return np.frombuffer(buf, dtype=geodesic_struct_dtype(), count=1)[0]


@needs_ref
def test_geodesic_init_src_against_reference_and_roundtrip(libs):
    """sim5unittests.c:171-255 with a valid observer (its own inclination of 170 deg is rejected by geodesic_init_inf, so the
    reference's loop never runs): a geodesic set up from infinity, re-initialised from its equatorial crossing point and local
    momentum, must come back with the same constants of motion and the same observer -- and our geodesic_init_src must return the
    struct the reference returns for the same (a, r, m, k, ppc)."""
    mine, ref = libs
    a, inc, rmax, N = 0.7, abi.deg2rad(60.0), 14.0, 12
    tested = back = 0
    for y in range(N):
        for x in range(N):
            alpha = ((x + .5) / N - 0.5) * 2.0 * rmax
            beta = ((y + .5) / N - 0.5) * 2.0 * rmax
            g1 = (C.c_char * 240)()
            e = C.c_int(0)
            if not mine.geodesic_init_inf(inc, a, alpha, beta, g1, C.byref(e)):
                continue
            P = mine.geodesic_find_midplane_crossing(g1, 0)
            if np.isnan(P):
                continue
            G1 = geod(g1)
            pa = int(P > G1["Rpc"])
            r = mine.geodesic_position_rad(g1, P)
            if np.isnan(r) or r < 1.0 + np.sqrt(1.0 - a * a):
                continue
            k = D4()
            mine.photon_momentum(a, r, 0.0, float(G1["l"]), float(G1["q"]), -1.0 if pa else +1.0, -1.0, k)
            g2, g3 = (C.c_char * 240)(), (C.c_char * 240)()
            e2, e3 = C.c_int(0), C.c_int(0)
            ok2 = mine.geodesic_init_src(a, r, 0.0, k, pa, g2, C.byref(e2))
            ok3 = ref.geodesic_init_src(a, r, 0.0, k, pa, g3, C.byref(e3))
            assert ok2 == ok3 and e2.value == e3.value, (alpha, beta, ok2, ok3, e2.value, e3.value)
            if not ok2:
                continue
            G2, G3 = geod(g2), geod(g3)
            assert G2["type"] == G3["type"] and G2["nrr"] == G3["nrr"]
            for f in ("a", "l", "q", "m2p", "m2m", "mm", "mK", "rp", "Rpc", "Tpp", "Tip", "cos_i", "incl", "alpha", "beta"):
                assert close(G2[f], G3[f], 1e-9) or (np.isnan(G2[f]) and np.isnan(G3[f])), (f, alpha, beta, float(G2[f]), float(G3[f]))
            for f in ("r1", "r2", "r3", "r4"):
                assert close(G2[f], G3[f], 1e-9, floor=1e-6), (f, G2[f], G3[f])
            tested += 1
            # the round trip (reference bar: 1e-5 on cos_i); l and q come back through photon_motion_constants
            assert abs(G2["l"] - G1["l"]) <= 1e-9 * max(1.0, abs(G1["l"])) and abs(G2["q"] - G1["q"]) <= 1e-9 * max(1.0, abs(G1["q"]))
            if not np.isnan(G2["cos_i"]):
                assert abs(G2["cos_i"] - G1["cos_i"]) <= 1e-5, (alpha, beta, float(G2["cos_i"]), float(G1["cos_i"]))
                assert abs(G2["alpha"] - G1["alpha"]) <= 1e-4 * max(1.0, abs(alpha)) and abs(G2["beta"] - G1["beta"]) <= 1e-4 * max(1.0, abs(beta))
                back += 1
    assert tested >= 60 and back >= 60, (tested, back)
    print("geodesic_init_src: %d structs equal to the reference's, %d round trips back to the observer" % (tested, back))


def _points_on_geodesics(mine, a, inc, n, seed):
    """(metric, k) at two positions of each of n geodesics from infinity: analytic positions and momenta (geodesic_momentum)"""
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        alpha, beta = rng.uniform(-9, 9), rng.uniform(-9, 9)
        g = (C.c_char * 240)()
        e = C.c_int(0)
        if not mine.geodesic_init_inf(inc, a, alpha, beta, g, C.byref(e)):
            continue
        G = geod(g)
        if G["type"] != 40 or not (G["rp"] > 2.2):          # RR geodesics that stay outside the horizon
            continue
        pts = []
        for r in (40.0, float(G["rp"]) * 1.3):
            P = mine.geodesic_P_int(g, r, 0)
            m = mine.geodesic_position_pol(g, P)
            k = D4()
            mine.geodesic_momentum(g, P, r, m, k)
            if np.isnan(P) or np.isnan(m) or any(np.isnan(v) for v in k):
                pts = []
                break
            M = Metric()
            mine.kerr_metric(a, r, m, C.byref(M))
            pts.append((M, k, r, m))
        if len(pts) == 2:
            out.append((alpha, beta, pts))
    return out


def test_polarization_vector_walker_penrose_invariants(libs):
    """sim5unittests.c:113-154 on analytic geodesics: f built from kappa at a second point of the same geodesic satisfies
    k.f = 0 (<= 1e-9), f.f = 1 and reproduces kappa; with oracle/_ref present, polarization_vector / polarization_constant /
    polarization_constant_infinity / polarization_angle_rotation return the reference's numbers for the same inputs."""
    mine, ref = libs
    a, inc = 0.94, abi.deg2rad(75.0)
    worst_kf = worst_ff = worst_wp = 0.0
    n_ref = 0
    for alpha, beta, pts in _points_on_geodesics(mine, a, inc, 24, 3):
        (M1, k1, r1, m1), (M2, k2, r2, m2) = pts
        # initial polarization vector: in the frame of a ZAMO, perpendicular to k (sim5unittests.c:90-101)
        t = Tetrad()
        mine.tetrad_zamo(C.byref(M1), C.byref(t))
        kl, fl, f1 = D4(), D4(), D4()
        mine.bl2on(k1, kl, C.byref(t))
        fl[0], fl[1], fl[2], fl[3] = 0.0, -kl[2], kl[1], 0.0
        mine.on2bl(fl, f1, C.byref(t))
        mine.vector_norm_to(f1, 1.0, C.byref(M1))
        assert abs(mine.dotprod(k1, f1, C.byref(M1))) <= 1e-9 and abs(mine.dotprod(f1, f1, C.byref(M1)) - 1.0) <= 1e-9
        wp1 = mine.polarization_constant(k1, f1, C.byref(M1))
        f2 = D4()
        mine.polarization_vector(k2, wp1, C.byref(M2), f2)
        wp2 = mine.polarization_constant(k2, f2, C.byref(M2))
        kf = abs(mine.dotprod(k2, f2, C.byref(M2)))
        ff = abs(mine.dotprod(f2, f2, C.byref(M2)) - 1.0)
        wp = abs(complex(wp2.re, wp2.im) - complex(wp1.re, wp1.im)) / abs(complex(wp1.re, wp1.im))
        worst_kf, worst_ff, worst_wp = max(worst_kf, kf), max(worst_ff, ff), max(worst_wp, wp)
        assert kf <= 1e-9 and ff <= 1e-9 and wp <= 1e-9, (alpha, beta, kf, ff, wp)
        if ref is not None:
            f2r = D4()
            ref.polarization_vector(k2, wp1, C.byref(M2), f2r)
            assert close(list(f2), list(f2r), 1e-9, floor=1e-6), (list(f2), list(f2r))
            wr = ref.polarization_constant(k1, f1, C.byref(M1))
            assert close([wp1.re, wp1.im], [wr.re, wr.im], 1e-9, floor=1e-6)
            ci, cr = mine.polarization_constant_infinity(a, alpha, beta, inc), ref.polarization_constant_infinity(a, alpha, beta, inc)
            assert close([ci.re, ci.im], [cr.re, cr.im], 1e-9, floor=1e-6)
            xa, xr = mine.polarization_angle_rotation(a, inc, alpha, beta, wp1), ref.polarization_angle_rotation(a, inc, alpha, beta, wp1)
            assert abs(xa - xr) <= 1e-7 * max(1.0, abs(xr))
            n_ref += 1
    print("polarization_vector on 24 geodesics: max k.f %.2e, |f.f - 1| %.2e, kappa drift %.2e; %d compared with the reference" % (worst_kf, worst_ff, worst_wp, n_ref))


@needs_ref
def test_metric_frames_and_connections_against_reference(libs):
    mine, ref = libs
    rng = np.random.default_rng(11)
    for _ in range(40):
        a, r, m = rng.uniform(0.0, 0.998), float(np.exp(rng.uniform(np.log(2.1), np.log(200.0)))), rng.uniform(-0.95, 0.95)
        for fn, args in (("kerr_metric", (a, r, m)), ("kerr_metric_contravariant", (a, r, m)), ("flat_metric", (r, m))):
            A, B = Metric(), Metric()
            getattr(mine, fn)(*args, C.byref(A))
            getattr(ref, fn)(*args, C.byref(B))
            va = [getattr(A, k) for k, _ in Metric._fields_]
            vb = [getattr(B, k) for k, _ in Metric._fields_]
            assert close(va, vb, 1e-12, floor=1e-12), (fn, args, va, vb)
        for fn, args in (("kerr_connection", (a, r, m)), ("flat_connection", (r, m))):
            GA, GB = (C.c_double * 64)(), (C.c_double * 64)()
            for i in range(64):
                GA[i] = GB[i] = 0.0
            getattr(mine, fn)(*args, C.cast(GA, C.c_void_p))
            getattr(ref, fn)(*args, C.cast(GB, C.c_void_p))
            assert close(list(GA), list(GB), 1e-12, floor=1e-12), (fn, args)
        M = Metric()
        ref.kerr_metric(a, r, m, C.byref(M))
        TA, TB = Tetrad(), Tetrad()
        mine.tetrad_zamo(C.byref(M), C.byref(TA))
        ref.tetrad_zamo(C.byref(M), C.byref(TB))
        ea = [TA.e[i][j] for i in range(4) for j in range(4)]
        eb = [TB.e[i][j] for i in range(4) for j in range(4)]
        assert close(ea, eb, 1e-12, floor=1e-12), ("tetrad_zamo", a, r, m)


@needs_ref
def test_raytrace_flat_space_option_against_reference(libs):
    """RTOPT_FLAT (sim5raytrace.c:64-70, 127-128): the stepper with the flat metric and connection.  Same step sequence (dl, pass
    counter, float step error) and the same x, k after every one of 120 steps as the reference, for 4 rays."""
    mine, ref = libs
    RTOPT_FLAT = 1
    rng = np.random.default_rng(5)
    for ray in range(4):
        r0, m0 = rng.uniform(20, 60), rng.uniform(-0.6, 0.6)
        M = Metric()
        ref.flat_metric(r0, m0, C.byref(M))
        # a null vector of the flat metric pointing inwards
        kr, km, kp = -1.0, rng.uniform(-0.01, 0.01), rng.uniform(-0.002, 0.002)
        spatial = M.g11 * kr * kr + M.g22 * km * km + M.g33 * kp * kp
        kt = np.sqrt(spatial / -M.g00)
        state = []
        for L in (mine, ref):
            x, k = D4(0.0, r0, m0, 0.0), D4(kt, kr, km, kp)
            rtd = RayData()
            L.raytrace_prepare(0.0, x, k, 0.01, RTOPT_FLAT, C.byref(rtd))
            assert rtd.opt_gr == 0
            trace = []
            for _ in range(120):
                dl = C.c_double(1e9)
                L.raytrace(x, k, C.byref(dl), C.byref(rtd))
                trace.append((dl.value, rtd.pass_, rtd.error) + tuple(x) + tuple(k))
            state.append((np.array(trace), L.raytrace_error(x, k, C.byref(rtd))))
        A, B = state[0][0], state[1][0]
        assert np.array_equal(A[:, 1], B[:, 1]), "pass counters differ"
        assert close(A[:, 0], B[:, 0], 1e-9), "step sizes differ"
        assert close(A[:, 3:], B[:, 3:], 1e-9, floor=1e-9), "positions / momenta differ"
        assert np.all(np.abs(A[:, 2] - B[:, 2]) <= 1e-9)
        assert abs(state[0][1] - state[1][1]) <= 1e-9
