#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libsim5ref.so, built from /root/reference by oracle/Makefile) in this container.

  python tools/make_golden.py

The reference's own tests hold no golden vectors, so these files are the pin:
outputs of the reference itself on seeded / deterministic inputs.  /root/reference does not exist on the
GPU box; the fixtures (and the built oracle/_ref) travel instead.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from sim5_b200 import abi  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def geodesic_struct_dtype():
    return np.dtype([("a", "f8"), ("alpha", "f8"), ("beta", "f8"), ("incl", "f8"), ("cos_i", "f8"), ("l", "f8"), ("q", "f8"),
                     ("r1", "f8", 2), ("r2", "f8", 2), ("r3", "f8", 2), ("r4", "f8", 2), ("nrr", "i4"), ("type", "i4"),
                     ("m2p", "f8"), ("m2m", "f8"), ("mm", "f8"), ("mK", "f8"), ("rp", "f8"), ("dmdp_inf", "f8"),
                     ("Rpc", "f8"), ("Tpp", "f8"), ("Tip", "f8"), ("k", "f8", 4), ("p", "f8")])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = H.load_ref()
    rng = np.random.default_rng(20261017)

    # ---- images of the BASELINE configs at fixture size
    for cfg, n in ((1, 64), (2, 64), (3, 64), (4, 16), (7, 48)):
        p = abi.default_params(cfg, n)
        if cfg == 4:
            p.outputs |= abi.OUT_QERR
        if cfg == 3:
            p.outputs |= abi.OUT_MUE
        pl, st, _ = H.run_ref(p)
        np.savez_compressed(os.path.join(OUT, "image_cfg%d_%d.npz" % (cfg, n)),
                            class_count=np.array(list(st.class_count)), gtype_count=np.array(list(st.gtype_count)),
                            **pl.arrays)
    # a non-square, ragged image (nx not a multiple of 32) incl. the beta==0 row and both hemispheres
    p = abi.default_params(2, 50, 37)
    pl, st, _ = H.run_ref(p)
    np.savez_compressed(os.path.join(OUT, "image_cfg2_50x37.npz"), **pl.arrays)

    # the surface finder on a second disk: truncated at R = 8 (rays through the hole reach the equatorial plane), a = 0.5, i = 30 deg
    p = abi.default_params(7, 40, 28)
    p.bh_spin, p.incl, p.surf_rin, p.surf_hr, p.rmax = 0.5, abi.deg2rad(30.0), 8.0, 0.2, 12.0
    pl, st, _ = H.run_ref(p)
    np.savez_compressed(os.path.join(OUT, "image_surface_rin8_40x28.npz"), class_count=np.array(list(st.class_count)), **pl.arrays)

    # ---- geodesic_timedelay: the B&F integrals behind it, delays between two radii on random geodesics, the DELAY plane
    import test_timedelay as TD
    trng = np.random.default_rng(20261018)
    out = {}
    for name, v in TD.integral_args(trng, 400).items():
        for k in range(7):
            out["%s_v%d" % (name, k)] = v[k]
        out[name] = TD.cpu_integral(ref, "ref_batch_integral", TD.OPS[name], v)
    al, be, ra, rb = TD.delay_args(trng, 600)
    out.update(alpha=al, beta=be, ra=ra, rb=rb, incl=abi.deg2rad(60.0), spin=0.9)
    out["delay"] = TD.cpu_timedelay(ref, "ref_batch_timedelay", abi.deg2rad(60.0), 0.9, al, be, ra, rb)
    p = TD.delay_params(48)
    pl, st, _ = H.run_ref(p)
    out.update(img_delay=pl["delay"], img_status=pl["status"])
    np.savez_compressed(os.path.join(OUT, "timedelay.npz"), **out)

    # ---- thermal spectrum (SPECTRUM preset shrunk), incl. a partial spectrum of an interleaved split
    p = abi.default_params(6, 96)
    spec, _ = H.run_spectrum("ref", p)
    p.split_count, p.split_index, p.split_rows = 3, 1, 8
    part, _ = H.run_spectrum("ref", p)
    np.savez_compressed(os.path.join(OUT, "spectrum_cfg6_96.npz"), spectrum=spec, part_3_1_8=part, energies=np.array(abi.spectrum_energies(abi.default_params(6, 96))))

    # ---- histogram lattice (cfg 5 shrunk)
    p = abi.default_params(5, 48)
    p.n_spin, p.n_incl, p.n_bins = 3, 2, 32
    hist = np.zeros(p.n_spin * p.n_incl * p.n_bins)
    dt = ref.ref_trace_histogram(C.byref(p), hist.ctypes.data_as(C.POINTER(C.c_double)), 0, 1)
    assert dt >= 0
    np.savez_compressed(os.path.join(OUT, "hist_cfg5_3x2x32_48.npz"), hist=hist)

    # ---- Carlson / Jacobi vectors
    n = 3000
    x = np.exp(rng.uniform(-8, 8, n)); y = np.exp(rng.uniform(-8, 8, n)); z = np.exp(rng.uniform(-8, 8, n))
    x[:300] = 0.0                                    # at most one argument may vanish
    pj = np.exp(rng.uniform(-6, 6, n)) * np.where(rng.uniform(size=n) < 0.3, -1.0, 1.0)
    yc = y * np.where(rng.uniform(size=n) < 0.3, -1.0, 1.0)
    u = rng.uniform(0, 3.0, n); m = rng.uniform(0, 1, n)
    m[:50] = rng.uniform(1.0, 1.5, 50)               # the m > 1 branch of sncndn
    sn, cn, dn = H.batch_call(ref, "ref_batch_sncndn", [u, m], nout=3)
    np.savez_compressed(os.path.join(OUT, "elliptic.npz"), x=x, y=y, z=z, p=pj, yc=yc, u=u, m=m,
                        rf=H.batch_call(ref, "ref_batch_rf", [x, y, z]),
                        rd=H.batch_call(ref, "ref_batch_rd", [x, y, np.maximum(z, 1e-3)]),
                        rc=H.batch_call(ref, "ref_batch_rc", [x, yc]),
                        rj=H.batch_call(ref, "ref_batch_rj", [x, y, z, pj]),
                        sn=sn, cn=cn, dn=dn)

    # ---- libm: glibc values (what the reference sees) for the argument ranges of the ray path
    n = 4000
    args = {
        "sin": rng.uniform(-7, 7, n), "cos": rng.uniform(-7, 7, n), "log": np.exp(rng.uniform(-12, 12, n)),
        "atan2": rng.normal(size=n), "acos": rng.uniform(-1, 1, n), "asin": rng.uniform(-1, 1, n),
        "atan": rng.normal(size=n) * 4, "pow_third": np.exp(rng.uniform(-20, 40, n)), "pow_1p5": np.exp(rng.uniform(-2, 8, n)),
        "pow_4": rng.uniform(0, 2, n),
    }
    b2 = rng.normal(size=n)
    opc = {"sin": 0, "cos": 1, "log": 2, "atan2": 3, "acos": 4, "asin": 5, "atan": 6, "pow_third": 7, "pow_1p5": 8, "pow_4": 9}
    out = {"atan2_x": b2}
    for k, a in args.items():
        out[k + "_in"] = a
        out[k + "_glibc"] = H.batch_call(ref, "ref_batch_libm", [a, b2 if k == "atan2" else np.zeros(n)], extra=(C.c_int(opc[k]),))
    np.savez_compressed(os.path.join(OUT, "libm_glibc.npz"), **out)

    # ---- geodesic structs from geodesic_init_inf (pins R_roots, the long-double T_roots, Rpc, Tpp, Tip)
    gd_t = geodesic_struct_dtype()
    assert gd_t.itemsize == 240
    n = 1500
    spin = rng.choice([0.0, 0.3, 0.9, 0.94, 0.998], n)
    incl = np.deg2rad(rng.uniform(2, 88, n))
    alpha = rng.uniform(-25, 25, n); beta = rng.uniform(-25, 25, n)
    beta[:20] = 0.0
    g = np.zeros(n, dtype=gd_t); ok = np.zeros(n, dtype=np.int32); err = np.zeros(n, dtype=np.int32)
    ref.geodesic_init_inf.restype = C.c_int
    ref.geodesic_init_inf.argtypes = [C.c_double] * 4 + [C.c_void_p, C.POINTER(C.c_int)]
    ref.geodesic_find_midplane_crossing.restype = C.c_double
    ref.geodesic_find_midplane_crossing.argtypes = [C.c_void_p, C.c_int]
    ref.geodesic_position_rad.restype = C.c_double
    ref.geodesic_position_rad.argtypes = [C.c_void_p, C.c_double]
    P0 = np.full(n, np.nan); r0 = np.full(n, np.nan)
    for i in range(n):
        e = C.c_int(0)
        buf = (C.c_char * 240)()
        ok[i] = ref.geodesic_init_inf(incl[i], spin[i], alpha[i], beta[i], buf, C.byref(e))
        err[i] = e.value
        if ok[i]:
            g[i] = np.frombuffer(buf, dtype=gd_t, count=1)[0]
            P0[i] = ref.geodesic_find_midplane_crossing(buf, 0)
            if not np.isnan(P0[i]):
                import contextlib
                r0[i] = ref.geodesic_position_rad(buf, P0[i])
    np.savez_compressed(os.path.join(OUT, "geodesic_init_inf.npz"), spin=spin, incl=incl, alpha=alpha, beta=beta,
                        ok=ok, err=err, g=g.view(np.uint8).reshape(n, 240), P0=P0, r0=r0)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-32s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
