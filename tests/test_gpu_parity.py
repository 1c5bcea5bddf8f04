"""GPU parity tests proper: everything goes through the C-ABI of libsim5b200.so (ctypes), on the device.
Checkers: the committed golden fixtures (from the unmodified reference) and, where it travelled with the
snapshot, oracle/_ref (the unmodified reference itself).  /root/reference is never read."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi

pytestmark = pytest.mark.gpu


def _gpu_planes(api, p):
    planes, st = api.trace_image(p, api.HostPlanes(p, pinned=True))
    return planes, st


@pytest.mark.parametrize("fname,cfg,nx,ny,extra", H.GOLDEN_IMAGES)
def test_images_against_golden(gpu_api, fname, cfg, nx, ny, extra):
    p = H.golden_params(cfg, nx, ny, extra)
    got, st = _gpu_planes(gpu_api, p)
    g = H.golden(fname)
    H.assert_image_parity(got.arrays, g, label=fname)
    if "class_count" in g:
        assert list(st.class_count) == list(g["class_count"])
        assert list(st.gtype_count) == list(g["gtype_count"])
    assert st.kernel_launches >= 1 and st.rays == nx * ny


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("cfg,n", [(1, 512), (2, 384), (3, 384), (4, 64), (7, 192)])
def test_images_against_reference(gpu_api, cfg, n):
    """SURVEY.md 8(d) configs at sizes the CPU reference finishes in seconds; cfg 1 at its full 512^2."""
    p = abi.default_params(cfg, n)
    got, st = _gpu_planes(gpu_api, p)
    ref, rst, _ = H.run_ref(p)
    rep = H.assert_image_parity(got.arrays, ref.arrays, label="cfg%d %d^2" % (cfg, n))
    assert list(st.class_count) == list(rst.class_count) and list(st.gtype_count) == list(rst.gtype_count)
    if cfg in (4, 7):
        assert st.total_steps == rst.total_steps
    print("cfg%d %dx%d vs reference:" % (cfg, n, n), {k: ("max %.2e p99.9 %.2e exact %.4f" % (v["max"], v["p999"], v["exact"])) for k, v in rep.items()})
    if cfg == 1:   # the reference's own statistics for BASELINE config 1
        assert list(st.class_count)[:5] == [241964, 372, 14512, 789, 4507]
        assert list(st.gtype_count)[1:4] == [213492, 48560, 92]


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref did not travel")
def test_edge_cases_against_reference(gpu_api):
    cases = []
    p = abi.default_params(1, 24); p.bh_spin = 1.5; cases.append(("spin>1", p))
    p = abi.default_params(1, 24); p.incl = 1.6; cases.append(("incl>pi/2", p))
    p = abi.default_params(1, 24); p.bh_spin = 0.0; p.rmax = 14.0; cases.append(("a=0", p))
    p = abi.default_params(2, 24); p.incl = abi.deg2rad(1.0); cases.append(("i=1deg", p))
    p = abi.default_params(2, 24); p.incl = abi.deg2rad(89.0); cases.append(("i=89deg", p))
    p = abi.default_params(2, 33, 17); p.max_order = 2; cases.append(("order2 ragged", p))
    p = abi.default_params(1, 1, 1); cases.append(("1x1", p))
    p = abi.default_params(1, 31, 64); p.r_emit_min = 3.0; cases.append(("r_emit_min", p))
    for label, p in cases:
        got, _ = _gpu_planes(gpu_api, p)
        ref, _, _ = H.run_ref(p)
        H.assert_image_parity(got.arrays, ref.arrays, label=label)


def test_row_ranges_and_interleaved_split_compose(gpu_api):
    """Tracing row ranges or interleaved row blocks (the multi-GPU split) gives exactly the rows of the full image."""
    p = abi.default_params(2, 128)
    full, _ = _gpu_planes(gpu_api, p)
    part = gpu_api.HostPlanes(p, pinned=True)
    for k in part.arrays:
        part[k][...] = 0
    for rb, re in ((0, 40), (40, 41), (41, 128)):
        p.row_begin, p.row_end = rb, re
        gpu_api.trace_image(p, part)
    for k in full.arrays:
        assert np.array_equal(full[k], part[k], equal_nan=True), k
    p = abi.default_params(2, 128)
    inter = gpu_api.HostPlanes(p, pinned=True)
    for k in inter.arrays:
        inter[k][...] = 0
    for r in range(4):
        p.split_count, p.split_index, p.split_rows = 4, r, 8
        _, st = gpu_api.trace_image(p, inter)
        assert st.rays == 128 * 128 // 4
    for k in full.arrays:
        assert np.array_equal(full[k], inter[k], equal_nan=True), k


def test_chunked_copy_overlap_equals_single_pass(gpu_api):
    """Host-plane calls trace in chunks and copy chunk k back under the kernels of chunk k+1: same planes and counters as the
    unchunked call, for contiguous rows, ragged row counts, row sub-ranges and the interleaved multi-GPU split."""
    L = gpu_api.lib()
    try:
        for cfg, nx, ny, split in ((2, 256, 256, 1), (2, 130, 77, 1), (3, 96, 160, 1), (2, 128, 256, 2), (4, 48, 40, 1), (1, 64, 192, 3)):
            p = abi.default_params(cfg, nx, ny)
            if split > 1:
                p.split_count, p.split_index, p.split_rows = split, 1, 8
            if cfg == 2 and split == 1 and nx == 130:
                p.row_begin, p.row_end = 5, 70
            p.flags = abi.FLAG_NO_OVERLAP
            L.sim5_set_chunk_rays(0)
            ref = gpu_api.HostPlanes(p, pinned=True)
            for k in ref.arrays:
                ref[k][...] = 0
            _, st0 = gpu_api.trace_image(p, ref)
            p.flags = 0
            L.sim5_set_chunk_rays(400)
            got = gpu_api.HostPlanes(p, pinned=True)
            for k in got.arrays:
                got[k][...] = 0
            _, st1 = gpu_api.trace_image(p, got)
            assert st1.kernel_launches > st0.kernel_launches, "the call was not chunked"
            assert list(st0.class_count) == list(st1.class_count) and st0.total_steps == st1.total_steps and st0.rays == st1.rays
            for k in ref.arrays:
                assert np.array_equal(ref[k], got[k], equal_nan=True), (cfg, nx, ny, split, k)
    finally:
        L.sim5_set_chunk_rays(0)


def test_empty_and_bad_parameters(gpu_api):
    L = gpu_api.lib()
    p = abi.default_params(1, 16)
    planes = gpu_api.HostPlanes(p, pinned=False)
    st = abi.TraceStats()
    p.row_begin, p.row_end = 5, 5            # empty row range: OK, nothing traced
    assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.OK and st.rays == 0
    p = abi.default_params(1, 16); p.nx = 0
    assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.ERR_BAD_PARAM
    p = abi.default_params(1, 16); p.struct_size = 8
    assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.ERR_BAD_PARAM
    p = abi.default_params(1, 16); p.outputs |= abi.OUT_PHI    # plane selected but NULL
    assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.ERR_NO_OUTPUT
    p = abi.default_params(1, 16); p.mode = 9
    assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.ERR_BAD_PARAM


def test_elliptic_batch_against_golden(gpu_api):
    g = H.golden("elliptic.npz")
    assert np.array_equal(gpu_api.batch_rf(g["x"], g["y"], g["z"]), g["rf"])
    assert np.array_equal(gpu_api.batch_rd(g["x"], g["y"], np.maximum(g["z"], 1e-3)), g["rd"])
    assert np.array_equal(gpu_api.batch_rc(g["x"], g["yc"]), g["rc"], equal_nan=True)
    assert np.array_equal(gpu_api.batch_rj(g["x"], g["y"], g["z"], g["p"]), g["rj"], equal_nan=True)
    sn, cn, dn = gpu_api.batch_sncndn(g["u"], g["m"])
    # sin/cos inside are ours (correctly rounded) vs glibc's: 1-ulp differences in ~0.1 % of calls
    for mine, ref in ((sn, g["sn"]), (cn, g["cn"]), (dn, g["dn"])):
        assert np.mean(mine == ref) > 0.99
        assert H.err_summary(mine, ref, floor=1e-3)["max"] < 1e-13


@pytest.mark.skipif(not H.have_oracle(), reason="oracle/libsim5oracle.so not built")
def test_fastfp_carlson_adversarial_operands(gpu_api):
    """The branch-free division / square-root bodies of R_F, R_C, R_J (fastfp.cuh) and their plain-operator fallback
    give the reference's bits on operands far outside the image's range: 2^-1000..2^1000, exact zeros, subnormals,
    equal arguments, negative p / y (Cauchy principal values).  Checker: the C restatement (bit-identical to the
    reference, tests/test_oracle.py)."""
    lib = H.load_oracle()
    rng = np.random.default_rng(7)
    n = 200000

    def wide(lo, hi):
        return np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(lo, hi, n))
    for lo, hi in ((-8, 8), (-90, 90), (-320, 320), (-1000, 1000)):
        x, y, z = wide(lo, hi), wide(lo, hi), wide(lo, hi)
        x[::7] = 0.0                                   # at most one zero argument is allowed
        y[3::11] = x[3::11]                            # equal arguments
        z[5::13] = 5e-324                              # subnormal
        assert np.array_equal(gpu_api.batch_rf(x, y, z), H.batch_call(lib, "orc_batch_rf", [x, y, z]), equal_nan=True), ("rf", lo, hi)
        yc = np.where(rng.random(n) < 0.3, -y, y)
        assert np.array_equal(gpu_api.batch_rc(x, yc), H.batch_call(lib, "orc_batch_rc", [x, yc]), equal_nan=True), ("rc", lo, hi)
        if hi <= 320:
            pp = np.where(rng.random(n) < 0.3, -wide(lo, hi), wide(lo, hi))
            assert np.array_equal(gpu_api.batch_rj(x, y, z, pp), H.batch_call(lib, "orc_batch_rj", [x, y, z, pp]), equal_nan=True), ("rj", lo, hi)
            zd = np.maximum(z, 1e-300)
            assert np.array_equal(gpu_api.batch_rd(x, y, zd), H.batch_call(lib, "orc_batch_rd", [x, y, zd]), equal_nan=True), ("rd", lo, hi)


def test_device_libm_equals_host_instantiation_and_glibc(gpu_api):
    """The sm_100a build of crmath.cuh gives the same bits as its host instantiation (so the CPU-side tests
    vouch for the device code) and matches glibc like the host build does."""
    hs = C.CDLL(os.path.join(H.ROOT, "tests", "_build", "libhostsim.so"))
    g = H.golden("libm_glibc.npz")
    ops = {"sin": 0, "cos": 1, "log": 2, "atan2": 3, "acos": 4, "asin": 5, "atan": 6, "pow_third": 7, "pow_1p5": 8, "pow_4": 9}
    for op, code in ops.items():
        a = g[op + "_in"]
        b = g["atan2_x"] if op == "atan2" else np.zeros_like(a)
        dev = gpu_api.batch_libm(op, a, b)
        host = H.batch_call(hs, "hs_batch_libm", [a, b], extra=(C.c_int(code),))
        if op == "pow_third":      # its cbrt()/log() seeds are platform libm calls; the rounded result still agrees
            assert np.mean(dev == host) > 0.9999
        else:
            assert np.array_equal(dev, host), op
        assert np.mean(dev == g[op + "_glibc"]) > 0.995, op


def test_scalar_api_geodesic_roundtrip(gpu_api):
    """The sim5lib.h scalar API (one-thread device launches) against the reference's structs: init_inf,
    crossing, position_rad -- the call sequence of examples/04 for single rays."""
    from tools_golden import geodesic_struct_dtype
    L = gpu_api.lib()
    L.geodesic_init_inf.restype = C.c_int
    L.geodesic_init_inf.argtypes = [C.c_double] * 4 + [C.c_void_p, C.POINTER(C.c_int)]
    L.geodesic_find_midplane_crossing.restype = C.c_double
    L.geodesic_find_midplane_crossing.argtypes = [C.c_void_p, C.c_int]
    L.geodesic_position_rad.restype = C.c_double
    L.geodesic_position_rad.argtypes = [C.c_void_p, C.c_double]
    L.rf.restype = C.c_double
    L.rf.argtypes = [C.c_double] * 3
    assert abs(L.rf(0.0, 1.0, 1.0) - np.pi / 2) < 1e-15
    g = H.golden("geodesic_init_inf.npz")
    dt = geodesic_struct_dtype()
    ref = g["g"].copy().view(dt).reshape(-1)
    checked = 0
    for i in range(0, 300):
        buf = (C.c_char * 240)()
        e = C.c_int(0)
        ok = L.geodesic_init_inf(g["incl"][i], g["spin"][i], g["alpha"][i], g["beta"][i], buf, C.byref(e))
        assert ok == g["ok"][i] and e.value == g["err"][i]
        if not ok:
            continue
        mine = np.frombuffer(buf, dtype=dt, count=1)[0]
        assert mine["type"] == ref["type"][i]
        for k in ("Rpc", "Tpp", "Tip", "m2p", "rp"):
            assert abs(float(mine[k]) - float(ref[k][i])) <= 1e-9 * abs(float(ref[k][i]))
        P = L.geodesic_find_midplane_crossing(buf, 0)
        if np.isnan(g["P0"][i]):
            assert np.isnan(P)
            continue
        assert abs(P - g["P0"][i]) <= 1e-9 * abs(g["P0"][i])
        r = L.geodesic_position_rad(buf, P)
        if np.isnan(g["r0"][i]):
            assert np.isnan(r)
        else:
            assert abs(r - g["r0"][i]) <= 1e-9 * abs(g["r0"][i])
            checked += 1
    assert checked > 100


def test_histogram_lattice_against_golden(gpu_api):
    p = abi.default_params(5, 48)
    p.n_spin, p.n_incl, p.n_bins = 3, 2, 32
    planes, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
    ref = H.golden("hist_cfg5_3x2x32_48.npz")["hist"]
    assert st.rays == 6 * 48 * 48
    assert np.allclose(planes["hist"], ref, rtol=1e-7, atol=0.0)
    # a lattice slice writes only its own images (the multi-GPU split of cfg 5)
    p.lattice_begin, p.lattice_end = 2, 5
    part = gpu_api.HostPlanes(p, pinned=False)
    gpu_api.trace_image(p, part)
    h = part["hist"].reshape(6, 32)
    assert np.all(h[:2] == 0) and np.all(h[5:] == 0)
    assert np.allclose(h[2:5], ref.reshape(6, 32)[2:5], rtol=1e-7)


def test_full_size_properties(gpu_api):
    """BASELINE.json's full sizes, through size-independent properties (the CPU reference would need minutes):
    cfg 2 at 4096^2 -- the image is mirror-consistent with a 4x coarser image of the same camera at the shared
    pixel centres?  No: centres differ.  Instead: (a) the class histogram of the device counters equals the
    histogram of the status plane, (b) hits have r >= r_ms, 0 < g < 2, flux >= 0, |phi| bounded; non-hits are
    all-zero, (c) tracing twice is bit-identical (idempotence, no race in the tile queue), (d) the lane-refill
    stepper returns the same planes as the no-refill stepper (cfg 4 at 192^2)."""
    p = abi.default_params(2)
    a, st = _gpu_planes(gpu_api, p)
    status = a["status"]
    cls = np.bincount(status & 31, minlength=32)
    assert list(cls) == list(st.class_count)
    hit = (status & 31) <= 1
    rms = abi.r_ms(p.bh_spin)
    assert np.all(a["r"][hit] >= rms) and np.all(a["r"][hit] < 200.0)
    assert np.all((a["g"][hit] > 0) & (a["g"][hit] < 2.0))
    assert np.all(a["flux"][hit] >= 0) and np.all(np.isfinite(a["phi"][hit])) and np.max(np.abs(a["phi"][hit])) < 40
    for k in ("r", "phi", "g", "flux"):
        assert np.all(a[k][~hit] == 0.0)
    assert hit.mean() > 0.95
    keep = {k: v.copy() for k, v in a.arrays.items()}
    b, _ = gpu_api.trace_image(p, a)
    for k in keep:
        assert np.array_equal(keep[k], b[k]), k
    p4 = abi.default_params(4, 192)
    p4.outputs |= abi.OUT_QERR
    x, sx = _gpu_planes(gpu_api, p4)
    keep = {k: v.copy() for k, v in x.arrays.items()}
    p4.flags = abi.FLAG_NO_REFILL
    y, sy = gpu_api.trace_image(p4, x)
    assert sx.total_steps == sy.total_steps
    for k in keep:
        assert np.array_equal(keep[k], y[k]), k
    done = (keep["status"] & 31)
    assert np.all((done == abi.ST_HORIZON) | (done == abi.ST_ESCAPE) | (done == abi.ST_ERRBREAK))
    esc = done == abi.ST_ESCAPE
    assert np.quantile(keep["qerr"][esc], 0.99) < 1e-3          # Carter-constant drift, the reference's own invariant


def test_full_index_split_writes_into_one_image(gpu_api):
    """SIM5_FLAG_FULL_INDEX: the calls of an interleaved split (one per GPU in a multi-GPU job) store their rows straight into
    ONE full-image set of device planes -- what the ranks do with rank 0's peer-mapped planes -- and the result is the image
    of a single unsplit call; CUDA IPC export/import entry points exist and reject null arguments."""
    p = abi.default_params(2, 160, 192)
    full, _ = _gpu_planes(gpu_api, p)
    img = gpu_api.DevicePlanes(p)
    try:
        for r in range(3):
            q = abi.default_params(2, 160, 192)
            q.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_FULL_INDEX
            q.split_count, q.split_index, q.split_rows = 3, r, 8
            st = gpu_api.trace_image_device(q, img.out)
            assert st.rays == 160 * 192 // 3
        for k in ("r", "phi", "g", "flux", "status"):
            assert np.array_equal(img.to_host(k), full[k], equal_nan=True), k
        h = img.handles()
        assert set(h) == {"r", "phi", "g", "flux", "status"} and all(len(v) == 64 for v in h.values())
    finally:
        img.close()
    L = gpu_api.lib()
    assert L.sim5_ipc_export(None, None) == abi.ERR_BAD_PARAM
    assert not L.sim5_ipc_import(None)


def test_async_train_and_phase_history(gpu_api):
    """A train of SIM5_FLAG_ASYNC calls enqueued back to back (what bench.py times): sim5_synchronize() completes them, the planes hold
    the image of a synchronous call, and sim5_phase_history() returns the per-kernel times of each call of the train."""
    p = abi.default_params(2, 256)
    full, _ = _gpu_planes(gpu_api, p)
    img = gpu_api.DevicePlanes(p)
    L = gpu_api.lib()
    try:
        q = abi.default_params(2, 256)
        q.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC
        st = abi.TraceStats()
        for _ in range(5):
            gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img.out), C.byref(st)), "sim5_trace_image")
        gpu_api.check(L.sim5_synchronize(), "sim5_synchronize")
        for k in ("r", "phi", "g", "flux", "status"):
            assert np.array_equal(img.to_host(k), full[k], equal_nan=True), k
        for back in range(5):
            ms, n = gpu_api.phase_history(back)
            assert n >= 3 and len(ms) == 3 and ms[0] > 0 and ms[1] > 0 and ms[2] >= 0, (back, ms, n)
        last, _, n_last = gpu_api.last_phase_ms()
        assert last == gpu_api.phase_history(0)[0]
        buf = (C.c_double * 3)()
        assert L.sim5_phase_history(64, buf, 3) == abi.ERR_BAD_PARAM and L.sim5_phase_history(-1, buf, 3) == abi.ERR_BAD_PARAM
        # SIM5_FLAG_DEFER_REDO: the redo passes of call k run beside call k+1 (two alternating queues); after sim5_synchronize() the
        # planes hold the same image.  The planes are cleared first so that a lost redo pass would show.
        for other in (abi.default_params(2, 256), abi.default_params(3, 192)):
            ref_img, _ = _gpu_planes(gpu_api, other)
            img2 = gpu_api.DevicePlanes(other, names=tuple(k for k in ("r", "phi", "g", "flux", "status") if k in ref_img.arrays))
            try:
                q = abi.ImageParams.from_buffer_copy(other)
                q.outputs = 0
                for name, bit, _ct in abi.PLANES:
                    if name in img2.ptrs:
                        q.outputs |= bit
                q.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC | abi.FLAG_DEFER_REDO
                for _ in range(5):
                    gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img2.out), C.byref(st)), "sim5_trace_image")
                gpu_api.check(L.sim5_synchronize(), "sim5_synchronize")
                for k in img2.ptrs:
                    assert np.array_equal(img2.to_host(k), ref_img[k], equal_nan=True), k
                # a synchronous call right after a deferred train drains it first
                gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img2.out), C.byref(st)), "sim5_trace_image")
                full2, _ = _gpu_planes(gpu_api, other)
                for k in img2.ptrs:
                    assert np.array_equal(full2[k], ref_img[k], equal_nan=True), k
            finally:
                gpu_api.check(L.sim5_synchronize(), "sim5_synchronize")
                img2.close()
    finally:
        img.close()
