"""the thick-disk surface finder (mode SURFACE).  The reference side is the Python layer's
DiskRaytrace.geodesic / __find_surface / image loop (python/sim5diskraytrace.py:163-391) spelled in C around the
unmodified reference library (oracle/ref_driver.c:pixel_surface); the product side is the per-lane state machine of
sim5_b200/csrc/pixel.cuh (SurfaceProg) in the lane-refill kernel.  Status byte and the number of geodesic_follow calls
are bit-exact artefacts; r, H, g, mu_e within 1e-9, flux within 1e-7."""
import numpy as np
import pytest

import harness as H
from sim5_b200 import abi


def _variants():
    out = []
    p = abi.default_params(7, 40); out.append(("preset", p))
    p = abi.default_params(7, 36); p.surf_hr = 0.0; out.append(("flat disk", p))
    p = abi.default_params(7, 40, 28)
    p.bh_spin, p.incl, p.surf_rin, p.surf_hr, p.rmax = 0.5, abi.deg2rad(30.0), 8.0, 0.2, 12.0
    out.append(("truncated at R=8", p))
    p = abi.default_params(7, 33); p.bh_spin, p.incl, p.surf_hr, p.rmax = 0.998, abi.deg2rad(25.0), 0.6, 12.0; out.append(("a=0.998 thick", p))
    p = abi.default_params(7, 24); p.rmax = 400.0; out.append(("wide field (restarts with a larger r0)", p))
    p = abi.default_params(7, 16); p.bh_spin = 1.5; out.append(("init error", p))
    p = abi.default_params(7, 16); p.incl, p.surf_hr = abi.deg2rad(80.0), 0.35; out.append(("observer below the surface", p))
    return out


def test_surface_constants_and_presets():
    p = abi.default_params(7)
    assert (p.mode, p.nx, p.ny) == (abi.MODE_SURFACE, 1024, 1024) and p.surf_hr == 0.2 and p.surf_rin == 0.0
    assert p.outputs & abi.OUT_HEIGHT and p.outputs & abi.OUT_STEPS


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("label,p", _variants(), ids=[v[0] for v in _variants()])
def test_hostsim_against_reference(label, p):
    got, _, _ = H.run_hostsim(p)
    ref, st, _ = H.run_ref(p)
    rep = H.assert_image_parity(got.arrays, ref.arrays, label=label)
    assert rep
    if label == "truncated at R=8":
        assert st.class_count[abi.ST_SURF_EQPLANE] > 0 and st.class_count[abi.ST_HIT0] > 0
    if label == "flat disk":
        assert st.total_steps == 0 and st.class_count[abi.ST_SURF_EQPLANE] > 0
    if label == "observer below the surface":
        assert st.class_count[abi.ST_SURF_BELOW] == p.nx * p.ny
    if label == "init error":
        assert st.class_count[abi.ST_INITERR + 12] == p.nx * p.ny


def test_hostsim_against_second_golden():
    g = H.golden("image_surface_rin8_40x28.npz")
    p = abi.default_params(7, 40, 28)
    p.bh_spin, p.incl, p.surf_rin, p.surf_hr, p.rmax = 0.5, abi.deg2rad(30.0), 8.0, 0.2, 12.0
    got, _, _ = H.run_hostsim(p)
    H.assert_image_parity(got.arrays, g, label="surface rin8 golden")


def test_surface_geometry_properties():
    """Size-independent properties of a surface hit: it lies on H(R) to the finder's accuracy, above the plane, and the
    g-factor / emission cosine are physical."""
    g = H.golden("image_cfg7_48.npz")
    p = abi.default_params(7, 48)
    hit = (g["status"] & 31) == abi.ST_HIT0
    assert hit.sum() > 2000
    r, Hh = g["r"][hit], g["height"][hit]
    R = np.sqrt(r * r - Hh * Hh)
    rin = abi.r_ms(p.bh_spin)
    Hs = np.where(R > rin, p.surf_hr * (R - rin) ** 2 / R, 0.0)
    assert np.all(Hh > -1e-2) and np.max(np.abs(Hh - Hs)) < 1e-2      # accuracy = 1e-2 in path length
    lit = hit & (g["flux"] > 0)
    assert np.all(g["g"][lit] > 0) and np.all(g["g"][lit] < 2) and np.all(g["mue"][lit] <= 1.0 + 1e-12)
    assert np.all(g["steps"][hit] > 10)


@pytest.mark.gpu
@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("label,p", _variants(), ids=[v[0] for v in _variants()])
def test_gpu_against_reference(gpu_api, label, p):
    got, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    ref, rst, _ = H.run_ref(p)
    H.assert_image_parity(got.arrays, ref.arrays, label=label)
    assert list(st.class_count) == list(rst.class_count) and st.total_steps == rst.total_steps


@pytest.mark.gpu
def test_gpu_against_second_golden_and_no_refill(gpu_api):
    g = H.golden("image_surface_rin8_40x28.npz")
    p = abi.default_params(7, 40, 28)
    p.bh_spin, p.incl, p.surf_rin, p.surf_hr, p.rmax = 0.5, abi.deg2rad(30.0), 8.0, 0.2, 12.0
    got, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    H.assert_image_parity(got.arrays, g, label="surface rin8 golden")
    assert list(st.class_count) == list(g["class_count"])
    p.flags |= abi.FLAG_NO_REFILL                     # lane refill changes scheduling, never results
    got2, _ = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    for k in got.arrays:
        assert np.array_equal(got[k], got2[k], equal_nan=True), k


@pytest.mark.gpu
def test_gpu_row_ranges_and_split_compose(gpu_api):
    """Row ranges and the interleaved multi-GPU split of a SURFACE image give exactly the rows of the full image."""
    p = abi.default_params(7, 64)
    full, _ = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    part = gpu_api.HostPlanes(p, pinned=True)
    for k in part.arrays:
        part[k][...] = 0
    for rb, re in ((0, 23), (23, 24), (24, 64)):
        p.row_begin, p.row_end = rb, re
        gpu_api.trace_image(p, part)
    for k in full.arrays:
        assert np.array_equal(full[k], part[k], equal_nan=True), k
    p = abi.default_params(7, 64)
    inter = gpu_api.HostPlanes(p, pinned=True)
    for k in inter.arrays:
        inter[k][...] = 0
    for idx in range(2):
        p.split_count, p.split_index, p.split_rows = 2, idx, 8
        gpu_api.trace_image(p, inter)
    for k in full.arrays:
        assert np.array_equal(full[k], inter[k], equal_nan=True), k
