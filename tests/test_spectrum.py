"""SPECTRUM mode (DiskRaytrace.spectrum of the reference's Python layer, python/sim5diskraytrace.py:43-134,
with blackbody() of sim5radiation.c:56-78): thermal spectrum of the thin disk summed over the image.
Checkers: the golden fixture written from the unmodified reference, the reference live, the C restatement.
Bar: 1e-7 relative to the spectrum's peak (north_star's flux tolerance); sums over ~1e4..1e7 positive terms in another order."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi


def _close(a, b, tol=1e-9):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * np.max(np.abs(b))


def test_host_instantiation_against_golden():
    g = H.golden("spectrum_cfg6_96.npz")
    p = abi.default_params(6, 96)
    assert np.array_equal(np.array(abi.spectrum_energies(p)), g["energies"])
    spec, _ = H.run_spectrum("hostsim", p)
    assert _close(spec, g["spectrum"]) and np.all(spec > 0)
    p.split_count, p.split_index, p.split_rows = 3, 1, 8
    part, _ = H.run_spectrum("hostsim", p)
    assert _close(part, g["part_3_1_8"])


@pytest.mark.skipif(not H.have_oracle(), reason="oracle/libsim5oracle.so not built")
def test_oracle_restatement_is_bit_identical_to_the_reference_fixture():
    g = H.golden("spectrum_cfg6_96.npz")
    p = abi.default_params(6, 96)
    spec, _ = H.run_spectrum("oracle", p)
    assert np.array_equal(spec, g["spectrum"])


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
def test_physics_sanity_and_reference_live():
    """Limb darkening off, no hardening, another camera: still the reference's numbers; the spectrum is a multi-colour black body:
    Rayleigh-Jeans slope ~ E^2 at the soft end... (here: monotone rise below the peak, exponential cut-off above)."""
    p = abi.default_params(6, 160)
    p.bh_spin, p.incl, p.spec_limb, p.spec_hardf, p.n_energy = 0.5, abi.deg2rad(40.0), 0, 1.0, 77
    p.rmax = abi.r_ms(0.5) + 30.0
    ref, _ = H.run_spectrum("ref", p)
    got, _ = H.run_spectrum("hostsim", p)
    assert _close(got, ref)
    k = int(np.argmax(ref))
    assert 5 < k < 70 and np.all(np.diff(ref[:k]) > 0) and np.all(np.diff(ref[k:]) < 0)
    assert ref[-1] < 1e-6 * ref[k]
    # partial spectra of an interleaved 2-way split add up to the whole (the multi-GPU reduction)
    parts = []
    for r in range(2):
        q = abi.default_params(6, 160)
        for name, _t in abi.ImageParams._fields_:
            setattr(q, name, getattr(p, name))
        q.split_count, q.split_index, q.split_rows = 2, r, 16
        parts.append(H.run_spectrum("hostsim", q)[0])
    assert _close(parts[0] + parts[1], ref)


@pytest.mark.gpu
def test_gpu_against_golden_and_reference(gpu_api):
    g = H.golden("spectrum_cfg6_96.npz")
    p = abi.default_params(6, 96)
    planes, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
    assert st.rays == 96 * 96 and st.kernel_launches == 1
    assert _close(planes["spectrum"], g["spectrum"], 1e-7)
    assert _close(planes["spectrum"], g["spectrum"], 1e-11)          # observed: ~1e-15
    p.split_count, p.split_index, p.split_rows = 3, 1, 8
    part, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
    assert st.rays == 96 * 96 // 3 and _close(part["spectrum"], g["part_3_1_8"], 1e-11)
    if H.have_ref():
        p = abi.default_params(6, 640)
        p.n_energy, p.spec_limb = 200, 0
        got, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
        ref, _ = H.run_spectrum("ref", p)
        assert _close(got["spectrum"], ref, 1e-11)
        print("spectrum 640^2 x 200 energies: kernel %.3f ms, max |d|/peak %.2e" % (st.kernel_ms, np.max(np.abs(got["spectrum"] - ref)) / ref.max()))


@pytest.mark.gpu
def test_gpu_bad_spectrum_parameters(gpu_api):
    L = gpu_api.lib()
    st = abi.TraceStats()
    p = abi.default_params(6, 16)
    hp = gpu_api.HostPlanes(p, pinned=False)
    p.n_energy = 257
    assert L.sim5_trace_image(C.byref(p), C.byref(hp.out), C.byref(st)) == abi.ERR_BAD_PARAM
    p = abi.default_params(6, 16); p.e_min_kev = 0.0
    assert L.sim5_trace_image(C.byref(p), C.byref(hp.out), C.byref(st)) == abi.ERR_BAD_PARAM
    p = abi.default_params(6, 16)
    hp.out.spectrum = None
    assert L.sim5_trace_image(C.byref(p), C.byref(hp.out), C.byref(st)) == abi.ERR_NO_OUTPUT
