"""sim5_trace_image_multi: one C call, several GPUs of the box (SURVEY.md 8b "device list", 8e).  On a one-GPU
box the device list is [0] (the split logic, the ragged-row segments and the reduction path still run); with >= 2 GPUs
(`gpurun --gpus 2`) the rows / lattice images are really dealt out and the planes of devices[0] are written over NVLink."""
import ctypes as C

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi

pytestmark = pytest.mark.gpu


def _devices(gpu_api):
    n = gpu_api.lib().sim5_gpu_device_count()
    return list(range(min(n, 8)))


def _single(gpu_api, p):
    got, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=True))
    return got, st


@pytest.mark.parametrize("cfg,nx,ny", [(2, 192, 256), (2, 130, 77), (3, 96, 64), (4, 40, 70), (7, 48, 33), (1, 64, 5)])
def test_multi_equals_single_call_host_planes(gpu_api, cfg, nx, ny):
    devs = _devices(gpu_api)
    p = abi.default_params(cfg, nx, ny)
    ref, st0 = _single(gpu_api, p)
    for dl in ([0], devs):
        planes = gpu_api.HostPlanes(p, pinned=True)
        for k in planes.arrays:
            planes[k][...] = 0
        got, st = gpu_api.trace_image_multi(p, dl, planes)
        assert st.rays == nx * ny and list(st.class_count) == list(st0.class_count) and st.total_steps == st0.total_steps
        for k in ref.arrays:
            assert np.array_equal(ref[k], got[k], equal_nan=True), (cfg, nx, ny, dl, k)
    gpu_api.init(0)


def test_multi_device_planes_on_first_device(gpu_api):
    """SIM5_FLAG_DEVICE_PTRS: the planes live on devices[0]; the other GPUs store their rows into them through peer access."""
    devs = _devices(gpu_api)
    gpu_api.init(0)
    p = abi.default_params(2, 160, 200)
    ref, _ = _single(gpu_api, p)
    img = gpu_api.DevicePlanes(p)
    try:
        q = abi.ImageParams.from_buffer_copy(p)
        q.flags |= abi.FLAG_DEVICE_PTRS
        st = abi.TraceStats()
        dl = (C.c_int * len(devs))(*devs)
        gpu_api.check(gpu_api.lib().sim5_trace_image_multi(C.byref(q), C.byref(img.out), C.byref(st), dl, len(devs)), "sim5_trace_image_multi")
        assert st.rays == 160 * 200
        for k in ("r", "phi", "g", "flux", "status"):
            assert np.array_equal(img.to_host(k), ref[k], equal_nan=True), k
    finally:
        img.close()


def test_multi_histogram_reduced_on_first_device(gpu_api):
    """cfg 5 over the device list: lattice images dealt out one by one, partial lattices added up on devices[0] (k_sum_peers)."""
    devs = _devices(gpu_api)
    p = abi.default_params(5, 64)
    p.n_spin, p.n_incl, p.n_bins = 5, 3, 32
    ref, st0 = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
    for dl in ([0], devs):
        got, st = gpu_api.trace_image_multi(p, dl, gpu_api.HostPlanes(p, pinned=False))
        assert st.rays == st0.rays and list(st.class_count) == list(st0.class_count)
        assert np.allclose(got["hist"], ref["hist"], rtol=1e-7, atol=0.0)
    # the interleaved deal by hand: the calls of a 3-way split add up to the lattice, each one zero outside its own images
    acc = np.zeros_like(ref["hist"])
    for r in range(3):
        q = abi.ImageParams.from_buffer_copy(p)
        q.split_count, q.split_index = 3, r
        part, st = gpu_api.trace_image(q, gpu_api.HostPlanes(q, pinned=False))
        h = part["hist"].reshape(15, 32)
        own = np.arange(15) % 3 == r
        assert np.all(h[~own] == 0) and st.rays == own.sum() * 64 * 64
        acc += part["hist"]
    assert np.allclose(acc, ref["hist"], rtol=1e-7, atol=0.0)
    g = H.golden("hist_cfg5_3x2x32_48.npz")["hist"]
    p = abi.default_params(5, 48)
    p.n_spin, p.n_incl, p.n_bins = 3, 2, 32
    got, _ = gpu_api.trace_image_multi(p, devs, gpu_api.HostPlanes(p, pinned=False))
    assert np.allclose(got["hist"], g, rtol=1e-7, atol=0.0)


def test_multi_rejects_bad_device_lists(gpu_api):
    L = gpu_api.lib()
    p = abi.default_params(1, 16)
    planes = gpu_api.HostPlanes(p, pinned=False)
    st = abi.TraceStats()
    two = (C.c_int * 2)(0, 0)
    assert L.sim5_trace_image_multi(C.byref(p), C.byref(planes.out), C.byref(st), two, 2) == abi.ERR_BAD_PARAM      # listed twice
    far = (C.c_int * 1)(99)
    assert L.sim5_trace_image_multi(C.byref(p), C.byref(planes.out), C.byref(st), far, 1) == abi.ERR_BAD_PARAM
    assert L.sim5_trace_image_multi(C.byref(p), C.byref(planes.out), C.byref(st), None, 0) == abi.ERR_BAD_PARAM


def test_chunks_ignore_split_rows_without_a_split(gpu_api):
    """advisor r01: split_rows > 1 with split_count <= 1 and a row count that is not a multiple of it shifted the last chunk."""
    L = gpu_api.lib()
    try:
        p = abi.default_params(2, 96, 101)
        p.flags = abi.FLAG_NO_OVERLAP
        ref, _ = _single(gpu_api, p)
        p.flags = 0
        p.split_count, p.split_rows = 1, 32
        L.sim5_set_chunk_rays(96 * 20)
        got = gpu_api.HostPlanes(p, pinned=True)
        for k in got.arrays:
            got[k][...] = 0
        _, st = gpu_api.trace_image(p, got)
        assert st.kernel_launches > 5
        for k in ref.arrays:
            assert np.array_equal(ref[k], got[k], equal_nan=True), k
    finally:
        L.sim5_set_chunk_rays(0)


def test_stage_copy_assembles_the_image_by_dma(gpu_api):
    """SIM5_FLAG_STAGE_COPY: the calls of an interleaved split trace into library-owned compact planes and the finished row blocks are
    moved into ONE full-image set of planes by strided 2-D copies on the copy stream (what the ranks of a multi-GPU job do with the
    assembling rank's peer-mapped planes); a train of such calls, with and without deferred redo passes, gives the image of one call."""
    L = gpu_api.lib()
    for cfg, nx, ny in ((2, 160, 192), (3, 96, 128), (4, 40, 64)):
        p = abi.default_params(cfg, nx, ny)
        full, _ = _single(gpu_api, p)
        names = tuple(k for k in full.arrays)
        img = gpu_api.DevicePlanes(p, names=names)
        try:
            st = abi.TraceStats()
            for defer in (0, abi.FLAG_DEFER_REDO, abi.FLAG_DEFER_REDO | abi.FLAG_ALT_STREAMS):
                for rep in range(3):                      # a train: the two scratch sets alternate
                    for r in range(4):
                        q = abi.ImageParams.from_buffer_copy(p)
                        q.flags = abi.FLAG_DEVICE_PTRS | abi.FLAG_ASYNC | abi.FLAG_FULL_INDEX | abi.FLAG_STAGE_COPY | defer
                        q.split_count, q.split_index, q.split_rows = 4, r, 8
                        gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img.out), C.byref(st)), "sim5_trace_image")
                gpu_api.check(L.sim5_synchronize(), "sim5_synchronize")
                for k in names:
                    assert np.array_equal(img.to_host(k), full[k], equal_nan=True), (cfg, k, defer)
        finally:
            gpu_api.check(L.sim5_synchronize(), "sim5_synchronize")
            img.close()


ALL_PLANES = tuple(name for name, _, _ in abi.PLANES)


def _lane_image(gpu_api, cfg, nx, ny, counter_start=None, devices=None):
    """cfg through SIM5_FLAG_SHARED_QUEUE into device planes pre-filled with 0xFF bytes; returns ({plane: flat array}, stats)."""
    L = gpu_api.lib()
    p = abi.default_params(cfg, nx, ny)
    p.flags |= abi.FLAG_DEVICE_PTRS | abi.FLAG_SHARED_QUEUE
    img = gpu_api.DevicePlanes(p, names=ALL_PLANES)
    ctr = L.sim5_device_alloc(64)
    try:
        for name, ptr in img.ptrs.items():
            gpu_api.check(L.sim5_device_memset(C.c_void_p(ptr), 0xFF, img.n * np.dtype(img.dtypes[name]).itemsize), "sim5_device_memset")
        gpu_api.check(L.sim5_device_memset(C.c_void_p(ctr), 0, 64), "sim5_device_memset")
        if counter_start is not None:
            v = np.array([counter_start], dtype=np.uint64)
            gpu_api.check(L.sim5_host_to_device(C.c_void_p(ctr), v.ctypes.data, 8), "sim5_host_to_device")
        st = abi.TraceStats()
        if devices is None:
            img.out.shared_counter = ctr
            gpu_api.check(L.sim5_trace_image(C.byref(p), C.byref(img.out), C.byref(st)), "sim5_trace_image")
        else:
            dl = (C.c_int * len(devices))(*devices)
            gpu_api.check(L.sim5_trace_image_multi(C.byref(p), C.byref(img.out), C.byref(st), dl, len(devices)), "sim5_trace_image_multi")
        return {k: img.to_host(k) for k in img.ptrs}, st
    finally:
        L.sim5_device_free(C.c_void_p(ctr))
        img.close()


@pytest.mark.parametrize("cfg,nx,ny", [(4, 48, 40), (7, 56, 33)])
def test_shared_queue_image_equals_the_private_queue_image(gpu_api, cfg, nx, ny):
    """SIM5_FLAG_SHARED_QUEUE: the rays come from the caller's counter (system-scope atomics) instead of the call's own -- same image;
    through sim5_trace_image_multi every device pulls from one counter on devices[0] and stores into devices[0]'s planes."""
    gpu_api.init(0)
    p = abi.default_params(cfg, nx, ny)
    ref, st0 = _single(gpu_api, p)
    for devices in (None, [0], _devices(gpu_api)):
        got, st = _lane_image(gpu_api, cfg, nx, ny, devices=devices)
        assert st.rays == nx * ny and list(st.class_count) == list(st0.class_count) and st.total_steps == st0.total_steps, devices
        for k in ref.arrays:
            assert np.array_equal(got[k], np.asarray(ref[k]).reshape(-1), equal_nan=True), (cfg, devices, k)
    gpu_api.init(0)


def test_shared_queue_skips_the_rays_another_device_took(gpu_api):
    """the counter starts at K, as if another GPU had pulled the first K rays: those pixels stay untouched, the rest is the image"""
    gpu_api.init(0)
    nx, ny, K = 64, 32, 700
    p = abi.default_params(4, nx, ny)
    p.flags |= abi.FLAG_ROW_MAJOR                 # ray k == pixel k
    ref, _ = _single(gpu_api, p)
    L = gpu_api.lib()
    q = abi.ImageParams.from_buffer_copy(p)
    q.flags |= abi.FLAG_DEVICE_PTRS | abi.FLAG_SHARED_QUEUE
    img = gpu_api.DevicePlanes(q, names=ALL_PLANES)
    ctr = L.sim5_device_alloc(64)
    try:
        for name, ptr in img.ptrs.items():
            gpu_api.check(L.sim5_device_memset(C.c_void_p(ptr), 0xFF, img.n * np.dtype(img.dtypes[name]).itemsize), "sim5_device_memset")
        v = np.array([K], dtype=np.uint64)
        gpu_api.check(L.sim5_host_to_device(C.c_void_p(ctr), v.ctypes.data, 8), "sim5_host_to_device")
        img.out.shared_counter = ctr
        st = abi.TraceStats()
        gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img.out), C.byref(st)), "sim5_trace_image")
        assert sum(st.class_count) == nx * ny - K
        s = img.to_host("status")
        assert np.all(s[:K] == 0xFF) and np.array_equal(s[K:], np.asarray(ref["status"]).reshape(-1)[K:])
        for k in ("intensity", "tau"):
            if k in img.ptrs:
                assert np.array_equal(img.to_host(k)[K:], np.asarray(ref[k]).reshape(-1)[K:], equal_nan=True), k
        # drained queue: a second call on the same counter traces nothing
        st2 = abi.TraceStats()
        gpu_api.check(L.sim5_trace_image(C.byref(q), C.byref(img.out), C.byref(st2)), "sim5_trace_image")
        assert sum(st2.class_count) == 0
    finally:
        L.sim5_device_free(C.c_void_p(ctr))
        img.close()


def test_shared_queue_is_refused_where_it_cannot_work(gpu_api):
    L = gpu_api.lib()
    gpu_api.init(0)
    for cfg, extra in ((2, abi.FLAG_DEVICE_PTRS), (4, 0)):
        p = abi.default_params(cfg, 32, 32)
        p.flags |= abi.FLAG_SHARED_QUEUE | extra
        planes = gpu_api.HostPlanes(p, pinned=False)
        st = abi.TraceStats()
        assert L.sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)) == abi.ERR_BAD_PARAM
        dl = (C.c_int * 1)(0)
        assert L.sim5_trace_image_multi(C.byref(p), C.byref(planes.out), C.byref(st), dl, 1) == abi.ERR_BAD_PARAM
