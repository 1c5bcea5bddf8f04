"""ctypes/numpy binding of libsim5b200.so (a Python binding for the batched
entry that needs no SWIG).  Mirrors the reference's Python-side usage (python/sim5diskraytrace.py:
DiskRaytrace.image) at image granularity: one call traces the whole image on the GPU.

There is no CPU path: if the CUDA library is missing or no device is usable, calls raise.
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIM5_B200_LIB") or os.path.join(_HERE, "libsim5b200.so")   # env override: kernel-variant A/B runs only

_lib = None


class Sim5Error(RuntimeError):
    pass


def lib():
    """Load libsim5b200.so (built in-tree by sim5_b200/csrc/Makefile or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Sim5Error("CUDA library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "sim5_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.sim5_gpu_init.argtypes = [C.c_int]
        L.sim5_gpu_init.restype = C.c_int
        L.sim5_gpu_device_count.restype = C.c_int
        L.sim5_set_stream.argtypes = [C.c_void_p]
        L.sim5_set_stream.restype = C.c_int
        L.sim5_synchronize.restype = C.c_int
        L.sim5_join.restype = C.c_int
        L.sim5_set_chunk_rays.argtypes = [C.c_int64]
        L.sim5_set_chunk_rays.restype = C.c_int
        L.sim5_last_error.restype = C.c_char_p
        L.sim5_version.restype = C.c_char_p
        L.sim5_host_alloc.argtypes = [C.c_size_t]
        L.sim5_host_alloc.restype = C.c_void_p
        L.sim5_host_free.argtypes = [C.c_void_p]
        L.sim5_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.sim5_host_register.restype = C.c_int
        L.sim5_host_unregister.argtypes = [C.c_void_p]
        L.sim5_host_unregister.restype = C.c_int
        L.sim5_device_alloc.argtypes = [C.c_size_t]
        L.sim5_device_alloc.restype = C.c_void_p
        L.sim5_device_free.argtypes = [C.c_void_p]
        L.sim5_device_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
        L.sim5_device_memset.restype = C.c_int
        L.sim5_host_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sim5_host_to_device.restype = C.c_int
        L.sim5_device_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sim5_device_to_host.restype = C.c_int
        L.sim5_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
        L.sim5_ipc_export.restype = C.c_int
        L.sim5_ipc_import.argtypes = [C.c_void_p]
        L.sim5_ipc_import.restype = C.c_void_p
        L.sim5_ipc_release.argtypes = [C.c_void_p]
        L.sim5_ipc_release.restype = C.c_int
        L.sim5_default_params.argtypes = [C.c_int, C.POINTER(abi.ImageParams)]
        L.sim5_default_params.restype = C.c_int
        L.sim5_trace_image.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(abi.ImageOut), C.POINTER(abi.TraceStats)]
        L.sim5_trace_image.restype = C.c_int
        L.sim5_trace_image_multi.argtypes = [C.POINTER(abi.ImageParams), C.POINTER(abi.ImageOut), C.POINTER(abi.TraceStats), C.POINTER(C.c_int), C.c_int]
        L.sim5_trace_image_multi.restype = C.c_int
        L.sim5_fp64_peak_tflops.argtypes = [C.c_int, C.c_int]
        L.sim5_fp64_peak_tflops.restype = C.c_double
        L.sim5_last_phase_ms.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int64)]
        L.sim5_last_phase_ms.restype = C.c_int
        L.sim5_phase_history.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
        L.sim5_phase_history.restype = C.c_int
        dp = C.POINTER(C.c_double)
        L.sim5_batch_rf.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.sim5_batch_rd.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.sim5_batch_rc.argtypes = [C.c_int64, dp, dp, dp]
        L.sim5_batch_rj.argtypes = [C.c_int64, dp, dp, dp, dp, dp]
        L.sim5_batch_rf_hi.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.sim5_batch_rj_hi.argtypes = [C.c_int64, dp, dp, dp, dp, dp]
        L.sim5_batch_sncndn.argtypes = [C.c_int64, dp, dp, dp, dp, dp]
        L.sim5_batch_libm.argtypes = [C.c_int, C.c_int64, dp, dp, dp]
        for f in ("sim5_batch_rf_hi", "sim5_batch_rj_hi", "sim5_batch_rf", "sim5_batch_rd", "sim5_batch_rc", "sim5_batch_rj", "sim5_batch_sncndn", "sim5_batch_libm"):
            getattr(L, f).restype = C.c_int
        _lib = L
    return _lib


def check(rc, what="sim5 call"):
    if rc != abi.OK:
        msg = lib().sim5_last_error()
        raise Sim5Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def init(device=0):
    check(lib().sim5_gpu_init(device), "sim5_gpu_init")


def last_error():
    m = lib().sim5_last_error()
    return m.decode() if m else ""


_NP = {C.c_double: np.float64, C.c_int32: np.int32, C.c_uint8: np.uint8}


def _free_pinned(ptr):
    try:
        lib().sim5_host_free(C.c_void_p(ptr))
    except Exception:
        pass


class PinnedArray:
    """numpy view over pinned host memory from sim5_host_alloc.  The memory belongs to the ctypes buffer every view is based on
    (numpy keeps it alive through `.base`), and is freed when the LAST view goes away -- an array handed out by HostPlanes may
    outlive the HostPlanes object."""

    def __init__(self, n, dtype):
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        self.ptr = lib().sim5_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise Sim5Error("sim5_host_alloc failed: " + last_error())
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        weakref.finalize(buf, _free_pinned, self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))


class HostPlanes:
    """Output planes of one image in pinned host memory + the sim5_image_out struct pointing at them."""

    def __init__(self, p, pinned=True):
        n = p.nx * p.ny
        self.out = abi.ImageOut()
        self.arrays = {}
        self._keep = []
        self.shape = (p.ny, p.nx)
        for name, bit, ct in abi.PLANES:
            if p.outputs & bit:
                if pinned:
                    pa = PinnedArray(n, _NP[ct])
                    self._keep.append(pa)
                    a = pa.array
                else:
                    a = np.zeros(n, dtype=_NP[ct])
                self.arrays[name] = a
                setattr(self.out, name, a.ctypes.data)
        if p.mode == abi.MODE_HISTOGRAM:
            nh = p.n_spin * p.n_incl * p.n_bins
            a = np.zeros(nh, dtype=np.float64)
            self.arrays["hist"] = a
            self.out.hist = a.ctypes.data
        if p.mode == abi.MODE_SPECTRUM:
            a = np.zeros(p.n_energy, dtype=np.float64)
            self.arrays["spectrum"] = a
            self.out.spectrum = a.ctypes.data

    def __getitem__(self, k):
        return self.arrays[k]

    def image(self, k):
        return self.arrays[k].reshape(self.shape)


def trace_image(p, planes=None, pinned=True):
    """Trace one image (or its rows [row_begin,row_end)) on the GPU.  Returns (HostPlanes, TraceStats)."""
    if planes is None:
        planes = HostPlanes(p, pinned=pinned)
    st = abi.TraceStats()
    check(lib().sim5_trace_image(C.byref(p), C.byref(planes.out), C.byref(st)), "sim5_trace_image")
    return planes, st


def trace_image_device(p, out_struct):
    """Trace with caller-owned DEVICE planes (e.g. torch tensors' data_ptr()); nothing is copied to the host.
    The caller's params are not modified.  Synchronous unless p.flags carries FLAG_ASYNC; after an ASYNC or DEFER_REDO call the
    planes are complete once synchronize() has returned (DevicePlanes.to_host waits by itself)."""
    q = abi.ImageParams.from_buffer_copy(p)
    q.flags |= abi.FLAG_DEVICE_PTRS
    st = abi.TraceStats()
    check(lib().sim5_trace_image(C.byref(q), C.byref(out_struct), C.byref(st)), "sim5_trace_image")
    return st


def trace_image_multi(p, devices, planes=None, pinned=True):
    """One call on several GPUs of the box (sim5_trace_image_multi): rows in interleaved 32-row blocks, lattice images one by one,
    the histogram reduced on devices[0].  Returns (HostPlanes, TraceStats)."""
    if planes is None:
        planes = HostPlanes(p, pinned=pinned)
    st = abi.TraceStats()
    devs = (C.c_int * len(devices))(*devices)
    check(lib().sim5_trace_image_multi(C.byref(p), C.byref(planes.out), C.byref(st), devs, len(devices)), "sim5_trace_image_multi")
    return planes, st


def synchronize():
    check(lib().sim5_synchronize(), "sim5_synchronize")


class DevicePlanes:
    """Full-image output planes in DEVICE memory owned by this process (sim5_device_alloc), exportable to the other
    ranks of a one-process-per-GPU job as CUDA IPC handles (peer-mapped planes: every rank stores its rows into them)."""

    def __init__(self, p, names=("r", "phi", "g", "flux", "status")):
        self.n = p.nx * p.ny
        self.shape = (p.ny, p.nx)
        self.ptrs, self.dtypes = {}, {}
        self.out = abi.ImageOut()
        for name, bit, ct in abi.PLANES:
            if name in names and (p.outputs & bit):
                nbytes = self.n * C.sizeof(ct)
                ptr = lib().sim5_device_alloc(nbytes)
                if not ptr:
                    raise Sim5Error("sim5_device_alloc failed: " + last_error())
                self.ptrs[name], self.dtypes[name] = ptr, _NP[ct]
                setattr(self.out, name, ptr)
        self._owned = True

    def handles(self):
        hs = {}
        for name, ptr in self.ptrs.items():
            buf = C.create_string_buffer(64)
            check(lib().sim5_ipc_export(C.c_void_p(ptr), buf), "sim5_ipc_export")
            hs[name] = buf.raw
        return hs

    @classmethod
    def from_handles(cls, p, hs):
        self = cls.__new__(cls)
        self.n, self.shape = p.nx * p.ny, (p.ny, p.nx)
        self.ptrs, self.dtypes, self.out, self._owned = {}, {}, abi.ImageOut(), False
        for name, bit, ct in abi.PLANES:
            if name in hs:
                ptr = lib().sim5_ipc_import(C.create_string_buffer(hs[name], 64))
                if not ptr:
                    raise Sim5Error("sim5_ipc_import failed: " + last_error())
                self.ptrs[name], self.dtypes[name] = ptr, _NP[ct]
                setattr(self.out, name, ptr)
        return self

    def to_host(self, name):
        a = np.empty(self.n, dtype=self.dtypes[name])
        check(lib().sim5_device_to_host(a.ctypes.data, C.c_void_p(self.ptrs[name]), a.nbytes), "sim5_device_to_host")
        return a

    def close(self):
        for ptr in self.ptrs.values():
            if self._owned:
                lib().sim5_device_free(C.c_void_p(ptr))
            else:
                lib().sim5_ipc_release(C.c_void_p(ptr))
        self.ptrs = {}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _batch(fn, ins, nout=1, pre=()):
    ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
    n = ins[0].size
    outs = [np.empty(n, dtype=np.float64) for _ in range(nout)]
    check(fn(*pre, C.c_int64(n), *[_dp(a) for a in ins], *[_dp(o) for o in outs]), fn.__name__)
    return outs[0] if nout == 1 else outs


def batch_rf(x, y, z):
    return _batch(lib().sim5_batch_rf, [x, y, z])


def batch_rd(x, y, z):
    return _batch(lib().sim5_batch_rd, [x, y, z])


def batch_rc(x, y):
    return _batch(lib().sim5_batch_rc, [x, y])


def batch_rj(x, y, z, p):
    return _batch(lib().sim5_batch_rj, [x, y, z, p])


def batch_rf_hi(x, y, z):
    return _batch(lib().sim5_batch_rf_hi, [x, y, z])


def batch_rj_hi(x, y, z, p):
    return _batch(lib().sim5_batch_rj_hi, [x, y, z, p])


def batch_sncndn(u, m):
    return _batch(lib().sim5_batch_sncndn, [u, m], nout=3)


LIBM_OPS = {"sin": 0, "cos": 1, "log": 2, "atan2": 3, "acos": 4, "asin": 5, "atan": 6,
            "pow_third": 7, "pow_1p5": 8, "pow_4": 9, "exp": 10}


def batch_libm(op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if b is None:
        b = np.zeros_like(a)
    return _batch(lib().sim5_batch_libm, [a, b], pre=(C.c_int(LIBM_OPS[op]),))


def fp64_peak_tflops(device=0, iters=8192):
    v = lib().sim5_fp64_peak_tflops(device, iters)
    if v < 0:
        raise Sim5Error("sim5_fp64_peak_tflops failed: " + last_error())
    return v


def last_phase_ms():
    """Device time (ms) of each kernel of the most recent sim5_trace_image call, [trace] or [trace, azimuth RR,
    azimuth RC], the (RR, RC) disk-hit counts the azimuth kernels integrated, and the number of kernels launched."""
    buf = (C.c_double * 3)()
    items = (C.c_int64 * 2)()
    n = lib().sim5_last_phase_ms(buf, 3, items)
    if n < 0:
        raise Sim5Error("sim5_last_phase_ms failed: " + last_error())
    return [buf[i] for i in range(min(n, 3))], (items[0], items[1]), n


def phase_history(back):
    """Per-kernel device times (ms) of the image call `back` calls ago (0 = most recent) and the number of kernels it launched."""
    buf = (C.c_double * 3)()
    n = lib().sim5_phase_history(back, buf, 3)
    if n < 0:
        raise Sim5Error("sim5_phase_history failed: " + last_error())
    return [buf[i] for i in range(min(n, 3))], n


def write_text_dump(path, planes, p):
    """Text dump compatible with the reference example (disk-image.c:116-120): `y x flux g` per line,
    blank line after each row."""
    f_img = planes.image("flux")
    g_img = planes.image("g")
    with open(path, "w") as fh:
        for iy in range(p.ny):
            for ix in range(p.nx):
                fh.write("%d  %d  %e  %e\n" % (iy, ix, np.float32(f_img[iy, ix]), np.float32(g_img[iy, ix])))
            fh.write("\n")
