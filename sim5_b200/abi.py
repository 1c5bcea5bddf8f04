"""ctypes mirror of include/sim5_b200.h (structs, constants, BASELINE config presets).

Pure declarations: nothing here computes rays.  The presets follow SURVEY.md 8(d);
`sim5_default_params` in the C library returns the same numbers (tests assert it).
"""
import ctypes as C
import math

# modes
MODE_EQPLANE, MODE_POLARIZED, MODE_STEPWISE, MODE_HISTOGRAM, MODE_SPECTRUM, MODE_SURFACE = 0, 1, 2, 3, 4, 5
# outputs
OUT_R, OUT_PHI, OUT_G, OUT_FLUX = 0x001, 0x002, 0x004, 0x008
OUT_CHI, OUT_DELTA, OUT_MUE = 0x010, 0x020, 0x040
OUT_INTENSITY, OUT_TAU, OUT_STEPS, OUT_STATUS, OUT_QERR = 0x080, 0x100, 0x200, 0x400, 0x800
OUT_HEIGHT, OUT_DELAY = 0x1000, 0x2000
# flags
FLAG_DEVICE_PTRS, FLAG_NO_REFILL, FLAG_ASYNC, FLAG_SINGLE_PASS, FLAG_NO_OVERLAP, FLAG_EXACT_AZIMUTH, FLAG_FULL_INDEX = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40
FLAG_DEFER_REDO = 0x80
FLAG_STAGE_COPY = 0x100
FLAG_ROW_MAJOR = 0x200
FLAG_ALT_STREAMS = 0x400
FLAG_SHARED_QUEUE = 0x800
# status classes
ST_HIT0, ST_HIT1, ST_MISS, ST_NOCROSS0, ST_NOCROSS1, ST_HIT2, ST_NOCROSS2 = 0, 1, 2, 3, 4, 5, 6
ST_HORIZON, ST_ESCAPE, ST_ERRBREAK, ST_MAXSTEPS, ST_NOSTART = 8, 9, 10, 11, 12
ST_SURF_UNDER, ST_SURF_EQPLANE, ST_SURF_BELOW, ST_SURF_LOST = 7, 13, 14, 15
ST_INITERR = 16
GT_NONE, GT_RR, GT_RC, GT_CC, GT_RR_DBL, GT_RR_BH = 0, 1, 2, 3, 4, 5

# error codes
OK, ERR_NO_DEVICE, ERR_BAD_PARAM, ERR_CUDA, ERR_NOT_IMPL, ERR_NO_OUTPUT = 0, -1, -2, -3, -4, -5

PLANES = (  # (field, output bit, ctype)
    ("r", OUT_R, C.c_double), ("phi", OUT_PHI, C.c_double), ("g", OUT_G, C.c_double),
    ("flux", OUT_FLUX, C.c_double), ("chi", OUT_CHI, C.c_double), ("delta", OUT_DELTA, C.c_double),
    ("mue", OUT_MUE, C.c_double), ("intensity", OUT_INTENSITY, C.c_double), ("tau", OUT_TAU, C.c_double),
    ("qerr", OUT_QERR, C.c_double), ("steps", OUT_STEPS, C.c_int32), ("status", OUT_STATUS, C.c_uint8),
    ("height", OUT_HEIGHT, C.c_double), ("delay", OUT_DELAY, C.c_double),
)


class ImageParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("mode", C.c_int32), ("outputs", C.c_uint32), ("flags", C.c_uint32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
        ("max_order", C.c_int32), ("device", C.c_int32),
        ("bh_spin", C.c_double), ("incl", C.c_double), ("rmax", C.c_double), ("r_emit_min", C.c_double),
        ("disk_mass", C.c_double), ("disk_mdot", C.c_double), ("disk_alpha", C.c_double),
        ("precision_factor", C.c_double), ("r_start", C.c_double), ("step_max", C.c_double),
        ("max_steps", C.c_int32), ("reserved0", C.c_int32),
        ("torus_rc", C.c_double), ("torus_w", C.c_double), ("torus_h", C.c_double),
        ("torus_ell", C.c_double), ("torus_j0", C.c_double), ("torus_k0", C.c_double),
        ("n_spin", C.c_int32), ("n_incl", C.c_int32), ("n_bins", C.c_int32),
        ("lattice_begin", C.c_int32), ("lattice_end", C.c_int32), ("reserved1", C.c_int32),
        ("spin_max", C.c_double), ("incl_min_deg", C.c_double), ("incl_max_deg", C.c_double),
        ("g_min", C.c_double), ("g_max", C.c_double), ("rmax_offset", C.c_double),
        ("split_count", C.c_int32), ("split_index", C.c_int32), ("split_rows", C.c_int32), ("reserved2", C.c_int32),
        ("n_energy", C.c_int32), ("spec_limb", C.c_int32),
        ("e_min_kev", C.c_double), ("e_max_kev", C.c_double), ("spec_hardf", C.c_double),
        ("surf_hr", C.c_double), ("surf_rin", C.c_double), ("delay_r_ref", C.c_double),
    ]


class ImageOut(C.Structure):
    _fields_ = [
        ("r", C.c_void_p), ("phi", C.c_void_p), ("g", C.c_void_p), ("flux", C.c_void_p),
        ("chi", C.c_void_p), ("delta", C.c_void_p), ("mue", C.c_void_p), ("intensity", C.c_void_p),
        ("tau", C.c_void_p), ("qerr", C.c_void_p), ("steps", C.c_void_p), ("status", C.c_void_p),
        ("hist", C.c_void_p), ("spectrum", C.c_void_p), ("height", C.c_void_p), ("delay", C.c_void_p),
        ("shared_counter", C.c_void_p),
    ]


class TraceStats(C.Structure):
    _fields_ = [
        ("rays", C.c_int64), ("class_count", C.c_int64 * 32), ("gtype_count", C.c_int64 * 8),
        ("total_steps", C.c_int64), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
        ("kernel_launches", C.c_int32), ("sm_count", C.c_int32), ("grid_ctas", C.c_int32),
        ("cta_threads", C.c_int32),
    ]


def r_ms(a):
    """Marginally stable orbit, spelled as sim5kerr.c:993-1004 (libm cbrt, same operation order)."""
    z1 = 1. + math.cbrt(1. - a * a) * (math.cbrt(1. + a) + math.cbrt(1. - a))
    z2 = math.sqrt(3. * a * a + z1 * z1)
    return 3. + z2 - math.sqrt((3. - z1) * (3. + z1 + 2. * z2))


def deg2rad(d):
    """sim5math.h:50 -- (a)/180.0*M_PI, in this order."""
    return d / 180.0 * math.pi


def ell_kepler(r, a):
    """Keplerian specific angular momentum, sim5kerr.c:1050-1071 (Komissarov 2008 form)."""
    return (r * r - 2. * a * math.sqrt(r) + a * a) / (math.sqrt(r) * r - 2. * math.sqrt(r) + a)


def spectrum_energies(p):
    """The detector energies [keV] of a SPECTRUM call, spelled as the library computes them (libm pow)."""
    n = p.n_energy
    if n == 1:
        return [p.e_min_kev]
    return [p.e_min_kev * math.pow(p.e_max_kev / p.e_min_kev, k / (n - 1)) for k in range(n)]


def default_params(cfg, nx=None, ny=None):
    """BASELINE.json configs 1..5 with the open parameters fixed as in SURVEY.md 8(d); 6 = the SPECTRUM preset, 7 = the SURFACE preset."""
    p = ImageParams()
    p.struct_size = C.sizeof(ImageParams)
    p.device = -1                  # the calling thread's current context (api.init(device))
    p.max_order = 1
    p.disk_mass, p.disk_mdot, p.disk_alpha = 10.0, 0.1, 0.1
    p.precision_factor, p.r_start, p.step_max, p.max_steps = 0.01, 50.0, 1e9, 100000
    p.torus_rc, p.torus_w, p.torus_h = 10.0, 2.0, 0.3
    p.torus_j0, p.torus_k0 = 1.0, 0.05
    p.n_spin, p.n_incl, p.n_bins = 64, 32, 256
    p.spin_max, p.incl_min_deg, p.incl_max_deg = 0.998, 5.0, 85.0
    p.g_min, p.g_max, p.rmax_offset = 0.0, 2.0, 20.0
    p.n_energy, p.spec_limb, p.e_min_kev, p.e_max_kev, p.spec_hardf = 128, 1, 0.05, 50.0, 1.7
    p.surf_hr, p.surf_rin, p.delay_r_ref = 0.2, 0.0, 1000.0
    if cfg == 1:
        p.mode, p.nx, p.ny = MODE_EQPLANE, 512, 512
        p.bh_spin, p.incl = 0.9, deg2rad(70.0)
        p.rmax = r_ms(p.bh_spin) + 8.0
        p.outputs = OUT_R | OUT_G | OUT_FLUX | OUT_STATUS
    elif cfg == 2:
        p.mode, p.nx, p.ny = MODE_EQPLANE, 4096, 4096
        p.bh_spin, p.incl = 0.998, deg2rad(75.0)
        p.rmax = r_ms(p.bh_spin) + 20.0
        p.outputs = OUT_R | OUT_PHI | OUT_G | OUT_FLUX | OUT_STATUS
    elif cfg == 3:
        p.mode, p.nx, p.ny = MODE_POLARIZED, 2048, 2048
        p.bh_spin, p.incl = 0.94, deg2rad(75.0)
        p.rmax = r_ms(p.bh_spin) + 20.0
        p.outputs = OUT_R | OUT_PHI | OUT_G | OUT_FLUX | OUT_CHI | OUT_DELTA | OUT_STATUS
    elif cfg == 4:
        p.mode, p.nx, p.ny = MODE_STEPWISE, 1024, 1024
        p.bh_spin, p.incl = 0.9, deg2rad(60.0)
        p.rmax = 25.0
        p.outputs = OUT_INTENSITY | OUT_TAU | OUT_STEPS | OUT_STATUS
    elif cfg == 5:
        p.mode, p.nx, p.ny = MODE_HISTOGRAM, 1024, 1024
        p.bh_spin, p.incl = 0.998, deg2rad(75.0)   # unused: the lattice defines spin/inclination
        p.rmax = 0.0
        p.outputs = 0
    elif cfg == 6:
        p.mode, p.nx, p.ny = MODE_SPECTRUM, 2048, 2048
        p.bh_spin, p.incl = 0.998, deg2rad(75.0)
        p.rmax = r_ms(p.bh_spin) + 20.0
        p.outputs = 0
    elif cfg == 7:
        p.mode, p.nx, p.ny = MODE_SURFACE, 1024, 1024
        p.bh_spin, p.incl = 0.9, deg2rad(60.0)
        p.rmax = 30.0
        p.outputs = OUT_R | OUT_HEIGHT | OUT_G | OUT_MUE | OUT_FLUX | OUT_STEPS | OUT_STATUS
    else:
        raise ValueError("cfg must be 1..7")
    p.torus_ell = ell_kepler(p.torus_rc, p.bh_spin)
    if nx is not None:
        p.nx = nx
        p.ny = ny if ny is not None else nx
    return p
