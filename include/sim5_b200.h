/*
 * sim5_b200.h -- batched C-ABI of the B200-native SIM5 photon hot path.
 *
 * This is the NEW entry point that replaces the caller-owned serial pixel loop
 * of the reference (examples/04-disk-image-eqplane/disk-image.c:53-105 and
 * python/sim5diskraytrace.py:163-205): one call traces a whole image (or a row
 * block of it) on one B200 with hand-written sm_100a FP64 kernels.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types cross the boundary.
 * There is no CPU fallback: every entry point returns SIM5_ERR_NO_DEVICE when no
 * CUDA device is usable.
 *
 * The scalar reference API (geodesic_init_inf, raytrace, rf, ...) is declared in
 * sim5lib.h next to this file and is implemented on top of the same device code.
 */
#ifndef SIM5_B200_H
#define SIM5_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ */
/* error codes (negative = failure), cf. the reference's TRUE/FALSE +  */
/* optional int* error convention (sim5kerr-geod.c:59-98)              */
/* ------------------------------------------------------------------ */
#define SIM5_OK                  0
#define SIM5_ERR_NO_DEVICE      -1   /* no CUDA device / driver: the library never computes on the CPU */
#define SIM5_ERR_BAD_PARAM      -2
#define SIM5_ERR_CUDA           -3   /* a CUDA runtime call failed; see sim5_last_error() */
#define SIM5_ERR_NOT_IMPL       -4
#define SIM5_ERR_NO_OUTPUT      -5   /* a plane selected in `outputs` has a NULL pointer */

/* ------------------------------------------------------------------ */
/* trace modes                                                         */
/* ------------------------------------------------------------------ */
#define SIM5_MODE_EQPLANE        0   /* thin-disk image: init_inf + midplane crossing (+ azimuth) + gfactorK + NT flux */
#define SIM5_MODE_POLARIZED      1   /* EQPLANE + local tetrad g-factor, emission angle, Walker-Penrose chi, Chandrasekhar delta */
#define SIM5_MODE_STEPWISE       2   /* raytrace() stepping through an optically thin torus */
#define SIM5_MODE_HISTOGRAM      3   /* transfer function: g-factor histograms over a (spin, inclination) lattice */
#define SIM5_MODE_SPECTRUM       4   /* observed thermal spectrum of the thin disk: sum over the image of the (limb-darkened, colour-corrected)
                                        black-body intensity of every disk hit -- DiskRaytrace.spectrum of the reference's Python layer
                                        (python/sim5diskraytrace.py:43-134) on the image grid, blackbody() of sim5radiation.c:56-78 */
#define SIM5_MODE_SURFACE        5   /* image of a geometrically THICK disk: the ray is followed analytically (geodesic_P_int -> geodesic_follow
                                        with the adaptive step of DiskRaytrace.__find_surface, python/sim5diskraytrace.py:214-336) until it meets the
                                        surface H(R); g-factor and emission cosine from tetrad_surface (python/sim5diskraytrace.py:340-391) */

/* output plane selector bits (all planes are row-major [ny][nx], index iy*nx+ix) */
#define SIM5_OUT_R            0x001  /* radius of the disk hit                      (double) */
#define SIM5_OUT_PHI          0x002  /* azimuth travelled from infinity             (double) */
#define SIM5_OUT_G            0x004  /* g-factor                                    (double) */
#define SIM5_OUT_FLUX         0x008  /* observed flux F(r)*g^4                      (double) */
#define SIM5_OUT_CHI          0x010  /* polarization angle at the observer          (double) */
#define SIM5_OUT_DELTA        0x020  /* polarization degree at emission             (double) */
#define SIM5_OUT_MUE          0x040  /* cosine of the emission angle                (double) */
#define SIM5_OUT_INTENSITY    0x080  /* STEPWISE: integrated intensity              (double) */
#define SIM5_OUT_TAU          0x100  /* STEPWISE: integrated optical depth          (double) */
#define SIM5_OUT_STEPS        0x200  /* STEPWISE: number of raytrace() calls        (int32)  */
#define SIM5_OUT_STATUS       0x400  /* classification / termination byte           (uint8)  */
#define SIM5_OUT_QERR         0x800  /* STEPWISE: raytrace_error() at the end       (double) */
#define SIM5_OUT_HEIGHT      0x1000  /* SURFACE: height r*cos(theta) of the hit     (double) */
#define SIM5_OUT_DELAY       0x2000  /* EQPLANE/POLARIZED: geodesic_timedelay() between the disk hit and the sphere r = delay_r_ref (double) */
#define SIM5_NPLANES             14

/* flags */
#define SIM5_FLAG_DEVICE_PTRS   0x1  /* pointers in sim5_image_out are device pointers (no staging, no D2H) */
#define SIM5_FLAG_NO_REFILL     0x2  /* debugging: disable warp-level lane refill / compaction */
#define SIM5_FLAG_SINGLE_PASS    0x8  /* A/B testing: compute the azimuth inside the tracing kernel instead of the queued second phase */
#define SIM5_FLAG_NO_OVERLAP    0x10 /* A/B testing: host planes are copied back after the whole image instead of chunk by chunk under the kernels */
#define SIM5_FLAG_EXACT_AZIMUTH 0x20 /* compute phi of RR disk hits with the bit-faithful Carlson routines (the reference's iteration counts and
                                        series, same bits as its CPU path) instead of the default tolerance-mode kernel, which evaluates the same
                                        integrals to ~1e-14 with the 7th-order series in fewer steps.  r, g, flux and every status flag are
                                        bit-faithful in both modes. */
#define SIM5_FLAG_FULL_INDEX    0x40 /* with DEVICE_PTRS and split_count > 1: the planes are FULL-image planes (index iy*nx+ix) instead of this call's
                                        compact rows -- e.g. the peer-mapped planes of the rank that assembles the image (sim5_ipc_import), so
                                        every GPU stores its rows directly into the final image over NVLink and no gather is needed */
#define SIM5_FLAG_DEFER_REDO    0x80 /* with DEVICE_PTRS | ASYNC, for a train of images traced back to back: the bit-faithful redo passes of the azimuth (one
                                        latency-bound wave over ~0.3 % of the rays) are left running on the library's auxiliary stream while the NEXT
                                        call's tracing kernel starts (two alternating work queues).  phi of a call is complete once sim5_join() has been
                                        enqueued on / sim5_synchronize() has returned for the launch stream */
#define SIM5_FLAG_STAGE_COPY   0x100 /* with DEVICE_PTRS | ASYNC | FULL_INDEX and split_count > 1 (the planes are another GPU's, mapped with sim5_ipc_import): trace into
                                        library-owned compact planes in LOCAL memory and move the finished row blocks into the caller's full-image planes with the
                                        copy engine (one strided 2-D copy per plane on the library's copy stream, two alternating scratch sets), so the transfer of
                                        image k rides under the kernels of image k+1 instead of every plane store crossing NVLink.  With 8 GPUs the ~70 MB per GPU
                                        and image otherwise arrive at the assembling GPU as 8-byte stores from 7 kernels at once and stretch their tracing kernels
                                        by 50 % (profiles/r05m_bench_cfg2_n8.json).  The planes are complete after sim5_join() / sim5_synchronize() */
#define SIM5_FLAG_ROW_MAJOR    0x200 /* A/B testing, STEPWISE: hand the rows to the lanes top to bottom instead of from the middle of the image outwards */
#define SIM5_FLAG_ALT_STREAMS  0x400 /* A/B testing, with DEFER_REDO: overlap consecutive calls of a train -- the tracing kernels alternate between two internal
                                        low-priority launch streams (image k+1 starts tracing when image k has finished tracing, not when its azimuth has), the
                                        azimuth kernels run on a high-priority stream.  Measured: it does NOT pay.  Without priorities the two kernels share the SMs
                                        and slow each other down (2 M-ray slice: phase A 0.499 -> 0.514 ms, azimuth 0.356 -> 0.451 ms, step 0.870 -> 0.890 ms;
                                        profiles/r05q_train_ab.log); with them the step is 0.866 -> 0.885 ms (profiles/r05r_train_ab_priority.log).  A train
                                        therefore stays on the caller's stream by default */
#define SIM5_FLAG_SHARED_QUEUE 0x800 /* STEPWISE / SURFACE with DEVICE_PTRS, no row split: the rays of the image are handed out from the caller's counter
                                        out->shared_counter -- a zeroed 64-bit word in memory every participating GPU can reach (the assembling GPU's, mapped
                                        with sim5_ipc_import in the other processes) -- with system-scope atomics.  All GPUs of a job then pull rays of ONE
                                        image from ONE queue, a warp refill at a time, and store into the full-image planes: the static row split plus work
                                        stealing of north_star taken to its limit (everything is "tail").  For the step-wise modes, whose rays cost 1 600 ...
                                        8 200 steps each: a static split leaves the GPUs up to 12 % apart (profiles/r05o_bench_cfg4_n4_centerout.json).
                                        One counter word per call: a train of calls uses consecutive words */
#define SIM5_FLAG_ASYNC         0x4  /* with DEVICE_PTRS: enqueue on the library stream (sim5_set_stream) and return without
                                        synchronising; stats are not filled.  Pair with sim5_synchronize(). */

/* ------------------------------------------------------------------ */
/* per-pixel status byte                                               */
/*   bits 0..4 : class / termination code                              */
/*   bits 5..7 : geodesic type code                                    */
/* ------------------------------------------------------------------ */
#define SIM5_ST_HIT0             0   /* order-0 crossing at r >= r_emit_min */
#define SIM5_ST_HIT1             1   /* order-1 crossing at r >= r_emit_min */
#define SIM5_ST_MISS             2   /* crossings exist but none at r >= r_emit_min (incl. r = NaN) */
#define SIM5_ST_NOCROSS0         3   /* geodesic_find_midplane_crossing(order 0) returned NaN */
#define SIM5_ST_NOCROSS1         4   /* order 0 fell inside r_emit_min, order 1 returned NaN */
#define SIM5_ST_HIT2             5   /* order-2 crossing (only if max_order >= 2) */
#define SIM5_ST_NOCROSS2         6
/* stepwise terminations */
#define SIM5_ST_HORIZON          8   /* r < 1.05 r_bh */
#define SIM5_ST_ESCAPE           9   /* r > 1.01 r_start */
#define SIM5_ST_ERRBREAK        10   /* rtd.error > 1e-2 */
#define SIM5_ST_MAXSTEPS        11
#define SIM5_ST_NOSTART         12   /* ray never reaches r_start (r_start < pericentre) or start position undefined */
/* SURFACE terminations (HIT0 = the surface was found; HORIZON, ESCAPE = left 1.1 r0 four times, MAXSTEPS = step underflow are shared) */
#define SIM5_ST_SURF_UNDER       7   /* the ray passed below the equatorial plane (m < 0) */
#define SIM5_ST_SURF_EQPLANE    13   /* the ray reached the equatorial plane (H < 1e-4) before any surface: result = midplane crossing */
#define SIM5_ST_SURF_BELOW      14   /* the start point of the search already lies below the surface */
#define SIM5_ST_SURF_LOST       15   /* geodesic_follow() reported status 0 (horizon or end of the position integral) */
/* init errors: 16 + GD_ERROR_* of geodesic_init_inf */
#define SIM5_ST_INITERR         16
#define SIM5_ST_CLASS(s)        ((s) & 0x1f)
#define SIM5_ST_GTYPE(s)        (((s) >> 5) & 0x7)
/* geodesic type codes in bits 5..7 (GEOD_TYPE_* of sim5kerr-geod.h:19-23 compressed) */
#define SIM5_GT_NONE             0
#define SIM5_GT_RR               1
#define SIM5_GT_RC               2
#define SIM5_GT_CC               3
#define SIM5_GT_RR_DBL           4
#define SIM5_GT_RR_BH            5

/* ------------------------------------------------------------------ */
/* harness-defined emission models. They are NOT part   */
/* of the reference; oracle driver and kernels share these numbers.    */
/* ------------------------------------------------------------------ */
/* Chandrasekhar (1960) pure-scattering limb polarization degree, mu = 0, 0.05, ..., 1 ;
 * linear interpolation between the knots. */
#define SIM5_CHANDRA_N 21
static const double SIM5_CHANDRA_DELTA[SIM5_CHANDRA_N] = {
    0.11713, 0.08979, 0.07448, 0.06311, 0.05410, 0.04667, 0.04041,
    0.03502, 0.03033, 0.02619, 0.02252, 0.01923, 0.01627, 0.01358,
    0.01112, 0.00888, 0.00682, 0.00492, 0.00316, 0.00152, 0.00000
};

typedef struct sim5_image_params {
    int32_t  struct_size;        /* = sizeof(sim5_image_params); ABI guard */
    int32_t  mode;               /* SIM5_MODE_* */
    uint32_t outputs;            /* SIM5_OUT_* mask */
    uint32_t flags;              /* SIM5_FLAG_* */
    int32_t  nx, ny;             /* image size in pixels */
    int32_t  row_begin, row_end; /* rows [row_begin,row_end) traced by this call; 0,0 = all rows.
                                    HOST planes are always addressed with the FULL image index iy*nx+ix.  DEVICE planes
                                    (SIM5_FLAG_DEVICE_PTRS) too, except in an interleaved split (split_count > 1) without
                                    SIM5_FLAG_FULL_INDEX: those are COMPACT, [rows of this call][nx] (see split_count below) */
    int32_t  max_order;          /* highest crossing order tried (reference example: 1) */
    int32_t  device;             /* CUDA device ordinal; < 0 (what sim5_default_params sets): the calling thread's current context,
                                    i.e. the device of its last sim5_gpu_init / explicit ordinal, else the process default.  Every
                                    device has a context of its own; naming another device never tears one down */
    /* geometry -- pixel rule of disk-image.c:57-58:
     *   alpha = ((ix+.5)/nx-.5)*2*rmax ; beta = ((iy+.5)/ny-.5)*2*rmax*(ny/nx) */
    double   bh_spin;
    double   incl;               /* radians */
    double   rmax;
    double   r_emit_min;         /* <= 0 : r_ms(bh_spin) as in disk-image.c:41,83 */
    /* Novikov-Thorne disk, disk_nt_setup(M, a, mdot, alpha, 0) sim5disk-nt.c:37 */
    double   disk_mass;
    double   disk_mdot;
    double   disk_alpha;
    /* STEPWISE */
    double   precision_factor;   /* raytrace_prepare() argument */
    double   r_start;            /* rays are started at this radius on the way in */
    double   step_max;           /* *step passed to raytrace() each call */
    int32_t  max_steps;          /* safety bound on raytrace() calls per ray */
    int32_t  reserved0;
    double   torus_rc;           /* torus centre radius */
    double   torus_w;            /* radial width */
    double   torus_h;            /* vertical scale height (in units of cylindrical radius) */
    double   torus_ell;          /* constant specific angular momentum of the torus */
    double   torus_j0;           /* emissivity normalisation */
    double   torus_k0;           /* absorption normalisation */
    /* HISTOGRAM */
    int32_t  n_spin, n_incl, n_bins;
    int32_t  lattice_begin, lattice_end; /* images [begin,end) of the n_spin*n_incl lattice; 0,0 = all.  With split_count > 1 the call traces
                                            the images begin + split_index, begin + split_index + split_count, ... of that range and zeroes the bins of
                                            the others, so the histograms of the split_count GPUs ADD UP to the lattice (one reduce at the end) */
    int32_t  reserved1;
    double   spin_max;           /* a_j = spin_max*j/(n_spin-1) */
    double   incl_min_deg, incl_max_deg; /* i_k = min + (max-min)*k/(n_incl-1) degrees */
    double   g_min, g_max;       /* histogram range */
    double   rmax_offset;        /* rmax = r_ms(a_j) + rmax_offset */
    /* multi-GPU interleaved row split: with split_count > 1 this call traces only the row blocks
     * b = (iy - row_begin) / split_rows with b % split_count == split_index (one process per GPU, each with its
     * own index).  Host planes keep full-image indexing; DEVICE planes are compact: local row
     * lr = (b / split_count) * split_rows + (iy - row_begin) % split_rows, index lr*nx + ix. */
    int32_t  split_count, split_index, split_rows, reserved2;
    /* SPECTRUM: n_energy <= 256 energies E_k = e_min_kev * (e_max_kev/e_min_kev)^(k/(n_energy-1)) at the detector;
     * per disk hit T = (F(r)/sigma_SB)^(1/4), g and the emission cosine mu_e from the Keplerian emitter frame,
     * spectrum[k] += B_E(T; hardening spec_hardf, limb darkening 0.5+0.75 mu_e if spec_limb) at E_k/g * g^3 * dalpha*dbeta
     * [erg cm^-2 s^-1 keV^-1 srad^-1 x (GM/c^2)^2]; the caller multiplies by (GM/c^2 / D)^2 */
    int32_t  n_energy, spec_limb;
    double   e_min_kev, e_max_kev, spec_hardf;
    /* SURFACE: harness-defined disk surface H(R) = surf_hr (R - surf_rin)^2 / R for R > surf_rin, 0 inside (dH/dR = surf_hr (1 - surf_rin^2/R^2)),
     * R = r sin(theta); Keplerian rotation ell = ellK(R, a), no radial drift, Novikov-Thorne flux at R.  surf_hr = 0: flat disk
     * (the `flat` branch of DiskRaytrace.geodesic).  surf_rin <= 0: r_ms(bh_spin) */
    double   surf_hr, surf_rin;
    /* SIM5_OUT_DELAY: radius of the sphere the travel time is measured to (> every disk radius in the image) */
    double   delay_r_ref;
} sim5_image_params;

typedef struct sim5_image_out {
    double  *r;
    double  *phi;
    double  *g;
    double  *flux;
    double  *chi;
    double  *delta;
    double  *mue;
    double  *intensity;
    double  *tau;
    double  *qerr;
    int32_t *steps;
    uint8_t *status;
    double  *hist;               /* [n_spin][n_incl][n_bins] */
    double  *spectrum;           /* SPECTRUM: [n_energy], always a HOST array (it is 2 KB) */
    double  *height;
    double  *delay;
    uint64_t *shared_counter;    /* SIM5_FLAG_SHARED_QUEUE: the ray queue of this call (device-reachable, zero before the first participant starts) */
} sim5_image_out;

typedef struct sim5_trace_stats {
    int64_t  rays;               /* rays traced by this call */
    int64_t  class_count[32];    /* histogram of SIM5_ST_CLASS over the traced rays */
    int64_t  gtype_count[8];     /* histogram of SIM5_ST_GTYPE */
    int64_t  total_steps;        /* STEPWISE: sum of raytrace() calls */
    double   kernel_ms;          /* device time of the trace kernel(s), CUDA events on the launch stream */
    double   total_ms;           /* device time of the whole call incl. copies */
    int32_t  kernel_launches;    /* kernels launched by this call */
    int32_t  sm_count;
    int32_t  grid_ctas, cta_threads;
} sim5_trace_stats;

/* lifecycle ------------------------------------------------------------- */
int  sim5_gpu_init(int device);        /* create context/streams/scratch on `device` (idempotent) and make it the calling thread's current
                                          context; the first successful call also sets the process default.  Every entry point makes its
                                          context's device current (cudaSetDevice) before it touches the CUDA runtime */
int  sim5_set_stream(void* cuda_stream); /* launch on the caller's cudaStream_t (e.g. torch's current stream); NULL = library stream */
int  sim5_set_chunk_rays(int64_t rays); /* host-plane calls trace and copy back in chunks of about this many rays (copy of chunk k under the kernels of chunk k+1); <= 0 restores the default (2^21) */
int  sim5_synchronize(void);            /* wait for everything enqueued by SIM5_FLAG_ASYNC calls */
int  sim5_join(void);                   /* make the launch stream wait (on the device, no host sync) for the deferred redo passes of earlier SIM5_FLAG_DEFER_REDO calls */
void sim5_gpu_shutdown(void);          /* tears down the contexts of all devices */
int  sim5_gpu_device_count(void);      /* 0 when no usable device */
const char* sim5_last_error(void);
const char* sim5_version(void);

/* pinned host memory for output planes (so the D2H leg of the call can overlap tracing) */
void* sim5_host_alloc(size_t bytes);
void  sim5_host_free(void* p);
/* page-lock / release caller-owned host memory (cudaHostRegister): e.g. one shared-memory image that the ranks of a one-process-per-GPU
 * job all map, each copying its own rows into it over its own PCIe link */
int   sim5_host_register(void* p, size_t bytes);
int   sim5_host_unregister(void* p);
/* device memory helpers for SIM5_FLAG_DEVICE_PTRS users (e.g. a torch tensor's data_ptr works too) */
void* sim5_device_alloc(size_t bytes);
void  sim5_device_free(void* p);
int   sim5_device_memset(void* p, int value, size_t bytes);            /* stream-ordered on the context's launch stream, then synchronised */
int   sim5_host_to_device(void* dst, const void* src, size_t bytes);   /* same ordering; for small control words (a pre-set ray queue) */
int   sim5_device_to_host(void* dst, const void* src, size_t bytes);   /* waits for the context's launch stream and deferred redo passes first */
/* CUDA IPC for one-process-per-GPU jobs on one node: export a sim5_device_alloc'd plane as a 64-byte handle, import it in
 * another process (peer access over NVLink is enabled on first use), release the mapping before the owner frees the plane */
int   sim5_ipc_export(const void* device_ptr, void* handle64);
void* sim5_ipc_import(const void* handle64);
int   sim5_ipc_release(void* imported_ptr);

/* defaults: fills every field with the SURVEY.md 8(d) definition of BASELINE config `cfg` (1..5); 6 = the SPECTRUM preset
 * (the camera of config 2 at 2048^2, 128 energies 0.05..50 keV, hardening 1.7, limb darkening on); 7 = the SURFACE preset (a = 0.9,
 * i = 60 deg, 1024^2, rmax = 30, H/R -> 0.2, inner edge r_ms) */
int  sim5_default_params(int cfg, sim5_image_params* p);

/* THE batched entry: replaces the per-pixel loop of disk-image.c:53-105 */
int  sim5_trace_image(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats);

/* The same call on `ndev` GPUs of the box (devices[] = distinct CUDA ordinals; p->device is ignored): one host thread per device
 * drives that device's context.  Image rows are dealt out in interleaved 32-row blocks (split_rows overrides the 32), ragged row
 * counts included; lattice images of HISTOGRAM mode one by one.  Host planes: every GPU copies its own rows into the caller's planes
 * (pinned planes from sim5_host_alloc let the copies overlap the kernels).  SIM5_FLAG_DEVICE_PTRS: the planes live on devices[0] and
 * the other GPUs store their rows into them over NVLink (peer access).  HISTOGRAM: devices[0] adds the partial lattices up with loads
 * from its peers' memory, in a fixed order, before out->hist is written.  stats: counters summed over the devices, kernel_ms /
 * total_ms of the slowest device.  SIM5_FLAG_ASYNC and SIM5_FLAG_DEFER_REDO do not apply.
 * SIM5_FLAG_SHARED_QUEUE (STEPWISE / SURFACE with DEVICE_PTRS): no row split at all -- every device runs the whole row range and pulls
 * its rays from ONE counter on devices[0] (the library's own word when out->shared_counter is NULL), so the devices finish within one
 * ray of each other whatever the image looks like; stats->rays of each device = the rays it pulled. */
int  sim5_trace_image_multi(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats, const int* devices, int ndev);

/* device time in ms of each kernel of the most recent sim5_trace_image call (CUDA events on the launch stream; waits for
 * the call to finish, so it also works after SIM5_FLAG_ASYNC): ms[0] trace kernel (phase A), ms[1] azimuth of the RR
 * and RC hits by the tolerance-mode kernel (default) or of the RR hits by the bit-faithful kernel (SIM5_FLAG_EXACT_AZIMUTH),
 * ms[2] the bit-faithful redo passes (default; normally empty lists) or the bit-faithful RC kernel; items (may be NULL):
 * [0] RR and [1] RC disk hits integrated.  Returns the number of kernels the call launched (1, 3 or 4), <0 on error. */
int  sim5_last_phase_ms(double* ms, int n, int64_t* items);
/* the same times for the image call `back` calls ago (0 = most recent, the last 63 are kept): a caller that enqueues a train of
 * SIM5_FLAG_ASYNC calls reads every call's per-kernel times afterwards instead of synchronising after each */
int  sim5_phase_history(int back, double* ms, int n);

/* FP64 DFMA-chain microbenchmark: returns measured TFLOP/s (2 flop per DFMA) of the device, <0 on error.
 * This is the roofline denominator for the compute-bound FP64 path (MEASURED_PEAKS.json has no FP64 entry). */
double sim5_fp64_peak_tflops(int device, int iters);

/* development aid: kernel-only time in ms of `reps` calls per thread of one device routine over n threads
 * (which: 0 rf, 1 rj, 2 rc, 3 sncndn, 4 sincos, 5 log, 6 atan2, 7 pow(.,1/3), 8 four divisions, 9 four sqrt, 10 complete rj) */
double sim5_micro_bench(int which, int64_t n, int reps);

/* batched element-wise access to the device functions (parity tests call these through the C-ABI).
 * Each mirrors one reference function; n elements, SoA arrays. */
int sim5_batch_rf(int64_t n, const double* x, const double* y, const double* z, double* out);           /* sim5elliptic.c:18 */
int sim5_batch_rd(int64_t n, const double* x, const double* y, const double* z, double* out);           /* sim5elliptic.c:58 */
int sim5_batch_rc(int64_t n, const double* x, const double* y, double* out);                            /* sim5elliptic.c:104 */
int sim5_batch_rj(int64_t n, const double* x, const double* y, const double* z, const double* p, double* out); /* sim5elliptic.c:144 */
/* tolerance-mode Carlson integrals of the azimuth phase (sim5_b200/csrc/ellfast.cuh): the same R_F / R_J to a few ulp, not bit for bit;
 * NaN for arguments outside their domain (x zero or in [2^-60,2^60]; y, z, p in [2^-60,2^60]) */
int sim5_batch_rf_hi(int64_t n, const double* x, const double* y, const double* z, double* out);
int sim5_batch_rj_hi(int64_t n, const double* x, const double* y, const double* z, const double* p, double* out);
int sim5_batch_sncndn(int64_t n, const double* u, const double* m, double* sn, double* cn, double* dn); /* sim5elliptic.c:535 */
/* unary/binary libm-compatible device functions (correctly-rounded double-double implementations):
 * op: 0 sin, 1 cos, 2 log, 3 atan2(y=a,x=b), 4 acos, 5 asin, 6 atan, 7 pow(a,1./3.), 8 pow(a,1.5), 9 pow(a,4.), 10 exp */
int sim5_batch_libm(int op, int64_t n, const double* a, const double* b, double* out);
/* the Byrd & Friedman integrals behind geodesic_timedelay (sim5elliptic.c:645-1139), n elements each; v[0..6] are the argument arrays
 * (NULL = zeros).  op: 0 integral_C1(u,m)  1 C2(u,m)  2 C2_cos(cn,m)  3 Z2(a,b,u,m)  4 Rm1(a,u,m)  5 Rm2(a,u,m)  6 R2(a,u,m)
 * 7 R_r0_re(a,b,c,d,X)  8 R_r0_re_inf(a,b,c,d)  9 R_r1_re(a,b,c,d,X)  10 R_r2_re(a,b,c,d,X)  11 T_m0(a2,b2,X)  12 T_m2(a2,b2,X)
 * and with the complex pair c = v[2] + i v[3]:  13 R_r0_cc(a,b,c,X=v[4])  14 R_r0_cc_inf(a,b,c)  15 R_r1_cc(a,b,c,X1=v[4],X2=v[5])
 * 16 R_r2_cc(a,b,c,X1,X2)  17 R_rp_cc2(a,b,c,p=v[6],X1,X2) */
int sim5_batch_integral(int op, int64_t n, const double* const* v, double* out);
/* geodesic_timedelay (sim5kerr-geod.c:559-731) between the radii ra[i] and rb[i] on the incoming branch of the geodesic with
 * impact parameters (alpha[i], beta[i]) of an observer at inclination incl [rad] around a hole of spin a */
int sim5_batch_timedelay(int64_t n, double incl, double a, const double* alpha, const double* beta, const double* ra, const double* rb, double* out);

#ifdef __cplusplus
}
#endif
#endif /* SIM5_B200_H */
