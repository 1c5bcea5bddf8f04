/*
 * kernels.cuh -- the sm_100a kernels of the SIM5 photon hot path.
 *
 * Execution model (B200: 148 SMs, FP64 non-tensor pipe is the bound, HBM traffic is write-only and tiny):
 *   - one ray per thread; persistent grid of (SM count x resident CTAs) CTAs;
 *   - every WARP pulls tiles of 32 consecutive pixels of a row from one global atomic counter
 *     (dynamic balance between the cheap outer image and the expensive shadow edge, no block syncs);
 *   - per-image constants (S5ImageConsts, ~0.7 KB incl. the Novikov-Thorne and Chandrasekhar tables)
 *     travel as a __grid_constant__ kernel parameter and are staged in shared memory once per CTA
 *     (the table lookups are lane-divergent, which shared memory serves without serialisation);
 *   - outputs are SoA planes; a warp stores 32 consecutive doubles per plane (one 256-byte
 *     fully-coalesced transaction), no reads from HBM at all;
 *   - the stepwise kernel keeps one live ray per lane and refills finished lanes from the queue
 *     (ballot + one aggregated atomic per warp) so 300..8000-step rays do not idle their warp.
 */
#ifndef SIM5_KERNELS_CUH
#define SIM5_KERNELS_CUH

#include <cuda_runtime.h>
#include "pixel.cuh"

namespace s5 {

struct DevOut {
    double *r, *phi, *g, *flux, *chi, *delta, *mue, *intensity, *tau, *qerr, *height, *delay;
    int* steps;
    unsigned char* status;
    int compact;                  /* 1: plane index = local_row*nx + ix ; 0: full-image index iy*nx + ix */
};

struct DevStats {                 /* device-side counters, flushed once per CTA */
    unsigned long long cls[32];
    unsigned long long gtype[8];
    unsigned long long steps;
    unsigned long long rays;
};

/* queue of azimuth work items between phase A (k_trace_eqplane) and phase B (k_azimuth): SoA, RR items from the
 * front, RC items from the back of the same arrays, so each phase-B launch sees one geodesic type only */
struct AzQueue {
    double* f;                    /* [S5_AZ_NFIELDS][cap] */
    unsigned long long* key;      /* [cap]: bits 0..47 output index, 48..51 nrr, 52 ppc, 53 beta >= 0, 54 turn (AzIn), 56 rf_ok */
    unsigned long long* count;    /* [0] RR items, [1] RC items, [2] RR / [3] RC items the tolerance-mode kernel handed back (redo lists) */
    unsigned* redo;               /* [cap] slots of those items: RR from the front, RC from the back */
    long long cap;                /* 0: no queue -> the azimuth is computed inline by phase A */
};

S5_HD S5_INL unsigned long long az_key(size_t i, const AzIn& z)
{
    return (unsigned long long)i | ((unsigned long long)(z.nrr & 15) << 48) | ((unsigned long long)(z.ppc ? 1 : 0) << 52) |
           ((unsigned long long)(z.beta_nonneg ? 1 : 0) << 53) | ((unsigned long long)(z.turn ? 1 : 0) << 54) | ((unsigned long long)(z.rf_ok ? 1 : 0) << 56);
}
S5_HD S5_INL void az_unkey(unsigned long long key, AzIn* z)
{
    z->nrr = (int)((key >> 48) & 15);
    z->ppc = ((key >> 52) & 1) != 0;
    z->beta_nonneg = ((key >> 53) & 1) != 0;
    z->turn = ((key >> 54) & 1) != 0;
    z->rf_ok = ((key >> 56) & 1) != 0;
}

#define S5_CTA_THREADS 128
#ifndef S5_EQ_THREADS
#define S5_EQ_THREADS 512          /* CTA size of the eq-plane trace kernels.  Re-swept on the round-2 routine (profiles/r06e_proxy.log, phase A of the 4096^2
                                      image, 2 CTAs/SM): 384 (80 regs) 3.73 ms, 416 (72) 4.12, 448 (72) 3.75, 480 (64) 3.75, 512 (64) 3.59 */
#endif
#ifndef S5_MIN_CTAS_EQ
#define S5_MIN_CTAS_EQ 2          /* resident CTAs per SM the eq-plane kernel is compiled for.  With the CTA-lockstep tile loop the kernel is bound by
                                     dependent-issue latency, not by registers: profiles/r01x_sweep.log (phase A of cfg 2, ms): 512 x 1 (128 regs, 16
                                     warps/SM) 5.33, 640 x 1 (96) 4.95, 768 x 1 (80) 4.80, 896 x 1 (72) 4.76, 1024 x 1 (64) 4.72, 384 x 3 (56) 4.84,
                                     256 x 4 (64) 5.11, 512 x 2 (64 regs, 32 warps/SM, ~800 B of spills per thread served by L1) 4.63 */
#endif
#if !defined(S5_EQ_NO_SMEM_GD)
#define S5_EQ_SMEM_GD 1           /* the ray's geodesic struct lives in a per-thread shared-memory slot (pixel.cuh SGD): spill stores of the
                                     64-register build 1106 -> 482 B; cfg 2 without phi 4.14 -> 4.02 ms, cfg 3 2.50 -> 2.46 ms (profiles/r02d_sweep.log) */
#endif
#if defined(S5_EQ_SMEM_GD) && !defined(S5_EQ_FREERUN) && !defined(S5_EQ_NO_STAGE_SYNC)
#define S5_EQ_DYN_SMEM ((size_t)S5_EQ_THREADS * 200 + 64)     /* per-thread geodesic slots (S5_GD_SLOT_BYTES) */
#define S5_EQ_DYN_SMEM_ON 1
#else
#define S5_EQ_DYN_SMEM ((size_t)0)
#define S5_EQ_DYN_SMEM_ON 0
#endif
#ifndef S5_EQ_TILES_PER_SYNC
#define S5_EQ_TILES_PER_SYNC 1    /* tiles a warp traces between two CTA barriers of the lockstep tile loop */
#endif
#ifndef S5_MIN_CTAS_STEP
#define S5_MIN_CTAS_STEP 4        /* stepwise lane kernel, cfg 4 at 1024^2 after the angle carry took cr_acos out of the step (profiles/r05g_step_sweep.log, ms):
                                     2 CTAs/SM 397, 3 (160 regs, 12 warps/SM) 324, 4 (128 regs, 16 warps) 308, 5 (96) 320; 64-thread CTAs x 6: 325.
                                     (before: 1 CTA/SM 153.7, 3 139.5, 4 141.2, 5 156.6 at 512^2, profiles/r02w_step_sweep.log) */
#endif
#ifndef S5_MIN_CTAS_AZ
#define S5_MIN_CTAS_AZ 4
#endif

__device__ __forceinline__ void stage_consts(S5ImageConsts* dst, const S5ImageConsts* src)
{
    const int nwords = (int)(sizeof(S5ImageConsts) / sizeof(int));
    const int* s = reinterpret_cast<const int*>(src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) d[i] = s[i];
    __syncthreads();
}

/* Output planes and queue items are written once and never read back by the writing kernel: streaming stores (st.global.cs,
 * evict-first) keep them from displacing the kernels' register spills from L2 (64 registers per thread at 32 warps/SM: ~100 MB of
 * local memory in flight next to 3 GB of output per image). */
#if defined(S5_NO_STREAM_STORES)
#define S5_ST(ptr, v) (*(ptr) = (v))
#else
#define S5_ST(ptr, v) __stcs((ptr), (v))
#endif
__device__ __forceinline__ void store_pixel(const DevOut& out, unsigned outputs, size_t i, const PixelOut& o)
{
    if (outputs & SIM5_OUT_R)         S5_ST(&out.r[i], o.r);
    if (outputs & SIM5_OUT_PHI)       S5_ST(&out.phi[i], o.phi);
    if (outputs & SIM5_OUT_G)         S5_ST(&out.g[i], o.g);
    if (outputs & SIM5_OUT_FLUX)      S5_ST(&out.flux[i], o.flux);
    if (outputs & SIM5_OUT_CHI)       S5_ST(&out.chi[i], o.chi);
    if (outputs & SIM5_OUT_DELTA)     S5_ST(&out.delta[i], o.delta);
    if (outputs & SIM5_OUT_MUE)       S5_ST(&out.mue[i], o.mue);
    if (outputs & SIM5_OUT_INTENSITY) S5_ST(&out.intensity[i], o.intensity);
    if (outputs & SIM5_OUT_TAU)       S5_ST(&out.tau[i], o.tau);
    if (outputs & SIM5_OUT_QERR)      S5_ST(&out.qerr[i], o.qerr);
    if (outputs & SIM5_OUT_HEIGHT)    S5_ST(&out.height[i], o.height);
    if (outputs & SIM5_OUT_DELAY)     S5_ST(&out.delay[i], o.delay);
    if (outputs & SIM5_OUT_STEPS)     S5_ST(&out.steps[i], o.steps);
    if (outputs & SIM5_OUT_STATUS)    S5_ST(&out.status[i], (unsigned char)o.status);
}

__device__ __forceinline__ void flush_stats(const unsigned int* s_cnt, unsigned long long s_steps, DevStats* gs)
{
    __syncthreads();
    if (threadIdx.x < 32) { if (s_cnt[threadIdx.x]) atomicAdd(&gs->cls[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]); }
    else if (threadIdx.x < 40) { if (s_cnt[threadIdx.x]) atomicAdd(&gs->gtype[threadIdx.x - 32], (unsigned long long)s_cnt[threadIdx.x]); }
    (void)s_steps;
}

/* ------------------------------------------------------------------ */
/* modes EQPLANE / POLARIZED : analytic geodesic per pixel             */
/* ------------------------------------------------------------------ */
template <bool DEFER, bool DELAY = false>
__global__ void __launch_bounds__(S5_EQ_THREADS, S5_MIN_CTAS_EQ)
k_trace_eqplane(const __grid_constant__ S5ImageConsts gconsts, DevOut out, AzQueue q, unsigned long long* __restrict__ tile_counter, DevStats* __restrict__ gstats)
{
    __shared__ S5ImageConsts c;
    __shared__ unsigned int s_cnt[40];
    if (threadIdx.x < 40) s_cnt[threadIdx.x] = 0;
    stage_consts(&c, &gconsts);

    const int lane = threadIdx.x & 31;
    const long long nx = c.nx;
    const long long npix = (long long)c.nrows_local * nx;
    const long long ntiles = (npix + 31) >> 5;

#if !defined(S5_EQ_FREERUN)
    __shared__ unsigned long long s_tile;
#if defined(S5_EQ_PREFETCH_TILE)
    /* A/B (-DS5_EQ_PREFETCH_TILE): the index of the NEXT batch is fetched while the current one is traced (the global atomic's latency and one
     * of the two barriers per batch leave the critical path); a CTA claims one batch past the end, which nobody traces.  Measured: no
     * difference (phase A 3.602 vs 3.607 ms at 4096^2, 0.4987 vs 0.4995 ms on a 2 M-ray slice, profiles/r05z_proxy.log): the second CTA of
     * the SM already hides that latency.  Off by default. */
    unsigned long long nxt = 0;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, (unsigned long long)(S5_EQ_THREADS / 32 * S5_EQ_TILES_PER_SYNC));
    __syncthreads();
#endif
#endif
    for (;;) {
        unsigned long long t = 0;
#if !defined(S5_EQ_FREERUN)
        /* the CTA takes S5_EQ_TILES_PER_SYNC tiles per warp at a time and passes a barrier per batch, so its 16 warps walk the (183 KB)
         * routine together and share instruction-cache lines: r01q sweep 5.33 vs 5.40 ms free-running (cfg 3: 2.67 vs 2.81 ms);
         * -DS5_EQ_FREERUN restores per-warp pulls */
#if defined(S5_EQ_PREFETCH_TILE)
        const unsigned long long cur = s_tile;
        if ((long long)cur >= ntiles) break;
        if (threadIdx.x == 0) nxt = atomicAdd(tile_counter, (unsigned long long)(S5_EQ_THREADS / 32 * S5_EQ_TILES_PER_SYNC));
#define S5_CUR_TILE cur
#else
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, (unsigned long long)(S5_EQ_THREADS / 32 * S5_EQ_TILES_PER_SYNC));
        __syncthreads();
        if ((long long)s_tile >= ntiles) break;
#define S5_CUR_TILE s_tile
#endif
      #pragma unroll 1
      for (int sub = 0; sub < S5_EQ_TILES_PER_SYNC; sub++) {
        t = S5_CUR_TILE + (unsigned long long)(sub * (S5_EQ_THREADS / 32)) + (threadIdx.x >> 5);
#else
        if (lane == 0) t = atomicAdd(tile_counter, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((long long)t >= ntiles) break;
      {
#endif
        long long p = ((long long)t << 5) + lane;
        AzIn z;
        bool deferred = false;
        size_t i = 0;
#if S5_EQ_DYN_SMEM_ON
        extern __shared__ double s_dyn[];
        Geodesic* const gslot = reinterpret_cast<Geodesic*>(reinterpret_cast<char*>(s_dyn) + (size_t)threadIdx.x * S5_GD_SLOT_BYTES);
#endif
#if !defined(S5_EQ_FREERUN) && !defined(S5_EQ_NO_STAGE_SYNC)
        {   /* every thread runs the routine (it has CTA barriers); threads past the end of the image trace its last pixel and drop the result */
            const bool valid = (long long)t < ntiles && p < npix;
            const long long pc = valid ? p : npix - 1;
            int lr = (int)(pc / nx);
            int ix = (int)(pc - (long long)lr * nx);
            int iy = s5_local_to_image_row(&c, lr);
            PixelOut o;
#if defined(S5_EQ_SMEM_GD)
            deferred = trace_eqplane_pixel_t<DEFER, DELAY, true, true>(c, ix, iy, &o, &z, gslot) && valid;
#else
            deferred = trace_eqplane_pixel_t<DEFER, DELAY, true>(c, ix, iy, &o, &z) && valid;
#endif
            if (valid) {
                i = out.compact ? (size_t)p : (size_t)iy * (size_t)nx + (size_t)ix;
                store_pixel(out, c.outputs, i, o);
                atomicAdd(&s_cnt[o.status & 31], 1u);
                atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
            }
        }
#else
        if ((long long)t < ntiles && p < npix) {
            int lr = (int)(p / nx);
            int ix = (int)(p - (long long)lr * nx);
            int iy = s5_local_to_image_row(&c, lr);
            PixelOut o;
            deferred = trace_eqplane_pixel_t<DEFER, DELAY>(c, ix, iy, &o, &z);
            i = out.compact ? (size_t)p : (size_t)iy * (size_t)nx + (size_t)ix;
            store_pixel(out, c.outputs, i, o);
            atomicAdd(&s_cnt[o.status & 31], 1u);
            atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
        }
#endif
        if (DEFER) {
            /* hand the azimuth of this tile's hits to phase B: one aggregated atomic per warp and geodesic type */
            bool is_rr = deferred && z.type == GEOD_TYPE_RR;
            bool is_rc = deferred && z.type == GEOD_TYPE_RC;
            unsigned m_rr = __ballot_sync(0xffffffffu, is_rr);
            unsigned m_rc = __ballot_sync(0xffffffffu, is_rc);
            long long slot = -1;
            if (m_rr) {
                int leader = __ffs(m_rr) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(&q.count[0], (unsigned long long)__popc(m_rr));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (is_rr) slot = (long long)base + __popc(m_rr & ((1u << lane) - 1u));
            }
            if (m_rc) {
                int leader = __ffs(m_rc) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(&q.count[1], (unsigned long long)__popc(m_rc));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (is_rc) slot = q.cap - 1 - ((long long)base + __popc(m_rc & ((1u << lane) - 1u)));
            }
            if (slot >= 0) {
                double* f = q.f + slot;
                const long long cap = q.cap;
#if S5_EQ_DYN_SMEM_ON && !defined(S5_EQ_FREERUN) && !defined(S5_EQ_NO_STAGE_SYNC)
                /* the geodesic's share of the item goes from its shared-memory slot straight to the queue (az_make_tail filled the rest of z) */
                const Geodesic* g = gslot;
                S5_ST(&f[0 * cap], g->r1.re);  S5_ST(&f[1 * cap], g->r2.re);  S5_ST(&f[2 * cap], g->r3.re);   S5_ST(&f[3 * cap], is_rr ? g->r4.re : g->r3.im);
                S5_ST(&f[4 * cap], g->l);   S5_ST(&f[5 * cap], g->m2m); S5_ST(&f[6 * cap], g->m2p);
#else
                S5_ST(&f[0 * cap], z.e0);  S5_ST(&f[1 * cap], z.e1);  S5_ST(&f[2 * cap], z.e2);   S5_ST(&f[3 * cap], z.e3);
                S5_ST(&f[4 * cap], z.l);   S5_ST(&f[5 * cap], z.m2m); S5_ST(&f[6 * cap], z.m2p);
#endif
                S5_ST(&f[7 * cap], z.r); S5_ST(&f[8 * cap], z.K_mm); S5_ST(&f[9 * cap], z.rf_u); S5_ST(&f[10 * cap], z.isn_inf);
                S5_ST(&q.key[slot], az_key(i, z));
            }
        }
      }
#if !defined(S5_EQ_FREERUN) && defined(S5_EQ_PREFETCH_TILE)
        if (threadIdx.x == 0) s_tile = nxt;
        __syncthreads();
#endif
    }
    flush_stats(s_cnt, 0, gstats);
}

/* phase B: azimuth of the queued disk hits of ONE geodesic type (TYPE = GEOD_TYPE_RR or GEOD_TYPE_RC).
 * One CTA of S5_AZ_THREADS threads per SM; the CTA pulls S5_AZ_THREADS items at a time and passes a barrier per batch,
 * so its warps run the same routines at about the same time and share the SM's instruction cache lines
 * (the azimuth code is ~120 KB of SASS; with free-running warps ncu shows 74 % i-cache hit rate and
 * stall_no_instruction as large as the FP64 dependency stall). */
#ifndef S5_AZ_THREADS
#define S5_AZ_THREADS 512
#endif
template <int TYPE>
__global__ void __launch_bounds__(S5_AZ_THREADS, 1)
k_azimuth(const __grid_constant__ S5ImageConsts gconsts, AzQueue q, double* __restrict__ phi, unsigned long long* __restrict__ tile_counter, int from_redo)
{
    __shared__ unsigned long long s_base;
    const long long count = (long long)q.count[(from_redo ? 2 : 0) + (TYPE == GEOD_TYPE_RR ? 0 : 1)];
    const long long cap = q.cap;
    const double a_eff = fmax(1e-4, gconsts.a);
    const double cos_i = gconsts.cos_i;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(tile_counter, (unsigned long long)S5_AZ_THREADS);
        __syncthreads();
        long long base = (long long)s_base;
        if (base >= count) break;
        long long it = base + threadIdx.x;
        const bool valid = it < count;
        if (!valid) it = count - 1;          /* every thread runs the routine (it has barriers); the surplus ones redo the last item */
        {
            long long slot = (TYPE == GEOD_TYPE_RR) ? it : cap - 1 - it;
            if (from_redo) slot = (long long)q.redo[slot];
            const double* f = q.f + slot;
            AzIn z;
            z.e0 = f[0 * cap];  z.e1 = f[1 * cap];  z.e2 = f[2 * cap];   z.e3 = f[3 * cap];
            z.l = f[4 * cap];   z.m2m = f[5 * cap]; z.m2p = f[6 * cap];  z.r = f[7 * cap];
            z.K_mm = f[8 * cap]; z.rf_u = f[9 * cap]; z.isn_inf = f[10 * cap];
            z.mm = z.m2p / (z.m2p + z.m2m);          /* == g->mm (geodesic_T_roots, q > 0) */
            unsigned long long key = q.key[slot];
            z.a = a_eff; z.cos_i = cos_i;
            z.type = TYPE;
            az_unkey(key, &z);
            double v = azimuth_from_t<true>(z);
            if (valid) phi[key & 0xffffffffffffULL] = v;
        }
    }
}

/* phase B, tolerance mode (the default): azimuth of the queued hits with azimuth_fast_rr / azimuth_fast_rc (pixel.cuh).
 * One launch covers both queues (RR items 0..n_rr-1 from the front, then the RC items from the back; a warp is of one
 * type except the single warp at the boundary).  The routines are ~20 KB of SASS each, so free-running warps stay inside
 * the instruction cache and no CTA barriers are needed; items of a type cost the same to within one duplication step, so
 * a static grid-stride split is balanced.  Items whose arguments leave the fast routines' domain go to the redo lists and
 * are integrated by the bit-faithful k_azimuth<TYPE> afterwards. */
#ifndef S5_AZF_THREADS
#define S5_AZF_THREADS 512        /* 512 x 2 CTAs/SM at 64 registers (32 warps/SM, 300 B of spills): round-2 sweeps on the final routine (profiles/r06a_proxy.log,
                                     r06b_proxy.log; ms at 4096^2 / on a 2 M-ray slice, same box): 256 x 2 (128 regs, no spills) 2.69 / 0.348, 256 x 3 (80) 2.64 / 0.350,
                                     384 x 2 (80) 2.63 / 0.340, 416 x 2 (72) 2.74 / 0.364, 448 x 2 (72) 2.57 / 0.355, 512 x 2 (64) 2.50 / 0.340; 192 x 3 and 224 x 2
                                     are slower than 256 x 2.  (round 1, on the 4-sequence routine: 256 x 2 3.14 ms, 128 x 4 3.32, 128 x 6 3.31, 128 x 3 3.46) */
#endif
#ifndef S5_MIN_CTAS_AZF
#define S5_MIN_CTAS_AZF 2
#endif
/* WHICH: 0 both queues in one launch, 1 the RR queue only, 2 the RC queue only */
template <int WHICH>
__global__ void __launch_bounds__(S5_AZF_THREADS, S5_MIN_CTAS_AZF)
k_azimuth_fast(const __grid_constant__ S5ImageConsts gconsts, AzQueue q, double* __restrict__ phi)
{
    const long long n_rr = (WHICH == 2) ? 0 : (long long)q.count[0];
    const long long count = n_rr + ((WHICH == 1) ? 0 : (long long)q.count[1]);
    const long long cap = q.cap;
    const double a_eff = fmax(1e-4, gconsts.a);
    const double cos_i = gconsts.cos_i;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < count; it += stride) {
        const bool is_rr = (WHICH == 1) || (WHICH == 0 && it < n_rr);
        const long long slot = is_rr ? it : cap - 1 - (it - n_rr);
        const double* f = q.f + slot;
        AzIn z;
        z.e0 = f[0 * cap];  z.e1 = f[1 * cap];  z.e2 = f[2 * cap];   z.e3 = f[3 * cap];
        z.l = f[4 * cap];   z.m2m = f[5 * cap]; z.m2p = f[6 * cap];  z.r = f[7 * cap];
#if defined(S5_POLAR_DUPLICATION)
        z.K_mm = f[8 * cap]; z.mm = z.m2p / (z.m2p + z.m2m);
#else
        z.K_mm = 0.0; z.mm = 0.0;                    /* not used by the tolerance-mode routines (the complete Pi is an AGM of its own) */
#endif
        z.rf_u = 0.0; z.isn_inf = 0.0;
        unsigned long long key = q.key[slot];
        z.a = a_eff; z.cos_i = cos_i;
        z.type = is_rr ? GEOD_TYPE_RR : GEOD_TYPE_RC;
        az_unkey(key, &z);
        z.rf_ok = false;
        bool ok;
        double v = is_rr ? azimuth_fast_rr(z, &ok) : azimuth_fast_rc(z, &ok);
        if (ok) {
            phi[key & 0xffffffffffffULL] = v;
        } else if (is_rr) {
            q.redo[atomicAdd(&q.count[2], 1ULL)] = (unsigned)slot;
        } else {
            q.redo[cap - 1 - (long long)atomicAdd(&q.count[3], 1ULL)] = (unsigned)slot;
        }
    }
}

/* ------------------------------------------------------------------ */
/* modes STEPWISE and SURFACE : one live ray per lane, warp-level refill */
/* ------------------------------------------------------------------ */
/* PROG = StepwiseProg (a raytrace() call + the torus per step) or SurfaceProg (one pass of geodesic_follow's loop per step) */
#define S5_STEPS_PER_ROUND 16
#define S5_REFILL_MIN 4

/* n rays from the queue: the kernel's own counter, or (SIM5_FLAG_SHARED_QUEUE) one word that several GPUs pull from over NVLink */
__device__ __forceinline__ unsigned long long take_rays(unsigned long long* counter, unsigned long long n, bool shared)
{
    return shared ? atomicAdd_system(counter, n) : atomicAdd(counter, n);
}

template <class PROG>
__global__ void __launch_bounds__(PROG::THREADS, PROG::MIN_CTAS)
k_trace_lanes(const __grid_constant__ S5ImageConsts gconsts, DevOut out, unsigned long long* __restrict__ ray_counter, DevStats* __restrict__ gstats)
{
    __shared__ S5ImageConsts c;
    __shared__ unsigned int s_cnt[40];
    __shared__ unsigned long long s_steps;
    if (threadIdx.x < 40) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_steps = 0;
    stage_consts(&c, &gconsts);

    const int lane = threadIdx.x & 31;
    const long long nx = c.nx;
    const long long npix = (long long)c.nrows_local * nx;
    const bool refill = !(c.flags & SIM5_FLAG_NO_REFILL);
    const bool shared_queue = (c.flags & SIM5_FLAG_SHARED_QUEUE) != 0;

    typename PROG::State s;
    long long mypix = -1;
    bool live = false;
    bool pend = false;             /* PROG::DEFERS: this lane's step waits for its expensive part */
    bool drained = false;          /* queue exhausted (warp-uniform) */
    unsigned long long my_steps = 0;

    if (PROG::BATCH > 0) {
        /* CTA-batch variant: the CTA takes blockDim.x consecutive pixels, all lanes start together and the CTA passes a barrier every
         * PROG::BATCH steps until its last ray has finished -- the warps then share instruction-cache lines (the step routine runs from
         * L2 otherwise) at the price of waiting for the slowest ray of the batch.  For programs whose neighbouring rays take similar
         * numbers of steps. */
        __shared__ unsigned long long s_base;
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_base = take_rays(ray_counter, (unsigned long long)blockDim.x, shared_queue);
            __syncthreads();
            const long long base = (long long)s_base;
            if (base >= npix) break;
            const long long p = base + threadIdx.x;
            live = false;
            if (p < npix) {
                int lr = (int)(p / nx);
                int ix = (int)(p - (long long)lr * nx);
                int iy = s5_local_to_image_row(&c, lr);
                PixelOut o;
                mypix = p;
                if (PROG::start(c, ix, iy, &s, &o)) {
                    live = true;
                } else {
                    size_t i = out.compact ? (size_t)p : (size_t)iy * (size_t)nx + (size_t)ix;
                    store_pixel(out, c.outputs, i, o);
                    atomicAdd(&s_cnt[o.status & 31], 1u);
                    atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
                }
            }
            for (;;) {
                #pragma unroll 1
                for (int it = 0; it < (PROG::BATCH > 0 ? PROG::BATCH : 1); it++) {
                    if (live) {
                        int cls = PROG::step(c, &s);
                        if (cls) {
                            PixelOut o;
                            PROG::finish(c, &s, cls, &o);
                            int lr = (int)(mypix / nx);
                            int ix = (int)(mypix - (long long)lr * nx);
                            int iy = s5_local_to_image_row(&c, lr);
                            size_t i = out.compact ? (size_t)mypix : (size_t)iy * (size_t)nx + (size_t)ix;
                            store_pixel(out, c.outputs, i, o);
                            atomicAdd(&s_cnt[o.status & 31], 1u);
                            atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
                            my_steps += (unsigned long long)o.steps;
                            live = false;
                        }
                    }
                }
                if (!__syncthreads_or(live ? 1 : 0)) break;
            }
        }
    } else
    for (;;) {
        unsigned idle = __ballot_sync(0xffffffffu, !live);
        if (idle == 0xffffffffu && drained) break;
        /* refill finished lanes: always when the whole warp is idle; otherwise once enough lanes wait */
        bool do_fill = !drained && (idle == 0xffffffffu || (refill && __popc(idle) >= PROG::REFILL_MIN));
        if (do_fill) {
            int nreq = __popc(idle);
            unsigned long long base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = take_rays(ray_counter, (unsigned long long)nreq, shared_queue);
            base = __shfl_sync(0xffffffffu, base, leader);
            if ((long long)base >= npix) drained = true;
            if (!live) {
                long long p = (long long)base + __popc(idle & ((1u << lane) - 1u));
                if (p < npix) {
                    int lr = (int)(p / nx);
                    int ix = (int)(p - (long long)lr * nx);
                    if (PROG::CENTER_OUT && !(c.flags & SIM5_FLAG_ROW_MAJOR)) {
                        /* longest rays first: the k-th row handed out is the k-th closest to the middle of this call's rows (mid, mid-1, mid+1, ...).
                         * The rays that take the most steps pass closest to the hole, i.e. sit in the middle rows; started last they ARE the tail of the
                         * kernel (a ray cannot be split: 8000 steps x ~7 us), started first the cheap outer rows fill in behind them */
                        const int nr = c.nrows_local, mid = nr >> 1;
                        lr = (lr & 1) ? mid - ((lr + 1) >> 1) : mid + (lr >> 1);
                        p = (long long)lr * nx + ix;
                    }
                    int iy = s5_local_to_image_row(&c, lr);
                    PixelOut o;
                    mypix = p;
                    if (PROG::start(c, ix, iy, &s, &o)) {
                        live = true;
                    } else {
                        size_t i = out.compact ? (size_t)p : (size_t)iy * (size_t)nx + (size_t)ix;
                        store_pixel(out, c.outputs, i, o);
                        atomicAdd(&s_cnt[o.status & 31], 1u);
                        atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
                    }
                }
            }
            if ((long long)base + nreq >= npix) drained = true;
        }
        /* advance the live lanes */
        #pragma unroll 1
        for (int it = 0; it < S5_STEPS_PER_ROUND; it++) {
            bool stepped = live;
            if (PROG::DEFERS) {
                /* the cheap part of the step for every lane that is not waiting; the expensive part (PROG::slow_step) for all waiting lanes at
                 * once, as soon as PROG::DEFER_MIN of them wait or nothing else is left to do in this warp */
                stepped = false;
                if (live && !pend) { pend = PROG::try_step(c, &s); stepped = !pend; }
                const unsigned pm = __ballot_sync(0xffffffffu, live && pend);
                const unsigned rm = __ballot_sync(0xffffffffu, live && !pend);
                if (pm && (__popc(pm) >= PROG::DEFER_MIN || rm == 0)) {
                    if (live && pend) { PROG::slow_step(c, &s); pend = false; stepped = true; }
                }
            }
            if (stepped) {
                int cls = PROG::DEFERS ? PROG::post(c, &s) : PROG::step(c, &s);
                if (cls) {
                    PixelOut o;
                    PROG::finish(c, &s, cls, &o);
                    int lr = (int)(mypix / nx);
                    int ix = (int)(mypix - (long long)lr * nx);
                    int iy = s5_local_to_image_row(&c, lr);
                    size_t i = out.compact ? (size_t)mypix : (size_t)iy * (size_t)nx + (size_t)ix;
                    store_pixel(out, c.outputs, i, o);
                    atomicAdd(&s_cnt[o.status & 31], 1u);
                    atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
                    my_steps += (unsigned long long)o.steps;
                    live = false;
                }
            }
            if (refill && !__any_sync(0xffffffffu, live)) break;
        }
    }
    /* total step count */
    for (int off = 16; off > 0; off >>= 1) my_steps += __shfl_down_sync(0xffffffffu, my_steps, off);
    if (lane == 0 && my_steps) atomicAdd(&s_steps, my_steps);
    flush_stats(s_cnt, 0, gstats);
    if (threadIdx.x == 0 && s_steps) atomicAdd(&gstats->steps, s_steps);
}

/* ------------------------------------------------------------------ */
/* mode HISTOGRAM : g-factor transfer function over a (spin, incl) lattice */
/* ------------------------------------------------------------------ */
/* Same execution shape as k_trace_eqplane: CTAs of S5_EQ_THREADS threads in lockstep, one tile per warp and batch, the staged pixel
 * routine with its CTA barriers and the geodesic in shared-memory slots (r02f sweep: the former 128-thread free-running kernel
 * traced 2.5e9 rays/s, 2.8e9 at 64 registers).  A batch never straddles two lattice images, so the staged constants stay valid. */
/* The launch traces the n_img lattice images img_first, img_first + img_stride, ... (one GPU's share of an interleaved deal of the
 * lattice over split_count GPUs: neighbouring images -- neighbouring inclinations of one spin -- cost about the same, so every GPU
 * gets the same mix). */
__global__ void __launch_bounds__(S5_EQ_THREADS, S5_MIN_CTAS_EQ)
k_trace_histogram(const S5ImageConsts* __restrict__ gconsts /* one per lattice image */, int img_first, int img_stride, int n_img,
                  double* __restrict__ hist, unsigned long long* __restrict__ tile_counter, DevStats* __restrict__ gstats)
{
    __shared__ S5ImageConsts c;
    __shared__ unsigned int s_cnt[40];
    __shared__ int s_img;
    if (threadIdx.x < 40) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_img = -1;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wpc = S5_EQ_THREADS / 32;                 /* tiles per batch */
    /* all images share nx, ny */
    const long long nx = gconsts[img_first].nx;
    const long long npix = (long long)gconsts[img_first].ny * nx;
    const long long tiles_per_img = (npix + 31) >> 5;
    const long long chunks_per_img = (tiles_per_img + wpc - 1) / wpc;
    const long long nchunks = chunks_per_img * (long long)n_img;
    __shared__ unsigned long long s_chunk;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(tile_counter, 1ULL);
        __syncthreads();
        long long ch = (long long)s_chunk;
        if (ch >= nchunks) break;
        int img = img_first + img_stride * (int)(ch / chunks_per_img);
        long long t = (ch % chunks_per_img) * wpc + (threadIdx.x >> 5);
        if (img != s_img) {
            __syncthreads();
            stage_consts(&c, gconsts + img);
            if (threadIdx.x == 0) s_img = img;
            __syncthreads();
        }
        long long p = (t << 5) + lane;
        bool hit = false;
        int bin = -1;
        double w = 0.0;
        {
            const bool valid = t < tiles_per_img && p < npix;
            const long long pc = valid ? p : npix - 1;
            int iy = (int)(pc / nx);
            int ix = (int)(pc - (long long)iy * nx);
            PixelOut o;
            AzIn z;                                     /* never filled: the lattice images carry no phi (DEFER only keeps the azimuth code out) */
#if S5_EQ_DYN_SMEM_ON
            extern __shared__ double s_dyn[];
            Geodesic* gslot = reinterpret_cast<Geodesic*>(reinterpret_cast<char*>(s_dyn) + (size_t)threadIdx.x * S5_GD_SLOT_BYTES);
            trace_eqplane_pixel_t<true, false, true, true>(c, ix, iy, &o, &z, gslot);
#else
            trace_eqplane_pixel_t<true, false, true>(c, ix, iy, &o, &z);
#endif
            if (valid) {
                atomicAdd(&s_cnt[o.status & 31], 1u);
                atomicAdd(&s_cnt[32 + ((o.status >> 5) & 7)], 1u);
                unsigned cls = o.status & 31;
                if (cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2) {
                    double tt = (o.g - c.g_min) / (c.g_max - c.g_min) * (double)c.n_bins;
                    if (tt >= 0.0 && tt < (double)c.n_bins) { hit = true; bin = (int)tt; w = o.flux * c.da * c.db; }
                }
            }
        }
        /* warp-aggregated accumulation: one atomic per distinct bin in the warp */
        unsigned active = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            unsigned peers = __match_any_sync(active, bin);
            int leader = __ffs(peers) - 1;
            double sum = 0.0;
            /* fixed lane order inside the group */
            for (unsigned m = peers; m; m &= m - 1) {
                int src = __ffs(m) - 1;
                sum += __shfl_sync(peers, w, src);
            }
            if (lane == leader) atomicAdd(&hist[(size_t)img * c.n_bins + bin], sum);
        }
    }
    flush_stats(s_cnt, 0, gstats);
}

/* Reduction of the lattice histograms of a multi-GPU call (sim5_trace_image_multi): the GPU that owns the result adds up the partial
 * histograms of all GPUs with plain loads from their memory (peer access over NVLink / NVSwitch; 4 MB per GPU) in a FIXED order, so the
 * reduced histogram does not depend on timing.  dst may alias src.p[0]. */
struct PeerPtrs { const double* p[16]; };
__global__ void __launch_bounds__(256) k_sum_peers(double* dst, PeerPtrs src, int nsrc, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = src.p[0][i];
        for (int k = 1; k < nsrc; k++) s += src.p[k][i];
        dst[i] = s;
    }
}

/* ------------------------------------------------------------------ */
/* mode SPECTRUM : thermal disk spectrum summed over the image          */
/* ------------------------------------------------------------------ */
/* Phase A per pixel (polarized flavour: g and mu_e from the emitter frame), then the black-body sum.  The sum is done
 * TRANSPOSED inside the warp: lane j owns the energies k = j, j+32, ... (<= 8 per lane) and walks the warp's 32 hits, whose
 * three per-hit numbers arrive by shuffle -- so every lane accumulates in registers and there is no per-term reduction.
 * Lanes flush once per kernel: shared-memory atomics per CTA, then one global atomic per CTA and energy. */
#define S5_SPEC_MAX_E 256
#if defined(S5_SPEC_FREERUN)
#define S5_SPEC_THREADS S5_CTA_THREADS
#define S5_SPEC_CTAS 4
#define S5_SPEC_DYN_SMEM ((size_t)0)
#else
/* the same execution shape as k_trace_eqplane / k_trace_histogram: CTAs of S5_EQ_THREADS threads in lockstep, one tile per warp and
 * batch, the staged pixel routine with the geodesic in shared-memory slots (the free-running 128-thread kernel was bound by
 * instruction supply: stall_no_instruction 2.7 per issue, FP64 pipe 36 % active, profiles/r04b_ncu_modes_summary.csv) */
#define S5_SPEC_THREADS S5_EQ_THREADS
#define S5_SPEC_CTAS S5_MIN_CTAS_EQ
#define S5_SPEC_DYN_SMEM S5_EQ_DYN_SMEM
#endif
__global__ void __launch_bounds__(S5_SPEC_THREADS, S5_SPEC_CTAS)
k_trace_spectrum(const __grid_constant__ S5ImageConsts gconsts, const double* __restrict__ energies, double* __restrict__ spec,
                 unsigned long long* __restrict__ tile_counter, DevStats* __restrict__ gstats)
{
    __shared__ S5ImageConsts c;
    __shared__ unsigned int s_cnt[40];
    __shared__ double s_spec[S5_SPEC_MAX_E];
    if (threadIdx.x < 40) s_cnt[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < S5_SPEC_MAX_E; i += blockDim.x) s_spec[i] = 0.0;
    stage_consts(&c, &gconsts);

    const int lane = threadIdx.x & 31;
    const long long nx = c.nx;
    const long long npix = (long long)c.nrows_local * nx;
    const long long ntiles = (npix + 31) >> 5;
    const int ne = c.n_energy;
    double Ek[S5_SPEC_MAX_E / 32], acc[S5_SPEC_MAX_E / 32];
    #pragma unroll
    for (int j = 0; j < S5_SPEC_MAX_E / 32; j++) { int k = lane + 32 * j; Ek[j] = (k < ne) ? energies[k] : 1.0; acc[j] = 0.0; }
#if !defined(S5_SPEC_FREERUN)
    __shared__ unsigned long long s_tile;
#endif
    for (;;) {
        unsigned long long t = 0;
        SpecHit h;
        h.amp3 = 0.0; h.ginv = 1.0; h.xs = 1.0;
        bool hit = false;
#if !defined(S5_SPEC_FREERUN)
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, (unsigned long long)(S5_SPEC_THREADS / 32));
        __syncthreads();
        if ((long long)s_tile >= ntiles) break;
        t = s_tile + (threadIdx.x >> 5);
        {
            long long p = ((long long)t << 5) + lane;
            const bool valid = (long long)t < ntiles && p < npix;
            const long long pc = valid ? p : npix - 1;
            int lr = (int)(pc / nx);
            int ix = (int)(pc - (long long)lr * nx);
            int iy = s5_local_to_image_row(&c, lr);
            unsigned status;
#if S5_EQ_DYN_SMEM_ON
            extern __shared__ double s_dyn[];
            Geodesic* gslot = reinterpret_cast<Geodesic*>(reinterpret_cast<char*>(s_dyn) + (size_t)threadIdx.x * S5_GD_SLOT_BYTES);
            hit = spectrum_pixel_t<true>(c, ix, iy, &h, &status, gslot) && valid;
#else
            hit = spectrum_pixel_t<false>(c, ix, iy, &h, &status, nullptr) && valid;
#endif
            if (valid) {
                atomicAdd(&s_cnt[status & 31], 1u);
                atomicAdd(&s_cnt[32 + ((status >> 5) & 7)], 1u);
            }
        }
#else
        if (lane == 0) t = atomicAdd(tile_counter, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((long long)t >= ntiles) break;
        long long p = ((long long)t << 5) + lane;
        if (p < npix) {
            int lr = (int)(p / nx);
            int ix = (int)(p - (long long)lr * nx);
            int iy = s5_local_to_image_row(&c, lr);
            unsigned status;
            hit = spectrum_pixel(c, ix, iy, &h, &status);
            atomicAdd(&s_cnt[status & 31], 1u);
            atomicAdd(&s_cnt[32 + ((status >> 5) & 7)], 1u);
        }
#endif
        unsigned hits = __ballot_sync(0xffffffffu, hit);
        for (unsigned m = hits; m; m &= m - 1) {
            int src = __ffs(m) - 1;
            SpecHit hs;
            hs.amp3 = __shfl_sync(0xffffffffu, h.amp3, src);
            hs.ginv = __shfl_sync(0xffffffffu, h.ginv, src);
            hs.xs = __shfl_sync(0xffffffffu, h.xs, src);
            #pragma unroll
            for (int j = 0; j < S5_SPEC_MAX_E / 32; j++)
                if (lane + 32 * j < ne) acc[j] += spectrum_term(hs, Ek[j]);
        }
    }
    #pragma unroll
    for (int j = 0; j < S5_SPEC_MAX_E / 32; j++)
        if (lane + 32 * j < ne && acc[j] != 0.0) atomicAdd(&s_spec[lane + 32 * j], acc[j]);
    __syncthreads();
    for (int k = threadIdx.x; k < ne; k += blockDim.x)
        if (s_spec[k] != 0.0) atomicAdd(&spec[k], s_spec[k]);
    flush_stats(s_cnt, 0, gstats);
}

/* ------------------------------------------------------------------ */
/* FP64 DFMA-chain microbenchmark (roofline denominator)               */
/* ------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, b = 1e-9;
    #pragma unroll 1
    for (int i = 0; i < iters; i++) {
        #pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __fma_rn(a0, m, b); a1 = __fma_rn(a1, m, b); a2 = __fma_rn(a2, m, b); a3 = __fma_rn(a3, m, b);
            a4 = __fma_rn(a4, m, b); a5 = __fma_rn(a5, m, b); a6 = __fma_rn(a6, m, b); a7 = __fma_rn(a7, m, b);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;      /* keep the chains alive */
}

/* micro-benchmark of one device routine in isolation (arguments generated in registers, one store per thread):
 * the ceiling a routine reaches without the rest of the pixel pipeline around it */
__global__ void __launch_bounds__(128) k_micro(int which, int reps, double* out)
{
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    /* cheap per-thread arguments in the ranges the azimuth path uses */
    double u = (double)((i * 2654435761ULL) & 0xfffff) * (1.0 / 1048576.0);      /* [0,1) */
    double v = (double)((i * 40503ULL + 12345ULL) & 0xfffff) * (1.0 / 1048576.0);
    double c2 = 0.05 + 0.9 * u, m = 0.1 + 0.8 * v;
    double acc = 0.0;
    for (int r = 0; r < reps; r++) {
        double q = 1.0 - (1.0 - c2) * m;
        switch (which) {
            case 0: acc += rf(c2, q, 1.0); break;
            case 1: acc += rj(c2, q, 1.0, 1.0 + 0.7 * (1.0 - c2)); break;
            case 2: acc += rc(c2, q); break;
            case 3: { double sn, cn, dn; jacobi_sncndn(1.5 * u, m, &sn, &cn, &dn); acc += sn + cn + dn; break; }
            case 4: { double sn, cn; crm::cr_sincos(3.0 * u, &sn, &cn); acc += sn + cn; break; }
            case 5: acc += crm::cr_log(0.5 + 4.0 * u); break;
            case 6: acc += crm::cr_atan2(u - 0.5, v - 0.3); break;
            case 7: acc += crm::cr_pow_third(0.1 + 100.0 * u); break;
            case 8: acc += (1.0 + u) / (1.0 + v) + (2.0 + u) / (3.0 + v) + (0.5 + v) / (1.5 + u) + (4.0 + v) / (1.1 + u); break;   /* 4 independent divisions */
            case 9: acc += sqrt(1.0 + u) + sqrt(2.0 + v) + sqrt(0.5 + u) + sqrt(3.0 + v); break;                                   /* 4 independent square roots */
            case 10: acc += rj(0.0, 1.0 - m, 1.0, 1.0 + 0.3 * u); break;
        }
        c2 += 1e-6; m += 1e-7; u += 1e-6; v += 1e-6;
    }
    out[i] = acc;
}

/* ------------------------------------------------------------------ */
/* element-wise batch kernels (unit parity tests through the C-ABI)    */
/* ------------------------------------------------------------------ */
__global__ void k_batch_rf(long long n, const double* x, const double* y, const double* z, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = rf(x[i], y[i], z[i]); }
__global__ void k_batch_rd(long long n, const double* x, const double* y, const double* z, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = rd(x[i], y[i], z[i]); }
__global__ void k_batch_rc(long long n, const double* x, const double* y, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = rc(x[i], y[i]); }
__global__ void k_batch_rj(long long n, const double* x, const double* y, const double* z, const double* p, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = rj(x[i], y[i], z[i], p[i]); }
/* tolerance-mode Carlson routines of ellfast.cuh (arguments outside their domain give NaN here; the kernels hand such items to the bit-faithful path) */
__global__ void k_batch_rf_hi(long long n, const double* x, const double* y, const double* z, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = hi_domain(x[i], y[i], z[i]) ? rf_hi(x[i], y[i], z[i]) : NAN; }
__global__ void k_batch_rj_hi(long long n, const double* x, const double* y, const double* z, const double* p, double* o)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = (hi_domain(x[i], y[i], z[i]) && hi_domain_p(p[i])) ? rj_hi(x[i], y[i], z[i], p[i]) : NAN; }
__global__ void k_batch_sncndn(long long n, const double* u, const double* m, double* sn, double* cn, double* dn)
{ for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) jacobi_sncndn(u[i], m[i], &sn[i], &cn[i], &dn[i]); }
__global__ void k_batch_libm(int op, long long n, const double* a, const double* b, double* o)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v;
        switch (op) {
            case 0: v = crm::cr_sin(a[i]); break;
            case 1: v = crm::cr_cos(a[i]); break;
            case 2: v = crm::cr_log(a[i]); break;
            case 3: v = crm::cr_atan2(a[i], b[i]); break;
            case 4: v = crm::cr_acos(a[i]); break;
            case 5: v = crm::cr_asin(a[i]); break;
            case 6: v = crm::cr_atan(a[i]); break;
            case 7: v = crm::cr_pow_third(a[i]); break;
            case 8: v = crm::cr_pow_1p5(a[i]); break;
            case 9: v = crm::cr_pow_4(a[i]); break;
            case 10: v = exp(a[i]); break;
            default: v = NAN;
        }
        o[i] = v;
    }
}

/* the integrals behind geodesic_timedelay (op codes: sim5_b200.h sim5_batch_integral) */
__global__ void k_batch_integral(int op, long long n, const double* v0, const double* v1, const double* v2, const double* v3,
                                 const double* v4, const double* v5, const double* v6, double* o)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double r;
        switch (op) {
            case 0: r = integral_C1(v0[i], v1[i]); break;
            case 1: r = integral_C2(v0[i], v1[i]); break;
            case 2: r = integral_C2_cos(v0[i], v1[i]); break;
            case 3: r = integral_Z2(v0[i], v1[i], v2[i], v3[i]); break;
            case 4: r = integral_Rm1(v0[i], v1[i], v2[i]); break;
            case 5: r = integral_Rm2(v0[i], v1[i], v2[i]); break;
            case 6: r = integral_R2(v0[i], v1[i], v2[i]); break;
            case 7: r = integral_R_r0_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 8: r = integral_R_r0_re_inf(v0[i], v1[i], v2[i], v3[i]); break;
            case 9: r = integral_R_r1_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 10: r = integral_R_r2_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 11: r = integral_T_m0(v0[i], v1[i], v2[i]); break;
            case 12: r = integral_T_m2(v0[i], v1[i], v2[i]); break;
            case 13: r = integral_R_r0_cc(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 14: r = integral_R_r0_cc_inf(v0[i], v1[i], v2[i], v3[i]); break;
            case 15: r = integral_R_r1_cc(v0[i], v1[i], v2[i], v3[i], v4[i], v5[i]); break;
            case 16: r = integral_R_r2_cc(v0[i], v1[i], v2[i], v3[i], v4[i], v5[i]); break;
            case 17: r = integral_R_rp_cc2(v0[i], v1[i], v2[i], v3[i], v6[i], v4[i], v5[i]); break;
            default: r = NAN;
        }
        o[i] = r;
    }
}
/* geodesic_timedelay between two radii on the way in, for n geodesics from infinity (sin_i, cos_i: host libm, as every image call) */
__global__ void k_batch_timedelay(long long n, double incl, double sin_i, double cos_i, double a, const double* alpha, const double* beta,
                                  const double* ra, const double* rb, double* o)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Geodesic gd;
        int error = 0;
        double r = NAN;
        if (geodesic_init_inf_sc(incl, sin_i, cos_i, a, alpha[i], beta[i], &gd, &error)) {
            double Pa = geodesic_P_int(&gd, ra[i], 0), Pb = geodesic_P_int(&gd, rb[i], 0);
            r = geodesic_timedelay(&gd, Pa, 0.0, 0.0, Pb, 0.0, 0.0);
        }
        o[i] = r;
    }
}

} /* namespace s5 */
#endif
