/*
 * capi.cu -- the C-ABI of libsim5b200.so: context, the batched sim5_trace_image entry, batched
 * element-wise entries, the FP64 peak microbenchmark, and the scalar sim5lib.h API executed as
 * one-thread device launches.
 *
 * No function in this file computes ray physics on the host.  Host code only (a) evaluates the
 * per-image constants of image_consts.h with the host libm, (b) moves bytes, (c) launches kernels.
 * Without a usable CUDA device every entry fails (SIM5_ERR_NO_DEVICE / NaN / FALSE) and says so on
 * stderr once.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sim5_b200.h"
#include "kernels.cuh"

using s5::DevOut;
using s5::DevStats;
using s5::AzQueue;

/* ------------------------------------------------------------------ */
/* context                                                             */
/* ------------------------------------------------------------------ */
namespace {

#ifndef S5_DEFER_REDO_CTAS
#define S5_DEFER_REDO_CTAS 32      /* grid of a deferred redo pass: it runs beside the next call's tracing kernel */
#endif
#define S5_MAX_CHUNKS 32
#define S5_BLK_WORDS (8 + sizeof(DevStats) / sizeof(unsigned long long))      /* counters + stats of one call, in 64-bit words */
#define S5_CHUNK_RAYS (1 << 21)     /* rays per chunk of a host-plane call: ~1.1 ms of kernels, ~1.2 ms of PCIe.  4096^2 r/phi/g/flux/status end to end
                                       (profiles/r01x_sweep.log, ms): 2^18 13.6, 2^19 13.6, 2^20 11.9, 2^21 11.4, 2^22 12.2, 2^23 14.4 */

struct Scratch {                  /* mapped pinned memory shared by host and device for scalar calls */
    s5::Geodesic g;
    s5::RayData rtd;
    s5::Metric m;
    s5::Tetrad t;
    double v[96];
    int iv[8];
};

struct Plane { void* p = nullptr; size_t bytes = 0; };

struct Context {
    bool ready = false;
    bool warned = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, own_stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;      /* the RC azimuth chain runs on aux_stream beside the RR chain */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t evp[3] = {nullptr, nullptr, nullptr};   /* phase boundaries of the last image call: after phase A, after azimuth RR, after azimuth RC */
    /* history of the phase events of the last S5_RING image calls ({ev1, evp[0..2], ev2} alias the current slot), so a caller that
     * enqueues many SIM5_FLAG_ASYNC calls back to back can read every call's per-kernel times afterwards without a sync per call */
#define S5_RING 64
    cudaEvent_t ring[S5_RING][5] = {{nullptr}};
    int ring_phases[S5_RING] = {0};
    bool ring_defer[S5_RING] = {false};      /* the slot's call left its redo passes on the auxiliary stream: it ends at evp[1], no third phase */
    bool last_defer = false;
    int ring_pos = 0;
    cudaEvent_t ev_chunk[S5_MAX_CHUNKS] = {nullptr};      /* chunk k traced -> its device->host copy may start */
    cudaEvent_t ev_copy_done = nullptr;
    long long chunk_rays = 0;             /* 0: S5_CHUNK_RAYS */
    int phases = 0;                       /* kernels launched by the last single-chunk image call: 0 none recorded, 1 trace only, 3 or 4 trace + azimuth kernels */
    Scratch* h_scr = nullptr;
    Scratch* d_scr = nullptr;
    S5ImageConsts* h_consts = nullptr;    /* pinned staging */
    S5ImageConsts* d_consts = nullptr;
    size_t consts_cap = 0;
    unsigned long long* d_counter = nullptr;
    DevStats* d_stats = nullptr;
    DevStats* h_stats = nullptr;          /* pinned */
    Plane planes[SIM5_NPLANES];
    Plane hist;
    Plane azq_redo;
    Plane azq_f, azq_key;                 /* azimuth work-item queue (phase A -> phase B) */
    /* SIM5_FLAG_DEFER_REDO: a second queue and two counter blocks of its own, so the redo passes of call k can run beside call k+1 */
    Plane azq2_f, azq2_key, azq2_redo;
    Plane shared_q;                       /* sim5_trace_image_multi + SIM5_FLAG_SHARED_QUEUE: the ray queue all devices pull from (on devices[0]) */
    unsigned long long* d_counter2 = nullptr;             /* [2][8] */
    cudaEvent_t ev_redo_done[2] = {nullptr, nullptr}, ev_fast_done = nullptr;
    bool redo_pending[2] = {false, false};
    int defer_buf = 0;
    /* trains of SIM5_FLAG_DEFER_REDO calls: the kernels of consecutive calls alternate between two internal launch streams; the caller's stream only
     * carries the entry / exit events */
    cudaStream_t train_stream[2] = {nullptr, nullptr}, hi_stream = nullptr;      /* low priority: tracing kernels of a train; high priority: its azimuth kernels */
    cudaEvent_t ev_entry[2] = {nullptr, nullptr}, ev_a_done[2] = {nullptr, nullptr}, ev_train_done[2] = {nullptr, nullptr};
    bool train_pending[2] = {false, false};
    /* SIM5_FLAG_STAGE_COPY: two alternating sets of local compact planes; the copy of set s to the caller's (peer) planes ends at ev_stage_done[s] */
    Plane stage[2][SIM5_NPLANES];
    cudaEvent_t ev_stage_done[2] = {nullptr, nullptr}, ev_stage_go = nullptr;
    bool stage_pending[2] = {false, false};
    int stage_buf = 0;
    unsigned long long* last_counts = nullptr;           /* queue counts of the most recent image call (sim5_last_phase_ms) */
    void* batch[8] = {nullptr};
    size_t batch_bytes[8] = {0};
    std::string last_error;
    std::mutex mu;
    /* disk_nt_setup() state of the scalar API */
    sim5_image_params disk_params;
    bool disk_set = false;
};

/* One context per CUDA device.  A host thread works on ONE of them at a time: the one it selected last (sim5_gpu_init(device), or
 * sim5_image_params.device >= 0), else the process default (the device of the first sim5_gpu_init, else device 0).  Selecting another
 * device never tears a context down: its stream, scratch planes and deferred work stay where they are.  sim5_trace_image_multi runs
 * one worker thread per device, each on its own context. */
#define S5_MAX_DEVICES 16
Context g_pool[S5_MAX_DEVICES];
int g_default_dev = 0;
bool g_default_set = false;
bool g_warned = false;
thread_local Context* t_ctx = nullptr;
inline Context& cur_ctx() { return t_ctx ? *t_ctx : g_pool[g_default_dev]; }
#define g_ctx (cur_ctx())
/* pick the context of `device` for this thread (device < 0: keep the current one) */
inline Context& select_ctx(int device)
{
    if (device >= 0 && device < S5_MAX_DEVICES) t_ctx = &g_pool[device];
    return cur_ctx();
}

void set_error(const std::string& s) { g_ctx.last_error = s; }

bool cuda_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}
#define CK(call) do { if (!cuda_ok((call), #call)) return SIM5_ERR_CUDA; } while (0)

int no_device(const char* why)
{
    set_error(std::string("no usable CUDA device (") + why + "); libsim5b200 has no CPU path");
    if (!g_warned) {
        fprintf(stderr, "sim5_b200: %s\n", g_ctx.last_error.c_str());
        g_warned = true;
    }
    return SIM5_ERR_NO_DEVICE;
}

/* Make the context of `device` (< 0: this thread's current one) ready and current for the CUDA runtime.  Called -- under the
 * context's mutex -- at the top of every entry point, so kernels, events and allocations always go to the context's own device
 * whatever the caller (or torch) made current in between. */
int ensure_init(int device)
{
    if (device >= S5_MAX_DEVICES) { set_error("device ordinal out of range"); return SIM5_ERR_BAD_PARAM; }
    Context& c = select_ctx(device);
    if (c.ready) { CK(cudaSetDevice(c.device)); return SIM5_OK; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return no_device(e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"); }
    device = (int)(&c - g_pool);
    if (device >= n) { set_error("device ordinal out of range"); return SIM5_ERR_BAD_PARAM; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return no_device(buf);
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    CK(cudaEventCreate(&c.ev0)); CK(cudaEventCreate(&c.ev1)); CK(cudaEventCreate(&c.ev2)); CK(cudaEventCreate(&c.ev3));
    for (int i = 0; i < 3; i++) CK(cudaEventCreate(&c.evp[i]));
    c.ring[0][0] = c.ev1; c.ring[0][1] = c.evp[0]; c.ring[0][2] = c.evp[1]; c.ring[0][3] = c.evp[2]; c.ring[0][4] = c.ev2;
    for (int k = 1; k < S5_RING; k++) for (int i = 0; i < 5; i++) CK(cudaEventCreate(&c.ring[k][i]));
    c.ring_pos = 0;
    for (int k = 0; k < S5_RING; k++) c.ring_phases[k] = 0;
    for (int i = 0; i < S5_MAX_CHUNKS; i++) CK(cudaEventCreateWithFlags(&c.ev_chunk[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_copy_done, cudaEventDisableTiming));
    CK(cudaHostAlloc((void**)&c.h_scr, sizeof(Scratch), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&c.d_scr, c.h_scr, 0));
    c.consts_cap = 1;
    CK(cudaHostAlloc((void**)&c.h_consts, sizeof(S5ImageConsts), cudaHostAllocDefault));
    CK(cudaMalloc((void**)&c.d_consts, sizeof(S5ImageConsts)));
    /* three blocks of {8 counters ([0..2] tile counters, [4..7] queue counts), DevStats}: the default one and the two alternating sets of
     * SIM5_FLAG_DEFER_REDO trains.  Counters and stats of a call are adjacent, so ONE memset per call clears both */
    static_assert(sizeof(DevStats) % sizeof(unsigned long long) == 0, "DevStats is made of 64-bit counters");
    CK(cudaMalloc((void**)&c.d_counter, 3 * S5_BLK_WORDS * sizeof(unsigned long long)));
    c.d_counter2 = c.d_counter + S5_BLK_WORDS;
    CK(cudaEventCreateWithFlags(&c.ev_redo_done[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_redo_done[1], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_fast_done, cudaEventDisableTiming));
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);       /* numerically lower = higher priority */
    CK(cudaStreamCreateWithPriority(&c.hi_stream, cudaStreamNonBlocking, prio_hi));
    for (int b = 0; b < 2; b++) {
        CK(cudaStreamCreateWithPriority(&c.train_stream[b], cudaStreamNonBlocking, prio_lo));
        CK(cudaEventCreateWithFlags(&c.ev_entry[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_a_done[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_train_done[b], cudaEventDisableTiming));
        c.train_pending[b] = false;
    }
    CK(cudaEventCreateWithFlags(&c.ev_stage_done[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_stage_done[1], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.ev_stage_go, cudaEventDisableTiming));
    c.stage_pending[0] = c.stage_pending[1] = false; c.stage_buf = 0;
    c.redo_pending[0] = c.redo_pending[1] = false; c.defer_buf = 0; c.last_counts = c.d_counter + 4;
    c.d_stats = (DevStats*)(c.d_counter + 8);
    CK(cudaHostAlloc((void**)&c.h_stats, sizeof(DevStats), cudaHostAllocDefault));
    /* the stepper keeps ~100 doubles of live state per thread and calls non-inlined Carlson routines */
    cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    if (S5_EQ_DYN_SMEM > 0) {
        CK(cudaFuncSetAttribute(s5::k_trace_eqplane<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_EQ_DYN_SMEM));
        CK(cudaFuncSetAttribute(s5::k_trace_eqplane<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_EQ_DYN_SMEM));
        CK(cudaFuncSetAttribute(s5::k_trace_eqplane<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_EQ_DYN_SMEM));
        CK(cudaFuncSetAttribute(s5::k_trace_eqplane<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_EQ_DYN_SMEM));
        CK(cudaFuncSetAttribute(s5::k_trace_histogram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_EQ_DYN_SMEM));
        if (S5_SPEC_DYN_SMEM > 0) CK(cudaFuncSetAttribute(s5::k_trace_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S5_SPEC_DYN_SMEM));
    }
    /* (cudaFuncAttributePreferredSharedMemoryCarveout = 0, i.e. all 256 KB to L1, was tried for the spilling kernels: phase A unchanged,
     * SURFACE 42.6 -> 45.4 ms -- profiles/r01z_sweep.log -- so the driver's default carve-out stays) */
    c.ready = true;
    return SIM5_OK;
}

int reserve(Plane& pl, size_t bytes)
{
    if (pl.bytes >= bytes && pl.p) return SIM5_OK;
    if (pl.p) { cudaFree(pl.p); pl.p = nullptr; pl.bytes = 0; }
    CK(cudaMalloc(&pl.p, bytes));
    pl.bytes = bytes;
    return SIM5_OK;
}

int reserve_consts(size_t n)
{
    Context& c = g_ctx;
    if (c.consts_cap >= n) return SIM5_OK;
    cudaFreeHost(c.h_consts); cudaFree(c.d_consts);
    CK(cudaHostAlloc((void**)&c.h_consts, n * sizeof(S5ImageConsts), cudaHostAllocDefault));
    CK(cudaMalloc((void**)&c.d_consts, n * sizeof(S5ImageConsts)));
    c.consts_cap = n;
    return SIM5_OK;
}

template <class K>
int persistent_grid(K kernel, int threads, size_t dyn_smem = 0)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return g_ctx.sm_count * per_sm;
}

const struct { unsigned bit; size_t elem; } kPlaneInfo[SIM5_NPLANES] = {
    {SIM5_OUT_R, 8}, {SIM5_OUT_PHI, 8}, {SIM5_OUT_G, 8}, {SIM5_OUT_FLUX, 8}, {SIM5_OUT_CHI, 8}, {SIM5_OUT_DELTA, 8},
    {SIM5_OUT_MUE, 8}, {SIM5_OUT_INTENSITY, 8}, {SIM5_OUT_TAU, 8}, {SIM5_OUT_QERR, 8}, {SIM5_OUT_STEPS, 4}, {SIM5_OUT_STATUS, 1},
    {SIM5_OUT_HEIGHT, 8}, {SIM5_OUT_DELAY, 8},
};
void* host_plane(const sim5_image_out* o, int i)
{
    switch (i) {
        case 0: return o->r; case 1: return o->phi; case 2: return o->g; case 3: return o->flux;
        case 4: return o->chi; case 5: return o->delta; case 6: return o->mue; case 7: return o->intensity;
        case 8: return o->tau; case 9: return o->qerr; case 10: return o->steps; case 11: return o->status;
        case 12: return o->height; case 13: return o->delay;
    }
    return nullptr;
}
void set_dev_plane(DevOut* d, int i, void* p)
{
    switch (i) {
        case 0: d->r = (double*)p; break; case 1: d->phi = (double*)p; break; case 2: d->g = (double*)p; break;
        case 3: d->flux = (double*)p; break; case 4: d->chi = (double*)p; break; case 5: d->delta = (double*)p; break;
        case 6: d->mue = (double*)p; break; case 7: d->intensity = (double*)p; break; case 8: d->tau = (double*)p; break;
        case 9: d->qerr = (double*)p; break; case 10: d->steps = (int*)p; break; case 11: d->status = (unsigned char*)p; break;
        case 12: d->height = (double*)p; break; case 13: d->delay = (double*)p; break;
    }
}

void lattice_params(const sim5_image_params* p, int img, sim5_image_params* q)
{
    int js = img / p->n_incl, ki = img % p->n_incl;
    *q = *p;
    q->mode = SIM5_MODE_EQPLANE;
    q->row_begin = 0; q->row_end = 0;
    q->outputs = SIM5_OUT_G | SIM5_OUT_FLUX;
    q->bh_spin = (p->n_spin > 1) ? p->spin_max * (double)js / (double)(p->n_spin - 1) : p->spin_max;
    if (q->bh_spin < 1e-4) q->bh_spin = 1e-4;
    double ideg = (p->n_incl > 1) ? p->incl_min_deg + (p->incl_max_deg - p->incl_min_deg) * (double)ki / (double)(p->n_incl - 1) : p->incl_min_deg;
    q->incl = ideg / 180.0 * M_PI;
    q->rmax = s5_host_r_ms(q->bh_spin) + p->rmax_offset;
    q->r_emit_min = 0.0;
}

int trace_histogram(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats)
{
    Context& c = g_ctx;
    if (!out->hist) { set_error("HISTOGRAM mode needs out->hist"); return SIM5_ERR_NO_OUTPUT; }
    if (p->n_spin < 1 || p->n_incl < 1 || p->n_bins < 1 || !(p->g_max > p->g_min)) { set_error("bad lattice"); return SIM5_ERR_BAD_PARAM; }
    int nimg = p->n_spin * p->n_incl;
    int lb = p->lattice_begin, le = p->lattice_end;
    if (lb == 0 && le == 0) le = nimg;
    if (lb < 0 || le > nimg || lb >= le) { set_error("bad lattice range"); return SIM5_ERR_BAD_PARAM; }
    /* interleaved deal of the slice over split_count GPUs (one process or worker thread per GPU): this call traces the images
     * lb + split_index, lb + split_index + split_count, ...; the bins of the other images of [lb,le) are zeroed, so the histograms of
     * all GPUs add up to the lattice (ncclReduce / dist.reduce between processes, k_sum_peers inside sim5_trace_image_multi) */
    const int split = p->split_count > 1 ? p->split_count : 1;
    if (split > 1 && (p->split_index < 0 || p->split_index >= split)) { set_error("bad split_index"); return SIM5_ERR_BAD_PARAM; }
    const int first = lb + (split > 1 ? p->split_index : 0);
    const int n_own = first < le ? (le - first + split - 1) / split : 0;
    int rc = reserve_consts((size_t)nimg);
    if (rc) return rc;
    for (int k = 0; k < n_own; k++) {
        int img = first + k * split;
        sim5_image_params q;
        lattice_params(p, img, &q);
        q.split_count = 0; q.split_index = 0;
        s5_fill_image_consts(&q, &c.h_consts[img]);
    }
    size_t hbytes = (size_t)nimg * p->n_bins * sizeof(double);
    bool devptr = (p->flags & SIM5_FLAG_DEVICE_PTRS) != 0;
    double* d_hist = out->hist;
    if (!devptr) { rc = reserve(c.hist, hbytes); if (rc) return rc; d_hist = (double*)c.hist.p; }
    CK(cudaEventRecord(c.ev0, c.stream));
    CK(cudaMemcpyAsync(c.d_consts + lb, c.h_consts + lb, (size_t)(le - lb) * sizeof(S5ImageConsts), cudaMemcpyHostToDevice, c.stream));
    CK(cudaMemsetAsync(d_hist + (size_t)lb * p->n_bins, 0, (size_t)(le - lb) * p->n_bins * sizeof(double), c.stream));
    CK(cudaMemsetAsync(c.d_counter, 0, sizeof(unsigned long long), c.stream));
    CK(cudaMemsetAsync(c.d_stats, 0, sizeof(DevStats), c.stream));
    int grid = persistent_grid(s5::k_trace_histogram, S5_EQ_THREADS, S5_EQ_DYN_SMEM);
    CK(cudaEventRecord(c.ev1, c.stream));
    if (n_own > 0) s5::k_trace_histogram<<<grid, S5_EQ_THREADS, S5_EQ_DYN_SMEM, c.stream>>>(c.d_consts, first, split, n_own, d_hist, c.d_counter, c.d_stats);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev2, c.stream));
    if (!devptr) CK(cudaMemcpyAsync(out->hist + (size_t)lb * p->n_bins, d_hist + (size_t)lb * p->n_bins, (size_t)(le - lb) * p->n_bins * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(c.h_stats, c.d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaEventRecord(c.ev3, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->rays = (int64_t)n_own * p->nx * p->ny;
        for (int i = 0; i < 32; i++) stats->class_count[i] = (int64_t)c.h_stats->cls[i];
        for (int i = 0; i < 8; i++) stats->gtype_count[i] = (int64_t)c.h_stats->gtype[i];
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev1, c.ev2); stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, c.ev0, c.ev3); stats->total_ms = ms;
        stats->kernel_launches = 1; stats->sm_count = c.sm_count; stats->grid_ctas = grid; stats->cta_threads = S5_EQ_THREADS;
    }
    return SIM5_OK;
}

/* mode SPECTRUM: rows [row_begin,row_end) (and the interleaved split) of one image summed into out->spectrum[n_energy] (host) */
int trace_spectrum(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats)
{
    Context& c = g_ctx;
    if (!out->spectrum) { set_error("SPECTRUM mode needs out->spectrum"); return SIM5_ERR_NO_OUTPUT; }
    if (p->n_energy < 1 || p->n_energy > S5_SPEC_MAX_E || !(p->e_min_kev > 0.0) || !(p->e_max_kev >= p->e_min_kev) || !(p->spec_hardf > 0.0)) {
        set_error("bad spectrum grid (1 <= n_energy <= 256, 0 < e_min_kev <= e_max_kev, spec_hardf > 0)"); return SIM5_ERR_BAD_PARAM;
    }
    int rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = p->ny;
    if (rb < 0 || re > p->ny || rb > re) { set_error("bad row range"); return SIM5_ERR_BAD_PARAM; }
    if (p->max_order < 0 || p->max_order > 2) { set_error("max_order must be 0..2"); return SIM5_ERR_BAD_PARAM; }
    int split = p->split_count > 1 ? p->split_count : 1;
    int srows = p->split_rows > 0 ? p->split_rows : 1;
    if (split > 1 && ((p->split_index < 0 || p->split_index >= split) || (re - rb) % (split * srows) != 0)) { set_error("bad split"); return SIM5_ERR_BAD_PARAM; }
    sim5_image_params q = *p;
    q.mode = SIM5_MODE_POLARIZED;                      /* the pixel pipeline the spectrum sits on: g and mu_e from the emitter frame */
    q.outputs = SIM5_OUT_R | SIM5_OUT_G | SIM5_OUT_MUE;
    S5ImageConsts consts;
    s5_fill_image_consts(&q, &consts);
    const int ne = p->n_energy;
    int rc = reserve(c.hist, 2 * (size_t)S5_SPEC_MAX_E * sizeof(double));
    if (rc) return rc;
    double* d_e = (double*)c.hist.p;
    double* d_spec = d_e + S5_SPEC_MAX_E;
    double h_e[S5_SPEC_MAX_E];
    for (int k = 0; k < ne; k++) h_e[k] = s5_spectrum_energy(p, k);
    CK(cudaEventRecord(c.ev0, c.stream));
    CK(cudaMemcpyAsync(d_e, h_e, ne * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    CK(cudaMemsetAsync(d_spec, 0, S5_SPEC_MAX_E * sizeof(double), c.stream));
    CK(cudaMemsetAsync(c.d_counter, 0, sizeof(unsigned long long), c.stream));
    CK(cudaMemsetAsync(c.d_stats, 0, sizeof(DevStats), c.stream));
    int grid = persistent_grid(s5::k_trace_spectrum, S5_SPEC_THREADS, S5_SPEC_DYN_SMEM);
    CK(cudaEventRecord(c.ev1, c.stream));
    if (consts.nrows_local > 0) s5::k_trace_spectrum<<<grid, S5_SPEC_THREADS, S5_SPEC_DYN_SMEM, c.stream>>>(consts, d_e, d_spec, c.d_counter, c.d_stats);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev2, c.stream));
    CK(cudaMemcpyAsync(out->spectrum, d_spec, ne * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaMemcpyAsync(c.h_stats, c.d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaEventRecord(c.ev3, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    c.phases = 0;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->rays = (int64_t)consts.nrows_local * p->nx;
        for (int i = 0; i < 32; i++) stats->class_count[i] = (int64_t)c.h_stats->cls[i];
        for (int i = 0; i < 8; i++) stats->gtype_count[i] = (int64_t)c.h_stats->gtype[i];
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev1, c.ev2); stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, c.ev0, c.ev3); stats->total_ms = ms;
        stats->kernel_launches = 1; stats->sm_count = c.sm_count; stats->grid_ctas = grid; stats->cta_threads = S5_SPEC_THREADS;
    }
    return SIM5_OK;
}

} /* anonymous namespace */

/* ------------------------------------------------------------------ */
/* lifecycle                                                           */
/* ------------------------------------------------------------------ */
extern "C" int sim5_gpu_init(int device)
{
    if (device >= S5_MAX_DEVICES) { set_error("device ordinal out of range"); return SIM5_ERR_BAD_PARAM; }
    Context& c = select_ctx(device);
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = ensure_init(device);
    if (rc == SIM5_OK && !g_default_set) { g_default_dev = c.device; g_default_set = true; }   /* threads that never select use this one */
    return rc;
}

namespace {
void shutdown_ctx(Context& c)
{
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    for (auto& pl : c.planes) { if (pl.p) cudaFree(pl.p); pl = Plane(); }
    if (c.hist.p) cudaFree(c.hist.p); c.hist = Plane();
    if (c.azq_f.p) cudaFree(c.azq_f.p); c.azq_f = Plane();
    if (c.azq_key.p) cudaFree(c.azq_key.p); c.azq_key = Plane();
    if (c.azq_redo.p) cudaFree(c.azq_redo.p); c.azq_redo = Plane();
    cudaStreamSynchronize(c.aux_stream);
    if (c.azq2_f.p) cudaFree(c.azq2_f.p); c.azq2_f = Plane();
    if (c.azq2_key.p) cudaFree(c.azq2_key.p); c.azq2_key = Plane();
    if (c.azq2_redo.p) cudaFree(c.azq2_redo.p); c.azq2_redo = Plane();
    if (c.shared_q.p) cudaFree(c.shared_q.p); c.shared_q = Plane();
    c.d_counter2 = nullptr;
    cudaStreamSynchronize(c.hi_stream); cudaStreamDestroy(c.hi_stream); c.hi_stream = nullptr;
    for (int b = 0; b < 2; b++) {
        cudaStreamSynchronize(c.train_stream[b]); cudaStreamDestroy(c.train_stream[b]); c.train_stream[b] = nullptr;
        cudaEventDestroy(c.ev_entry[b]); cudaEventDestroy(c.ev_a_done[b]); cudaEventDestroy(c.ev_train_done[b]);
        c.train_pending[b] = false;
    }
    cudaStreamSynchronize(c.copy_stream);
    for (int b = 0; b < 2; b++) for (auto& pl : c.stage[b]) { if (pl.p) cudaFree(pl.p); pl = Plane(); }
    cudaEventDestroy(c.ev_stage_done[0]); cudaEventDestroy(c.ev_stage_done[1]); cudaEventDestroy(c.ev_stage_go);
    c.stage_pending[0] = c.stage_pending[1] = false;
    cudaEventDestroy(c.ev_redo_done[0]); cudaEventDestroy(c.ev_redo_done[1]); cudaEventDestroy(c.ev_fast_done);
    for (int i = 0; i < 8; i++) { if (c.batch[i]) cudaFree(c.batch[i]); c.batch[i] = nullptr; c.batch_bytes[i] = 0; }
    cudaFreeHost(c.h_scr); cudaFreeHost(c.h_consts); cudaFree(c.d_consts); cudaFree(c.d_counter); c.d_counter = nullptr; c.d_stats = nullptr; cudaFreeHost(c.h_stats);
    cudaEventDestroy(c.ev0); cudaEventDestroy(c.ev3);
    for (int k = 0; k < S5_RING; k++) for (int i = 0; i < 5; i++) cudaEventDestroy(c.ring[k][i]);
    for (int i = 0; i < S5_MAX_CHUNKS; i++) cudaEventDestroy(c.ev_chunk[i]);
    cudaEventDestroy(c.ev_copy_done);
    cudaStreamDestroy(c.own_stream); cudaStreamDestroy(c.copy_stream); cudaStreamDestroy(c.aux_stream);
    cudaEventDestroy(c.ev_fork); cudaEventDestroy(c.ev_join);
    c.stream = c.own_stream = c.copy_stream = c.aux_stream = nullptr;
    c.redo_pending[0] = c.redo_pending[1] = false;
    c.phases = 0;
    c.ready = false;
}
}

/* tears down the contexts of ALL devices */
extern "C" void sim5_gpu_shutdown(void)
{
    for (int d = 0; d < S5_MAX_DEVICES; d++) {
        std::lock_guard<std::mutex> lk(g_pool[d].mu);
        shutdown_ctx(g_pool[d]);
    }
    g_default_set = false; g_default_dev = 0;
    t_ctx = nullptr;
}

extern "C" int sim5_set_stream(void* cuda_stream)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    int rc = ensure_init(-1);
    if (rc) return rc;
    g_ctx.stream = cuda_stream ? (cudaStream_t)cuda_stream : g_ctx.own_stream;
    return SIM5_OK;
}
extern "C" int sim5_set_chunk_rays(int64_t rays)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    g_ctx.chunk_rays = rays > 0 ? (long long)rays : 0;
    return SIM5_OK;
}

namespace {
/* the launch stream waits for the redo passes that SIM5_FLAG_DEFER_REDO calls left running on the auxiliary stream */
int join_deferred(Context& c, int only_buf /* -1: all */, cudaStream_t on = nullptr)
{
    if (!on) on = c.stream;
    for (int b = 0; b < 2; b++) {
        if (!c.redo_pending[b] || (only_buf >= 0 && b != only_buf)) continue;
        CK(cudaStreamWaitEvent(on, c.ev_redo_done[b], 0));
        c.redo_pending[b] = false;
    }
    if (only_buf < 0) {
        for (int b = 0; b < 2; b++) {                  /* ... for the internal launch streams of a train */
            if (!c.train_pending[b]) continue;
            CK(cudaStreamWaitEvent(on, c.ev_train_done[b], 0));
            c.train_pending[b] = false;
        }
        for (int b = 0; b < 2; b++) {                  /* ... and for the staged plane copies of SIM5_FLAG_STAGE_COPY calls */
            if (!c.stage_pending[b]) continue;
            CK(cudaStreamWaitEvent(on, c.ev_stage_done[b], 0));
            c.stage_pending[b] = false;
        }
    }
    return SIM5_OK;
}
}

extern "C" int sim5_join(void)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (!g_ctx.ready) return SIM5_OK;
    return join_deferred(g_ctx, -1);
}

extern "C" int sim5_synchronize(void)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (!g_ctx.ready) return SIM5_OK;
    int rc = join_deferred(g_ctx, -1);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_ctx.stream));
    return SIM5_OK;
}

extern "C" int sim5_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* sim5_last_error(void) { return g_ctx.last_error.c_str(); }
extern "C" const char* sim5_version(void) { return "sim5_b200 0.1 (sm_100a, fp64)"; }

extern "C" void* sim5_host_alloc(size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return nullptr;
    void* p = nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
    return p;
}
extern "C" void sim5_host_free(void* p) { if (p) cudaFreeHost(p); }
/* page-lock caller-owned host memory (e.g. a shared-memory segment that several one-process-per-GPU ranks map: every rank registers it and
 * copies its own rows into the one image), so that the device->host copies of a host-plane call run asynchronously under its kernels */
extern "C" int sim5_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) { set_error("sim5_host_register: null argument"); return SIM5_ERR_BAD_PARAM; }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    int rc = ensure_init(-1);
    if (rc) return rc;
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return SIM5_OK;
}
extern "C" int sim5_host_unregister(void* p)
{
    if (!p) return SIM5_OK;
    CK(cudaHostUnregister(p));
    return SIM5_OK;
}
extern "C" void* sim5_device_alloc(size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return nullptr;
    void* p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc")) return nullptr;
    return p;
}
extern "C" void sim5_device_free(void* p) { if (p) cudaFree(p); }
/* CUDA IPC: lets the ranks of a one-process-per-GPU job store their rows straight into the image planes of the rank that
 * assembles the image (peer memory over NVLink / NVSwitch), instead of gathering compact planes afterwards */
extern "C" int sim5_ipc_export(const void* device_ptr, void* handle64)
{
    if (!device_ptr || !handle64) { set_error("sim5_ipc_export: null argument"); return SIM5_ERR_BAD_PARAM; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    memcpy(handle64, &h, 64);
    return SIM5_OK;
}
extern "C" void* sim5_ipc_import(const void* handle64)
{
    if (!handle64) { set_error("sim5_ipc_import: null handle"); return nullptr; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
    return p;
}
extern "C" int sim5_ipc_release(void* imported_ptr)
{
    if (!imported_ptr) return SIM5_OK;
    CK(cudaIpcCloseMemHandle(imported_ptr));
    return SIM5_OK;
}

/* Ordered behind everything the library has enqueued for this context: the launch stream (which is cudaStreamNonBlocking, so a plain
 * cudaMemcpy on the legacy stream would NOT wait for it) and the deferred redo passes of SIM5_FLAG_DEFER_REDO calls. */
extern "C" int sim5_device_memset(void* p, int value, size_t bytes)
{
    if (!p) { set_error("sim5_device_memset: null pointer"); return SIM5_ERR_BAD_PARAM; }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return SIM5_ERR_NO_DEVICE;
    CK(cudaMemsetAsync(p, value, bytes, g_ctx.stream));
    CK(cudaStreamSynchronize(g_ctx.stream));
    return SIM5_OK;
}
extern "C" int sim5_host_to_device(void* dst, const void* src, size_t bytes)
{
    if (!dst || !src) { set_error("sim5_host_to_device: null pointer"); return SIM5_ERR_BAD_PARAM; }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return SIM5_ERR_NO_DEVICE;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    CK(cudaStreamSynchronize(g_ctx.stream));
    return SIM5_OK;
}
extern "C" int sim5_device_to_host(void* dst, const void* src, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return SIM5_ERR_NO_DEVICE;
    int rc = join_deferred(g_ctx, -1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
    CK(cudaStreamSynchronize(g_ctx.stream));
    return SIM5_OK;
}

/* the open parameters of the five BASELINE configs */
extern "C" int sim5_default_params(int cfg, sim5_image_params* p)
{
    if (!p || cfg < 1 || cfg > 7) return SIM5_ERR_BAD_PARAM;
    memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(*p);
    p->device = -1;                /* the calling thread's current context (sim5_gpu_init), never a silent move to device 0 */
    p->max_order = 1;
    p->disk_mass = 10.0; p->disk_mdot = 0.1; p->disk_alpha = 0.1;
    p->precision_factor = 0.01; p->r_start = 50.0; p->step_max = 1e9; p->max_steps = 100000;
    p->torus_rc = 10.0; p->torus_w = 2.0; p->torus_h = 0.3; p->torus_j0 = 1.0; p->torus_k0 = 0.05;
    p->n_spin = 64; p->n_incl = 32; p->n_bins = 256;
    p->spin_max = 0.998; p->incl_min_deg = 5.0; p->incl_max_deg = 85.0;
    p->g_min = 0.0; p->g_max = 2.0; p->rmax_offset = 20.0;
    p->n_energy = 128; p->spec_limb = 1; p->e_min_kev = 0.05; p->e_max_kev = 50.0; p->spec_hardf = 1.7;
    p->surf_hr = 0.2; p->surf_rin = 0.0; p->delay_r_ref = 1000.0;
    switch (cfg) {
        case 1: p->mode = SIM5_MODE_EQPLANE; p->nx = p->ny = 512; p->bh_spin = 0.9; p->incl = 70.0 / 180.0 * M_PI;
                p->rmax = s5_host_r_ms(p->bh_spin) + 8.0; p->outputs = SIM5_OUT_R | SIM5_OUT_G | SIM5_OUT_FLUX | SIM5_OUT_STATUS; break;
        case 2: p->mode = SIM5_MODE_EQPLANE; p->nx = p->ny = 4096; p->bh_spin = 0.998; p->incl = 75.0 / 180.0 * M_PI;
                p->rmax = s5_host_r_ms(p->bh_spin) + 20.0; p->outputs = SIM5_OUT_R | SIM5_OUT_PHI | SIM5_OUT_G | SIM5_OUT_FLUX | SIM5_OUT_STATUS; break;
        case 3: p->mode = SIM5_MODE_POLARIZED; p->nx = p->ny = 2048; p->bh_spin = 0.94; p->incl = 75.0 / 180.0 * M_PI;
                p->rmax = s5_host_r_ms(p->bh_spin) + 20.0;
                p->outputs = SIM5_OUT_R | SIM5_OUT_PHI | SIM5_OUT_G | SIM5_OUT_FLUX | SIM5_OUT_CHI | SIM5_OUT_DELTA | SIM5_OUT_STATUS; break;
        case 4: p->mode = SIM5_MODE_STEPWISE; p->nx = p->ny = 1024; p->bh_spin = 0.9; p->incl = 60.0 / 180.0 * M_PI;
                p->rmax = 25.0; p->outputs = SIM5_OUT_INTENSITY | SIM5_OUT_TAU | SIM5_OUT_STEPS | SIM5_OUT_STATUS; break;
        case 5: p->mode = SIM5_MODE_HISTOGRAM; p->nx = p->ny = 1024; p->bh_spin = 0.998; p->incl = 75.0 / 180.0 * M_PI;
                p->rmax = 0.0; p->outputs = 0; break;
        case 6: p->mode = SIM5_MODE_SPECTRUM; p->nx = p->ny = 2048; p->bh_spin = 0.998; p->incl = 75.0 / 180.0 * M_PI;
                p->rmax = s5_host_r_ms(p->bh_spin) + 20.0; p->outputs = 0; break;
        case 7: p->mode = SIM5_MODE_SURFACE; p->nx = p->ny = 1024; p->bh_spin = 0.9; p->incl = 60.0 / 180.0 * M_PI;
                p->rmax = 30.0; p->outputs = SIM5_OUT_R | SIM5_OUT_HEIGHT | SIM5_OUT_G | SIM5_OUT_MUE | SIM5_OUT_FLUX | SIM5_OUT_STEPS | SIM5_OUT_STATUS; break;
    }
    {   /* ellK(torus_rc, a), sim5kerr.c:1050-1071 */
        double r = p->torus_rc, a = p->bh_spin;
        p->torus_ell = (r * r - 2. * a * sqrt(r) + a * a) / (sqrt(r) * r - 2. * sqrt(r) + a);
    }
    return SIM5_OK;
}

/* ------------------------------------------------------------------ */
/* the batched entry                                                   */
/* ------------------------------------------------------------------ */
extern "C" int sim5_trace_image(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats)
{
    if (!p || !out) { set_error("null params/out"); return SIM5_ERR_BAD_PARAM; }
    if (p->struct_size != (int32_t)sizeof(sim5_image_params)) { set_error("sim5_image_params.struct_size mismatch"); return SIM5_ERR_BAD_PARAM; }
    if (p->nx <= 0 || p->ny <= 0) { set_error("empty image"); return SIM5_ERR_BAD_PARAM; }
    if (p->mode < SIM5_MODE_EQPLANE || p->mode > SIM5_MODE_SURFACE) { set_error("unknown mode"); return SIM5_ERR_BAD_PARAM; }
    if (p->device >= S5_MAX_DEVICES) { set_error("device ordinal out of range"); return SIM5_ERR_BAD_PARAM; }
    Context& c = select_ctx(p->device);              /* device < 0: this thread's current context */
    std::lock_guard<std::mutex> lk(c.mu);
    int rc = ensure_init(p->device);
    if (rc) return rc;
    if (p->mode == SIM5_MODE_HISTOGRAM) return trace_histogram(p, out, stats);
    if (p->mode == SIM5_MODE_SPECTRUM) return trace_spectrum(p, out, stats);

    int rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = p->ny;
    if (rb < 0 || re > p->ny || rb > re) { set_error("bad row range"); return SIM5_ERR_BAD_PARAM; }
    if (p->max_order < 0 || p->max_order > 2) { set_error("max_order must be 0..2"); return SIM5_ERR_BAD_PARAM; }
    if (p->mode == SIM5_MODE_STEPWISE && (p->max_steps < 1 || !(p->precision_factor > 0))) { set_error("bad stepper parameters"); return SIM5_ERR_BAD_PARAM; }
    for (int i = 0; i < SIM5_NPLANES; i++)
        if ((p->outputs & kPlaneInfo[i].bit) && !host_plane(out, i)) { set_error("selected output plane is NULL"); return SIM5_ERR_NO_OUTPUT; }

    int split = p->split_count > 1 ? p->split_count : 1;
    int srows = (split > 1 && p->split_rows > 0) ? p->split_rows : 1;     /* split_rows means nothing without a split: chunks may then cut at any row */
    if (split > 1) {
        if (p->split_index < 0 || p->split_index >= split) { set_error("bad split_index"); return SIM5_ERR_BAD_PARAM; }
        if ((re - rb) % (split * srows) != 0) { set_error("row range must be a multiple of split_count*split_rows"); return SIM5_ERR_BAD_PARAM; }
    }
    int nrows_local = (re - rb) / split;
    size_t npix = (size_t)nrows_local * (size_t)p->nx;
    bool devptr = (p->flags & SIM5_FLAG_DEVICE_PTRS) != 0;
    bool async = devptr && (p->flags & SIM5_FLAG_ASYNC);
    DevOut d;
    memset(&d, 0, sizeof d);
    d.compact = (!devptr || (split > 1 && !(p->flags & SIM5_FLAG_FULL_INDEX))) ? 1 : 0;
    const bool stage_copy = devptr && async && split > 1 && npix > 0 && (p->flags & SIM5_FLAG_FULL_INDEX) && (p->flags & SIM5_FLAG_STAGE_COPY);
    int sb = 0;
    if (stage_copy) {
        sb = c.stage_buf;
        c.stage_buf ^= 1;
        d.compact = 1;
    }
    for (int i = 0; i < SIM5_NPLANES; i++) {
        if (!(p->outputs & kPlaneInfo[i].bit)) continue;
        if (stage_copy) {
            rc = reserve(c.stage[sb][i], npix * kPlaneInfo[i].elem);
            if (rc) return rc;
            set_dev_plane(&d, i, c.stage[sb][i].p);
            continue;
        }
        if (devptr) { set_dev_plane(&d, i, host_plane(out, i)); continue; }
        rc = reserve(c.planes[i], (npix ? npix : 1) * kPlaneInfo[i].elem);
        if (rc) return rc;
        set_dev_plane(&d, i, c.planes[i].p);
    }

    S5ImageConsts consts;                  /* travels by value as a kernel parameter: no H2D copy, async-safe */
    s5_fill_image_consts(p, &consts);
    /* two-phase azimuth: phase A queues the disk hits, phase B integrates phi per geodesic type */
    bool lanes = (p->mode == SIM5_MODE_STEPWISE || p->mode == SIM5_MODE_SURFACE);
    const bool shared_queue = (p->flags & SIM5_FLAG_SHARED_QUEUE) != 0;
    if (shared_queue && !(lanes && devptr && split == 1 && out->shared_counter)) {
        set_error("SIM5_FLAG_SHARED_QUEUE needs a lane mode (STEPWISE / SURFACE), DEVICE_PTRS, no row split and out->shared_counter");
        return SIM5_ERR_BAD_PARAM;
    }
    bool two_phase = !lanes && (p->outputs & SIM5_OUT_PHI) && !(p->flags & SIM5_FLAG_SINGLE_PASS) && npix > 0;

    /* Host planes: the rows are traced in CHUNKS so the device->host copy of chunk k (copy stream, copy engine) runs
     * under the kernels of chunk k+1; only the last chunk's copy is exposed.  Device planes: one chunk. */
    int nchunks = 1;
    if (!devptr && !(p->flags & SIM5_FLAG_NO_OVERLAP) && npix > 0) {
        long long want = (long long)(npix / (size_t)(c.chunk_rays > 0 ? c.chunk_rays : S5_CHUNK_RAYS));
        nchunks = (int)(want < 1 ? 1 : (want > S5_MAX_CHUNKS ? S5_MAX_CHUNKS : want));
    }
    int chunk_rows = (nrows_local + nchunks - 1) / nchunks;
    chunk_rows = ((chunk_rows + srows - 1) / srows) * srows;
    if (chunk_rows < srows) chunk_rows = srows;
    nchunks = nrows_local > 0 ? (nrows_local + chunk_rows - 1) / chunk_rows : 1;
    /* chunk boundaries in local rows.  The first chunk's kernels and the last chunk's copy are the only exposed pieces of the
     * pipeline, so with >= 4 chunks the first and the last are cut in half (one more chunk in total) */
    int chunk_lr[S5_MAX_CHUNKS + 2];
    {
        int n = 0, lr = 0;
        bool ramp = nchunks >= 4 && nchunks + 1 <= S5_MAX_CHUNKS && (chunk_rows / 2 / srows) * srows >= srows;
        int half = ramp ? (chunk_rows / 2 / srows) * srows : chunk_rows;
        chunk_lr[n++] = 0;
        if (ramp) { lr = half; chunk_lr[n++] = lr; }
        while (lr < nrows_local && n <= S5_MAX_CHUNKS) {
            int left = nrows_local - lr;
            int take = (ramp && left <= chunk_rows + half && left > half) ? left - half : (left < chunk_rows ? left : chunk_rows);
            if (n == S5_MAX_CHUNKS) take = left;
            lr += take; chunk_lr[n++] = lr;
        }
        nchunks = nrows_local > 0 ? n - 1 : 1;
        if (nrows_local == 0) chunk_lr[1] = 0;
    }

    AzQueue q;
    memset(&q, 0, sizeof q);
    if (two_phase) {
        int max_rows = 0;
        for (int ch = 0; ch < nchunks; ch++) if (chunk_lr[ch + 1] - chunk_lr[ch] > max_rows) max_rows = chunk_lr[ch + 1] - chunk_lr[ch];
        size_t qpix = (size_t)max_rows * (size_t)p->nx;
        rc = reserve(c.azq_f, qpix * S5_AZ_NFIELDS * sizeof(double)); if (rc) return rc;
        rc = reserve(c.azq_key, qpix * sizeof(unsigned long long)); if (rc) return rc;
        rc = reserve(c.azq_redo, qpix * sizeof(unsigned)); if (rc) return rc;
        q.f = (double*)c.azq_f.p; q.key = (unsigned long long*)c.azq_key.p; q.count = c.d_counter + 4; q.cap = (long long)qpix;
        q.redo = (unsigned*)c.azq_redo.p;
    }
    /* deferred redo passes: this call's queue and counters alternate between two sets; the set it takes must be free again (the call
     * before the previous one), every other kind of call first waits for all of them */
    const bool defer = two_phase && async && nchunks == 1 && (p->flags & SIM5_FLAG_DEFER_REDO) && !(p->flags & SIM5_FLAG_EXACT_AZIMUTH);
    unsigned long long* cnt = c.d_counter;
    cudaStream_t ls = c.stream;                  /* the stream this call's kernels are launched on */
    const bool alt = defer && (p->flags & SIM5_FLAG_ALT_STREAMS);      /* A/B only: co-running the two kernels slows both (sim5_b200.h) */
    if (defer) {
        const int b = c.defer_buf;
        c.defer_buf ^= 1;
        if (alt) {
            /* behind everything the caller has enqueued so far, and behind the tracing kernel of the previous call of the train */
            ls = c.train_stream[b];
            CK(cudaEventRecord(c.ev_entry[b], c.stream));
            CK(cudaStreamWaitEvent(ls, c.ev_entry[b], 0));
            CK(cudaStreamWaitEvent(ls, c.ev_a_done[b ^ 1], 0));
        }
        rc = join_deferred(c, b, ls); if (rc) return rc;
        cnt = c.d_counter2 + S5_BLK_WORDS * b;
        if (b == 1) {
            size_t qpix = (size_t)q.cap;
            rc = reserve(c.azq2_f, qpix * S5_AZ_NFIELDS * sizeof(double)); if (rc) return rc;
            rc = reserve(c.azq2_key, qpix * sizeof(unsigned long long)); if (rc) return rc;
            rc = reserve(c.azq2_redo, qpix * sizeof(unsigned)); if (rc) return rc;
            q.f = (double*)c.azq2_f.p; q.key = (unsigned long long*)c.azq2_key.p; q.redo = (unsigned*)c.azq2_redo.p;
        }
        q.count = cnt + 4;
    } else {
        rc = join_deferred(c, -1); if (rc) return rc;
    }
    c.last_counts = cnt + 4;
    c.ring_pos = (c.ring_pos + 1) % S5_RING;
    c.ev1 = c.ring[c.ring_pos][0]; c.evp[0] = c.ring[c.ring_pos][1]; c.evp[1] = c.ring[c.ring_pos][2]; c.evp[2] = c.ring[c.ring_pos][3];
    c.ev2 = c.ring[c.ring_pos][4];
    c.ring_phases[c.ring_pos] = 0;
    DevStats* const d_stats = (DevStats*)(cnt + 8);          /* this call's stats sit behind its counters */
    int grid = 0, launches = 0;
    if (stage_copy && c.stage_pending[sb]) { CK(cudaStreamWaitEvent(ls, c.ev_stage_done[sb], 0)); c.stage_pending[sb] = false; }      /* the scratch set is free again */
    if (npix == 0) CK(cudaMemsetAsync(cnt, 0, S5_BLK_WORDS * sizeof(unsigned long long), ls));
    CK(cudaEventRecord(c.ev1, ls));                          /* start of the call (one memset of 400 bytes precedes the first kernel) */
    for (int ch = 0; ch < nchunks && npix > 0; ch++) {
        int lr0 = chunk_lr[ch];
        int lrows = chunk_lr[ch + 1] - lr0;
        size_t pix0 = (size_t)lr0 * (size_t)p->nx;
        S5ImageConsts cc = consts;
        DevOut dd = d;
        if (nchunks > 1) {
            /* local rows [lr0, lr0+lrows) of this call == a call of its own whose row range starts (lr0/srows) split periods later */
            cc.row_begin = consts.row_begin + (lr0 / srows) * split * srows;
            cc.nrows_local = lrows;
            cc.row_end = cc.row_begin + lrows * split;
            for (int i = 0; i < SIM5_NPLANES; i++) {
                if (!(p->outputs & kPlaneInfo[i].bit)) continue;
                set_dev_plane(&dd, i, (char*)c.planes[i].p + pix0 * kPlaneInfo[i].elem);
            }
        }
        CK(cudaMemsetAsync(cnt, 0, (ch == 0 ? S5_BLK_WORDS : 8) * sizeof(unsigned long long), ls));      /* counters per chunk, stats once */
        /* the ray queue: this call's own counter, or the caller's shared word (several GPUs pull rays of one image from it) */
        unsigned long long* rayq = shared_queue ? (unsigned long long*)out->shared_counter : cnt;
        if (p->mode == SIM5_MODE_STEPWISE) {
            grid = persistent_grid(s5::k_trace_lanes<s5::StepwiseProg>, s5::StepwiseProg::THREADS);
            s5::k_trace_lanes<s5::StepwiseProg><<<grid, s5::StepwiseProg::THREADS, 0, ls>>>(cc, dd, rayq, d_stats);
        } else if (p->mode == SIM5_MODE_SURFACE) {
            grid = persistent_grid(s5::k_trace_lanes<s5::SurfaceProg>, s5::SurfaceProg::THREADS);
            s5::k_trace_lanes<s5::SurfaceProg><<<grid, s5::SurfaceProg::THREADS, 0, ls>>>(cc, dd, rayq, d_stats);
        } else if (two_phase) {
            if (p->outputs & SIM5_OUT_DELAY) {
                grid = persistent_grid(s5::k_trace_eqplane<true, true>, S5_EQ_THREADS, S5_EQ_DYN_SMEM);
                s5::k_trace_eqplane<true, true><<<grid, S5_EQ_THREADS, S5_EQ_DYN_SMEM, ls>>>(cc, dd, q, cnt, d_stats);
            } else {
                grid = persistent_grid(s5::k_trace_eqplane<true>, S5_EQ_THREADS, S5_EQ_DYN_SMEM);
                s5::k_trace_eqplane<true><<<grid, S5_EQ_THREADS, S5_EQ_DYN_SMEM, ls>>>(cc, dd, q, cnt, d_stats);
            }
            if (ch == 0) CK(cudaEventRecord(c.evp[0], ls));
            if (alt) CK(cudaEventRecord(c.ev_a_done[(cnt == c.d_counter2) ? 0 : 1], ls));      /* the next call of the train may start its tracing kernel */
            int g_rr = persistent_grid(s5::k_azimuth<s5::GEOD_TYPE_RR>, S5_AZ_THREADS);
            int g_rc = persistent_grid(s5::k_azimuth<s5::GEOD_TYPE_RC>, S5_AZ_THREADS);
            /* the redo passes are one latency-bound wave (~0.1 ms whatever the number of items).  Tried and dropped (profiles/r02y_redo_sweep.log):
             * smaller CTAs (64 / 128 threads: more serial passes, 0.31 / 0.21 ms on an eighth of the image against 0.24) and an even deal of the
             * items over all CTAs without barriers (0.14 / 0.22 ms against 0.10 / 0.24): the wave is bound by walking ~120 KB of code, not by the pipe */
            const int redo_threads = S5_AZ_THREADS;
            if (p->flags & SIM5_FLAG_EXACT_AZIMUTH) {
                s5::k_azimuth<s5::GEOD_TYPE_RR><<<g_rr, S5_AZ_THREADS, 0, ls>>>(cc, q, dd.phi, cnt + 1, 0);
                if (ch == 0) CK(cudaEventRecord(c.evp[1], ls));
                s5::k_azimuth<s5::GEOD_TYPE_RC><<<g_rc, S5_AZ_THREADS, 0, ls>>>(cc, q, dd.phi, cnt + 2, 0);
            } else {
                if (defer) {
                    /* a train of images: RR hits on the launch stream, the RC chain (3 % of the hits; 0.2 ms alone on a full image) beside it on
                     * the auxiliary stream, where both redo passes stay too, in a few CTAs (a 512-thread CTA of the bit-faithful kernel fills the
                     * register file of its SM, and the next call's tracing kernel is about to want the SMs); nothing joins the launch stream here.
                     * (One launch for both queues, k_azimuth_fast<0>, saves a stream operation but serialises the RC part: 2 GPUs 1.35 -> 1.53 ms,
                     * profiles/r05h_bench_cfg2_n2.json.) */
                    CK(cudaStreamWaitEvent(c.aux_stream, c.evp[0], 0));              /* evp[0]: phase A done */
                    if (alt) { CK(cudaStreamWaitEvent(c.hi_stream, c.evp[0], 0)); ls = c.hi_stream; }      /* A/B: the azimuth on the high-priority stream */
                    int g_f = persistent_grid(s5::k_azimuth_fast<1>, S5_AZF_THREADS);
                    s5::k_azimuth_fast<1><<<g_f, S5_AZF_THREADS, 0, ls>>>(cc, q, dd.phi);
                    g_f = persistent_grid(s5::k_azimuth_fast<2>, S5_AZF_THREADS);
                    s5::k_azimuth_fast<2><<<g_f, S5_AZF_THREADS, 0, c.aux_stream>>>(cc, q, dd.phi);
                    const int gs_rc = g_rc < S5_DEFER_REDO_CTAS ? g_rc : S5_DEFER_REDO_CTAS, gs_rr = g_rr < S5_DEFER_REDO_CTAS ? g_rr : S5_DEFER_REDO_CTAS;
                    CK(cudaEventRecord(c.evp[1], ls));             /* end of the call on the launch stream; the RR redo pass waits for it */
                    s5::k_azimuth<s5::GEOD_TYPE_RC><<<gs_rc, redo_threads, 0, c.aux_stream>>>(cc, q, dd.phi, cnt + 2, 1);
                    CK(cudaStreamWaitEvent(c.aux_stream, c.evp[1], 0));
                    s5::k_azimuth<s5::GEOD_TYPE_RR><<<gs_rr, redo_threads, 0, c.aux_stream>>>(cc, q, dd.phi, cnt + 1, 1);
                    const int b = (cnt == c.d_counter2) ? 0 : 1;
                    CK(cudaEventRecord(c.ev_redo_done[b], c.aux_stream));
                    c.redo_pending[b] = true;
                } else {
                /* RR chain on the launch stream, RC chain (3 % of the hits) on the auxiliary stream: the two bit-faithful redo
                 * passes are latency-bound single waves, so they run side by side instead of back to back */
                CK(cudaEventRecord(c.ev_fork, ls));
                CK(cudaStreamWaitEvent(c.aux_stream, c.ev_fork, 0));
                int g_f = persistent_grid(s5::k_azimuth_fast<1>, S5_AZF_THREADS);
                s5::k_azimuth_fast<1><<<g_f, S5_AZF_THREADS, 0, ls>>>(cc, q, dd.phi);
                g_f = persistent_grid(s5::k_azimuth_fast<2>, S5_AZF_THREADS);
                s5::k_azimuth_fast<2><<<g_f, S5_AZF_THREADS, 0, c.aux_stream>>>(cc, q, dd.phi);
                {
                    s5::k_azimuth<s5::GEOD_TYPE_RC><<<g_rc, redo_threads, 0, c.aux_stream>>>(cc, q, dd.phi, cnt + 2, 1);
                    CK(cudaEventRecord(c.ev_join, c.aux_stream));
                    if (ch == 0) CK(cudaEventRecord(c.evp[1], ls));
                    /* the redo list: items outside the fast routines' domain or flagged by the conditioning guard (~0.3 %) */
                    s5::k_azimuth<s5::GEOD_TYPE_RR><<<g_rr, redo_threads, 0, ls>>>(cc, q, dd.phi, cnt + 1, 1);
                    CK(cudaStreamWaitEvent(ls, c.ev_join, 0));
                }
                }
                launches += 2;
            }
            if (ch == 0 && !defer) CK(cudaEventRecord(c.evp[2], ls));
            launches += 2;
        } else {
            if (p->outputs & SIM5_OUT_DELAY) {
                grid = persistent_grid(s5::k_trace_eqplane<false, true>, S5_EQ_THREADS, S5_EQ_DYN_SMEM);
                s5::k_trace_eqplane<false, true><<<grid, S5_EQ_THREADS, S5_EQ_DYN_SMEM, ls>>>(cc, dd, q, cnt, d_stats);
            } else {
                grid = persistent_grid(s5::k_trace_eqplane<false>, S5_EQ_THREADS, S5_EQ_DYN_SMEM);
                s5::k_trace_eqplane<false><<<grid, S5_EQ_THREADS, S5_EQ_DYN_SMEM, ls>>>(cc, dd, q, cnt, d_stats);
            }
        }
        launches += 1;
        CK(cudaGetLastError());
        if (devptr) continue;
        /* this chunk's rows go home on the copy stream while the next chunk computes */
        cudaStream_t cs = ls;
        if (nchunks > 1) {
            CK(cudaEventRecord(c.ev_chunk[ch], ls));
            CK(cudaStreamWaitEvent(c.copy_stream, c.ev_chunk[ch], 0));
            cs = c.copy_stream;
        } else {
            CK(cudaEventRecord(c.ev2, ls));
        }
        for (int i = 0; i < SIM5_NPLANES; i++) {
            if (!(p->outputs & kPlaneInfo[i].bit)) continue;
            size_t es = kPlaneInfo[i].elem;
            const char* src = (const char*)c.planes[i].p + pix0 * es;
            if (split == 1) {
                char* dst = (char*)host_plane(out, i) + ((size_t)rb + (size_t)lr0) * p->nx * es;
                CK(cudaMemcpyAsync(dst, src, (size_t)lrows * p->nx * es, cudaMemcpyDeviceToHost, cs));
            } else {
                size_t blk = (size_t)srows * p->nx * es;
                for (int b = 0; b < lrows / srows; b++) {
                    int iy0 = rb + ((lr0 / srows + b) * split + p->split_index) * srows;
                    CK(cudaMemcpyAsync((char*)host_plane(out, i) + (size_t)iy0 * p->nx * es, src + (size_t)b * blk, blk, cudaMemcpyDeviceToHost, cs));
                }
            }
        }
    }
    if (stage_copy) {
        /* the finished row blocks go to the caller's full-image planes by DMA: block j of this call is srows rows, `split` blocks apart in the image */
        CK(cudaEventRecord(c.ev_stage_go, ls));
        CK(cudaStreamWaitEvent(c.copy_stream, c.ev_stage_go, 0));
        if (defer) CK(cudaStreamWaitEvent(c.copy_stream, c.ev_redo_done[(cnt == c.d_counter2) ? 0 : 1], 0));      /* phi is complete after the redo passes */
        for (int i = 0; i < SIM5_NPLANES; i++) {
            if (!(p->outputs & kPlaneInfo[i].bit)) continue;
            const size_t es = kPlaneInfo[i].elem;
            const size_t blk = (size_t)srows * p->nx * es;
            char* dst = (char*)host_plane(out, i) + ((size_t)rb + (size_t)p->split_index * srows) * p->nx * es;
            CK(cudaMemcpy2DAsync(dst, blk * split, c.stage[sb][i].p, blk, blk, (size_t)(nrows_local / srows), cudaMemcpyDeviceToDevice, c.copy_stream));
        }
        CK(cudaEventRecord(c.ev_stage_done[sb], c.copy_stream));
        c.stage_pending[sb] = true;
    }
    c.phases = (nchunks == 1) ? launches : 0;      /* per-kernel times are defined for single-chunk calls only */
    c.ring_phases[c.ring_pos] = c.phases;
    c.ring_defer[c.ring_pos] = defer;
    c.last_defer = defer;
    if ((devptr || npix == 0) && !defer) CK(cudaEventRecord(c.ev2, ls));      /* (a deferred call ends at evp[1]) */
    if (alt) {
        const int b = (cnt == c.d_counter2) ? 0 : 1;
        CK(cudaEventRecord(c.ev_train_done[b], ls));
        c.train_pending[b] = true;
    }
    if (async) return SIM5_OK;
    if (nchunks > 1) {
        CK(cudaEventRecord(c.ev2, c.stream));
        CK(cudaEventRecord(c.ev_copy_done, c.copy_stream));
        CK(cudaStreamWaitEvent(c.stream, c.ev_copy_done, 0));
    }
    CK(cudaMemcpyAsync(c.h_stats, d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaEventRecord(c.ev3, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->rays = (int64_t)npix;
        for (int i = 0; i < 32; i++) stats->class_count[i] = (int64_t)c.h_stats->cls[i];
        for (int i = 0; i < 8; i++) stats->gtype_count[i] = (int64_t)c.h_stats->gtype[i];
        stats->total_steps = (int64_t)c.h_stats->steps;
        if (shared_queue) {          /* the rays THIS device pulled from the shared queue */
            stats->rays = 0;
            for (int i = 0; i < 32; i++) stats->rays += stats->class_count[i];
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev1, c.ev2); stats->kernel_ms = ms;
        cudaEventElapsedTime(&ms, c.ev1, c.ev3); stats->total_ms = ms;
        stats->kernel_launches = launches;
        stats->sm_count = c.sm_count; stats->grid_ctas = grid; stats->cta_threads = (p->mode == SIM5_MODE_SURFACE) ? s5::SurfaceProg::THREADS : lanes ? S5_CTA_THREADS : S5_EQ_THREADS;
    }
    return SIM5_OK;
}

/* ------------------------------------------------------------------ */
/* one call, several GPUs of the box                                   */
/* ------------------------------------------------------------------ */
namespace {
void add_stats(sim5_trace_stats* acc, const sim5_trace_stats& s, bool same_device)
{
    acc->rays += s.rays;
    for (int i = 0; i < 32; i++) acc->class_count[i] += s.class_count[i];
    for (int i = 0; i < 8; i++) acc->gtype_count[i] += s.gtype_count[i];
    acc->total_steps += s.total_steps;
    acc->kernel_launches += s.kernel_launches;
    if (same_device) { acc->kernel_ms += s.kernel_ms; acc->total_ms += s.total_ms; }     /* calls of one device run back to back */
    else { if (s.kernel_ms > acc->kernel_ms) acc->kernel_ms = s.kernel_ms; if (s.total_ms > acc->total_ms) acc->total_ms = s.total_ms; }
    acc->sm_count += same_device ? 0 : s.sm_count;
    if (s.grid_ctas > acc->grid_ctas) acc->grid_ctas = s.grid_ctas;
    acc->cta_threads = s.cta_threads;
}

struct MultiJob {
    int device = 0, index = 0, ndev = 1;
    const sim5_image_params* p = nullptr;
    const sim5_image_out* out = nullptr;
    int dev0 = 0;
    int rc = SIM5_OK;
    std::string err;
    sim5_trace_stats st;
    double* d_hist = nullptr;            /* HISTOGRAM: this device's partial lattice */
    uint64_t* shared_q = nullptr;        /* SIM5_FLAG_SHARED_QUEUE: the one ray queue of the call, on devices[0] */
    double spec[S5_SPEC_MAX_E];          /* SPECTRUM: this device's partial sum */
};

/* the share of device job->index: runs on a thread of its own, on that device's context */
void multi_worker(MultiJob* job)
{
    const sim5_image_params* p = job->p;
    memset(&job->st, 0, sizeof job->st);
    sim5_image_params q = *p;
    q.device = job->device;
    q.flags &= ~(uint32_t)(SIM5_FLAG_ASYNC | SIM5_FLAG_DEFER_REDO);
    const bool devptr = (p->flags & SIM5_FLAG_DEVICE_PTRS) != 0;
    auto fail = [&](int rc) { job->rc = rc; job->err = g_ctx.last_error; };
    if (devptr && job->device != job->dev0 && p->mode != SIM5_MODE_HISTOGRAM && p->mode != SIM5_MODE_SPECTRUM) {
        /* this GPU stores its rows straight into the planes of devices[0] over NVLink */
        Context& c = select_ctx(job->device);
        std::lock_guard<std::mutex> lk(c.mu);
        int rc = ensure_init(job->device);
        if (rc) return fail(rc);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, job->device, job->dev0);
        if (!can) { set_error("sim5_trace_image_multi: no peer access to the device that owns the planes"); return fail(SIM5_ERR_NOT_IMPL); }
        cudaError_t e = cudaDeviceEnablePeerAccess(job->dev0, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cuda_ok(e, "cudaDeviceEnablePeerAccess"); return fail(SIM5_ERR_CUDA); }
        cudaGetLastError();
    }
    if (p->mode == SIM5_MODE_HISTOGRAM) {
        /* partial lattice in this device's own buffer; sim5_trace_image_multi adds the buffers up on devices[0] */
        size_t hbytes = (size_t)p->n_spin * p->n_incl * p->n_bins * sizeof(double);
        {
            Context& c = select_ctx(job->device);
            std::lock_guard<std::mutex> lk(c.mu);
            int rc = ensure_init(job->device);
            if (!rc) rc = reserve(c.azq_f, hbytes);
            if (rc) return fail(rc);
            job->d_hist = (double*)c.azq_f.p;
        }
        sim5_image_out o = *job->out;
        o.hist = job->d_hist;
        q.flags |= SIM5_FLAG_DEVICE_PTRS;
        q.split_count = job->ndev; q.split_index = job->index;
        int rc = sim5_trace_image(&q, &o, &job->st);
        if (rc) fail(rc);
        return;
    }
    int rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = p->ny;
    /* interleaved row blocks: 32-row blocks while they divide evenly, single rows for what is left, the last < ndev rows to device 0 */
    const int nd = job->ndev;
    const int srows_a = p->split_rows > 0 ? p->split_rows : 32;
    const int rows = re - rb;
    const int n_a = rows / (nd * srows_a) * (nd * srows_a);
    const int n_b = (rows - n_a) / nd * nd;
    const int seg_begin[3] = {rb, rb + n_a, rb + n_a + n_b};
    const int seg_end[3]   = {rb + n_a, rb + n_a + n_b, re};
    const int seg_srows[3] = {srows_a, 1, 0};
    sim5_image_out o = *job->out;
    if (job->shared_q) {
        /* one queue for all devices: every GPU runs the whole row range and pulls rays from the counter on devices[0], a warp refill at a
         * time, until it is empty -- no split to balance, the GPUs finish within one ray of each other */
        q.row_begin = rb; q.row_end = re;
        q.split_count = 0; q.split_index = 0; q.split_rows = 0;
        o.shared_counter = job->shared_q;
        int rc = sim5_trace_image(&q, &o, &job->st);
        if (rc) fail(rc);
        return;
    }
    q.flags &= ~(uint32_t)SIM5_FLAG_SHARED_QUEUE;
    if (p->mode == SIM5_MODE_SPECTRUM) { for (int k = 0; k < S5_SPEC_MAX_E; k++) job->spec[k] = 0.0; }
    for (int sgm = 0; sgm < 3; sgm++) {
        if (seg_end[sgm] <= seg_begin[sgm]) continue;
        if (sgm == 2 && job->index != 0) continue;
        q.row_begin = seg_begin[sgm]; q.row_end = seg_end[sgm];
        if (sgm < 2) { q.split_count = nd; q.split_index = job->index; q.split_rows = seg_srows[sgm]; if (devptr) q.flags |= SIM5_FLAG_FULL_INDEX; }
        else         { q.split_count = 0; q.split_index = 0; q.split_rows = 0; }
        sim5_trace_stats st;
        memset(&st, 0, sizeof st);
        double part[S5_SPEC_MAX_E];
        if (p->mode == SIM5_MODE_SPECTRUM) o.spectrum = part;
        int rc = sim5_trace_image(&q, &o, &st);
        if (rc) return fail(rc);
        if (p->mode == SIM5_MODE_SPECTRUM) for (int k = 0; k < p->n_energy; k++) job->spec[k] += part[k];
        add_stats(&job->st, st, true);
    }
}
}

/* THE batched entry on several GPUs of one box: the same arguments as sim5_trace_image plus the device list.  One host thread per
 * device drives that device's context; image rows are dealt out in interleaved 32-row blocks (the expensive rows around the shadow go
 * to all GPUs), lattice images of HISTOGRAM mode one by one.  Host planes: every GPU copies its rows to the caller's planes itself
 * (its own PCIe link; chunks under its kernels).  SIM5_FLAG_DEVICE_PTRS: the planes live on devices[0] and the other GPUs store into
 * them over NVLink (peer access, SIM5_FLAG_FULL_INDEX).  HISTOGRAM: each GPU accumulates its images, devices[0] adds the partial
 * lattices up with loads from its peers' memory (k_sum_peers, fixed order).  SPECTRUM: partial sums added on the host.  p->device is
 * ignored; SIM5_FLAG_ASYNC / _DEFER_REDO do not apply.  stats: counters summed, times = the slowest device. */
extern "C" int sim5_trace_image_multi(const sim5_image_params* p, const sim5_image_out* out, sim5_trace_stats* stats, const int* devices, int ndev)
{
    if (!p || !out || !devices || ndev < 1 || ndev > S5_MAX_DEVICES) { set_error("sim5_trace_image_multi: bad arguments"); return SIM5_ERR_BAD_PARAM; }
    if (p->struct_size != (int32_t)sizeof(sim5_image_params)) { set_error("sim5_image_params.struct_size mismatch"); return SIM5_ERR_BAD_PARAM; }
    const int navail = sim5_gpu_device_count();
    if (navail <= 0) return no_device("device count is 0");
    for (int i = 0; i < ndev; i++) {
        if (devices[i] < 0 || devices[i] >= navail || devices[i] >= S5_MAX_DEVICES) { set_error("sim5_trace_image_multi: device ordinal out of range"); return SIM5_ERR_BAD_PARAM; }
        for (int j = 0; j < i; j++) if (devices[j] == devices[i]) { set_error("sim5_trace_image_multi: a device is listed twice"); return SIM5_ERR_BAD_PARAM; }
    }
    if (p->mode == SIM5_MODE_SPECTRUM && (p->n_energy < 1 || p->n_energy > S5_SPEC_MAX_E)) { set_error("bad spectrum grid"); return SIM5_ERR_BAD_PARAM; }
    if (p->mode == SIM5_MODE_HISTOGRAM && !out->hist) { set_error("HISTOGRAM mode needs out->hist"); return SIM5_ERR_NO_OUTPUT; }
    Context* caller = &g_ctx;
    uint64_t* shared_q = nullptr;
    if (p->flags & SIM5_FLAG_SHARED_QUEUE) {
        const bool lanes = (p->mode == SIM5_MODE_STEPWISE || p->mode == SIM5_MODE_SURFACE);
        if (!lanes || !(p->flags & SIM5_FLAG_DEVICE_PTRS)) { set_error("SIM5_FLAG_SHARED_QUEUE needs a lane mode (STEPWISE / SURFACE) and DEVICE_PTRS"); return SIM5_ERR_BAD_PARAM; }
        shared_q = out->shared_counter;
        if (!shared_q) {
            /* the library's own queue word on devices[0], zeroed before the first device starts */
            Context& c0 = select_ctx(devices[0]);
            std::lock_guard<std::mutex> lk(c0.mu);
            int rc0 = ensure_init(devices[0]);
            if (!rc0) rc0 = reserve(c0.shared_q, 64);
            if (!rc0 && !(cuda_ok(cudaMemsetAsync(c0.shared_q.p, 0, 64, c0.stream), "shared queue memset") && cuda_ok(cudaStreamSynchronize(c0.stream), "shared queue memset"))) rc0 = SIM5_ERR_CUDA;
            if (rc0) { caller->last_error = c0.last_error; t_ctx = caller; return rc0; }
            shared_q = (uint64_t*)c0.shared_q.p;
            t_ctx = caller;
        }
    }
    std::vector<MultiJob> jobs((size_t)ndev);
    std::vector<std::thread> th;
    for (int i = 0; i < ndev; i++) {
        jobs[i].device = devices[i]; jobs[i].index = i; jobs[i].ndev = ndev; jobs[i].p = p; jobs[i].out = out; jobs[i].dev0 = devices[0]; jobs[i].shared_q = shared_q;
    }
    for (int i = 1; i < ndev; i++) th.emplace_back(multi_worker, &jobs[i]);
    multi_worker(&jobs[0]);                                  /* the calling thread drives devices[0] */
    for (auto& t : th) t.join();
    int rc = SIM5_OK;
    for (int i = 0; i < ndev; i++) if (jobs[i].rc != SIM5_OK && rc == SIM5_OK) { rc = jobs[i].rc; caller->last_error = jobs[i].err; }
    if (rc == SIM5_OK && p->mode == SIM5_MODE_HISTOGRAM) {
        /* reduce on devices[0]: loads from the peers' partial lattices (staged through a peer copy where there is no peer access) */
        Context& c = select_ctx(devices[0]);
        std::lock_guard<std::mutex> lk(c.mu);
        rc = ensure_init(devices[0]);
        const long long n = (long long)p->n_spin * p->n_incl * p->n_bins;
        s5::PeerPtrs src;
        for (int i = 0; i < 16; i++) src.p[i] = nullptr;
        size_t staged = 0;
        for (int i = 0; i < ndev && rc == SIM5_OK; i++) {
            src.p[i] = jobs[i].d_hist;
            if (i == 0) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[0], devices[i]);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[i], 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); continue; }
            cudaGetLastError();
            rc = reserve(c.azq2_f, (size_t)(ndev - 1) * n * sizeof(double));
            if (rc) break;
            double* stage = (double*)c.azq2_f.p + staged * n;
            staged++;
            if (!cuda_ok(cudaMemcpyPeerAsync(stage, devices[0], jobs[i].d_hist, devices[i], n * sizeof(double), c.stream), "cudaMemcpyPeerAsync")) { rc = SIM5_ERR_CUDA; break; }
            src.p[i] = stage;
        }
        if (rc == SIM5_OK) {
            double* dst = jobs[0].d_hist;
            s5::k_sum_peers<<<c.sm_count * 4, 256, 0, c.stream>>>(dst, src, ndev, n);
            bool okc = cuda_ok(cudaGetLastError(), "k_sum_peers");
            if (okc) okc = cuda_ok(cudaMemcpyAsync(out->hist, dst, n * sizeof(double), (p->flags & SIM5_FLAG_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream), "histogram copy");
            if (okc) okc = cuda_ok(cudaStreamSynchronize(c.stream), "histogram reduce");
            if (!okc) rc = SIM5_ERR_CUDA;
        }
        if (rc != SIM5_OK) caller->last_error = c.last_error;
    }
    if (rc == SIM5_OK && p->mode == SIM5_MODE_SPECTRUM) {
        if (!out->spectrum) { set_error("SPECTRUM mode needs out->spectrum"); return SIM5_ERR_NO_OUTPUT; }
        for (int k = 0; k < p->n_energy; k++) { double s = 0.0; for (int i = 0; i < ndev; i++) s += jobs[i].spec[k]; out->spectrum[k] = s; }
    }
    t_ctx = caller;
    cudaSetDevice(caller->ready ? caller->device : devices[0]);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int i = 0; i < ndev; i++) add_stats(stats, jobs[i].st, false);
    }
    return rc;
}

extern "C" int sim5_last_phase_ms(double* ms, int n, int64_t* items)
{
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready) CK(cudaSetDevice(c.device));
    if (!c.ready || !ms || n < 1) { set_error("sim5_last_phase_ms: no context / no output"); return SIM5_ERR_BAD_PARAM; }
    if (c.phases < 1) { set_error("sim5_last_phase_ms: no image call recorded"); return SIM5_ERR_BAD_PARAM; }
    CK(cudaEventSynchronize(c.last_defer ? c.evp[1] : c.ev2));
    float t = 0;
    for (int i = 0; i < n; i++) ms[i] = 0.0;
    if (items) items[0] = items[1] = 0;
    if (c.phases == 1) { CK(cudaEventElapsedTime(&t, c.ev1, c.ev2)); ms[0] = t; return 1; }
    cudaEvent_t seq[4] = {c.ev1, c.evp[0], c.evp[1], c.evp[2]};
    for (int i = 0; i < (c.last_defer ? 2 : 3) && i < n; i++) { CK(cudaEventElapsedTime(&t, seq[i], seq[i + 1])); ms[i] = t; }
    if (items) {
        unsigned long long cnt[2] = {0, 0};
        CK(cudaMemcpy(cnt, c.last_counts, sizeof cnt, cudaMemcpyDeviceToHost));
        items[0] = (int64_t)cnt[0]; items[1] = (int64_t)cnt[1];
    }
    return c.phases;
}

/* per-kernel times of the image call `back` calls ago (0 = the most recent; up to S5_RING - 1 calls are kept): same layout as
 * sim5_last_phase_ms.  Waits for that call only.  Lets a caller time a train of SIM5_FLAG_ASYNC calls without a sync per call. */
extern "C" int sim5_phase_history(int back, double* ms, int n)
{
    Context& c = g_ctx;
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.ready) CK(cudaSetDevice(c.device));
    if (!c.ready || !ms || n < 1 || back < 0 || back >= S5_RING) { set_error("sim5_phase_history: bad arguments"); return SIM5_ERR_BAD_PARAM; }
    int pos = ((c.ring_pos - back) % S5_RING + S5_RING) % S5_RING;
    int phases = c.ring_phases[pos];
    if (phases < 1) { set_error("sim5_phase_history: no image call recorded in that slot"); return SIM5_ERR_BAD_PARAM; }
    cudaEvent_t* e = c.ring[pos];
    const bool dfr = c.ring_defer[pos];
    CK(cudaEventSynchronize(dfr ? e[2] : e[4]));
    float t = 0;
    for (int i = 0; i < n; i++) ms[i] = 0.0;
    if (phases == 1) { CK(cudaEventElapsedTime(&t, e[0], e[4])); ms[0] = t; return 1; }
    for (int i = 0; i < (dfr ? 2 : 3) && i < n; i++) { CK(cudaEventElapsedTime(&t, e[i], e[i + 1])); ms[i] = t; }
    return phases;
}

/* ------------------------------------------------------------------ */
/* FP64 peak                                                           */
/* ------------------------------------------------------------------ */
extern "C" double sim5_fp64_peak_tflops(int device, int iters)
{
    if (device >= S5_MAX_DEVICES) return -1.0;
    Context& c = select_ctx(device);
    std::lock_guard<std::mutex> lk(c.mu);
    if (ensure_init(device) != SIM5_OK) return -1.0;
    if (iters < 1) iters = 4096;
    int grid = c.sm_count * 8;
    if (reserve(c.planes[0], 4096) != SIM5_OK) return -1.0;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(c.ev0, c.stream);
        s5::k_dfma_peak<<<grid, 256, 0, c.stream>>>((double*)c.planes[0].p, iters, 1.0);
        cudaEventRecord(c.ev1, c.stream);
        if (!cuda_ok(cudaStreamSynchronize(c.stream), "dfma peak")) return -1.0;
        float ms = 0;
        cudaEventElapsedTime(&ms, c.ev0, c.ev1);
        double flops = (double)grid * 256.0 * (double)iters * 64.0 * 2.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    return best;
}

/* kernel-only time (ms) of `reps` calls per thread of one device routine over n threads; <0 on error */
extern "C" double sim5_micro_bench(int which, int64_t n, int reps)
{
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    if (ensure_init(-1) != SIM5_OK) return -1.0;
    Context& c = g_ctx;
    n = (n + 127) / 128 * 128;
    if (reserve(c.planes[0], (size_t)n * sizeof(double)) != SIM5_OK) return -1.0;
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
        cudaEventRecord(c.ev0, c.stream);
        s5::k_micro<<<(unsigned)(n / 128), 128, 0, c.stream>>>(which, reps, (double*)c.planes[0].p);
        cudaEventRecord(c.ev1, c.stream);
        if (!cuda_ok(cudaStreamSynchronize(c.stream), "micro bench")) return -1.0;
        float ms = 0; cudaEventElapsedTime(&ms, c.ev0, c.ev1);
        if (ms < best) best = ms;
    }
    return best;
}

/* ------------------------------------------------------------------ */
/* batched element-wise entries                                        */
/* ------------------------------------------------------------------ */
namespace {
int stage_in(int slot, const double* h, size_t n)
{
    Context& c = g_ctx;
    size_t bytes = (n ? n : 1) * sizeof(double);
    if (c.batch_bytes[slot] < bytes) {
        if (c.batch[slot]) cudaFree(c.batch[slot]);
        c.batch[slot] = nullptr; c.batch_bytes[slot] = 0;
        CK(cudaMalloc(&c.batch[slot], bytes));
        c.batch_bytes[slot] = bytes;
    }
    if (h && n) CK(cudaMemcpyAsync(c.batch[slot], h, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    return SIM5_OK;
}
int stage_out(int slot, double* h, size_t n)
{
    if (h && n) CK(cudaMemcpyAsync(h, g_ctx.batch[slot], n * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
    return SIM5_OK;
}
int batch_grid(int64_t n) { int64_t b = (n + 127) / 128; int64_t cap = (int64_t)g_ctx.sm_count * 16; return (int)(b < 1 ? 1 : (b > cap ? cap : b)); }
#define B(i) ((double*)g_ctx.batch[i])
#define BATCH_BEGIN() std::lock_guard<std::mutex> lk(g_ctx.mu); if (n < 0) return SIM5_ERR_BAD_PARAM; { int rc_ = ensure_init(-1); if (rc_) return rc_; }
#define BATCH_END() CK(cudaGetLastError()); CK(cudaStreamSynchronize(g_ctx.stream)); return SIM5_OK;
}

extern "C" int sim5_batch_rf(int64_t n, const double* x, const double* y, const double* z, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(2, z, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rf<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(2), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_rf_hi(int64_t n, const double* x, const double* y, const double* z, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(2, z, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rf_hi<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(2), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_rj_hi(int64_t n, const double* x, const double* y, const double* z, const double* p, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(2, z, n)) || (rc = stage_in(4, p, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rj_hi<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(2), B(4), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_rd(int64_t n, const double* x, const double* y, const double* z, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(2, z, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rd<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(2), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_rc(int64_t n, const double* x, const double* y, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rc<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_rj(int64_t n, const double* x, const double* y, const double* z, const double* p, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, x, n)) || (rc = stage_in(1, y, n)) || (rc = stage_in(2, z, n)) || (rc = stage_in(4, p, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    s5::k_batch_rj<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(2), B(4), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_sncndn(int64_t n, const double* u, const double* m, double* sn, double* cn, double* dn)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, u, n)) || (rc = stage_in(1, m, n)) || (rc = stage_in(3, nullptr, n)) || (rc = stage_in(4, nullptr, n)) || (rc = stage_in(5, nullptr, n))) return rc;
    s5::k_batch_sncndn<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, B(0), B(1), B(3), B(4), B(5));
    if ((rc = stage_out(3, sn, n)) || (rc = stage_out(4, cn, n)) || (rc = stage_out(5, dn, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_libm(int op, int64_t n, const double* a, const double* b, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, a, n)) || (rc = stage_in(1, b, b ? n : 0)) || (rc = stage_in(1, nullptr, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    if (!b) CK(cudaMemsetAsync(g_ctx.batch[1], 0, (n ? n : 1) * sizeof(double), g_ctx.stream));
    s5::k_batch_libm<<<batch_grid(n), 128, 0, g_ctx.stream>>>(op, n, B(0), B(1), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}

extern "C" int sim5_batch_integral(int op, int64_t n, const double* const* v, double* out)
{
    BATCH_BEGIN();
    if (!v) { set_error("null argument table"); return SIM5_ERR_BAD_PARAM; }
    int rc;
    for (int k = 0; k < 7; k++) {
        if ((rc = stage_in(k, v[k], v[k] ? n : 0)) || (rc = stage_in(k, nullptr, n))) return rc;
        if (!v[k]) CK(cudaMemsetAsync(g_ctx.batch[k], 0, (n ? n : 1) * sizeof(double), g_ctx.stream));
    }
    if ((rc = stage_in(7, nullptr, n))) return rc;
    s5::k_batch_integral<<<batch_grid(n), 128, 0, g_ctx.stream>>>(op, n, B(0), B(1), B(2), B(3), B(4), B(5), B(6), B(7));
    if ((rc = stage_out(7, out, n))) return rc;
    BATCH_END();
}
extern "C" int sim5_batch_timedelay(int64_t n, double incl, double a, const double* alpha, const double* beta, const double* ra, const double* rb, double* out)
{
    BATCH_BEGIN();
    int rc; if ((rc = stage_in(0, alpha, n)) || (rc = stage_in(1, beta, n)) || (rc = stage_in(2, ra, n)) || (rc = stage_in(4, rb, n)) || (rc = stage_in(3, nullptr, n))) return rc;
    double si, ci;
    sincos(incl, &si, &ci);
    s5::k_batch_timedelay<<<batch_grid(n), 128, 0, g_ctx.stream>>>(n, incl, si, ci, a, B(0), B(1), B(2), B(4), B(3));
    if ((rc = stage_out(3, out, n))) return rc;
    BATCH_END();
}

/* ------------------------------------------------------------------ */
/* scalar sim5lib.h API: one-thread device launches                    */
/* ------------------------------------------------------------------ */
template <class F> __global__ void k_call(F f) { f(); }

namespace {

template <class F> bool dev_call(F f)
{
    k_call<<<1, 1, 0, g_ctx.stream>>>(f);
    if (!cuda_ok(cudaGetLastError(), "scalar launch")) return false;
    return cuda_ok(cudaStreamSynchronize(g_ctx.stream), "scalar sync");
}
bool scalar_ready() { return ensure_init(-1) == SIM5_OK; }
#define H (g_ctx.h_scr)
#define D (g_ctx.d_scr)
#define LOCK std::lock_guard<std::mutex> lk(g_ctx.mu)
const double kNaN = NAN;

/* double f(double...) helpers */
double call_d1(int op, double a0, double a1, double a2, double a3, double a4, double a5)
{
    LOCK;
    if (!scalar_ready()) return kNaN;
    Scratch* d = D;
    bool ok = dev_call([=] __device__ () {
        double r = NAN;
        switch (op) {
            case 0: r = s5::rf(a0, a1, a2); break;
            case 1: r = s5::rd(a0, a1, a2); break;
            case 2: r = s5::rc(a0, a1); break;
            case 3: r = s5::rj(a0, a1, a2, a3); break;
            case 4: r = s5::elliptic_k(a0); break;
            case 5: r = s5::elliptic_f(a0, a1); break;
            case 6: r = s5::elliptic_f_cos(a0, a1); break;
            case 7: r = s5::elliptic_f_sin(a0, a1); break;
            case 8: r = s5::elliptic_e_cos(a0, a1); break;
            case 9: r = s5::elliptic_e_sin(a0, a1); break;
            case 10: r = s5::elliptic_pi_complete(a0, a1); break;
            case 11: r = s5::elliptic_pi_cos(a0, a1, a2); break;
            case 12: r = s5::elliptic_pi_sin(a0, a1, a2); break;
            case 13: r = s5::jacobi_isn(a0, a1); break;
            case 14: r = s5::jacobi_icn(a0, a1); break;
            case 15: r = s5::jacobi_itn(a0, a1); break;
            case 16: r = s5::jacobi_sn(a0, a1); break;
            case 17: r = s5::jacobi_cn(a0, a1); break;
            case 18: r = s5::jacobi_dn(a0, a1); break;
            case 19: r = s5::integral_Z1(a0, a1, a2, a3); break;
            case 20: r = s5::integral_R1(a0, a1, a2); break;
            case 21: r = s5::integral_R_rp_re(a0, a1, a2, a3, a4, a5); break;
            case 22: r = s5::integral_R_rp_re_inf(a0, a1, a2, a3, a4); break;
            case 23: r = s5::integral_T_mp(a0, a1, a2, a3); break;
            case 24: r = s5::r_bh(a0); break;
            case 25: r = s5::OmegaK(a0, a1); break;
            case 26: r = s5::gfactorK(a0, a1, a2); break;
            case 27: r = s5::ellK(a0, a1); break;
            case 28: r = s5::integral_C1(a0, a1); break;
            case 29: r = s5::integral_C2(a0, a1); break;
            case 30: r = s5::integral_C2_cos(a0, a1); break;
            case 31: r = s5::integral_Z2(a0, a1, a2, a3); break;
            case 32: r = s5::integral_Rm1(a0, a1, a2); break;
            case 33: r = s5::integral_Rm2(a0, a1, a2); break;
            case 34: r = s5::integral_R2(a0, a1, a2); break;
            case 35: r = s5::integral_R_r0_re(a0, a1, a2, a3, a4); break;
            case 36: r = s5::integral_R_r0_re_inf(a0, a1, a2, a3); break;
            case 37: r = s5::integral_R_r1_re(a0, a1, a2, a3, a4); break;
            case 38: r = s5::integral_R_r2_re(a0, a1, a2, a3, a4); break;
            case 39: r = s5::integral_T_m0(a0, a1, a2); break;
            case 40: r = s5::integral_T_m2(a0, a1, a2); break;
            case 41: r = s5::integral_R0(a0, a1); break;
            case 42: r = a0; break;                                      /* integral_C0, sim5elliptic.c:636-642 */
        }
        d->v[0] = r;
    });
    return ok ? H->v[0] : kNaN;
}

} /* anonymous namespace */

#define DFUN1(name, op) extern "C" double name(double a) { return call_d1(op, a, 0, 0, 0, 0, 0); }
#define DFUN2(name, op) extern "C" double name(double a, double b) { return call_d1(op, a, b, 0, 0, 0, 0); }
#define DFUN3(name, op) extern "C" double name(double a, double b, double c) { return call_d1(op, a, b, c, 0, 0, 0); }
#define DFUN4(name, op) extern "C" double name(double a, double b, double c, double d) { return call_d1(op, a, b, c, d, 0, 0); }
DFUN3(rf, 0) DFUN3(rd, 1) DFUN2(rc, 2) DFUN4(rj, 3)
DFUN1(elliptic_k, 4) DFUN2(elliptic_f, 5) DFUN2(elliptic_f_cos, 6) DFUN2(elliptic_f_sin, 7)
DFUN2(elliptic_e_cos, 8) DFUN2(elliptic_e_sin, 9) DFUN2(elliptic_pi_complete, 10)
DFUN3(elliptic_pi_cos, 11) DFUN3(elliptic_pi_sin, 12)
DFUN2(jacobi_isn, 13) DFUN2(jacobi_icn, 14) DFUN2(jacobi_itn, 15)
DFUN2(jacobi_sn, 16) DFUN2(jacobi_cn, 17) DFUN2(jacobi_dn, 18)
DFUN4(integral_Z1, 19) DFUN3(integral_R1, 20)
extern "C" double integral_R_rp_re(double a, double b, double c, double d, double p, double X) { return call_d1(21, a, b, c, d, p, X); }
extern "C" double integral_R_rp_re_inf(double a, double b, double c, double d, double p) { return call_d1(22, a, b, c, d, p, 0); }
DFUN4(integral_T_mp, 23)
DFUN1(r_bh, 24) DFUN2(OmegaK, 25) DFUN3(gfactorK, 26) DFUN2(ellK, 27)
extern "C" double integral_R_r0_re(double a, double b, double c, double d, double X) { return call_d1(35, a, b, c, d, X, 0); }
DFUN4(integral_R_r0_re_inf, 36)
extern "C" double integral_R_r1_re(double a, double b, double c, double d, double X) { return call_d1(37, a, b, c, d, X, 0); }
extern "C" double integral_R_r2_re(double a, double b, double c, double d, double X) { return call_d1(38, a, b, c, d, X, 0); }
DFUN3(integral_T_m0, 39) DFUN3(integral_T_m2, 40)

/* r_ms is a per-image constant in every caller; it needs cbrt of the host libm to match the reference bit for bit */
extern "C" double r_ms(double a) { return s5_host_r_ms(a); }

struct c_complex { double re, im; };

#define CCFUN(name, call, ...) extern "C" double name(__VA_ARGS__) \
{ LOCK; if (!scalar_ready()) return kNaN; Scratch* d = D; bool ok = dev_call([=] __device__ () { d->v[0] = call; }); return ok ? H->v[0] : kNaN; }
CCFUN(integral_R_r0_cc, s5::integral_R_r0_cc(a, b, c.re, c.im, X), double a, double b, c_complex c, double X)
CCFUN(integral_R_r0_cc_inf, s5::integral_R_r0_cc_inf(a, b, c.re, c.im), double a, double b, c_complex c)
CCFUN(integral_R_r1_cc, s5::integral_R_r1_cc(a, b, c.re, c.im, X1, X2), double a, double b, c_complex c, double X1, double X2)
CCFUN(integral_R_r2_cc, s5::integral_R_r2_cc(a, b, c.re, c.im, X1, X2), double a, double b, c_complex c, double X1, double X2)
CCFUN(integral_R_rp_cc2, s5::integral_R_rp_cc2(a, b, c.re, c.im, p, X1, X2), double a, double b, c_complex c, double p, double X1, double X2)

extern "C" double integral_R_rp_cc2_inf(double a, double b, c_complex c, double p, double X1)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D;
    bool ok = dev_call([=] __device__ () { d->v[0] = s5::integral_R_rp_cc2_inf(a, b, c.re, c.im, p, X1); });
    return ok ? H->v[0] : kNaN;
}

extern "C" void jacobi_sncndn(double u, double m, double* sn, double* cn, double* dn)
{
    LOCK;
    if (!scalar_ready()) { *sn = *cn = *dn = kNaN; return; }
    Scratch* d = D;
    bool ok = dev_call([=] __device__ () { s5::jacobi_sncndn(u, m, &d->v[0], &d->v[1], &d->v[2]); });
    *sn = ok ? H->v[0] : kNaN; *cn = ok ? H->v[1] : kNaN; *dn = ok ? H->v[2] : kNaN;
}

/* ---- metric / tetrads -------------------------------------------------------------------------- */
extern "C" void kerr_metric(double a, double r, double m, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) { memset(g, 0xff, sizeof *g); return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::kerr_metric(a, r, m, &d->m); });
    *g = H->m;
}
extern "C" void kerr_metric_contravariant(double a, double r, double m, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) { memset(g, 0xff, sizeof *g); return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::kerr_metric_contravariant(a, r, m, &d->m); });
    *g = H->m;
}
extern "C" void flat_metric(double r, double m, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) { memset(g, 0xff, sizeof *g); return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::flat_metric(r, m, &d->m); });
    *g = H->m;
}
extern "C" void kerr_connection(double a, double r, double m, double G[4][4][4])
{
    LOCK; if (!scalar_ready()) { for (int i = 0; i < 64; i++) (&G[0][0][0])[i] = kNaN; return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::Conn c; s5::kerr_connection(a, r, m, &c); s5::conn_to_array(&c, d->v); });
    memcpy(G, H->v, 64 * sizeof(double));
}
extern "C" void flat_connection(double r, double m, double G[4][4][4])
{
    LOCK; if (!scalar_ready()) { for (int i = 0; i < 64; i++) (&G[0][0][0])[i] = kNaN; return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::Conn c; s5::flat_connection(r, m, &c); s5::conn_to_array(&c, d->v); });
    memcpy(G, H->v, 64 * sizeof(double));
}
/* generic contraction with a caller-supplied G[4][4][4] (all 40 upper-triangle terms, reference order sim5kerr.c:421-440) */
extern "C" void Gamma(double G[4][4][4], double U[4], double V[4], double result[4])
{
    LOCK; if (!scalar_ready()) { result[0] = result[1] = result[2] = result[3] = kNaN; return; }
    Scratch* d = D;
    memcpy(H->v, G, 64 * sizeof(double));
    memcpy(H->v + 64, U, 32); memcpy(H->v + 68, V, 32);
    dev_call([=] __device__ () {
        const double* g = d->v; const double* u = d->v + 64; const double* v = d->v + 68;
        for (int i = 0; i < 4; i++) {
            double acc = 0.0;
            for (int j = 0; j < 4; j++) for (int k = j; k < 4; k++) acc -= 0.5 * g[i * 16 + j * 4 + k] * (u[j] * v[k] + u[k] * v[j]);
            d->v[72 + i] = acc;
        }
    });
    memcpy(result, H->v + 72, 32);
}
extern "C" double dotprod(double V1[4], double V2[4], s5::Metric* g)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D;
    memcpy(H->v, V1, 32); memcpy(H->v + 4, V2, 32);
    int flat = (g == nullptr);
    if (g) H->m = *g;
    dev_call([=] __device__ () {
        if (flat) d->v[8] = -d->v[0] * d->v[4] + d->v[1] * d->v[5] + d->v[2] * d->v[6] + d->v[3] * d->v[7];
        else d->v[8] = s5::dotprod(d->v, d->v + 4, &d->m);
    });
    return H->v[8];
}
extern "C" void vector_norm_to(double V[4], double norm, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) { V[0] = V[1] = V[2] = V[3] = kNaN; return; }
    Scratch* d = D;
    memcpy(H->v, V, 32); H->m = *g;
    dev_call([=] __device__ () { s5::vector_norm_to(d->v, norm, &d->m); });
    memcpy(V, H->v, 32);
}
extern "C" void tetrad_zamo(s5::Metric* g, s5::Tetrad* t)
{
    LOCK; if (!scalar_ready()) { memset(t, 0xff, sizeof *t); return; }
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { s5::tetrad_zamo(&d->m, &d->t); });
    *t = H->t;
}
extern "C" void tetrad_azimuthal(s5::Metric* g, double Omega, s5::Tetrad* t)
{
    LOCK; if (!scalar_ready()) { memset(t, 0xff, sizeof *t); return; }
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { s5::tetrad_azimuthal(&d->m, Omega, &d->t); });
    *t = H->t;
}
extern "C" void tetrad_surface(s5::Metric* g, double Omega, double V, double dhdr, s5::Tetrad* t)
{
    LOCK; if (!scalar_ready()) { memset(t, 0xff, sizeof *t); return; }
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { s5::tetrad_surface(&d->m, Omega, V, dhdr, &d->t); });
    *t = H->t;
}
extern "C" void bl2on(double Vin[4], double Vout[4], s5::Tetrad* t)
{
    LOCK; if (!scalar_ready()) { Vout[0] = Vout[1] = Vout[2] = Vout[3] = kNaN; return; }
    Scratch* d = D; memcpy(H->v, Vin, 32); H->t = *t;
    dev_call([=] __device__ () { s5::bl2on(d->v, d->v + 4, &d->t); });
    memcpy(Vout, H->v + 4, 32);
}
extern "C" void on2bl(double Vin[4], double Vout[4], s5::Tetrad* t)
{
    LOCK; if (!scalar_ready()) { Vout[0] = Vout[1] = Vout[2] = Vout[3] = kNaN; return; }
    Scratch* d = D; memcpy(H->v, Vin, 32); H->t = *t;
    dev_call([=] __device__ () { s5::on2bl(d->v, d->v + 4, &d->t); });
    memcpy(Vout, H->v + 4, 32);
}
extern "C" double Omega_from_ell(double ell, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { d->v[0] = s5::Omega_from_ell(ell, &d->m); });
    return H->v[0];
}
extern "C" double ell_from_Omega(double Omega, s5::Metric* g)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { d->v[0] = s5::ell_from_Omega(Omega, &d->m); });
    return H->v[0];
}
extern "C" void photon_momentum(double a, double r, double m, double l, double q, double r_sign, double m_sign, double k[4])
{
    LOCK; if (!scalar_ready()) { k[0] = k[1] = k[2] = k[3] = kNaN; return; }
    Scratch* d = D;
    dev_call([=] __device__ () { s5::photon_momentum(a, r, m, l, q, r_sign, m_sign, d->v); });
    memcpy(k, H->v, 32);
}
extern "C" void photon_motion_constants(double a, double r, double m, double k[4], double* L, double* Q)
{
    LOCK; if (!scalar_ready()) { *L = *Q = kNaN; return; }
    Scratch* d = D; memcpy(H->v, k, 32);
    dev_call([=] __device__ () { s5::photon_motion_constants(a, r, m, d->v, &d->v[4], &d->v[5]); });
    *L = H->v[4]; *Q = H->v[5];
}
extern "C" double photon_carter_const(double k[4], s5::Metric* g)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D; memcpy(H->v, k, 32); H->m = *g;
    dev_call([=] __device__ () { d->v[4] = s5::photon_carter_const(d->v, &d->m); });
    return H->v[4];
}
extern "C" void fourvelocity_azimuthal(double Omega, s5::Metric* g, double U[4])
{
    LOCK; if (!scalar_ready()) { U[0] = U[1] = U[2] = U[3] = kNaN; return; }
    Scratch* d = D; H->m = *g;
    dev_call([=] __device__ () { s5::fourvelocity_azimuthal(Omega, &d->m, d->v); });
    memcpy(U, H->v, 32);
}

/* ---- geodesics ----------------------------------------------------------------------------------- */
extern "C" int geodesic_init_inf(double i, double a, double alpha, double beta, s5::Geodesic* g, int* error)
{
    LOCK; if (!scalar_ready()) { if (error) *error = -1; return 0; }
    Scratch* d = D;
    H->g = *g;
    H->iv[1] = -1;
    dev_call([=] __device__ () { d->iv[0] = s5::geodesic_init_inf(i, a, alpha, beta, &d->g, &d->iv[1]); });
    *g = H->g;
    if (error && H->iv[1] != -1) *error = H->iv[1];
    return H->iv[0];
}
extern "C" int geodesic_init_src(double a, double r, double m, double k[4], int ppc, s5::Geodesic* g, int* error)
{
    LOCK; if (!scalar_ready()) { if (error) *error = -1; return 0; }
    Scratch* d = D;
    H->g = *g; memcpy(H->v, k, 32);
    H->iv[1] = -1;
    dev_call([=] __device__ () { d->iv[0] = s5::geodesic_init_src(a, r, m, d->v, ppc, &d->g, &d->iv[1]); });
    *g = H->g;
    if (error && H->iv[1] != -1) *error = H->iv[1];
    return H->iv[0];
}
#define GEO_D(name, expr, ...) extern "C" double name(s5::Geodesic* g, __VA_ARGS__) { \
    LOCK; if (!scalar_ready()) return kNaN; Scratch* d = D; H->g = *g; \
    dev_call([=] __device__ () { d->v[0] = expr; }); return H->v[0]; }
GEO_D(geodesic_P_int, s5::geodesic_P_int(&d->g, r, ppc), double r, int ppc)
GEO_D(geodesic_position_rad, s5::geodesic_position_rad(&d->g, P), double P)
GEO_D(geodesic_timedelay, s5::geodesic_timedelay(&d->g, P1, r1, m1, P2, r2, m2), double P1, double r1, double m1, double P2, double r2, double m2)
GEO_D(geodesic_position_pol, s5::geodesic_position_pol(&d->g, P), double P)
GEO_D(geodesic_position_pol_sign_k_theta, s5::geodesic_position_pol_sign_k_theta(&d->g, P), double P)
GEO_D(geodesic_position_azm, s5::geodesic_position_azm(&d->g, r, m, P), double r, double m, double P)
GEO_D(geodesic_dm_sign, s5::geodesic_dm_sign(&d->g, P), double P)
GEO_D(geodesic_find_midplane_crossing, s5::geodesic_find_midplane_crossing(&d->g, order), int order)
extern "C" void geodesic_position(s5::Geodesic*, double, double*) { /* empty in the reference too: sim5kerr-geod.c:268-285 */ }
extern "C" void geodesic_momentum(s5::Geodesic* g, double P, double r, double m, double k[])
{
    LOCK; if (!scalar_ready()) { k[0] = k[1] = k[2] = k[3] = kNaN; return; }
    Scratch* d = D; H->g = *g;
    memcpy(H->v, k, 32);
    dev_call([=] __device__ () { s5::geodesic_momentum(&d->g, P, r, m, d->v); });
    memcpy(k, H->v, 32);
}
extern "C" void geodesic_follow(s5::Geodesic* g, double step, double* P, double* r, double* m, int* status)
{
    LOCK; if (!scalar_ready()) { if (status) *status = 0; *r = *m = kNaN; return; }
    Scratch* d = D; H->g = *g;
    H->v[0] = *P; H->v[1] = *r; H->v[2] = *m; H->iv[0] = status ? *status : 0;
    dev_call([=] __device__ () { s5::geodesic_follow(&d->g, step, &d->v[0], &d->v[1], &d->v[2], &d->iv[0]); });
    *P = H->v[0]; *r = H->v[1]; *m = H->v[2];
    if (status) *status = H->iv[0];
}

/* ---- stepwise integrator ------------------------------------------------------------------------- */
extern "C" void raytrace_prepare(double bh_spin, double x[4], double k[4], double precision_factor, int options, s5::RayData* rtd)
{
    LOCK; if (!scalar_ready()) { memset(rtd, 0xff, sizeof *rtd); return; }
    Scratch* d = D; H->rtd = *rtd;
    memcpy(H->v, x, 32); memcpy(H->v + 4, k, 32);
    dev_call([=] __device__ () { s5::raytrace_prepare(bh_spin, d->v, d->v + 4, precision_factor, options, &d->rtd); });
    *rtd = H->rtd;
}
extern "C" void raytrace(double x[4], double k[4], double* step, s5::RayData* rtd)
{
    LOCK; if (!scalar_ready()) { x[1] = kNaN; return; }
    Scratch* d = D; H->rtd = *rtd;
    memcpy(H->v, x, 32); memcpy(H->v + 4, k, 32); H->v[8] = *step;
    dev_call([=] __device__ () { s5::raytrace(d->v, d->v + 4, &d->v[8], &d->rtd); });
    memcpy(x, H->v, 32); memcpy(k, H->v + 4, 32); *step = H->v[8];
    *rtd = H->rtd;
}
extern "C" double raytrace_error(double x[4], double k[4], s5::RayData* rtd)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D; H->rtd = *rtd;
    memcpy(H->v, x, 32); memcpy(H->v + 4, k, 32);
    dev_call([=] __device__ () { d->v[8] = s5::raytrace_error(d->v, d->v + 4, &d->rtd); });
    return H->v[8];
}

/* ---- polarization -------------------------------------------------------------------------------- */
extern "C" void polarization_vector(double k[4], c_complex wp, s5::Metric* g, double f[4])
{
    LOCK; if (!scalar_ready()) { f[0] = f[1] = f[2] = f[3] = kNaN; return; }
    Scratch* d = D; H->m = *g; memcpy(H->v, k, 32);
    dev_call([=] __device__ () { s5::polarization_vector(d->v, s5::Cplx{wp.re, wp.im}, &d->m, d->v + 4); });
    memcpy(f, H->v + 4, 32);
}
extern "C" c_complex polarization_constant(double k[4], double f[4], s5::Metric* g)
{
    LOCK; if (!scalar_ready()) return c_complex{kNaN, kNaN};
    Scratch* d = D; H->m = *g; memcpy(H->v, k, 32); memcpy(H->v + 4, f, 32);
    dev_call([=] __device__ () { s5::Cplx c = s5::polarization_constant(d->v, d->v + 4, &d->m); d->v[8] = c.re; d->v[9] = c.im; });
    return c_complex{H->v[8], H->v[9]};
}
extern "C" c_complex polarization_constant_infinity(double a, double alpha, double beta, double incl)
{
    LOCK; if (!scalar_ready()) return c_complex{kNaN, kNaN};
    Scratch* d = D;
    dev_call([=] __device__ () { s5::Cplx c = s5::polarization_constant_infinity_s(a, alpha, beta, crm::cr_sin(incl)); d->v[8] = c.re; d->v[9] = c.im; });
    return c_complex{H->v[8], H->v[9]};
}
extern "C" double polarization_angle_rotation(double a, double inc, double alpha, double beta, c_complex kappa)
{
    LOCK; if (!scalar_ready()) return kNaN;
    Scratch* d = D;
    dev_call([=] __device__ () { d->v[0] = s5::polarization_angle_rotation_s(a, crm::cr_sin(inc), alpha, beta, s5::Cplx{kappa.re, kappa.im}); });
    return H->v[0];
}

/* ---- Novikov-Thorne flux (per-pixel part only) ----------------------------------------------------- */
extern "C" int disk_nt_setup(double M, double a, double mdot_or_L, double alpha, int options)
{
    LOCK;
    if (options != 0) { set_error("disk_nt_setup: only options=0 (mdot given) is on the GPU path"); return -1; }
    sim5_image_params p;
    memset(&p, 0, sizeof p);
    p.nx = p.ny = 1; p.bh_spin = a; p.incl = 1.0; p.rmax = 1.0;
    p.disk_mass = M; p.disk_mdot = mdot_or_L; p.disk_alpha = alpha;
    g_ctx.disk_params = p;
    g_ctx.disk_set = true;
    return 0;
}
extern "C" double disk_nt_r_min(void)
{
    LOCK;
    double a = g_ctx.disk_set ? (double)(float)g_ctx.disk_params.bh_spin : 0.0;
    return s5_host_disk_nt_r_min(a);
}
extern "C" double disk_nt_flux(double r)
{
    LOCK; if (!scalar_ready()) return kNaN;
    if (!g_ctx.disk_set) {
        sim5_image_params p; memset(&p, 0, sizeof p);
        p.nx = p.ny = 1; p.incl = 1.0; p.rmax = 1.0; p.disk_mass = 10.0; p.disk_mdot = 0.1; p.disk_alpha = 0.1;
        g_ctx.disk_params = p; g_ctx.disk_set = true;
    }
    s5_fill_image_consts(&g_ctx.disk_params, g_ctx.h_consts);
    if (!cuda_ok(cudaMemcpyAsync(g_ctx.d_consts, g_ctx.h_consts, sizeof(S5ImageConsts), cudaMemcpyHostToDevice, g_ctx.stream), "consts")) return kNaN;
    Scratch* d = D;
    const S5ImageConsts* dc = g_ctx.d_consts;
    dev_call([=] __device__ () { d->v[0] = s5::disk_nt_flux(*dc, r); });
    return H->v[0];
}
