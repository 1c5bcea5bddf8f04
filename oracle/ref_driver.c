/*
 * ref_driver.c -- pixel-loop driver around the UNMODIFIED reference library.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with
 * /root/reference/src/sim5lib.c (where it lies; never copied) into
 * oracle/_ref/libsim5ref.so by oracle/Makefile.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load that library.
 *
 * The loops below are the caller-side loops the reference leaves to its users:
 *   eq-plane image  : examples/04-disk-image-eqplane/disk-image.c:53-105
 *   polarized pixel : python/sim5diskraytrace.py:250,340-391 + sim5polarization.c:144,271
 *   stepwise ray    : README.md:184-193, src/sim5unittests.c:116-127
 * generalised to the five BASELINE configs exactly as SURVEY.md 8(d) fixes them.
 * Every physics call goes to the reference's own functions; the harness-defined
 * pieces (Chandrasekhar table, torus, histogram binning) are spelled the same way
 * in the CUDA kernels (sim5_b200/csrc) and in the oracle port (oracle/sim5_oracle.c).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <unistd.h>
#include <fcntl.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sim5lib.h"          /* the reference header, from /root/reference/src */
#include "sim5_b200.h"        /* params/out structs shared with the product */

static int gtype_code(int type, int have)
{
    if (!have) return SIM5_GT_NONE;
    switch (type) {
        case GEOD_TYPE_RR:     return SIM5_GT_RR;
        case GEOD_TYPE_RC:     return SIM5_GT_RC;
        case GEOD_TYPE_CC:     return SIM5_GT_CC;
        case GEOD_TYPE_RR_DBL: return SIM5_GT_RR_DBL;
        case GEOD_TYPE_RR_BH:  return SIM5_GT_RR_BH;
    }
    return SIM5_GT_NONE;
}

typedef struct pixel_result {
    double r, phi, g, flux, chi, delta, mue, intensity, tau, qerr, height, delay;
    int steps;
    unsigned char status;
} pixel_result;

static double chandra_delta(double mue)
{
    double mu = fmin(fmax(mue, 0.0), 1.0);
    double t  = mu * (double)(SIM5_CHANDRA_N - 1);
    int    i0 = (int)t;
    if (i0 > SIM5_CHANDRA_N - 2) i0 = SIM5_CHANDRA_N - 2;
    double w  = t - (double)i0;
    return SIM5_CHANDRA_DELTA[i0] + (SIM5_CHANDRA_DELTA[i0+1] - SIM5_CHANDRA_DELTA[i0]) * w;
}

/* one pixel of the equatorial-plane image (modes EQPLANE and POLARIZED) */
static void pixel_eqplane(const sim5_image_params* p, double rmin, double alpha, double beta, pixel_result* o)
{
    geodesic gd;
    int error = 0;
    memset(o, 0, sizeof(*o));

    geodesic_init_inf(p->incl, p->bh_spin, alpha, beta, &gd, &error);
    if (error) {
        /* type is only meaningful when R_roots ran; report it for the codes raised after it */
        int have = (error == GD_ERROR_TYPE_RR_DOUBLE);
        o->status = (unsigned char)((SIM5_ST_INITERR + error) | (gtype_code(gd.type, have) << 5));
        return;
    }
    int gt = gtype_code(gd.type, 1) << 5;

    int order;
    for (order = 0; order <= p->max_order; order++) {
        double P = geodesic_find_midplane_crossing(&gd, order);
        if (isnan(P)) {
            o->status = (unsigned char)((order == 0 ? SIM5_ST_NOCROSS0 : order == 1 ? SIM5_ST_NOCROSS1 : SIM5_ST_NOCROSS2) | gt);
            return;
        }
        double r = geodesic_position_rad(&gd, P);
        if (r >= rmin) {
            o->status = (unsigned char)((order == 0 ? SIM5_ST_HIT0 : order == 1 ? SIM5_ST_HIT1 : SIM5_ST_HIT2) | gt);
            o->r = r;
            if (p->outputs & SIM5_OUT_PHI) o->phi = geodesic_position_azm(&gd, r, 0.0, P);
            if (p->outputs & SIM5_OUT_DELAY) {       /* travel time between the sphere r = delay_r_ref and the disk hit */
                double Pref = geodesic_P_int(&gd, p->delay_r_ref, 0);
                o->delay = geodesic_timedelay(&gd, Pref, p->delay_r_ref, 0.0, P, r, 0.0);
            }
            if (p->mode == SIM5_MODE_POLARIZED) {
                double a = p->bh_spin;
                double k[4], U[4], N[4], kl[4], fl[4], f[4];
                double e0[4] = {1.0, 0.0, 0.0, 0.0};
                double e2[4] = {0.0, 0.0, 1.0, 0.0};
                sim5metric m;
                sim5tetrad t;
                photon_momentum(a, r, 0.0, gd.l, gd.q, gd.Rpc - P, 1.0, k);
                kerr_metric(a, r, 0.0, &m);
                tetrad_azimuthal(&m, OmegaK(r, a), &t);
                on2bl(e0, U, &t);
                on2bl(e2, N, &t);
                double kU = dotprod(k, U, &m);
                double g  = (k[0]*m.g00 + k[3]*m.g03) / kU;
                double mue = dotprod(k, N, &m) / kU;
                /* polarization vector: parallel to the disk plane, perpendicular to the ray, in the fluid frame */
                bl2on(k, kl, &t);
                fl[0] = 0.0; fl[1] = -kl[3]; fl[2] = 0.0; fl[3] = kl[1];
                on2bl(fl, f, &t);
                vector_norm_to(f, 1.0, &m);
                sim5complex kappa = polarization_constant(k, f, &m);
                o->chi   = polarization_angle_rotation(a, p->incl, gd.alpha, gd.beta, kappa);
                o->mue   = mue;
                o->delta = chandra_delta(mue);
                o->g     = g;
                o->flux  = disk_nt_flux(r) * pow(g, 4.);
            } else {
                double g = gfactorK(r, p->bh_spin, gd.l);
                double f = disk_nt_flux(r);
                o->g    = g;
                o->flux = f * pow(g, 4.);
            }
            return;
        }
    }
    o->status = (unsigned char)(SIM5_ST_MISS | gt);
}

/* one stepwise ray through the harness torus (mode STEPWISE) */
static void pixel_stepwise(const sim5_image_params* p, double alpha, double beta, pixel_result* o)
{
    geodesic gd;
    int error = 0;
    memset(o, 0, sizeof(*o));
    double a = p->bh_spin;

    geodesic_init_inf(p->incl, a, alpha, beta, &gd, &error);
    if (error) {
        int have = (error == GD_ERROR_TYPE_RR_DOUBLE);
        o->status = (unsigned char)((SIM5_ST_INITERR + error) | (gtype_code(gd.type, have) << 5));
        return;
    }
    int gt = gtype_code(gd.type, 1) << 5;
    double r0 = p->r_start;
    if (!(r0 > gd.rp)) { o->status = (unsigned char)(SIM5_ST_NOSTART | gt); return; }

    double x[4], k[4];
    double P = geodesic_P_int(&gd, r0, 0);
    x[0] = 0.0;
    x[1] = r0;
    x[2] = geodesic_position_pol(&gd, P);
    x[3] = 0.0;                          /* torus is axisymmetric: azimuth origin is irrelevant */
    geodesic_momentum(&gd, P, r0, x[2], k);
    if (isnan(P) || isnan(x[2]) || isnan(k[1]) || isnan(k[2])) { o->status = (unsigned char)(SIM5_ST_NOSTART | gt); return; }

    raytrace_data rtd;
    raytrace_prepare(a, x, k, p->precision_factor, RTOPT_NONE, &rtd);

    double rh   = 1.05 * r_bh(a);
    double rout = 1.01 * r0;
    double Iacc = 0.0, tau = 0.0;
    int steps = 0, st;
    while (1) {
        double dl = p->step_max;
        raytrace(x, k, &dl, &rtd);
        steps++;
        /* emission / absorption of the harness torus at the new position */
        {
            double r = x[1], m = x[2];
            double R = r * sqrt(1.0 - m*m);
            double z = r * m;
            double s = sqr((R - p->torus_rc) / p->torus_w) + sqr(z / (p->torus_h * R));
            if (s < 13.8) {
                sim5metric M;
                kerr_metric(a, r, m, &M);
                double Om  = Omega_from_ell(p->torus_ell, &M);
                double den = M.g00 + 2.*Om*M.g03 + sqr(Om)*M.g33;
                if (den < 0.0) {
                    double U[4];
                    fourvelocity_azimuthal(Om, &M, U);
                    double g   = (k[0]*M.g00 + k[3]*M.g03) / dotprod(k, U, &M);
                    double rho = exp(-s);
                    double j   = p->torus_j0 * rho * rho;
                    double al  = p->torus_k0 * rho;
                    Iacc += j * g*g*g * exp(-tau) * dl;
                    tau += al * dl;
                }
            }
        }
        if (x[1] < rh)            { st = SIM5_ST_HORIZON;  break; }
        if (x[1] > rout)          { st = SIM5_ST_ESCAPE;   break; }
        if (rtd.error > 1e-2)     { st = SIM5_ST_ERRBREAK; break; }
        if (steps >= p->max_steps){ st = SIM5_ST_MAXSTEPS; break; }
    }
    o->intensity = Iacc;
    o->tau   = tau;
    o->steps = steps;
    o->qerr  = raytrace_error(x, k, &rtd);
    o->status = (unsigned char)(st | gt);
}


/* ------------------------------------------------------------------------------------------------------------
 * mode SURFACE: DiskRaytrace.geodesic / __find_surface / image of the reference's Python layer
 * (python/sim5diskraytrace.py:214-336, 163-205, 340-391) spelled in C, every physics call the reference's own.
 * The disk model (python: a dlopen'ed plug-in, src/sim5disk.c) is the harness surface of sim5_b200.h.
 * ------------------------------------------------------------------------------------------------------------ */
static double surf_rin(const sim5_image_params* p) { return (p->surf_rin > 0.0) ? p->surf_rin : r_ms(p->bh_spin); }
static double surf_h(const sim5_image_params* p, double R)
{
    double rin = surf_rin(p);
    return (R > rin) ? p->surf_hr * sqr(R - rin) / R : 0.0;
}
static double surf_dhdr(const sim5_image_params* p, double R)
{
    double rin = surf_rin(p);
    return (R > rin) ? p->surf_hr * (1.0 - sqr(rin) / sqr(R)) : 0.0;
}

/* __find_surface (python/sim5diskraytrace.py:259-336); the recursion on `iteration` is the outer loop.
 * Returns the termination class; on SIM5_ST_HIT0 / SIM5_ST_SURF_EQPLANE (P, r, m) is the result. */
static int find_surface(const sim5_image_params* p, geodesic* gd, double* Po, double* ro, double* mo, int* nfollow)
{
    const double accuracy = 1e-2;
    double disk_theta = atan(surf_h(p, 1e6) / 1e6);
    double rbh = r_bh(p->bh_spin);
    int iteration;
    for (iteration = 0; ; iteration++) {
        if (iteration > 3) return SIM5_ST_ESCAPE;
        double r0 = fmax(fmax(200.0, 1.1 * gd->rp), (0.5 + iteration) * sqrt(sqr(gd->alpha) + sqr(gd->beta)) / cos(gd->incl + disk_theta));
        double P1, r1, m1, R1, H1, Hd;
        while (1) {
            P1 = geodesic_P_int(gd, r0, 0);
            r1 = geodesic_position_rad(gd, P1);
            m1 = geodesic_position_pol(gd, P1);
            R1 = r1 * sqrt(1. - m1 * m1);
            H1 = r1 * m1;
            Hd = surf_h(p, R1);
            if ((Hd < H1) || (r0 > 5e6)) break;
            r0 = 2.0 * r0;
        }
        if (!(Hd < H1)) return SIM5_ST_SURF_BELOW;          /* python: if (Hd >= H1) -- NaN compares false there and goes on; see below */
        double P = P1, r = r1, m = m1;
        int status = 0, restart = 0;
        double step_factor = 1.0;
        while (1) {
            double step = fmax(accuracy / 2., fmin((H1 - Hd) / 2., 0.5 * (sqrt(r) - 0.99) * step_factor));
            geodesic_follow(gd, step, &P, &r, &m, &status); (*nfollow)++;
            if (!status) return SIM5_ST_SURF_LOST;
            R1 = r * sqrt(1. - m * m);
            H1 = r * m;
            Hd = surf_h(p, R1);
            if (H1 <= Hd) {
                if (step < accuracy) {
                    geodesic_follow(gd, -step / 2., &P, &r, &m, &status); (*nfollow)++;
                    *Po = P; *ro = r; *mo = m;
                    return SIM5_ST_HIT0;
                }
                geodesic_follow(gd, -step, &P, &r, &m, &status); (*nfollow)++;
                step_factor = step_factor / 5.;
                continue;
            }
            if (H1 < 1e-4) {
                *Po = geodesic_find_midplane_crossing(gd, 0);
                *ro = geodesic_position_rad(gd, *Po);
                *mo = geodesic_position_pol(gd, *Po);
                return SIM5_ST_SURF_EQPLANE;
            }
            if (r < 1.05 * rbh) return SIM5_ST_HORIZON;
            if (r > 1.1 * r0) { restart = 1; break; }
            if (m < 0.0) return SIM5_ST_SURF_UNDER;
            if (step < accuracy / 2.) break;
        }
        if (!restart) return SIM5_ST_MAXSTEPS;
    }
}

static void pixel_surface(const sim5_image_params* p, double alpha, double beta, pixel_result* o)
{
    geodesic gd;
    int error = 0;
    memset(o, 0, sizeof(*o));
    double a = p->bh_spin;

    geodesic_init_inf(p->incl, a, alpha, beta, &gd, &error);
    if (error) {
        int have = (error == GD_ERROR_TYPE_RR_DOUBLE);
        o->status = (unsigned char)((SIM5_ST_INITERR + error) | (gtype_code(gd.type, have) << 5));
        return;
    }
    int gt = gtype_code(gd.type, 1) << 5;
    double P, r, m;
    int cls, nfollow = 0;
    if (surf_h(p, 1e5) == 0.0) {                        /* flat=(self.disk.h(1e5)==0.0), python/sim5diskraytrace.py:170,240-243 */
        P = geodesic_find_midplane_crossing(&gd, 0);
        r = geodesic_position_rad(&gd, P);
        m = 0.0;
        cls = SIM5_ST_SURF_EQPLANE;
    } else {
        cls = find_surface(p, &gd, &P, &r, &m, &nfollow);
    }
    o->steps = nfollow;
    if (cls != SIM5_ST_HIT0 && cls != SIM5_ST_SURF_EQPLANE) { o->status = (unsigned char)(cls | gt); return; }
    if (isnan(r)) { o->status = (unsigned char)(SIM5_ST_MISS | gt); return; }
    o->status = (unsigned char)(cls | gt);

    double k[4];
    photon_momentum(a, r, m, gd.l, gd.q, gd.Rpc - P, 1.0, k);
    double R = r * sqrt(1. - m * m);
    o->r = r;
    o->height = r * m;
    double F = disk_nt_flux(R);
    if (F == 0.0) return;                                /* image(): if (F == 0.0): continue */

    sim5metric M;
    sim5tetrad t;
    double U[4], N[4];
    double e0[4] = {1.0, 0.0, 0.0, 0.0};
    double e2[4] = {0.0, 0.0, 1.0, 0.0};
    kerr_metric(a, r, m, &M);                            /* __tetrad, :340-349; disk.l = Keplerian, disk.vr = 0 */
    tetrad_surface(&M, Omega_from_ell(ellK(R, a), &M), 0.0, (m > 0.0) ? surf_dhdr(p, R) : 0.0, &t);
    on2bl(e0, U, &t);
    on2bl(e2, N, &t);
    double g = (k[0] * M.g00 + k[3] * M.g03) / dotprod(k, U, &M);      /* __gfactor, :353-362 */
    if (!(g > 0.0)) g = 0.0;
    double mue = dotprod(k, N, &M) / dotprod(k, U, &M);                 /* __emission_angle, :378-391 */
    if (mue < 0.0 && mue > -1e-2) mue = 1e-3;
    double limb = 0.5 + 0.75 * mue;
    if (!(g > 0.0)) return;                              /* image(): invalid g -> continue */
    o->g = g;
    o->mue = mue;
    o->flux = F * pow(g, 4.) * limb;
}

static int silence_stderr(void)
{
    fflush(stderr);
    int saved = dup(2);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 2); close(nul); }
    return saved;
}
static void restore_stderr(int saved)
{
    fflush(stderr);
    if (saved >= 0) { dup2(saved, 2); close(saved); }
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Trace rows [row_begin,row_end) of one image with the reference library.
 * nthreads <= 0 : all OpenMP threads.  Returns elapsed seconds of the pixel loop (>=0) or <0 on error.
 * The reference prints a line to stderr for some NaN results (sim5utils.c:41-54); `quiet` silences that.
 */
double ref_trace_image(const sim5_image_params* p, const sim5_image_out* out, int nthreads, int quiet, sim5_trace_stats* stats)
{
    if (!p || !out) return -1.0;
    if (p->mode == SIM5_MODE_HISTOGRAM) return -2.0;   /* use ref_trace_histogram */
    int nx = p->nx, ny = p->ny;
    int rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = ny;
    double a = p->bh_spin;
    double rms  = r_ms(a);
    double rmin = (p->r_emit_min > 0.0) ? p->r_emit_min : rms;
    double rmax = p->rmax;

    disk_nt_setup(p->disk_mass, a, p->disk_mdot, p->disk_alpha, 0);

    int saved = quiet ? silence_stderr() : -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
#else
    struct timespec ts0, ts1; clock_gettime(CLOCK_MONOTONIC, &ts0);
#endif
    int iy;
    #pragma omp parallel for schedule(dynamic,4)
    for (iy = rb; iy < re; iy++) {
        int ix;
        if (p->split_count > 1 && ((iy - rb) / (p->split_rows > 0 ? p->split_rows : 1)) % p->split_count != p->split_index) continue;
        for (ix = 0; ix < nx; ix++) {
            double alpha = (((double)(ix)+.5)/(double)(nx)-0.5)*2.0*rmax;
            double beta  = (((double)(iy)+.5)/(double)(ny)-0.5)*2.0*rmax * ((double)ny/(double)nx);
            pixel_result o;
            if (p->mode == SIM5_MODE_STEPWISE)     pixel_stepwise(p, alpha, beta, &o);
            else if (p->mode == SIM5_MODE_SURFACE) pixel_surface(p, alpha, beta, &o);
            else                                   pixel_eqplane(p, rmin, alpha, beta, &o);
            size_t i = (size_t)iy*(size_t)nx + (size_t)ix;
            if ((p->outputs & SIM5_OUT_R)         && out->r)         out->r[i] = o.r;
            if ((p->outputs & SIM5_OUT_PHI)       && out->phi)       out->phi[i] = o.phi;
            if ((p->outputs & SIM5_OUT_G)         && out->g)         out->g[i] = o.g;
            if ((p->outputs & SIM5_OUT_FLUX)      && out->flux)      out->flux[i] = o.flux;
            if ((p->outputs & SIM5_OUT_CHI)       && out->chi)       out->chi[i] = o.chi;
            if ((p->outputs & SIM5_OUT_DELTA)     && out->delta)     out->delta[i] = o.delta;
            if ((p->outputs & SIM5_OUT_MUE)       && out->mue)       out->mue[i] = o.mue;
            if ((p->outputs & SIM5_OUT_INTENSITY) && out->intensity) out->intensity[i] = o.intensity;
            if ((p->outputs & SIM5_OUT_TAU)       && out->tau)       out->tau[i] = o.tau;
            if ((p->outputs & SIM5_OUT_QERR)      && out->qerr)      out->qerr[i] = o.qerr;
            if ((p->outputs & SIM5_OUT_HEIGHT)    && out->height)    out->height[i] = o.height;
            if ((p->outputs & SIM5_OUT_DELAY)     && out->delay)     out->delay[i] = o.delay;
            if ((p->outputs & SIM5_OUT_STEPS)     && out->steps)     out->steps[i] = o.steps;
            if ((p->outputs & SIM5_OUT_STATUS)    && out->status)    out->status[i] = o.status;
        }
    }
#ifdef _OPENMP
    double dt = omp_get_wtime() - t0;
#else
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    double dt = (ts1.tv_sec-ts0.tv_sec) + 1e-9*(ts1.tv_nsec-ts0.tv_nsec);
#endif
    if (quiet) restore_stderr(saved);

    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->rays = (int64_t)(re-rb)*nx;
        if ((p->outputs & SIM5_OUT_STATUS) && out->status) {
            for (iy = rb; iy < re; iy++) for (int ix = 0; ix < nx; ix++) {
                unsigned char s = out->status[(size_t)iy*nx+ix];
                stats->class_count[SIM5_ST_CLASS(s)]++;
                stats->gtype_count[SIM5_ST_GTYPE(s)]++;
            }
        }
        if ((p->outputs & SIM5_OUT_STEPS) && out->steps)
            for (iy = rb; iy < re; iy++) for (int ix = 0; ix < nx; ix++) stats->total_steps += out->steps[(size_t)iy*nx+ix];
        stats->total_ms = dt*1e3;
    }
    return dt;
}

/*
 * Transfer-function lattice (mode HISTOGRAM, SURVEY.md 8d cfg 5): for every (spin j, inclination k)
 * image of nx*ny rays, a histogram of g in [g_min,g_max) weighted by F*g^4*dalpha*dbeta.
 * hist is [n_spin][n_incl][n_bins] doubles; only lattice images [lattice_begin,lattice_end) are
 * computed (0,0 = all) and only their slots written.
 */
double ref_trace_histogram(const sim5_image_params* p, double* hist, int nthreads, int quiet)
{
    if (!p || !hist) return -1.0;
    int nx = p->nx, ny = p->ny;
    int nimg = p->n_spin * p->n_incl;
    int lb = p->lattice_begin, le = p->lattice_end;
    if (lb == 0 && le == 0) le = nimg;
    int nb = p->n_bins;
    int saved = quiet ? silence_stderr() : -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
    int nthr = omp_get_max_threads();
#else
    int nthr = 1;
    struct timespec ts0, ts1; clock_gettime(CLOCK_MONOTONIC, &ts0);
#endif
    double* part = (double*)malloc(sizeof(double)*(size_t)nb*(size_t)ny);
    for (int img = lb; img < le; img++) {
        int js = img / p->n_incl, ki = img % p->n_incl;
        sim5_image_params q = *p;
        q.mode = SIM5_MODE_EQPLANE;
        q.bh_spin = (p->n_spin > 1) ? p->spin_max*(double)js/(double)(p->n_spin-1) : p->spin_max;
        if (q.bh_spin < 1e-4) q.bh_spin = 1e-4;
        double ideg = (p->n_incl > 1) ? p->incl_min_deg + (p->incl_max_deg-p->incl_min_deg)*(double)ki/(double)(p->n_incl-1) : p->incl_min_deg;
        q.incl = deg2rad(ideg);
        double rms = r_ms(q.bh_spin);
        q.rmax = rms + p->rmax_offset;
        double rmin = rms;
        double da = 2.0*q.rmax/(double)nx;
        double db = 2.0*q.rmax*((double)ny/(double)nx)/(double)ny;
        disk_nt_setup(q.disk_mass, q.bh_spin, q.disk_mdot, q.disk_alpha, 0);
        memset(part, 0, sizeof(double)*(size_t)nb*(size_t)ny);
        int iy;
        #pragma omp parallel for schedule(dynamic,4)
        for (iy = 0; iy < ny; iy++) {
            double* h = part + (size_t)iy*nb;      /* one partial histogram per row: fixed summation order */
            for (int ix = 0; ix < nx; ix++) {
                double alpha = (((double)(ix)+.5)/(double)(nx)-0.5)*2.0*q.rmax;
                double beta  = (((double)(iy)+.5)/(double)(ny)-0.5)*2.0*q.rmax * ((double)ny/(double)nx);
                pixel_result o;
                pixel_eqplane(&q, rmin, alpha, beta, &o);
                int cls = SIM5_ST_CLASS(o.status);
                if (cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2) {
                    double t = (o.g - p->g_min)/(p->g_max - p->g_min)*(double)nb;
                    if (t >= 0.0 && t < (double)nb) h[(int)t] += o.flux*da*db;
                }
            }
        }
        double* H = hist + (size_t)img*nb;
        for (int b = 0; b < nb; b++) { double s = 0.0; for (iy = 0; iy < ny; iy++) s += part[(size_t)iy*nb+b]; H[b] = s; }
    }
    free(part);
    (void)nthr;
#ifdef _OPENMP
    double dt = omp_get_wtime() - t0;
#else
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    double dt = (ts1.tv_sec-ts0.tv_sec) + 1e-9*(ts1.tv_nsec-ts0.tv_nsec);
#endif
    if (quiet) restore_stderr(saved);
    return dt;
}

/*
 * Thermal disk spectrum (mode SPECTRUM): DiskRaytrace.spectrum of the reference's Python layer
 * (python/sim5diskraytrace.py:43-134) on the image grid instead of its polar grid: for every disk hit
 * T = (F/sigma_SB)^(1/4) (python/sim5diskmodel.py:48), g and the emission cosine from the Keplerian emitter frame
 * (:340-391), spectrum += blackbody(T, hardf, mu_e or -1, energies/g) * g^3 * dalpha*dbeta (:124).  Every physics call
 * is the reference's own (blackbody() of sim5radiation.c:56-78).  spec is [n_energy]; rows [row_begin,row_end) and the
 * interleaved split are honoured so partial spectra of a multi-GPU job can be checked too.  One partial sum per row, rows
 * added in row order.
 */
double ref_trace_spectrum(const sim5_image_params* p, double* spec, int nthreads, int quiet)
{
    if (!p || !spec || p->n_energy < 1 || p->n_energy > 256) return -1.0;
    int nx = p->nx, ny = p->ny, ne = p->n_energy;
    int rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = ny;
    double a = p->bh_spin;
    double rmin = (p->r_emit_min > 0.0) ? p->r_emit_min : r_ms(a);
    double rmax = p->rmax;
    double da = 2.0*rmax/(double)nx;
    double db = 2.0*rmax*((double)ny/(double)nx)/(double)ny;
    double dA = da*db;
    double E[256];
    int k;
    for (k = 0; k < ne; k++) E[k] = (ne <= 1) ? p->e_min_kev : p->e_min_kev*pow(p->e_max_kev/p->e_min_kev, (double)k/(double)(ne-1));
    sim5_image_params q = *p;
    q.mode = SIM5_MODE_POLARIZED;
    q.outputs = SIM5_OUT_R | SIM5_OUT_G | SIM5_OUT_MUE;
    disk_nt_setup(p->disk_mass, a, p->disk_mdot, p->disk_alpha, 0);
    int saved = quiet ? silence_stderr() : -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
#else
    struct timespec ts0, ts1; clock_gettime(CLOCK_MONOTONIC, &ts0);
#endif
    double* part = (double*)calloc((size_t)ne*(size_t)ny, sizeof(double));
    int iy;
    #pragma omp parallel for schedule(dynamic,4)
    for (iy = rb; iy < re; iy++) {
        if (p->split_count > 1 && ((iy - rb) / (p->split_rows > 0 ? p->split_rows : 1)) % p->split_count != p->split_index) continue;
        double* h = part + (size_t)iy*ne;
        double Eg[256], Iv[256];
        for (int ix = 0; ix < nx; ix++) {
            double alpha = (((double)(ix)+.5)/(double)(nx)-0.5)*2.0*rmax;
            double beta  = (((double)(iy)+.5)/(double)(ny)-0.5)*2.0*rmax * ((double)ny/(double)nx);
            pixel_result o;
            pixel_eqplane(&q, rmin, alpha, beta, &o);
            int cls = SIM5_ST_CLASS(o.status);
            if (!(cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2)) continue;
            double T = sqrt(sqrt(disk_nt_flux(o.r)/5.670400e-05));
            if (!(T > 0.0) || !(o.g > 0.0)) continue;
            for (int j = 0; j < ne; j++) Eg[j] = E[j]/o.g;
            blackbody(T, p->spec_hardf, (p->spec_limb && o.mue >= 0.0) ? o.mue : -1.0, Eg, Iv, ne);
            for (int j = 0; j < ne; j++) h[j] += Iv[j]*(o.g*o.g*o.g)*dA;
        }
    }
    for (k = 0; k < ne; k++) { double s = 0.0; for (iy = rb; iy < re; iy++) s += part[(size_t)iy*ne+k]; spec[k] = s; }
    free(part);
#ifdef _OPENMP
    double dt = omp_get_wtime() - t0;
#else
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    double dt = (ts1.tv_sec-ts0.tv_sec) + 1e-9*(ts1.tv_nsec-ts0.tv_nsec);
#endif
    if (quiet) restore_stderr(saved);
    return dt;
}

/* element-wise wrappers (n-element SoA loops over single reference functions) for unit parity tests */
void ref_batch_rf(long n, const double* x, const double* y, const double* z, double* o) { for (long i=0;i<n;i++) o[i]=rf(x[i],y[i],z[i]); }
void ref_batch_rd(long n, const double* x, const double* y, const double* z, double* o) { for (long i=0;i<n;i++) o[i]=rd(x[i],y[i],z[i]); }
void ref_batch_rc(long n, const double* x, const double* y, double* o) { for (long i=0;i<n;i++) o[i]=rc(x[i],y[i]); }
void ref_batch_rj(long n, const double* x, const double* y, const double* z, const double* p, double* o) { for (long i=0;i<n;i++) o[i]=rj(x[i],y[i],z[i],p[i]); }
void ref_batch_sncndn(long n, const double* u, const double* m, double* sn, double* cn, double* dn) { for (long i=0;i<n;i++) jacobi_sncndn(u[i],m[i],&sn[i],&cn[i],&dn[i]); }
/* libm as the reference sees it (glibc of this image): op codes of sim5_batch_libm */
void ref_batch_libm(int op, long n, const double* a, const double* b, double* o)
{
    for (long i=0;i<n;i++) {
        switch (op) {
            case 0: o[i]=sin(a[i]); break;
            case 1: o[i]=cos(a[i]); break;
            case 2: o[i]=log(a[i]); break;
            case 3: o[i]=atan2(a[i],b[i]); break;
            case 4: o[i]=acos(a[i]); break;
            case 5: o[i]=asin(a[i]); break;
            case 6: o[i]=atan(a[i]); break;
            case 7: o[i]=pow(a[i],1./3.); break;
            case 8: o[i]=pow(a[i],1.5); break;
            case 9: o[i]=pow(a[i],4.); break;
            case 10: o[i]=exp(a[i]); break;
            default: o[i]=NAN;
        }
    }
}
double ref_r_ms(double a) { return r_ms(a); }
double ref_r_bh(double a) { return r_bh(a); }

/* the integrals behind geodesic_timedelay, element-wise (op codes shared with sim5_batch_integral of the product):
 * 0 C1(u,m) 1 C2(u,m) 2 C2_cos(c,m) 3 Z2(a,b,u,m) 4 Rm1(a,u,m) 5 Rm2(a,u,m) 6 R2(a,u,m) 7 R_r0_re(a,b,c,d,X) 8 R_r0_re_inf(a,b,c,d)
 * 9 R_r1_re(a,b,c,d,X) 10 R_r2_re(a,b,c,d,X) 11 T_m0(a2,b2,X) 12 T_m2(a2,b2,X) 13 R_r0_cc(a,b,(c,d),X) 14 R_r0_cc_inf(a,b,(c,d))
 * 15 R_r1_cc(a,b,(c,d),X1,X2) 16 R_r2_cc(..) 17 R_rp_cc2(a,b,(c,d),p,X1,X2) with X1 = v[4], X2 = v[5], p = v[6] */
/* internal to sim5elliptic.c (no prototype in sim5elliptic.h); the unity build emits them as external symbols */
double integral_C1(double u, double m);
double integral_C2(double u, double m);
double integral_C2_cos(double cn_u, double m);
double integral_Z2(double a, double b, double u, double m);
double integral_Rm1(double a, double u, double m);
double integral_Rm2(double a, double u, double m);
double integral_R2(double a, double u, double m);
void ref_batch_integral(int op, long n, const double* v0, const double* v1, const double* v2, const double* v3, const double* v4,
                        const double* v5, const double* v6, double* o)
{
    for (long i = 0; i < n; i++) {
        sim5complex c = makeComplex(v2[i], v3[i]);
        switch (op) {
            case 0: o[i] = integral_C1(v0[i], v1[i]); break;
            case 1: o[i] = integral_C2(v0[i], v1[i]); break;
            case 2: o[i] = integral_C2_cos(v0[i], v1[i]); break;
            case 3: o[i] = integral_Z2(v0[i], v1[i], v2[i], v3[i]); break;
            case 4: o[i] = integral_Rm1(v0[i], v1[i], v2[i]); break;
            case 5: o[i] = integral_Rm2(v0[i], v1[i], v2[i]); break;
            case 6: o[i] = integral_R2(v0[i], v1[i], v2[i]); break;
            case 7: o[i] = integral_R_r0_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 8: o[i] = integral_R_r0_re_inf(v0[i], v1[i], v2[i], v3[i]); break;
            case 9: o[i] = integral_R_r1_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 10: o[i] = integral_R_r2_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 11: o[i] = integral_T_m0(v0[i], v1[i], v2[i]); break;
            case 12: o[i] = integral_T_m2(v0[i], v1[i], v2[i]); break;
            case 13: o[i] = integral_R_r0_cc(v0[i], v1[i], c, v4[i]); break;
            case 14: o[i] = integral_R_r0_cc_inf(v0[i], v1[i], c); break;
            case 15: o[i] = integral_R_r1_cc(v0[i], v1[i], c, v4[i], v5[i]); break;
            case 16: o[i] = integral_R_r2_cc(v0[i], v1[i], c, v4[i], v5[i]); break;
            case 17: o[i] = integral_R_rp_cc2(v0[i], v1[i], c, v6[i], v4[i], v5[i]); break;
            default: o[i] = NAN;
        }
    }
}
/* geodesic_timedelay between two radii on the way in (P = geodesic_P_int(r, 0)) of n geodesics from infinity */
void ref_batch_timedelay(long n, double incl, double a, const double* alpha, const double* beta, const double* ra, const double* rb, double* o)
{
    int saved = silence_stderr();
    for (long i = 0; i < n; i++) {
        geodesic gd; int error = 0;
        o[i] = NAN;
        geodesic_init_inf(incl, a, alpha[i], beta[i], &gd, &error);
        if (error) continue;
        double Pa = geodesic_P_int(&gd, ra[i], 0), Pb = geodesic_P_int(&gd, rb[i], 0);
        o[i] = geodesic_timedelay(&gd, Pa, 0.0, 0.0, Pb, 0.0, 0.0);     /* r = 0: radii and latitudes recomputed from P */
    }
    restore_stderr(saved);
}
