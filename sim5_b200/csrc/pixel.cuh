/*
 * pixel.cuh -- what ONE ray does: the per-pixel pipelines of the trace modes, written once
 * as device functions and called by the kernels in kernels.cu (one ray per thread).
 *
 * The pipelines replace the caller-side pixel loops of the reference
 * (examples/04-disk-image-eqplane/disk-image.c:53-105, python/sim5diskraytrace.py:163-205,340-391,
 * README.md:184-193).  Each result equals what those loops produce with the reference library, but
 * the redundant work of the reference call chain is removed where that is bit-exact:
 *   - K(mm) and cn^-1(cos_i/sqrt(m2p), mm) are evaluated once and shared by Tpp, Tip and every
 *     crossing order (geodesic_find_midplane_crossing recomputes both, sim5kerr-geod.c:870-878);
 *   - the Novikov-Thorne coefficients are per-image constants (image_consts.h).
 */
#ifndef SIM5_PIXEL_CUH
#define SIM5_PIXEL_CUH

#include "geod.cuh"
#include "raytrace.cuh"
#include "polar.cuh"
#include "image_consts.h"

namespace s5 {

struct PixelOut {
    double r, phi, g, flux, chi, delta, mue, intensity, tau, qerr;
    int steps;
    unsigned status;
};

S5_HD S5_INL int gtype_code(int type)
{
    switch (type) {
        case GEOD_TYPE_RR:     return SIM5_GT_RR;
        case GEOD_TYPE_RC:     return SIM5_GT_RC;
        case GEOD_TYPE_CC:     return SIM5_GT_CC;
        case GEOD_TYPE_RR_DBL: return SIM5_GT_RR_DBL;
        case GEOD_TYPE_RR_BH:  return SIM5_GT_RR_BH;
    }
    return SIM5_GT_NONE;
}

/* pixel -> impact parameters, disk-image.c:57-58 */
S5_HD S5_INL void pixel_impact(const S5ImageConsts& c, int ix, int iy, double* alpha, double* beta)
{
    *alpha = (((double)(ix) + .5) / (double)(c.nx) - 0.5) * 2.0 * c.rmax;
    *beta  = (((double)(iy) + .5) / (double)(c.ny) - 0.5) * 2.0 * c.rmax * c.aspect;
}

/* Page-Thorne flux with the per-image pieces hoisted.  sim5disk-nt.c:109-146 */
S5_HD S5_INL double disk_nt_flux(const S5ImageConsts& c, double r)
{
    if (r <= c.nt_rms) return 0.0;
    double x = sqrt(r);
    double f0 = x - c.nt_x0 - c.nt_k0 * cr_log(x / c.nt_x0);
    double f1 = c.nt_k1 * cr_log((x - c.nt_x1) / c.nt_d1);
    double f2 = c.nt_k2 * cr_log((x - c.nt_x2) / c.nt_d2);
    double f3 = c.nt_k3 * cr_log((x - c.nt_x3) / c.nt_d3);
    double F = 1. / (c.nt_4pi * r) * 1.5 / (x * x * (x * x * x - 3. * x + c.nt_2a)) * (f0 - f1 - f2 - f3);
    return 9.1721376255e+28 * F * c.nt_mdot / c.nt_mass;
}

/* harness: Chandrasekhar limb polarization degree, linear interpolation (sim5_b200.h table) */
S5_HD S5_INL double chandra_delta(const double* tab, double mue)
{
    double mu = fmin(fmax(mue, 0.0), 1.0);
    double t = mu * (double)(SIM5_CHANDRA_N - 1);
    int i0 = (int)t;
    if (i0 > SIM5_CHANDRA_N - 2) i0 = SIM5_CHANDRA_N - 2;
    double w = t - (double)i0;
    return tab[i0] + (tab[i0 + 1] - tab[i0]) * w;
}

/* geodesic_init_inf with the two T-integrals' Carlson values kept for the crossings.
 * Same results as geodesic_init_inf_sc (geod.cuh); K_mm = K(mm), icn_u = cn^-1(cos_i/sqrt(m2p), mm). */
S5_HD S5_INL int init_inf_cached(const S5ImageConsts& c, double alpha, double beta, Geodesic* g, int* error, double* K_mm, double* icn_u)
{
    double a = c.a, i = c.incl;
    if ((a < 0.0) || (a > 1. - 1e-6)) { *error = GD_ERROR_SPIN_RANGE; return 0; }
    if ((i <= 0.0) || (i >= S5_PI_HALF)) { *error = GD_ERROR_INCL_RANGE; return 0; }
    if (beta == 0.0) beta = +1e-6;
    g->a = fmax(1e-4, a);
    g->incl = i;
    g->cos_i = c.cos_i;
    g->alpha = alpha;
    g->beta = beta;
    g->l = -alpha * c.sin_i;
    g->q = sq(beta) + sq(c.cos_i) * (sq(alpha) - sq(a));
    if (g->q == 0.0) { *error = GD_ERROR_Q_RANGE; return 0; }
    if (!geodesic_R_roots(g, 1.7976931348623157e308, error)) return 0;
    if (!geodesic_T_roots(g, g->cos_i, error)) return 0;
    /* theta_int(0) == mK*K(mm): jacobi_icn(0/sqrt(m2p), mm) takes its z == 0 exit (0 <= mm < 1 here) */
    *K_mm = elliptic_k(g->mm);
    *icn_u = jacobi_icn(g->cos_i / sqrt(g->m2p), g->mm);
    g->Tpp = 2. * (g->mK * (*K_mm));
    g->Tip = g->mK * (*icn_u);
    *error = GD_OK;
    return 1;
}
/* geodesic_find_midplane_crossing on the cached values.  sim5kerr-geod.c:845-885 */
S5_HD S5_INL double crossing_cached(const Geodesic* g, int order, double K_mm, double icn_u)
{
    if (g->q <= 0.0) return NAN;
    double u = g->cos_i / sqrt(g->m2p);
    double u0 = u;
    if (!ensure_range(&u, -1.0, +1.0, 1e-4)) return NAN;
    double icn = (u == u0) ? icn_u : jacobi_icn(u, g->mm);
    double pos;
    if (g->beta > 0.0)      pos = g->mK * ((2. * (double)order + 1.) * K_mm + icn);
    else if (g->beta < 0.0) pos = g->mK * ((2. * (double)order + 1.) * K_mm - icn);
    else                    pos = g->mK * ((2. * (double)order + 1.) * K_mm);
    if (pos > 2. * g->Rpc) pos = NAN;
    return pos;
}

/* emission-side quantities of the polarized mode for a disk hit at (r, m=0), position parameter P */
S5_HD S5_INL void polarized_hit(const S5ImageConsts& c, const Geodesic* gd, double r, double P, PixelOut* o)
{
    double a = c.a;
    double k[4], U[4], N[4], kl[4], fl[4], f[4];
    const double e0[4] = {1.0, 0.0, 0.0, 0.0};
    const double e2[4] = {0.0, 0.0, 1.0, 0.0};
    Metric m;
    Tetrad t;
    photon_momentum(a, r, 0.0, gd->l, gd->q, gd->Rpc - P, 1.0, k);
    kerr_metric(a, r, 0.0, &m);
    tetrad_azimuthal(&m, OmegaK(r, a), &t);
    on2bl(e0, U, &t);
    on2bl(e2, N, &t);
    double kU = dotprod(k, U, &m);
    double g = (k[0] * m.g00 + k[3] * m.g03) / kU;
    double mue = dotprod(k, N, &m) / kU;
    bl2on(k, kl, &t);
    fl[0] = 0.0; fl[1] = -kl[3]; fl[2] = 0.0; fl[3] = kl[1];
    on2bl(fl, f, &t);
    vector_norm_to(f, 1.0, &m);
    Cplx kappa = polarization_constant(k, f, &m);
    o->chi = polarization_angle_rotation_s(a, c.sin_i, gd->alpha, gd->beta, kappa);
    o->mue = mue;
    o->delta = chandra_delta(c.chandra, mue);
    o->g = g;
    o->flux = disk_nt_flux(c, r) * crm::cr_pow_4(g);
}

/* modes EQPLANE and POLARIZED */
S5_HD S5_INL void trace_eqplane_pixel(const S5ImageConsts& c, int ix, int iy, PixelOut* o)
{
    double alpha, beta;
    pixel_impact(c, ix, iy, &alpha, &beta);
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = 0.0;
    o->intensity = o->tau = o->qerr = 0.0; o->steps = 0;

    Geodesic gd;
    int error = 0;
    double K_mm = 0.0, icn_u = 0.0;
    gd.type = -1;
    if (!init_inf_cached(c, alpha, beta, &gd, &error, &K_mm, &icn_u)) {
        int gt = (error == GD_ERROR_TYPE_RR_DOUBLE) ? gtype_code(gd.type) : SIM5_GT_NONE;
        o->status = (unsigned)((SIM5_ST_INITERR + error) | (gt << 5));
        return;
    }
    unsigned gt = (unsigned)gtype_code(gd.type) << 5;
    for (int order = 0; order <= c.max_order; order++) {
        double P = crossing_cached(&gd, order, K_mm, icn_u);
        if (isnan(P)) {
            o->status = (order == 0 ? SIM5_ST_NOCROSS0 : order == 1 ? SIM5_ST_NOCROSS1 : SIM5_ST_NOCROSS2) | gt;
            return;
        }
        double r = geodesic_position_rad(&gd, P);
        if (r >= c.rmin_emit) {
            o->status = (order == 0 ? SIM5_ST_HIT0 : order == 1 ? SIM5_ST_HIT1 : SIM5_ST_HIT2) | gt;
            o->r = r;
            if (c.outputs & SIM5_OUT_PHI) o->phi = geodesic_position_azm(&gd, r, 0.0, P);
            if (c.mode == SIM5_MODE_POLARIZED) {
                polarized_hit(c, &gd, r, P, o);
            } else {
                double g = gfactorK(r, c.a, gd.l);
                double f = disk_nt_flux(c, r);
                o->g = g;
                o->flux = f * crm::cr_pow_4(g);
            }
            return;
        }
    }
    o->status = SIM5_ST_MISS | gt;
}

/* mode STEPWISE: raytrace() through the harness torus (SURVEY.md 8d cfg 4; oracle/ref_driver.c pixel_stepwise) */
struct StepRay {           /* live state of one stepwise ray (the persistent kernel keeps this per lane) */
    double x[4], k[4];
    RayData rtd;
    double I, tau;
    int steps;
    unsigned gt;
};
/* returns true if the ray is live (needs stepping); otherwise o->status is final */
S5_HD S5_INL bool stepwise_start(const S5ImageConsts& c, int ix, int iy, StepRay* s, PixelOut* o)
{
    double alpha, beta;
    pixel_impact(c, ix, iy, &alpha, &beta);
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = 0.0;
    o->intensity = o->tau = o->qerr = 0.0; o->steps = 0;
    Geodesic gd;
    int error = 0;
    gd.type = -1;
    if (!geodesic_init_inf_sc(c.incl, c.sin_i, c.cos_i, c.a, alpha, beta, &gd, &error)) {
        int gt = (error == GD_ERROR_TYPE_RR_DOUBLE) ? gtype_code(gd.type) : SIM5_GT_NONE;
        o->status = (unsigned)((SIM5_ST_INITERR + error) | (gt << 5));
        return false;
    }
    s->gt = (unsigned)gtype_code(gd.type) << 5;
    double r0 = c.r_start;
    if (!(r0 > gd.rp)) { o->status = SIM5_ST_NOSTART | s->gt; return false; }
    double P = geodesic_P_int(&gd, r0, 0);
    s->x[0] = 0.0;
    s->x[1] = r0;
    s->x[2] = geodesic_position_pol(&gd, P);
    s->x[3] = 0.0;
    geodesic_momentum(&gd, P, r0, s->x[2], s->k);
    if (isnan(P) || isnan(s->x[2]) || isnan(s->k[1]) || isnan(s->k[2])) { o->status = SIM5_ST_NOSTART | s->gt; return false; }
    raytrace_prepare(c.a, s->x, s->k, c.pf, 0, &s->rtd);
    s->I = 0.0; s->tau = 0.0; s->steps = 0;
    return true;
}
/* one raytrace() call + torus emission; returns 0 while the ray is live, else the termination class */
S5_HD S5_INL int stepwise_step(const S5ImageConsts& c, StepRay* s)
{
    double dl = c.step_max;
    raytrace(s->x, s->k, &dl, &s->rtd);
    s->steps++;
    {
        double r = s->x[1], m = s->x[2];
        double R = r * sqrt(1.0 - m * m);
        double z = r * m;
        double sv = sq((R - c.torus_rc) / c.torus_w) + sq(z / (c.torus_h * R));
        if (sv < 13.8) {
            Metric M;
            kerr_metric(c.a, r, m, &M);
            double Om = Omega_from_ell(c.torus_ell, &M);
            double den = M.g00 + 2. * Om * M.g03 + sq(Om) * M.g33;
            if (den < 0.0) {
                double U[4];
                fourvelocity_azimuthal(Om, &M, U);
                double g = (s->k[0] * M.g00 + s->k[3] * M.g03) / dotprod(s->k, U, &M);
                double rho = exp(-sv);
                double j = c.torus_j0 * rho * rho;
                double al = c.torus_k0 * rho;
                s->I += j * g * g * g * exp(-s->tau) * dl;
                s->tau += al * dl;
            }
        }
    }
    if (s->x[1] < c.rh_stop) return SIM5_ST_HORIZON;
    if (s->x[1] > c.rout_stop) return SIM5_ST_ESCAPE;
    if (s->rtd.error > 1e-2) return SIM5_ST_ERRBREAK;
    if (s->steps >= c.max_steps) return SIM5_ST_MAXSTEPS;
    return 0;
}
S5_HD S5_INL void stepwise_finish(const S5ImageConsts& c, StepRay* s, int cls, PixelOut* o)
{
    (void)c;
    o->intensity = s->I;
    o->tau = s->tau;
    o->steps = s->steps;
    o->qerr = raytrace_error(s->x, s->k, &s->rtd);
    o->status = (unsigned)cls | s->gt;
}

} /* namespace s5 */
#endif
