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
#include "ellfast.cuh"

namespace s5 {

struct PixelOut {
    double r, phi, g, flux, chi, delta, mue, intensity, tau, qerr, height, delay;
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
template <class OPS>
S5_HD S5_INL double disk_nt_flux_t(OPS& o, const S5ImageConsts& c, double r)
{
    if (r <= c.nt_rms) return 0.0;
    double x = o.sqrt(r);
    double f0 = x - c.nt_x0 - c.nt_k0 * cr_log(o.div(x, c.nt_x0));
    double f1 = c.nt_k1 * cr_log(o.div(x - c.nt_x1, c.nt_d1));
    double f2 = c.nt_k2 * cr_log(o.div(x - c.nt_x2, c.nt_d2));
    double f3 = c.nt_k3 * cr_log(o.div(x - c.nt_x3, c.nt_d3));
    double F = o.div(o.div(1., c.nt_4pi * r) * 1.5, x * x * (x * x * x - 3. * x + c.nt_2a)) * (f0 - f1 - f2 - f3);
    return o.div(9.1721376255e+28 * F * c.nt_mdot, c.nt_mass);
}
S5_HD S5_MID double disk_nt_flux(const S5ImageConsts& c, double r)
{
    ff::Quick f;
    double F = disk_nt_flux_t(f, c, r);
    if (!f.ok) { ff::Plain p; F = disk_nt_flux_t(p, c, r); }
    return F;
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

/* per-ray values that the reference recomputes several times (identical inputs -> identical bits) */
struct RayCache {
    double K_mm;        /* K(mm) = rf(0, 1-mm, 1): Tpp, every crossing order, complete Pi in the azimuth */
    double icn_u;       /* cn^-1(cos_i/sqrt(m2p), mm): Tip and every crossing order */
    double rf_u;        /* the R_F inside icn_u; equals the R_F of elliptic_pi_cos(cos_i/sqrt(m2p), ., mm) */
    double rf_u_z, rf_u_m;
    double isn_inf;     /* sn^-1 at r = infinity (RR rays): Rpc and both integral_R_rp_re_inf calls */
    double u_pol;       /* cos_i / sqrt(m2p): the argument of icn_u, of every crossing order and of the rf_u test (the reference forms it 3-4 times) */
    bool have_rf_u;
};
S5_HD S5_INL double polar_amplitude(const Geodesic* g)
{
    ff::Quick o;
    double u = o.div(g->cos_i, o.sqrt(g->m2p));
    if (!o.ok) u = g->cos_i / sqrt(g->m2p);
    return u;
}

/* geodesic_init_inf with the T-integrals' Carlson values kept for the crossings and the azimuth.
 * Same results as geodesic_init_inf_sc (geod.cuh). */
S5_HD S5_INL int init_inf_cached(const S5ImageConsts& c, double alpha, double beta, Geodesic* g, int* error, RayCache* k)
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
    k->isn_inf = 0.0;
    if (!geodesic_R_roots(g, 1.7976931348623157e308, error, &k->isn_inf)) return 0;
    if (!geodesic_T_roots(g, g->cos_i, error)) return 0;
    /* theta_int(0) == mK*K(mm): jacobi_icn(0/sqrt(m2p), mm) takes its z == 0 exit (0 <= mm < 1 here) */
    k->K_mm = elliptic_k(g->mm);
    k->u_pol = polar_amplitude(g);
    k->icn_u = jacobi_icn_ex(k->u_pol, g->mm, &k->rf_u, &k->rf_u_z, &k->rf_u_m, &k->have_rf_u);
    g->Tpp = 2. * (g->mK * k->K_mm);
    g->Tip = g->mK * k->icn_u;
    *error = GD_OK;
    return 1;
}
/* geodesic_find_midplane_crossing on the cached values.  sim5kerr-geod.c:845-885 */
S5_HD S5_INL double crossing_cached(const Geodesic* g, int order, const RayCache& k)
{
    if (g->q <= 0.0) return NAN;
    double u = k.u_pol;
    double u0 = u;
    if (!ensure_range(&u, -1.0, +1.0, 1e-4)) return NAN;
    double icn = (u == u0) ? k.icn_u : jacobi_icn(u, g->mm);
    double pos;
    if (g->beta > 0.0)      pos = g->mK * ((2. * (double)order + 1.) * k.K_mm + icn);
    else if (g->beta < 0.0) pos = g->mK * ((2. * (double)order + 1.) * k.K_mm - icn);
    else                    pos = g->mK * ((2. * (double)order + 1.) * k.K_mm);
    if (pos > 2. * g->Rpc) pos = NAN;
    return pos;
}

/* Everything the azimuth of an equatorial disk hit needs from its geodesic (17 doubles + flags).  The image
 * kernels run the azimuth as a SECOND PHASE over a queue of these items (kernels.cuh): phase A (roots, crossing,
 * radius, g, flux) and phase B (3 rf + 6 rj + 2 sncndn) each fit the instruction cache and run at higher
 * occupancy than one fused kernel, and phase B is free of RR/RC divergence. */
struct AzIn {
    double e0, e1, e2, e3;      /* RR: r1..r4 (real) ; RC: r1, r2, Re r3, Im r3 */
    double l, m2m, m2p, mm;
    double K_mm, rf_u, isn_inf;
    double r;
    double a, cos_i;            /* per-image: the clamped spin g->a and cos(i) */
    int type, nrr;
    bool rf_ok;                 /* rf_u is the R_F of elliptic_pi_cos(cos_i/sqrt(m2p), ., m2p/(m2m+m2p)) */
    /* the three decisions geodesic_position_azm takes from P, Rpc, Tpp, Tip and beta (sim5kerr-geod.c:476, 528-546), taken where those
     * numbers are (phase A) and queued as three bits of the key instead of five doubles */
    bool ppc;                   /* (nrr > 0) && (P > Rpc): the hit lies behind the radial turning point */
    bool beta_nonneg;           /* beta >= 0 */
    bool turn;                  /* P >= T + Tpp with T = -(Tpp - Tip) (beta >= 0) or -Tip: the hit lies behind the first polar turning point */
};
/* queue item: e0 e1 e2 e3 l m2m m2p r | K_mm rf_u isn_inf -- the tolerance-mode kernel reads the first 8, the bit-faithful one all 11;
 * mm = m2p / (m2m + m2p) is re-formed by the reader (the same quotient of the same operands as geodesic_T_roots forms for q > 0, and only
 * q > 0 rays cross the equatorial plane) */
#define S5_AZ_NFIELDS 11
#define S5_AZ_NFAST 8
S5_HD S5_INL void az_decide(const Geodesic* g, double P, AzIn* z)
{
    z->ppc = (g->nrr > 0) && (P > g->Rpc);
    z->beta_nonneg = (g->beta >= 0.0);
    double T = z->beta_nonneg ? -(g->Tpp - g->Tip) : -g->Tip;
    z->turn = (P >= T + g->Tpp);
}

/* the part of the item that is NOT in the geodesic struct (the lockstep kernels keep the geodesic in a shared-memory slot and copy its
 * fields into the queue from there: the item never exists as a 160-byte local copy) */
S5_HD S5_INL void az_make_tail(const Geodesic* g, const RayCache& k, double r, double P, AzIn* z)
{
    z->K_mm = k.K_mm; z->rf_u = k.rf_u; z->isn_inf = k.isn_inf;
    z->r = r;
    az_decide(g, P, z);
    z->type = g->type; z->nrr = g->nrr;
    /* m2p / (m2m + m2p) is g->mm for q > 0 (geodesic_T_roots: the same quotient of the same operands); q < 0 rays never hit the disk */
    double tm = (g->q > 0.0) ? g->mm : g->m2p / (g->m2m + g->m2p);
    z->rf_ok = k.have_rf_u && (k.rf_u_z == k.u_pol) && (k.rf_u_m == tm);
}
S5_HD S5_INL void az_make(const Geodesic* g, const RayCache& k, double r, double P, AzIn* z)
{
    if (g->type == GEOD_TYPE_RR) { z->e0 = g->r1.re; z->e1 = g->r2.re; z->e2 = g->r3.re; z->e3 = g->r4.re; }
    else                          { z->e0 = g->r1.re; z->e1 = g->r2.re; z->e2 = g->r3.re; z->e3 = g->r3.im; }
    z->l = g->l; z->m2m = g->m2m; z->m2p = g->m2p; z->mm = g->mm;
    az_decide(g, P, z);
    z->K_mm = k.K_mm; z->rf_u = k.rf_u; z->isn_inf = k.isn_inf;
    z->r = r; z->a = g->a; z->cos_i = g->cos_i;
    z->type = g->type; z->nrr = g->nrr;
    double tm = g->m2p / (g->m2m + g->m2p);
    z->rf_ok = k.have_rf_u && (k.rf_u_z == g->cos_i / sqrt(g->m2p)) && (k.rf_u_m == tm);
}

/* geodesic_position_azm(g, r, m = 0, P) for an equatorial hit, without the reference's repeated Carlson calls
 * (sim5kerr-geod.c:462-555 -> sim5elliptic.c:1017-1044, 676-690, 425-450, 1142-1159):
 *   - both poles (r+, r-) share sn^-1, sn/cn/dn and the R_F term of Pi at each of the two limits (infinity, r);
 *   - sn^-1 at infinity is the value already behind Rpc; integral_Z1(u=0) is a signed zero;
 *   - K(mm) and the R_F behind Tip serve the polar integrals; phi_mp(m=0) is the same call as phi_pp/2.
 * 3 rf + 6 rj + 2 sncndn instead of ~14 rf + 7 rj + 4 sncndn (reference: ~33 rf incl. its argument checks), same bits. */
#if defined(__CUDA_ARCH__)
#define S5_STAGE_SYNC() do { if (SYNC) __syncthreads(); } while (0)
#else
#define S5_STAGE_SYNC() do { } while (0)
#endif
/* SYNC: the caller is a CTA whose threads ALL run this routine on items of ONE geodesic type; the barriers between
 * the stages keep its warps inside the same routine (the same instruction-cache lines) at the same time */
template <bool SYNC>
S5_HD S5_MID double azimuth_from_t(const AzIn& z)
{
    double phi = 0.0;
    int ppc = z.ppc;
    double a2 = sq(z.a);
    double rp = 1. + sqrt(1. - a2);
    double rm = 1. - sqrt(1. - a2);
    double A, B;
    const double r = z.r;

    if (z.type == GEOD_TYPE_RR) {
        double a = z.e0, b = z.e1, c = z.e2, d = z.e3;
        double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
        double pre = -2.0 / sqrt((a - c) * (b - d));
        double aa2 = (a - d) / (b - d);
        double sn_, dn_, cn_inf, cn_r;
        double u_inf = z.isn_inf;
        jacobi_sncndn(u_inf, m2, &sn_, &cn_inf, &dn_);
        double u_r = jacobi_isn(sqrt(((b - d) * (r - a)) / ((a - d) * (r - b))), m2);
        S5_STAGE_SYNC();
        jacobi_sncndn(u_r, m2, &sn_, &cn_r, &dn_);
        double c2p = ((rp - b) * (a - d)) / ((rp - a) * (b - d));
        double c2m = ((rm - b) * (a - d)) / ((rm - a) * (b - d));
        double Pinf_p, Pinf_m, Pr_p, Pr_m;
        S5_STAGE_SYNC();
        pi_cos_pair(cn_inf, m2, c2p, c2m, &Pinf_p, &Pinf_m);
        S5_STAGE_SYNC();
        pi_cos_pair(cn_r, m2, c2p, c2m, &Pr_p, &Pr_m);
        S5_STAGE_SYNC();
        {
            double p = rp;
            double c2 = c2p;
            double z0 = integral_Z1_at0(c2, aa2);
            double Rinf = pre / (p - a) * (1. / c2 * ((c2 - aa2) * Pinf_p + aa2 * u_inf) - z0);
            double Rr   = pre / (p - a) * (1. / c2 * ((c2 - aa2) * Pr_p + aa2 * u_r) - z0);
            A = Rinf + (ppc ? +1 : -1) * Rr;
        }
        {
            double p = rm;
            double c2 = c2m;
            double z0 = integral_Z1_at0(c2, aa2);
            double Rinf = pre / (p - a) * (1. / c2 * ((c2 - aa2) * Pinf_m + aa2 * u_inf) - z0);
            double Rr   = pre / (p - a) * (1. / c2 * ((c2 - aa2) * Pr_m + aa2 * u_r) - z0);
            B = Rinf + (ppc ? +1 : -1) * Rr;
        }
        phi += 1. / sqrt(1. - a2) * (A * (z.a * rp - z.l * a2 / 2.) - B * (z.a * rm - z.l * a2 / 2.));
    } else if (z.type == GEOD_TYPE_RC) {
        A = integral_R_rp_cc2_inf(z.e0, z.e1, z.e2, z.e3, rp, r);
        B = integral_R_rp_cc2_inf(z.e0, z.e1, z.e2, z.e3, rm, r);
        phi += 1. / sqrt(1. - a2) * (A * (z.a * rp - z.l * a2 / 2.) - B * (z.a * rm - z.l * a2 / 2.));
    } else {
        return NAN;
    }

    /* polar part: integral_T_mp(m2m, m2p, 1.0, X) for X = 0 (twice) and X = cos_i */
    double tm = z.m2p / (z.m2m + z.m2p);
    double tn = z.m2p / (z.m2p - 1.0);
    double tpre = 1. / sqrt(z.m2m + z.m2p) / (1.0 - z.m2p);
    double comp;                                     /* elliptic_pi_cos(0, n, m) -> elliptic_pi_complete(n, m) */
    if (isinf(tn)) {
        comp = 0.0;
    } else {
        double mc = (tm == 1.0) ? 0.99999999 : tm;
        double nc = (tn == 1.0) ? 0.99999999 : tn;
        double q = 1.0 - mc;
        double rfK = (tm == z.mm && tm != 1.0) ? z.K_mm : rf(0.0, q, 1.0);
        comp = rfK + nc * rj(0.0, q, 1.0, 1.0 - nc) / 3.0;
    }
    S5_STAGE_SYNC();
    double T0 = tpre * comp;
    double phi_pp = 2.0 * z.l / z.a * T0;
    double phi_mp = z.l / z.a * T0;
    double phi_ip;
    if (z.cos_i >= 0.0) {
        double cu = z.cos_i / sqrt(z.m2p);
        PiShare sh = z.rf_ok ? pi_share_with(cu, tm, z.rf_u) : pi_share(cu, tm);
        double v = (cu == 0.0 && !isinf(tn)) ? comp : pi_cos_shared(sh, tn);
        phi_ip = z.l / z.a * (tpre * v);
    } else {
        phi_ip = z.l / z.a * integral_T_mp(z.m2m, z.m2p, 1.0, z.cos_i);
    }

    double sign_dm = z.beta_nonneg ? +1.0 : -1.0;
    if (sign_dm > 0.0) {
        phi -= phi_pp - phi_ip;
    } else {
        phi -= phi_ip;
    }
    if (z.turn) {
        phi += phi_pp;
        sign_dm = -sign_dm;
    }
    phi += (sign_dm < 0) ? phi_mp : phi_pp - phi_mp;
    return phi;
}
S5_HD S5_INL double azimuth_from(const AzIn& z) { return azimuth_from_t<false>(z); }
S5_HD S5_INL double azimuth_equatorial(const Geodesic* g, const RayCache& k, double r, double P)
{
    AzIn z;
    az_make(g, k, r, P, &z);
    return azimuth_from(z);
}
S5_HD S5_MID double azimuth_fast_rr(const AzIn& z, bool* ok);
S5_HD S5_MID double azimuth_fast_rc(const AzIn& z, bool* ok);
/* the azimuth as the image kernels compute it by default: tolerance mode for RR hits, bit-faithful for everything else */
S5_HD S5_INL double azimuth_equatorial_default(const Geodesic* g, const RayCache& k, double r, double P)
{
    AzIn z;
    az_make(g, k, r, P, &z);
    if (z.type == GEOD_TYPE_RR || z.type == GEOD_TYPE_RC) {
        bool ok;
        double v = (z.type == GEOD_TYPE_RR) ? azimuth_fast_rr(z, &ok) : azimuth_fast_rc(z, &ok);
        if (ok) return v;
    }
    return azimuth_from(z);
}

/* Cauchy principal value of R_J for p < 0 on top of the shared sequence: R_J(x,y,z,p) = a (b R_J(x,y,z,pt) + 3 (R_C(rho,tau) - R_F(x,y,z)))
 * with x <= y <= z (sim5elliptic.c:166-177, 204).  pv_prepare turns p into the positive pt the sequence is run with. */
struct PvTerm { double a, b, rcx; bool neg; };
S5_HD S5_INL double pv_prepare(double x, double y, double z, double p, PvTerm* t)
{
    t->neg = !(p > 0.0);
    if (!t->neg) { t->a = t->b = t->rcx = 0.0; return p; }
    t->a = ff::rcp_ap(y - p);
    t->b = t->a * (z - y) * (y - x);
    double pt = y + t->b;
    double ry = ff::rcp_ap(y);
    double rho = x * z * ry;
    double tau = p * pt * ry;                       /* < 0: R_C(rho, tau) = sqrt(rho / (rho - tau)) R_C(rho - tau, -tau) */
    double xs = rho - tau;
    t->rcx = (rho > 0.0) ? ff::sqrt_ap(rho * ff::rcp_ap(xs)) * rc_hi(xs, -tau) : 0.0;
    return pt;
}
S5_HD S5_INL double pv_finish(double J, double F, const PvTerm& t) { return t.neg ? t.a * (t.b * J + 3.0 * (t.rcx - F)) : J; }

/* polar part of the azimuth in tolerance mode: integral_T_mp(m2m, m2p, 1, X) for X = 0 and X = cos_i (sim5elliptic.c:1142-1159)
 * and the turning-point bookkeeping of geodesic_position_azm (sim5kerr-geod.c:528-553) for an equatorial hit */
/* Parity is measured against the REFERENCE's doubles, and the reference evaluates phi as a sum of separately rounded terms
 * that can be much larger than phi (1/c2 ((c2-a2) Pi + a2 u) with |c2| << a2 when a root of R(r) sits next to r+ or r-;
 * A - B; the polar turning-point terms).  Its own rounding noise is then ~1e-16 * mag, which a different (even a more
 * accurate) evaluation cannot reproduce.  Where that noise could approach the 1e-9 bar -- mag > S5_AZ_COND_LIMIT max(|phi|,1),
 * a few pixels per million -- the item is handed to the bit-faithful kernel instead. */
#ifndef S5_AZ_COND_LIMIT
#define S5_AZ_COND_LIMIT 3.0e4
#endif
#if defined(S5_AZ_DIAG)
#undef S5_AZ_COND_LIMIT
#define S5_AZ_COND_LIMIT 1.0e300      /* calibration build: never fall back, so the unguarded deviation is visible */
#endif
#if defined(S5_AZ_DIAG) && !defined(__CUDA_ARCH__)
static thread_local double s5_last_kappa;       /* calibration builds (tools/calibrate_azimuth_guard.py): the indicator of the last item */
#endif
/* The reference passes the amplitude of each incomplete integral through 1 - cn^2 and 1 - sn^2 conversions (elliptic_f_cos,
 * elliptic_pi_cos, jacobi_isn: sim5elliptic.c:262-270, 436-448, 485) and through cn(sn^-1(.)) round trips, which put an ABSOLUTE
 * noise of ~1e-16 on both cn^2 and sn^2.  Its F then carries a noise of ~1e-16 / (2 cn sn dn) and its Pi(n) one of
 * ~1e-16 / (2 cn sn dn |1 - n sn^2|) -- large next to the turning point (sn -> 0), at cn -> 0 and at the pole of Pi.
 * amp_noise returns these in units of 1e-16 with a safety factor of 8 (p = 1 for F). */
S5_HD S5_INL float amp_noise(float csd, float p)        /* csd = cn * sn * dn */
{
    return 4.0f * ff::rcpf_ap(fabsf(csd * p));                         /* inf / NaN for degenerate amplitudes: the guard then fails, as it should */
}
/* the guard is bookkeeping, not a result: it runs in FP32 (the FP32 and SFU pipes idle next to the FP64 work) */
S5_HD S5_INL float gf(double x) { return fabsf((float)x); }

S5_HD S5_INL bool azimuth_well_conditioned(float mag, double phi)
{
#if defined(S5_AZ_DIAG) && !defined(__CUDA_ARCH__)
    s5_last_kappa = (double)mag / fmax(fabs(phi), 1.0);
#endif
    return mag <= (float)S5_AZ_COND_LIMIT * fmaxf(gf(phi), 1.0f);       /* false for NaN, for inf */
}

S5_HD S5_MID double azimuth_fast_polar(const AzIn& z, bool* ok, float* mag)
{
    bool good = true;
    double phi = 0.0;
    double msum = z.m2m + z.m2p;
    double tm = ff::div_ap(z.m2p, msum);
    double tn = ff::div_ap(z.m2p, z.m2p - 1.0);
    double tpre = -ff::rsqrt_nc(msum) * ff::rcp_ap(z.m2p - 1.0);
    double qc = 1.0 - tm, pc = 1.0 - tn;
    double cu = z.cos_i * ff::rsqrt_nc(z.m2p);
    double cu2 = cu * cu;
    double ns2 = -tn * (1.0 - cu2);
    double qu = 1.0 - (1.0 - cu2) * tm;
    double pu = 1.0 + ns2;
    bool gp = (z.cos_i > 0.0) && (cu2 < 1.0) && (tm < 1.0) && !(tn == 1.0) && hi_domain(0.0, qc, 1.0) && hi_domain_p(pc) && hi_domain(cu2, qu, 1.0) && hi_domain_p(pu);
    if (!gp) { qc = pc = cu2 = qu = pu = 1.0; good = false; }
#if defined(S5_POLAR_DUPLICATION)
    double rfK = (tm == z.mm) ? z.K_mm : rf_hi(0.0, qc, 1.0);
    double comp = rfK + tn * rj_hi(0.0, qc, 1.0, pc) * HK(THIRD);
#else
    double comp = cel_pi_hi(qc, pc);                 /* complete Pi(tn | tm): AGM instead of a duplication sequence */
#endif
    double Fu, Ju;
    rfj_hi<1, true>(cu2, qu, 1.0, &pu, &Fu, &Ju);
    double vu = ff::sqrt_ap0(1.0 - cu2) * (Fu - ns2 * Ju * HK(THIRD));
    double la = ff::div_ap(z.l, z.a);
    double T0 = tpre * comp;
    double phi_pp = 2.0 * la * T0;
    double phi_mp = la * T0;
    double phi_ip = la * (tpre * vu);

    double sign_dm = z.beta_nonneg ? +1.0 : -1.0;
    if (sign_dm > 0.0) {
        phi -= phi_pp - phi_ip;
    } else {
        phi -= phi_ip;
    }
    if (z.turn) {
        phi += phi_pp;
        sign_dm = -sign_dm;
    }
    phi += (sign_dm < 0) ? phi_mp : phi_pp - phi_mp;
    *mag = 2.0f * gf(phi_pp) + gf(phi_ip) + gf(phi_mp);
    *ok = good;
    return phi;
}

/*
 * Tolerance-mode azimuth of an RR disk hit (phase B, the dominant kernel): the same integral as azimuth_from_t, to
 * ~1e-14 instead of bit for bit (bar: 1e-9, BASELINE.json north_star), with less work:
 *   - the amplitudes are algebraic: sn^2 at infinity is (b-d)/(a-d) and cn^2 = (a-b)/(a-d); at r,
 *     sn^2 = (b-d)(r-a)/((a-d)(r-b)) and cn^2 = (a-b)(r-d)/((a-d)(r-b)) (no cancellation).  The reference gets them as
 *     cn(sn^-1(sn)) through two jacobi_sncndn calls (sim5elliptic.c:676-690 via :1026,1041); here neither is needed;
 *   - sn^-1(sn, m) = sn * R_F(cn^2, 1 - m sn^2, 1) is the R_F that the Pi of the same amplitude needs anyway, so each
 *     limit is ONE duplication sequence giving R_F and the R_J of both poles (rfj_hi<2, true>);
 *   - the 7th-order Carlson series of ellfast.cuh (about 4 duplication steps instead of 6.4).
 * 4 duplication sequences and 6 R_J per hit (2 x {R_F, R_J, R_J}, the complete R_J and one R_J + R_F of the polar part)
 * instead of 3 R_F + 6 R_J + 2 sncndn over 9 sequences.  *ok = false (and a NaN result) when an argument leaves the
 * domain of the fast routines or the reference would take one of its special branches; the caller then has the
 * bit-faithful kernel redo the item.
 */
/* one limit (amplitude sn^2 = s2, cn^2 = c2) of the RR radial integrals for both poles: adds sgn * ((n_k - aa2) Pi(n_k) + aa2 u)
 * to br[k] and the matching magnitudes (terms + the reference's amplitude noise) to gm[k]; consumed at once so that nothing
 * but the two brackets stays live across the second duplication sequence */
S5_HD S5_INL bool rr_limit(double s2, double c2, double m2, double aa2, double c2p, double c2m, double sgn, double* br, float* gm)
{
    bool good = true;
    double q = 1.0 - s2 * m2;
    double pp[2] = {1.0 - c2p * s2, 1.0 - c2m * s2};
    const float p0f[2] = {(float)pp[0], (float)pp[1]};
    bool g = (s2 >= 0.0) && hi_domain(c2, q, 1.0) && hi_domain_p(fabs(pp[0])) && hi_domain_p(fabs(pp[1]));
    if (!g) { c2 = q = pp[0] = pp[1] = 1.0; good = false; }
    PvTerm t0, t1;                                     /* p < 0 (principal value): rays whose turning point lies inside r- */
    pp[0] = pv_prepare(c2, q, 1.0, pp[0], &t0);
    pp[1] = pv_prepare(c2, q, 1.0, pp[1], &t1);
    if (!(hi_domain_p(pp[0]) && hi_domain_p(pp[1]))) { pp[0] = pp[1] = 1.0; good = false; }
    double F, J[2];
    rfj_hi<2, true>(c2, q, 1.0, pp, &F, J);
    double sn = ff::sqrt_ap0(s2);
    double u = sn * F;
    double Pp = sn * (F + c2p * s2 * pv_finish(J[0], F, t0) * HK(THIRD));
    double Pm = sn * (F + c2m * s2 * pv_finish(J[1], F, t1) * HK(THIRD));
    br[0] += sgn * ((c2p - aa2) * Pp + aa2 * u);
    br[1] += sgn * ((c2m - aa2) * Pm + aa2 * u);
    float csd = ff::sqrtf_ap(gf(c2) * gf(s2) * gf(q));
    float gu = gf(aa2) * (gf(u) + amp_noise(csd, 1.0f));
    gm[0] += gf(c2p - aa2) * (gf(Pp) + amp_noise(csd, p0f[0])) + gu;
    gm[1] += gf(c2m - aa2) * (gf(Pm) + amp_noise(csd, p0f[1])) + gu;
    return good;
}

S5_HD S5_MID double azimuth_fast_rr(const AzIn& z, bool* ok)
{
    const double r = z.r;
    int ppc = z.ppc;
    double a2 = sq(z.a);
    double sq1 = ff::sqrt_ap(1. - a2);
    double rp = 1. + sq1, rm = 1. - sq1;
    double a = z.e0, b = z.e1, c = z.e2, d = z.e3;
    double ad = a - d, bd = b - d, ab = a - b, ac = a - c;
    double rbd = ff::rcp_ap(bd), rad = ff::rcp_ap(ad);
    double pre = -2.0 * ff::rsqrt_nc(ac * bd);
    double m2 = ((b - c) * ad) * (0.25 * pre * pre);          /* 1 / (ac bd) = pre^2 / 4 */
    double aa2 = ad * rbd;
    double c2p = ff::div_ap((rp - b) * aa2, rp - a);
    double c2m = ff::div_ap((rm - b) * aa2, rm - a);
    double br[2] = {0.0, 0.0};
    float gm[2] = {0.0f, 0.0f};
    /* limit at infinity: sn^2 = (b-d)/(a-d); limit at r: sn^2 = (b-d)(r-a)/((a-d)(r-b)) */
    bool good = rr_limit(bd * rad, ab * rad, m2, aa2, c2p, c2m, 1.0, br, gm);
    double rrb = rad * ff::rcp_ap(r - b);
    good = rr_limit((bd * (r - a)) * rrb, (ab * (r - d)) * rrb, m2, aa2, c2p, c2m, ppc ? +1.0 : -1.0, br, gm) && good;

    double rsq1 = ff::rcp_ap(sq1);
    double wA = (z.a * rp - z.l * a2 * 0.5) * rsq1, wB = (z.a * rm - z.l * a2 * 0.5) * rsq1;
    double oA = ff::div_ap(pre, (rp - a) * c2p), oB = ff::div_ap(pre, (rm - a) * c2m);
    double phi = (oA * br[0]) * wA - (oB * br[1]) * wB;
    /* size of the rounded terms the reference adds up, plus the noise its amplitudes carry (azimuth_well_conditioned, amp_noise) */
    float mag = gf(wA * oA) * gm[0] + gf(wB * oB) * gm[1];

    bool okp;
    float pmag;
    phi += azimuth_fast_polar(z, &okp, &pmag);
    good = good && okp && azimuth_well_conditioned(mag + pmag, phi);
    *ok = good;
    return good ? phi : NAN;
}

/*
 * Tolerance-mode azimuth of an RC disk hit (two real roots a > b and the pair u +- iv): B&F 260.04 and 341.03 as the
 * reference combines them (sim5elliptic.c:1081-1112, 755-792), with the same economies as azimuth_fast_rr: the amplitudes
 * cn at r and at infinity are the algebraic arguments of elliptic_f_cos themselves, F and the Pi of both poles share one
 * duplication sequence per limit, the complete integrals are evaluated only when the two limits lie on different sides
 * of cn = 0, and the complex f1 term of integral_R1 is spelled as the real atan / log it reduces to.
 */
S5_HD S5_MID double azimuth_fast_rc(const AzIn& z, bool* ok)
{
    const double r = z.r;
    double a2s = sq(z.a);
    double sq1 = ff::sqrt_ap(1. - a2s);
    double rpm[2] = {1. + sq1, 1. - sq1};
    double a = z.e0, b = z.e1, u = z.e2, v2 = sq(z.e3);
    double A_ = ff::sqrt_ap(sq(a - u) + v2), B_ = ff::sqrt_ap(sq(b - u) + v2);
    double g = ff::rsqrt_nc(A_ * B_);
    double m = (sq(A_ + B_) - sq(a - b)) * (0.25 * g * g);
    double alpha2 = ff::div_ap(B_ + A_, B_ - A_);
    double cl[2] = {ff::div_ap(r * (A_ - B_) + a * B_ - b * A_, r * (A_ + B_) - a * B_ - b * A_), ff::div_ap(A_ - B_, A_ + B_)};   /* cn at r, at infinity */
    bool good = (m > 0.0) && (m < 1.0) && (cl[0] != 0.0) && (cl[1] != 0.0);
    double alpha1[2], nn[2], mma[2];
    #pragma unroll
    for (int k = 0; k < 2; k++) {
        double p = rpm[k];
        alpha1[k] = ff::div_ap(B_ * a + b * A_ - p * A_ - p * B_, B_ * a - b * A_ + p * A_ - p * B_);
        double al2 = sq(alpha1[k]);
        double ral = ff::rcp_ap(al2 - 1.);
        nn[k] = al2 * ral;
        mma[k] = -(m + (1. - m) * al2) * ral;
        good = good && (nn[k] == nn[k]) && (fabs(nn[k]) < 1e18) && (nn[k] != 1.0);
    }
    /* per limit: F(|c|) and the two Pi(|c|, n_k), f1 */
    double Fh[2], Ph[2][2], f1[2][2];
    float nP[2][2], nF[2];
    #pragma unroll
    for (int j = 0; j < 2; j++) {
        double c = fabs(cl[j]);
        double c2 = c * c, s2 = 1.0 - c2;
        double q = 1.0 - s2 * m;
        double pp[2] = {1.0 - nn[0] * s2, 1.0 - nn[1] * s2};
        bool gj = (c2 < 1.0) && hi_domain(c2, q, 1.0) && (pp[0] != 0.0) && (pp[1] != 0.0) && hi_domain_p(fabs(pp[0])) && hi_domain_p(fabs(pp[1]));
        if (!gj) { c2 = q = 1.0; s2 = 0.0; pp[0] = pp[1] = 1.0; good = false; }
        PvTerm t0, t1;
        double pt[2] = {pv_prepare(c2, q, 1.0, pp[0], &t0), pv_prepare(c2, q, 1.0, pp[1], &t1)};
        if (!(hi_domain_p(pt[0]) && hi_domain_p(pt[1]))) { pt[0] = pt[1] = 1.0; good = false; }
        double F, J[2];
        rfj_hi<2, true>(c2, q, 1.0, pt, &F, J);
        double s = ff::sqrt_ap0(s2), dn = ff::sqrt_ap(q);
        float csd = ff::sqrtf_ap(gf(c2) * gf(s2) * gf(q));     /* the reference's amplitude noise, as in azimuth_fast_rr */
        nF[j] = amp_noise(csd, 1.0f);
        nP[j][0] = amp_noise(csd, (float)pp[0]);
        nP[j][1] = amp_noise(csd, (float)pp[1]);
        Fh[j] = s * F;
        Ph[j][0] = s * (F + nn[0] * s2 * pv_finish(J[0], F, t0) * HK(THIRD));
        Ph[j][1] = s * (F + nn[1] * s2 * pv_finish(J[1], F, t1) * HK(THIRD));
        #pragma unroll
        for (int k = 0; k < 2; k++) {
            double am = fabs(mma[k]);
            if (am > 1e-5) {
                double rsm = ff::rsqrt_nc(am);
                double y = am * rsm * s * ff::rcp_ap(dn);
                f1[j][k] = (mma[k] > 0.0) ? atan(y) * rsm : -0.5 * log(fabs(ff::div_ap(1.0 + y, 1.0 - y))) * rsm;
            } else {
                f1[j][k] = ff::div_ap(s, dn);
            }
        }
    }
    /* complete integrals, needed when the limits straddle cn = 0 */
    bool straddle = (cl[0] >= 0.0) != (cl[1] >= 0.0);
    double Kc = 0.0, Pc[2] = {0.0, 0.0};
    if (straddle) {
        double qc = 1.0 - m;
        double pp[2] = {1.0 - nn[0], 1.0 - nn[1]};
        bool gc = hi_domain(0.0, qc, 1.0) && (pp[0] != 0.0) && (pp[1] != 0.0) && hi_domain_p(fabs(pp[0])) && hi_domain_p(fabs(pp[1]));
        if (!gc) { qc = 1.0; pp[0] = pp[1] = 1.0; good = false; }
        PvTerm t0, t1;
        double pt[2] = {pv_prepare(0.0, qc, 1.0, pp[0], &t0), pv_prepare(0.0, qc, 1.0, pp[1], &t1)};
        if (!(hi_domain_p(pt[0]) && hi_domain_p(pt[1]))) { pt[0] = pt[1] = 1.0; good = false; }
        double J[2];
        rfj_hi<2, true>(0.0, qc, 1.0, pt, &Kc, J);
        Pc[0] = Kc + nn[0] * pv_finish(J[0], Kc, t0) * HK(THIRD);
        Pc[1] = Kc + nn[1] * pv_finish(J[1], Kc, t1) * HK(THIRD);
    }
    /* u(c) = F_cos(c): F(|c|) or 2K - F(|c|); Pi_cos alike.  index 0 = at r (u1), 1 = at infinity (u2) */
    double uu[2], AB[2];
    float MG[2];
    #pragma unroll
    for (int j = 0; j < 2; j++) uu[j] = (cl[j] >= 0.0) ? Fh[j] : 2.0 * Kc - Fh[j];
    /* when both limits are on the negative side the 2K (2 Pi_c) terms cancel in the differences: Kc = Pc = 0 above is then exact */
    #pragma unroll
    for (int k = 0; k < 2; k++) {
        double p = rpm[k];
        double R1v[2];
        float R1m = 0.0f;
        double r1a = ff::rcp_ap(1. - sq(alpha1[k]));
        #pragma unroll
        for (int j = 0; j < 2; j++) {
            double Pi = (cl[j] > 0.0) ? Ph[j][k] : ((cl[j] == 0.0) ? Pc[k] : 2.0 * Pc[k] - Ph[j][k]);
            R1v[j] = r1a * (Pi + alpha1[k] * f1[j][k]);
            /* the reference forms Pi(c<0) = 2 Pi_c - Pi(|c|) and F(c<0) = 2K - F(|c|) from separately rounded pieces */
            R1m += (gf(Ph[j][k]) + nP[j][k] + 2.0f * gf(Pc[k]) + gf(alpha1[k] * f1[j][k])) * gf(r1a);
        }
        double t0 = alpha2 * (uu[1] - uu[0]);
        double t1 = (alpha1[k] - alpha2) * (R1v[1] - R1v[0]);
        double outer = ff::div_ap((B_ - A_) * g, B_ * a + b * A_ - p * A_ - p * B_);
        AB[k] = outer * (t0 + t1);
        MG[k] = gf(outer) * (gf(alpha2) * (gf(Fh[0]) + nF[0] + gf(Fh[1]) + nF[1] + 4.0f * gf(Kc)) + gf(alpha1[k] - alpha2) * R1m);
    }
    double rsq1 = ff::rcp_ap(sq1);
    double phi = rsq1 * (AB[0] * (z.a * rpm[0] - z.l * a2s * 0.5) - AB[1] * (z.a * rpm[1] - z.l * a2s * 0.5));
    float mag = (MG[0] * gf(z.a * rpm[0] - z.l * a2s * 0.5) + MG[1] * gf(z.a * rpm[1] - z.l * a2s * 0.5)) * gf(rsq1);
    bool okp;
    float pmag;
    phi += azimuth_fast_polar(z, &okp, &pmag);
    *ok = good && okp && azimuth_well_conditioned(mag + pmag, phi);
    return *ok ? phi : NAN;
}

/* emission-side quantities of the polarized mode for a disk hit at (r, m=0), position parameter P */
S5_HD S5_MID void polarized_hit(const S5ImageConsts& c, const Geodesic* gd, double r, double P, PixelOut* o)
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
/* DEFER: when the azimuth is requested, a hit fills *defer (returns true) instead of computing phi here
 * (the instantiation then carries no azimuth code at all) */
/* DELAY: the instantiation also evaluates geodesic_timedelay between the hit and the sphere r = delay_r_ref (SIM5_OUT_DELAY,
 * SURVEY 8f N2); the default instantiation carries none of that code */
/* SYNC: EVERY thread of the CTA calls this routine (threads without a pixel trace a clamped one and drop the result), and the
 * routine passes CTA barriers between its stages (roots | polar integrals | order-0 crossing and radius | emission).  The routine
 * is ~190 KB of SASS against a 32 KB L1.5 instruction cache, so it runs from L2; the barriers keep all warps of the CTA inside
 * the same stage, i.e. on the same instruction lines, and every line is fetched once per batch instead of once per warp.  The
 * retry of higher crossing orders (a few rays per thousand) runs without barriers.  Same calls in the same order: same bits. */
/* SGD: the geodesic lives in the caller's slot *gslot (a per-thread slice of shared memory, S5_GD_SLOT_BYTES apart: the first
 * 200 bytes of the struct are all the routine touches) instead of in registers / local memory: the ~20 doubles of it that stay
 * live from the roots to the emission stage then cost a conflict-free ld.shared instead of a register each or a spill */
#define S5_GD_SLOT_BYTES 200      /* offsetof(Geodesic, k): 50 words per thread = conflict-free 64-bit accesses per half-warp */
template <bool DEFER, bool DELAY = false, bool SYNC = false, bool SGD = false>
S5_HD S5_INL bool trace_eqplane_pixel_t(const S5ImageConsts& c, int ix, int iy, PixelOut* o, AzIn* defer, Geodesic* gslot = nullptr)
{
    double alpha, beta;
    pixel_impact(c, ix, iy, &alpha, &beta);
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = 0.0;
    o->intensity = o->tau = o->qerr = o->height = o->delay = 0.0; o->steps = 0;

    Geodesic gd_local;
    Geodesic& gd = SGD ? *gslot : gd_local;
    int error = 0;
    RayCache k;
    gd.type = -1;
    /* stage 1: geodesic_init_inf up to the radial roots (sim5kerr-geod.c:59-86) */
    bool alive = true;
    {
        double a = c.a, i = c.incl;
        if ((a < 0.0) || (a > 1. - 1e-6)) { error = GD_ERROR_SPIN_RANGE; alive = false; }
        else if ((i <= 0.0) || (i >= S5_PI_HALF)) { error = GD_ERROR_INCL_RANGE; alive = false; }
        else {
            if (beta == 0.0) beta = +1e-6;
            gd.a = fmax(1e-4, a);
            gd.incl = i;
            gd.cos_i = c.cos_i;
            gd.alpha = alpha;
            gd.beta = beta;
            gd.l = -alpha * c.sin_i;
            gd.q = sq(beta) + sq(c.cos_i) * (sq(alpha) - sq(a));
            if (gd.q == 0.0) { error = GD_ERROR_Q_RANGE; alive = false; }
            else {
                k.isn_inf = 0.0;
                if (!geodesic_R_roots(&gd, 1.7976931348623157e308, &error, &k.isn_inf)) alive = false;
            }
        }
    }
    S5_STAGE_SYNC();
    /* stage 2: polar roots and the two polar Carlson integrals (sim5kerr-geod.c:88-96), kept for the crossings and the azimuth */
    if (alive) {
        if (!geodesic_T_roots(&gd, gd.cos_i, &error)) alive = false;
        else {
            /* theta_int(0) == mK*K(mm): jacobi_icn(0/sqrt(m2p), mm) takes its z == 0 exit (0 <= mm < 1 here) */
            k.K_mm = elliptic_k(gd.mm);
            k.u_pol = polar_amplitude(&gd);
            k.icn_u = jacobi_icn_ex(k.u_pol, gd.mm, &k.rf_u, &k.rf_u_z, &k.rf_u_m, &k.have_rf_u);
            gd.Tpp = 2. * (gd.mK * k.K_mm);
            gd.Tip = gd.mK * k.icn_u;
        }
    }
    S5_STAGE_SYNC();
    unsigned gt = 0;
    int hit_order = -1;
    double P = 0.0, r = 0.0;
    if (!alive) {
        int g0 = (error == GD_ERROR_TYPE_RR_DOUBLE) ? gtype_code(gd.type) : SIM5_GT_NONE;
        o->status = (unsigned)((SIM5_ST_INITERR + error) | (g0 << 5));
    } else {
        /* stage 3: the order-0 crossing and its radius */
        gt = (unsigned)gtype_code(gd.type) << 5;
        o->status = SIM5_ST_MISS | gt;
        P = crossing_cached(&gd, 0, k);
        if (isnan(P)) { o->status = SIM5_ST_NOCROSS0 | gt; alive = false; }
        else {
            r = geodesic_position_rad(&gd, P);
            if (r >= c.rmin_emit) hit_order = 0;
        }
    }
    S5_STAGE_SYNC();
    if (alive && hit_order < 0) {
        for (int order = 1; order <= c.max_order; order++) {
            P = crossing_cached(&gd, order, k);
            if (isnan(P)) { o->status = (order == 1 ? SIM5_ST_NOCROSS1 : SIM5_ST_NOCROSS2) | gt; alive = false; break; }
            r = geodesic_position_rad(&gd, P);
            if (r >= c.rmin_emit) { hit_order = order; break; }
        }
    }
    /* stage 4: emission side of the hit */
    bool deferred = false;
    if (alive && hit_order >= 0) {
        o->status = (hit_order == 0 ? SIM5_ST_HIT0 : hit_order == 1 ? SIM5_ST_HIT1 : SIM5_ST_HIT2) | gt;
        o->r = r;
        if (c.outputs & SIM5_OUT_PHI) {
            if (DEFER) {
                if (gd.type == GEOD_TYPE_RR || gd.type == GEOD_TYPE_RC) {
                    if (SGD) az_make_tail(&gd, k, r, P, defer); else az_make(&gd, k, r, P, defer);
                    deferred = true;
                }
                else o->phi = NAN;               /* geodesic_position_azm returns NaN for the other types */
            } else {
                o->phi = (c.flags & SIM5_FLAG_EXACT_AZIMUTH) ? azimuth_equatorial(&gd, k, r, P) : azimuth_equatorial_default(&gd, k, r, P);
            }
        }
        if (DELAY) {
            if (c.outputs & SIM5_OUT_DELAY) {
                double Pref = geodesic_P_int(&gd, c.delay_r_ref, 0);
                o->delay = geodesic_timedelay(&gd, Pref, c.delay_r_ref, 0.0, P, r, 0.0);
            }
        }
        if (c.mode == SIM5_MODE_POLARIZED) {
            polarized_hit(c, &gd, r, P, o);
        } else {
            double g = gfactorK(r, c.a, gd.l);
            double f = disk_nt_flux(c, r);
            o->g = g;
            o->flux = f * crm::cr_pow_4(g);
        }
    }
    return deferred;
}
S5_HD S5_INL void trace_eqplane_pixel(const S5ImageConsts& c, int ix, int iy, PixelOut* o)
{
    if (c.outputs & SIM5_OUT_DELAY) trace_eqplane_pixel_t<false, true>(c, ix, iy, o, nullptr);
    else                            trace_eqplane_pixel_t<false>(c, ix, iy, o, nullptr);
}

/* mode SPECTRUM: what one disk hit contributes.  Returns false for pixels without a hit (or with T = 0 / g <= 0, which
 * DiskRaytrace.spectrum skips, python/sim5diskraytrace.py:100,115).  The contribution to energy E_k is
 *     amp * E_k^3 / expm1(xs * E_k)        (blackbody() of sim5radiation.c:73-76 at E_k/g, times g^3 dalpha dbeta)
 * with the (1/g)^3 of the emitted energy and the g^3 of the intensity invariant kept as the reference has them. */
struct SpecHit { double amp3, ginv, xs; };      /* amp3 = limb factor * BB1 * g^3 * dalpha dbeta ; ginv = 1/g ; xs = BB2 (hardening and T included) */
/* SYNC: every thread of the CTA calls (lockstep kernels, see trace_eqplane_pixel_t); gslot = the thread's shared-memory geodesic slot */
template <bool SYNC>
S5_HD S5_INL bool spectrum_pixel_t(const S5ImageConsts& c, int ix, int iy, SpecHit* h, unsigned* status, Geodesic* gslot)
{
    PixelOut o;
    AzIn z;                                                    /* never filled: no phi is requested (DEFER only keeps the azimuth code out of the kernel) */
    if (SYNC) trace_eqplane_pixel_t<true, false, true, true>(c, ix, iy, &o, &z, gslot);
    else      trace_eqplane_pixel_t<true>(c, ix, iy, &o, &z);  /* c.mode == POLARIZED: g and mu_e from the Keplerian emitter frame */
    *status = o.status;
    unsigned cls = o.status & 31;
    if (!(cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2)) return false;
    double F = disk_nt_flux(c, o.r);
    double T = sqrt(sqrt(F / 5.670400e-05));                   /* T_eff = (F / sigma_SB)^(1/4), python/sim5diskmodel.py:48 */
    if (!(T > 0.0) || !(o.g > 0.0)) return false;
    double limbf = (c.spec_limb && o.mue >= 0.0) ? 0.5 + 0.75 * o.mue : 1.0;
    h->amp3 = limbf * c.bb1 * (o.g * o.g * o.g) * (c.da * c.db);
    h->ginv = 1.0 / o.g;
    h->xs = c.bb2 / T;
    return true;
}
S5_HD S5_INL bool spectrum_pixel(const S5ImageConsts& c, int ix, int iy, SpecHit* h, unsigned* status)
{
    return spectrum_pixel_t<false>(c, ix, iy, h, status, nullptr);
}
/* one (hit, energy) term: Iv = BB1 E^3 / expm1(BB2 E) at E = E_k / g, times g^3 dA -- with the per-hit factors hoisted
 * (the reference-side driver divides and multiplies per term; the two agree to rounding, far inside the 1e-7 bar) */
S5_HD S5_INL double spectrum_term(const SpecHit& h, double Ek)
{
    double E = Ek * h.ginv;
    double x = h.xs * E;
    /* far Wien tail: expm1(x) == e^x to every bit and overflows at x > 709, where the approximate reciprocal has no inf -> 0 path */
    double w = (x < 600.0) ? ff::rcp_ap(expm1(x)) : exp(-x);
    return h.amp3 * (E * E * E) * w;
}

/* mode STEPWISE: raytrace() through the harness torus (SURVEY.md 8d cfg 4; oracle/ref_driver.c pixel_stepwise) */
struct StepRay {           /* live state of one stepwise ray (the persistent kernel keeps this per lane) */
    double x[4], k[4];
    RayData rtd;
    double I, tau;
    crm::AngCarry ac;      /* the polar angle behind x[2], left by the previous step's cosine */
    double dl, th0;        /* a step whose RK4 fallback is pending (k_trace_lanes runs those for several lanes at once) */
    int steps;
    unsigned gt;
};
/* returns true if the ray is live (needs stepping); otherwise o->status is final */
S5_HD S5_INL bool stepwise_start(const S5ImageConsts& c, int ix, int iy, StepRay* s, PixelOut* o)
{
    double alpha, beta;
    pixel_impact(c, ix, iy, &alpha, &beta);
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = 0.0;
    o->intensity = o->tau = o->qerr = o->height = o->delay = 0.0; o->steps = 0;
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
    crm::carry_reset(&s->ac);
    s->I = 0.0; s->tau = 0.0; s->steps = 0;
    return true;
}
/* a raytrace() call in two parts: stepwise_try runs the Verlet step and returns true if the RK4 fallback is needed (stepwise_rk4);
 * stepwise_post = the torus emission along the finished step and the termination tests: 0 while the ray is live, else the class */
S5_HD S5_INL bool stepwise_try(const S5ImageConsts& c, StepRay* s)
{
    s->dl = c.step_max;
    return raytrace_verlet(s->x, s->k, &s->dl, &s->rtd, &s->ac, &s->th0);
}
S5_HD S5_INL void stepwise_rk4(const S5ImageConsts& c, StepRay* s)
{
    (void)c;
    raytrace_rk4(s->x, s->k, s->dl, &s->rtd, s->th0, &s->ac);
}
S5_HD S5_INL int stepwise_post(const S5ImageConsts& c, StepRay* s)
{
    const double dl = s->dl;
    s->steps++;
    {
        double r = s->x[1], m = s->x[2];
        double R = r * sqrt(1.0 - m * m);
        double z = r * m;
        double sv = sq((R - c.torus_rc) / c.torus_w) + sq(z / (c.torus_h * R));
        if (sv < 13.8) {
            S5_STAT(7);
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
/* one raytrace() call + torus emission; returns 0 while the ray is live, else the termination class */
S5_HD S5_INL int stepwise_step(const S5ImageConsts& c, StepRay* s)
{
    if (stepwise_try(c, s)) stepwise_rk4(c, s);
    return stepwise_post(c, s);
}
S5_HD S5_INL void stepwise_finish(const S5ImageConsts& c, StepRay* s, int cls, PixelOut* o)
{
    (void)c;
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = o->height = o->delay = 0.0;
    o->intensity = s->I;
    o->tau = s->tau;
    o->steps = s->steps;
    o->qerr = raytrace_error(s->x, s->k, &s->rtd);
    o->status = (unsigned)cls | s->gt;
}


/* ------------------------------------------------------------------------------------------------------------------
 * mode SURFACE: the thick-disk surface finder of the reference's Python layer (DiskRaytrace.geodesic / __find_surface /
 * image, python/sim5diskraytrace.py:214-336, 163-205, 340-391; reference-side spelling: oracle/ref_driver.c pixel_surface)
 * as a per-lane state machine.  The unit of work is ONE pass of geodesic_follow's do-while body (sim5kerr-geod.c:907-921:
 * a bounded advance of P, then r(P) and cos theta(P), two Jacobi sn/cn/dn evaluations), so the lanes of a warp stay
 * together although every ray takes a different number of follow calls and of sub-steps inside each call
 * (~400 sub-steps per ray from r0 >= 200 down to the disk); finished lanes are refilled by the persistent kernel.
 * ------------------------------------------------------------------------------------------------------------------ */
struct SurfRay {
    Geodesic gd;
    double P, r, m;             /* current position on the geodesic */
    double H1, Hd;              /* height of the ray and of the surface at the last completed forward follow */
    double r0, step_factor;
    double step, rem;           /* argument of the running geodesic_follow call, and what is left of it */
    int phase;                  /* running call: 0 forward probe, 1 step back by -step, 2 final step back by -step/2 */
    int iteration, nfollow;
    unsigned gt;
};
S5_HD S5_INL double surf_h(const S5ImageConsts& c, double R) { return (R > c.surf_rin) ? c.surf_hr * sq(R - c.surf_rin) / R : 0.0; }
S5_HD S5_INL double surf_dhdr(const S5ImageConsts& c, double R) { return (R > c.surf_rin) ? c.surf_hr * (1.0 - sq(c.surf_rin) / sq(R)) : 0.0; }
S5_HD S5_INL void surface_follow(SurfRay* s, double step, int phase) { s->step = (phase == 0) ? step : s->step; s->rem = step; s->phase = phase; s->nfollow++; }
S5_HD S5_INL double surface_next_step(const SurfRay* s) { return fmax(1e-2 / 2., fmin((s->H1 - s->Hd) / 2., 0.5 * (sqrt(s->r) - 0.99) * s->step_factor)); }

/* start (or restart with a larger r0) the search: python/sim5diskraytrace.py:259-289.  -1 = live, else the termination class (SIM5_ST_HIT0 is 0) */
S5_HD S5_MID int surface_begin(const S5ImageConsts& c, SurfRay* s)
{
    if (s->iteration > 3) return SIM5_ST_ESCAPE;
    const Geodesic* gd = &s->gd;
    double r0 = fmax(fmax(200.0, 1.1 * gd->rp), (0.5 + (double)s->iteration) * sqrt(sq(gd->alpha) + sq(gd->beta)) / c.surf_cos_it);
    double P1, r1, m1, H1, Hd;
    for (;;) {
        P1 = geodesic_P_int(gd, r0, 0);
        r1 = geodesic_position_rad(gd, P1);
        m1 = geodesic_position_pol(gd, P1);
        double R1 = r1 * sqrt(1. - m1 * m1);
        H1 = r1 * m1;
        Hd = surf_h(c, R1);
        if ((Hd < H1) || (r0 > 5e6)) break;
        r0 = 2.0 * r0;
    }
    if (!(Hd < H1)) return SIM5_ST_SURF_BELOW;
    s->r0 = r0; s->P = P1; s->r = r1; s->m = m1; s->H1 = H1; s->Hd = Hd;
    s->step_factor = 1.0;
    surface_follow(s, surface_next_step(s), 0);
    return -1;
}
S5_HD S5_INL void surface_fail(PixelOut* o, unsigned status, int nfollow)
{
    o->r = o->phi = o->g = o->flux = o->chi = o->delta = o->mue = o->intensity = o->tau = o->qerr = o->height = o->delay = 0.0;
    o->steps = nfollow;
    o->status = status;
}
/* emission-side quantities at the surface point (P, r, m): DiskRaytrace.image + __tetrad/__gfactor/__emission_angle */
S5_HD S5_MID void surface_finish(const S5ImageConsts& c, SurfRay* s, int cls, PixelOut* o)
{
    surface_fail(o, (unsigned)cls | s->gt, s->nfollow);
    if (cls != SIM5_ST_HIT0 && cls != SIM5_ST_SURF_EQPLANE) return;
    double r = s->r, m = s->m, a = c.a;
    if (isnan(r)) { o->status = SIM5_ST_MISS | s->gt; return; }
    double k[4];
    photon_momentum(a, r, m, s->gd.l, s->gd.q, s->gd.Rpc - s->P, 1.0, k);
    double R = r * sqrt(1. - m * m);
    o->r = r;
    o->height = r * m;
    double F = disk_nt_flux(c, R);
    if (F == 0.0) return;
    Metric M;
    Tetrad t;
    double U[4], N[4];
    const double e0[4] = {1.0, 0.0, 0.0, 0.0};
    const double e2[4] = {0.0, 0.0, 1.0, 0.0};
    kerr_metric(a, r, m, &M);
    tetrad_surface(&M, Omega_from_ell(ellK(R, a), &M), 0.0, (m > 0.0) ? surf_dhdr(c, R) : 0.0, &t);
    on2bl(e0, U, &t);
    on2bl(e2, N, &t);
    double g = (k[0] * M.g00 + k[3] * M.g03) / dotprod(k, U, &M);
    if (!(g > 0.0)) g = 0.0;
    double mue = dotprod(k, N, &M) / dotprod(k, U, &M);
    if (mue < 0.0 && mue > -1e-2) mue = 1e-3;
    double limb = 0.5 + 0.75 * mue;
    if (!(g > 0.0)) return;
    o->g = g;
    o->mue = mue;
    o->flux = F * crm::cr_pow_4(g) * limb;
}
/* returns true if the ray is live; otherwise *o is final */
S5_HD S5_INL bool surface_start(const S5ImageConsts& c, int ix, int iy, SurfRay* s, PixelOut* o)
{
    double alpha, beta;
    pixel_impact(c, ix, iy, &alpha, &beta);
    int error = 0;
    s->gd.type = -1;
    s->nfollow = 0;
    if (!geodesic_init_inf_sc(c.incl, c.sin_i, c.cos_i, c.a, alpha, beta, &s->gd, &error)) {
        int gt = (error == GD_ERROR_TYPE_RR_DOUBLE) ? gtype_code(s->gd.type) : SIM5_GT_NONE;
        surface_fail(o, (unsigned)((SIM5_ST_INITERR + error) | (gt << 5)), 0);
        return false;
    }
    s->gt = (unsigned)gtype_code(s->gd.type) << 5;
    if (c.surf_flat) {                                   /* flat disk: the order-0 equatorial crossing, nothing to follow */
        s->P = geodesic_find_midplane_crossing(&s->gd, 0);
        s->r = geodesic_position_rad(&s->gd, s->P);
        s->m = 0.0;
        surface_finish(c, s, SIM5_ST_SURF_EQPLANE, o);
        return false;
    }
    s->iteration = 0;
    int cls = surface_begin(c, s);
    if (cls >= 0) { surface_finish(c, s, cls, o); return false; }
    return true;
}
/* one pass of geodesic_follow's loop body; when that completes a follow call, the decision of __find_surface that follows
 * it (python/sim5diskraytrace.py:296-331).  -1 while the ray is live, else the termination class */
S5_HD S5_MID int surface_step(const S5ImageConsts& c, SurfRay* s)
{
    const Geodesic* gd = &s->gd;
    bool ok = true, done = false;
    {
        const double MAXSTEP_FACTOR = 5e-2;
        double truestep = s->rem / fabs(s->rem) * fmin(fabs(s->rem), MAXSTEP_FACTOR * sqrt(s->r));
        s->P = s->P + truestep / (sq(s->r) + sq(gd->a * s->m));
        s->r = geodesic_position_rad(gd, s->P);
        s->m = geodesic_position_pol(gd, s->P);
        if (s->r < 1.01 * r_bh(gd->a)) { ok = false; done = true; }
        else if ((s->P < 0.0) || (s->P > 2. * gd->Rpc)) { ok = false; done = true; }
        else { s->rem -= truestep; done = !(fabs(s->rem) > 1e-5); }
    }
    if (!done) return -1;
    if (s->phase == 2) return SIM5_ST_HIT0;              /* the caller does not look at the status of the step back */
    if (s->phase == 1) {
        s->step_factor = s->step_factor / 5.;
    } else {
        const double accuracy = 1e-2;
        if (!ok) return SIM5_ST_SURF_LOST;
        double R1 = s->r * sqrt(1. - s->m * s->m);
        s->H1 = s->r * s->m;
        s->Hd = surf_h(c, R1);
        if (s->H1 <= s->Hd) {
            if (s->step < accuracy) surface_follow(s, -s->step / 2., 2);
            else                    surface_follow(s, -s->step, 1);
            return -1;
        }
        if (s->H1 < 1e-4) {
            s->P = geodesic_find_midplane_crossing(gd, 0);
            s->r = geodesic_position_rad(gd, s->P);
            s->m = geodesic_position_pol(gd, s->P);
            return SIM5_ST_SURF_EQPLANE;
        }
        if (s->r < 1.05 * c.r_bh) return SIM5_ST_HORIZON;
        if (s->r > 1.1 * s->r0) { s->iteration++; return surface_begin(c, s); }
        if (s->m < 0.0) return SIM5_ST_SURF_UNDER;
        if (s->step < accuracy / 2.) return SIM5_ST_MAXSTEPS;
    }
    surface_follow(s, surface_next_step(s), 0);
    return -1;
}

/* the two ray programs of the lane-refill kernel (kernels.cuh:k_trace_lanes) */
#ifndef S5_STEP_THREADS
#define S5_STEP_THREADS 128
#endif
#ifndef S5_STEP_BATCH
#define S5_STEP_BATCH 0           /* 0: free-running warps with lane refill (steps per ray vary by an order of magnitude) */
#endif
#ifndef S5_MIN_CTAS_STEP
#define S5_MIN_CTAS_STEP 4
#endif
#ifndef S5_STEP_REFILL_MIN
#define S5_STEP_REFILL_MIN 8      /* idle lanes of a warp that trigger a refill: 2 -> 354 ms, 4 -> 324, 8 -> 311 (cfg 4 at 1024^2, profiles/r05g_step_sweep.log):
                                     a refill runs the ray start (init_inf, P_int: ~30 steps' worth) with only the idle lanes active */
#endif
#ifndef S5_RK4_BATCH
#define S5_RK4_BATCH 1            /* lanes of a warp that must wait for the RK4 fallback before it runs; 1: run it at once, as raytrace() does.
                                     Holding the fallback back does NOT pay (cfg 4 at 512^2 / 1024^2, ms, profiles/r05f_step_sweep.log): 1 -> 98.3 / 324,
                                     3 -> 121 / 371, 4 -> 133 / 389, 6 -> 140 / 419, 8 -> 151 / 434.  RK4 steps come in runs (a ray needs the fallback for
                                     many consecutive steps around a turning point, and so do its neighbours in the warp at about the same time), so the
                                     lanes of a warp already take it together, and a lane that waits holds up a long run */
#endif
struct StepwiseProg {
    typedef StepRay State;
    static const int REFILL_MIN = S5_STEP_REFILL_MIN;   /* idle lanes of a warp that trigger a refill */
    static const int MIN_CTAS = S5_MIN_CTAS_STEP; /* resident 128-thread CTAs per SM the kernel is compiled for */
    static const int THREADS = S5_STEP_THREADS;
    static const int BATCH = S5_STEP_BATCH;
    static S5_HD S5_INL bool start(const S5ImageConsts& c, int ix, int iy, State* s, PixelOut* o) { return stepwise_start(c, ix, iy, s, o); }
    static S5_HD S5_INL int step(const S5ImageConsts& c, State* s) { return stepwise_step(c, s); }
    static S5_HD S5_INL void finish(const S5ImageConsts& c, State* s, int cls, PixelOut* o) { stepwise_finish(c, s, cls, o); }
    /* the step in two parts: `slow` (the RK4 fallback, 0.3 % of the steps, ~4 steps' worth of work) is held back until DEFER_MIN lanes of the
     * warp want it (or no lane can do anything else).  Per-ray arithmetic and order are unchanged.  tools/step_stats.cpp: with 32 lanes a warp
     * met an RK4 step in 8.9 % of its rounds and idled 31 lanes for it */
#if defined(S5_STEP_ROW_MAJOR)
    static const bool CENTER_OUT = false;
#else
    static const bool CENTER_OUT = true;          /* hand the rows out from the middle of the image outwards (k_trace_lanes) */
#endif
    static const bool DEFERS = (S5_RK4_BATCH > 1);
    static const int DEFER_MIN = S5_RK4_BATCH;
    static S5_HD S5_INL bool try_step(const S5ImageConsts& c, State* s) { return stepwise_try(c, s); }
    static S5_HD S5_INL void slow_step(const S5ImageConsts& c, State* s) { stepwise_rk4(c, s); }
    static S5_HD S5_INL int post(const S5ImageConsts& c, State* s) { return stepwise_post(c, s); }
};
/* Launch shape of the SURFACE lane kernel (1024^2 preset, ms; profiles/r01x_sweep.log, r02z_surf_sweep.log).  Lane refill does not pay
 * here: a ray's start (init_inf, P_int) costs ~30 sub-steps, so refilling a few idle lanes while the rest of the warp waits loses more
 * than it gains -- refill at 4 / 8 / 16 / 24 idle lanes 61.9 / 60.1 / 57.9 / 56.4, only when the whole warp is idle 56.0 (neighbouring
 * pixels take similar numbers of sub-steps).  Occupancy pays: 1 CTA/SM of 128 threads (226 regs) 56.0, 3 (168) 47.6, 4 (128) 44.1,
 * 5 (96 regs, 20 warps/SM) 42.7, 6 (80) 44.2, 8 (64) 46.5.  And so does the plain CTA-batch loop (k_trace_lanes, PROG::BATCH: the CTA
 * takes blockDim.x consecutive pixels, starts them together and steps until the last one is done -- no refill bookkeeping in the
 * hot loop): 128 x 5 CTAs 34.0, 64 x 10 33.0, 32 x 20 33.7, 128 x 4 37.1, 128 x 6 36.6, 192 x 4 37.0, 256 x 3 37.9, 384 x 2 39.4, 512 x 1 45.5,
 * 640 x 1 41.5 (the larger the batch, the longer it waits for its slowest ray). */
#ifndef S5_SURF_REFILL_MIN
#define S5_SURF_REFILL_MIN 32
#endif
#ifndef S5_SURF_MIN_CTAS
#define S5_SURF_MIN_CTAS 10
#endif
#ifndef S5_SURF_THREADS
#define S5_SURF_THREADS 64
#endif
#ifndef S5_SURF_BATCH
#define S5_SURF_BATCH 1           /* > 0: CTA-batch variant of the lane kernel with a barrier every S5_SURF_BATCH sub-steps; 0: warps with lane refill */
#endif
#ifndef S5_MIN_CTAS_STEP
#define S5_MIN_CTAS_STEP 4
#endif
struct SurfaceProg {
    typedef SurfRay State;
    static const int REFILL_MIN = S5_SURF_REFILL_MIN;
    static const int MIN_CTAS = S5_SURF_MIN_CTAS;
    static const int THREADS = S5_SURF_THREADS;
    static const int BATCH = S5_SURF_BATCH;
    static S5_HD S5_INL bool start(const S5ImageConsts& c, int ix, int iy, State* s, PixelOut* o) { return surface_start(c, ix, iy, s, o); }
    /* the kernel's protocol is "0 while live"; SIM5_ST_HIT0 is 0, so the class travels with bit 8 set */
    static const bool CENTER_OUT = false;
    static const bool DEFERS = false;
    static const int DEFER_MIN = 1;
    static S5_HD S5_INL bool try_step(const S5ImageConsts&, State*) { return false; }
    static S5_HD S5_INL void slow_step(const S5ImageConsts&, State*) { }
    static S5_HD S5_INL int post(const S5ImageConsts&, State*) { return 0; }
    static S5_HD S5_INL int step(const S5ImageConsts& c, State* s) { int r = surface_step(c, s); return r < 0 ? 0 : (r | 0x100); }
    static S5_HD S5_INL void finish(const S5ImageConsts& c, State* s, int cls, PixelOut* o) { surface_finish(c, s, cls & 0xff, o); }
};

} /* namespace s5 */
#endif
