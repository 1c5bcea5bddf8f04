/*
 * geod.cuh -- analytic Kerr null geodesics (roots of R and Theta, position integral and its inversions,
 * equatorial crossing, azimuth, momentum) as sm_100a device code.
 *
 * Bit-for-bit behavioural twin of the listed functions of the reference's src/sim5kerr-geod.c
 * (operation order follows the reference; see elliptic.cuh for the contract).  Differences in form:
 *   - the 4 complex roots are classified and ordered without the generic array sort
 *     (sim5polyroots.c:277-324); the two csqrt pairs are either both real or both complex, which
 *     leaves four cases with a fixed outcome;
 *   - the `long double` arithmetic of geodesic_priv_T_roots (sim5kerr-geod.c:1125-1131) is
 *     emulated to the bit (crm::x87_mu_roots);
 *   - nothing prints; NaN / FALSE + error code are returned exactly where the reference does.
 */
#ifndef SIM5_GEOD_CUH
#define SIM5_GEOD_CUH

#include "kerr.cuh"

namespace s5 {

enum {
    GEOD_TYPE_RR = 40, GEOD_TYPE_RR_DBL = 41, GEOD_TYPE_RR_BH = 42, GEOD_TYPE_RC = 2, GEOD_TYPE_CC = 0
};
enum {
    GD_OK = 0, GD_ERROR_Q_ZERO = 1, GD_ERROR_BOUND_GEODESIC = 2, GD_ERROR_UNKNOWN_SOLUTION = 3,
    GD_ERROR_TYPE_RR_DOUBLE = 4, GD_ERROR_TYPE_CC = 5, GD_ERROR_Q_RANGE = 7, GD_ERROR_MUPLUS_RANGE = 8,
    GD_ERROR_MU0_RANGE = 9, GD_ERROR_MM_RANGE = 10, GD_ERROR_INCL_RANGE = 11, GD_ERROR_SPIN_RANGE = 12
};

#define S5_PI_HALF 1.57079632679          /* the reference's truncated constant, sim5math.h:39 */

struct Cplx { double re, im; };
struct Geodesic {          /* == struct geodesic, sim5kerr-geod.h:42-68 (240 bytes) */
    double a, alpha, beta, incl, cos_i;
    double l, q;
    Cplx r1, r2, r3, r4;
    int nrr, type;
    double m2p, m2m, mm, mK;
    double rp, dmdp_inf;
    double Rpc, Tpp, Tip;
    double k[4];
    double p;
};
static_assert(sizeof(Geodesic) == 240, "geodesic ABI");

/* csqrt of a real number (imaginary part +0), glibc semantics */
template <class OPS>
S5_HD S5_INL Cplx csqrt_real(OPS& o, double x)
{
    if (x != x) return Cplx{x, x};
    if (x < 0.0) return Cplx{0.0, o.sqrt(-x)};
    return Cplx{fabs(o.sqrt(x)), 0.0};
}

/* sim5math.c:49-58 */
S5_HD S5_INL int ensure_range(double* v, double lo, double hi, double acc)
{
    if (*v < lo - acc) return 0;
    if (*v > hi + acc) return 0;
    if (*v < lo) *v = lo;
    if (*v > hi) *v = hi;
    return 1;
}

/* roots of R(r), geodesic class, pericentre and R-integral to the pericentre.  sim5kerr-geod.c:985-1104 */
template <class OPS>
S5_HD S5_INL int geodesic_R_roots_t(OPS& o, Geodesic* g, double r0, int* error, double* isn_inf)
{
    double a = g->a, l = g->l, q = g->q;
    double a2 = sq(a), l2 = sq(l);
    double A, B, C, D, E, F, X, Z, z;

    C = sq(a - l) + q;
    D = 2. / 3. * (q + l2 - a2);
    E = 9. / 4. * sq(D) - 12. * a2 * q;
    F = -27. / 4. * (D * D * D) - 108. * a2 * q * D + 108. * sq(C);
    X = sq(F) - 4. * (E * E * E);
    if (X >= 0) {
        double sX = o.sqrt(X);
        A = (F > sX ? +1 : -1) * 1. / 3. * crm::cr_pow_third(fabs(F - sX) / 2.) +
            (F > -sX ? +1 : -1) * 1. / 3. * crm::cr_pow_third(fabs(F + sX) / 2.);
    } else {
        const double sX54 = o.div(o.sqrt(-X), 54.), F54 = o.div(F, 54.);
        Z = o.sqrt(sq(F54) + sq(sX54));
        z = cr_atan2(sX54, F54);
        A = crm::cr_pow_third(Z) * 2. * crm::cr_cos(o.div(z, 3.));
    }
    B = o.sqrt(A + D);
    const double C4B = o.div(4. * C, B);
    Cplx s12 = csqrt_real(o, -A + 2. * D - C4B);
    Cplx s34 = csqrt_real(o, -A + 2. * D + C4B);
    Cplx p1 = Cplx{+B / 2. + .5 * s12.re,  .5 * s12.im};
    Cplx p2 = Cplx{+B / 2. - .5 * s12.re, -(.5 * s12.im)};
    Cplx p3 = Cplx{-B / 2. + .5 * s34.re,  .5 * s34.im};
    Cplx p4 = Cplx{-B / 2. - .5 * s34.re, -(.5 * s34.im)};

    /* sort_roots(): real roots first in descending order, then the complex ones in input order */
    bool real12 = (p1.im == 0.) && (p2.im == 0.);
    bool real34 = (p3.im == 0.) && (p4.im == 0.);
    if (real12 && real34) {
        /* p1 >= p2 and p3 >= p4 already; merge the two ordered pairs (selection sort of the reference
           gives plain descending order; ties are value-identical) */
        double v0 = p1.re, v1 = p2.re, v2 = p3.re, v3 = p4.re;
        double t;
        /* 4-element descending sorting network */
        if (v2 > v0) { t = v0; v0 = v2; v2 = t; }
        if (v3 > v1) { t = v1; v1 = v3; v3 = t; }
        if (v1 > v0) { t = v0; v0 = v1; v1 = t; }
        if (v3 > v2) { t = v2; v2 = v3; v3 = t; }
        if (v2 > v1) { t = v1; v1 = v2; v2 = t; }
        g->r1 = Cplx{v0, 0.0}; g->r2 = Cplx{v1, 0.0}; g->r3 = Cplx{v2, 0.0}; g->r4 = Cplx{v3, 0.0};
        /* keep the sign of the zero imaginary parts as the reference would: irrelevant to every consumer */
        g->nrr = 4;
    } else if (real12) {
        g->r1 = p1; g->r2 = p2; g->r3 = p3; g->r4 = p4; g->nrr = 2;
    } else if (real34) {
        g->r1 = p3; g->r2 = p4; g->r3 = p1; g->r4 = p2; g->nrr = 2;
    } else {
        g->r1 = p1; g->r2 = p2; g->r3 = p3; g->r4 = p4; g->nrr = 0;
    }

    switch (g->nrr) {
        case 4:
            g->type = GEOD_TYPE_RR;
            if ((r0 < g->r3.re) || ((r0 > g->r2.re) && (r0 < g->r1.re))) {
                if (error) *error = GD_ERROR_UNKNOWN_SOLUTION;
                return 0;
            }
            if (fabs(g->r1.re - g->r2.re) < 1e-8) {
                g->type = GEOD_TYPE_RR_DBL;
                if (error) *error = GD_ERROR_TYPE_RR_DOUBLE;
                return 0;
            }
            if ((r0 >= g->r3.re) && (r0 <= g->r2.re)) g->type = GEOD_TYPE_RR_BH;
            break;
        case 2:  g->type = GEOD_TYPE_RC; break;
        default: g->type = GEOD_TYPE_CC; break;
    }

    double r1, r2, r3, r4, u, v, mm;
    switch (g->type) {
        case GEOD_TYPE_RR:
            r1 = g->r1.re; r2 = g->r2.re; r3 = g->r3.re; r4 = g->r4.re;
            mm = o.div((r2 - r3) * (r1 - r4), (r2 - r4) * (r1 - r3));
            g->rp = r1;
            {
                double u = jacobi_isn(o.sqrt(o.div(r2 - r4, r1 - r4)), mm);
                if (isn_inf) *isn_inf = u;          /* == the u1 of integral_R_rp_re_inf, sim5elliptic.c:1039-1040 */
                g->Rpc = o.div(2., o.sqrt((r1 - r3) * (r2 - r4))) * u;
            }
            break;
        case GEOD_TYPE_RR_BH:
            r1 = g->r1.re; r2 = g->r2.re; r3 = g->r3.re; r4 = g->r4.re;
            mm = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
            g->rp = r2;
            g->Rpc = 2. / sqrt((r1 - r3) * (r2 - r4)) * elliptic_k(mm);
            break;
        case GEOD_TYPE_RC:
            r1 = g->r1.re; r2 = g->r2.re; u = g->r3.re; v = g->r3.im;
            A = o.sqrt(sq(r1 - u) + sq(v));
            B = o.sqrt(sq(r2 - u) + sq(v));
            mm = o.div(sq(A + B) - sq(r1 - r2), 4. * A * B);
            g->rp = r1;
            g->Rpc = o.div(1., o.sqrt(A * B)) * jacobi_icn(o.div(A - B, A + B), mm);
            break;
        default: {  /* CC */
            r1 = g->r1.re; r2 = g->r3.re; r3 = g->r1.im; r4 = g->r3.im;
            A = sqrt(sq(r1 - r2) + sq(r3 + r4));
            B = sqrt(sq(r1 - r2) + sq(r3 - r4));
            double g1 = sqrt((4. * sq(r3) - sq(A - B)) / (sq(A + B) - 4. * sq(r3)));
            mm = 4. * A * B / sq(A + B);
            g->rp = r1 - r3 * g1;
            g->Rpc = 2. / (A + B) * jacobi_itn(-1. / g1, mm);
            break;
        }
    }
    return 1;
}
S5_HD S5_MID int geodesic_R_roots(Geodesic* g, double r0, int* error, double* isn_inf = nullptr)
{
    ff::Quick f;
    int rc = geodesic_R_roots_t(f, g, r0, error, isn_inf);
    if (!f.ok) { ff::Plain p; rc = geodesic_R_roots_t(p, g, r0, error, isn_inf); }
    return rc;
}

/* roots of Theta(mu).  sim5kerr-geod.c:1109-1184 (CPU branch: extended-precision m2m, m2p) */
template <class OPS>
S5_HD S5_INL int geodesic_T_roots_t(OPS& o, Geodesic* g, double m, int* error)
{
    double a = g->a, l = g->l, q = g->q;
    double a2 = sq(a), l2 = sq(l);
    crm::x87_mu_roots(q, l2, a2, &g->m2m, &g->m2p);

    if ((g->m2p <= 0.0) || (g->m2p >= 1.0)) {
        if (error) *error = GD_ERROR_MUPLUS_RANGE;
        return 0;
    }
    if (q > 0.0) {
        g->mm = o.div(g->m2p, g->m2p + g->m2m);
        if ((g->mm < 0.0) || (g->mm >= 1.0)) { if (error) *error = GD_ERROR_MM_RANGE; return 0; }
        if (fabs(m) > o.sqrt(g->m2p)) { if (error) *error = GD_ERROR_MU0_RANGE; return 0; }
        g->mK = o.div(1., o.sqrt(a2 * (g->m2p + g->m2m)));
    } else if (q < 0.0) {
        g->mm = (g->m2p + g->m2m) / g->m2p;
        if ((g->mm < 0.0) || (g->mm >= 1.0)) { if (error) *error = GD_ERROR_MM_RANGE; return 0; }
        if ((fabs(m) > sqrt(g->m2p)) || (fabs(m) < sqrt(-g->m2m))) { if (error) *error = GD_ERROR_MU0_RANGE; return 0; }
        g->mK = 1. / sqrt(a2 * g->m2p);
    } else {
        if (error) *error = GD_ERROR_Q_RANGE;
        return 0;
    }
    return 1;
}
S5_HD S5_MID int geodesic_T_roots(Geodesic* g, double m, int* error)
{
    ff::Quick f;
    int rc = geodesic_T_roots_t(f, g, m, error);
    if (!f.ok) { ff::Plain p; rc = geodesic_T_roots_t(p, g, m, error); }
    return rc;
}

S5_HD S5_INL double theta_int(const Geodesic* g, double x) { return g->mK * jacobi_icn(x / sqrt(g->m2p), g->mm); }
S5_HD S5_INL double theta_inv(const Geodesic* g, double x) { return sqrt(g->m2p) * jacobi_cn(x / g->mK, g->mm); }

/* geodesic from impact parameters at infinity.  sim5kerr-geod.c:41-100.
 * sin_i / cos_i are sin(i), cos(i) as the caller's libm gives them (per-image host constants in the
 * image kernels; cr_sincos in the scalar API). */
S5_HD S5_INL int geodesic_init_inf_sc(double i, double sin_i, double cos_i, double a, double alpha, double beta, Geodesic* g, int* error)
{
    if ((a < 0.0) || (a > 1. - 1e-6)) { if (error) *error = GD_ERROR_SPIN_RANGE; return 0; }
    if ((i <= 0.0) || (i >= S5_PI_HALF)) { if (error) *error = GD_ERROR_INCL_RANGE; return 0; }
    if (beta == 0.0) beta = +1e-6;

    g->a = fmax(1e-4, a);
    g->incl = i;
    g->cos_i = cos_i;
    g->alpha = alpha;
    g->beta = beta;
    g->l = -alpha * sin_i;
    g->q = sq(beta) + sq(cos_i) * (sq(alpha) - sq(a));
    if (g->q == 0.0) { if (error) *error = GD_ERROR_Q_RANGE; return 0; }

    if (!geodesic_R_roots(g, 1.7976931348623157e308, error)) return 0;
    if (!geodesic_T_roots(g, g->cos_i, error)) return 0;

    g->Tpp = 2. * theta_int(g, 0.0);
    g->Tip = theta_int(g, g->cos_i);
    if (error) *error = GD_OK;
    return 1;
}
S5_HD S5_INL int geodesic_init_inf(double i, double a, double alpha, double beta, Geodesic* g, int* error)
{
    double s, c;
    cr_sincos(i, &s, &c);
    return geodesic_init_inf_sc(i, s, c, a, alpha, beta, g, error);
}

/* r -> P.  sim5kerr-geod.c:178-263 */
S5_HD S5_MID double geodesic_P_int(const Geodesic* g, double r, int ppc)
{
    double r1, r2, r3, r4, u, v, mm, R, A, B;
    if (r == g->rp) return g->Rpc;
    switch (g->type) {
        case GEOD_TYPE_RR:
            r1 = g->r1.re; r2 = g->r2.re; r3 = g->r3.re; r4 = g->r4.re;
            mm = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
            R = 2. / sqrt((r1 - r3) * (r2 - r4)) * jacobi_isn(sqrt(((r2 - r4) * (r - r1)) / ((r1 - r4) * (r - r2))), mm);
            return (ppc) ? g->Rpc + R : g->Rpc - R;
        case GEOD_TYPE_RR_DBL:
            return NAN;
        case GEOD_TYPE_RR_BH:
            r1 = g->r1.re; r2 = g->r2.re; r3 = g->r3.re; r4 = g->r4.re;
            mm = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
            R = 2. / sqrt((r1 - r3) * (r2 - r4)) * jacobi_isn(sqrt((r1 - r3) / (r2 - r3) * (r2 - r) / (r1 - r)), mm);
            return (ppc) ? g->Rpc + R : g->Rpc - R;
        case GEOD_TYPE_RC:
            r1 = g->r1.re; r2 = g->r2.re; u = g->r3.re; v = g->r3.im;
            A = sqrt(sq(r1 - u) + sq(v));
            B = sqrt(sq(r2 - u) + sq(v));
            mm = (sq(A + B) - sq(r1 - r2)) / (4. * A * B);
            R = 1. / sqrt(A * B) * jacobi_icn(((A - B) * r + r1 * B - r2 * A) / ((A + B) * r - r1 * B - r2 * A), mm);
            return g->Rpc - R;
        case GEOD_TYPE_CC: {
            r1 = g->r1.re; r2 = g->r3.re; r3 = g->r1.im; r4 = g->r3.im;
            A = sqrt(sq(r1 - r2) + sq(r3 + r4));
            B = sqrt(sq(r1 - r2) + sq(r3 - r4));
            double g1 = sqrt((4. * sq(r3) - sq(A - B)) / (sq(A + B) - 4. * sq(r3)));
            mm = 4. * A * B / sq(A + B);
            R = 2. / (A + B) * jacobi_itn((r - r1 + r3 * g1) / (r3 + r1 * g1 - g1 * r), mm);
            return g->Rpc - R;
        }
    }
    return NAN;
}

/* P -> r.  sim5kerr-geod.c:290-357 */
template <class OPS>
S5_HD S5_INL double geodesic_position_rad_t(OPS& o, const Geodesic* g, double P)
{
    if ((P <= 0.0) || (P >= 2. * g->Rpc)) return NAN;
    if (P == g->Rpc) return g->rp;
    if (g->type == GEOD_TYPE_RR) {
        double r1 = g->r1.re, r2 = g->r2.re, r3 = g->r3.re, r4 = g->r4.re;
        double m4 = o.div((r2 - r3) * (r1 - r4), (r2 - r4) * (r1 - r3));
        double x4 = 0.5 * fabs(P - g->Rpc) * o.sqrt((r2 - r4) * (r1 - r3));
        double sn2 = sq(jacobi_sn(x4, m4));
        return o.div(r1 * (r2 - r4) - r2 * (r1 - r4) * sn2, r2 - r4 - (r1 - r4) * sn2);
    }
    if (g->type == GEOD_TYPE_RC) {
        if (P > g->Rpc) return NAN;
        double r1 = g->r1.re, r2 = g->r2.re, u = g->r3.re, v = g->r3.im;
        double A = sqrt(sq(r1 - u) + sq(v));
        double B = sqrt(sq(r2 - u) + sq(v));
        double m2 = (sq(A + B) - sq(r1 - r2)) / (4. * A * B);
        double cn = jacobi_cn(sqrt(A * B) * (g->Rpc - P), m2);
        return (r2 * A - r1 * B - (r2 * A + r1 * B) * cn) / ((A - B) - (A + B) * cn);
    }
    return NAN;
}
S5_HD S5_MID double geodesic_position_rad(const Geodesic* g, double P)
{
    ff::Quick f;
    double r = geodesic_position_rad_t(f, g, P);
    if (!f.ok) { ff::Plain p; r = geodesic_position_rad_t(p, g, P); }
    return r;
}

/* helper shared by the three polar routines: sign of d(mu)/dP at P and the start of the current theta-oscillation.
 * sim5kerr-geod.c:362-407, 412-457, 736-781 */
S5_HD S5_INL bool polar_phase(const Geodesic* g, double P, double* sign_dm, double* T)
{
    if (!(g->type == GEOD_TYPE_RR || g->type == GEOD_TYPE_RC || g->type == GEOD_TYPE_CC)) return false;
    double s = (g->beta >= 0.0) ? +1.0 : -1.0;
    double t = (s > 0.0) ? -(g->Tpp - g->Tip) : -(g->Tip);
    while (P > t + g->Tpp) {
        t += g->Tpp;
        s = -s;
    }
    *sign_dm = s; *T = t;
    return true;
}
S5_HD S5_INL double geodesic_position_pol(const Geodesic* g, double P)
{
    double s, T;
    if (!polar_phase(g, P, &s, &T)) return NAN;
    return -s * theta_inv(g, P - T);
}
S5_HD S5_INL double geodesic_position_pol_sign_k_theta(const Geodesic* g, double P)
{
    double s, T;
    if (!polar_phase(g, P, &s, &T)) return NAN;
    return (s < 0) ? +1 : -1;
}
S5_HD S5_INL double geodesic_dm_sign(const Geodesic* g, double P)
{
    double s, T;
    if (!polar_phase(g, P, &s, &T)) return NAN;
    return s;
}

/* azimuth travelled between infinity and (r, m) at position P.  sim5kerr-geod.c:462-555 */
S5_HD S5_INL double geodesic_position_azm(const Geodesic* g, double r, double m, double P)
{
    double phi = 0.0;
    int ppc = (g->nrr > 0) && (P > g->Rpc);
    double a2 = sq(g->a);
    double rp = 1. + sqrt(1. - a2);
    double rm = 1. - sqrt(1. - a2);
    double r1, r2, r3, r4, A, B;

    if (g->type == GEOD_TYPE_RR) {
        r1 = g->r1.re; r2 = g->r2.re; r3 = g->r3.re; r4 = g->r4.re;
        A = integral_R_rp_re_inf(r1, r2, r3, r4, rp) + (ppc ? +1 : -1) * integral_R_rp_re(r1, r2, r3, r4, rp, r);
        B = integral_R_rp_re_inf(r1, r2, r3, r4, rm) + (ppc ? +1 : -1) * integral_R_rp_re(r1, r2, r3, r4, rm, r);
        phi += 1. / sqrt(1. - a2) * (A * (g->a * rp - g->l * a2 / 2.) - B * (g->a * rm - g->l * a2 / 2.));
    } else if (g->type == GEOD_TYPE_RC) {
        r1 = g->r1.re; r2 = g->r2.re;
        A = integral_R_rp_cc2_inf(r1, r2, g->r3.re, g->r3.im, rp, r);
        B = integral_R_rp_cc2_inf(r1, r2, g->r3.re, g->r3.im, rm, r);
        phi += 1. / sqrt(1. - a2) * (A * (g->a * rp - g->l * a2 / 2.) - B * (g->a * rm - g->l * a2 / 2.));
    } else {
        return NAN;
    }

    double phi_pp = 2.0 * g->l / g->a * integral_T_mp(g->m2m, g->m2p, 1.0, 0.0);
    double phi_ip =       g->l / g->a * integral_T_mp(g->m2m, g->m2p, 1.0, g->cos_i);
    double phi_mp =       g->l / g->a * integral_T_mp(g->m2m, g->m2p, 1.0, m);

    double T;
    double sign_dm = (g->beta >= 0.0) ? +1.0 : -1.0;
    if (sign_dm > 0.0) {
        T = -(g->Tpp - g->Tip);
        phi -= phi_pp - phi_ip;
    } else {
        T = -g->Tip;
        phi -= phi_ip;
    }
    if (P >= T + g->Tpp) {          /* the reference's `while` body ends in `break` */
        T += g->Tpp;
        phi += phi_pp;
        sign_dm = -sign_dm;
    }
    phi += (sign_dm < 0) ? phi_mp : phi_pp - phi_mp;
    return phi;
}

/* sim5kerr-geod.c:786-840 */
S5_HD S5_INL void geodesic_momentum(const Geodesic* g, double P, double r, double m, double k[4])
{
    if ((r == 0.0) && (m == 0.0)) {
        r = geodesic_position_rad(g, P);
        m = geodesic_position_pol(g, P);
    }
    switch (g->type) {
        case GEOD_TYPE_RR:
        case GEOD_TYPE_RC:
        case GEOD_TYPE_CC: {
            double dm = geodesic_dm_sign(g, P);
            photon_momentum(g->a, r, m, g->l, g->q, (P < g->Rpc ? -1 : +1), dm, k);
            return;
        }
        case GEOD_TYPE_RR_DBL:
        case GEOD_TYPE_RR_BH:
            k[0] = k[1] = k[2] = k[3] = NAN;
            return;
    }
}

/* position parameter of the n-th crossing of the equatorial plane.  sim5kerr-geod.c:845-885 */
S5_HD S5_INL double geodesic_find_midplane_crossing(const Geodesic* g, int order)
{
    if (g->q <= 0.0) return NAN;
    double u = g->cos_i / sqrt(g->m2p);
    if (!ensure_range(&u, -1.0, +1.0, 1e-4)) return NAN;
    double pos;
    if (g->beta > 0.0)
        pos = g->mK * ((2. * (double)order + 1.) * elliptic_k(g->mm) + jacobi_icn(u, g->mm));
    else if (g->beta < 0.0)
        pos = g->mK * ((2. * (double)order + 1.) * elliptic_k(g->mm) - jacobi_icn(u, g->mm));
    else
        pos = g->mK * ((2. * (double)order + 1.) * elliptic_k(g->mm));
    if (pos > 2. * g->Rpc) pos = NAN;
    return pos;
}

/* analytic stepping along the geodesic.  sim5kerr-geod.c:890-925 */
S5_HD S5_INL void geodesic_follow(const Geodesic* g, double step, double* P, double* r, double* m, int* status)
{
    const double MAXSTEP_FACTOR = 5e-2;
    do {
        double truestep = step / fabs(step) * fmin(fabs(step), MAXSTEP_FACTOR * sqrt(*r));
        (*P) = (*P) + truestep / (sq(*r) + sq((g->a) * (*m)));
        (*r) = geodesic_position_rad(g, *P);
        (*m) = geodesic_position_pol(g, *P);
        if ((*r) < 1.01 * r_bh(g->a)) { if (status) *status = 0; return; }
        if ((*P < 0.0) || (*P > 2. * g->Rpc)) { if (status) *status = 0; return; }
        step -= truestep;
    } while (fabs(step) > 1e-5);
    if (status) *status = 1;
}

/* geodesic from a local position and direction.  sim5kerr-geod.c:105-173 */
S5_HD S5_INL int geodesic_init_src(double a, double r, double m, const double k[4], int ppc, Geodesic* g, int* error)
{
    double l, q;
    photon_motion_constants(a, r, m, k, &l, &q);
    g->a = fmax(1e-8, a);
    g->l = l;
    g->q = q;
    g->cos_i = g->alpha = g->beta = NAN;
    if (!geodesic_R_roots(g, r, error)) return 0;
    if (!geodesic_T_roots(g, m, error)) return 0;
    if (isnan(g->cos_i) && (r > g->rp)) {
        double T, Tmp, Tpp, sign_dm;
        Tmp = theta_int(g, m);
        Tpp = 2. * theta_int(g, 0.0);
        T = geodesic_P_int(g, r, ppc);
        sign_dm = (k[2] < 0.0) ? +1.0 : -1.0;
        T += (sign_dm > 0.0) ? Tpp - Tmp : Tmp;
        while (T > Tpp) {
            T -= Tpp;
            sign_dm = -sign_dm;
        }
        g->cos_i = -sign_dm * theta_inv(g, T);
        g->incl = cr_acos(g->cos_i);
        g->alpha = -g->l / sqrt(1.0 - sq(g->cos_i));
        g->beta = -sign_dm * sqrt(g->q - sq(g->cos_i) * (sq(g->alpha) - sq(g->a)));
    }
    g->Tpp = 2. * theta_int(g, 0.0);
    g->Tip = theta_int(g, g->cos_i);
    if (error) *error = GD_OK;
    return 1;
}

/* travel time between two positions on the geodesic (the radial integrals only: the reference's polar part is commented
 * out).  sim5kerr-geod.c:559-731 */
S5_HD S5_MID double geodesic_timedelay(const Geodesic* g, double P1, double r1, double m1, double P2, double r2, double m2)
{
    double time = 0.0;
    if (P1 > P2) {
        double tmp;
        tmp = P2; P2 = P1; P1 = tmp;
        tmp = r2; r2 = r1; r1 = tmp;
        tmp = m2; m2 = m1; m1 = tmp;
    }
    if (r1 == 0) { r1 = geodesic_position_rad(g, P1); m1 = geodesic_position_pol(g, P1); }
    if (r2 == 0) { r2 = geodesic_position_rad(g, P2); m2 = geodesic_position_pol(g, P2); }
    (void)m1; (void)m2;
    double a2 = sq(g->a);
    double rp = 1. + sqrt(1. - a2);
    double rm = 1. - sqrt(1. - a2);
    double ra = g->r1.re, rb = g->r2.re, rc = g->r3.re, rd = g->r4.re;
    double R0, R1, R2, RA, RB, A, B, s;
    switch (g->type) {
        case GEOD_TYPE_RR:
            s = (((P1 > g->Rpc) && (P2 < g->Rpc)) || ((P1 < g->Rpc) && (P2 > g->Rpc))) ? +1 : -1;
            R0 = integral_R_r0_re(ra, rb, rc, rd, r1) + s * integral_R_r0_re(ra, rb, rc, rd, r2);
            R1 = integral_R_r1_re(ra, rb, rc, rd, r1) + s * integral_R_r1_re(ra, rb, rc, rd, r2);
            R2 = integral_R_r2_re(ra, rb, rc, rd, r1) + s * integral_R_r2_re(ra, rb, rc, rd, r2);
            RA = integral_R_rp_re(ra, rb, rc, rd, rp, r1) + s * integral_R_rp_re(ra, rb, rc, rd, rp, r2);
            RB = integral_R_rp_re(ra, rb, rc, rd, rm, r1) + s * integral_R_rp_re(ra, rb, rc, rd, rm, r2);
            A = (-g->a * g->l + 4.) * rp - 2. * a2;
            B = (+g->a * g->l - 4.) * rm + 2. * a2;
            time += 4. * fabs(R0) + 2. * fabs(R1) + fabs(R2) + (A * fabs(RA) + B * fabs(RB)) / sqrt(1. - a2);
            break;
        case GEOD_TYPE_RC: {
            double cre = g->r3.re, cim = g->r3.im;
            R0 = integral_R_r0_cc(ra, rb, cre, cim, r1) - integral_R_r0_cc(ra, rb, cre, cim, r2);
            R1 = (r1 < r2) ? integral_R_r1_cc(ra, rb, cre, cim, r1, r2) : integral_R_r1_cc(ra, rb, cre, cim, r2, r1);
            R2 = (r1 < r2) ? integral_R_r2_cc(ra, rb, cre, cim, r1, r2) : integral_R_r2_cc(ra, rb, cre, cim, r2, r1);
            RA = (r1 < r2) ? integral_R_rp_cc2(ra, rb, cre, cim, rp, r1, r2) : integral_R_rp_cc2(ra, rb, cre, cim, rp, r2, r1);
            RB = (r1 < r2) ? integral_R_rp_cc2(ra, rb, cre, cim, rm, r1, r2) : integral_R_rp_cc2(ra, rb, cre, cim, rm, r2, r1);
            A = (-g->a * g->l + 4.) * rp - 2. * a2;
            B = (+g->a * g->l - 4.) * rm + 2. * a2;
            time += 4. * fabs(R0) + 2. * fabs(R1) + fabs(R2) + (A * fabs(RA) + B * fabs(RB)) / sqrt(1. - a2);
            break;
        }
        default:                            /* RR_DBL, RR_BH, CC: not implemented in the reference either */
            return NAN;
    }
    return time;
}

} /* namespace s5 */
#endif
