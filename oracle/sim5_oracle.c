/*
 * sim5_oracle.c -- CPU restatement ("port") of the reference's per-pixel photon path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/libsim5oracle.so by oracle/Makefile; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (sim5_b200/libsim5b200.so) never links, loads or calls anything in this file and has no CPU path.
 *
 * PARITY PINNING: the reference's own tests hold no golden vectors for this path, so this
 * restatement is pinned against OUTPUTS OF THE UNMODIFIED REFERENCE:
 *   - the .npz fixtures under tests/golden (written by tools/make_golden.py from oracle/_ref = the reference sources
 *     compiled where they lie); tests/test_oracle.py requires BIT-IDENTICAL planes and status bytes
 *     (both sides are the same call-for-call arithmetic on the same glibc; observed: 0 differing doubles);
 *   - oracle/_ref itself, live, whenever it is built (same test file, larger images).
 *
 * Scope: the analytic path of BASELINE configs 1, 2, 3 and 5 (modes EQPLANE, POLARIZED, HISTOGRAM) and the SPECTRUM mode.  The
 * step-wise integrator (config 4) is checked against oracle/_ref and its golden fixture only;
 * orc_trace_image returns -3 for SIM5_MODE_STEPWISE.
 *
 * This is the algorithm AS THE REFERENCE WRITES IT (call for call, with its repeated Carlson
 * evaluations and its use of the platform libm) -- not the fused form the CUDA kernels use.  Every
 * function names the reference lines it follows (paths relative to the reference's src/).
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sim5_b200.h"
#include "sim5_oracle.h"

#define SQ(x) ((x) * (x))
static const double TOL_DUP = 0.0003;           /* ERRTOL of all four Carlson routines */
static const double HALF_PI_TRUNC = 1.57079632679;   /* sim5math.h:39 (PI_half, truncated) */

/* ------------------------------------------------------------------------------------------ */
/* Carlson integrals by duplication (Numerical Recipes form).  sim5elliptic.c:18-206            */
/* ------------------------------------------------------------------------------------------ */

/* R_F.  sim5elliptic.c:18-52 (the argument check only prints) */
double orc_rf(double x, double y, double z)
{
    double mu, dx, dy, dz;
    for (;;) {
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
        double lam = sx * (sy + sz) + sy * sz;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        z = 0.25 * (z + lam);
        mu = (1.0 / 3.0) * (x + y + z);
        dx = (mu - x) / mu;
        dy = (mu - y) / mu;
        dz = (mu - z) / mu;
        if (!(fmax(fmax(fabs(dx), fabs(dy)), fabs(dz)) > TOL_DUP)) break;
    }
    double e2 = dx * dy - dz * dz;
    double e3 = dx * dy * dz;
    return (1.0 + ((1.0 / 24.0) * e2 - 0.1 - (3.0 / 44.0) * e3) * e2 + (1.0 / 14.0) * e3) / sqrt(mu);
}

/* R_D.  sim5elliptic.c:58-98 */
double orc_rd(double x, double y, double z)
{
    const double k1 = 3.0 / 14.0, k2 = 1.0 / 6.0, k3 = 9.0 / 22.0, k4 = 3.0 / 26.0, k5 = 0.25 * k3, k6 = 1.5 * k4;
    double acc = 0.0, w = 1.0, mu, dx, dy, dz;
    for (;;) {
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
        double lam = sx * (sy + sz) + sy * sz;
        acc += w / (sz * (z + lam));
        w = 0.25 * w;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        z = 0.25 * (z + lam);
        mu = 0.2 * (x + y + 3.0 * z);
        dx = (mu - x) / mu;
        dy = (mu - y) / mu;
        dz = (mu - z) / mu;
        if (!(fmax(fmax(fabs(dx), fabs(dy)), fabs(dz)) > TOL_DUP)) break;
    }
    double ea = dx * dy, eb = dz * dz, ec = ea - eb, ed = ea - 6.0 * eb, ee = ed + ec + ec;
    return 3.0 * acc + w * (1.0 + ed * (-k1 + k5 * ed - k6 * dz * ee) + dz * (k2 * ee + dz * (-k3 * ec + dz * k4 * ea))) / (mu * sqrt(mu));
}

/* R_C, Cauchy principal value for y < 0.  sim5elliptic.c:104-137 */
double orc_rc(double x, double y)
{
    double w, mu, s;
    if (y > 0.0) {
        w = 1.0;
    } else {
        double x0 = x;
        x = x - y;
        y = -y;
        w = sqrt(x0) / sqrt(x);
    }
    for (;;) {
        double lam = 2.0 * sqrt(x) * sqrt(y) + y;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        mu = (1.0 / 3.0) * (x + y + y);
        s = (y - mu) / mu;
        if (!(fabs(s) > TOL_DUP)) break;
    }
    return w * (1.0 + s * s * (0.3 + s * ((1.0 / 7.0) + s * (0.375 + s * (9.0 / 22.0))))) / sqrt(mu);
}

/* R_J, Cauchy principal value for p < 0; 0 for arguments outside its range.  sim5elliptic.c:144-206 */
double orc_rj(double x, double y, double z, double p)
{
    const double tiny = pow(5.0 * DBL_MIN, 1. / 3.), big = 0.3 * pow(0.1 * DBL_MAX, 1. / 3.);
    const double k1 = 3.0 / 14.0, k2 = 1.0 / 3.0, k3 = 3.0 / 22.0, k4 = 3.0 / 26.0, k5 = 0.75 * k3, k6 = 1.5 * k4, k7 = 0.5 * k2, k8 = k3 + k3;
    if ((fmin(fmin(x, y), z) < 0.0) || (fmin(fmin(x + y, x + z), fmin(y + z, fabs(p))) < tiny) ||
        (fmax(fmax(x, y), fmax(z, fabs(p))) > big))
        return 0.0;
    double a = 0.0, b = 0.0, rcx = 0.0, xt, yt, zt, pt;
    if (p > 0.0) {
        xt = x; yt = y; zt = z; pt = p;
    } else {
        xt = fmin(fmin(x, y), z);
        zt = fmax(fmax(x, y), z);
        yt = x + y + z - xt - zt;
        a = 1.0 / (yt - p);
        b = a * (zt - yt) * (yt - xt);
        pt = yt + b;
        double rho = xt * zt / yt;
        double tau = p * pt / yt;
        rcx = orc_rc(rho, tau);
    }
    double acc = 0.0, w = 1.0, mu, dx, dy, dz, dp;
    for (;;) {
        double sx = sqrt(xt), sy = sqrt(yt), sz = sqrt(zt);
        double lam = sx * (sy + sz) + sy * sz;
        double al = SQ(pt * (sx + sy + sz) + sx * sy * sz);
        double be = pt * SQ(pt + lam);
        acc += w * orc_rc(al, be);
        w = 0.25 * w;
        xt = 0.25 * (xt + lam);
        yt = 0.25 * (yt + lam);
        zt = 0.25 * (zt + lam);
        pt = 0.25 * (pt + lam);
        mu = 0.2 * (xt + yt + zt + pt + pt);
        dx = (mu - xt) / mu;
        dy = (mu - yt) / mu;
        dz = (mu - zt) / mu;
        dp = (mu - pt) / mu;
        if (!(fmax(fmax(fabs(dx), fabs(dy)), fmax(fabs(dz), fabs(dp))) > TOL_DUP)) break;
    }
    double ea = dx * (dy + dz) + dy * dz;
    double eb = dx * dy * dz;
    double ec = dp * dp;
    double ed = ea - 3.0 * ec;
    double ee = eb + 2.0 * dp * (ea - ec);
    double v = 3.0 * acc + w * (1.0 + ed * (-k1 + k5 * ed - k6 * ee) + eb * (k7 + dp * (-k8 + dp * k4)) + dp * ea * (k2 - dp * k3) - k2 * dp * ec) / (mu * sqrt(mu));
    if (p <= 0.0) v = a * (b * v + 3.0 * (rcx - orc_rf(xt, yt, zt)));
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* Legendre / Jacobi forms.  sim5elliptic.c:217-630                                             */
/* ------------------------------------------------------------------------------------------ */

/* K(m).  sim5elliptic.c:217-225 */
static double ell_K(double m)
{
    if (m == 1.0) m = 1.0 - 1e-8;
    return orc_rf(0, 1.0 - m, 1.0);
}
/* F from sin(phi).  sim5elliptic.c:273-284 */
static double ell_F_sin(double s, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (s == 0.0) return 0.0;
    double s2 = SQ(s);
    return s * orc_rf(1. - s2, 1.0 - s2 * m, 1.0);
}
/* F from cos(phi).  sim5elliptic.c:254-271 */
static double ell_F_cos(double c, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (c == 1.0) return 0.0;
    double X = 0.0;
    if (c < 0.0) {
        c = -c;
        X = 2.0 * orc_rf(0.0, 1.0 - m, 1.0);
    }
    double s2 = 1.0 - SQ(c);
    return X + ((X == 0.0) ? (+1) : (-1)) * sqrt(s2) * orc_rf(1.0 - s2, 1.0 - s2 * m, 1.0);
}
/* complete Pi(n, m).  sim5elliptic.c:365-378 */
static double ell_Pi_complete(double n, double m)
{
    if (isinf(n)) return 0.0;
    if (m == 1.0) m = 0.99999999;
    if (n == 1.0) n = 0.99999999;
    double q = 1.0 - m;
    return orc_rf(0.0, q, 1.0) + n * orc_rj(0.0, q, 1.0, 1.0 - n) / 3.0;
}
/* Pi from cos(phi).  sim5elliptic.c:425-450 */
static double ell_Pi_cos(double c, double n, double m)
{
    if (isinf(n)) return 0.0;
    if (c == 1.0) return 0.0;
    if (c == 0.0) return ell_Pi_complete(n, m);
    if (m == 1.0) m = 0.99999999;
    double X = 0.0;
    if (c < 0.0) {
        c = -c;
        X = 2.0 * ((orc_rf(0.0, 1.0 - m, 1.0) + n * orc_rj(0.0, 1.0 - m, 1.0, 1.0 - n) / 3.0));
    }
    double c2 = SQ(c);
    double s = sqrt(1.0 - c2);
    double ns2 = -n * (1.0 - c2);
    double q = 1.0 - (1.0 - c2) * m;
    return X + ((X == 0.0) ? (+1) : (-1)) * s * (orc_rf(c2, q, 1.0) - ns2 * orc_rj(c2, q, 1.0, 1.0 + ns2) / 3.0);
}
/* sn^-1.  sim5elliptic.c:480-486 */
static double jac_isn(double z, double m)
{
    if (fabs(m - 0.0) < 1e-8) return asin(z);
    if (fabs(m - 1.0) < 1e-8) return log(sqrt((1. + z) / (1. - z)));
    return z * orc_rf(1.0 - z * z, 1.0 - m * z * z, 1.0);
}
/* cn^-1 (negative z through the imaginary-modulus formula).  sim5elliptic.c:492-514 */
static double jac_icn(double z, double m)
{
    if ((z > +1.0) && (z < +1.0 + 1e-8)) z = +1.0;
    if ((z < -1.0) && (z > -1.0 - 1e-8)) z = -1.0;
    if ((m > +1.0) && (m < +1.0 + 1e-8)) m = 1.0;
    if ((m < 0.0) && (m > 0.0 - 1e-8)) m = 0.0;
    if (z == 0.0) return ell_K(m);
    if (z == 1.0) return 0.0;
    if (m == 0.0) return acos(z);
    if (m == 1.0) return log((1. + sqrt(1. - z)) / z);
    double pos = sqrt(1. - z * z) * orc_rf(z * z, 1.0 - m * (1. - z * z), 1.0);
    return (z > 0.0) ? pos : 2. / sqrt(1. - m) * ell_F_sin(-z, m / (m - 1.)) + pos;
}
/* tn^-1.  sim5elliptic.c:522-528 */
static double jac_itn(double z, double m)
{
    if (m == 0.0) return atan(z);
    if (m == 1.0) return log(z + sqrt(1. + z * z));
    return jac_isn(sqrt(z * z / (1. + z * z)), m);
}
/* sn, cn, dn by the descending Landen / AGM scheme.  sim5elliptic.c:535-606 */
void orc_sncndn(double u, double m, double* sn, double* cn, double* dn)
{
    if (m == 1.0) m = 0.999999999;
    const double agm_tol = 1.0e-8;
    double lev_a[13], lev_g[13];
    double a, b, c = 0.0, d = 1.0;
    double mc = 1.0 - m;
    int flip, i, last = 0;
    if (mc == 0.0) {
        *cn = 1.0 / cosh(u);
        *dn = *cn;
        *sn = tanh(u);
        return;
    }
    flip = (mc < 0.0);
    if (flip) {
        d = 1.0 - mc;
        mc /= -1.0 / d;
        u *= (d = sqrt(d));
    }
    a = 1.0;
    *dn = 1.0;
    for (i = 0; i < 13; i++) {
        last = i;
        lev_a[i] = a;
        lev_g[i] = (mc = sqrt(mc));
        c = 0.5 * (a + mc);
        if (fabs(a - mc) <= agm_tol * a) break;
        mc *= a;
        a = c;
    }
    u *= c;
    *sn = sin(u);
    *cn = cos(u);
    if (*sn != 0.0) {
        a = (*cn) / (*sn);
        c *= a;
        for (i = last; i >= 0; i--) {
            b = lev_a[i];
            a *= c;
            c *= *dn;
            *dn = (lev_g[i] + a) / (b + a);
            a = c / b;
        }
        a = 1.0 / sqrt(c * c + 1.0);
        *sn = ((*sn) >= 0.0 ? a : -a);
        *cn = c * (*sn);
    }
    if (flip) {
        a = *dn;
        *dn = *cn;
        *cn = a;
        *sn /= d;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Byrd & Friedman integrals used by the azimuth.  sim5elliptic.c:676-690, 755-792, 1017-1159    */
/* ------------------------------------------------------------------------------------------ */

/* B&F 340.01.  sim5elliptic.c:676-690 */
static double bf_Z1(double a, double b, double u, double m)
{
    double sn, cn, dn;
    orc_sncndn(u, m, &sn, &cn, &dn);
    return 1. / a * ((a - b) * ell_Pi_cos(cn, a, m) + b * u);
}
/* B&F 341.03 / 361.54; the complex f1 term uses the platform csqrt / catan.  sim5elliptic.c:755-792 */
static double bf_R1(double a, double u, double m)
{
    double a2 = SQ(a);
    double n = a2 / (a2 - 1.);
    double sn, cn, dn;
    orc_sncndn(u, m, &sn, &cn, &dn);
    double mma = (m + (1. - m) * a2) / (1. - a2);
    double complex f1 = (fabs(mma) > 1e-5) ? csqrt(CMPLX(1. / mma, 0.0)) * catan(csqrt(CMPLX(mma, 0.0)) * sn / dn) : CMPLX(sn / dn, 0.0);
    double complex ellpi = ell_Pi_cos(cn, n, m);
    double complex res = 1. / (1. - a2) * (ellpi + a * f1);
    return creal(res);
}
/* B&F 258.39, four real roots a>b>c>d, upper limit X.  sim5elliptic.c:1017-1029 */
static double bf_R_pole_rr(double a, double b, double c, double d, double p, double X)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double s = sqrt(((b - d) * (X - a)) / ((a - d) * (X - b)));
    double u1 = jac_isn(s, m2);
    double a2 = (a - d) / (b - d);
    double c2 = ((p - b) * (a - d)) / ((p - a) * (b - d));
    return -2.0 / sqrt((a - c) * (b - d)) / (p - a) * (bf_Z1(c2, a2, u1, m2) - bf_Z1(c2, a2, 0.0, m2));
}
/* the same with X -> infinity.  sim5elliptic.c:1032-1044 */
static double bf_R_pole_rr_inf(double a, double b, double c, double d, double p)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double s = sqrt((b - d) / (a - d));
    double u1 = jac_isn(s, m2);
    double a2 = (a - d) / (b - d);
    double c2 = ((p - b) * (a - d)) / ((p - a) * (b - d));
    return -2.0 / sqrt((a - c) * (b - d)) / (p - a) * (bf_Z1(c2, a2, u1, m2) - bf_Z1(c2, a2, 0.0, m2));
}
/* B&F 260.04: two real roots a>b and the pair c, c*, from X1 to infinity.  sim5elliptic.c:1081-1112 */
static double bf_R_pole_rc_inf(double a, double b, double complex c, double p, double X1)
{
    double u = creal(c);
    double v2 = SQ(cimag(c));
    double A = sqrt(SQ(a - u) + v2);
    double B = sqrt(SQ(b - u) + v2);
    double m = (SQ(A + B) - SQ(a - b)) / (4. * A * B);
    double g = 1. / sqrt(A * B);
    double al1 = (B * a + b * A - p * A - p * B) / (B * a - b * A + p * A - p * B);
    double al2 = (B + A) / (B - A);
    double u1 = ell_F_cos((X1 * (A - B) + a * B - b * A) / (X1 * (A + B) - a * B - b * A), m);
    double u2 = ell_F_cos((A - B) / (A + B), m);
    double t0 = al2 * (u2 - u1);
    double t1 = (al1 - al2) * (bf_R1(al1, u2, m) - bf_R1(al1, u1, m));
    return (B - A) * g / (B * a + b * A - p * A - p * B) * (t0 + t1);
}
/* B&F 213.02.  sim5elliptic.c:1142-1159 */
static double bf_T_pole(double a2, double b2, double p, double X)
{
    double m = b2 / (a2 + b2);
    double n = b2 / (b2 - p);
    if (X >= 0.0)
        return 1. / sqrt(a2 + b2) / (p - b2) * ell_Pi_cos(X / sqrt(b2), n, m);
    return 1. / sqrt(a2 + b2) / (p - b2) * (2. * ell_Pi_complete(n, m) - ell_Pi_cos(-X / sqrt(b2), n, m));
}

/* ------------------------------------------------------------------------------------------ */
/* analytic Kerr geodesic.  sim5kerr-geod.c                                                     */
/* ------------------------------------------------------------------------------------------ */
enum { TY_RR = 40, TY_RR_DBL = 41, TY_RR_BH = 42, TY_RC = 2, TY_CC = 0 };          /* sim5kerr-geod.h:19-23 */
enum { E_OK = 0, E_UNKNOWN = 3, E_RR_DOUBLE = 4, E_Q_RANGE = 7, E_MUPLUS = 8, E_MU0 = 9, E_MM = 10, E_INCL = 11, E_SPIN = 12 };   /* :26-37 */

typedef struct orc_ray {          /* the fields of `geodesic` (sim5kerr-geod.h:42-68) this path reads */
    double a, alpha, beta, cos_i, l, q;
    double complex root[4];
    int nreal, type;
    double m2p, m2m, mm, mK, rp, Rpc, Tpp, Tip;
} orc_ray;

/* sim5math.c:49-58 */
static int clamp_into(double* v, double lo, double hi, double slack)
{
    if (*v < lo - slack) return 0;
    if (*v > hi + slack) return 0;
    if (*v < lo) *v = lo;
    if (*v > hi) *v = hi;
    return 1;
}

/* real roots first (descending), complex ones after in input order.  sim5polyroots.c:277-324 */
static void order_roots(orc_ray* g)
{
    double complex in[4], out[4];
    int i, j, k, nr = 0;
    for (i = 0; i < 4; i++) in[i] = g->root[i];
    for (i = 0; i < 4; i++) if (cimag(in[i]) == 0.) out[nr++] = in[i];
    k = nr;
    for (i = 0; i < 4; i++) if (cimag(in[i]) != 0.) out[k++] = in[i];
    for (i = 0; i < nr; i++)
        for (j = 0; j < nr - i; j++)
            if (creal(out[i + j]) > creal(out[i])) { double complex t = out[i + j]; out[i + j] = out[i]; out[i] = t; }
    for (i = 0; i < 4; i++) g->root[i] = out[i];
    g->nreal = nr;
}

/* quartic roots of R(r) by the Cadez et al. resolvent, classification, pericentre, R-integral to it.
 * sim5kerr-geod.c:985-1104 */
static int radial_roots(orc_ray* g, double r0, int* err)
{
    double a = g->a, l = g->l, q = g->q;
    double a2 = SQ(a), l2 = SQ(l);
    double A, B, C, D, E, F, X;
    C = SQ(a - l) + q;
    D = 2. / 3. * (q + l2 - a2);
    E = 9. / 4. * SQ(D) - 12. * a2 * q;
    F = -27. / 4. * (D * D * D) - 108. * a2 * q * D + 108. * SQ(C);
    X = SQ(F) - 4. * (E * E * E);
    if (X >= 0) {
        A = (F > sqrt(X) ? +1 : -1) * 1. / 3. * pow(fabs(F - sqrt(X)) / 2., 1. / 3.) + (F > -sqrt(X) ? +1 : -1) * 1. / 3. * pow(fabs(F + sqrt(X)) / 2., 1. / 3.);
    } else {
        double Z = sqrt(pow(F / 54., 2) + pow(sqrt(-X) / 54., 2));
        double z = atan2(sqrt(-X) / 54., F / 54.);
        A = pow(Z, 1. / 3.) * 2. * cos(z / 3.);
    }
    B = sqrt(A + D);
    g->root[0] = +B / 2. + .5 * csqrt(CMPLX(-A + 2. * D - 4. * C / B, 0.0));
    g->root[1] = +B / 2. - .5 * csqrt(CMPLX(-A + 2. * D - 4. * C / B, 0.0));
    g->root[2] = -B / 2. + .5 * csqrt(CMPLX(-A + 2. * D + 4. * C / B, 0.0));
    g->root[3] = -B / 2. - .5 * csqrt(CMPLX(-A + 2. * D + 4. * C / B, 0.0));
    order_roots(g);

    double r1 = creal(g->root[0]), r2 = creal(g->root[1]), r3 = creal(g->root[2]), r4 = creal(g->root[3]);
    switch (g->nreal) {
        case 4:
            g->type = TY_RR;
            if ((r0 < r3) || ((r0 > r2) && (r0 < r1))) { *err = E_UNKNOWN; return 0; }
            if (fabs(r1 - r2) < 1e-8) { g->type = TY_RR_DBL; *err = E_RR_DOUBLE; return 0; }
            if ((r0 >= r3) && (r0 <= r2)) g->type = TY_RR_BH;
            break;
        case 2: g->type = TY_RC; break;
        case 0: g->type = TY_CC; break;
        default: *err = E_UNKNOWN; return 0;
    }
    double mm, u, v;
    switch (g->type) {
        case TY_RR:
            mm = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
            g->rp = r1;
            g->Rpc = 2. / sqrt((r1 - r3) * (r2 - r4)) * jac_isn(sqrt((r2 - r4) / (r1 - r4)), mm);
            break;
        case TY_RR_BH:
            mm = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
            g->rp = r2;
            g->Rpc = 2. / sqrt((r1 - r3) * (r2 - r4)) * ell_K(mm);
            break;
        case TY_RC:
            u = creal(g->root[2]); v = cimag(g->root[2]);
            A = sqrt(SQ(r1 - u) + SQ(v));
            B = sqrt(SQ(r2 - u) + SQ(v));
            mm = (SQ(A + B) - SQ(r1 - r2)) / (4. * A * B);
            g->rp = r1;
            g->Rpc = 1. / sqrt(A * B) * jac_icn((A - B) / (A + B), mm);
            break;
        default: {        /* CC: b1 = Re r1, b2 = Re r3, a1 = Im r1, a2 = Im r3 */
            double b1 = creal(g->root[0]), b2 = creal(g->root[2]), i1 = cimag(g->root[0]), i2 = cimag(g->root[2]);
            A = sqrt(SQ(b1 - b2) + SQ(i1 + i2));
            B = sqrt(SQ(b1 - b2) + SQ(i1 - i2));
            double g1 = sqrt((4. * SQ(i1) - SQ(A - B)) / (SQ(A + B) - 4. * SQ(i1)));
            mm = 4. * A * B / SQ(A + B);
            g->rp = b1 - i1 * g1;
            g->Rpc = 2. / (A + B) * jac_itn(-1. / g1, mm);
            break;
        }
    }
    return 1;
}

/* roots of Theta(mu) through the product identity, in x87 extended precision like the CPU build.
 * sim5kerr-geod.c:1109-1184 */
static int polar_roots(orc_ray* g, double m, int* err)
{
    double a2 = SQ(g->a), l2 = SQ(g->l), q = g->q;
    long double qla = q + l2 - a2;
    long double X = sqrt(SQ(qla) + 4. * q * a2) + qla;
    long double dbla = a2 + a2;
    long double dblq = q + q;
    g->m2m = X / dbla;
    g->m2p = dblq / X;
    if ((g->m2p <= 0.0) || (g->m2p >= 1.0)) { *err = E_MUPLUS; return 0; }
    if (q > 0.0) {
        g->mm = g->m2p / (g->m2p + g->m2m);
        if ((g->mm < 0.0) || (g->mm >= 1.0)) { *err = E_MM; return 0; }
        if (fabs(m) > sqrt(g->m2p)) { *err = E_MU0; return 0; }
        g->mK = 1. / sqrt(a2 * (g->m2p + g->m2m));
    } else if (q < 0.0) {
        g->mm = (g->m2p + g->m2m) / g->m2p;
        if ((g->mm < 0.0) || (g->mm >= 1.0)) { *err = E_MM; return 0; }
        if ((fabs(m) > sqrt(g->m2p)) || (fabs(m) < sqrt(-g->m2m))) { *err = E_MU0; return 0; }
        g->mK = 1. / sqrt(a2 * g->m2p);
    } else {
        *err = E_Q_RANGE;
        return 0;
    }
    return 1;
}

/* theta_int of sim5kerr-geod.c (static helper of the reference: mK * cn^-1(x / sqrt(m2p), mm)) */
static double polar_integral(const orc_ray* g, double x) { return g->mK * jac_icn(x / sqrt(g->m2p), g->mm); }

/* geodesic from impact parameters at infinity.  sim5kerr-geod.c:41-100 */
static int ray_from_infinity(double i, double a, double alpha, double beta, orc_ray* g, int* err)
{
    if ((a < 0.0) || (a > 1. - 1e-6)) { *err = E_SPIN; return 0; }
    if ((i <= 0.0) || (i >= HALF_PI_TRUNC)) { *err = E_INCL; return 0; }
    if (beta == 0.0) beta = +1e-6;
    g->a = fmax(1e-4, a);
    g->cos_i = cos(i);
    g->alpha = alpha;
    g->beta = beta;
    g->l = -alpha * sin(i);
    g->q = SQ(beta) + SQ(cos(i)) * (SQ(alpha) - SQ(a));
    if (g->q == 0.0) { *err = E_Q_RANGE; return 0; }
    if (!radial_roots(g, DBL_MAX, err)) return 0;
    if (!polar_roots(g, g->cos_i, err)) return 0;
    g->Tpp = 2. * polar_integral(g, 0.0);
    g->Tip = polar_integral(g, g->cos_i);
    *err = E_OK;
    return 1;
}

/* position parameter of the n-th equatorial crossing.  sim5kerr-geod.c:845-885 */
static double midplane_crossing(const orc_ray* g, int order)
{
    if (g->q <= 0.0) return NAN;
    double u = g->cos_i / sqrt(g->m2p);
    if (!clamp_into(&u, -1.0, +1.0, 1e-4)) return NAN;
    double pos;
    if (g->beta > 0.0)      pos = g->mK * ((2. * (double)order + 1.) * ell_K(g->mm) + jac_icn(u, g->mm));
    else if (g->beta < 0.0) pos = g->mK * ((2. * (double)order + 1.) * ell_K(g->mm) - jac_icn(u, g->mm));
    else                    pos = g->mK * ((2. * (double)order + 1.) * ell_K(g->mm));
    if (pos > 2. * g->Rpc) pos = NAN;
    return pos;
}

/* P -> r.  sim5kerr-geod.c:290-357 */
static double radius_at(const orc_ray* g, double P)
{
    if ((P <= 0.0) || (P >= 2. * g->Rpc)) return NAN;
    if (P == g->Rpc) return g->rp;
    double sn, cn, dn;
    if (g->type == TY_RR) {
        double r1 = creal(g->root[0]), r2 = creal(g->root[1]), r3 = creal(g->root[2]), r4 = creal(g->root[3]);
        double m4 = ((r2 - r3) * (r1 - r4)) / ((r2 - r4) * (r1 - r3));
        double x4 = 0.5 * fabs(P - g->Rpc) * sqrt((r2 - r4) * (r1 - r3));
        orc_sncndn(x4, m4, &sn, &cn, &dn);
        double sn2 = pow(sn, 2.0);
        return (r1 * (r2 - r4) - r2 * (r1 - r4) * sn2) / (r2 - r4 - (r1 - r4) * sn2);
    }
    if (g->type == TY_RC) {
        if (P > g->Rpc) return NAN;       /* unphysical second branch */
        double r1 = creal(g->root[0]), r2 = creal(g->root[1]), u = creal(g->root[2]), v = cimag(g->root[2]);
        double A = sqrt(SQ(r1 - u) + SQ(v));
        double B = sqrt(SQ(r2 - u) + SQ(v));
        double m2 = (SQ(A + B) - SQ(r1 - r2)) / (4. * A * B);
        orc_sncndn(sqrt(A * B) * (g->Rpc - P), m2, &sn, &cn, &dn);
        return (r2 * A - r1 * B - (r2 * A + r1 * B) * cn) / ((A - B) - (A + B) * cn);
    }
    return NAN;
}

/* azimuth travelled from infinity to (r, m) at P.  sim5kerr-geod.c:462-555 */
static double azimuth_at(const orc_ray* g, double r, double m, double P)
{
    double phi = 0.0;
    int past_peri = (g->nreal > 0) && (P > g->Rpc);
    double a2 = SQ(g->a);
    double rp = 1. + sqrt(1. - a2);
    double rm = 1. - sqrt(1. - a2);
    double A, B;
    if (g->type == TY_RR) {
        double r1 = creal(g->root[0]), r2 = creal(g->root[1]), r3 = creal(g->root[2]), r4 = creal(g->root[3]);
        A = bf_R_pole_rr_inf(r1, r2, r3, r4, rp) + (past_peri ? +1 : -1) * bf_R_pole_rr(r1, r2, r3, r4, rp, r);
        B = bf_R_pole_rr_inf(r1, r2, r3, r4, rm) + (past_peri ? +1 : -1) * bf_R_pole_rr(r1, r2, r3, r4, rm, r);
        phi += 1. / sqrt(1. - a2) * (A * (g->a * rp - g->l * a2 / 2.) - B * (g->a * rm - g->l * a2 / 2.));
    } else if (g->type == TY_RC) {
        double r1 = creal(g->root[0]), r2 = creal(g->root[1]);
        A = bf_R_pole_rc_inf(r1, r2, g->root[2], rp, r);
        B = bf_R_pole_rc_inf(r1, r2, g->root[2], rm, r);
        phi += 1. / sqrt(1. - a2) * (A * (g->a * rp - g->l * a2 / 2.) - B * (g->a * rm - g->l * a2 / 2.));
    } else {
        return NAN;
    }
    double phi_pp = 2.0 * g->l / g->a * bf_T_pole(g->m2m, g->m2p, 1.0, 0.0);
    double phi_ip =       g->l / g->a * bf_T_pole(g->m2m, g->m2p, 1.0, g->cos_i);
    double phi_mp =       g->l / g->a * bf_T_pole(g->m2m, g->m2p, 1.0, m);
    double T;
    double dm = (g->beta >= 0.0) ? +1.0 : -1.0;
    if (dm > 0.0) {
        T = -(g->Tpp - g->Tip);
        phi -= phi_pp - phi_ip;
    } else {
        T = -g->Tip;
        phi -= phi_ip;
    }
    while (P >= T + g->Tpp) {           /* the reference's loop body ends in `break` */
        T += g->Tpp;
        phi += phi_pp;
        dm = -dm;
        break;
    }
    phi += (dm < 0) ? phi_mp : phi_pp - phi_mp;
    return phi;
}

/* ------------------------------------------------------------------------------------------ */
/* Kerr metric helpers, frames, disk.  sim5kerr.c, sim5disk-nt.c, sim5polarization.c             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double a, r, m, tt, rr, hh, ff, tf; } orc_metric;      /* sim5kerr.h:18-26 */

/* sim5kerr.c:993-1004 (sqrt3 is cbrt, sim5math.h:45) */
double orc_r_ms(double a)
{
    double z1 = 1. + cbrt(1. - a * a) * (cbrt(1. + a) + cbrt(1. - a));
    double z2 = sqrt(3. * (a * a) + (z1 * z1));
    return 3. + z2 - sqrt((3. - z1) * (3. + z1 + 2. * z2));
}
/* sim5kerr.c:74-101 */
static void metric_at(double a, double r, double m, orc_metric* g)
{
    double r2 = SQ(r), a2 = SQ(a), m2 = SQ(m);
    double S = r2 + a2 * m2;
    double s2S = (1.0 - m2) / S;
    g->a = a; g->r = r; g->m = m;
    g->tt = -1. + 2.0 * r / S;
    g->rr = S / (r2 - 2. * r + a2);
    g->hh = S;
    g->ff = ((a2 + r2) * S + 2. * r * a2 * s2S * S) * s2S;
    g->tf = -2. * a * r * s2S;
}
/* sim5kerr.c:608-625 */
static double dot4(const double A[4], const double B[4], const orc_metric* g)
{
    return A[0] * B[0] * g->tt + A[1] * B[1] * g->rr + A[2] * B[2] * g->hh + A[3] * B[3] * g->ff + A[0] * B[3] * g->tf + A[3] * B[0] * g->tf;
}
/* sim5kerr.c:1036-1046 */
static double kepler_omega(double r, double a) { return 1. / (a + pow(r, 1.5)); }
/* sim5kerr.c:1127-1141 */
static double g_kepler(double r, double a, double l)
{
    double Om = 1. / (a + pow(r, 1.5));
    return sqrt(1. - 2. / r * SQ(1. - a * Om) - (r * r + a * a) * SQ(Om)) / (1. - Om * l);
}
/* k^mu from the constants of motion (CPU branch).  sim5kerr.c:1150-1213 */
static void momentum_from_constants(double a, double r, double m, double l, double q, double rs, double ms, double k[4])
{
    double a2 = SQ(a), l2 = SQ(l), r2 = SQ(r), m2 = SQ(m);
    double S = r2 + a2 * m2;
    double D = r2 - 2. * r + a2;
    double R = SQ(r2 + a2 - a * l) - D * (SQ(l - a) + q);
    double M = q - l2 * m2 / (1. - m2) + a2 * m2;
    if ((M < 0.0) && (-M < 1e-8)) M = 0.0;
    if ((R < 0.0) && (-R < 1e-8)) R = 0.0;
    if (M < 0.0) { k[0] = k[1] = k[2] = k[3] = NAN; return; }
    k[0] = +1 / S * (-a * (a * (1. - m2) - l) + (r2 + a2) / D * (r2 + a2 - a * l));
    k[1] = +1 / S * sqrt(R);
    k[2] = +1 / S * sqrt(M);
    k[3] = +1 / S * (-a + l / (1. - m2) + a / D * (r2 + a2 - a * l));
    if (rs < 0.0) k[1] = -k[1];
    if (ms < 0.0) k[2] = -k[2];
}
/* frame of an observer on a circular orbit with angular velocity Omega (Omega != 0 here).  sim5kerr.c:765-813 */
static void frame_azimuthal(const orc_metric* g, double Om, double e[4][4])
{
    double tt = g->tt, ff = g->ff, tf = g->tf;
    double U0 = sqrt(-1.0 / (tt + 2. * Om * tf + SQ(Om) * ff));
    double U3 = U0 * Om;
    memset(e, 0, 16 * sizeof(double));
    e[0][0] = U0;
    e[0][3] = U3;
    e[1][1] = sqrt(1. / g->rr);
    e[2][2] = -sqrt(1. / g->hh);
    double k1 = (tf * U3 + tt * U0);
    double k2 = (ff * U3 + tf * U0);
    e[3][0] = -((k1) >= 0.0 ? (+1.0) : (-1.0)) * k2 / sqrt((ff * tt - tf * tf) * (tt * U0 * U0 + ff * U3 * U3 + 2.0 * tf * U0 * U3));
    e[3][3] = e[3][0] * (-k1 / k2);
}
/* sim5kerr.c:925-943 and :947-970 */
static void to_frame(const double V[4], double out[4], double e[4][4], const orc_metric* g)
{
    out[0] = -dot4(e[0], V, g);
    out[1] = +dot4(e[1], V, g);
    out[2] = +dot4(e[2], V, g);
    out[3] = +dot4(e[3], V, g);
}
static void from_frame(const double V[4], double out[4], double e[4][4])
{
    for (int i = 0; i < 4; i++) {
        out[i] = 0.0;
        for (int j = 0; j < 4; j++) out[i] += V[j] * e[j][i];
    }
}
/* sim5kerr.c:552-572 */
static void rescale_to(double V[4], double norm, const orc_metric* g)
{
    double N = dot4(V, V, g);
    for (int i = 0; i < 4; i++) V[i] *= sqrt(norm / N);
}
/* Walker-Penrose constant, Connors, Piran & Stark (1980).  sim5polarization.c:144-168 */
static double complex walker_penrose(const double k[4], const double f[4], const orc_metric* g)
{
    double a = g->a, m = g->m, r = g->r;
    double A1 = (k[0] * f[1] - k[1] * f[0]) + a * (1. - m * m) * (k[1] * f[3] - k[3] * f[1]);
    double A2 = sqrt(1. - m * m) * ((r * r + a * a) * (k[3] * f[2] - k[2] * f[3]) - a * (k[0] * f[2] - k[2] * f[0]));
    double wp1 = +r * A1 - a * m * A2;
    double wp2 = -r * A2 - a * m * A1;
    return CMPLX(wp1, wp2);
}
/* rotation of the polarization angle between emitter and observer at infinity.  sim5polarization.c:271-285 */
static double angle_rotation(double a, double inc, double alpha, double beta, double complex kappa)
{
    double k1 = creal(kappa), k2 = cimag(kappa);
    double S = -alpha - a * sin(inc);
    double T = +beta;
    double X = (-S * k2 - T * k1) / (S * S + T * T);
    double Y = (-S * k1 + T * k2) / (S * S + T * T);
    return atan2(Y, X);
}

/* Novikov-Thorne disk: disk_nt_setup stores mass, spin, mdot and r_min as FLOAT statics (sim5disk-nt.c:27-32, 37-77);
 * r_min = r_ms of the float spin + 1e-3 (:90-105); Page-Thorne flux (:109-146) */
typedef struct { float mass, spin, mdot, rms; } orc_disk;
static void disk_setup(orc_disk* d, double M, double a, double mdot)
{
    d->mass = (float)M;
    d->spin = (float)a;
    d->mdot = (float)mdot;
    double as = d->spin;
    double sga = (as >= 0.0) ? +1. : -1.;
    double z1 = 1. + pow(1. - as * as, 1. / 3.) * (pow(1. + as, 1. / 3.) + pow(1. - as, 1. / 3.));
    double z2 = sqrt(3. * as * as + z1 * z1);
    d->rms = (float)(3. + z2 - sga * sqrt((3. - z1) * (3. + z1 + 2. * z2)) + 1e-3);
}
static double disk_flux(const orc_disk* d, double r)
{
    if (r <= d->rms) return 0.0;
    double a = d->spin;
    double x = sqrt(r);
    double x0 = sqrt(d->rms);
    double x1 = +2. * cos(1. / 3. * acos(a) - M_PI / 3.);
    double x2 = +2. * cos(1. / 3. * acos(a) + M_PI / 3.);
    double x3 = -2. * cos(1. / 3. * acos(a));
    double f0 = x - x0 - 1.5 * a * log(x / x0);
    double f1 = 3. * SQ(x1 - a) / (x1 * (x1 - x2) * (x1 - x3)) * log((x - x1) / (x0 - x1));
    double f2 = 3. * SQ(x2 - a) / (x2 * (x2 - x1) * (x2 - x3)) * log((x - x2) / (x0 - x2));
    double f3 = 3. * SQ(x3 - a) / (x3 * (x3 - x1) * (x3 - x2)) * log((x - x3) / (x0 - x3));
    double F = 1. / (4. * M_PI * r) * 1.5 / (x * x * (x * x * x - 3. * x + 2. * a)) * (f0 - f1 - f2 - f3);
    return 9.1721376255e+28 * F * d->mdot / d->mass;
}

/* ------------------------------------------------------------------------------------------ */
/* the pixel loops the reference leaves to its callers (same definitions as oracle/ref_driver.c) */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double r, phi, g, flux, chi, delta, mue; unsigned char status; } orc_pixel;

static int type_bits(int type, int valid)
{
    if (!valid) return SIM5_GT_NONE;
    switch (type) {
        case TY_RR: return SIM5_GT_RR;
        case TY_RC: return SIM5_GT_RC;
        case TY_CC: return SIM5_GT_CC;
        case TY_RR_DBL: return SIM5_GT_RR_DBL;
        case TY_RR_BH: return SIM5_GT_RR_BH;
    }
    return SIM5_GT_NONE;
}
/* harness: Chandrasekhar limb polarization table of sim5_b200.h, linear interpolation */
static double limb_polarization(double mue)
{
    double mu = fmin(fmax(mue, 0.0), 1.0);
    double t = mu * (double)(SIM5_CHANDRA_N - 1);
    int i0 = (int)t;
    if (i0 > SIM5_CHANDRA_N - 2) i0 = SIM5_CHANDRA_N - 2;
    return SIM5_CHANDRA_DELTA[i0] + (SIM5_CHANDRA_DELTA[i0 + 1] - SIM5_CHANDRA_DELTA[i0]) * (t - (double)i0);
}

/* examples/04-disk-image-eqplane/disk-image.c:60-104 generalised to crossing orders 0..max_order, plus the
 * polarized composition of SURVEY.md 8(d) cfg 3 */
static void trace_pixel(const sim5_image_params* p, const orc_disk* disk, double rmin, double alpha, double beta, orc_pixel* o)
{
    orc_ray g;
    int err = 0;
    memset(o, 0, sizeof *o);
    memset(&g, 0, sizeof g);
    g.type = -1;
    ray_from_infinity(p->incl, p->bh_spin, alpha, beta, &g, &err);
    if (err) {
        o->status = (unsigned char)((SIM5_ST_INITERR + err) | (type_bits(g.type, err == E_RR_DOUBLE) << 5));
        return;
    }
    int tb = type_bits(g.type, 1) << 5;
    for (int order = 0; order <= p->max_order; order++) {
        double P = midplane_crossing(&g, order);
        if (isnan(P)) {
            o->status = (unsigned char)((order == 0 ? SIM5_ST_NOCROSS0 : order == 1 ? SIM5_ST_NOCROSS1 : SIM5_ST_NOCROSS2) | tb);
            return;
        }
        double r = radius_at(&g, P);
        if (!(r >= rmin)) continue;
        o->status = (unsigned char)((order == 0 ? SIM5_ST_HIT0 : order == 1 ? SIM5_ST_HIT1 : SIM5_ST_HIT2) | tb);
        o->r = r;
        if (p->outputs & SIM5_OUT_PHI) o->phi = azimuth_at(&g, r, 0.0, P);
        if (p->mode == SIM5_MODE_POLARIZED) {
            double a = p->bh_spin;
            double k[4], U[4], N[4], kl[4], fl[4], f[4], e[4][4];
            const double e_t[4] = {1.0, 0.0, 0.0, 0.0}, e_z[4] = {0.0, 0.0, 1.0, 0.0};
            orc_metric M;
            momentum_from_constants(a, r, 0.0, g.l, g.q, g.Rpc - P, 1.0, k);
            metric_at(a, r, 0.0, &M);
            frame_azimuthal(&M, kepler_omega(r, a), e);
            from_frame(e_t, U, e);
            from_frame(e_z, N, e);
            double kU = dot4(k, U, &M);
            double gf = (k[0] * M.tt + k[3] * M.tf) / kU;
            double mue = dot4(k, N, &M) / kU;
            to_frame(k, kl, e, &M);
            fl[0] = 0.0; fl[1] = -kl[3]; fl[2] = 0.0; fl[3] = kl[1];
            from_frame(fl, f, e);
            rescale_to(f, 1.0, &M);
            o->chi = angle_rotation(a, p->incl, g.alpha, g.beta, walker_penrose(k, f, &M));
            o->mue = mue;
            o->delta = limb_polarization(mue);
            o->g = gf;
            o->flux = disk_flux(disk, r) * pow(gf, 4.);
        } else {
            double gf = g_kepler(r, p->bh_spin, g.l);
            o->g = gf;
            o->flux = disk_flux(disk, r) * pow(gf, 4.);
        }
        return;
    }
    o->status = (unsigned char)(SIM5_ST_MISS | tb);
}

static double now_s(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

double orc_trace_image(const sim5_image_params* p, const sim5_image_out* out, int nthreads, sim5_trace_stats* stats)
{
    if (!p || !out) return -1.0;
    if (p->mode == SIM5_MODE_HISTOGRAM) return -2.0;
    if (p->mode == SIM5_MODE_STEPWISE) return -3.0;        /* not restated here: checked against oracle/_ref + golden */
    if (p->mode == SIM5_MODE_SURFACE) return -3.0;         /* likewise (the surface finder is a caller-side loop over the reference API: oracle/ref_driver.c) */
    int nx = p->nx, ny = p->ny, rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = ny;
    double rmin = (p->r_emit_min > 0.0) ? p->r_emit_min : orc_r_ms(p->bh_spin);
    double rmax = p->rmax;
    orc_disk disk;
    disk_setup(&disk, p->disk_mass, p->bh_spin, p->disk_mdot);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double t0 = now_s();
    int iy;
    #pragma omp parallel for schedule(dynamic, 4)
    for (iy = rb; iy < re; iy++) {
        if (p->split_count > 1 && ((iy - rb) / (p->split_rows > 0 ? p->split_rows : 1)) % p->split_count != p->split_index) continue;
        for (int ix = 0; ix < nx; ix++) {
            double alpha = (((double)(ix) + .5) / (double)(nx) - 0.5) * 2.0 * rmax;                           /* disk-image.c:57 */
            double beta = (((double)(iy) + .5) / (double)(ny) - 0.5) * 2.0 * rmax * ((double)ny / (double)nx);  /* disk-image.c:58 */
            orc_pixel o;
            trace_pixel(p, &disk, rmin, alpha, beta, &o);
            size_t i = (size_t)iy * (size_t)nx + (size_t)ix;
            if ((p->outputs & SIM5_OUT_R) && out->r) out->r[i] = o.r;
            if ((p->outputs & SIM5_OUT_PHI) && out->phi) out->phi[i] = o.phi;
            if ((p->outputs & SIM5_OUT_G) && out->g) out->g[i] = o.g;
            if ((p->outputs & SIM5_OUT_FLUX) && out->flux) out->flux[i] = o.flux;
            if ((p->outputs & SIM5_OUT_CHI) && out->chi) out->chi[i] = o.chi;
            if ((p->outputs & SIM5_OUT_DELTA) && out->delta) out->delta[i] = o.delta;
            if ((p->outputs & SIM5_OUT_MUE) && out->mue) out->mue[i] = o.mue;
            if ((p->outputs & SIM5_OUT_STATUS) && out->status) out->status[i] = o.status;
        }
    }
    double dt = now_s() - t0;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = (int64_t)(re - rb) * nx;
        if ((p->outputs & SIM5_OUT_STATUS) && out->status)
            for (iy = rb; iy < re; iy++)
                for (int ix = 0; ix < nx; ix++) {
                    unsigned char s = out->status[(size_t)iy * nx + ix];
                    stats->class_count[SIM5_ST_CLASS(s)]++;
                    stats->gtype_count[SIM5_ST_GTYPE(s)]++;
                }
        stats->total_ms = dt * 1e3;
    }
    return dt;
}

/* transfer-function lattice of SURVEY.md 8(d) cfg 5: per image a histogram of g weighted by F g^4 dalpha dbeta,
 * accumulated row by row and then over rows in row order (the order oracle/ref_driver.c uses) */
double orc_trace_histogram(const sim5_image_params* p, double* hist, int nthreads)
{
    if (!p || !hist) return -1.0;
    int nx = p->nx, ny = p->ny, nb = p->n_bins;
    int nimg = p->n_spin * p->n_incl, lb = p->lattice_begin, le = p->lattice_end;
    if (lb == 0 && le == 0) le = nimg;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double t0 = now_s();
    double* rows = (double*)malloc(sizeof(double) * (size_t)nb * (size_t)ny);
    for (int img = lb; img < le; img++) {
        int js = img / p->n_incl, ki = img % p->n_incl;
        sim5_image_params q = *p;
        q.mode = SIM5_MODE_EQPLANE;
        q.outputs = SIM5_OUT_G | SIM5_OUT_FLUX | SIM5_OUT_STATUS;
        q.bh_spin = (p->n_spin > 1) ? p->spin_max * (double)js / (double)(p->n_spin - 1) : p->spin_max;
        if (q.bh_spin < 1e-4) q.bh_spin = 1e-4;
        double ideg = (p->n_incl > 1) ? p->incl_min_deg + (p->incl_max_deg - p->incl_min_deg) * (double)ki / (double)(p->n_incl - 1) : p->incl_min_deg;
        q.incl = ideg / 180.0 * M_PI;                      /* deg2rad, sim5math.h:50 */
        double rms = orc_r_ms(q.bh_spin);
        q.rmax = rms + p->rmax_offset;
        double da = 2.0 * q.rmax / (double)nx;
        double db = 2.0 * q.rmax * ((double)ny / (double)nx) / (double)ny;
        orc_disk disk;
        disk_setup(&disk, q.disk_mass, q.bh_spin, q.disk_mdot);
        memset(rows, 0, sizeof(double) * (size_t)nb * (size_t)ny);
        int iy;
        #pragma omp parallel for schedule(dynamic, 4)
        for (iy = 0; iy < ny; iy++) {
            double* h = rows + (size_t)iy * nb;
            for (int ix = 0; ix < nx; ix++) {
                double alpha = (((double)(ix) + .5) / (double)(nx) - 0.5) * 2.0 * q.rmax;
                double beta = (((double)(iy) + .5) / (double)(ny) - 0.5) * 2.0 * q.rmax * ((double)ny / (double)nx);
                orc_pixel o;
                trace_pixel(&q, &disk, rms, alpha, beta, &o);
                int cls = SIM5_ST_CLASS(o.status);
                if (cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2) {
                    double t = (o.g - p->g_min) / (p->g_max - p->g_min) * (double)nb;
                    if (t >= 0.0 && t < (double)nb) h[(int)t] += o.flux * da * db;
                }
            }
        }
        double* H = hist + (size_t)img * nb;
        for (int b = 0; b < nb; b++) {
            double s = 0.0;
            for (iy = 0; iy < ny; iy++) s += rows[(size_t)iy * nb + b];
            H[b] = s;
        }
    }
    free(rows);
    return now_s() - t0;
}

/* black-body specific intensity on an energy grid.  sim5radiation.c:56-78 (constants of sim5const.h:33-87) */
static void planck_grid(double T, double hardf, double cos_mu, const double* E, double* Iv, int n)
{
    const double h = 6.626069e-27, c = 2.997925e+10, kB = 1.380650e-16, kev2hz = 2.417990e+17;
    if (T <= 0.0) return;
    double limbf = (cos_mu >= 0.0) ? 0.5 + 0.75 * cos_mu : 1.0;
    double BB1 = limbf * 2.0 * h / SQ(c) / (hardf * hardf * hardf * hardf) * (kev2hz * kev2hz * kev2hz * kev2hz);
    double BB2 = (h * kev2hz) / (kB * hardf * T);
    for (int i = 0; i < n; i++) Iv[i] = BB1 * (E[i] * E[i] * E[i]) / expm1(BB2 * E[i]);
}

/* thermal disk spectrum (mode SPECTRUM; same definition as oracle/ref_driver.c ref_trace_spectrum, which follows
 * python/sim5diskraytrace.py:43-134 on the image grid) */
double orc_trace_spectrum(const sim5_image_params* p, double* spec, int nthreads)
{
    if (!p || !spec || p->n_energy < 1 || p->n_energy > 256) return -1.0;
    int nx = p->nx, ny = p->ny, ne = p->n_energy, rb = p->row_begin, re = p->row_end;
    if (rb == 0 && re == 0) re = ny;
    double rmin = (p->r_emit_min > 0.0) ? p->r_emit_min : orc_r_ms(p->bh_spin);
    double rmax = p->rmax;
    double dA = (2.0 * rmax / (double)nx) * (2.0 * rmax * ((double)ny / (double)nx) / (double)ny);
    double E[256];
    for (int k = 0; k < ne; k++) E[k] = (ne <= 1) ? p->e_min_kev : p->e_min_kev * pow(p->e_max_kev / p->e_min_kev, (double)k / (double)(ne - 1));
    sim5_image_params q = *p;
    q.mode = SIM5_MODE_POLARIZED;
    q.outputs = SIM5_OUT_R | SIM5_OUT_G | SIM5_OUT_MUE;
    orc_disk disk;
    disk_setup(&disk, p->disk_mass, p->bh_spin, p->disk_mdot);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    double t0 = now_s();
    double* rows = (double*)calloc((size_t)ne * (size_t)ny, sizeof(double));
    int iy;
    #pragma omp parallel for schedule(dynamic, 4)
    for (iy = rb; iy < re; iy++) {
        if (p->split_count > 1 && ((iy - rb) / (p->split_rows > 0 ? p->split_rows : 1)) % p->split_count != p->split_index) continue;
        double* h = rows + (size_t)iy * ne;
        double Eg[256], Iv[256];
        for (int ix = 0; ix < nx; ix++) {
            double alpha = (((double)(ix) + .5) / (double)(nx) - 0.5) * 2.0 * rmax;
            double beta = (((double)(iy) + .5) / (double)(ny) - 0.5) * 2.0 * rmax * ((double)ny / (double)nx);
            orc_pixel o;
            trace_pixel(&q, &disk, rmin, alpha, beta, &o);
            int cls = SIM5_ST_CLASS(o.status);
            if (!(cls == SIM5_ST_HIT0 || cls == SIM5_ST_HIT1 || cls == SIM5_ST_HIT2)) continue;
            double T = sqrt(sqrt(disk_flux(&disk, o.r) / 5.670400e-05));
            if (!(T > 0.0) || !(o.g > 0.0)) continue;
            for (int j = 0; j < ne; j++) Eg[j] = E[j] / o.g;
            planck_grid(T, p->spec_hardf, (p->spec_limb && o.mue >= 0.0) ? o.mue : -1.0, Eg, Iv, ne);
            for (int j = 0; j < ne; j++) h[j] += Iv[j] * (o.g * o.g * o.g) * dA;
        }
    }
    for (int k = 0; k < ne; k++) {
        double s = 0.0;
        for (iy = rb; iy < re; iy++) s += rows[(size_t)iy * ne + k];
        spec[k] = s;
    }
    free(rows);
    return now_s() - t0;
}

/* element-wise entries for unit tests */
void orc_batch_rf(long n, const double* x, const double* y, const double* z, double* o) { for (long i = 0; i < n; i++) o[i] = orc_rf(x[i], y[i], z[i]); }
void orc_batch_rd(long n, const double* x, const double* y, const double* z, double* o) { for (long i = 0; i < n; i++) o[i] = orc_rd(x[i], y[i], z[i]); }
void orc_batch_rc(long n, const double* x, const double* y, double* o) { for (long i = 0; i < n; i++) o[i] = orc_rc(x[i], y[i]); }
void orc_batch_rj(long n, const double* x, const double* y, const double* z, const double* p, double* o) { for (long i = 0; i < n; i++) o[i] = orc_rj(x[i], y[i], z[i], p[i]); }
void orc_batch_sncndn(long n, const double* u, const double* m, double* sn, double* cn, double* dn) { for (long i = 0; i < n; i++) orc_sncndn(u[i], m[i], &sn[i], &cn[i], &dn[i]); }
