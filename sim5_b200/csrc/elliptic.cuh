/*
 * elliptic.cuh -- Carlson symmetric integrals, Legendre/Jacobi wrappers and the Byrd&Friedman
 * radial/polar integrals used by the analytic Kerr geodesic solver, as sm_100a device code.
 *
 * Behavioural contract: every function returns bit-for-bit what the reference's CPU function of the
 * same name returns (src/sim5elliptic.c, lines cited per function), given IEEE add/mul/div/sqrt with
 * no FMA contraction (the kernels are built with -fmad=false) and the correctly-rounded elementary
 * functions of crmath.cuh in place of glibc's libm.  Operation ORDER therefore follows the reference
 * expressions exactly; everything else (structure, naming, loop shape) is ours.
 */
#ifndef SIM5_ELLIPTIC_CUH
#define SIM5_ELLIPTIC_CUH

#include "crmath.cuh"
#include "fastfp.cuh"

namespace s5 {

using crm::cr_sincos;
using crm::cr_log;
using crm::cr_atan;
using crm::cr_atan2;
using crm::cr_acos;
using crm::cr_asin;

S5_HD S5_INL double sq(double x) { return x * x; }
S5_HD S5_INL double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

#define S5_CARLSON_TOL 0.0003

/* Carlson series constants as a constant-bank table: a 64-bit literal costs two UMOV per use in SASS (the kernels are
 * instruction-issue bound), a c[bank][offset] operand costs nothing.  Same values as the literals of the plain versions. */
#define S5_KC_INIT { 1.0 / 3.0, 0.0003, 0.2, \
    1.0 / 24.0, 0.1, 3.0 / 44.0, 1.0 / 14.0, \
    0.3, 1.0 / 7.0, 0.375, 9.0 / 22.0, \
    3.0 / 14.0, 1.0 / 3.0, 3.0 / 22.0, 3.0 / 26.0, 0.75 * (3.0 / 22.0), 1.5 * (3.0 / 26.0), 0.5 * (1.0 / 3.0), (3.0 / 22.0) + (3.0 / 22.0) }
enum { KC_THIRD = 0, KC_TOL, KC_FIFTH, KC_F1, KC_F2, KC_F3, KC_F4, KC_C1, KC_C2, KC_C3, KC_C4,
       KC_J1, KC_J2, KC_J3, KC_J4, KC_J5, KC_J6, KC_J7, KC_J8, KC_COUNT };
static const double s5_kc_host[KC_COUNT] = S5_KC_INIT;
#if defined(__CUDACC__)
static __constant__ double s5_kc_dev[KC_COUNT] = S5_KC_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define S5KC(i) s5_kc_dev[i]
#else
#define S5KC(i) s5_kc_host[i]
#endif

/* ---- plain-operator versions: the reference spelling with `/` and sqrt(); the fallback of the fast versions below ---- */
/* R_F(x,y,z), duplication theorem with the 5th-order series tail.  sim5elliptic.c:18-52 */
S5_HD S5_NOINL double rf_plain(double x, double y, double z)
{
    constexpr double THIRD = 1.0 / 3.0;
    constexpr double K1 = 1.0 / 24.0, K2 = 0.1, K3 = 3.0 / 44.0, K4 = 1.0 / 14.0;
    double mu, dx, dy, dz;
    do {
        double sx = sqrt(x), sy = sqrt(y), sz = sqrt(z);
        double lam = sx * (sy + sz) + sy * sz;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        z = 0.25 * (z + lam);
        mu = THIRD * (x + y + z);
        dx = (mu - x) / mu;
        dy = (mu - y) / mu;
        dz = (mu - z) / mu;
    } while (max3(fabs(dx), fabs(dy), fabs(dz)) > S5_CARLSON_TOL);
    double e2 = dx * dy - dz * dz;
    double e3 = dx * dy * dz;
    return (1.0 + (K1 * e2 - K2 - K3 * e3) * e2 + K4 * e3) / sqrt(mu);
}

/* R_D(x,y,z).  sim5elliptic.c:58-98 */
S5_HD S5_NOINL double rd(double x, double y, double z)
{
    constexpr double K1 = 3.0 / 14.0, K2 = 1.0 / 6.0, K3 = 9.0 / 22.0, K4 = 3.0 / 26.0, K5 = 0.25 * K3, K6 = 1.5 * K4;
    double acc = 0.0, w = 1.0, mu, dx, dy, dz;
    do {
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
    } while (max3(fabs(dx), fabs(dy), fabs(dz)) > S5_CARLSON_TOL);
    double ea = dx * dy, eb = dz * dz;
    double ec = ea - eb, ed = ea - 6.0 * eb;
    double ee = ed + ec + ec;
    return 3.0 * acc + w * (1.0 + ed * (-K1 + K5 * ed - K6 * dz * ee)
        + dz * (K2 * ee + dz * (-K3 * ec + dz * K4 * ea))) / (mu * sqrt(mu));
}

/* R_C(x,y), Cauchy principal value for y < 0.  sim5elliptic.c:104-137 */
S5_HD S5_NOINL double rc_plain(double x, double y)
{
    constexpr double THIRD = 1.0 / 3.0, K1 = 0.3, K2 = 1.0 / 7.0, K3 = 0.375, K4 = 9.0 / 22.0;
    double pre, mu, s;
    if (y > 0.0) {
        pre = 1.0;
    } else {
        double xs = x - y;
        pre = sqrt(x) / sqrt(xs);
        x = xs;
        y = -y;
    }
    do {
        double lam = 2.0 * sqrt(x) * sqrt(y) + y;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        mu = THIRD * (x + y + y);
        s = (y - mu) / mu;
    } while (fabs(s) > S5_CARLSON_TOL);
    return pre * (1.0 + s * s * (K1 + s * (K2 + s * (K3 + s * K4)))) / sqrt(mu);
}

/* R_J(x,y,z,p), principal value for p < 0; returns 0 on out-of-range arguments like the reference.
 * sim5elliptic.c:144-206 */
S5_HD S5_NOINL double rj_plain(double x, double y, double z, double p)
{
    constexpr double K1 = 3.0 / 14.0, K2 = 1.0 / 3.0, K3 = 3.0 / 22.0, K4 = 3.0 / 26.0,
                     K5 = 0.75 * K3, K6 = 1.5 * K4, K7 = 0.5 * K2, K8 = K3 + K3;
    /* pow(5.0*DBL_MIN,1./3.) and 0.3*pow(0.1*DBL_MAX,1./3.) */
    const double LO = 0x1.13c484138708ep-340, HI = 0x1.674da50c1a606p+338;
    if ((fmin(fmin(x, y), z) < 0.0) || (fmin(fmin(x + y, x + z), fmin(y + z, fabs(p))) < LO) ||
        (fmax(fmax(x, y), fmax(z, fabs(p))) > HI)) return 0.0;

    double ca = 0.0, cb = 0.0, rcx = 0.0, acc = 0.0, w = 1.0;
    double xt, yt, zt, pt;
    if (p > 0.0) {
        xt = x; yt = y; zt = z; pt = p;
    } else {
        xt = fmin(fmin(x, y), z);
        zt = fmax(fmax(x, y), z);
        yt = x + y + z - xt - zt;
        ca = 1.0 / (yt - p);
        cb = ca * (zt - yt) * (yt - xt);
        pt = yt + cb;
        double rho = xt * zt / yt;
        double tau = p * pt / yt;
        rcx = rc_plain(rho, tau);
    }
    double mu, dx, dy, dz, dp;
    do {
        double sx = sqrt(xt), sy = sqrt(yt), sz = sqrt(zt);
        double lam = sx * (sy + sz) + sy * sz;
        double al = sq(pt * (sx + sy + sz) + sx * sy * sz);
        double be = pt * sq(pt + lam);
        acc += w * rc_plain(al, be);
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
    } while (fmax(fmax(fabs(dx), fabs(dy)), fmax(fabs(dz), fabs(dp))) > S5_CARLSON_TOL);
    double ea = dx * (dy + dz) + dy * dz;
    double eb = dx * dy * dz;
    double ec = dp * dp;
    double ed = ea - 3.0 * ec;
    double ee = eb + 2.0 * dp * (ea - ec);
    double res = 3.0 * acc + w * (1.0 + ed * (-K1 + K5 * ed - K6 * ee) + eb * (K7 + dp * (-K8 + dp * K4))
        + dp * ea * (K2 - dp * K3) - K2 * dp * ec) / (mu * sqrt(mu));
    if (p <= 0.0) res = ca * (cb * res + 3.0 * (rcx - rf_plain(xt, yt, zt)));
    return res;
}


/* ---- fast versions (what the kernels call): the same arithmetic, operation for operation, on the branch-free
 * division / square root of fastfp.cuh; quotients by a common divisor share one refined reciprocal.
 *
 * Domain: the fast bodies run only when every argument is zero-or-within [2^-100, 2^100) (R_C: [2^-320, 2^320)), which
 * the entry test establishes with one integer compare per argument.  Inside that domain every operand of every
 * sqrt/division of the duplication loops is a positive normal number far from the exponent limits (the iterates stay
 * between min and max of the arguments; numerators mu-x are exact zeros or >= one ulp of a number >= 2^-322), so the
 * unchecked primitives are exact replacements of the operators.  Anything else (negative p aside, which has its own
 * prologue) takes the plain version.  tests: test_gpu_parity.py::test_carlson_fast_vs_plain. ---- */
#if defined(S5_NO_FASTFP)
S5_HD S5_INL double rf(double x, double y, double z) { return rf_plain(x, y, z); }
S5_HD S5_INL double rc(double x, double y) { return rc_plain(x, y); }
S5_HD S5_INL double rj(double x, double y, double z, double p) { return rj_plain(x, y, z, p); }
#else
S5_HD S5_INL bool above_tol(double d) { return fabs(d) > S5KC(KC_TOL); }
#define S5_EXP_MID 100          /* exponent window of the fast R_F / R_J bodies */
#define S5_EXP_WIDE 320         /* ... of R_C (it receives squares and cubes of R_J's iterates) */

/* R_F body: x zero-or-mid, y and z mid */
S5_HD S5_INL double rf_core(double x, double y, double z)
{
    const double THIRD = S5KC(KC_THIRD);
    const double K1 = S5KC(KC_F1), K2 = S5KC(KC_F2), K3 = S5KC(KC_F3), K4 = S5KC(KC_F4);
    double mu, dx, dy, dz;
    double sx = ff::fsqrt0_nc(x), sy = ff::fsqrt_nc(y), sz = ff::fsqrt_nc(z);
    for (;;) {
        double lam = sx * (sy + sz) + sy * sz;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        z = 0.25 * (z + lam);
        mu = THIRD * (x + y + z);
        const double ex = mu - x, ey = mu - y, ez = mu - z;
        if (!ff::surely_above_tol(ff::imax_(ff::imax_(ff::hi_abs(ex), ff::hi_abs(ey)), ff::hi_abs(ez)), mu)) {
            ff::Rcp rmu = ff::rcp_of(mu);
            dx = ff::fdiv_nc(ex, rmu);
            dy = ff::fdiv_nc(ey, rmu);
            dz = ff::fdiv_nc(ez, rmu);
            if (!(above_tol(dx) || above_tol(dy) || above_tol(dz))) break;     /* == !(max3(|dx|,|dy|,|dz|) > tol) */
        }
        sx = ff::fsqrt_nc(x); sy = ff::fsqrt_nc(y); sz = ff::fsqrt_nc(z);
    }
    double e2 = dx * dy - dz * dz;
    double e3 = dx * dy * dz;
    return ff::fdiv_nc(1.0 + (K1 * e2 - K2 - K3 * e3) * e2 + K4 * e3, ff::fsqrt_nc(mu));
}
S5_HD S5_INL bool rf_fast_domain(double x, double y, double z)
{
    return ff::zero_or_pos_within<S5_EXP_MID>(x) && ff::pos_within<S5_EXP_MID>(y) && ff::pos_within<S5_EXP_MID>(z);
}
S5_HD S5_NOINL double rf(double x, double y, double z)
{
    if (!rf_fast_domain(x, y, z)) return rf_plain(x, y, z);
    return rf_core(x, y, z);
}

/* R_C body for y > 0: x zero-or-wide, y wide; sx = sqrt(x) comes from the caller, who may know it exactly
 * (x = RN(v*v) => sqrt(x) = |v| for every binary64 v whose square is a normal number) */
S5_HD S5_INL double rc_pos_core(double x, double y, double sx)
{
    const double THIRD = S5KC(KC_THIRD), K1 = S5KC(KC_C1), K2 = S5KC(KC_C2), K3 = S5KC(KC_C3), K4 = S5KC(KC_C4);
    double mu, s;
    double sy = ff::fsqrt_nc(y);
    for (;;) {
        double lam = 2.0 * sx * sy + y;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        mu = THIRD * (x + y + y);
        const double ey = y - mu;
        if (!ff::surely_above_tol(ff::hi_abs(ey), mu)) {
            s = ff::fdiv_nc(ey, mu);
            if (!above_tol(s)) break;
        }
        sx = ff::fsqrt_nc(x); sy = ff::fsqrt_nc(y);
    }
    return ff::fdiv_nc(1.0 + s * s * (K1 + s * (K2 + s * (K3 + s * K4))), ff::fsqrt_nc(mu));      /* pre == 1: 1.0 * v == v */
}
/* R_C body for y < 0 (Cauchy principal value): x zero-or-wide, -y wide */
S5_HD S5_INL double rc_neg_core(double x, double y)
{
    const double THIRD = S5KC(KC_THIRD), K1 = S5KC(KC_C1), K2 = S5KC(KC_C2), K3 = S5KC(KC_C3), K4 = S5KC(KC_C4);
    double mu, s;
    double xs = x - y;
    double sx = ff::fsqrt_nc(xs);
    double pre = ff::fdiv_nc(ff::fsqrt0_nc(x), sx);
    x = xs;
    y = -y;
    double sy = ff::fsqrt_nc(y);
    for (;;) {
        double lam = 2.0 * sx * sy + y;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        mu = THIRD * (x + y + y);
        const double ey = y - mu;
        if (!ff::surely_above_tol(ff::hi_abs(ey), mu)) {
            s = ff::fdiv_nc(ey, mu);
            if (!above_tol(s)) break;
        }
        sx = ff::fsqrt_nc(x); sy = ff::fsqrt_nc(y);
    }
    return ff::fdiv_nc(pre * (1.0 + s * s * (K1 + s * (K2 + s * (K3 + s * K4)))), ff::fsqrt_nc(mu));
}
S5_HD S5_NOINL double rc(double x, double y)
{
    if (ff::zero_or_pos_within<S5_EXP_WIDE>(x)) {
        if (ff::pos_within<S5_EXP_WIDE>(y)) return rc_pos_core(x, y, ff::fsqrt0_nc(x));
        if (ff::pos_within<S5_EXP_WIDE>(-y)) return rc_neg_core(x, y);
    }
    return rc_plain(x, y);
}

/* tail of R_J after convergence.  sim5elliptic.c:197-203 */
S5_HD S5_INL double rj_tail(double acc, double w, double mu, double dx, double dy, double dz, double dp)
{
    const double K1 = S5KC(KC_J1), K2 = S5KC(KC_J2), K3 = S5KC(KC_J3), K4 = S5KC(KC_J4),
                 K5 = S5KC(KC_J5), K6 = S5KC(KC_J6), K7 = S5KC(KC_J7), K8 = S5KC(KC_J8);
    double ea = dx * (dy + dz) + dy * dz;
    double eb = dx * dy * dz;
    double ec = dp * dp;
    double ed = ea - 3.0 * ec;
    double ee = eb + 2.0 * dp * (ea - ec);
    return 3.0 * acc + ff::fdiv_nc(w * (1.0 + ed * (-K1 + K5 * ed - K6 * ee) + eb * (K7 + dp * (-K8 + dp * K4))
        + dp * ea * (K2 - dp * K3) - K2 * dp * ec), mu * ff::fsqrt_nc(mu));
}

/* R_F(x,y,z) (if WANT_RF) and R_J(x,y,z,p_k), k < NJ, over ONE duplication sequence: sqrt(x), sqrt(y), sqrt(z), lambda,
 * the sums and the x,y,z updates do not depend on p, so the elliptic_pi_cos pairs of the azimuth (both poles r+ and r-
 * at the same amplitude, plus the R_F of the same amplitude) pay for them once.  Each function keeps its own
 * convergence test and leaves the loop at the iteration the stand-alone routine would.  x zero-or-mid; y, z, p_k mid. */
template <int NJ, bool WANT_RF>
S5_HD S5_INL void rfj_shared_core(double x, double y, double z, const double* p, double* rf_out, double* rj_out)
{
    const double THIRD = S5KC(KC_THIRD);
    const double F1 = S5KC(KC_F1), F2 = S5KC(KC_F2), F3 = S5KC(KC_F3), F4 = S5KC(KC_F4);
    double pt[NJ], acc[NJ];
    bool jdone[NJ];
    #pragma unroll
    for (int k = 0; k < NJ; k++) { pt[k] = p[k]; acc[k] = 0.0; jdone[k] = false; }
    bool fdone = !WANT_RF;
    double w = 1.0;
    double sx = ff::fsqrt0_nc(x), sy = ff::fsqrt_nc(y), sz = ff::fsqrt_nc(z);
    for (;;) {
        double lam = sx * (sy + sz) + sy * sz;
        double ssum = sx + sy + sz;
        double sprod = sx * sy * sz;
        constexpr bool SOLO = (NJ == 1) && !WANT_RF;      /* one function: it leaves the loop when it converges, no flags needed */
        #pragma unroll
        for (int k = 0; k < NJ; k++) {
            if (SOLO || !jdone[k]) {
                double v = pt[k] * ssum + sprod;
                double al = sq(v);
                double be = pt[k] * sq(pt[k] + lam);
                acc[k] += w * rc_pos_core(al, be, fabs(v));
                pt[k] = 0.25 * (pt[k] + lam);
            }
        }
        w = 0.25 * w;
        x = 0.25 * (x + lam);
        y = 0.25 * (y + lam);
        z = 0.25 * (z + lam);
        double s3 = x + y + z;
        if (WANT_RF && !fdone) {
            double mu = THIRD * s3;
            const double ex = mu - x, ey = mu - y, ez = mu - z;
            if (!ff::surely_above_tol(ff::imax_(ff::imax_(ff::hi_abs(ex), ff::hi_abs(ey)), ff::hi_abs(ez)), mu)) {
            ff::Rcp rmu = ff::rcp_of(mu);
            double dx = ff::fdiv_nc(ex, rmu), dy = ff::fdiv_nc(ey, rmu), dz = ff::fdiv_nc(ez, rmu);
            if (!(above_tol(dx) || above_tol(dy) || above_tol(dz))) {
                double e2 = dx * dy - dz * dz;
                double e3 = dx * dy * dz;
                *rf_out = ff::fdiv_nc(1.0 + (F1 * e2 - F2 - F3 * e3) * e2 + F4 * e3, ff::fsqrt_nc(mu));
                fdone = true;
            }
            }
        }
        bool all = fdone;
        #pragma unroll
        for (int k = 0; k < NJ; k++) {
            if (SOLO || !jdone[k]) {
                double mu = S5KC(KC_FIFTH) * (s3 + pt[k] + pt[k]);
                const double ex = mu - x, ey = mu - y, ez = mu - z, ep = mu - pt[k];
                if (!ff::surely_above_tol(ff::imax_(ff::imax_(ff::hi_abs(ex), ff::hi_abs(ey)), ff::imax_(ff::hi_abs(ez), ff::hi_abs(ep))), mu)) {
                    ff::Rcp rmu = ff::rcp_of(mu);
                    double dx = ff::fdiv_nc(ex, rmu), dy = ff::fdiv_nc(ey, rmu), dz = ff::fdiv_nc(ez, rmu), dp = ff::fdiv_nc(ep, rmu);
                    if (!(above_tol(dx) || above_tol(dy) || above_tol(dz) || above_tol(dp))) {
                        rj_out[k] = rj_tail(acc[k], w, mu, dx, dy, dz, dp);
                        jdone[k] = true;
                    }
                }
            }
            all = all && jdone[k];
        }
        if (all) break;
        sx = ff::fsqrt_nc(x); sy = ff::fsqrt_nc(y); sz = ff::fsqrt_nc(z);
    }
}

S5_HD S5_INL bool rj_args_bad(double x, double y, double z, double p)
{
    /* pow(5.0*DBL_MIN,1./3.) and 0.3*pow(0.1*DBL_MAX,1./3.) */
    const double LO = 0x1.13c484138708ep-340, HI = 0x1.674da50c1a606p+338;
    return (fmin(fmin(x, y), z) < 0.0) || (fmin(fmin(x + y, x + z), fmin(y + z, fabs(p))) < LO) ||
           (fmax(fmax(x, y), fmax(z, fabs(p))) > HI);
}
/* R_J, p < 0 (principal value): the prologue/epilogue of sim5elliptic.c:166-177, 204 around the shared body */
S5_HD S5_INL double rj_neg_core(double x, double y, double z, double p)
{
    double xt = fmin(fmin(x, y), z);
    double zt = fmax(fmax(x, y), z);
    double yt = x + y + z - xt - zt;
    double ca = ff::fdiv_nc(1.0, yt - p);
    double cb = ca * (zt - yt) * (yt - xt);
    double pt = yt + cb;
    ff::Rcp ryt = ff::rcp_of(yt);
    double rho = ff::fdiv_nc(xt * zt, ryt);
    double tau = ff::fdiv_nc(p * pt, ryt);
    double rcx = rc_neg_core(rho, tau);
    /* body: identical to the p > 0 loop; the final x,y,z are also needed for the R_F term */
    constexpr double THIRD = 1.0 / 3.0;
    (void)THIRD;
    double acc = 0.0, w = 1.0, mu, dx, dy, dz, dp;
    double sx = ff::fsqrt0_nc(xt), sy = ff::fsqrt_nc(yt), sz = ff::fsqrt_nc(zt);
    for (;;) {
        double lam = sx * (sy + sz) + sy * sz;
        double v = pt * (sx + sy + sz) + sx * sy * sz;
        double al = sq(v);
        double be = pt * sq(pt + lam);
        acc += w * rc_pos_core(al, be, fabs(v));
        w = 0.25 * w;
        xt = 0.25 * (xt + lam);
        yt = 0.25 * (yt + lam);
        zt = 0.25 * (zt + lam);
        pt = 0.25 * (pt + lam);
        mu = S5KC(KC_FIFTH) * (xt + yt + zt + pt + pt);
        const double ex = mu - xt, ey = mu - yt, ez = mu - zt, ep = mu - pt;
        if (!ff::surely_above_tol(ff::imax_(ff::imax_(ff::hi_abs(ex), ff::hi_abs(ey)), ff::imax_(ff::hi_abs(ez), ff::hi_abs(ep))), mu)) {
            ff::Rcp rmu = ff::rcp_of(mu);
            dx = ff::fdiv_nc(ex, rmu); dy = ff::fdiv_nc(ey, rmu); dz = ff::fdiv_nc(ez, rmu); dp = ff::fdiv_nc(ep, rmu);
            if (!(above_tol(dx) || above_tol(dy) || above_tol(dz) || above_tol(dp))) break;
        }
        sx = ff::fsqrt_nc(xt); sy = ff::fsqrt_nc(yt); sz = ff::fsqrt_nc(zt);
    }
    double res = rj_tail(acc, w, mu, dx, dy, dz, dp);
    return ca * (cb * res + 3.0 * (rcx - rf_core(xt, yt, zt)));
}
S5_HD S5_INL bool rj_fast_domain(double x, double y, double z, double p)
{
    return ff::zero_or_pos_within<S5_EXP_MID>(x) && ff::pos_within<S5_EXP_MID>(y) && ff::pos_within<S5_EXP_MID>(z) && ff::pos_within<S5_EXP_MID>(p);
}
S5_HD S5_NOINL double rj(double x, double y, double z, double p)
{
    if (rj_fast_domain(x, y, z, p)) {             /* implies the reference's argument test passes */
        double v;
        rfj_shared_core<1, false>(x, y, z, &p, nullptr, &v);
        return v;
    }
    if (rj_args_bad(x, y, z, p)) return 0.0;
    /* p < 0 with x,y,z,|p| in range and the sorted middle/largest strictly positive and distinct from a zero smallest */
    if (p < 0.0 && ff::pos_within<S5_EXP_MID>(-p) && ff::zero_or_pos_within<S5_EXP_MID>(x) && ff::zero_or_pos_within<S5_EXP_MID>(y) && ff::zero_or_pos_within<S5_EXP_MID>(z)
        && ((x > 0.0) + (y > 0.0) + (z > 0.0) >= 2)) {
        double xt = fmin(fmin(x, y), z), zt = fmax(fmax(x, y), z);
        double yt = x + y + z - xt - zt;
        double pt = yt + ff::fdiv_nc(1.0, yt - p) * (zt - yt) * (yt - xt);
        if (ff::pos_within<S5_EXP_MID>(yt) && ff::pos_within<S5_EXP_MID>(pt)) return rj_neg_core(x, y, z, p);
    }
    return rj_plain(x, y, z, p);
}
/* R_F(x,y,z), R_J(x,y,z,p1), R_J(x,y,z,p2) -- bit-identical to the three stand-alone calls */
S5_HD S5_NOINL void rf_rj2(double x, double y, double z, double p1, double p2, double* f, double* j1, double* j2)
{
    if (rj_fast_domain(x, y, z, p1) && ff::pos_within<S5_EXP_MID>(p2)) {
        double p[2] = {p1, p2}, j[2];
        rfj_shared_core<2, true>(x, y, z, p, f, j);
        *j1 = j[0]; *j2 = j[1];
        return;
    }
    *f = rf(x, y, z);
    *j1 = rj(x, y, z, p1);
    *j2 = rj(x, y, z, p2);
}
#endif

/* K(m).  sim5elliptic.c:217-225 */
S5_HD S5_INL double elliptic_k(double m)
{
    if (m == 1.0) m = 1.0 - 1e-8;
    return rf(0.0, 1.0 - m, 1.0);
}

/* F(phi,m) from sin(phi) and from cos(phi).  sim5elliptic.c:273-284, 254-271 */
S5_HD S5_MID double elliptic_f_sin(double s, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (s == 0.0) return 0.0;
    double s2 = sq(s);
    return s * rf(1. - s2, 1.0 - s2 * m, 1.0);
}
S5_HD S5_MID double elliptic_f_cos(double c, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (c == 1.0) return 0.0;
    double base = 0.0;
    if (c < 0.0) {
        c = -c;
        base = 2.0 * rf(0.0, 1.0 - m, 1.0);
    }
    double s2 = 1.0 - sq(c);
    return base + ((base == 0.0) ? (+1) : (-1)) * sqrt(s2) * rf(1.0 - s2, 1.0 - s2 * m, 1.0);
}
/* F(phi,m).  sim5elliptic.c:236-252 */
S5_HD S5_INL double elliptic_f(double phi, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (phi == 0.0) return 0.0;
    int k = 0;
    while (fabs(phi) > M_PI / 2.) { (phi > 0) ? k++ : k--; phi += (phi > 0) ? -M_PI : +M_PI; }
    double s, c;
    cr_sincos(phi, &s, &c);
    double s2 = s * s;
    double v = (phi > 0 ? +1 : -1) * sqrt(s2) * rf(1 - s2, 1.0 - s2 * m, 1.0);
    if (k != 0) v += 2. * k * elliptic_k(m);
    return v;
}

/* E(phi,m) from cos(phi).  sim5elliptic.c:319-337 */
S5_HD S5_INL double elliptic_e_cos(double c, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (c == 1.0) return 0.0;
    double base = 0.0;
    if (c < 0.0) {
        c = -c;
        base = 2.0 * (rf(0.0, 1.0 - m, 1.0) - m * rd(0.0, 1.0 - m, 1.0) / 3.0);
    }
    double c2 = sq(c);
    double s = sqrt(1.0 - c2);
    double q = 1.0 - m + c2 * m;
    return base + ((base == 0.0) ? (+1) : (-1)) * s * (rf(c2, q, 1.0) - sq(s * sqrt(m)) * rd(c2, q, 1.0) / 3.0);
}
/* E(phi,m) from sin(phi).  sim5elliptic.c:339-355 */
S5_HD S5_INL double elliptic_e_sin(double s, double m)
{
    if (m == 1.0) m = 0.99999999;
    if (s == 0.0) return 0.0;
    double s2 = s * s;
    double c2 = 1.0 - s2;
    double q = 1.0 - s2 * m;
    return s * (rf(c2, q, 1.0) - sq(s * sqrt(m)) * rd(c2, q, 1.0) / 3.0);
}

/* complete Pi(n,m).  sim5elliptic.c:365-378 */
S5_HD S5_MID double elliptic_pi_complete(double n, double m)
{
    if (isinf(n)) return 0.0;
    if (m == 1.0) m = 0.99999999;
    if (n == 1.0) n = 0.99999999;
    double q = 1.0 - m;
    return rf(0.0, q, 1.0) + n * rj(0.0, q, 1.0, 1.0 - n) / 3.0;
}
/* Pi(phi,n,m) from cos(phi).  sim5elliptic.c:425-450 */
S5_HD S5_MID double elliptic_pi_cos(double c, double n, double m)
{
    if (isinf(n)) return 0.0;
    if (c == 1.0) return 0.0;
    if (c == 0.0) return elliptic_pi_complete(n, m);
    if (m == 1.0) m = 0.99999999;
    double base = 0.0;
    if (c < 0.0) {
        c = -c;
        base = 2.0 * ((rf(0.0, 1.0 - m, 1.0) + n * rj(0.0, 1.0 - m, 1.0, 1.0 - n) / 3.0));
    }
    double c2 = sq(c);
    double s = sqrt(1.0 - c2);
    double ns2 = -n * (1.0 - c2);
    double q = 1.0 - (1.0 - c2) * m;
    return base + ((base == 0.0) ? (+1) : (-1)) * s * (rf(c2, q, 1.0) - ns2 * rj(c2, q, 1.0, 1.0 + ns2) / 3.0);
}
/* elliptic_pi_cos for several characteristics n at one (cos_phi, m): the R_F term rf(c^2, 1-(1-c^2)m, 1) does not
 * depend on n, so it is evaluated once (or taken from a caller that already has it).  Same bits as elliptic_pi_cos. */
struct PiShare { double c, m, c2, s, q, rfv; bool fast; };
S5_HD S5_MID PiShare pi_share(double c, double m)
{
    PiShare sh;
    sh.c = c; sh.m = m;
    sh.fast = (c > 0.0) && (c != 1.0) && (m != 1.0);
    sh.c2 = sq(c);
    sh.s = sqrt(1.0 - sh.c2);
    sh.q = 1.0 - (1.0 - sh.c2) * m;
    sh.rfv = sh.fast ? rf(sh.c2, sh.q, 1.0) : 0.0;
    return sh;
}
S5_HD S5_INL PiShare pi_share_with(double c, double m, double rfv)      /* rfv == rf(c^2, 1-(1-c^2)m, 1) already known */
{
    PiShare sh;
    sh.c = c; sh.m = m;
    sh.fast = (c > 0.0) && (c != 1.0) && (m != 1.0);
    sh.c2 = sq(c);
    sh.s = sqrt(1.0 - sh.c2);
    sh.q = 1.0 - (1.0 - sh.c2) * m;
    sh.rfv = rfv;
    return sh;
}
S5_HD S5_MID double pi_cos_shared(const PiShare& sh, double n)
{
    if (!sh.fast) return elliptic_pi_cos(sh.c, n, sh.m);
    if (isinf(n)) return 0.0;
    double ns2 = -n * (1.0 - sh.c2);
    return 0.0 + (+1) * sh.s * (sh.rfv - ns2 * rj(sh.c2, sh.q, 1.0, 1.0 + ns2) / 3.0);
}

/* elliptic_pi_cos(c, n1, m) and elliptic_pi_cos(c, n2, m): one amplitude, two characteristics (the two poles r+ and r- of
 * the azimuth integrand).  The R_F term and the x,y,z duplication sequence of both R_J are shared (rf_rj2). */
S5_HD S5_MID void pi_cos_pair(double c, double m, double n1, double n2, double* P1, double* P2)
{
    bool fast = (c > 0.0) && (c != 1.0) && (m != 1.0) && !isinf(n1) && !isinf(n2);
    if (!fast) { *P1 = elliptic_pi_cos(c, n1, m); *P2 = elliptic_pi_cos(c, n2, m); return; }
    double c2 = sq(c);
    double s = sqrt(1.0 - c2);
    double q = 1.0 - (1.0 - c2) * m;
    double ns1 = -n1 * (1.0 - c2), ns2 = -n2 * (1.0 - c2);
    double f, j1, j2;
#if defined(S5_NO_FASTFP)
    f = rf(c2, q, 1.0); j1 = rj(c2, q, 1.0, 1.0 + ns1); j2 = rj(c2, q, 1.0, 1.0 + ns2);
#else
    rf_rj2(c2, q, 1.0, 1.0 + ns1, 1.0 + ns2, &f, &j1, &j2);
#endif
    *P1 = 0.0 + (+1) * s * (f - ns1 * j1 / 3.0);
    *P2 = 0.0 + (+1) * s * (f - ns2 * j2 / 3.0);
}

/* Pi(phi,n,m) from sin(phi).  sim5elliptic.c:453-474 */
S5_HD S5_INL double elliptic_pi_sin(double s, double n, double m)
{
    double s2 = s * s;
    if (isinf(n)) return 0.0;
    if (m == 1.0) m = 0.99999999;
    if (s == 0.0) return 0.0;
    if (s == 1.0) return elliptic_pi_complete(n, m);
    double c2 = 1.0 - s2;
    double ns2 = -n * s2;
    double q = 1.0 - s2 * m;
    return s * (rf(c2, q, 1.0) - ns2 * rj(c2, q, 1.0, 1.0 + ns2) / 3.0);
}

/* inverse Jacobi functions.  sim5elliptic.c:480-486, 492-514, 522-528 */
S5_HD S5_MID double jacobi_isn(double z, double m)
{
    if (fabs(m - 0.0) < 1e-8) return cr_asin(z);
    if (fabs(m - 1.0) < 1e-8) return cr_log(sqrt((1. + z) / (1. - z)));
    return z * rf(1.0 - z * z, 1.0 - m * z * z, 1.0);
}
/* jacobi_icn that also hands out its Carlson value rf(z^2, 1-m(1-z^2), 1) (rfv, valid when *have) so that
 * callers needing the same R_F again (elliptic_pi_cos with the same modulus and cosine) can share it */
S5_HD S5_MID double jacobi_icn_ex(double z, double m, double* rfv, double* z_used, double* m_used, bool* have)
{
    *have = false;
    if ((z > +1.0) && (z < +1.0 + 1e-8)) z = +1.0;
    if ((z < -1.0) && (z > -1.0 - 1e-8)) z = -1.0;
    if ((m > +1.0) && (m < +1.0 + 1e-8)) m = 1.0;
    if ((m < 0.0) && (m > 0.0 - 1e-8)) m = 0.0;

    if (z == 0.0) return elliptic_k(m);
    if (z == 1.0) return 0.0;
    if (m == 0.0) return cr_acos(z);
    if (m == 1.0) return cr_log((1. + sqrt(1. - z)) / z);

    double f = rf(z * z, 1.0 - m * (1. - z * z), 1.0);
    *rfv = f; *z_used = z; *m_used = m; *have = true;
    ff::Quick o;
    double sz = o.sqrt(1. - z * z);
    if (!o.ok) sz = sqrt(1. - z * z);
    double v = sz * f;
    return (z > 0.0) ? v : 2. / sqrt(1. - m) * elliptic_f_sin(-z, m / (m - 1.)) + v;
}
S5_HD S5_INL double jacobi_icn(double z, double m)
{
    double f, zu, mu; bool have;
    return jacobi_icn_ex(z, m, &f, &zu, &mu, &have);
}
S5_HD S5_INL double jacobi_itn(double z, double m)
{
    if (m == 0.0) return cr_atan(z);
    if (m == 1.0) return cr_log(z + sqrt(1. + z * z));
    return jacobi_isn(sqrt(z * z / (1. + z * z)), m);
}

/* sn, cn, dn by descending Landen (AGM) + back substitution.  sim5elliptic.c:535-596 */
template <class OPS>
S5_HD S5_INL void jacobi_sncndn_t(OPS& o, double u, double m, double* sn_, double* cn_, double* dn_)
{
    if (m == 1.0) m = 0.999999999;
    const double CA = 1.0e-8;
    double sn, cn, dn;
    double emc = 1.0 - m;
    double d = 1.0;
    if (emc != 0.0) {
        bool neg = (emc < 0.0);
        if (neg) {
            d = 1.0 - emc;
            emc = o.div(emc, o.div(-1.0, d));
            u *= (d = o.sqrt(d));
        }
        double a = 1.0, c = 0.0;
        double am[13], gm[13];
        int last = 0;
        dn = 1.0;
        for (int i = 0; i < 13; i++) {
            last = i;
            am[i] = a;
            gm[i] = (emc = o.sqrt(emc));
            c = 0.5 * (a + emc);
            if (fabs(a - emc) <= CA * a) break;
            emc *= a;
            a = c;
        }
        u *= c;
        cr_sincos(u, &sn, &cn);
        if (sn != 0.0) {
            a = o.div(cn, sn);
            c *= a;
            for (int i = last; i >= 0; i--) {
                double b = am[i];
                a *= c;
                c *= dn;
                dn = o.div(gm[i] + a, b + a);
                a = o.div(c, b);
            }
            a = o.div(1.0, o.sqrt(c * c + 1.0));
            sn = (sn >= 0.0 ? a : -a);
            cn = c * sn;
        }
        if (neg) {
            a = dn;
            dn = cn;
            cn = a;
            sn = o.div(sn, d);
        }
    } else {
        cn = 1.0 / cosh(u);
        dn = cn;
        sn = tanh(u);
    }
    *sn_ = sn; *cn_ = cn; *dn_ = dn;
}
S5_HD S5_NOINL void jacobi_sncndn(double u, double m, double* sn_, double* cn_, double* dn_)
{
    ff::Quick f;
    jacobi_sncndn_t(f, u, m, sn_, cn_, dn_);
    if (!f.ok) { ff::Plain p; jacobi_sncndn_t(p, u, m, sn_, cn_, dn_); }
}
S5_HD S5_INL double jacobi_sn(double u, double m) { double s, c, d; jacobi_sncndn(u, m, &s, &c, &d); return s; }
S5_HD S5_INL double jacobi_cn(double u, double m) { double s, c, d; jacobi_sncndn(u, m, &s, &c, &d); return c; }
S5_HD S5_INL double jacobi_dn(double u, double m) { double s, c, d; jacobi_sncndn(u, m, &s, &c, &d); return d; }

/* int (1-b sn^2)/(1-a sn^2) du, B&F 340.01.  sim5elliptic.c:676-690 */
S5_HD S5_INL double integral_Z1(double a, double b, double u, double m)
{
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    return 1. / a * ((a - b) * elliptic_pi_cos(cn, a, m) + b * u);
}
/* the same at u = 0: sn=0, cn=1 => Pi = 0 (or 0 for infinite a); keeps the reference's signed-zero/NaN result */
S5_HD S5_INL double integral_Z1_at0(double a, double b)
{
    return 1. / a * ((a - b) * 0.0 + b * 0.0);
}

/* glibc's catan(), restricted to the two shapes integral_R1 can produce: a real argument (x, +-0)
 * and a purely imaginary one (+-0, y).  Returns the complex result.  (glibc s_catan_template.c) */
S5_HD S5_INL void catan_axis(double re, double im, double* ore, double* oim)
{
    if (re == 0.0 && im == 0.0) { *ore = re; *oim = im; return; }
    double absx = fabs(re), absy = fabs(im);
    if (absx < absy) { double t = absx; absx = absy; absy = t; }
    double den = (1 - absx) * (1 + absx);          /* absy == 0 < eps/2 */
    if (den == 0) den = 0;
    *ore = 0.5 * cr_atan2(2 * re, den);
    if (fabs(im) == 1) {
        *oim = copysign(0.5, im) * (0.6931471805599453 - cr_log(fabs(re)));
    } else {
        double r2 = 0.0;                           /* |re| is 0 or the argument is real: r2 = re*re */
        if (fabs(re) >= 4.930380657631324e-32) r2 = re * re;
        double num = im + 1;
        num = r2 + num * num;
        double dd_ = im - 1;
        dd_ = r2 + dd_ * dd_;
        double f = num / dd_;
        if (f < 0.5) {
            *oim = 0.25 * cr_log(f);
        } else {
            num = 4 * im;
            *oim = 0.25 * crm::cr_log1p(num / dd_);
        }
    }
}

/* int du/(1+a cn u), B&F 341.03 / 361.54.  sim5elliptic.c:755-792 (complex arithmetic spelled out) */
S5_HD S5_MID double integral_R1(double a, double u, double m)
{
    double a2 = sq(a);
    double n = a2 / (a2 - 1.);
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    double mma = (m + (1. - m) * a2) / (1. - a2);
    double f1re;
    if (fabs(mma) > 1e-5) {
        /* csqrt(1/mma) * catan( csqrt(mma)*sn/dn ) */
        double inv = 1. / mma;
        double s1re, s1im, wre, wim;
        if (inv < 0.0) { s1re = 0.0; s1im = sqrt(-inv); } else { s1re = fabs(sqrt(inv)); s1im = 0.0; }
        if (mma < 0.0) { wre = 0.0; wim = sqrt(-mma); } else { wre = fabs(sqrt(mma)); wim = 0.0; }
        wre = wre * sn / dn;
        wim = wim * sn / dn;
        double cre, cim;
        catan_axis(wre, wim, &cre, &cim);
        f1re = s1re * cre - s1im * cim;
    } else {
        f1re = sn / dn;
    }
    double ellpi = elliptic_pi_cos(cn, n, m);
    return 1. / (1. - a2) * (ellpi + a * f1re);
}

/* int_a^X dx/((x-p) sqrt((x-a)(x-b)(x-c)(x-d))), B&F 258.39.  sim5elliptic.c:1017-1029 */
S5_HD S5_INL double integral_R_rp_re(double a, double b, double c, double d, double p, double X)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt(((b - d) * (X - a)) / ((a - d) * (X - b)));
    double u1 = jacobi_isn(sn, m2);
    double a2 = (a - d) / (b - d);
    double c2 = ((p - b) * (a - d)) / ((p - a) * (b - d));
    return -2.0 / sqrt((a - c) * (b - d)) / (p - a) * (integral_Z1(c2, a2, u1, m2) - integral_Z1_at0(c2, a2));
}
/* X -> infinity.  sim5elliptic.c:1032-1044 */
S5_HD S5_INL double integral_R_rp_re_inf(double a, double b, double c, double d, double p)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt((b - d) / (a - d));
    double u1 = jacobi_isn(sn, m2);
    double a2 = (a - d) / (b - d);
    double c2 = ((p - b) * (a - d)) / ((p - a) * (b - d));
    return -2.0 / sqrt((a - c) * (b - d)) / (p - a) * (integral_Z1(c2, a2, u1, m2) - integral_Z1_at0(c2, a2));
}
/* int_X1^inf dx/((x-p) sqrt((x-a)(x-b)(x-c)(x-c*))), c = u+iv, B&F 260.04.  sim5elliptic.c:1081-1112 */
S5_HD S5_MID double integral_R_rp_cc2_inf(double a, double b, double cre, double cim, double p, double X1)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double g = 1. / sqrt(A * B);
    double alpha1 = (B * a + b * A - p * A - p * B) / (B * a - b * A + p * A - p * B);
    double alpha2 = (B + A) / (B - A);
    double u1 = elliptic_f_cos((X1 * (A - B) + a * B - b * A) / (X1 * (A + B) - a * B - b * A), m);
    double u2 = elliptic_f_cos((A - B) / (A + B), m);
    double t0 = alpha2 * (u2 - u1);
    double t1 = (alpha1 - alpha2) * (integral_R1(alpha1, u2, m) - integral_R1(alpha1, u1, m));
    return (B - A) * g / (B * a + b * A - p * A - p * B) * (t0 + t1);
}
/* int_X^b dx/((p-x^2) sqrt((a^2+x^2)(b^2-x^2))), B&F 213.02.  sim5elliptic.c:1142-1159 */
S5_HD S5_MID double integral_T_mp(double a2, double b2, double p, double X)
{
    double m = b2 / (a2 + b2);
    double n = b2 / (b2 - p);
    if (X >= 0.0)
        return 1. / sqrt(a2 + b2) / (p - b2) * elliptic_pi_cos(X / sqrt(b2), n, m);
    else
        return 1. / sqrt(a2 + b2) / (p - b2) * (2. * elliptic_pi_complete(n, m) - elliptic_pi_cos(-X / sqrt(b2), n, m));
}

/* ---- the integrals behind geodesic_timedelay ------------------------------------------------------ */
/* int cn u du = acos(dn)/sqrt(m), B&F 312.01.  sim5elliptic.c:645-653 */
S5_HD S5_INL double integral_C1(double u, double m)
{
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    return crm::cr_acos(dn) / sqrt(m);
}
/* int cn^2 u du, B&F 312.02.  sim5elliptic.c:656-673 */
S5_HD S5_INL double integral_C2(double u, double m)
{
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    return 1. / m * (elliptic_e_cos(cn, m) - (1. - m) * u);
}
S5_HD S5_INL double integral_C2_cos(double cn_u, double m)
{
    return 1. / m * (elliptic_e_cos(cn_u, m) - (1. - m) * elliptic_f_cos(cn_u, m));
}
/* int (1-b sn^2)^2/(1-a sn^2)^2 du, B&F 340.02.  sim5elliptic.c:693-715 */
S5_HD S5_MID double integral_Z2(double a, double b, double u, double m)
{
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    double V1 = elliptic_pi_cos(cn, a, m);
    double V2 = 0.5 / ((a - 1.) * (m - a)) * (
                    a * elliptic_e_cos(cn, m) + (m - a) * u +
                    (2. * a * m + 2. * a - a * a - 3. * m) * V1 -
                    (a * a * sn * cn * dn) / (1. - a * sn * sn)
                );
    double ab = a - b;
    return 1. / sq(a) * (sq(b) * u + 2. * b * ab * V1 + ab * ab * V2);
}
/* int (1 + a cn u) du, B&F 341.00.  sim5elliptic.c:718-729 */
S5_HD S5_INL double integral_Rm1(double a, double u, double m)
{
    return u + a / sqrt(m) * crm::cr_acos(jacobi_dn(u, m));
}
/* int (1 + a cn u)^2 du, B&F 341.01.  sim5elliptic.c:732-745 */
S5_HD S5_MID double integral_Rm2(double a, double u, double m)
{
    double a2 = sq(a);
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    return 1 / m * ((m - a2 * (1. - m)) * u + a2 * elliptic_e_cos(cn, m) + 2 * a * sqrt(m) * crm::cr_acos(dn));
}
S5_HD S5_INL double integral_R0(double u, double m) { (void)m; return u; }
/* int du/(1 + a cn u)^2, B&F 341.04.  sim5elliptic.c:795-815 */
S5_HD S5_MID double integral_R2(double a, double u, double m)
{
    double a2 = sq(a);
    double mma = (m + (1. - m) * a2);
    double sn, cn, dn;
    jacobi_sncndn(u, m, &sn, &cn, &dn);
    return 1 / (a2 - 1.) / mma * (
               (a2 * (2. * m - 1.) - 2. * m) * integral_R1(a, u, m) +
               2. * m * integral_Rm1(a, u, m) -
               m * integral_Rm2(a, u, m) +
               a * a2 * sn * dn / (1. + a * cn)
           );
}
/* int_a^X dx/sqrt((x-a)(x-b)(x-c)(x-d)), B&F 258.00.  sim5elliptic.c:825-838, 841-854 */
S5_HD S5_INL double integral_R_r0_re(double a, double b, double c, double d, double X)
{
    double m4 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt(((b - d) * (X - a)) / ((a - d) * (X - b)));
    return 2.0 / sqrt((a - c) * (b - d)) * jacobi_isn(sn, m4);
}
S5_HD S5_INL double integral_R_r0_re_inf(double a, double b, double c, double d)
{
    double m4 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt((b - d) / (a - d));
    return 2.0 / sqrt((a - c) * (b - d)) * jacobi_isn(sn, m4);
}
/* the same with a complex pair c = u + iv, d = c*: B&F 260.00.  sim5elliptic.c:857-872, 875-889 */
S5_HD S5_INL double integral_R_r0_cc(double a, double b, double cre, double cim, double X)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m2 = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double cn = (X * (A - B) + a * B - b * A) / (X * (A + B) - a * B - b * A);
    return 1. / sqrt(A * B) * jacobi_icn(cn, m2);
}
S5_HD S5_INL double integral_R_r0_cc_inf(double a, double b, double cre, double cim)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m2 = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double cn = (A - B) / (A + B);
    return 1. / sqrt(A * B) * jacobi_icn(cn, m2);
}
/* int_a^X x dx/sqrt(...), B&F 258.11.  sim5elliptic.c:892-905 */
S5_HD S5_MID double integral_R_r1_re(double a, double b, double c, double d, double X)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt(((b - d) * (X - a)) / ((a - d) * (X - b)));
    double u = jacobi_isn(sn, m2);
    double a2 = (a - d) / (b - d);
    double b2 = ((a - d) * b) / (a * (b - d));
    double Z = integral_Z1(a2, b2, u, m2) - integral_Z1(a2, b2, 0, m2);
    return a * 2.0 / sqrt((a - c) * (b - d)) * Z;
}
/* int_X1^X2 x dx/sqrt((x-a)(x-b)(x-c)(x-c*)), B&F 260.03.  sim5elliptic.c:908-931 */
S5_HD S5_MID double integral_R_r1_cc(double a, double b, double cre, double cim, double X1, double X2)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double g = 1. / sqrt(A * B);
    double alpha1 = (B * a + b * A) / (B * a - b * A);
    double alpha2 = (B + A) / (B - A);
    double u1 = elliptic_f_cos((X1 * (A - B) + a * B - b * A) / (X1 * (A + B) - a * B - b * A), m);
    double u2 = elliptic_f_cos((X2 * (A - B) + a * B - b * A) / (X2 * (A + B) - a * B - b * A), m);
    double t0 = alpha1 * (integral_R0(u2, m) - integral_R0(u1, m));
    double t1 = (alpha2 - alpha1) * (integral_R1(alpha2, u2, m) - integral_R1(alpha2, u1, m));
    return (B * a - b * A) / (B + A) * g * (t0 + t1);
}
/* int_a^X x^2 dx/sqrt(...), B&F 258.11.  sim5elliptic.c:954-967 */
S5_HD S5_MID double integral_R_r2_re(double a, double b, double c, double d, double X)
{
    double m2 = ((b - c) * (a - d)) / ((a - c) * (b - d));
    double sn = sqrt(((b - d) * (X - a)) / ((a - d) * (X - b)));
    double u = jacobi_isn(sn, m2);
    double a2 = (a - d) / (b - d);
    double b2 = ((a - d) * b) / (a * (b - d));
    double Z = integral_Z2(a2, b2, u, m2) - integral_Z2(a2, b2, 0, m2);
    return sq(a) * 2.0 / sqrt((a - c) * (b - d)) * Z;
}
/* int_X1^X2 x^2 dx/sqrt((x-a)(x-b)(x-c)(x-c*)), B&F 260.03.  sim5elliptic.c:970-995; pow(., 2.) is an exact square in glibc's
 * pow to the last bit of a correctly rounded product */
S5_HD S5_MID double integral_R_r2_cc(double a, double b, double cre, double cim, double X1, double X2)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double g = 1. / sqrt(A * B);
    double alpha1 = (B * a + b * A) / (B * a - b * A);
    double alpha2 = (B + A) / (B - A);
    double u1 = elliptic_f_cos((X1 * (A - B) + a * B - b * A) / (X1 * (A + B) - a * B - b * A), m);
    double u2 = elliptic_f_cos((X2 * (A - B) + a * B - b * A) / (X2 * (A + B) - a * B - b * A), m);
    double t0 = sq(alpha1) * (integral_R0(u2, m) - integral_R0(u1, m));
    double t1 = 2. * alpha1 * (alpha2 - alpha1) * (integral_R1(alpha2, u2, m) - integral_R1(alpha2, u1, m));
    double t2 = sq(alpha2 - alpha1) * (integral_R2(alpha2, u2, m) - integral_R2(alpha2, u1, m));
    return sq((B * a - b * A) / (B + A)) * g * (t0 + t1 + t2);
}
/* int_X1^X2 dx/((x-p) sqrt((x-a)(x-b)(x-c)(x-c*))), B&F 260.04.  sim5elliptic.c:1047-1078 */
S5_HD S5_MID double integral_R_rp_cc2(double a, double b, double cre, double cim, double p, double X1, double X2)
{
    double u = cre;
    double v2 = sq(cim);
    double A = sqrt(sq(a - u) + v2);
    double B = sqrt(sq(b - u) + v2);
    double m = (sq(A + B) - sq(a - b)) / (4. * A * B);
    double g = 1. / sqrt(A * B);
    double alpha1 = (B * a + b * A - p * A - p * B) / (B * a - b * A + p * A - p * B);
    double alpha2 = (B + A) / (B - A);
    double u1 = elliptic_f_cos((X1 * (A - B) + a * B - b * A) / (X1 * (A + B) - a * B - b * A), m);
    double u2 = elliptic_f_cos((X2 * (A - B) + a * B - b * A) / (X2 * (A + B) - a * B - b * A), m);
    double t0 = alpha2 * (integral_R0(u2, m) - integral_R0(u1, m));
    double t1 = (alpha1 - alpha2) * (integral_R1(alpha1, u2, m) - integral_R1(alpha1, u1, m));
    return (B - A) * g / (B * a + b * A - p * A - p * B) * (t0 + t1);
}
/* int_X^b dx/sqrt((a^2+x^2)(b^2-x^2)) and with x^2 in the numerator, B&F 213.00 / 213.06.  sim5elliptic.c:1119-1139 */
S5_HD S5_INL double integral_T_m0(double a2, double b2, double X)
{
    double m = b2 / (a2 + b2);
    return 1. / sqrt(a2 + b2) * jacobi_icn(X / sqrt(b2), m);
}
S5_HD S5_INL double integral_T_m2(double a2, double b2, double X)
{
    double m = b2 / (a2 + b2);
    double cn = X / sqrt(b2);
    return b2 / sqrt(a2 + b2) * (integral_C2_cos(cn, m) - integral_C2(0, m));
}

} /* namespace s5 */
#endif
