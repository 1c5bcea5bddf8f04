/*
 * crmath.cuh -- double-double ("almost always correctly rounded") elementary functions
 * for the sm_100a photon kernels.
 *
 * Why this exists: the reference computes on the CPU with glibc's libm
 * (call sites: sim5kerr-geod.c:73,78,1004-1008; sim5elliptic.c:483-484,502-503,525-526,571-572;
 * sim5kerr.c:1045,1139; sim5disk-nt.c:131-134; sim5polarization.c:284; sim5raytrace.c:177),
 * whose results are the correctly rounded ones in >= 99.85 % of calls.  CUDA's libm
 * is 1-2 ulp and would disagree in tens of % of calls; on the ill-conditioned photon-ring pixels a
 * 1-ulp input change is amplified to > 1e-9.  Evaluating each function in double-double (~2^-68
 * relative or better) and rounding once reproduces glibc bit-for-bit except where glibc itself is
 * not correctly rounded or the exact value lies within ~2^-15 ulp of a rounding boundary.
 *
 * Everything is __host__ __device__ so the very same code can be exercised on the build box
 * (which has no GPU) by tests/hostsim; the product library only ever runs it on the device.
 * All error-free transforms use explicit fma / non-contracted mul/add, so results do not depend
 * on the compiler's contraction flags.
 */
#ifndef SIM5_CRMATH_CUH
#define SIM5_CRMATH_CUH

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define S5_HD __host__ __device__
#define S5_INL __forceinline__
#define S5_NOINL __noinline__
#else
#define S5_HD
#define S5_INL inline
#define S5_NOINL __attribute__((noinline))
#endif
#define S5_CONST static constexpr
/* mid-size routines (a few hundred instructions, several call sites): inlining them multiplies the kernel's code size
 * and the instruction cache starts missing (ncu: stall_no_inst); S5_MID lets the build choose */
#if defined(S5_MID_NOINLINE)
#define S5_MID S5_NOINL
#else
#define S5_MID S5_INL
#endif

namespace crm {

#include "crmath_tables.h"

static const double crm_logtab_host[128][3] = CRM_LOGTAB_INIT;
static const double crm_atantab_host[65][2] = CRM_ATANTAB_INIT;
#if defined(__CUDACC__)
static __device__ const double crm_logtab_dev[128][3] = CRM_LOGTAB_INIT;
static __device__ const double crm_atantab_dev[65][2] = CRM_ATANTAB_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define CRM_LOGTAB(i, k)  __ldg(&crm_logtab_dev[i][k])
#define CRM_ATANTAB(i, k) __ldg(&crm_atantab_dev[i][k])
#else
#define CRM_LOGTAB(i, k)  crm_logtab_host[i][k]
#define CRM_ATANTAB(i, k) crm_atantab_host[i][k]
#endif

/* The polynomial coefficients and split constants as a constant-bank table: in SASS a 64-bit literal costs two UMOV / MOV per use
 * (profiles/r04k: 5 percent of phase A's executed instructions were UMOV, and the kernels are instruction-issue bound), a c[bank][offset]
 * operand of DFMA / DADD / DMUL costs nothing.  Same values: the table is initialised from the constexpr literals of crmath_tables.h. */
#define CRM_KLIST(X) X(PI_H) X(PI_L) X(PI_2_H) X(PI_2_L) X(2_PI) X(PIO2_1) X(PIO2_2) X(PIO2_3) X(PIO2_4) X(LN2_HEAD) X(LN2_MID) X(LN2_TAIL) X(S1_H) X(S1_L) X(S2_H) X(S2_L) X(S3_H) X(S3_L) X(SQ0) X(SQ1) X(SQ2) X(SQ3) X(SQ4) X(SQ5) X(SQ6) X(SQ7) X(C2_H) X(C2_L) X(C3_H) X(C3_L) X(CQ0) X(CQ1) X(CQ2) X(CQ3) X(CQ4) X(CQ5) X(CQ6) X(CQ7) X(LP1) X(LP2) X(LP3) X(LP4) X(LP5) X(LP6) X(LP7) X(LP8) X(THIRD_H) X(THIRD_L) X(AT0) X(AT1) X(AT2) X(AT3) X(AT4) X(AT5)
enum {
#define X(n) KCI_##n,
    CRM_KLIST(X)
#undef X
    KCI_COUNT
};
#if defined(__CUDACC__)
static __constant__ double crm_kc_dev[KCI_COUNT] = {
#define X(n) CRM_##n,
    CRM_KLIST(X)
#undef X
};
#endif
#if defined(__CUDA_ARCH__) && !defined(S5_CRM_LITERALS)
#define KC(n) crm_kc_dev[KCI_##n]
#else
#define KC(n) CRM_##n
#endif

/* ---- exact building blocks (never contracted, never reassociated) ---- */
S5_HD S5_INL double fma_(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
S5_HD S5_INL double mul_(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
S5_HD S5_INL double add_(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
S5_HD S5_INL int64_t bits_of(double x)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t b; memcpy(&b, &x, 8); return b;
#endif
}
S5_HD S5_INL double from_bits(int64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x; memcpy(&x, &b, 8); return x;
#endif
}
S5_HD S5_INL double inf_() { return from_bits(0x7ff0000000000000LL); }
S5_HD S5_INL double pow2i(int e) { return from_bits((int64_t)(e + 1023) << 52); }   /* 2^e, -1022 <= e <= 1023 */

struct dd { double h, l; };

S5_HD S5_INL dd two_sum(double a, double b)
{
    double s = add_(a, b);
    double bb = add_(s, -a);
    double e = add_(add_(a, -add_(s, -bb)), add_(b, -bb));
    return dd{s, e};
}
S5_HD S5_INL dd fast_two_sum(double a, double b)      /* needs |a| >= |b| (or a == 0) */
{
    double s = add_(a, b);
    double e = add_(b, -add_(s, -a));
    return dd{s, e};
}
S5_HD S5_INL dd two_prod(double a, double b)
{
    double p = mul_(a, b);
    return dd{p, fma_(a, b, -p)};
}
S5_HD S5_INL dd dd_neg(dd a) { return dd{-a.h, -a.l}; }
S5_HD S5_INL dd dd_add(dd a, dd b)                   /* accurate, any magnitudes */
{
    dd s = two_sum(a.h, b.h);
    dd t = two_sum(a.l, b.l);
    s.l = add_(s.l, t.h);
    s = fast_two_sum(s.h, s.l);
    s.l = add_(s.l, t.l);
    return fast_two_sum(s.h, s.l);
}
S5_HD S5_INL dd dd_add_fast(dd a, dd b)              /* needs |a.h| >= |b.h|, no heavy cancellation */
{
    dd s = fast_two_sum(a.h, b.h);
    s.l = add_(s.l, add_(a.l, b.l));
    return fast_two_sum(s.h, s.l);
}
S5_HD S5_INL dd dd_add_d(dd a, double b)
{
    dd s = two_sum(a.h, b);
    s.l = add_(s.l, a.l);
    return fast_two_sum(s.h, s.l);
}
S5_HD S5_INL dd dd_mul(dd a, dd b)
{
    dd p = two_prod(a.h, b.h);
    p.l = fma_(a.h, b.l, fma_(a.l, b.h, p.l));
    return fast_two_sum(p.h, p.l);
}
S5_HD S5_INL dd dd_mul_d(dd a, double b)
{
    dd p = two_prod(a.h, b);
    p.l = fma_(a.l, b, p.l);
    return fast_two_sum(p.h, p.l);
}
S5_HD S5_INL dd dd_sqr(dd a)
{
    dd p = two_prod(a.h, a.h);
    p.l = fma_(add_(a.h, a.h), a.l, p.l);
    return fast_two_sum(p.h, p.l);
}
#if defined(S5_DD_DIV3)
S5_HD S5_INL dd dd_div(dd a, dd b)
{
    double q1 = a.h / b.h;
    dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    double q2 = r.h / b.h;
    r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
    double q3 = r.h / b.h;
    dd q = fast_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
#else
/* a / b to ~2^-102 (two quotient digits; the third one of the long-division form above changes the result below 2^-104, far under
 * the 2^-68 this library's error budget needs): one division, one double-double product and two double-double sums less per call */
#if defined(S5_DD_DIV_2DIV)
S5_HD S5_INL dd dd_div(dd a, dd b)
{
    double q1 = a.h / b.h;
    dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    double q2 = r.h / b.h;
    return fast_two_sum(q1, q2);
}
#else
/* both quotient digits from ONE correctly rounded reciprocal of b.h (the same bits on host and device): a digit that is off by an
 * ulp only moves work into the next digit, the remainder is formed exactly either way */
S5_HD S5_INL double rcp_(double b)
{
#if defined(__CUDA_ARCH__)
    return __drcp_rn(b);
#else
    return 1.0 / b;
#endif
}
S5_HD S5_INL dd dd_div(dd a, dd b)
{
    double rb = rcp_(b.h);
    double q1 = mul_(a.h, rb);
    dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    double q2 = mul_(r.h, rb);
    return fast_two_sum(q1, q2);
}
#endif
#endif
S5_HD S5_INL dd dd_div_dd_d(double a, double b)      /* a/b for plain doubles, as a dd */
{
    double q = a / b;
    double r = fma_(-q, b, a);                       /* exact remainder */
    return fast_two_sum(q, r / b);
}
S5_HD S5_INL dd dd_sqrt(dd a)
{
    if (!(a.h > 0.0)) return dd{a.h == 0.0 ? 0.0 : sqrt(a.h), 0.0};
    double s = sqrt(a.h);
    dd p = two_prod(s, s);
    double r = add_(add_(add_(a.h, -p.h), -p.l), a.l);
    return fast_two_sum(s, r / (s + s));
}

/* ================================================================== */
/* sin / cos                                                           */
/* ================================================================== */
S5_HD S5_INL dd sin_kernel(dd r)                      /* |r| <= pi/4 (+eps) */
{
    dd z = dd_sqr(r);
    double zh = z.h;
    double q = KC(SQ7);
    q = fma_(q, zh, KC(SQ6)); q = fma_(q, zh, KC(SQ5)); q = fma_(q, zh, KC(SQ4));
    q = fma_(q, zh, KC(SQ3)); q = fma_(q, zh, KC(SQ2)); q = fma_(q, zh, KC(SQ1));
    q = fma_(q, zh, KC(SQ0));
    dd t = fast_two_sum(KC(S3_H), fma_(zh, q, KC(S3_L)));
    t = dd_add_fast(dd{KC(S2_H), KC(S2_L)}, dd_mul(z, t));
    t = dd_add_fast(dd{KC(S1_H), KC(S1_L)}, dd_mul(z, t));
    dd w = dd_mul(dd_mul(r, z), t);
    return dd_add_fast(r, w);
}
S5_HD S5_INL dd cos_kernel(dd r)
{
    dd z = dd_sqr(r);
    double zh = z.h;
    double q = KC(CQ7);
    q = fma_(q, zh, KC(CQ6)); q = fma_(q, zh, KC(CQ5)); q = fma_(q, zh, KC(CQ4));
    q = fma_(q, zh, KC(CQ3)); q = fma_(q, zh, KC(CQ2)); q = fma_(q, zh, KC(CQ1));
    q = fma_(q, zh, KC(CQ0));
    dd t = fast_two_sum(KC(C3_H), fma_(zh, q, KC(C3_L)));
    t = dd_add_fast(dd{KC(C2_H), KC(C2_L)}, dd_mul(z, t));
    dd v = dd_mul(dd_sqr(z), t);
    dd u = fast_two_sum(1.0, -0.5 * z.h);
    u.l = add_(u.l, -0.5 * z.l);
    u = fast_two_sum(u.h, u.l);
    return dd_add_fast(u, v);
}
/* x = k*pi/2 + r ; returns k mod 4.  Valid for |x| < 2^20*pi/2. */
S5_HD S5_INL int reduce_pio2(double x, dd& r)
{
    if (fabs(x) <= 0.78539816339744828) { r = dd{x, 0.0}; return 0; }
    double k = rint(x * KC(2_PI));
    double a = fma_(-k, KC(PIO2_1), x);               /* exact */
    dd s = two_sum(a, -mul_(k, KC(PIO2_2)));          /* k*PIO2_2 exact */
    s = dd_add_d(s, -mul_(k, KC(PIO2_3)));            /* exact product */
    dd p4 = two_prod(k, KC(PIO2_4));
    s = dd_add(s, dd_neg(p4));
    r = s;
    return (int)((long long)k & 3);
}
/* which: bit0 = want sin, bit1 = want cos */
S5_HD S5_NOINL void cr_sincos(double x, double* sn, double* cs)
{
    double ax = fabs(x);
    if (!(ax < 1.0e6)) {                               /* huge / inf / nan: outside the ray path's domain */
        if (ax != ax || ax > 1.7976931348623157e308) { *sn = x - x; *cs = x - x; return; }
        *sn = sin(x); *cs = cos(x); return;
    }
    if (ax < 7.450580596923828e-09) { *sn = x; *cs = 1.0; return; }   /* 2^-27 */
    dd r;
    int k = reduce_pio2(x, r);
    dd s = sin_kernel(r);
    dd c = cos_kernel(r);
    double S = s.h, Cc = c.h;
    switch (k) {
        case 0:  *sn =  S;  *cs =  Cc; break;
        case 1:  *sn =  Cc; *cs = -S;  break;
        case 2:  *sn = -S;  *cs = -Cc; break;
        default: *sn = -Cc; *cs =  S;  break;
    }
}
S5_HD S5_INL double cr_sin(double x) { double s, c; cr_sincos(x, &s, &c); return s; }
/* cos alone evaluates ONE of the two kernels (the quadrant says which): the same value as the cosine of cr_sincos, half its work
 * (the stepper calls it three times per raytrace() step, the radial roots once per ray) */
S5_HD S5_NOINL double cr_cos(double x)
{
#if defined(S5_COS_VIA_SINCOS)
    double s, c; cr_sincos(x, &s, &c); return c;
#else
    double ax = fabs(x);
    if (!(ax < 1.0e6)) {
        if (ax != ax || ax > 1.7976931348623157e308) return x - x;
        return cos(x);
    }
    if (ax < 7.450580596923828e-09) return 1.0;
    dd r;
    int k = reduce_pio2(x, r);
    double v = (k & 1) ? sin_kernel(r).h : cos_kernel(r).h;
    return (k == 1 || k == 2) ? -v : v;          /* k = 0: +cos r, 1: -sin r, 2: -cos r, 3: +sin r */
#endif
}

/* ================================================================== */
/* log                                                                 */
/* ================================================================== */
/* log(xh + xl), xl a tiny correction (|xl| <= ulp(xh)) */
S5_HD S5_NOINL double log_core(double x, double xl)
{
    if (!(x > 0.0)) return (x == 0.0) ? -inf_() : (x - x) / (x - x);
    if (x > 1.7976931348623157e308) return x;
    int64_t b = bits_of(x);
    int e = (int)(b >> 52);
    if (e == 0) { x *= 18014398509481984.0; xl *= 18014398509481984.0; b = bits_of(x); e = (int)(b >> 52) - 54; }   /* subnormal: * 2^54 */
    e -= 1023;
    int i = (int)((b >> 45) & 127);
    double m = from_bits((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    double scale = from_bits(b & 0x7ff0000000000000LL);  /* x / m: the power of two of x's binade, taken from its exponent field (x is normal here) */
    if (i >= CRM_LOG_SPLIT) { m *= 0.5; e += 1; scale += scale; }
    double c = CRM_LOGTAB(i, 0), lch = CRM_LOGTAB(i, 1), lcl = CRM_LOGTAB(i, 2);
    double r = fma_(m, c, -1.0);                       /* exact, |r| < 2^-7 */
    dd r2 = two_prod(r, r);
    double p = KC(LP8);
    p = fma_(p, r, KC(LP7)); p = fma_(p, r, KC(LP6)); p = fma_(p, r, KC(LP5)); p = fma_(p, r, KC(LP4));
    p = fma_(p, r, KC(LP3)); p = fma_(p, r, KC(LP2)); p = fma_(p, r, KC(LP1));
    /* 1/3 + r*p with 1/3 in double-double so the r^3/3 term is good to ~2^-70 of the result */
    double t3h = fma_(p, r, KC(THIRD_H));
    double r3  = mul_(r, r2.h);
    double t3  = fma_(r3, t3h, mul_(r3, KC(THIRD_L)));
    if (xl != 0.0) t3 = add_(t3, (xl / scale) * c / (1.0 + r));   /* d/dr log1p(r) * dr */
    double ed = (double)e;
    dd s1 = two_sum(mul_(ed, KC(LN2_HEAD)), lch);       /* e*HEAD exact (42-bit head) */
    dd s2 = two_sum(s1.h, r);
    dd s3 = two_sum(s2.h, -0.5 * r2.h);
    double low = add_(add_(s1.l, s2.l), s3.l);
    low = add_(low, fma_(ed, KC(LN2_MID), lcl));
    low = add_(low, fma_(-0.5, r2.l, t3));
    low = fma_(ed, KC(LN2_TAIL), low);
    return add_(s3.h, low);
}
S5_HD S5_INL double cr_log(double x) { return log_core(x, 0.0); }
S5_HD S5_INL double cr_log1p(double x)
{
    if (fabs(x) < 5.551115123125783e-17) return x;       /* 2^-54 */
    dd s = two_sum(1.0, x);
    return log_core(s.h, s.l);
}

/* ================================================================== */
/* atan family                                                         */
/* ================================================================== */
S5_HD S5_INL dd atan_kernel(dd x)                      /* 0 <= x <= 1 (+eps) */
{
    int j = (int)(x.h * 64.0 + 0.5);
    if (j > 64) j = 64;
    dd t;
    if (j == 0) {
        t = x;
    } else {
        double c = (double)j * 0.015625;
        dd num = two_sum(add_(x.h, -c), x.l);           /* x.h - c exact */
        dd xc = two_prod(x.h, c);
        xc.l = fma_(x.l, c, xc.l);
        dd den = fast_two_sum(1.0, xc.h);
        den.l = add_(den.l, xc.l);
        den = fast_two_sum(den.h, den.l);
        t = dd_div(num, den);
    }
    double th = t.h, z = mul_(th, th);
    double p = KC(AT5);
    p = fma_(p, z, KC(AT4)); p = fma_(p, z, KC(AT3)); p = fma_(p, z, KC(AT2));
    p = fma_(p, z, KC(AT1)); p = fma_(p, z, KC(AT0));
    double w = mul_(mul_(th, z), p);
    if (j == 0) return fast_two_sum(t.h, add_(t.l, w));
    dd s = two_sum(CRM_ATANTAB(j, 0), t.h);
    s.l = add_(s.l, add_(add_(CRM_ATANTAB(j, 1), t.l), w));
    return fast_two_sum(s.h, s.l);
}
/* atan(num/den) for non-negative double-double num, den (not both zero), result in [0, pi/2] */
S5_HD S5_INL dd atan_ratio(dd num, dd den)
{
    if (num.h > den.h || (num.h == den.h && num.l > den.l)) {
        dd a = atan_kernel(dd_div(den, num));
        return dd_add(dd{KC(PI_2_H), KC(PI_2_L)}, dd_neg(a));
    }
    return atan_kernel(dd_div(num, den));
}
S5_HD S5_NOINL double cr_atan2(double y, double x)
{
    if (x != x || y != y) return x + y;
    double ay = fabs(y), ax = fabs(x);
    bool xneg = bits_of(x) < 0;
    if (ay == 0.0) return xneg ? copysign(KC(PI_H), y) : y;
    if (ax == 0.0) return copysign(KC(PI_2_H), y);
    const double INF = inf_();
    if (ax == INF || ay == INF) {
        double v;
        if (ax == INF && ay == INF) v = xneg ? 2.356194490192345 : 0.7853981633974483;
        else if (ax == INF)         v = xneg ? KC(PI_H) : 0.0;
        else                        v = KC(PI_2_H);
        return copysign(v, y);
    }
    dd a;
    if (ay > ax) {
        a = atan_kernel(dd_div_dd_d(ax, ay));
        a = dd_add(dd{KC(PI_2_H), KC(PI_2_L)}, dd_neg(a));
    } else {
        a = atan_kernel(dd_div_dd_d(ay, ax));
    }
    if (xneg) a = dd_add(dd{KC(PI_H), KC(PI_L)}, dd_neg(a));
    return copysign(a.h, y);
}
S5_HD S5_NOINL double cr_atan(double x)
{
    if (x != x) return x;
    double ax = fabs(x);
    if (ax < 7.450580596923828e-09) return x;
    dd a;
    if (ax > 1.0) {
        if (ax > 1.7976931348623157e308) return copysign(KC(PI_2_H), x);
        a = atan_kernel(dd_div_dd_d(1.0, ax));
        a = dd_add(dd{KC(PI_2_H), KC(PI_2_L)}, dd_neg(a));
    } else {
        a = atan_kernel(dd{ax, 0.0});
    }
    return copysign(a.h, x);
}
S5_HD S5_NOINL double cr_acos(double x);
/* The stepper alternates m -> theta = acos(m) -> theta + dtheta -> m' = cos(theta + dtheta) every step (sim5raytrace.c:177), so the acos
 * of a step is the inverse of the cos of the step before.  AngCarry keeps what that cos knew: its argument th and the low word of its
 * double-double result, i.e. cos(th) = m + lo to ~2^-68.  Then acos(m) = th + lo / sqrt(1 - m^2) + O(lo^2): one FMA, one reciprocal
 * square root and a handful of operations instead of a double-double sqrt, two double-double divisions and the arctangent kernel
 * (cr_acos was 28 % of the stepwise kernel's stall samples, profiles/r04b_hotspots_stepwise.txt).  The shortcut is taken only when the
 * rounding of th + correction is beyond doubt (both ends of the error interval round to the same double); otherwise -- near a rounding
 * boundary, near the poles, after a restart -- cr_acos runs.  Either way the value is the correctly rounded one that cr_acos returns. */
#if defined(S5_STEP_STATS) && !defined(__CUDA_ARCH__)
extern "C" long long s5_stat[16];                    /* tools/step_stats.cpp: how often each path of the stepper runs */
#define S5_STAT(i) __atomic_fetch_add(&crm::s5_stat[i], 1LL, __ATOMIC_RELAXED)
#else
#define S5_STAT(i) ((void)0)
#endif
struct AngCarry { double th, lo, m; };
S5_HD S5_INL void carry_reset(AngCarry* c) { c->th = 0.0; c->lo = 0.0; c->m = 2.0; }      /* m = 2 matches no cosine */
S5_HD S5_NOINL double cr_cos_carry(double x, AngCarry* c)          /* == cr_cos(x) */
{
    double ax = fabs(x);
    c->m = 2.0;
    if (!(ax < 1.0e6)) {
        if (ax != ax || ax > 1.7976931348623157e308) return x - x;
        return cos(x);
    }
    if (ax < 7.450580596923828e-09) return 1.0;
    dd r;
    int k = reduce_pio2(x, r);
    dd v = (k & 1) ? sin_kernel(r) : cos_kernel(r);
    const bool neg = (k == 1 || k == 2);
    const double m = neg ? -v.h : v.h;
    c->th = x; c->lo = neg ? -v.l : v.l; c->m = m;
    return m;
}
S5_HD S5_INL double cr_acos_carry(double m, const AngCarry* c)   /* == cr_acos(m) */
{
#if !defined(S5_NO_ANGLE_CARRY)
    if (c->m == m && c->th > 0.05 && c->th < 3.09) {
        double s2 = fma_(-m, m, 1.0);
        if (s2 > 0.00390625) {
#if defined(__CUDA_ARCH__)
            double rs = rsqrt(s2);
#else
            double rs = 1.0 / sqrt(s2);
#endif
            double cc = mul_(c->lo, rs);
            double e = fma_(mul_(fabs(m), rs), 0x1p-66, mul_(fabs(cc), 0x1p-30));
            double r1 = add_(c->th, add_(cc, e)), r2 = add_(c->th, add_(cc, -e));
            if (r1 == r2) return r1;
            S5_STAT(5);
        }
    }
#endif
    S5_STAT(4);
    return cr_acos(m);
}

/* sqrt(1 - x^2) as a double-double, 0 <= x <= 1 */
S5_HD S5_INL dd sqrt_1mx2(double ax)
{
    dd om = two_sum(1.0, -ax);
    dd op = two_sum(1.0, ax);
    return dd_sqrt(dd_mul(om, op));
}
S5_HD S5_NOINL double cr_acos(double x)
{
    double ax = fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x);
    if (ax == 1.0) return x > 0.0 ? 0.0 : KC(PI_H);
    dd s = sqrt_1mx2(ax);
    dd a = atan_ratio(s, dd{ax, 0.0});
    if (x < 0.0) a = dd_add(dd{KC(PI_H), KC(PI_L)}, dd_neg(a));
    return a.h;
}
S5_HD S5_NOINL double cr_asin(double x)
{
    double ax = fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x);
    if (ax < 7.450580596923828e-09) return x;
    if (ax == 1.0) return copysign(KC(PI_2_H), x);
    dd s = sqrt_1mx2(ax);
    dd a = atan_ratio(dd{ax, 0.0}, s);
    return copysign(a.h, x);
}

/* ================================================================== */
/* the three pow() shapes the ray path uses                            */
/* ================================================================== */
/* pow(x, 1./3.) -- note 1./3. is the double 0x3FD5555555555555 = 1/3 - 2^-54/3, NOT cbrt
 * (sim5kerr-geod.c:1004,1008).  x^(1/3-d) = cbrt(x) * (1 - d*ln x + O(d^2)). */
S5_HD S5_NOINL double cr_pow_third(double x)
{
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : (x - x) / (x - x);
    if (x > 1.7976931348623157e308) return x;
    double y0 = cbrt(x);
    dd p = two_prod(y0, y0);
    dd t = two_prod(p.h, y0);
    double r = add_(add_(add_(x, -t.h), -t.l), -mul_(p.l, y0));
    double e = r / (3.0 * p.h);
    const double delta = 1.850371707708594e-17;        /* 1/3 - (double)(1./3.) = 2^-54/3 */
    double corr = -delta * log(x) * y0;
    return add_(y0, add_(e, corr));
}
/* pow(x, 1.5)  (sim5kerr.c:1045,1139) */
S5_HD S5_INL double cr_pow_1p5(double x)
{
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : (x - x) / (x - x);
    if (x > 1.7976931348623157e308) return x;
    double s = sqrt(x);
    double sl = fma_(-s, s, x) / (s + s);              /* x - s*s is exact */
    dd p = two_prod(x, s);
    return add_(p.h, fma_(x, sl, p.l));
}
/* pow(x, 4.)  (disk-image.c:86) */
S5_HD S5_INL double cr_pow_4(double x)
{
    dd p = two_prod(x, x);
    dd q = two_prod(p.h, p.h);
    return add_(q.h, fma_(add_(p.h, p.h), p.l, q.l));
}

/* ================================================================== */
/* x87 80-bit extended arithmetic, emulated (sim5kerr-geod.c:1125-1131 uses `long double`)   */
/* A value with a 64-bit significand is kept as an unevaluated sum h + l of two doubles.     */
/* ================================================================== */
S5_HD S5_INL dd round_to_64(dd v)                       /* v normalised: h = RN(h+l) */
{
    if (v.h == 0.0 || v.l == 0.0) return v;
    int64_t b = bits_of(v.h);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    bool pow2 = (b & 0x000fffffffffffffLL) == 0;
    if (pow2 && ((v.l < 0.0) != (v.h < 0.0))) e -= 1;   /* h+l lies in the binade below h */
    double M = 1.5 * pow2i(e - 11);                     /* 1.5 * 2^52 * ulp64 */
    double lr = add_(add_(v.l, M), -M);                 /* RN-even to a multiple of ulp64 = 2^(e-63) */
    return dd{v.h, lr};
}
S5_HD S5_INL double x87_to_double(dd v) { return add_(v.h, v.l); }

/* The T-integral roots exactly as the CPU reference computes them (sim5kerr-geod.c:1124-1131):
 *   long double qla = q + l2 - a2;  X = sqrt(sqr(qla)+4.*q*a2) + qla;  m2m = X/(a2+a2);  m2p = (q+q)/X */
S5_HD S5_MID void x87_mu_roots(double q, double l2, double a2, double* m2m, double* m2p)
{
    double qla = add_(add_(q, l2), -a2);
    dd sq = round_to_64(two_prod(qla, qla));            /* fmul in extended precision */
    double fqa = mul_(mul_(4.0, q), a2);                /* double arithmetic (SSE) */
    dd sum = round_to_64(dd_add(sq, dd{fqa, 0.0}));     /* fadd extended */
    double arg = x87_to_double(sum);                    /* sqrt() takes a double */
    double sr = sqrt(arg);
    dd X = round_to_64(two_sum(sr, qla));               /* fadd extended */
    dd Xn = fast_two_sum(X.h, X.l);
    double dbla = add_(a2, a2), dblq = add_(q, q);
    *m2m = x87_to_double(round_to_64(dd_div(Xn, dd{dbla, 0.0})));
    *m2p = x87_to_double(round_to_64(dd_div(dd{dblq, 0.0}, Xn)));
}

} /* namespace crm */
#endif
