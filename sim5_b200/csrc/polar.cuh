/*
 * polar.cuh -- Walker-Penrose polarization transport as sm_100a device code.
 * Bit-for-bit behavioural twin of src/sim5polarization.c of the reference (complex numbers spelled
 * out as (re, im) pairs; operation order follows the reference).
 */
#ifndef SIM5_POLAR_CUH
#define SIM5_POLAR_CUH

#include "geod.cuh"

namespace s5 {

/* kappa = (A1 - i A2)(r - i a cos(theta)), Connors, Piran & Stark (1980).  sim5polarization.c:144-168 */
S5_HD S5_INL Cplx polarization_constant(const double k[4], const double f[4], const Metric* g)
{
    double a = g->a, m = g->m, r = g->r;
    double A1 = (k[0] * f[1] - k[1] * f[0]) + a * (1. - m * m) * (k[1] * f[3] - k[3] * f[1]);
    double A2 = sqrt(1. - m * m) * ((r * r + a * a) * (k[3] * f[2] - k[2] * f[3]) - a * (k[0] * f[2] - k[2] * f[0]));
    double wp1 = +r * A1 - a * m * A2;
    double wp2 = -r * A2 - a * m * A1;
    return Cplx{wp1, wp2};
}

/* polarization vector (f^0 = 0 gauge) from kappa.  sim5polarization.c:13-105 */
S5_HD S5_INL void polarization_vector(const double k[4], Cplx wp, const Metric* g, double f[4])
{
    double a = g->a, m = g->m, r = g->r;
    double s = sqrt(1.0 - m * m);
    double ra2 = r * r + a * a;
    double r2 = r * r;
    double a2 = a * a;
    double s2 = 1.0 - m * m;
    if (s < 1e-12) {
        s = 1e-12;
        s2 = 1e-24;
        m = 1.0 - 0.5 * s;
    }
    double A1 = (+r * wp.re - a * m * wp.im) / (r * r + a * a * m * m);
    double A2 = (-r * wp.im - a * m * wp.re) / (r * r + a * a * m * m);
    f[0] = 0.0;
    f[3] = (
               +g->g11 * A1 * k[1] * (s * r2 * k[3] + s * a2 * k[3] - s * a * k[0])
               + g->g22 * A2 * k[2] * (k[0] - a * s2 * k[3])
           ) / (
               +sq(k[0]) * g->g33 * (s * k[3] * a)
               + sq(k[0]) * g->g03 * (s * k[0] * a - s * r2 * k[3] - s * a2 * k[3] - a2 * s * s2 * k[3])
               + sq(k[1]) * g->g11 * a * s * s2 * (+r2 * k[3] + a2 * k[3] - a * k[0])
               + sq(k[2]) * g->g22 * (a2 * a * s * s2 * k[3] + r2 * a * s * s2 * k[3] - s * r2 * k[0] - s * a2 * k[0])
               + sq(k[3]) * g->g33 * s * (k[3] * a * s2 * r2 + k[3] * a2 * a * s2 - k[0] * r2 - k[0] * a2 - a2 * s2 * k[0])
               + sq(k[3]) * g->g03 * a * s * s2 * (r2 * k[0] + a2 * k[0])
           );
    f[1] = (A1 - a * s * s * k[1] * f[3]) / (k[0] - a * s * s * k[3]);
    f[2] = (A2 + s * k[2] * f[3] * ra2) / (s * k[3] * ra2 - s * a * k[0]);
    vector_norm_to(f, 1.0, g);
}

/* sim5polarization.c:248-268; sin_incl = sin(incl) */
S5_HD S5_INL Cplx polarization_constant_infinity_s(double a, double alpha, double beta, double sin_incl)
{
    double gamma = -alpha - a * sin_incl;
    return Cplx{-gamma, -beta};
}
/* rotation of the polarization angle between emitter and observer.  sim5polarization.c:271-285 */
S5_HD S5_INL double polarization_angle_rotation_s(double a, double sin_inc, double alpha, double beta, Cplx kappa)
{
    double kappa1 = kappa.re, kappa2 = kappa.im;
    double S = -alpha - a * sin_inc;
    double T = +beta;
    double X = (-S * kappa2 - T * kappa1) / (S * S + T * T);
    double Y = (-S * kappa1 + T * kappa2) / (S * S + T * T);
    return cr_atan2(Y, X);
}

} /* namespace s5 */
#endif
