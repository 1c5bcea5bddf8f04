/*
 * raytrace.cuh -- step-wise null-geodesic integration (velocity-Verlet after Dolence+09 with a
 * fixed-point momentum solve, step-size control and an RK4 fallback) as sm_100a device code.
 *
 * Bit-for-bit behavioural twin of src/sim5raytrace.c of the reference, including its `float`
 * error bookkeeping (sim5raytrace.c:140,219; sim5raytrace.h:42) which decides accept/reject of a
 * step and therefore the whole step sequence.  The connection lives in registers (kerr.cuh Conn).
 */
#ifndef SIM5_RAYTRACE_CUH
#define SIM5_RAYTRACE_CUH

#include "geod.cuh"

namespace s5 {

#define S5_RTOPT_FLAT 1
#define S5_TINY 1e-40

struct RayData {           /* == raytrace_data, sim5raytrace.h:26-43 (144 bytes) */
    int opt_gr, opt_pol;
    double step_epsilon;
    double bh_spin, E, Q;
    Cplx WP;
    int pass, refines;
    double dk[4], df[4];
    double kt;
    float error;
};
static_assert(sizeof(RayData) == 144, "raytrace_data ABI");

S5_HD S5_INL double frac_err(double a, double b) { return fabs(b - a) / (fabs(b) + 1e-40); }

/* sim5raytrace.c:43-94 */
S5_HD S5_INL void raytrace_prepare(double bh_spin, const double x[4], const double k[4], double precision_factor, int options, RayData* rtd)
{
    rtd->opt_gr = !((options & S5_RTOPT_FLAT) == S5_RTOPT_FLAT);
    rtd->step_epsilon = sqrt(precision_factor) / 10.;
    Metric m;
    Conn G;
    if (rtd->opt_gr) { kerr_metric(bh_spin, x[1], x[2], &m); kerr_connection(bh_spin, x[1], x[2], &G); }
    else             { flat_metric(x[1], x[2], &m);          flat_connection(x[1], x[2], &G); }
    rtd->bh_spin = bh_spin;
    rtd->E = k[0] * m.g00 + k[3] * m.g03;
    rtd->Q = photon_carter_const(k, &m);
    rtd->pass = 0;
    rtd->refines = 0;
    rtd->kt = rtd->E;
    rtd->error = 0.0;
    Gamma(&G, k, k, rtd->dk);
}

/* sim5raytrace.c:250-323.  th0 = acos(x[2]) (the caller has it); *ac receives the angle behind the new x[2] */
S5_HD S5_INL void raytrace_rk4(double x[4], double k[4], double dl, RayData* rtd, double th0, crm::AngCarry* ac)
{
    Metric m;
    Conn G;
    double xp[4];
    double k1[4], dk1[4], k2[4], dk2[4], k3[4], dk3[4], k4[4], dk4[4];
    double dl_2 = 0.5 * dl;
    double kt0 = rtd->kt;
    int i;

    x[2] = th0;

#define S5_CONN_AT(xx) do { if (rtd->opt_gr) kerr_connection(rtd->bh_spin, (xx)[1], crm::cr_cos((xx)[2]), &G); \
                            else flat_connection((xx)[1], crm::cr_cos((xx)[2]), &G); } while (0)
    for (i = 0; i < 4; i++) xp[i] = x[i];
    S5_CONN_AT(xp);
    for (i = 0; i < 4; i++) k1[i] = k[i];
    Gamma(&G, k1, k1, dk1);

    for (i = 0; i < 4; i++) xp[i] = x[i] + k1[i] * dl_2;
    S5_CONN_AT(xp);
    for (i = 0; i < 4; i++) k2[i] = k[i] + dk1[i] * dl_2;
    Gamma(&G, k2, k2, dk2);

    for (i = 0; i < 4; i++) xp[i] = x[i] + k2[i] * dl_2;
    S5_CONN_AT(xp);
    for (i = 0; i < 4; i++) k3[i] = k[i] + dk2[i] * dl_2;
    Gamma(&G, k3, k3, dk3);

    for (i = 0; i < 4; i++) xp[i] = x[i] + k3[i] * dl;
    S5_CONN_AT(xp);
    for (i = 0; i < 4; i++) k4[i] = k[i] + dk3[i] * dl;
    Gamma(&G, k4, k4, dk4);
#undef S5_CONN_AT

    for (i = 0; i < 4; i++) {
        x[i] += dl / 6. * (k1[i] + 2. * k2[i] + 2. * k3[i] + k4[i]);
        k[i] += dl / 6. * (dk1[i] + 2. * dk2[i] + 2. * dk3[i] + dk4[i]);
    }
    x[2] = crm::cr_cos_carry(x[2], ac);

    if (rtd->opt_gr) kerr_connection(rtd->bh_spin, x[1], x[2], &G); else flat_connection(x[1], x[2], &G);
    Gamma(&G, k, k, rtd->dk);

    kerr_metric(rtd->bh_spin, x[1], x[2], &m);
    double kt1 = k[0] * m.g00 + k[3] * m.g03;
    rtd->error = (float)frac_err(kt1, kt0);
}

/* one step, first part: the velocity-Verlet step with its fixed-point momentum solve (sim5raytrace.c:108-226).  Returns false when
 * the step was accepted (x, k, rtd advanced).  Returns true when the reference would now fall back to RK4 (sim5raytrace.c:219-226):
 * x, k are back at their values, *step holds the step length and *th0 = acos(x[2]) for raytrace_rk4.  The split lets the lane kernel
 * run the (five times more expensive, 0.3 % of the steps) RK4 fallback for several lanes of a warp at once instead of idling 31 lanes
 * every time one lane needs it; the operations of a ray and their order are those of raytrace().
 * *ac: the angle behind x[2] if the previous step of this ray left one (crm::AngCarry) */
S5_HD S5_INL bool raytrace_verlet(double x[4], double k[4], double* step, RayData* rtd, crm::AngCarry* ac, double* th0_out)
{
    int i;
    Metric m;
    Conn G;
    double* dk = rtd->dk;
    double x_orig[4], k_orig[4];
    double xp[4], kp[4], kp_prev[4];
    double kk = 0.0, kt = rtd->kt;
    float k_frac_error;

    for (i = 0; i < 4; i++) { x_orig[i] = x[i]; k_orig[i] = k[i]; }

    double stepsize = rtd->step_epsilon / (fabs(dk[0]) / (fabs(k[0]) + S5_TINY) + fabs(dk[1]) / (fabs(k[1]) + S5_TINY) +
                                           fabs(dk[2]) / (fabs(k[2]) + S5_TINY) + fabs(dk[3]) / (fabs(k[3]) + S5_TINY) + S5_TINY);
    double dl = fmin(*step, stepsize);
    if (dl < 1e-3) dl = 1e-3;

    rtd->pass++;

    double half_dl = 0.5 * dl;
    double half_dl2 = 0.5 * dl * dl;
    xp[0] = x[0] + k[0] * dl + dk[0] * half_dl2;
    xp[1] = x[1] + k[1] * dl + dk[1] * half_dl2;
    const double th0 = crm::cr_acos_carry(x[2], ac);
    crm::AngCarry acp;
    xp[2] = crm::cr_cos_carry(th0 + (k[2] * dl + dk[2] * half_dl2), &acp);
    xp[3] = x[3] + k[3] * dl + dk[3] * half_dl2;

    for (i = 0; i < 4; i++) k[i] += dk[i] * half_dl;

    if (rtd->opt_gr) {
        kerr_metric(rtd->bh_spin, xp[1], xp[2], &m);
        kerr_connection(rtd->bh_spin, xp[1], xp[2], &G);
    } else {
        flat_metric(xp[1], xp[2], &m);
        flat_connection(xp[1], xp[2], &G);
    }

    for (i = 0; i < 4; i++) kp[i] = k[i] + dk[i] * half_dl;

    int k_iter = 0;
    do {
        k_frac_error = 0.0;
        for (i = 0; i < 4; i++) kp_prev[i] = kp[i];
        kp[0] = k[0] + k_deriv0(&G, kp_prev) * half_dl;  k_frac_error += frac_err(kp[0], kp_prev[0]);
        kp[1] = k[1] + k_deriv1(&G, kp_prev) * half_dl;  k_frac_error += frac_err(kp[1], kp_prev[1]);
        kp[2] = k[2] + k_deriv2(&G, kp_prev) * half_dl;  k_frac_error += frac_err(kp[2], kp_prev[2]);
        kp[3] = k[3] + k_deriv3(&G, kp_prev) * half_dl;  k_frac_error += frac_err(kp[3], kp_prev[3]);
        k_iter++;
    } while (k_frac_error > 1e-2 * 1e-3 && k_iter < 3);
    S5_STAT(0); S5_STAT(k_iter);

    kt = kp[0] * m.g00 + kp[3] * m.g03;
    kk = fabs(dotprod(kp, kp, &m));
    rtd->error = (float)fmax(frac_err(kt, rtd->kt), kk);
    if ((k_frac_error > 1e-2 * 1e-2) || (rtd->error > 1e-2 * 1e-2)) {
        for (i = 0; i < 4; i++) { x[i] = x_orig[i]; k[i] = k_orig[i]; }
        S5_STAT(6);
        *th0_out = th0;
        *step = dl;
        return true;
    }

    for (i = 0; i < 4; i++) { x[i] = xp[i]; k[i] = kp[i]; }
    *ac = acp;
    dk[0] = k_deriv0(&G, kp);
    dk[1] = k_deriv1(&G, kp);
    dk[2] = k_deriv2(&G, kp);
    dk[3] = k_deriv3(&G, kp);
    rtd->kt = kt;
    *step = dl;
    return false;
}
/* one step.  sim5raytrace.c:108-245 */
S5_HD S5_INL void raytrace(double x[4], double k[4], double* step, RayData* rtd, crm::AngCarry* ac)
{
    double th0;
    if (raytrace_verlet(x, k, step, rtd, ac, &th0)) raytrace_rk4(x, k, *step, rtd, th0, ac);
}
S5_HD S5_INL void raytrace(double x[4], double k[4], double* step, RayData* rtd)
{
    crm::AngCarry ac;
    crm::carry_reset(&ac);
    raytrace(x, k, step, rtd, &ac);
}

/* relative drift of the Carter constant.  sim5raytrace.c:327-343 */
S5_HD S5_INL double raytrace_error(const double x[4], const double k[4], const RayData* rtd)
{
    Metric m;
    if (rtd->opt_gr) kerr_metric(rtd->bh_spin, x[1], x[2], &m); else flat_metric(x[1], x[2], &m);
    return frac_err(rtd->Q, photon_carter_const(k, &m));
}

} /* namespace s5 */
#endif
