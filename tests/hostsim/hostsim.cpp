// hostsim.cpp -- TEST-ONLY host instantiation of the device headers in sim5_b200/csrc (.cuh files).
//
// The build container has no GPU.  The device math is written as __host__ __device__ code with
// explicit fma / non-contracted arithmetic, so compiling the same headers with g++ gives results
// bit-identical to the sm_100a build (IEEE add/mul/div/sqrt/fma on both sides); the only
// platform-dependent calls are the non-critical cbrt()/log() seeds in cr_pow_third and exp() in the
// harness torus.  This library lets `pytest -m "not gpu"` check the kernel math against the
// reference before any GPU time is spent.  It is NOT part of the product: libsim5b200.so never
// links or loads it and has no CPU path.
#include <omp.h>
#include <stdio.h>
#include <vector>
#include "../../sim5_b200/csrc/pixel.cuh"

using namespace s5;

extern "C" {

void hs_batch_libm(int op, long n, const double* a, const double* b, double* o)
{
    for (long i = 0; i < n; i++) {
        switch (op) {
            case 0: o[i] = crm::cr_sin(a[i]); break;
            case 1: o[i] = crm::cr_cos(a[i]); break;
            case 2: o[i] = crm::cr_log(a[i]); break;
            case 3: o[i] = crm::cr_atan2(a[i], b[i]); break;
            case 4: o[i] = crm::cr_acos(a[i]); break;
            case 5: o[i] = crm::cr_asin(a[i]); break;
            case 6: o[i] = crm::cr_atan(a[i]); break;
            case 7: o[i] = crm::cr_pow_third(a[i]); break;
            case 8: o[i] = crm::cr_pow_1p5(a[i]); break;
            case 9: o[i] = crm::cr_pow_4(a[i]); break;
            case 10: o[i] = exp(a[i]); break;
            default: o[i] = NAN;
        }
    }
}

/* the angle carry of the stepper (crmath.cuh): m = cos(th) through cr_cos_carry, then acos(m) through the carry and through cr_acos;
 * took[i] = 1 where the shortcut (not the full routine) produced the value */
void hs_acos_carry(long n, const double* th, double* m, double* via_carry, double* via_acos, int* took)
{
    for (long i = 0; i < n; i++) {
        crm::AngCarry c;
        crm::carry_reset(&c);
        m[i] = crm::cr_cos_carry(th[i], &c);
        via_acos[i] = crm::cr_acos(m[i]);
        via_carry[i] = crm::cr_acos_carry(m[i], &c);
        /* the shortcut is taken iff the carry is valid, away from the poles, and the rounding is beyond doubt: redo its decision */
        took[i] = 0;
        if (c.m == m[i] && c.th > 0.05 && c.th < 3.09) {
            double s2 = fma(-m[i], m[i], 1.0);
            if (s2 > 0.00390625) {
                double rs = 1.0 / sqrt(s2), cc = c.lo * rs;
                double e = fma(fabs(m[i]) * rs, 0x1p-66, fabs(cc) * 0x1p-30);
                took[i] = (c.th + (cc + e) == c.th + (cc - e)) ? 1 : 0;
            }
        }
    }
}
/* the conservative high-word form of the Carlson convergence test (fastfp.cuh) */
void hs_surely_above_tol(long n, const double* e, const double* mu, int* out)
{
    for (long i = 0; i < n; i++) out[i] = ff::surely_above_tol(ff::hi_abs(e[i]), mu[i]) ? 1 : 0;
}

void hs_mu_roots(long n, const double* q, const double* l2, const double* a2, double* m2m, double* m2p)
{
    for (long i = 0; i < n; i++) crm::x87_mu_roots(q[i], l2[i], a2[i], &m2m[i], &m2p[i]);
}

void hs_batch_rf(long n, const double* x, const double* y, const double* z, double* o) { for (long i = 0; i < n; i++) o[i] = rf(x[i], y[i], z[i]); }
void hs_batch_rd(long n, const double* x, const double* y, const double* z, double* o) { for (long i = 0; i < n; i++) o[i] = rd(x[i], y[i], z[i]); }
void hs_batch_rc(long n, const double* x, const double* y, double* o) { for (long i = 0; i < n; i++) o[i] = rc(x[i], y[i]); }
void hs_batch_rj(long n, const double* x, const double* y, const double* z, const double* p, double* o) { for (long i = 0; i < n; i++) o[i] = rj(x[i], y[i], z[i], p[i]); }
void hs_batch_rf_hi(long n, const double* x, const double* y, const double* z, double* o) { for (long i = 0; i < n; i++) o[i] = hi_domain(x[i], y[i], z[i]) ? rf_hi(x[i], y[i], z[i]) : NAN; }
void hs_batch_rj_hi(long n, const double* x, const double* y, const double* z, const double* p, double* o) { for (long i = 0; i < n; i++) o[i] = (hi_domain(x[i], y[i], z[i]) && hi_domain_p(p[i])) ? rj_hi(x[i], y[i], z[i], p[i]) : NAN; }
void hs_batch_cel_pi(long n, const double* qc, const double* pc, double* o) { for (long i = 0; i < n; i++) o[i] = cel_pi_hi(qc[i], pc[i]); }
void hs_batch_sncndn(long n, const double* u, const double* m, double* sn, double* cn, double* dn) { for (long i = 0; i < n; i++) jacobi_sncndn(u[i], m[i], &sn[i], &cn[i], &dn[i]); }

void hs_batch_integral(int op, long n, const double* v0, const double* v1, const double* v2, const double* v3, const double* v4,
                       const double* v5, const double* v6, double* o)
{
    for (long i = 0; i < n; i++) {
        double r;
        switch (op) {
            case 0: r = integral_C1(v0[i], v1[i]); break;
            case 1: r = integral_C2(v0[i], v1[i]); break;
            case 2: r = integral_C2_cos(v0[i], v1[i]); break;
            case 3: r = integral_Z2(v0[i], v1[i], v2[i], v3[i]); break;
            case 4: r = integral_Rm1(v0[i], v1[i], v2[i]); break;
            case 5: r = integral_Rm2(v0[i], v1[i], v2[i]); break;
            case 6: r = integral_R2(v0[i], v1[i], v2[i]); break;
            case 7: r = integral_R_r0_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 8: r = integral_R_r0_re_inf(v0[i], v1[i], v2[i], v3[i]); break;
            case 9: r = integral_R_r1_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 10: r = integral_R_r2_re(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 11: r = integral_T_m0(v0[i], v1[i], v2[i]); break;
            case 12: r = integral_T_m2(v0[i], v1[i], v2[i]); break;
            case 13: r = integral_R_r0_cc(v0[i], v1[i], v2[i], v3[i], v4[i]); break;
            case 14: r = integral_R_r0_cc_inf(v0[i], v1[i], v2[i], v3[i]); break;
            case 15: r = integral_R_r1_cc(v0[i], v1[i], v2[i], v3[i], v4[i], v5[i]); break;
            case 16: r = integral_R_r2_cc(v0[i], v1[i], v2[i], v3[i], v4[i], v5[i]); break;
            case 17: r = integral_R_rp_cc2(v0[i], v1[i], v2[i], v3[i], v6[i], v4[i], v5[i]); break;
            default: r = NAN;
        }
        o[i] = r;
    }
}
void hs_batch_timedelay(long n, double incl, double a, const double* alpha, const double* beta, const double* ra, const double* rb, double* o)
{
    double si, ci;
    sincos(incl, &si, &ci);
    for (long i = 0; i < n; i++) {
        Geodesic gd; int error = 0;
        o[i] = NAN;
        if (!geodesic_init_inf_sc(incl, si, ci, a, alpha[i], beta[i], &gd, &error)) continue;
        double Pa = geodesic_P_int(&gd, ra[i], 0), Pb = geodesic_P_int(&gd, rb[i], 0);
        o[i] = geodesic_timedelay(&gd, Pa, 0.0, 0.0, Pb, 0.0, 0.0);
    }
}

// geodesic_init_inf through the scalar-API path (cr_sincos of the inclination); g is the 240-byte reference struct
int hs_geodesic_init_inf(double i, double a, double alpha, double beta, void* g, int* error)
{
    return geodesic_init_inf(i, a, alpha, beta, (Geodesic*)g, error);
}

// fused azimuth (pixel.cuh azimuth_equatorial) vs the call-for-call port of geodesic_position_azm (geod.cuh) on every
// hit pixel of an image: returns the number of pixels where the two differ in any bit (must be 0)
long hs_azimuth_mismatches(const sim5_image_params* p)
{
    S5ImageConsts c;
    s5_fill_image_consts(p, &c);
    long bad = 0;
    #pragma omp parallel for schedule(dynamic, 4) reduction(+:bad)
    for (int iy = 0; iy < c.ny; iy++) {
        for (int ix = 0; ix < c.nx; ix++) {
            double alpha, beta;
            pixel_impact(c, ix, iy, &alpha, &beta);
            Geodesic gd; int err = 0; RayCache k;
            if (!init_inf_cached(c, alpha, beta, &gd, &err, &k)) continue;
            for (int order = 0; order <= 1; order++) {
                double P = crossing_cached(&gd, order, k);
                if (isnan(P)) break;
                double r = geodesic_position_rad(&gd, P);
                if (!(r >= c.rmin_emit)) continue;
                double a = azimuth_equatorial(&gd, k, r, P);
                double b = geodesic_position_azm(&gd, r, 0.0, P);
                if (memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) bad++;
                break;
            }
        }
    }
    return bad;
}

// tolerance-mode azimuth coverage: counts[0] RR hits, [1] of them handed back to the bit-faithful path, [2] RC hits, [3] handed back
void hs_fast_azimuth_coverage(const sim5_image_params* p, long* counts)
{
    S5ImageConsts c;
    s5_fill_image_consts(p, &c);
    long n0 = 0, n1 = 0, n2 = 0, n3 = 0;
    #pragma omp parallel for schedule(dynamic, 4) reduction(+:n0,n1,n2,n3)
    for (int iy = 0; iy < c.ny; iy++) {
        for (int ix = 0; ix < c.nx; ix++) {
            double alpha, beta;
            pixel_impact(c, ix, iy, &alpha, &beta);
            Geodesic gd; int err = 0; RayCache k;
            if (!init_inf_cached(c, alpha, beta, &gd, &err, &k)) continue;
            for (int order = 0; order <= c.max_order; order++) {
                double P = crossing_cached(&gd, order, k);
                if (isnan(P)) break;
                double r = geodesic_position_rad(&gd, P);
                if (!(r >= c.rmin_emit)) continue;
                AzIn z;
                az_make(&gd, k, r, P, &z);
                bool ok = true;
                if (gd.type == GEOD_TYPE_RR) { azimuth_fast_rr(z, &ok); n0++; n1 += !ok; }
                else if (gd.type == GEOD_TYPE_RC) { azimuth_fast_rc(z, &ok); n2++; n3 += !ok; }
                break;
            }
        }
    }
    counts[0] = n0; counts[1] = n1; counts[2] = n2; counts[3] = n3;
}

// mode SPECTRUM on the host instantiation: per-row partial sums added in row order (the kernel's atomics add in another order)
double hs_trace_spectrum(const sim5_image_params* p, double* spec, int nthreads)
{
    sim5_image_params q = *p;
    q.mode = SIM5_MODE_POLARIZED;
    q.outputs = SIM5_OUT_R | SIM5_OUT_G | SIM5_OUT_MUE;
    S5ImageConsts c;
    s5_fill_image_consts(&q, &c);
    int ne = p->n_energy;
    std::vector<double> E(ne), part((size_t)ne * c.ny, 0.0);
    for (int k = 0; k < ne; k++) E[k] = s5_spectrum_energy(p, k);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
    #pragma omp parallel for schedule(dynamic, 4)
    for (int lr = 0; lr < c.nrows_local; lr++) {
        int iy = s5_local_to_image_row(&c, lr);
        double* h = part.data() + (size_t)iy * ne;
        for (int ix = 0; ix < c.nx; ix++) {
            SpecHit sh; unsigned status;
            if (!spectrum_pixel(c, ix, iy, &sh, &status)) continue;
            for (int k = 0; k < ne; k++) h[k] += spectrum_term(sh, E[k]);
        }
    }
    for (int k = 0; k < ne; k++) { double s = 0.0; for (int iy = 0; iy < c.ny; iy++) s += part[(size_t)iy * ne + k]; spec[k] = s; }
    return omp_get_wtime() - t0;
}

double hs_trace_image(const sim5_image_params* p, const sim5_image_out* out, int nthreads)
{
    S5ImageConsts c;
    s5_fill_image_consts(p, &c);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
    #pragma omp parallel for schedule(dynamic, 4)
    for (int lr = 0; lr < c.nrows_local; lr++) {
        int iy = s5_local_to_image_row(&c, lr);
        for (int ix = 0; ix < c.nx; ix++) {
            PixelOut o;
            if (c.mode == SIM5_MODE_STEPWISE) {
                StepRay s;
                if (stepwise_start(c, ix, iy, &s, &o)) {
                    int cls;
                    while ((cls = stepwise_step(c, &s)) == 0) {}
                    stepwise_finish(c, &s, cls, &o);
                }
            } else if (c.mode == SIM5_MODE_SURFACE) {
                SurfRay s;
                if (SurfaceProg::start(c, ix, iy, &s, &o)) {
                    int cls;
                    while ((cls = SurfaceProg::step(c, &s)) == 0) {}
                    SurfaceProg::finish(c, &s, cls, &o);
                }
            } else {
                trace_eqplane_pixel(c, ix, iy, &o);
            }
            size_t i = (size_t)iy * c.nx + ix;
            if ((c.outputs & SIM5_OUT_R) && out->r) out->r[i] = o.r;
            if ((c.outputs & SIM5_OUT_PHI) && out->phi) out->phi[i] = o.phi;
            if ((c.outputs & SIM5_OUT_G) && out->g) out->g[i] = o.g;
            if ((c.outputs & SIM5_OUT_FLUX) && out->flux) out->flux[i] = o.flux;
            if ((c.outputs & SIM5_OUT_CHI) && out->chi) out->chi[i] = o.chi;
            if ((c.outputs & SIM5_OUT_DELTA) && out->delta) out->delta[i] = o.delta;
            if ((c.outputs & SIM5_OUT_MUE) && out->mue) out->mue[i] = o.mue;
            if ((c.outputs & SIM5_OUT_INTENSITY) && out->intensity) out->intensity[i] = o.intensity;
            if ((c.outputs & SIM5_OUT_TAU) && out->tau) out->tau[i] = o.tau;
            if ((c.outputs & SIM5_OUT_QERR) && out->qerr) out->qerr[i] = o.qerr;
            if ((c.outputs & SIM5_OUT_HEIGHT) && out->height) out->height[i] = o.height;
            if ((c.outputs & SIM5_OUT_DELAY) && out->delay) out->delay[i] = o.delay;
            if ((c.outputs & SIM5_OUT_STEPS) && out->steps) out->steps[i] = o.steps;
            if ((c.outputs & SIM5_OUT_STATUS) && out->status) out->status[i] = (uint8_t)o.status;
        }
    }
    return omp_get_wtime() - t0;
}

}

#if defined(S5_COUNT_ITERS)
// op-counting build only: the tolerance-mode Carlson counters of ellfast.cuh
extern "C" void hs_hi_counts(long long* out, int reset)
{
    for (int i = 0; i < 4; i++) { out[i] = s5::s5_hi_counts[i]; if (reset) s5::s5_hi_counts[i] = 0; }
}
#endif

#if defined(S5_AZ_DIAG)
// calibration build only: the conditioning indicator of azimuth_well_conditioned for every disk hit (0 elsewhere) and the
// tolerance-mode phi WITHOUT the guard's fallback, so tools/calibrate_azimuth_guard.py can plot deviation against indicator
extern "C" void hs_fast_azimuth_kappa(const sim5_image_params* p, double* kappa, double* phi_fast)
{
    S5ImageConsts c;
    s5_fill_image_consts(p, &c);
    #pragma omp parallel for schedule(dynamic, 4)
    for (int iy = 0; iy < c.ny; iy++) for (int ix = 0; ix < c.nx; ix++) {
        size_t i = (size_t)iy * c.nx + ix;
        kappa[i] = 0.0; phi_fast[i] = 0.0;
        double alpha, beta;
        pixel_impact(c, ix, iy, &alpha, &beta);
        Geodesic gd; int err = 0; RayCache k;
        if (!init_inf_cached(c, alpha, beta, &gd, &err, &k)) continue;
        for (int order = 0; order <= c.max_order; order++) {
            double P = crossing_cached(&gd, order, k);
            if (isnan(P)) break;
            double r = geodesic_position_rad(&gd, P);
            if (!(r >= c.rmin_emit)) continue;
            AzIn z; az_make(&gd, k, r, P, &z);
            bool ok = false; double v = NAN;
            s5_last_kappa = 0.0;
            if (gd.type == GEOD_TYPE_RR) v = azimuth_fast_rr(z, &ok);
            else if (gd.type == GEOD_TYPE_RC) v = azimuth_fast_rc(z, &ok);
            kappa[i] = s5_last_kappa;
            phi_fast[i] = v;
            break;
        }
    }
}
#endif
