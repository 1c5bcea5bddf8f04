/*
 * image_consts.h -- per-image constants, computed ONCE on the host with the host libm (glibc),
 * i.e. with exactly the arithmetic the reference uses for them, and then broadcast to the kernels
 * (they are staged in shared memory by every CTA).
 *
 * Hoisting is bit-exact: each field is a pure function of the image parameters, spelled in the
 * reference's operation order (citations per field).  The kernels never evaluate cbrt/acos of
 * per-image quantities themselves.
 */
#ifndef SIM5_IMAGE_CONSTS_H
#define SIM5_IMAGE_CONSTS_H

#include <math.h>
#include <string.h>
#include "sim5_b200.h"

struct S5ImageConsts {
    /* geometry */
    double a;            /* bh_spin as given */
    double incl, sin_i, cos_i;     /* sincos(incl) of the host libm, sim5kerr-geod.c:73,78 */
    double rmax, aspect; /* aspect = (double)ny/(double)nx, disk-image.c:58 */
    double rmin_emit;    /* r_ms(a) unless overridden, disk-image.c:41,83 */
    double r_bh;
    int nx, ny, row_begin, row_end;
    int nrows_local, split_count, split_index, split_rows;   /* rows this call traces and how they map to image rows */
    int max_order, mode;
    unsigned outputs, flags;
    /* Novikov-Thorne flux, sim5disk-nt.c:109-146 with the float statics of :27-32 */
    double nt_rms;       /* (double)(float)(disk_nt_r_min()) */
    double nt_a;         /* (double)(float)a */
    double nt_x0, nt_x1, nt_x2, nt_x3;
    double nt_k0;        /* 1.5*a */
    double nt_k1, nt_k2, nt_k3;   /* 3.*sqr(xi-a)/(xi*(xi-xj)*(xi-xk)) */
    double nt_d1, nt_d2, nt_d3;   /* x0-xi */
    double nt_4pi;       /* 4.*M_PI */
    double nt_2a;        /* 2.*a */
    double nt_mdot, nt_mass;      /* (double)(float) */
    /* stepwise */
    double pf, r_start, step_max, rh_stop, rout_stop;
    int max_steps, pad0;
    double torus_rc, torus_w, torus_h, torus_ell, torus_j0, torus_k0;
    /* histogram */
    double g_min, g_max, da, db;  /* pixel size in alpha and beta (histogram weight F*g^4*da*db) */
    int n_bins, pad1;
    /* spectrum (sim5radiation.c:56-78 blackbody(), constants of sim5const.h) */
    double bb1;          /* 2 h / c^2 / hardf^4 * kev2freq^4 */
    double bb2;          /* h kev2freq / (k_B hardf)  -- BB2 of blackbody() is bb2 / T */
    int n_energy, spec_limb;
    /* surface finder (harness surface of sim5_b200.h; python/sim5diskraytrace.py:259-275) */
    double surf_hr, surf_rin;
    double surf_cos_it;  /* cos(incl + atan(H(1e6)/1e6)), host libm */
    int surf_flat, pad2; /* H(1e5) == 0 */
    /* SIM5_OUT_DELAY */
    double delay_r_ref;
    /* Chandrasekhar table (harness, sim5_b200.h) */
    double chandra[SIM5_CHANDRA_N];
};

/* sim5kerr.c:993-1004 (sqrt3 == cbrt, sim5math.h:45) */
static inline double s5_host_r_ms(double a)
{
    double z1 = 1. + cbrt(1. - a * a) * (cbrt(1. + a) + cbrt(1. - a));
    double z2 = sqrt(3. * (a * a) + (z1 * z1));
    return 3. + z2 - sqrt((3. - z1) * (3. + z1 + 2. * z2));
}
static inline double s5_host_r_bh(double a) { return 1. + sqrt(1. - a * a); }

/* sim5disk-nt.c:90-105, evaluated on the float-truncated spin as disk_nt_setup leaves it */
static inline double s5_host_disk_nt_r_min(double a)
{
    double sga = (a >= 0.0) ? +1. : -1.;
    double z1 = 1. + pow(1. - a * a, 1. / 3.) * (pow(1. + a, 1. / 3.) + pow(1. - a, 1. / 3.));
    double z2 = sqrt(3. * a * a + z1 * z1);
    double r0 = 3. + z2 - sga * sqrt((3. - z1) * (3. + z1 + 2. * z2));
    return r0 + 1e-3;
}

/* image row of local row lr (contiguous when split_count == 1) */
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int s5_local_to_image_row(const S5ImageConsts* c, int lr)
{
    return c->row_begin + ((lr / c->split_rows) * c->split_count + c->split_index) * c->split_rows + lr % c->split_rows;
}

static inline void s5_fill_image_consts(const sim5_image_params* p, S5ImageConsts* c)
{
    memset(c, 0, sizeof(*c));
    c->a = p->bh_spin;
    c->incl = p->incl;
    sincos(p->incl, &c->sin_i, &c->cos_i);
    c->rmax = p->rmax;
    c->aspect = (double)p->ny / (double)p->nx;
    c->rmin_emit = (p->r_emit_min > 0.0) ? p->r_emit_min : s5_host_r_ms(p->bh_spin);
    c->r_bh = s5_host_r_bh(p->bh_spin);
    c->nx = p->nx; c->ny = p->ny;
    c->row_begin = p->row_begin; c->row_end = p->row_end;
    if (c->row_begin == 0 && c->row_end == 0) c->row_end = p->ny;
    c->split_count = p->split_count > 1 ? p->split_count : 1;
    c->split_index = p->split_count > 1 ? p->split_index : 0;
    c->split_rows  = (p->split_count > 1 && p->split_rows > 0) ? p->split_rows : 1;
    c->nrows_local = (c->row_end - c->row_begin) / c->split_count;
    c->max_order = p->max_order;
    c->mode = p->mode;
    c->outputs = p->outputs;
    c->flags = p->flags;

    /* disk_nt_setup(M, a, mdot, alpha, 0): float statics */
    float f_mass = (float)p->disk_mass;
    float f_spin = (float)p->bh_spin;
    float f_mdot = (float)p->disk_mdot;
    double as = (double)f_spin;
    float f_rms = (float)s5_host_disk_nt_r_min(as);
    c->nt_rms = (double)f_rms;
    c->nt_a = as;
    c->nt_x0 = sqrt((double)f_rms);
    c->nt_x1 = +2. * cos(1. / 3. * acos(as) - M_PI / 3.);
    c->nt_x2 = +2. * cos(1. / 3. * acos(as) + M_PI / 3.);
    c->nt_x3 = -2. * cos(1. / 3. * acos(as));
    {
        double x1 = c->nt_x1, x2 = c->nt_x2, x3 = c->nt_x3, x0 = c->nt_x0;
        c->nt_k0 = 1.5 * as;
        c->nt_k1 = 3. * ((x1 - as) * (x1 - as)) / (x1 * (x1 - x2) * (x1 - x3));
        c->nt_k2 = 3. * ((x2 - as) * (x2 - as)) / (x2 * (x2 - x1) * (x2 - x3));
        c->nt_k3 = 3. * ((x3 - as) * (x3 - as)) / (x3 * (x3 - x1) * (x3 - x2));
        c->nt_d1 = x0 - x1; c->nt_d2 = x0 - x2; c->nt_d3 = x0 - x3;
    }
    c->nt_4pi = 4. * M_PI;
    c->nt_2a = 2. * as;
    c->nt_mdot = (double)f_mdot;
    c->nt_mass = (double)f_mass;

    c->pf = p->precision_factor;
    c->r_start = p->r_start;
    c->step_max = p->step_max;
    c->rh_stop = 1.05 * s5_host_r_bh(p->bh_spin);
    c->rout_stop = 1.01 * p->r_start;
    c->max_steps = p->max_steps;
    c->torus_rc = p->torus_rc; c->torus_w = p->torus_w; c->torus_h = p->torus_h;
    c->torus_ell = p->torus_ell; c->torus_j0 = p->torus_j0; c->torus_k0 = p->torus_k0;

    c->g_min = p->g_min; c->g_max = p->g_max; c->n_bins = p->n_bins;
    {
        c->da = 2.0 * p->rmax / (double)p->nx;
        c->db = 2.0 * p->rmax * ((double)p->ny / (double)p->nx) / (double)p->ny;
    }
    {
        /* planck_h, speed_of_light, boltzmann_k, kev2freq of sim5const.h:33-87; expression order of sim5radiation.c:73-75 */
        const double planck_h = 6.626069e-27, speed_of_light = 2.997925e+10, boltzmann_k = 1.380650e-16, kev2freq = 2.417990e+17;
        double hf = p->spec_hardf > 0.0 ? p->spec_hardf : 1.0;
        c->bb1 = 2.0 * planck_h / (speed_of_light * speed_of_light) / (hf * hf * hf * hf) * (kev2freq * kev2freq * kev2freq * kev2freq);
        c->bb2 = (planck_h * kev2freq) / (boltzmann_k * hf);
        c->n_energy = p->n_energy;
        c->spec_limb = p->spec_limb;
    }
    {
        c->surf_hr = p->surf_hr;
        c->surf_rin = (p->surf_rin > 0.0) ? p->surf_rin : s5_host_r_ms(p->bh_spin);
        double h6 = (1e6 > c->surf_rin) ? c->surf_hr * ((1e6 - c->surf_rin) * (1e6 - c->surf_rin)) / 1e6 : 0.0;
        double h5 = (1e5 > c->surf_rin) ? c->surf_hr * ((1e5 - c->surf_rin) * (1e5 - c->surf_rin)) / 1e5 : 0.0;
        c->surf_cos_it = cos(p->incl + atan(h6 / 1e6));
        c->surf_flat = (h5 == 0.0) ? 1 : 0;
        c->delay_r_ref = p->delay_r_ref;
    }
    for (int i = 0; i < SIM5_CHANDRA_N; i++) c->chandra[i] = SIM5_CHANDRA_DELTA[i];
}

/* detector energies of a SPECTRUM call [keV], host libm: shared by the library, the reference driver and the oracle */
static inline double s5_spectrum_energy(const sim5_image_params* p, int k)
{
    if (p->n_energy <= 1) return p->e_min_kev;
    return p->e_min_kev * pow(p->e_max_kev / p->e_min_kev, (double)k / (double)(p->n_energy - 1));
}

#endif
