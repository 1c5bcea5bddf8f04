/*
 * sim5lib.h -- drop-in header for the photon hot path of SIM5 (mbursa/sim5), B200-native build.
 *
 * Same struct layouts and the same C signatures as the reference's src/sim5lib.h aggregate for the
 * functions on the per-pixel photon path (reference headers cited per block).  A program written
 * against the reference compiles against this header and links libsim5b200.so instead of
 * lib/sim5lib.o; every call below is executed by the sm_100a device code (a one-thread launch for
 * the scalar functions) -- there is no CPU implementation behind them.  For throughput use the
 * batched entry sim5_trace_image() of sim5_b200.h, which replaces the caller's pixel loop.
 *
 * Error convention is the reference's: int TRUE/FALSE plus optional int* error, NaN for numeric
 * failures; nothing is printed per call.  Without a usable CUDA device the functions return
 * NaN / FALSE and sim5_last_error() says why.
 */
#ifndef _SIM5LIB_H
#define _SIM5LIB_H

#include <math.h>
#include <stdint.h>
#include "sim5_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sim5const.h:24-27, sim5math.h:36-65 -------------------------------------------------- */
#ifndef TINY
#define TINY                1e-40
#endif
#ifndef TRUE
#define TRUE                1
#define FALSE               0
#endif
#define PI      3.14159265359
#define PI2     6.28318530718
#define PI4     12.5663706144
#define PI_half 1.57079632679
#define sqr(a)   ((a) * (a))
#define sqr2(a)  ((a) * (a))
#define sqr3(a)  ((a) * (a) * (a))
#define sqr4(a)  ((a) * (a) * (a) * (a))
#define deg2rad(a) ((a)/180.0*M_PI)
#define rad2deg(a) ((a)*180.0/M_PI)

#if defined(__cplusplus) || defined(SIM5_NO_C99_COMPLEX)
typedef struct sim5complex { double re, im; } sim5complex;   /* same size, alignment and calling convention as double _Complex */
#else
#include <complex.h>
typedef double _Complex sim5complex;                           /* sim5math.h:63 */
#endif

/* ---- sim5kerr.h:18-32 ----------------------------------------------------------------------- */
struct sim5metric {
    double a, r, m;
    double g00;
    double g11;
    double g22;
    double g33;
    double g03;
};
typedef struct sim5metric sim5metric;

struct sim5tetrad {
    double e[4][4];
    sim5metric metric;
};
typedef struct sim5tetrad sim5tetrad;

void kerr_metric(double a, double r, double m, sim5metric *metric);
void kerr_metric_contravariant(double a, double r, double m, sim5metric *metric);
void flat_metric(double r, double m, sim5metric *metric);
void kerr_connection(double a, double r, double m, double G[4][4][4]);
void flat_connection(double r, double m, double G[4][4][4]);
void Gamma(double G[4][4][4], double U[4], double V[4], double result[4]);
double dotprod(double V1[4], double V2[4], sim5metric* m);
void vector_norm_to(double V[4], double norm, sim5metric* m);
void tetrad_zamo(sim5metric *m, sim5tetrad *t);
void tetrad_azimuthal(sim5metric *m, double Omega, sim5tetrad *t);
void tetrad_surface(sim5metric *m, double Omega, double V, double dhdr, sim5tetrad *t);
void bl2on(double Vin[4], double Vout[4], sim5tetrad* t);
void on2bl(double Vin[4], double Vout[4], sim5tetrad* t);
double r_bh(double a);
double r_ms(double a);
double OmegaK(double r, double a);
double ellK(double r, double a);                                                  /* sim5kerr.c:1050 */
double Omega_from_ell(double ell, sim5metric *m);
double ell_from_Omega(double Omega, sim5metric *m);
double gfactorK(double r, double a, double l);
void photon_momentum(double a, double r, double m, double l, double q, double r_sign, double m_sign, double k[4]);
void photon_motion_constants(double a, double r, double m, double k[4], double* L, double* Q);
double photon_carter_const(double k[4], sim5metric *metric);
void fourvelocity_azimuthal(double Omega, sim5metric *m, double U[4]);

/* ---- sim5elliptic.h:20-56 --------------------------------------------------------------------- */
double rf(double x, double y, double z);
double rd(double x, double y, double z);
double rc(double x, double y);
double rj(double x, double y, double z, double p);
double elliptic_k(double m);
double elliptic_f(double phi, double m);
double elliptic_f_cos(double cos_phi, double m);
double elliptic_f_sin(double sin_phi, double m);
double elliptic_e_cos(double cos_phi, double m);
double elliptic_e_sin(double sin_phi, double m);
double elliptic_pi_complete(double n, double m);
double elliptic_pi_cos(double cos_phi, double n, double m);
double elliptic_pi_sin(double sin_phi, double n, double m);
double jacobi_isn(double y, double emmc);
double jacobi_icn(double z, double m);
double jacobi_itn(double z, double m);
void jacobi_sncndn(double uu, double emmc, double *sn, double *cn, double *dn);
double jacobi_sn(double uu, double emmc);
double jacobi_cn(double uu, double emmc);
double jacobi_dn(double u, double m);
double integral_Z1(double a, double b, double u, double m);
double integral_R1(double a, double u, double m);
double integral_R_rp_re(double a, double b, double c, double d, double p, double X);
double integral_R_rp_re_inf(double a, double b, double c, double d, double p);
double integral_R_rp_cc2_inf(double a, double b, sim5complex c, double p, double X1);
double integral_T_mp(double a2, double b2, double p, double X);
/* the integrals behind geodesic_timedelay that the reference exports (sim5elliptic.h:41-55; sim5elliptic.c:825-1139) */
double integral_R_r0_re(double a, double b, double c, double d, double X);
double integral_R_r0_re_inf(double a, double b, double c, double d);
double integral_R_r0_cc(double a, double b, sim5complex c, double X);
double integral_R_r0_cc_inf(double a, double b, sim5complex c);
double integral_R_r1_re(double a, double b, double c, double d, double X);
double integral_R_r1_cc(double a, double b, sim5complex c, double X1, double X2);
double integral_R_r2_re(double a, double b, double c, double d, double X);
double integral_R_r2_cc(double a, double b, sim5complex c, double X1, double X2);
double integral_R_rp_cc2(double a, double b, sim5complex c, double p, double X1, double X2);
double integral_T_m0(double a2, double b2, double X);
double integral_T_m2(double a2, double b2, double X);

/* ---- sim5kerr-geod.h:19-84 -------------------------------------------------------------------- */
#define GEOD_TYPE_RR               40
#define GEOD_TYPE_RR_DBL           41
#define GEOD_TYPE_RR_BH            42
#define GEOD_TYPE_RC                2
#define GEOD_TYPE_CC                0

#define GD_OK                           0
#define GD_ERROR_Q_ZERO                 1
#define GD_ERROR_BOUND_GEODESIC         2
#define GD_ERROR_UNKNOWN_SOLUTION       3
#define GD_ERROR_TYPE_RR_DOUBLE         4
#define GD_ERROR_TYPE_CC                5
#define GD_ERROR_Q_RANGE                7
#define GD_ERROR_MUPLUS_RANGE           8
#define GD_ERROR_MU0_RANGE              9
#define GD_ERROR_MM_RANGE              10
#define GD_ERROR_INCL_RANGE            11
#define GD_ERROR_SPIN_RANGE            12

typedef struct geodesic {
    double a;
    double alpha;
    double beta;
    double incl;
    double cos_i;
    double l;
    double q;
    sim5complex r1,r2,r3,r4;
    int    nrr;
    int    type;
    double m2p,m2m,mm,mK;
    double rp;
    double dmdp_inf;
    double Rpc;
    double Tpp;
    double Tip;
    double k[4];
    double p;
} geodesic;

int geodesic_init_inf(double i, double a, double alpha, double beta, geodesic *g, int *error);
int geodesic_init_src(double a, double r, double m, double k[4], int bpa, geodesic *g, int *error);
double geodesic_P_int(geodesic *g, double r, int bpa);
void geodesic_position(geodesic *g, double P, double x[4]);
double geodesic_position_rad(geodesic *g, double P);
double geodesic_position_pol(geodesic *g, double P);
double geodesic_position_pol_sign_k_theta(geodesic *g, double P);
double geodesic_position_azm(geodesic *g, double r, double m, double P);
double geodesic_dm_sign(geodesic *g, double P);
void geodesic_momentum(geodesic *g, double P, double r, double m, double k[]);
double geodesic_find_midplane_crossing(geodesic *g, int order);
void geodesic_follow(geodesic *g, double step, double *P, double *r, double *m, int *status);
double geodesic_timedelay(geodesic *g, double P1, double r1, double m1, double P2, double r2, double m2);   /* sim5kerr-geod.c:559 */

/* ---- sim5raytrace.h:21-53 --------------------------------------------------------------------- */
#define RTOPT_NONE              0
#define RTOPT_FLAT              1
#define RTOPT_POLARIZATION      2

typedef struct raytrace_data {
    int opt_gr;
    int opt_pol;
    double step_epsilon;
    double bh_spin;
    double E;
    double Q;
    sim5complex WP;
    int pass;
    int refines;
    double dk[4];
    double df[4];
    double kt;
    float error;
} raytrace_data;

void raytrace_prepare(double bh_spin, double x[4], double k[4], double presision_factor, int options, raytrace_data* rtd);
void raytrace(double x[4], double k[4], double *step, raytrace_data* rtd);
double raytrace_error(double x[4], double k[4], raytrace_data* rtd);

/* ---- sim5polarization.h:19-23 ----------------------------------------------------------------- */
void polarization_vector(double k[4], sim5complex wp, sim5metric *metric, double f[4]);
sim5complex polarization_constant(double k[4], double f[4], sim5metric *metric);
sim5complex polarization_constant_infinity(double a, double alpha, double beta, double incl);
double polarization_angle_rotation(double a, double inc, double alpha, double beta, sim5complex kappa);

/* ---- sim5disk-nt.h:23-34 (only the per-pixel flux path) --------------------------------------- */
int disk_nt_setup(double M, double a, double mdot_or_L, double alpha, int options);
double disk_nt_r_min(void);
double disk_nt_flux(double r);

#ifdef __cplusplus
}
#endif

#endif
