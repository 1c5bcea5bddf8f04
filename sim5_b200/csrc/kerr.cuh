/*
 * kerr.cuh -- Kerr metric, connection, tetrads, g-factor and photon momentum as sm_100a device code.
 *
 * Bit-for-bit behavioural twin of the listed functions of the reference's src/sim5kerr.c (operation
 * order follows the reference expressions; see elliptic.cuh for the contract).  The connection is
 * kept as its 20 structurally non-zero entries in registers instead of a zero-filled G[4][4][4]
 * (sim5kerr.c:232-316 memsets 512 B per call); sums over it visit the entries in the reference's
 * (a,b) order, and skipping an exact-zero term does not change an IEEE sum.
 */
#ifndef SIM5_KERR_CUH
#define SIM5_KERR_CUH

#include "elliptic.cuh"

namespace s5 {

struct Metric {            /* == sim5metric, sim5kerr.h:18-26 */
    double a, r, m;
    double g00, g11, g22, g33, g03;
};
struct Tetrad {            /* == sim5tetrad, sim5kerr.h:28-32 */
    double e[4][4];
    Metric metric;
};

/* sim5kerr.c:74-101 */
S5_HD S5_INL void kerr_metric(double a, double r, double m, Metric* g)
{
    double r2 = sq(r), a2 = sq(a), m2 = sq(m);
    double S = r2 + a2 * m2;
    double s2_S = (1.0 - m2) / S;
    g->a = a; g->r = r; g->m = m;
    g->g00 = -1. + 2.0 * r / S;
    g->g11 = S / (r2 - 2. * r + a2);
    g->g22 = S;
    g->g33 = ((a2 + r2) * S + 2. * r * a2 * s2_S * S) * s2_S;
    g->g03 = -2. * a * r * s2_S;
}
/* sim5kerr.c:105-131 */
S5_HD S5_INL void kerr_metric_contravariant(double a, double r, double m, Metric* g)
{
    double r2 = sq(r), a2 = sq(a), m2 = sq(m);
    double S = r2 + a2 * m2;
    double SD = S * (r2 - 2. * r + a2);
    g->a = a; g->r = r; g->m = m;
    g->g00 = -sq(r2 + a2) / SD + a2 * (1. - m2) / S;
    g->g11 = (r2 - 2. * r + a2) / S;
    g->g22 = 1. / S;
    g->g33 = 1. / S / (1. - m2) - a2 / SD;
    g->g03 = -2. * a * r / SD;
}
/* sim5kerr.c:30-48 */
S5_HD S5_INL void flat_metric(double r, double m, Metric* g)
{
    g->a = 0.0; g->r = r; g->m = m;
    g->g00 = -1.0; g->g11 = +1.0; g->g22 = +r * r; g->g33 = +r * r * (1. - m * m); g->g03 = 0.0;
}

/* the 20 non-zero Christoffel entries, symmetric pairs pre-doubled as in the reference */
struct Conn {
    double g001, g002, g013, g023;
    double g100, g103, g111, g112, g122, g133;
    double g200, g203, g211, g212, g222, g233;
    double g301, g302, g313, g323;
};
/* sim5kerr.c:232-316 */
S5_HD S5_MID void kerr_connection(double a, double r, double m, Conn* G)
{
    double rS = 2.0 * r;
    double s = sqrt(1. - m * m);
    double cs = s * m;
    double c2 = m * m;
    double s2 = s * s;
    double cc = c2 - s2;
    double CC = 8. * c2 * c2 - 8. * c2 + 1.;
    double a2 = a * a;
    double a4 = a2 * a2;
    double a2cc = a2 * cc;
    double a2c2 = a2 * c2;
    double a2cs = a2 * cs;
    double a4CC = a4 * CC;
    double r2 = r * r;
    double r3 = r2 * r;
    double r4 = r2 * r2;
    double a2r2 = a2 * r2;
    double a2_r2 = a2 + r2;
    double R = sq(a2 + 2. * r2 + a2cc);
    double D = r2 - 2. * r + a2;
    double S = r2 + a2c2;
    double S_1 = 1. / S;
    double S_3 = 1. / (S * S * S);
    double D_1 = 1. / D;
    double R_1 = 1. / R;
    double m_s = m / s;
    double DR_1 = D_1 * R_1;
    double DS_1 = D_1 * S_1;
    double dbl_r2 = 2. * r2;

    G->g001 = 2.0 * 4.0 * (a2_r2) * (r2 - a2c2) * DR_1;
    G->g002 = 2.0 * -4.0 * a2cs * rS * R_1;
    G->g013 = 2.0 * 2.0 * a * s2 * (a4 - 3. * a2r2 - 6. * r4 + a2cc * (a2 - r2)) * DR_1;
    G->g023 = -G->g002 * s2 * a;

    G->g100 = D * (r2 - a2c2) * S_3;
    G->g103 = -2.0 * G->g100 * a * s2;
    G->g111 = (r * (a2 - r) + a2 * (1. - r) * c2) * DS_1;
    G->g112 = -2.0 * a2cs * S_1;
    G->g122 = -r * D * S_1;
    G->g133 = -D * s2 * (2. * a2c2 * r3 + r2 * r3 + a2 * a2c2 * s2 + a2c2 * a2c2 * r - a2r2 * s2) * S_3;

    G->g200 = -2.0 * r * a2cs * S_3;
    G->g203 = 2.0 * -G->g200 * a2_r2 / a;
    G->g211 = +a2cs * DS_1;
    G->g212 = 2.0 * r * S_1;
    G->g222 = -a2cs * S_1;
    G->g233 = -cs * (a2_r2 * S * S + a2 * s2 * rS * (a2_r2 + S)) * S_3;

    G->g301 = 2.0 * a * (r2 - a2c2) * DS_1 * S_1;
    G->g302 = 2.0 * -4.0 * a * rS * m_s * R_1;
    G->g313 = (a4 + 3. * a4 * r - 12. * a2r2 + 8. * a2 * r3 -
               16. * r4 + 8. * r2 * r3 + 4. * r * (dbl_r2 - r + a2) * a2cc -
               a4CC * (1. - r)) * DR_1;
    G->g323 = ((3. * a4 + 8. * a2 * r + 8. * a2r2 + 8. * r4 +
                4. * (dbl_r2 - 2. * r + a2) * a2cc + a4CC) * m_s) * R_1;
}
/* sim5kerr.c:198-228 (flat space), same container */
S5_HD S5_INL void flat_connection(double r, double m, Conn* G)
{
    double s = sqrt(1. - m * m);
    *G = Conn{};
    G->g122 = -r;
    G->g133 = -r * s * s;
    G->g212 = 2.0 * 1. / r;
    G->g233 = -m * s;
    G->g313 = 2.0 * 1. / r;
    G->g323 = 2.0 * m / s;
}
S5_HD S5_INL void conn_to_array(const Conn* G, double* A /* [4][4][4] */)
{
    for (int i = 0; i < 64; i++) A[i] = 0.0;
#define S5_G(i, j, k) A[(i) * 16 + (j) * 4 + (k)]
    S5_G(0,0,1) = G->g001; S5_G(0,0,2) = G->g002; S5_G(0,1,3) = G->g013; S5_G(0,2,3) = G->g023;
    S5_G(1,0,0) = G->g100; S5_G(1,0,3) = G->g103; S5_G(1,1,1) = G->g111; S5_G(1,1,2) = G->g112; S5_G(1,2,2) = G->g122; S5_G(1,3,3) = G->g133;
    S5_G(2,0,0) = G->g200; S5_G(2,0,3) = G->g203; S5_G(2,1,1) = G->g211; S5_G(2,1,2) = G->g212; S5_G(2,2,2) = G->g222; S5_G(2,3,3) = G->g233;
    S5_G(3,0,1) = G->g301; S5_G(3,0,2) = G->g302; S5_G(3,1,3) = G->g313; S5_G(3,2,3) = G->g323;
#undef S5_G
}

/* -G^i_{jk} U^j V^k, symmetrised.  sim5kerr.c:421-440; term order = (j,k) lexicographic, j<=k */
#define S5_GT(Gc, j, k) (0.5 * (Gc) * (U[j] * V[k] + U[k] * V[j]))
S5_HD S5_INL void Gamma(const Conn* G, const double U[4], const double V[4], double res[4])
{
    double t;
    t = 0.0; t -= S5_GT(G->g001,0,1); t -= S5_GT(G->g002,0,2); t -= S5_GT(G->g013,1,3); t -= S5_GT(G->g023,2,3); res[0] = t;
    t = 0.0; t -= S5_GT(G->g100,0,0); t -= S5_GT(G->g103,0,3); t -= S5_GT(G->g111,1,1); t -= S5_GT(G->g112,1,2); t -= S5_GT(G->g122,2,2); t -= S5_GT(G->g133,3,3); res[1] = t;
    t = 0.0; t -= S5_GT(G->g200,0,0); t -= S5_GT(G->g203,0,3); t -= S5_GT(G->g211,1,1); t -= S5_GT(G->g212,1,2); t -= S5_GT(G->g222,2,2); t -= S5_GT(G->g233,3,3); res[2] = t;
    t = 0.0; t -= S5_GT(G->g301,0,1); t -= S5_GT(G->g302,0,2); t -= S5_GT(G->g313,1,3); t -= S5_GT(G->g323,2,3); res[3] = t;
}
#undef S5_GT
/* -G^j_{ab} k^a k^b for one j.  sim5raytrace.c:151-156 (k_deriv) */
#define S5_KT(Gc, a, b) ((Gc) * k[a] * k[b])
S5_HD S5_INL double k_deriv0(const Conn* G, const double k[4]) { double t = 0.0; t -= S5_KT(G->g001,0,1); t -= S5_KT(G->g002,0,2); t -= S5_KT(G->g013,1,3); t -= S5_KT(G->g023,2,3); return t; }
S5_HD S5_INL double k_deriv1(const Conn* G, const double k[4]) { double t = 0.0; t -= S5_KT(G->g100,0,0); t -= S5_KT(G->g103,0,3); t -= S5_KT(G->g111,1,1); t -= S5_KT(G->g112,1,2); t -= S5_KT(G->g122,2,2); t -= S5_KT(G->g133,3,3); return t; }
S5_HD S5_INL double k_deriv2(const Conn* G, const double k[4]) { double t = 0.0; t -= S5_KT(G->g200,0,0); t -= S5_KT(G->g203,0,3); t -= S5_KT(G->g211,1,1); t -= S5_KT(G->g212,1,2); t -= S5_KT(G->g222,2,2); t -= S5_KT(G->g233,3,3); return t; }
S5_HD S5_INL double k_deriv3(const Conn* G, const double k[4]) { double t = 0.0; t -= S5_KT(G->g301,0,1); t -= S5_KT(G->g302,0,2); t -= S5_KT(G->g313,1,3); t -= S5_KT(G->g323,2,3); return t; }
#undef S5_KT

/* g_{mu nu} A^mu B^nu.  sim5kerr.c:608-625 */
S5_HD S5_INL double dotprod(const double A[4], const double B[4], const Metric* g)
{
    return A[0] * B[0] * g->g00 + A[1] * B[1] * g->g11 + A[2] * B[2] * g->g22 +
           A[3] * B[3] * g->g33 + A[0] * B[3] * g->g03 + A[3] * B[0] * g->g03;
}
/* sim5kerr.c:552-572 */
S5_HD S5_INL void vector_norm_to(double V[4], double norm, const Metric* g)
{
    double N = dotprod(V, V, g);
    V[0] *= sqrt(norm / N);
    V[1] *= sqrt(norm / N);
    V[2] *= sqrt(norm / N);
    V[3] *= sqrt(norm / N);
}

/* sim5kerr.c:677-710 */
S5_HD S5_INL void tetrad_zamo(const Metric* g, Tetrad* t)
{
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t->e[i][j] = 0.0;
    t->e[0][0] = sqrt(g->g33 / (sq(g->g03) - g->g33 * g->g00));
    t->e[0][3] = -t->e[0][0] * g->g03 / g->g33;
    t->e[1][1] = 1. / sqrt(g->g11);
    t->e[2][2] = -1. / sqrt(g->g22);
    t->e[3][3] = 1. / sqrt(g->g33);
    t->metric = *g;
}
/* sim5kerr.c:765-813 */
S5_HD S5_INL void tetrad_azimuthal(const Metric* g, double Omega, Tetrad* t)
{
    if (Omega == 0.0) { tetrad_zamo(g, t); return; }
    double g00 = g->g00, g33 = g->g33, g03 = g->g03;
    double U0 = sqrt(-1.0 / (g00 + 2. * Omega * g03 + sq(Omega) * g33));
    double U3 = U0 * Omega;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) t->e[i][j] = 0.0;
    t->e[0][0] = U0;
    t->e[0][3] = U3;
    t->e[1][1] = sqrt(1. / g->g11);
    t->e[2][2] = -sqrt(1. / g->g22);
    double k1 = (g03 * U3 + g00 * U0);
    double k2 = (g33 * U3 + g03 * U0);
    t->e[3][0] = -((k1) >= 0.0 ? (+1.0) : (-1.0)) * k2 / sqrt((g33 * g00 - g03 * g03) * (g00 * U0 * U0 + g33 * U3 * U3 + 2.0 * g03 * U0 * U3));
    t->e[3][3] = t->e[3][0] * (-k1 / k2);
    t->metric = *g;
}
/* sim5kerr.c:817-921 */
S5_HD S5_INL void tetrad_surface(const Metric* g, double Omega, double V, double dhdr, Tetrad* t)
{
    double g00 = g->g00, g11 = g->g11, g22 = g->g22, g33 = g->g33, g03 = g->g03;
    double S0r = 1.0 / sqrt(g11 + g22 * sq(dhdr));
    double S0h = S0r * dhdr;
    double ur = V / sqrt(1. - V * V) / sqrt(g11);
    double v = ((V) >= 0.0 ? (+1.0) : (-1.0)) * sqrt((sq(ur / S0r) * (-g00 - 2. * Omega * g03 - sq(Omega) * g33)) / (1. + sq(ur / S0r)));
    t->e[0][0] = 1.0; t->e[0][1] = v * S0r; t->e[0][2] = v * S0h; t->e[0][3] = Omega;
    vector_norm_to(t->e[0], -1.0, g);
    t->e[1][0] = (v * t->e[0][0]);
    t->e[1][1] = (v * t->e[0][1] + S0r / t->e[0][0]);
    t->e[1][2] = (v * t->e[0][2] + S0h / t->e[0][0]);
    t->e[1][3] = (v * t->e[0][3]);
    vector_norm_to(t->e[1], 1.0, g);
    t->e[2][0] = 0.0; t->e[2][1] = dhdr; t->e[2][2] = -1.0; t->e[2][3] = 0.0;
    vector_norm_to(t->e[2], 1.0, g);
    t->e[3][0] = -(g03 + g33 * Omega) / (g00 + g03 * Omega);
    t->e[3][1] = 0.0; t->e[3][2] = 0.0; t->e[3][3] = 1.0;
    vector_norm_to(t->e[3], 1.0, g);
    t->metric = *g;
}
/* sim5kerr.c:925-943 */
S5_HD S5_INL void bl2on(const double Vin[4], double Vout[4], const Tetrad* t)
{
    Vout[0] = -dotprod(t->e[0], Vin, &t->metric);
    Vout[1] = +dotprod(t->e[1], Vin, &t->metric);
    Vout[2] = +dotprod(t->e[2], Vin, &t->metric);
    Vout[3] = +dotprod(t->e[3], Vin, &t->metric);
}
/* sim5kerr.c:947-970 */
S5_HD S5_INL void on2bl(const double Vin[4], double Vout[4], const Tetrad* t)
{
    for (int i = 0; i < 4; i++) {
        double acc = 0.0;
        for (int j = 0; j < 4; j++) acc += Vin[j] * t->e[j][i];
        Vout[i] = acc;
    }
}

/* sim5kerr.c:980-989 */
S5_HD S5_INL double r_bh(double a) { return 1. + sqrt(1. - sq(a)); }
/* sim5kerr.c:1036-1046 */
S5_HD S5_INL double OmegaK(double r, double a) { return 1. / (a + crm::cr_pow_1p5(r)); }
/* Keplerian specific angular momentum (Komissarov 2008 form).  sim5kerr.c:1050-1071 */
S5_HD S5_INL double ellK(double r, double a) { return (sq(r) - 2. * a * sqrt(r) + sq(a)) / (sqrt(r) * r - 2. * sqrt(r) + a); }
/* sim5kerr.c:1101-1111 */
S5_HD S5_INL double Omega_from_ell(double ell, const Metric* g) { return -(g->g03 + ell * g->g00) / (g->g33 + ell * g->g03); }
/* sim5kerr.c:1114-1124 */
S5_HD S5_INL double ell_from_Omega(double Omega, const Metric* g) { return -(g->g03 + g->g33 * Omega) / (g->g00 + g->g03 * Omega); }
/* sim5kerr.c:1127-1141 */
template <class OPS>
S5_HD S5_INL double gfactorK_t(OPS& o, double r, double a, double l)
{
    double Om = o.div(1., a + crm::cr_pow_1p5(r));
    return o.div(o.sqrt(1. - o.div(2., r) * sq(1. - a * Om) - (r * r + a * a) * sq(Om)), 1. - Om * l);
}
S5_HD S5_MID double gfactorK(double r, double a, double l)
{
    ff::Quick f;
    double g = gfactorK_t(f, r, a, l);
    if (!f.ok) { ff::Plain p; g = gfactorK_t(p, r, a, l); }
    return g;
}
/* sim5kerr.c:1295-1309 */
S5_HD S5_INL void fourvelocity_azimuthal(double Omega, const Metric* g, double U[4])
{
    U[0] = sqrt(-1.0 / (g->g00 + 2. * Omega * g->g03 + sq(Omega) * g->g33));
    U[1] = 0.0;
    U[2] = 0.0;
    U[3] = U[0] * Omega;
}

/* k^mu from the constants of motion.  sim5kerr.c:1150-1213 (the CPU branch: NaN vector when M < 0) */
S5_HD S5_INL void photon_momentum(double a, double r, double m, double l, double q, double r_sign, double m_sign, double k[4])
{
    double a2 = sq(a), l2 = sq(l), r2 = sq(r), m2 = sq(m);
    double S = r2 + a2 * m2;
    double D = r2 - 2. * r + a2;
    double R = sq(r2 + a2 - a * l) - D * (sq(l - a) + q);
    double M = q - l2 * m2 / (1. - m2) + a2 * m2;
    if ((M < 0.0) && (-M < 1e-8)) M = 0.0;
    if ((R < 0.0) && (-R < 1e-8)) R = 0.0;
    if (M < 0.0) {
        k[0] = k[1] = k[2] = k[3] = NAN;
        return;
    }
    k[0] = +1 / S * (-a * (a * (1. - m2) - l) + (r2 + a2) / D * (r2 + a2 - a * l));
    k[1] = +1 / S * sqrt(R);
    k[2] = +1 / S * sqrt(M);
    k[3] = +1 / S * (-a + l / (1. - m2) + a / D * (r2 + a2 - a * l));
    if (r_sign < 0.0) k[1] = -k[1];
    if (m_sign < 0.0) k[2] = -k[2];
}
/* sim5kerr.c:1216-1250 */
S5_HD S5_INL void photon_motion_constants(double a, double r, double m, const double k[4], double* L, double* Q)
{
    double a2 = sq(a), r2 = sq(r);
    double s2 = 1. - m * m;
    double D = r2 - 2. * r + a2;
    double l;
    double nf = k[3] / k[0];
    double nh = sq(k[2]) / sq(k[0]);
    *L = l = (-a * a2 + sq(a2) * nf + nf * sq(r2) + a * (D - r2) + a2 * nf * (2. * r2 - D * s2)) * s2 /
             (D - a * s2 * (a - a2 * nf + nf * (D - r2)));
    *Q = sq(a * (l - a * s2) + ((a2 + r2) * (a2 - a * l + r2)) / D) *
         (nh - (sq(D * m) * (sq(l) - a2 * s2)) / (-s2 * sq(sq(a2) - a * a2 * l + sq(r2) + a * l * (D - r2) + a2 * (2. * r2 - D * s2))));
}
/* sim5kerr.c:1254-1268 */
S5_HD S5_INL double photon_carter_const(const double k[4], const Metric* g)
{
    double m2 = sq(g->m);
    double kt = k[0] * g->g00 + k[3] * g->g03;
    double kh = k[2] * g->g22;
    double kf = k[3] * g->g33 + k[0] * g->g03;
    return sq(kh) + sq(kf) * m2 / (1. - m2) - sq(g->a) * sq(kt) * m2;
}

} /* namespace s5 */
#endif
