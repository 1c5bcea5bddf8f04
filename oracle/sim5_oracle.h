/*
 * sim5_oracle.h -- entry points of the CPU restatement (oracle/sim5_oracle.c).  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product.
 */
#ifndef SIM5_ORACLE_H
#define SIM5_ORACLE_H
#include "sim5_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* trace rows [row_begin,row_end) of an EQPLANE / POLARIZED image with the call-for-call algorithm of the reference.
 * Returns the elapsed seconds of the pixel loop; -1 bad arguments, -2 HISTOGRAM (use orc_trace_histogram),
 * -3 STEPWISE (not restated; checked against oracle/_ref and tests/golden/image_cfg4_16.npz). */
double orc_trace_image(const sim5_image_params* p, const sim5_image_out* out, int nthreads, sim5_trace_stats* stats);
double orc_trace_histogram(const sim5_image_params* p, double* hist, int nthreads);
double orc_trace_spectrum(const sim5_image_params* p, double* spec, int nthreads);
int    orc_max_threads(void);
double orc_r_ms(double a);
double orc_rf(double x, double y, double z);
double orc_rd(double x, double y, double z);
double orc_rc(double x, double y);
double orc_rj(double x, double y, double z, double p);
void   orc_sncndn(double u, double m, double* sn, double* cn, double* dn);
void orc_batch_rf(long n, const double* x, const double* y, const double* z, double* o);
void orc_batch_rd(long n, const double* x, const double* y, const double* z, double* o);
void orc_batch_rc(long n, const double* x, const double* y, double* o);
void orc_batch_rj(long n, const double* x, const double* y, const double* z, const double* p, double* o);
void orc_batch_sncndn(long n, const double* u, const double* m, double* sn, double* cn, double* dn);
#ifdef __cplusplus
}
#endif
#endif
