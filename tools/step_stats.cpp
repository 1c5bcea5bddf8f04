// tools/step_stats.cpp -- how often each path of the stepwise integrator runs on BASELINE config 4 (host instantiation of the device
// headers; development aid, not part of the product):  g++ -O2 -fopenmp -DS5_STEP_STATS -Iinclude tools/step_stats.cpp -o /tmp/step_stats && /tmp/step_stats [n]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../sim5_b200/csrc/pixel.cuh"
namespace crm { extern "C" { long long s5_stat[16]; } }
using namespace s5;
int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 128;
    sim5_image_params p;
    memset(&p, 0, sizeof p);
    p.struct_size = sizeof p; p.mode = SIM5_MODE_STEPWISE; p.nx = p.ny = n; p.bh_spin = 0.9; p.incl = 60.0 / 180.0 * M_PI; p.rmax = 25.0;
    p.max_order = 1; p.disk_mass = 10; p.disk_mdot = 0.1; p.disk_alpha = 0.1; p.precision_factor = 0.01; p.r_start = 50; p.step_max = 1e9; p.max_steps = 100000;
    p.torus_rc = 10; p.torus_w = 2; p.torus_h = 0.3; p.torus_j0 = 1; p.torus_k0 = 0.05;
    { double r = p.torus_rc, a = p.bh_spin; p.torus_ell = (r * r - 2. * a * sqrt(r) + a * a) / (sqrt(r) * r - 2. * sqrt(r) + a); }
    p.outputs = SIM5_OUT_INTENSITY | SIM5_OUT_TAU | SIM5_OUT_STEPS | SIM5_OUT_STATUS;
    S5ImageConsts c;
    s5_fill_image_consts(&p, &c);
    long long rays = 0, hist[8] = {0};
    #pragma omp parallel for schedule(dynamic, 1) reduction(+:rays)
    for (int iy = 0; iy < n; iy++) for (int ix = 0; ix < n; ix++) {
        StepRay s; PixelOut o;
        if (!stepwise_start(c, ix, iy, &s, &o)) continue;
        rays++;
        int cls;
        while (!(cls = stepwise_step(c, &s))) {}
    }
    printf("rays %lld  steps %lld (%.0f per ray)\n", rays, crm::s5_stat[0], (double)crm::s5_stat[0] / rays);
    printf("k_iter = 1: %.4f  2: %.4f  3: %.4f\n", (double)crm::s5_stat[1] / crm::s5_stat[0], (double)crm::s5_stat[2] / crm::s5_stat[0], (double)crm::s5_stat[3] / crm::s5_stat[0]);
    printf("RK4 fallback: %.4f of the steps\n", (double)crm::s5_stat[6] / crm::s5_stat[0]);
    printf("torus branch: %.4f of the steps\n", (double)crm::s5_stat[7] / crm::s5_stat[0]);
    printf("acos: full routine %.4f of the steps (of which in doubt after the shortcut %.4f)\n", (double)crm::s5_stat[4] / crm::s5_stat[0], (double)crm::s5_stat[5] / crm::s5_stat[0]);
    (void)hist;
    return 0;
}
