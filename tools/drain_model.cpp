// tools/drain_model.cpp -- why the step-wise image (BASELINE configs[3]) does not strong-scale: a replay of k_trace_lanes' hand-out
// policy on the REAL per-ray step counts of the image (host instantiation of the device headers), development aid, not product.
//
//   g++ -O2 -fopenmp -Iinclude tools/drain_model.cpp -o /tmp/drain_model
//   /tmp/drain_model [n=1024] [counts-cache=/tmp/s5_step_counts_<n>.bin]
//
// Step 1: the number of PROG::step calls of every ray (6 minutes on 8 cores at 1024^2; cached in the counts file).
// Step 2: for N = 1, 2, 4, 8 GPUs the kernel's scheduling is replayed: N * 148 SMs * 4 CTAs * 4 warps, rows handed out from the middle
//   of the image outwards, one queue for all warps (SIM5_FLAG_SHARED_QUEUE; the static split measures the same),
//   a warp refills at the start of a 16-step round when >= 8 of its lanes are idle (or all), and it advances while ANY lane is live.
//   Time model per SM: one round of steps of its resident warps takes max(t_lat, active_warps * t_issue) -- a dependent chain per ray
//   (latency) against the shared FP64 pipe (issue).  The two constants are fitted to the measured N = 1 and N = 4 times; N = 2 and
//   N = 8 are then predictions to hold against the measurements (profiles/r07b_queue_cfg4_n4.json, r05o_bench_cfg4_n8*.json).
// Step 3: the same replay with other policies, to see what a change could buy before anybody writes it.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <queue>
#include <algorithm>
#include <string>
#include "../sim5_b200/csrc/pixel.cuh"
namespace crm { extern "C" { long long s5_stat[16]; } }
using namespace s5;

static std::vector<int> ray_counts(int n, const std::string& cache)
{
    std::vector<int> cnt((size_t)n * n, 0);
    if (FILE* f = fopen(cache.c_str(), "rb")) {
        size_t got = fread(cnt.data(), sizeof(int), cnt.size(), f);
        fclose(f);
        if (got == cnt.size()) { fprintf(stderr, "per-ray step counts from %s\n", cache.c_str()); return cnt; }
    }
    sim5_image_params p;                 /* BASELINE configs[3] = sim5_default_params(4, .) */
    memset(&p, 0, sizeof p);
    p.struct_size = sizeof p; p.mode = SIM5_MODE_STEPWISE; p.nx = p.ny = n; p.bh_spin = 0.9; p.incl = 60.0 / 180.0 * M_PI; p.rmax = 25.0;
    p.max_order = 1; p.disk_mass = 10; p.disk_mdot = 0.1; p.disk_alpha = 0.1; p.precision_factor = 0.01; p.r_start = 50; p.step_max = 1e9; p.max_steps = 100000;
    p.torus_rc = 10; p.torus_w = 2; p.torus_h = 0.3; p.torus_j0 = 1; p.torus_k0 = 0.05;
    { double r = p.torus_rc, a = p.bh_spin; p.torus_ell = (r * r - 2. * a * sqrt(r) + a * a) / (sqrt(r) * r - 2. * sqrt(r) + a); }
    p.outputs = SIM5_OUT_INTENSITY | SIM5_OUT_TAU | SIM5_OUT_STEPS | SIM5_OUT_STATUS;
    S5ImageConsts c;
    s5_fill_image_consts(&p, &c);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int iy = 0; iy < n; iy++) for (int ix = 0; ix < n; ix++) {
        StepRay s; PixelOut o;
        if (!stepwise_start(c, ix, iy, &s, &o)) continue;
        int k = 1;
        while (!stepwise_step(c, &s)) k++;
        cnt[(size_t)iy * n + ix] = k;
    }
    if (FILE* f = fopen(cache.c_str(), "wb")) { fwrite(cnt.data(), sizeof(int), cnt.size(), f); fclose(f); }
    return cnt;
}

struct Policy {
    int refill_min = 8;          // S5_STEP_REFILL_MIN
    int round = 16;              // S5_STEPS_PER_ROUND
    int order = 0;               // 0: rows centre-out (the kernel), 1: row-major, 2: rays sorted by length, longest first (oracle order),
                                 // 3: pilot order -- tile x tile pixel blocks sorted by the step count of their centre ray (a 1/tile^2 pre-pass), longest first
    int tile = 8;
    bool repack = false;         // drain: the live rays of a CTA are packed into as few warps as possible at every round start
};

struct Result { double t_us; double lane_util; double warp_util; };

// One replay.  Every SM has a clock of its own; the SM with the earliest clock runs its next round.  Returns the kernel time for the
// given (t_lat, t_issue) in us per STEP.
static Result replay(const std::vector<int>& cnt, int n, int ngpu, const Policy& pol, double t_lat, double t_issue)
{
    const int WPC = 4, CPS = 4, SMS = 148 * ngpu, WPS = WPC * CPS;
    const size_t npix = cnt.size();
    std::vector<unsigned> order(npix);
    if (pol.order == 2) {
        for (size_t i = 0; i < npix; i++) order[i] = (unsigned)i;
        std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return cnt[a] > cnt[b]; });
    } else if (pol.order == 3) {
        const int T = pol.tile, nt = (n + T - 1) / T;
        std::vector<unsigned> tiles((size_t)nt * nt);
        for (size_t i = 0; i < tiles.size(); i++) tiles[i] = (unsigned)i;
        auto pilot = [&](unsigned t) { int ty = (int)(t / nt), tx = (int)(t % nt); int y = std::min(n - 1, ty * T + T / 2), x = std::min(n - 1, tx * T + T / 2); return cnt[(size_t)y * n + x]; };
        std::stable_sort(tiles.begin(), tiles.end(), [&](unsigned a, unsigned b) { return pilot(a) > pilot(b); });
        size_t k = 0;
        for (unsigned t : tiles) {
            int ty = (int)(t / nt), tx = (int)(t % nt);
            for (int y = ty * T; y < std::min(n, ty * T + T); y++) for (int x = tx * T; x < std::min(n, tx * T + T); x++) order[k++] = (unsigned)((size_t)y * n + x);
        }
    } else {
        for (size_t p = 0; p < npix; p++) {
            int lr = (int)(p / n), ix = (int)(p % n);
            if (pol.order == 0) { int mid = n >> 1; lr = (lr & 1) ? mid - ((lr + 1) >> 1) : mid + (lr >> 1); }
            order[p] = (unsigned)((size_t)lr * n + ix);
        }
    }
    size_t head = 0;
    std::vector<int> rem((size_t)SMS * WPS * 32, 0);           // remaining steps per lane; 0 = idle
    std::vector<char> drained((size_t)SMS * WPS, 0);
    typedef std::pair<double, int> Ev;
    std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev> > pq;
    for (int s = 0; s < SMS; s++) pq.push(Ev(0.0, s));
    double t_end = 0.0, lane_steps = 0.0, warp_steps = 0.0, warp_slots = 0.0;
    while (!pq.empty()) {
        Ev e = pq.top(); pq.pop();
        const int sm = e.second;
        int* R = &rem[(size_t)sm * WPS * 32];
        char* D = &drained[(size_t)sm * WPS];
        // round start: refills (and the optional re-packing inside each CTA once the queue is dry)
        if (pol.repack && head >= npix) {
            for (int c = 0; c < CPS; c++) {
                int* L = R + c * WPC * 32;
                int k = 0;
                for (int i = 0; i < WPC * 32; i++) if (L[i] > 0) { int v = L[i]; L[i] = 0; L[k++] = v; }
            }
        }
        int active = 0, max_round = 0;
        for (int w = 0; w < WPS; w++) {
            int* L = R + w * 32;
            int idle = 0;
            for (int i = 0; i < 32; i++) idle += (L[i] <= 0);
            if (!D[w] && (idle == 32 || idle >= pol.refill_min)) {
                for (int i = 0; i < 32 && idle > 0; i++) if (L[i] <= 0) {
                    while (head < npix && cnt[order[head]] == 0) head++;        // rays that end in start() cost nothing here
                    if (head >= npix) break;
                    L[i] = cnt[order[head++]];
                    idle--;
                }
                if (head >= npix) D[w] = 1;
            }
            int live = 0, longest = 0, shortest_all = 1 << 30;
            for (int i = 0; i < 32; i++) if (L[i] > 0) { live++; longest = std::max(longest, L[i]); shortest_all = std::min(shortest_all, L[i]); }
            if (!live) continue;
            // the warp runs `round` steps, or fewer if its last live lane finishes earlier
            int steps = std::min(pol.round, longest);
            for (int i = 0; i < 32; i++) if (L[i] > 0) { int d = std::min(L[i], steps); lane_steps += d; L[i] -= d; }
            warp_steps += steps;
            active++;
            max_round = std::max(max_round, steps);
        }
        if (!active) {
            bool all_drained = true;
            for (int w = 0; w < WPS; w++) all_drained = all_drained && D[w];
            if (!all_drained && head < npix) { pq.push(Ev(e.first + t_lat, sm)); }
            else t_end = std::max(t_end, e.first);
            continue;
        }
        const double per_step = std::max(t_lat, active * t_issue);
        const double dt = per_step * max_round;
        warp_slots += (double)WPS * max_round;
        pq.push(Ev(e.first + dt, sm));
    }
    Result r;
    r.t_us = t_end;
    r.lane_util = lane_steps / (warp_steps * 32.0);
    r.warp_util = warp_steps / warp_slots;
    return r;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const std::string cache = argc > 2 ? argv[2] : ("/tmp/s5_step_counts_" + std::to_string(n) + ".bin");
    std::vector<int> cnt = ray_counts(n, cache);
    long long total = 0; int longest = 0, shortest = 1 << 30; size_t live = 0;
    for (int v : cnt) if (v > 0) { total += v; longest = std::max(longest, v); shortest = std::min(shortest, v); live++; }
    printf("image %dx%d: %zu stepped rays, %lld steps, %d .. %d per ray (mean %.0f)\n", n, n, live, total, shortest, longest, (double)total / live);

    // fit (t_lat, t_issue) to the measured kernel times at N = 1 and N = 4 (ms; argv[3], argv[4] override)
    const double m1 = argc > 3 ? atof(argv[3]) : 300.6, m4 = argc > 4 ? atof(argv[4]) : 102.4;
    Policy kernel;
    // the replay is homogeneous in (t_lat, t_issue): the RATIO x = t_lat / (16 t_issue) shapes T(N=4) / T(N=1), the scale then sets T(N=1)
    double best = 1e30, bl = 0, bi = 0;
    for (double x = 0.05; x <= 2.0001; x += 0.05) {
        const double ti = 1.0 / 16.0, tl = x;
        double a = replay(cnt, n, 1, kernel, tl, ti).t_us, b = replay(cnt, n, 4, kernel, tl, ti).t_us;
        double err = fabs((b / a) / (m4 / m1) - 1.0);
        if (err < best) { best = err; const double k = m1 * 1e3 / a; bl = tl * k; bi = ti * k; }
    }
    printf("fit to N=1 %.1f ms and N=4 %.1f ms: t_lat %.2f us per step (one ray alone), t_issue %.4f us per warp-step (16 warps: %.2f us per step); residual %.3f\n",
           m1, m4, bl, bi, 16 * bi, best);
    struct Case { const char* name; Policy p; };
    std::vector<Case> cases;
    cases.push_back({"the kernel (centre-out rows, refill at 8 idle lanes, 16-step rounds)", kernel});
    { Policy p; p.order = 1; cases.push_back({"row-major hand-out", p}); }
    { Policy p; p.order = 2; cases.push_back({"rays sorted by length, longest first (needs an oracle)", p}); }
    { Policy p; p.refill_min = 1; cases.push_back({"refill at 1 idle lane", p}); }
    { Policy p; p.refill_min = 16; cases.push_back({"refill at 16 idle lanes", p}); }
    { Policy p; p.round = 4; cases.push_back({"4-step rounds", p}); }
    { Policy p; p.repack = true; cases.push_back({"+ re-pack the live rays of a CTA into full warps once the queue is dry", p}); }
    { Policy p; p.order = 3; p.tile = 8; cases.push_back({"pilot order: 8x8 tiles by the step count of their centre ray (1/64 pre-pass)", p}); }
    { Policy p; p.order = 3; p.tile = 4; cases.push_back({"pilot order: 4x4 tiles (1/16 pre-pass)", p}); }
    { Policy p; p.order = 3; p.tile = 16; cases.push_back({"pilot order: 16x16 tiles (1/256 pre-pass)", p}); }
    { Policy p; p.order = 3; p.tile = 8; p.refill_min = 1; cases.push_back({"pilot order 8x8 + refill at 1 idle lane", p}); }
    { Policy p; p.order = 3; p.tile = 8; p.refill_min = 16; cases.push_back({"pilot order 8x8 + refill at 16 idle lanes", p}); }
    printf("%-78s %9s %9s %9s %9s   lane/warp utilisation at N=1, N=8\n", "policy", "N=1 ms", "N=2 ms", "N=4 ms", "N=8 ms");
    for (const Case& c : cases) {
        Result r[4];
        int k = 0;
        for (int g : {1, 2, 4, 8}) r[k++] = replay(cnt, n, g, c.p, bl, bi);
        printf("%-78s %9.1f %9.1f %9.1f %9.1f   %.3f/%.3f  %.3f/%.3f\n", c.name, r[0].t_us * 1e-3, r[1].t_us * 1e-3, r[2].t_us * 1e-3, r[3].t_us * 1e-3,
               r[0].lane_util, r[0].warp_util, r[3].lane_util, r[3].warp_util);
    }
    printf("ideal (N=1 / N): %.1f %.1f %.1f %.1f ms; bound by the longest ray alone: %.1f ms\n", m1, m1 / 2, m1 / 4, m1 / 8, longest * bl * 1e-3);
    return 0;
}
