// Development microbenchmark 2: MUFU throughputs (64H vs f32), conversions, and prototype IEEE division / sqrt built on
// the f32 MUFU seeds, checked bit-for-bit against the compiler's division / sqrt on random + adversarial operands.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096

__device__ __forceinline__ double rcp64h(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rsq64h(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ float rcp32(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsq32(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- prototype: IEEE division with an f32 MUFU.RCP seed (fast path only for "safe" exponents) ----
__device__ __forceinline__ double div_f32seed(double a, double b)
{
    unsigned bh = (unsigned)__double2hiint(b), bl = (unsigned)__double2loint(b);
    unsigned ah = (unsigned)__double2hiint(a);
    unsigned eb = (bh >> 20) & 0x7ff, ea = (ah >> 20) & 0x7ff;
    // safe: both normal, quotient exponent comfortably inside the normal range, intermediate 1/b normal
    bool safe = (eb - 2u < 2043u) && (ea - 64u < 1919u) && ((int)(ea - eb) > -960) && ((int)(ea - eb) < 960);
    if (!safe) return a / b;
    unsigned fb = 0x3f800000u | ((bh & 0xfffffu) << 3) | (bl >> 29);
    unsigned fr = __float_as_uint(rcp32(__uint_as_float(fb)));            // (0.5, 1]
    unsigned yh = (fr >> 3) + 0x77f00000u - (bh & 0x7ff00000u);           // exponent: E_f + 1919 - E_b
    yh |= bh & 0x80000000u;
    double y = __hiloint2double((int)yh, (int)(fr << 29));
    double e = __fma_rn(y, -b, 1.0);
    e = __fma_rn(e, e, e);
    y = __fma_rn(y, e, y);
    e = __fma_rn(y, -b, 1.0);
    y = __fma_rn(y, e, y);
    double q = __dmul_rn(a, y);
    double r = __fma_rn(q, -b, a);
    return __fma_rn(y, r, q);
}
// ---- prototype: IEEE sqrt with an f32 MUFU.RSQ seed ----
__device__ __forceinline__ double sqrt_f32seed(double x)
{
    unsigned xh = (unsigned)__double2hiint(x), xl = (unsigned)__double2loint(x);
    if (!(xh - 0x03500000u < 0x7ca00000u)) return sqrt(x);
    unsigned odd = ((xh >> 20) & 1u) ^ 1u;                                 // unbiased exponent parity
    unsigned fb = (0x3f800000u + (odd << 23)) | ((xh & 0xfffffu) << 3) | (xl >> 29);   // [1,4)
    unsigned fr = __float_as_uint(rsq32(__uint_as_float(fb)));            // (0.5, 1]
    // x = f * 2^(k - odd), k = E-1023 ; rsqrt(x) = rsq(f) * 2^(-(k-odd)/2)
    unsigned E = (xh >> 20) & 0x7ffu;
    unsigned half = (E - 1023u - odd) >> 1;                                // arithmetic on the even number (two's complement ok via int)
    int hk = ((int)(E - 1023u - odd)) >> 1;
    (void)half;
    unsigned yh = (fr >> 3) + ((unsigned)(1023 - 127 - hk) << 20);
    double y = __hiloint2double((int)yh, (int)(fr << 29));
    double t = __dmul_rn(y, y);
    double e = __fma_rn(x, -t, 1.0);
    double p = __fma_rn(e, 0.375, 0.5);
    double ye = __dmul_rn(y, e);
    double y1 = __fma_rn(p, ye, y);
    double s = __dmul_rn(x, y1);
    double h = __hiloint2double(__double2hiint(y1) - 0x00100000, __double2loint(y1));
    double r = __fma_rn(s, -s, x);
    return __fma_rn(r, h, s);
}

template <int MODE, int NCH>
__global__ void __launch_bounds__(128) k(double* out, double s0, double s1)
{
    double a[NCH], b[NCH];
    float f[NCH];
    for (int i = 0; i < NCH; i++) { a[i] = s0 + threadIdx.x * 1e-3 + i; b[i] = s1 + i * 1e-7 + threadIdx.x * 1e-9; f[i] = (float)a[i]; }
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) a[i] = rcp64h(a[i]);
            if (MODE == 1) a[i] = rsq64h(a[i]);
            if (MODE == 2) f[i] = rcp32(f[i]);
            if (MODE == 3) f[i] = rsq32(f[i]);
            if (MODE == 4) { f[i] = (float)a[i]; a[i] = (double)f[i] + b[i]; }     // F2F both ways + DADD
            if (MODE == 5) a[i] = b[i] / a[i] + b[i];
            if (MODE == 6) a[i] = div_f32seed(b[i], a[i]) + b[i];
            if (MODE == 7) a[i] = sqrt(a[i]) + b[i];
            if (MODE == 8) a[i] = sqrt_f32seed(a[i]) + b[i];
        }
    }
    double s = 0; for (int i = 0; i < NCH; i++) s += a[i] + f[i];
    if (s == 12345.678) out[0] = s;
}
template <int MODE, int NCH>
void run(const char* name, int blocks_per_sm)
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = p.multiProcessorCount * blocks_per_sm, threads = 128;
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<MODE, NCH><<<grid, threads>>>(out, 1.0000001, 0.99999);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double n = (double)grid * threads * ITERS * NCH;
    printf("%-40s ch=%d blk/SM=%d : %8.3f ms  %6.2f thread-ops/clk/SM (at %.3f GHz)\n", name, NCH, blocks_per_sm, best,
           n / (best * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate * 1e-6);
    cudaFree(out);
}

// ---- correctness ----
__device__ __forceinline__ uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
__global__ void check(unsigned long long* bad_div, unsigned long long* bad_sqrt, int rounds, int mode, double* ex)
{
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x12345677ULL + 99 + mode * 7777;
    unsigned long long bd = 0, bs = 0;
    for (int i = 0; i < rounds; i++) {
        uint64_t u = splitmix(s), v = splitmix(s);
        double a, b;
        if (mode == 0) {          // random mantissas, moderate exponents
            a = __longlong_as_double((long long)((u & 0x800fffffffffffffULL) | ((uint64_t)(1023 - 40 + (u >> 52) % 80) << 52)));
            b = __longlong_as_double((long long)((v & 0x800fffffffffffffULL) | ((uint64_t)(1023 - 40 + (v >> 52) % 80) << 52)));
        } else if (mode == 1) {   // fully random bit patterns (incl. inf/nan/subnormals)
            a = __longlong_as_double((long long)u); b = __longlong_as_double((long long)v);
        } else if (mode == 2) {   // adversarial mantissas: few set bits / all ones / near powers of two
            uint64_t ma = (u & 1) ? (0x000fffffffffffffULL >> (u >> 58)) : (1ULL << ((u >> 8) % 52)) | ((u >> 20) & 3);
            uint64_t mb = (v & 1) ? (0x000fffffffffffffULL << ((v >> 58) % 52)) & 0x000fffffffffffffULL : (1ULL << ((v >> 8) % 52)) | ((v >> 20) & 7);
            a = __longlong_as_double((long long)(ma | ((uint64_t)(1023 + (int)((u >> 40) % 9) - 4) << 52)));
            b = __longlong_as_double((long long)(mb | ((uint64_t)(1023 + (int)((v >> 40) % 9) - 4) << 52)));
        } else {                  // quotients near 1 and small integers ratios
            a = (double)(1 + (u % 4096)) * (1.0 + (double)((u >> 20) & 0xff) * 2.220446049250313e-16);
            b = (double)(1 + (v % 4096)) * (1.0 + (double)((v >> 20) & 0xff) * 2.220446049250313e-16);
        }
        double q0 = a / b, q1 = div_f32seed(a, b);
        if (__double_as_longlong(q0) != __double_as_longlong(q1) && !(q0 != q0 && q1 != q1)) { if (!bd) { ex[0] = a; ex[1] = b; } bd++; }
        double xa = fabs(a);
        double r0 = sqrt(xa), r1 = sqrt_f32seed(xa);
        if (__double_as_longlong(r0) != __double_as_longlong(r1) && !(r0 != r0 && r1 != r1)) { if (!bs) ex[2] = xa; bs++; }
    }
    if (bd) atomicAdd(bad_div, bd);
    if (bs) atomicAdd(bad_sqrt, bs);
}
// seed accuracy: max relative error of the seeds over all 2^20 high-word mantissas (RCP64H sees only the high word)
__global__ void seed_err(double* out)
{
    unsigned m = blockIdx.x * blockDim.x + threadIdx.x;      // 2^20 values
    double worst[4] = {0, 0, 0, 0};
    for (unsigned lo = 0; lo < 4; lo++) {
        unsigned lw = lo == 0 ? 0u : lo == 1 ? 0xffffffffu : lo == 2 ? 0x80000000u : 0x12345678u;
        double x = __hiloint2double((int)(0x3ff00000u | m), (int)lw);
        double y = rcp64h(x);
        y = __hiloint2double(__double2hiint(y), 0);
        double e0 = fabs(__fma_rn(y, -x, 1.0));
        unsigned fb = 0x3f800000u | (m << 3) | (lw >> 29);
        unsigned fr = __float_as_uint(rcp32(__uint_as_float(fb)));
        double y2 = __hiloint2double((int)((fr >> 3) + 0x77f00000u - 0x3ff00000u), (int)(fr << 29));
        double e1 = fabs(__fma_rn(y2, -x, 1.0));
        worst[0] = fmax(worst[0], e0); worst[1] = fmax(worst[1], e1);
        for (int odd = 0; odd < 2; odd++) {
            double xx = __hiloint2double((int)((odd ? 0x40000000u : 0x3ff00000u) | m), (int)lw);
            double r = rsq64h(xx); r = __hiloint2double(__double2hiint(r), 0);
            double e2 = fabs(__fma_rn(__dmul_rn(r, r), -xx, 1.0));
            unsigned fb2 = (0x3f800000u + ((unsigned)odd << 23)) | (m << 3) | (lw >> 29);
            unsigned fr2 = __float_as_uint(rsq32(__uint_as_float(fb2)));
            double r2 = __hiloint2double((int)((fr2 >> 3) + ((unsigned)(1023 - 127) << 20)), (int)(fr2 << 29));
            double e3 = fabs(__fma_rn(__dmul_rn(r2, r2), -xx, 1.0));
            worst[2] = fmax(worst[2], e2); worst[3] = fmax(worst[3], e3);
        }
    }
    for (int i = 0; i < 4; i++) {
        double w = worst[i];
        for (int off = 16; off; off >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, off));
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&out[i], (unsigned long long)__double_as_longlong(w));
    }
}

int main()
{
    run<0, 8>("MUFU.RCP64H", 8); run<1, 8>("MUFU.RSQ64H", 8); run<2, 8>("MUFU.RCP f32", 8); run<3, 8>("MUFU.RSQ f32", 8);
    run<4, 8>("F2F f64->f32->f64 + DADD", 8);
    run<5, 4>("a/b + DADD (compiler)", 8); run<6, 4>("div_f32seed + DADD", 8); run<6, 4>("div_f32seed + DADD", 4); run<6, 2>("div_f32seed + DADD", 2);
    run<7, 4>("sqrt + DADD (compiler)", 8); run<8, 4>("sqrt_f32seed + DADD", 8); run<8, 4>("sqrt_f32seed + DADD", 4); run<8, 2>("sqrt_f32seed + DADD", 2);
    unsigned long long *bad; double* ex; cudaMallocManaged(&bad, 16); cudaMallocManaged(&ex, 64);
    for (int mode = 0; mode < 4; mode++) {
        bad[0] = bad[1] = 0; ex[0] = ex[1] = ex[2] = 0;
        check<<<148 * 16, 256>>>(bad, bad + 1, 4096, mode, ex);
        cudaDeviceSynchronize();
        printf("check mode %d: %llu tests, div mismatches %llu (e.g. %a / %a), sqrt mismatches %llu (e.g. %a)\n", mode, 148ULL * 16 * 256 * 4096, bad[0], ex[0], ex[1], bad[1], ex[2]);
    }
    double* w; cudaMallocManaged(&w, 32); w[0] = w[1] = w[2] = w[3] = 0;
    seed_err<<<(1 << 20) / 256, 256>>>(w); cudaDeviceSynchronize();
    printf("seed |1 - y*x| max: RCP64H %.3e (2^%.2f)  f32 seed %.3e (2^%.2f) ; |1 - r*r*x| max: RSQ64H %.3e (2^%.2f)  f32 seed %.3e (2^%.2f)\n",
           w[0], log2(w[0]), w[1], log2(w[1]), w[2], log2(w[2]), w[3], log2(w[3]));
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
