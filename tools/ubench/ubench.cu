// Development microbenchmarks: what the sm_100a FP64 pipe really sustains for the instruction shapes of the
// Carlson loops (register-operand DFMA/DADD/DMUL, IEEE division, IEEE sqrt, dependent-chain latency).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

template <int MODE, int NCH>
__global__ void __launch_bounds__(256) k(double* out, double s0, double s1, double s2)
{
    double a[NCH], b[NCH], c[NCH];
    for (int i = 0; i < NCH; i++) { a[i] = s0 + threadIdx.x * 1e-3 + i; b[i] = s1 + i * 1e-7 + threadIdx.x * 1e-12; c[i] = s2 + i * 1e-9 + threadIdx.x * 1e-13; }
    #pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0) a[i] = __fma_rn(a[i], b[i], c[i]);              // DFMA, 3 register operands
            if (MODE == 1) a[i] = __dadd_rn(a[i], c[i]);                   // DADD
            if (MODE == 2) a[i] = __dmul_rn(a[i], b[i]);                   // DMUL
            if (MODE == 3) a[i] = __fma_rn(a[i], 0.9999999, 1e-9);         // DFMA, constant operands
            if (MODE == 4) a[i] = c[i] / a[i] + b[i];                      // IEEE division (+1 DADD)
            if (MODE == 5) a[i] = sqrt(a[i]) + b[i];                       // IEEE sqrt (+1 DADD)
            if (MODE == 6) { a[i] = __dadd_rn(__dmul_rn(a[i], b[i]), c[i]); }   // DMUL+DADD dependent pair
        }
    }
    double s = 0; for (int i = 0; i < NCH; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE, int NCH>
void run(const char* name, int ops_per_it, int blocks_per_sm, int threads)
{
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = p.multiProcessorCount * blocks_per_sm;
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<MODE, NCH><<<grid, threads>>>(out, 1.0000001, 0.99999, 1e-7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    double n = (double)grid * threads * ITERS * NCH;
    double per_sm_clk = n * ops_per_it / (best * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3);
    printf("%-44s ch=%d blk/SM=%d thr=%3d : %8.3f ms  %7.2f Gop/s  %6.2f thread-ops/clk/SM\n", name, NCH, blocks_per_sm, threads, best,
           n * ops_per_it / (best * 1e-3) / 1e9, per_sm_clk);
    cudaFree(out);
}

int main()
{
    run<3, 8>("DFMA const operands (peak kernel shape)", 1, 8, 256);
    run<0, 8>("DFMA 3 register operands", 1, 8, 256);
    run<0, 4>("DFMA 3 register operands", 1, 8, 256);
    run<1, 8>("DADD register operands", 1, 8, 256);
    run<2, 8>("DMUL register operands", 1, 8, 256);
    run<6, 8>("DMUL->DADD pair", 2, 8, 256);
    run<0, 1>("DFMA dependent chain, 1 warp/SMSP", 1, 1, 128);
    run<0, 1>("DFMA dependent chain, 2 warps/SMSP", 1, 2, 128);
    run<0, 1>("DFMA dependent chain, 4 warps/SMSP", 1, 4, 128);
    run<0, 1>("DFMA dependent chain, 8 warps/SMSP", 1, 8, 128);
    run<0, 1>("DFMA dependent chain, 16 warps/SMSP", 1, 8, 256);
    run<0, 2>("DFMA 2 chains, 1 warp/SMSP", 1, 1, 128);
    run<0, 4>("DFMA 4 chains, 1 warp/SMSP", 1, 1, 128);
    run<0, 8>("DFMA 8 chains, 1 warp/SMSP", 1, 1, 128);
    run<4, 1>("div+add, 1 chain, 1 warp/SMSP", 1, 1, 128);
    run<4, 1>("div+add, 1 chain, 8 warps/SMSP", 1, 8, 128);
    run<4, 4>("div+add, 4 chains, 1 warp/SMSP", 1, 1, 128);
    run<4, 4>("div+add, 4 chains, 4 warps/SMSP", 1, 4, 128);
    run<4, 4>("div+add, 4 chains, 8 warps/SMSP", 1, 8, 128);
    run<4, 8>("div+add, 8 chains, 8 warps/SMSP", 1, 8, 128);
    run<5, 1>("sqrt+add, 1 chain, 1 warp/SMSP", 1, 1, 128);
    run<5, 4>("sqrt+add, 4 chains, 4 warps/SMSP", 1, 4, 128);
    run<5, 4>("sqrt+add, 4 chains, 8 warps/SMSP", 1, 8, 128);
    run<5, 8>("sqrt+add, 8 chains, 8 warps/SMSP", 1, 8, 128);
    return 0;
}
