// Micro-benchmark: what limits a pure DFMA stream to ~92 % of 64 FMA/clk/SM on B200?
// Varies operand pattern, chains per thread and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT, int CH>
__global__ void __launch_bounds__(256) k(double *sink, int iters, double a, double b)
{
    double v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 256 / CH; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (PAT == 0) v[i] = fma(v[i], a, b);          // 3 distinct regs, 2 shared
                if (PAT == 1) v[i] = fma(v[i], v[i], v[i]);    // 1 distinct reg
                if (PAT == 2) v[i] = fma(v[i], a, v[i]);       // 2 distinct
                if (PAT == 3) v[i] = fma(v[i], 0.9990234375, b);   // immediate-encodable multiplier
                if (PAT == 4) v[i] = v[i] + a;                 // DADD
                if (PAT == 5) v[i] = v[i] * a;                 // DMUL
                if (PAT == 6) v[i] = fma(v[i], v[(i + 1) % CH], v[(i + 2) % CH]);  // 3 distinct varying
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += v[i];
    if (s == 123.456) sink[0] = s;
}

template <int PAT, int CH>
void run(const char *name, double *sink, int sms, int blocks_per_sm, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = sms * blocks_per_sm;
    k<PAT, CH><<<blocks, threads>>>(sink, 50, 0.999, 1e-9);
    cudaEventRecord(e0);
    k<PAT, CH><<<blocks, threads>>>(sink, iters, 0.999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * (double)iters * 256;
    printf("%-34s ch=%2d warps/SM=%2d  %8.2f ms  %.2f T op/s  (%.1f /clk/SM @1.965)\n", name, CH,
           blocks_per_sm * threads / 32, ms, ops / ms / 1e9, ops / ms / 1e6 / 148 / 1.965e3);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 8);
    const int s = p.multiProcessorCount;
    run<0, 8>("fma(v,a,b)", sink, s, 8, 256);
    run<0, 8>("fma(v,a,b)", sink, s, 4, 256);
    run<0, 8>("fma(v,a,b)", sink, s, 2, 256);
    run<0, 8>("fma(v,a,b)", sink, s, 1, 256);
    run<0, 8>("fma(v,a,b)", sink, s, 1, 128);
    run<0, 16>("fma(v,a,b)", sink, s, 4, 256);
    run<0, 4>("fma(v,a,b)", sink, s, 8, 256);
    run<1, 8>("fma(v,v,v)", sink, s, 8, 256);
    run<2, 8>("fma(v,a,v)", sink, s, 8, 256);
    run<3, 8>("fma(v,imm,b)", sink, s, 8, 256);
    run<4, 8>("dadd(v,a)", sink, s, 8, 256);
    run<5, 8>("dmul(v,a)", sink, s, 8, 256);
    run<6, 8>("fma(v_i,v_i+1,v_i+2)", sink, s, 8, 256);
    run<6, 16>("fma(v_i,v_i+1,v_i+2)", sink, s, 4, 256);
    return 0;
}
