// Micro-benchmark: does FP64 DMMA (mma.sync m8n8k4 f64) run concurrently with vector DFMA on B200?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>   // NF DFMA chains x 8 per trip, NM DMMA accumulators per trip
__global__ void __launch_bounds__(256) probe(double *sink, int iters, double a, double b)
{
    double v[8], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (NF) {
#pragma unroll
                for (int i = 0; i < NF; ++i) v[i] = fma(v[i], a, b);
            }
            if (NM) {
#pragma unroll
                for (int i = 0; i < NM; ++i) dmma(c0[i], c1[i], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    if (s == 123.456) sink[0] = s;
}

template <int NF, int NM>
void run(const char *name, double *sink, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = sms * 8, threads = 256;
    probe<NF, NM><<<blocks, threads>>>(sink, 100, 0.999, 1e-9);
    cudaEventRecord(e0);
    probe<NF, NM><<<blocks, threads>>>(sink, iters, 0.999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double dfma = warps * 32 * (double)iters * 8 * NF;       // thread-level DFMA
    const double dm = warps * (double)iters * 8 * NM;               // warp-level DMMA (256 FMA each)
    printf("%-22s %8.2f ms  DFMA %.2f T/s  DMMA-FMA %.2f T/s  total FMA %.2f T/s\n", name, ms, dfma / ms / 1e9,
           dm * 256 / ms / 1e9, (dfma + dm * 256) / ms / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *sink; cudaMalloc(&sink, 8);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<8, 0>("dfma only (8 chains)", sink, p.multiProcessorCount);
    run<0, 4>("dmma only (4 acc)", sink, p.multiProcessorCount);
    run<0, 2>("dmma only (2 acc)", sink, p.multiProcessorCount);
    run<8, 1>("8 dfma + 1 dmma", sink, p.multiProcessorCount);
    run<8, 2>("8 dfma + 2 dmma", sink, p.multiProcessorCount);
    run<6, 1>("6 dfma + 1 dmma", sink, p.multiProcessorCount);
    run<4, 4>("4 dfma + 4 dmma", sink, p.multiProcessorCount);
    return 0;
}
