// Schedule tuning harness for gsf_sum_kernel: compile with -DGSF_TUNE_STYLE=s -DGSF_TUNE_UNROLL=u,
// run on the GPU, prints G point*modes/s for the main (D, NC, P) combinations.
#include <cstdio>
#include <vector>
#include <random>
#include "../../gstools-core_b200/csrc/gsf_kernels.cuh"

template <int D, int NC, int P>
double run(int64_t m, int n_modes, double tail_waves)
{
    using namespace gsf;
    constexpr int R = rec_doubles(D, NC);
    std::mt19937_64 rng(1);
    std::normal_distribution<double> nd;
    std::uniform_real_distribution<double> ud(0, 100);
    std::vector<double> rec((size_t)n_modes * R, 0.0), pos((size_t)D * m);
    for (int i = 0; i < n_modes; ++i) {
        for (int d = 0; d < D; ++d) rec[(size_t)i * R + d] = nd(rng) / 3.14159;
        rec[(size_t)i * R + D] = ud(rng) / 100.0;
        for (int c = 0; c < NC; ++c) rec[(size_t)i * R + D + 1 + c] = nd(rng);
    }
    for (auto &v : pos) v = ud(rng);
    double *drec, *dpos, *dout;
    cudaMalloc(&drec, rec.size() * 8); cudaMalloc(&dpos, pos.size() * 8); cudaMalloc(&dout, (size_t)NC * m * 8);
    cudaMemcpy(drec, rec.data(), rec.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dpos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice);
    SumArgs a{};
    a.rec = drec; a.n_modes = n_modes; a.pos = dpos; a.ps0 = m; a.ps1 = 1; a.n_points = m;
    a.out = dout; a.os0 = 1; a.os1 = NC;
    gsf::poly_constants(gsf::kFastDeg, a.coef);
    const int64_t tile = (int64_t)P * kThreads;
    int64_t tail_pts = (int64_t)(tail_waves * 148 * 8 * (double)tile);   // as launch_sum
    if (tail_pts > m) tail_pts = m;
    int64_t n_big = P > 1 ? (m - tail_pts) / tile : (m + tile - 1) / tile;
    int64_t n_small = P > 1 ? (m - n_big * tile + kThreads - 1) / kThreads : 0;
    a.n_big = n_big;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        gsf_sum_kernel<D, NC, P, 1, gsf::kFastDeg><<<(unsigned)(n_big + n_small), kThreads>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 1 && ms < best) best = ms;
    }
    cudaFree(drec); cudaFree(dpos); cudaFree(dout);
    return (double)m * n_modes / best / 1e6;
}

int main()
{
    printf("deg=%d style=%d unroll=%d |", gsf::kFastDeg, GSF_TUNE_STYLE, GSF_TUNE_UNROLL);
    printf(" d3s P3 1M %.0f 4M %.0f |", run<3, 1, 3>(1000000, 1000, 0.5), run<3, 1, 3>(4000000, 1000, 0.5));
    printf(" d3s P4 4M %.0f | d3s P2 4M %.0f |", run<3, 1, 4>(4000000, 1000, 0.5), run<3, 1, 2>(4000000, 1000, 0.5));
    printf(" d3i P3 1M %.0f 4M %.0f | d3i P2 4M %.0f |", run<3, 3, 3>(1000000, 1000, 0.5), run<3, 3, 3>(4000000, 1000, 0.5), run<3, 3, 2>(4000000, 1000, 0.5));
    printf(" d2s P2 1M %.0f P3 4M %.0f P4 4M %.0f |", run<2, 1, 2>(1000000, 1000, 0.5), run<2, 1, 3>(4000000, 1000, 0.5), run<2, 1, 4>(4000000, 1000, 0.5));
    printf(" d2i P3 4M %.0f | d1s P3 4M %.0f | d3s P1 1M %.0f\n", run<2, 2, 3>(4000000, 1000, 0.5), run<1, 1, 3>(4000000, 1000, 0.5), run<3, 1, 1>(1000000, 1000, 0.5));
    return 0;
}
