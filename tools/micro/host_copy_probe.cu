// host_copy_probe.cu -- what bounds the pageable-input path on the GPU box's host?
//   * thread scaling of the staging copy (pageable -> pinned): memcpy vs non-temporal AVX2 stores,
//     persistent workers released by a spin barrier (no wake-up cost in the measurement)
//   * cudaHostRegister / cudaHostUnregister cost of the caller's buffer
//   * cudaMemcpyAsync straight from pageable memory (driver-staged), pinned H2D / D2H for reference
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -mavx2 -o host_copy_probe host_copy_probe.cu
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void nt_copy(double *dst, const double *src, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        __m256d c = _mm256_loadu_pd(src + i + 8), d = _mm256_loadu_pd(src + i + 12);
        _mm256_stream_pd(dst + i, a); _mm256_stream_pd(dst + i + 4, b);
        _mm256_stream_pd(dst + i + 8, c); _mm256_stream_pd(dst + i + 12, d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

struct Crew {
    int T;
    std::vector<std::thread> th;
    std::atomic<int> gen{0}, done{0};
    std::atomic<bool> stop{false};
    double *dst = nullptr; const double *src = nullptr; size_t n = 0; int mode = 0; size_t part = 49152;
    std::atomic<size_t> next{0};
    explicit Crew(int t) : T(t)
    {
        for (int i = 1; i < T; ++i) th.emplace_back([this]() { loop(); });
    }
    ~Crew() { stop = true; gen.fetch_add(1); for (auto &t : th) t.join(); }
    void work()
    {
        for (;;) {
            size_t p = next.fetch_add(part);
            if (p >= n) break;
            size_t c = std::min(part, n - p);
            if (mode) nt_copy(dst + p, src + p, c); else memcpy(dst + p, src + p, c * 8);
        }
        done.fetch_add(1);
    }
    void loop()
    {
        int seen = 0;
        for (;;) {
            while (gen.load(std::memory_order_acquire) == seen) _mm_pause();
            if (stop) return;
            seen = gen.load();
            work();
        }
    }
    double run(double *d, const double *s, size_t count, int m)
    {
        dst = d; src = s; n = count; mode = m; next = 0; done = 0;
        double t0 = now_ms();
        gen.fetch_add(1, std::memory_order_release);
        work();
        while (done.load() < T) _mm_pause();
        return now_ms() - t0;
    }
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main()
{
    const size_t n = 3000000;   // 24 MB = C2 positions
    const size_t big = 300000000 / 8 * 8 / 8;   // 300 MB = one rank's C5 shard at 8 GPUs
    printf("hardware_concurrency=%u\n", std::thread::hardware_concurrency());
    double *src = (double *)aligned_alloc(4096, big * 8), *pin = nullptr, *pinout = nullptr, *dev = nullptr;
    for (size_t i = 0; i < big; ++i) src[i] = (double)i;
    CK(cudaMallocHost((void **)&pin, big * 8));
    CK(cudaMallocHost((void **)&pinout, big * 8));
    CK(cudaMalloc((void **)&dev, big * 8));
    memset(pin, 0, big * 8);
    for (size_t cnt : {n, big}) {
        for (int T : {1, 2, 3, 4, 6, 8, 12, 16}) {
            if (T > (int)std::thread::hardware_concurrency()) break;
            Crew crew(T);
            for (int mode = 0; mode < 2; ++mode) {
                double best = 1e9, sum = 0;
                const int reps = cnt == n ? 12 : 4;
                for (int r = 0; r < reps + 2; ++r) {
                    double ms = crew.run(pin, src, cnt, mode);
                    if (r >= 2) { best = std::min(best, ms); sum += ms; }
                }
                printf("copy %4zu MB  T=%2d %s  best %.3f ms (%.1f GB/s)  mean %.3f ms\n", cnt * 8 >> 20, T,
                       mode ? "nt-avx2" : "memcpy ", best, cnt * 8 / best / 1e6, sum / reps);
            }
        }
    }
    // host register cost
    for (size_t cnt : {n, big}) {
        for (int r = 0; r < 3; ++r) {
            double t0 = now_ms();
            CK(cudaHostRegister(src, cnt * 8, cudaHostRegisterDefault));
            double t1 = now_ms();
            CK(cudaHostUnregister(src));
            double t2 = now_ms();
            printf("cudaHostRegister %4zu MB: %.3f ms, unregister %.3f ms\n", cnt * 8 >> 20, t1 - t0, t2 - t1);
        }
    }
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (size_t cnt : {n, big}) {
        for (int r = 0; r < 3; ++r) {
            double t0 = now_ms();
            CK(cudaMemcpyAsync(dev, src, cnt * 8, cudaMemcpyHostToDevice, st));
            double t1 = now_ms();
            CK(cudaStreamSynchronize(st));
            double t2 = now_ms();
            printf("pageable cudaMemcpyAsync H2D %4zu MB: call %.3f ms, +sync %.3f ms (%.1f GB/s)\n", cnt * 8 >> 20,
                   t1 - t0, t2 - t0, cnt * 8 / (t2 - t0) / 1e6);
        }
        for (int r = 0; r < 3; ++r) {
            float ms;
            CK(cudaEventRecord(e0, st));
            CK(cudaMemcpyAsync(dev, pin, cnt * 8, cudaMemcpyHostToDevice, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("pinned H2D %4zu MB: %.3f ms (%.1f GB/s)\n", cnt * 8 >> 20, ms, cnt * 8 / ms / 1e6);
            CK(cudaEventRecord(e0, st));
            CK(cudaMemcpyAsync(pinout, dev, cnt * 8 / 3, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("pinned D2H %4zu MB: %.3f ms (%.1f GB/s)\n", cnt * 8 / 3 >> 20, ms, cnt * 8 / 3 / ms / 1e6);
        }
    }
    // scatter-out direction: pinned -> pageable
    {
        Crew crew(2);
        for (int mode = 0; mode < 2; ++mode) {
            double best = 1e9;
            for (int r = 0; r < 8; ++r) best = std::min(best, crew.run(src, pin, 1000000, mode));
            printf("scatter 8 MB pinned->pageable T=2 %s best %.3f ms (%.1f GB/s)\n", mode ? "nt-avx2" : "memcpy ", best,
                   8.0 / best);
        }
    }
    return 0;
}
