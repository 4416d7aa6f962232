// gsf_hostcopy.cpp -- the staging copy between the caller's pageable arrays and the pinned ring.
//
// Compiled by the host compiler alone (see Makefile) so that the AVX2 path can be selected at run
// time (function multiversioning via the target attribute) without building the whole library with
// -mavx2.  Streaming (non-temporal) stores: the destination of a staging copy is read next by the
// GPU's copy engine (or written next by it), never by this core, so allocating its lines in the
// cache only costs a read-for-ownership per line -- 3 memory transfers per byte instead of 2.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace {

__attribute__((target("avx2"))) void copy_nt_avx2(double *dst, const double *src, size_t n)
{
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        const __m256d a = _mm256_loadu_pd(src + i), b = _mm256_loadu_pd(src + i + 4);
        const __m256d c = _mm256_loadu_pd(src + i + 8), d = _mm256_loadu_pd(src + i + 12);
        _mm256_stream_pd(dst + i, a);
        _mm256_stream_pd(dst + i + 4, b);
        _mm256_stream_pd(dst + i + 8, c);
        _mm256_stream_pd(dst + i + 12, d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

void copy_nt_sse2(double *dst, const double *src, size_t n)
{
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15)) { dst[i] = src[i]; ++i; }
    for (; i + 8 <= n; i += 8) {
        const __m128d a = _mm_loadu_pd(src + i), b = _mm_loadu_pd(src + i + 2);
        const __m128d c = _mm_loadu_pd(src + i + 4), d = _mm_loadu_pd(src + i + 6);
        _mm_stream_pd(dst + i, a);
        _mm_stream_pd(dst + i + 2, b);
        _mm_stream_pd(dst + i + 4, c);
        _mm_stream_pd(dst + i + 6, d);
    }
    for (; i < n; ++i) dst[i] = src[i];
    _mm_sfence();
}

bool have_avx2()
{
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

}  // namespace

// Contiguous copy of n doubles.  Small copies stay on memcpy (cache-resident data, no sfence).
extern "C" void gsf_copy_stream(double *dst, const double *src, size_t n)
{
    if (n < 2048) {
        memcpy(dst, src, n * sizeof(double));
        return;
    }
    if (have_avx2())
        copy_nt_avx2(dst, src, n);
    else
        copy_nt_sse2(dst, src, n);
}

// dst[j] = src[j*stride], j < n  (gather of a strided view into the contiguous ring)
extern "C" void gsf_gather_strided(double *dst, const double *src, size_t n, ptrdiff_t stride)
{
    for (size_t j = 0; j < n; ++j) dst[j] = src[static_cast<ptrdiff_t>(j) * stride];
}

// dst[j*stride] = src[j], j < n
extern "C" void gsf_scatter_strided(double *dst, const double *src, size_t n, ptrdiff_t stride)
{
    for (size_t j = 0; j < n; ++j) dst[static_cast<ptrdiff_t>(j) * stride] = src[j];
}
