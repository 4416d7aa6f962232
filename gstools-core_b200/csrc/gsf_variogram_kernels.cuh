// gsf_variogram_kernels.cuh -- empirical variogram estimators, sm_100a.
//
// Reference: /root/reference/src/variogram.rs
//   variogram_structured :136-178, variogram_ma_structured :190-240,
//   variogram_directional :315-447 (dir_test :243-290), variogram_unstructured :465-545,
//   estimators Matheron / Cressie :41-65, distances Euclid / Haversine :92-123.
//
// The reference walks ALL point pairs once per bin (it parallelises over bins).  Here every pair
// is visited once: a CTA owns 128 "i" points (one per thread, in registers) and streams chunks of
// "j" points through shared memory; each in-range pair is binned by binary search and added to a
// PER-THREAD (bin, direction) accumulator in shared memory (layout [slot][thread]: no atomics, no
// bank conflicts, and every thread's summation order is fixed).  Threads, then CTAs, are combined
// in a fixed order, so the result is run-to-run deterministic.
//
// Bin membership is bit-faithful to the reference for Euclidean distances: instead of comparing
// sqrt(d2) with an edge e, the kernel compares the squared distance d2 -- accumulated exactly like
// Euclid::dist, left to right without FMA contraction -- with t(e) = min{x : sqrt(x) >= e}, which
// the host computes with the correctly rounded sqrt (sqrt is monotone, so sqrt(d2) >= e <=> d2 >=
// t(e)).  The same trick replaces acos in dir_test by a threshold on its argument.  Haversine
// distances go through CUDA's sin/cos/atan2 and agree with libm to the last ulp or two.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gsf {

constexpr int kVarThreads = 128;

enum VarioMode { kVarEuclid = 0, kVarHaversine = 1, kVarDirectional = 2 };

struct VarioArgs {
    const double *pos;     // [D][m]
    const double *f;       // [nf][m]
    int64_t m;
    int nf;
    const double *thr;     // [nb + 1] thresholds on d2 (Euclid / directional) or raw edges (Haversine)
    int nb;                // bins handled by this launch
    int monotone;          // thresholds are non-decreasing and free of NaN: one bin per pair
    int n_dir;             // directional only
    const double *dir;     // [n_dir][D]
    int use_bw;            // bandwidth > 0 (src/variogram.rs:262)
    double bw_thr;         // band distance^2 >= bw_thr  <=>  b_dist >= bandwidth
    double ang_thr;        // |s_prod| / dist <= ang_thr  <=>  acos(angle) >= angles_tol
    int separate;
    int cressie;
    int jc;                // points per j chunk (multiple of 128)
    int n_iblocks;
    int64_t n_tiles;
    const int64_t *tile_prefix;   // [n_iblocks + 1]: first tile of every i block
    double *part_v;               // [gridDim.x][slots]
    unsigned long long *part_c;
};

__device__ __forceinline__ double vario_estimate(int cressie, double d)
{
    return cressie ? __dsqrt_rn(fabs(d)) : __dmul_rn(d, d);   // src/variogram.rs:57-59 / 44-46
}

template <int D, int MODE>
__global__ void __launch_bounds__(kVarThreads) gsf_vario_pairs(VarioArgs a)
{
    extern __shared__ __align__(16) unsigned char vsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_dir = MODE == kVarDirectional ? a.n_dir : 1;
    const int nb = a.nb, slots = n_dir * nb, jc = a.jc;
    double *acc_v = reinterpret_cast<double *>(vsm);                                        // [slots][128]
    unsigned long long *acc_c = reinterpret_cast<unsigned long long *>(acc_v + (size_t)slots * kVarThreads);
    double *s_thr = reinterpret_cast<double *>(acc_c + (size_t)slots * kVarThreads);        // [nb + 1]
    double *s_dir = s_thr + nb + 1;                                                         // [n_dir][D]
    double *s_pos = s_dir + n_dir * D;                                                      // [D][jc]
    double *s_f = s_pos + D * jc;                                                           // [jc]
    double *s_cos = s_f + jc;                                                               // [jc] (Haversine)
    __shared__ int64_t s_tile[2];

    for (int s = 0; s < slots; ++s) {
        acc_v[s * kVarThreads + tid] = 0.0;
        acc_c[s * kVarThreads + tid] = 0ull;
    }
    for (int e = tid; e <= nb; e += kVarThreads) s_thr[e] = a.thr[e];
    if (MODE == kVarDirectional)
        for (int e = tid; e < n_dir * D; e += kVarThreads) s_dir[e] = a.dir[e];

    const double kRad = 0.017453292519943295;   // f64::to_radians: x * (PI / 180)
    const int64_t m = a.m;
    const bool one_field = a.nf == 1;

    for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        if (tid == 0) {   // largest i block whose first tile is <= t
            int lo = 0, hi = a.n_iblocks;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (a.tile_prefix[mid] <= t) lo = mid; else hi = mid;
            }
            s_tile[0] = lo;
            s_tile[1] = ((int64_t)lo * kVarThreads + 1) / jc + (t - a.tile_prefix[lo]);
        }
        __syncthreads();   // tile known; everybody is done with the previous chunk
        const int64_t i = s_tile[0] * kVarThreads + tid;
        const int64_t j0 = s_tile[1] * jc;
        const int cnt = (int)(m - j0 < jc ? m - j0 : jc);
        for (int e = tid; e < cnt; e += kVarThreads) {
#pragma unroll
            for (int q = 0; q < D; ++q) s_pos[q * jc + e] = a.pos[q * m + j0 + e];
            if (one_field) s_f[e] = a.f[j0 + e];
            if (MODE == kVarHaversine) s_cos[e] = cos(a.pos[j0 + e] * kRad);
        }
        const bool valid = i < m;
        double xi[D], fi = 0.0, cos_i = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) xi[q] = valid ? a.pos[q * m + i] : 0.0;
        if (valid && one_field) fi = a.f[i];
        if (valid && MODE == kVarHaversine) cos_i = cos(xi[0] * kRad);
        __syncthreads();
        if (!valid) continue;
        int jj = j0 <= i ? (int)(i + 1 - j0 < cnt ? i + 1 - j0 : cnt) : 0;   // only pairs j > i
#pragma unroll 2
        for (; jj < cnt; ++jj) {
            double df[D], key;
#pragma unroll
            for (int q = 0; q < D; ++q) df[q] = xi[q] - s_pos[q * jc + jj];
            if (MODE == kVarHaversine) {   // src/variogram.rs:108-117
                const double s1 = sin(__dmul_rn(df[0], kRad) / 2.0), s2 = sin(__dmul_rn(df[D - 1], kRad) / 2.0);
                const double arg = __dadd_rn(__dmul_rn(s1, s1),
                                             __dmul_rn(__dmul_rn(cos_i, s_cos[jj]), __dmul_rn(s2, s2)));
                key = 2.0 * atan2(__dsqrt_rn(arg), __dsqrt_rn(__dadd_rn(1.0, -arg)));
            } else {                       // Euclid::dist without the sqrt, src/variogram.rs:93-102
                key = __dmul_rn(df[0], df[0]);
#pragma unroll
                for (int q = 1; q < D; ++q) key = __dadd_rn(key, __dmul_rn(df[q], df[q]));
            }
            // bins b with !(key < thr[b] || key >= thr[b+1])  (src/variogram.rs:397 / :518)
            int b_lo, b_hi;
            const bool searched = a.monotone && key == key;
            if (searched) {
                if (key < s_thr[0] || key >= s_thr[nb]) continue;
                int lo = 0, hi = nb;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_thr[mid] <= key) lo = mid; else hi = mid;
                }
                b_lo = lo;
                b_hi = lo + 1;
            } else {
                b_lo = 0;
                b_hi = nb;
            }
            for (int b = b_lo; b < b_hi; ++b) {
                if (!searched && (key < s_thr[b] || key >= s_thr[b + 1])) continue;
                for (int r = 0; r < n_dir; ++r) {
                    if (MODE == kVarDirectional) {   // dir_test, src/variogram.rs:243-290
                        const double *dr = s_dir + r * D;
                        double s_prod = __dmul_rn(df[0], dr[0]);
#pragma unroll
                        for (int q = 1; q < D; ++q) s_prod = __dadd_rn(s_prod, __dmul_rn(df[q], dr[q]));
                        if (a.use_bw) {
                            double b2 = 0.0;
#pragma unroll
                            for (int q = 0; q < D; ++q) {
                                const double u = __dadd_rn(df[q], -__dmul_rn(s_prod, dr[q]));
                                b2 = q == 0 ? __dmul_rn(u, u) : __dadd_rn(b2, __dmul_rn(u, u));
                            }
                            if (b2 >= a.bw_thr) continue;
                        }
                        if (key > 0.0) {
                            const double angle = __ddiv_rn(fabs(s_prod), __dsqrt_rn(key));
                            if (angle <= a.ang_thr) continue;
                        }
                    }
                    const int slot = (r * nb + b) * kVarThreads + tid;
                    if (one_field) {
                        const double fij = fi - s_f[jj];
                        if (fij == fij) {   // skip no-data values, src/variogram.rs:413 / :524
                            acc_c[slot] += 1ull;
                            acc_v[slot] = __dadd_rn(acc_v[slot], vario_estimate(a.cressie, fij));
                        }
                    } else {
                        double v = acc_v[slot];
                        unsigned long long c = acc_c[slot];
                        for (int q = 0; q < a.nf; ++q) {
                            const double fij = a.f[q * m + i] - a.f[q * m + j0 + jj];
                            if (fij == fij) {
                                c += 1ull;
                                v = __dadd_rn(v, vario_estimate(a.cressie, fij));
                            }
                        }
                        acc_v[slot] = v;
                        acc_c[slot] = c;
                    }
                    if (MODE == kVarDirectional && a.separate) break;   // src/variogram.rs:424-426
                }
            }
        }
    }

    // threads -> CTA in a fixed order: 4 columns per lane, then the xor butterfly
    __syncthreads();
    for (int s = warp; s < slots; s += kVarThreads / 32) {
        const double *pv = acc_v + s * kVarThreads;
        const unsigned long long *pc = acc_c + s * kVarThreads;
        double v = ((pv[lane] + pv[lane + 32]) + pv[lane + 64]) + pv[lane + 96];
        unsigned long long c = pc[lane] + pc[lane + 32] + pc[lane + 64] + pc[lane + 96];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            a.part_v[(size_t)blockIdx.x * slots + s] = v;
            a.part_c[(size_t)blockIdx.x * slots + s] = c;
        }
    }
}

// CTAs -> result, in CTA order.  Pass-local slot (r, b) lands at out[r * nb_total + b0 + b].
__global__ void gsf_vario_reduce(const double *part_v, const unsigned long long *part_c, int n_parts, int n_dir,
                                 int nb, int nb_total, int b0, double *out_v, unsigned long long *out_c)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_dir * nb) return;
    double v = 0.0;
    unsigned long long c = 0ull;
    for (int p = 0; p < n_parts; ++p) {
        v += part_v[(size_t)p * n_dir * nb + s];
        c += part_c[(size_t)p * n_dir * nb + s];
    }
    const int r = s / nb, b = s % nb;
    out_v[r * nb_total + b0 + b] = v;
    out_c[r * nb_total + b0 + b] = c;
}

// ---- structured grids: value_k = sum_e est(f[e] - f[e + k*n1]) over the flattened (n0-k, n1) slab ----
constexpr int kVsThreads = 256;

template <bool MASKED>
__global__ void __launch_bounds__(kVsThreads) gsf_vario_struct(const double *f, const uint8_t *mask, int64_t n0,
                                                               int64_t n1, int cressie, int n_split,
                                                               double *part_v, unsigned long long *part_c)
{
    __shared__ double s_v[kVsThreads / 32];
    __shared__ unsigned long long s_c[kVsThreads / 32];
    const int64_t k = (int64_t)blockIdx.x + 1;
    const int64_t len = (n0 - k) * n1, off = k * n1;
    // split s owns [s*chunk, (s+1)*chunk), chunk a multiple of the CTA width
    int64_t chunk = (len + n_split - 1) / n_split;
    chunk = (chunk + kVsThreads - 1) / kVsThreads * kVsThreads;
    const int64_t e0 = (int64_t)blockIdx.y * chunk, e1 = e0 + chunk < len ? e0 + chunk : len;
    double v = 0.0;
    unsigned long long c = 0ull;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += kVsThreads) {
        if (MASKED && (mask[e] || mask[e + off])) continue;   // src/variogram.rs:223-225
        v = __dadd_rn(v, vario_estimate(cressie, f[e] - f[e + off]));
        c += 1ull;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = v;
        s_c[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kVsThreads / 32; ++w) {
            v += s_v[w];
            c += s_c[w];
        }
        part_v[(size_t)blockIdx.x * n_split + blockIdx.y] = v;
        part_c[(size_t)blockIdx.x * n_split + blockIdx.y] = c;
    }
}

__global__ void gsf_vario_struct_reduce(const double *part_v, const unsigned long long *part_c, int64_t n_k,
                                        int n_split, double *out_v, unsigned long long *out_c)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_k) return;
    double v = 0.0;
    unsigned long long c = 0ull;
    for (int s = 0; s < n_split; ++s) {
        v += part_v[k * n_split + s];
        c += part_c[k * n_split + s];
    }
    out_v[k] = v;
    out_c[k] = c;
}

}  // namespace gsf
