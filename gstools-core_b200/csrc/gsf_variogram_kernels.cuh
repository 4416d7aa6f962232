// gsf_variogram_kernels.cuh -- empirical variogram estimators, sm_100a.
//
// Reference: /root/reference/src/variogram.rs
//   variogram_structured :136-178, variogram_ma_structured :190-240,
//   variogram_directional :315-447 (dir_test :243-290), variogram_unstructured :465-545,
//   estimators Matheron / Cressie :41-65, distances Euclid / Haversine :92-123.
//
// The reference walks ALL point pairs once per bin (it parallelises over bins).  Here every pair
// is visited at most once: the host sorts the points along a Morton curve and lists the tiles
// (128 "i" points x one chunk of "j" points) whose bounding boxes are close enough to hold an
// in-range pair (gsf_variogram_host.inc).  A CTA keeps one i point per thread in registers and
// streams the j chunk through shared memory; the range test of 8 j points runs back to back (pure
// FP64 arithmetic, no divergence), then only the candidates are binned by binary search and added
// to a PER-THREAD (bin, direction) accumulator in shared memory (layout [slot][thread]: no
// atomics, no bank conflicts, and every thread's summation order is fixed).  Threads, then CTAs,
// are combined in a fixed order, so the result is run-to-run deterministic.  Every term of the
// estimators is symmetric in (i, j) bit for bit, so the reordering only changes the summation order.
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
constexpr int kVarBatch = 8;       // j points whose range test runs back to back before any binning

enum VarioMode { kVarEuclid = 0, kVarHaversine = 1, kVarDirectional = 2 };

// shared-memory record of one j point: D = 1: (x, f); D = 2: (x, y, f, cos(lat)); D = 3: (x, y, z, f);
// D >= 4 (generic pair kernel only): (x_0 .. x_{D-1}, f) padded to an even number of doubles
__host__ __device__ constexpr int vario_rec(int d) { return d == 1 ? 2 : d <= 3 ? 4 : (d + 2) & ~1; }
constexpr int kVarMaxDim = 8;      // gsf_vario_pairs is instantiated for D = 1..8 (Morton sort / window kernel: D <= 3)

// dir_test, src/variogram.rs:243-290, for one direction `dr`; df = x_i - x_j, key = |df|^2.
// The angle test is `|s_prod| / sqrt(key) <= ang_thr` (see above).  Away from the tolerance the
// sign of s_prod^2 - ang_thr^2 * key decides it without the square root and the division: the
// computed quotient is within 3e-16 of the real one, the margins a2_lo / a2_hi (ang_thr^2 * (1 -+
// 1e-13)) are far wider, and only pairs inside that sliver (or with operands near the ends of the
// double range) take the exact path -- the outcome is the same for every input.
struct VarioDirTest {
    int use_bw;            // bandwidth > 0 (src/variogram.rs:262)
    double bw_thr;         // band distance^2 >= bw_thr  <=>  b_dist >= bandwidth
    double ang_thr;        // reject when angle <= ang_thr; < 0: the test never rejects
    double a2_lo, a2_hi;
};

struct VarioArgs {
    const double *pos;     // [D][m]
    const double *f;       // [nf][m]
    int64_t m;
    int nf;
    const double *thr;     // [nb + 1] thresholds on d2 (Euclid / directional) or raw edges (Haversine)
    int nb;                // bins handled by this launch
    int monotone;          // thresholds are non-decreasing and free of NaN: one bin per pair
    double pre_lo, pre_hi; // range pre-filter: a pair with key < pre_lo or key >= pre_hi is in no bin
                           // (thr[0], thr[nb] when monotone, else NaN = filter nothing)
    int n_dir;             // directional only
    const double *dir;     // [n_dir][D]
    VarioDirTest dt;
    int separate;
    int cressie;
    int jc;                // points per j chunk (multiple of 128)
    int64_t n_tiles;
    const int2 *tiles;     // (i block, j chunk) of every tile that can hold an in-range pair
    int flush_every;       // tiles after which the 32-bit per-thread counts move to the 64-bit totals
    double *part_v;        // [gridDim.x][slots]
    unsigned long long *part_c;
};

template <int D>
__device__ __forceinline__ bool vario_dir_pass(const double (&df)[D], const double *dr, double key, const VarioDirTest &p)
{
    double s_prod = __dmul_rn(df[0], dr[0]);
#pragma unroll
    for (int q = 1; q < D; ++q) s_prod = __dadd_rn(s_prod, __dmul_rn(df[q], dr[q]));
    // Both tests are pure predicates, so the cheap angle decision goes first: it rejects most pairs
    // (a pi/8 cone holds 8 % of the directions) before the band distance is ever computed.
    bool angle_known = !(key > 0.0 && p.ang_thr >= 0.0);   // the test does not apply: passes
    if (!angle_known) {
        const double s2 = s_prod * s_prod;
        if (key > 1e-280 && key < 1e280 && s2 > 1e-280 && s2 < 1e280) {
            if (s2 < p.a2_lo * key) return false;
            angle_known = s2 > p.a2_hi * key;
        }
    }
    if (p.use_bw) {
        double b2 = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) {
            const double w = __dadd_rn(df[q], -__dmul_rn(s_prod, dr[q]));
            b2 = q == 0 ? __dmul_rn(w, w) : __dadd_rn(b2, __dmul_rn(w, w));
        }
        if (b2 >= p.bw_thr) return false;
    }
    if (!angle_known) {
        const double angle = __ddiv_rn(fabs(s_prod), __dsqrt_rn(key));
        if (angle <= p.ang_thr) return false;
    }
    return true;
}

__device__ __forceinline__ double vario_estimate(int cressie, double d)
{
    return cressie ? __dsqrt_rn(fabs(d)) : __dmul_rn(d, d);   // src/variogram.rs:57-59 / 44-46
}

// distance key of the pair (xi, record pj): squared Euclidean distance accumulated like Euclid::dist
// (src/variogram.rs:93-102, no FMA), or the Haversine distance itself (:108-117)
template <int D, int MODE>
__device__ __forceinline__ double vario_key(const double (&xi)[D], double cos_i, const double *pj, double (&df)[D])
{
#pragma unroll
    for (int q = 0; q < D; ++q) df[q] = xi[q] - pj[q];
    if (MODE == kVarHaversine) {
        const double kRad = 0.017453292519943295;   // f64::to_radians: x * (PI / 180)
        const double s1 = sin(__dmul_rn(df[0], kRad) / 2.0), s2 = sin(__dmul_rn(df[D - 1], kRad) / 2.0);
        const double arg = __dadd_rn(__dmul_rn(s1, s1), __dmul_rn(__dmul_rn(cos_i, pj[3 % vario_rec(D)]), __dmul_rn(s2, s2)));
        return 2.0 * atan2(__dsqrt_rn(arg), __dsqrt_rn(__dadd_rn(1.0, -arg)));
    }
    double key = __dmul_rn(df[0], df[0]);
#pragma unroll
    for (int q = 1; q < D; ++q) key = __dadd_rn(key, __dmul_rn(df[q], df[q]));
    return key;
}

template <int D, int MODE>
__global__ void __launch_bounds__(kVarThreads) gsf_vario_pairs(VarioArgs a)
{
    constexpr int W = vario_rec(D);
    extern __shared__ __align__(16) unsigned char vsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_dir = MODE == kVarDirectional ? a.n_dir : 1;
    const int nb = a.nb, slots = n_dir * nb, jc = a.jc;
    double *s_rec = reinterpret_cast<double *>(vsm);                                    // [jc][W]
    double *acc_v = s_rec + (size_t)jc * W;                                             // [slots][128]
    double *s_thr = acc_v + (size_t)slots * kVarThreads;                                // [nb + 1]
    double *s_dir = s_thr + nb + 1;                                                     // [n_dir][D]
    unsigned long long *tot_c = reinterpret_cast<unsigned long long *>(s_dir + n_dir * D);   // [slots]
    unsigned int *acc_c = reinterpret_cast<unsigned int *>(tot_c + slots);              // [slots][128]
    __shared__ double s_own[kVarThreads][W < 4 ? 4 : W];  // the CTA's i points: position, field value, cos(lat)
    __shared__ unsigned int s_q[kVarThreads / 32][64];    // per-warp queue of candidate pairs (lane << 16 | j)

    for (int s = 0; s < slots; ++s) {
        acc_v[s * kVarThreads + tid] = 0.0;
        acc_c[s * kVarThreads + tid] = 0u;
    }
    for (int s = tid; s < slots; s += kVarThreads) tot_c[s] = 0ull;
    for (int e = tid; e <= nb; e += kVarThreads) s_thr[e] = a.thr[e];
    if (MODE == kVarDirectional)
        for (int e = tid; e < n_dir * D; e += kVarThreads) s_dir[e] = a.dir[e];

    const int64_t m = a.m;
    const bool one_field = a.nf == 1;
    const double pre_lo = a.pre_lo, pre_hi = a.pre_hi;
    int since_flush = 0;

    // 32-bit per-thread counts -> 64-bit CTA totals (integer sums: any schedule gives the same result)
    auto flush_counts = [&]() {
        __syncthreads();
        for (int s = warp; s < slots; s += kVarThreads / 32) {
            unsigned int *pc = acc_c + s * kVarThreads;
            unsigned long long c = (unsigned long long)pc[lane] + pc[lane + 32] + pc[lane + 64] + pc[lane + 96];
            pc[lane] = pc[lane + 32] = pc[lane + 64] = pc[lane + 96] = 0u;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) tot_c[s] += c;
        }
        __syncthreads();
    };

    for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int2 tile = a.tiles[t];
        const int64_t i = (int64_t)tile.x * kVarThreads + tid;
        const int64_t j0 = (int64_t)tile.y * jc;
        const int cnt = (int)(m - j0 < jc ? m - j0 : jc);
        __syncthreads();   // everybody is done with the previous chunk
        for (int e = tid; e < cnt; e += kVarThreads) {
            double *r = s_rec + e * W;
#pragma unroll
            for (int q = 0; q < D; ++q) r[q] = a.pos[q * m + j0 + e];
            r[D] = one_field ? a.f[j0 + e] : 0.0;
            if (MODE == kVarHaversine) r[3] = cos(a.pos[j0 + e] * 0.017453292519943295);
        }
        const bool valid = i < m;
        double xi[D], fi = 0.0, cos_i = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) xi[q] = valid ? a.pos[q * m + i] : 0.0;
        if (valid && one_field) fi = a.f[i];
        if (valid && MODE == kVarHaversine) cos_i = cos(xi[0] * 0.017453292519943295);
#pragma unroll
        for (int q = 0; q < D; ++q) s_own[tid][q] = xi[q];
        s_own[tid][D] = fi;
        if (MODE == kVarHaversine || D < 3) s_own[tid][3] = cos_i;
        __syncthreads();
        // One candidate pair: bin it (binary search), test the directions, add it to THIS thread's
        // accumulator column.  `li` is the warp lane that owns the i point.
        auto process = [&](unsigned int entry) {
            const int li = (int)(entry >> 16), jj = (int)(entry & 0xffffu);
            const double *pi = s_own[warp * 32 + li];
            const double *pj = s_rec + jj * W;
            double xo[D], df[D];
#pragma unroll
            for (int q = 0; q < D; ++q) xo[q] = pi[q];
            const double key = vario_key<D, MODE>(xo, pi[3], pj, df);
            // bins b with !(key < thr[b] || key >= thr[b+1])
            int b_lo = 0, b_hi = nb;
            const bool searched = a.monotone && key == key;
            if (searched) {
                int lo = 0, hi = nb;   // thr[lo] <= key < thr[hi]
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_thr[mid] <= key) lo = mid; else hi = mid;
                }
                b_lo = lo;
                b_hi = lo + 1;
            }
            for (int b = b_lo; b < b_hi; ++b) {
                if (!searched && (key < s_thr[b] || key >= s_thr[b + 1])) continue;
                for (int r = 0; r < n_dir; ++r) {
                    if (MODE == kVarDirectional && !vario_dir_pass<D>(df, s_dir + r * D, key, a.dt)) continue;
                    const int slot = (r * nb + b) * kVarThreads + tid;
                    if (one_field) {
                        const double fij = pi[D] - pj[D];
                        if (fij == fij) {   // skip no-data values, src/variogram.rs:413 / :524
                            acc_c[slot] += 1u;
                            acc_v[slot] = __dadd_rn(acc_v[slot], vario_estimate(a.cressie, fij));
                        }
                    } else {
                        const int64_t gi = (int64_t)tile.x * kVarThreads + warp * 32 + li;
                        double v = acc_v[slot];
                        unsigned int c = acc_c[slot];
                        for (int q = 0; q < a.nf; ++q) {
                            const double fij = a.f[q * m + gi] - a.f[q * m + j0 + jj];
                            if (fij == fij) {
                                c += 1u;
                                v = __dadd_rn(v, vario_estimate(a.cressie, fij));
                            }
                        }
                        acc_v[slot] = v;
                        acc_c[slot] = c;
                    }
                    if (MODE == kVarDirectional && a.separate) break;   // src/variogram.rs:424-426
                }
            }
        };

        // only pairs j > i; lanes past the end of the data own no pair
        const int jstart = !valid ? cnt : j0 > i ? 0 : (int)(i + 1 - j0 < cnt ? i + 1 - j0 : cnt);
        const int jb0 = __shfl_sync(0xffffffffu, jstart, 0) / kVarBatch * kVarBatch;   // lane 0 starts first
        // The range test runs for every (lane, j) with all lanes busy.  The candidates -- a few per
        // j and warp -- go to a per-warp queue in ballot order (deterministic) and are binned 32 at
        // a time, again with all lanes busy: the expensive part no longer runs at the lane
        // utilisation of the hit rate.
        unsigned int *queue = s_q[warp];
        int qn = 0;
        for (int jb = jb0; jb < cnt; jb += kVarBatch) {
            unsigned mask = 0u;
#pragma unroll
            for (int u = 0; u < kVarBatch; ++u) {
                double df[D];
                const double key = vario_key<D, MODE>(xi, cos_i, s_rec + (jb + u) * W, df);
                // candidate unless excluded from every bin (src/variogram.rs:397 / :518); NaN stays in
                const bool in = jb + u >= jstart && jb + u < cnt && !(key < pre_lo) && !(key >= pre_hi);
                mask |= (unsigned)in << u;
            }
            if (!__any_sync(0xffffffffu, mask != 0u)) continue;
#pragma unroll 1
            for (int u = 0; u < kVarBatch; ++u) {
                const bool in = (mask >> u) & 1u;
                const unsigned int hits = __ballot_sync(0xffffffffu, in);
                if (!hits) continue;
                if (in) queue[qn + __popc(hits & ((1u << lane) - 1u))] = ((unsigned int)lane << 16) | (unsigned int)(jb + u);
                qn += __popc(hits);
                if (qn >= 32) {
                    __syncwarp();
                    process(queue[lane]);
                    qn -= 32;
                    const unsigned int moved = lane < qn ? queue[32 + lane] : 0u;
                    __syncwarp();
                    if (lane < qn) queue[lane] = moved;
                }
            }
        }
        __syncwarp();
        if (lane < qn) process(queue[lane]);
        if (++since_flush >= a.flush_every) {
            flush_counts();
            since_flush = 0;
        }
    }

    // threads -> CTA in a fixed order: 4 columns per lane, then the xor butterfly
    flush_counts();
    for (int s = warp; s < slots; s += kVarThreads / 32) {
        const double *pv = acc_v + s * kVarThreads;
        double v = ((pv[lane] + pv[lane + 32]) + pv[lane + 64]) + pv[lane + 96];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            a.part_v[(size_t)blockIdx.x * slots + s] = v;
            a.part_c[(size_t)blockIdx.x * slots + s] = tot_c[s];
        }
    }
}

// ---- isotropic Euclidean fast path -------------------------------------------------------------
// For sorted points the distances inside one tile span only a few bins.  The host lists every tile
// once per window of kVarWin consecutive bins its bounding boxes can reach; the kernel keeps that
// window's thresholds in (uniform) registers: the bin of a pair is the count of window thresholds
// below its key -- no search, no threshold loads -- and the per-thread accumulators shrink to
// kVarWin columns of shared memory, so occupancy no longer depends on the number of bins.
// After the tile the 128 threads are combined in a fixed order and added to the CTA's running
// per-bin totals (global memory, touched once per tile and bin).
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

constexpr int kVarWin = 8;
constexpr int kVarWinDirs = 4;     // most directions the window kernel handles (3 axes + 1)

struct VarioIsoArgs {
    const double *pos;     // [D][m], Morton order
    const double *f;       // [nf][m]
    int64_t m;
    int nf;
    const double *thr;     // [nb + 1], non-decreasing, no NaN
    int nb;
    int jc;
    int64_t n_tiles;
    const int3 *tiles;     // (i block, j chunk, first bin of the window)
    // directional only (see VarioArgs)
    int n_dir;
    const double *dir;
    VarioDirTest dt;
    int separate;
    double *part_v;        // [gridDim.x][n_dir][nb], zeroed by the host
    unsigned long long *part_c;
};

// Shared-memory access through 32-bit shared addresses held in registers: with plain pointers into
// the dynamic shared array the compiler re-derives the window base (S2UR + 4 uniform instructions)
// in front of every load of the pair loop.
__device__ __forceinline__ uint32_t vario_saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vario_lds2(uint32_t addr, double &x, double &y)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ double vario_lds1(uint32_t addr)
{
    double x;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
    return x;
}
// one accumulator cell: (sum, count) in 16 bytes
__device__ __forceinline__ void vario_acc_add(uint32_t addr, double e, unsigned int n)
{
    double v;
    unsigned long long c;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=d"(v), "=l"(c) : "r"(addr));
    v = __dadd_rn(v, e);
    c += n;
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "d"(v), "l"(c) : "memory");
}
// idx += (key >= t): one compare and one predicated add
__device__ __forceinline__ void vario_count_ge(int &idx, double key, double t)
{
    asm("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.s32 %0, %0, 1; }" : "+r"(idx) : "d"(key), "d"(t));
}

template <int D, bool CRESSIE, bool DIRECTIONAL>
__global__ void __launch_bounds__(kVarThreads) gsf_vario_iso(VarioIsoArgs a)
{
    constexpr int W = vario_rec(D), K = kVarWin, MAXS = K * (DIRECTIONAL ? kVarWinDirs : 1);
    extern __shared__ __align__(16) unsigned char vsm[];
    const int n_dir = DIRECTIONAL ? a.n_dir : 1, slots = n_dir * K;
    double *s_rec = reinterpret_cast<double *>(vsm);                        // [jc][W]
    double *w_acc = s_rec + (size_t)a.jc * W;                                // [n_dir][K][128] x (sum, count): one column per thread
    __shared__ double s_wv[kVarThreads / 32][MAXS];
    __shared__ unsigned long long s_wc[kVarThreads / 32][MAXS];
    __shared__ double s_dir[DIRECTIONAL ? kVarWinDirs * D : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = a.nb, jc = a.jc;
    const int64_t m = a.m;
    const bool one_field = a.nf == 1;
    double *cta_v = a.part_v + (size_t)blockIdx.x * n_dir * nb;
    unsigned long long *cta_c = a.part_c + (size_t)blockIdx.x * n_dir * nb;
    if (DIRECTIONAL && tid < n_dir * D) s_dir[tid] = a.dir[tid];
    const uint32_t rec_addr = vario_saddr(s_rec);
    const uint32_t acc_addr = vario_saddr(w_acc) + tid * 16;   // slot s at acc_addr + s * 2048

    for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int3 tile = a.tiles[t];
        const int64_t i = (int64_t)tile.x * kVarThreads + tid;
        const int64_t j0 = (int64_t)tile.y * jc;
        const int b0 = tile.z;
        const int cnt = (int)(m - j0 < jc ? m - j0 : jc);
        __syncthreads();   // everybody is done with the previous chunk
        for (int e = tid; e < cnt; e += kVarThreads) {
            double *r = s_rec + e * W;
#pragma unroll
            for (int q = 0; q < D; ++q) r[q] = a.pos[q * m + j0 + e];
            r[D] = one_field ? a.f[j0 + e] : 0.0;
        }
        for (int s = 0; s < slots; ++s) {
            w_acc[(s * kVarThreads + tid) * 2] = 0.0;
            reinterpret_cast<unsigned long long *>(w_acc)[(s * kVarThreads + tid) * 2 + 1] = 0ull;
        }
        const bool valid = i < m;
        double xi[D], fi = 0.0;
#pragma unroll
        for (int q = 0; q < D; ++q) xi[q] = valid ? a.pos[q * m + i] : 0.0;
        if (valid && one_field) fi = a.f[i];
        double th[K + 1];   // beyond the last edge: +inf, `key >= inf` never holds for a finite key
#pragma unroll
        for (int k = 0; k <= K; ++k) th[k] = b0 + k <= nb ? a.thr[b0 + k] : INFINITY;
        __syncthreads();

        const int jstart = !valid ? cnt : j0 > i ? 0 : (int)(i + 1 - j0 < cnt ? i + 1 - j0 : cnt);
        const int jfirst = __shfl_sync(0xffffffffu, jstart, 0);   // lane 0 starts first
        // the pair loop, specialised on "one field" and on "tile touches the diagonal" (only there
        // the per-lane start index matters; invalid lanes of the last i block also take that path)
        auto pair_loop = [&](auto one_tag, auto diag_tag) {
            constexpr bool ONE = decltype(one_tag)::value, DIAG = decltype(diag_tag)::value;
#pragma unroll 4
            for (int jj = jfirst; jj < cnt; ++jj) {
                const uint32_t pj = rec_addr + jj * (W * 8);
                double rj[4], df[D];
                vario_lds2(pj, rj[0], rj[1]);                       // D = 1: (x, f); D >= 2: (x, y)
                if (D == 3) vario_lds2(pj + 16, rj[2], rj[3]);      // (z, f)
                df[0] = xi[0] - rj[0];
                double key = __dmul_rn(df[0], df[0]);               // Euclid::dist, src/variogram.rs:93-102
                if (D >= 2) {
                    df[D >= 2 ? 1 : 0] = xi[D >= 2 ? 1 : 0] - rj[1];
                    key = __dadd_rn(key, __dmul_rn(df[D >= 2 ? 1 : 0], df[D >= 2 ? 1 : 0]));
                }
                if (D == 3) {
                    df[D - 1] = xi[D - 1] - rj[2];
                    key = __dadd_rn(key, __dmul_rn(df[D - 1], df[D - 1]));
                }
                if ((DIAG && jj < jstart) || !(key >= th[0]) || key >= th[K]) continue;   // src/variogram.rs:397 / :518
                // thresholds are sorted: the bin is the number of interior thresholds <= key
                int idx = 0;
#pragma unroll
                for (int k = 1; k < K; ++k) vario_count_ge(idx, key, th[k]);
                double e;
                unsigned int n;
                if (ONE) {
                    const double fj = D == 1 ? rj[1] : D == 2 ? vario_lds1(pj + 16) : rj[3];
                    const double fij = fi - fj;
                    if (fij != fij) continue;            // skip no-data values, src/variogram.rs:413 / :524
                    n = 1u;
                    e = CRESSIE ? __dsqrt_rn(fabs(fij)) : __dmul_rn(fij, fij);
                } else {
                    e = 0.0;
                    n = 0u;
                    for (int q = 0; q < a.nf; ++q) {
                        const double fij = a.f[q * m + i] - a.f[q * m + j0 + jj];
                        if (fij == fij) {
                            n += 1u;
                            e = __dadd_rn(e, CRESSIE ? __dsqrt_rn(fabs(fij)) : __dmul_rn(fij, fij));
                        }
                    }
                }
                if (!DIRECTIONAL) {
                    vario_acc_add(acc_addr + idx * (kVarThreads * 16), e, n);
                } else {
                    for (int r = 0; r < n_dir; ++r) {
                        if (!vario_dir_pass<D>(df, s_dir + r * D, key, a.dt)) continue;
                        vario_acc_add(acc_addr + (r * K + idx) * (kVarThreads * 16), e, n);
                        if (a.separate) break;   // src/variogram.rs:424-426
                    }
                }
            }
        };
        const bool diag = j0 <= (int64_t)tile.x * kVarThreads + kVarThreads - 1 || (int64_t)(tile.x + 1) * kVarThreads > m;
        if (one_field) {
            if (diag) pair_loop(TrueTag(), TrueTag()); else pair_loop(TrueTag(), FalseTag());
        } else {
            if (diag) pair_loop(FalseTag(), TrueTag()); else pair_loop(FalseTag(), FalseTag());
        }
        __syncwarp();
        // 128 threads -> one value per (direction, window bin), fixed order
        for (int s = 0; s < slots; ++s) {
            double vk = w_acc[(s * kVarThreads + tid) * 2];
            unsigned long long cw = reinterpret_cast<unsigned long long *>(w_acc)[(s * kVarThreads + tid) * 2 + 1];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                vk += __shfl_xor_sync(0xffffffffu, vk, o);
                cw += __shfl_xor_sync(0xffffffffu, cw, o);
            }
            if (lane == 0) {
                s_wv[warp][s] = vk;
                s_wc[warp][s] = cw;
            }
        }
        __syncthreads();
        if (tid < slots && b0 + tid % K < nb) {
            const double vv = ((s_wv[0][tid] + s_wv[1][tid]) + s_wv[2][tid]) + s_wv[3][tid];
            const unsigned long long cc = s_wc[0][tid] + s_wc[1][tid] + s_wc[2][tid] + s_wc[3][tid];
            const int o = (tid / K) * nb + b0 + tid % K;
            cta_v[o] += vv;
            cta_c[o] += cc;
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

// ---- structured grids: value_k = sum_{i < n0-k, c} est(f[i][c] - f[i+k][c]) ----------------------
// A CTA owns a block of 32 consecutive lags and walks (64-row x 32-column) tiles of the field: the
// tile rows i0..i0+64 ("A") and the 96 rows starting at i0 + k0 ("B") sit in shared memory, lane =
// column, and each of the 8 warps owns 4 consecutive lags.  Going down the rows a thread keeps the
// 4 B values of its lags in registers and slides that window by one row per step, so a pair costs
// half a shared-memory load and every field element is fetched from L2 ~13x less often than in a
// lag-by-lag sweep.  The per-lag sums stay in registers across tiles; lanes are combined by a fixed
// butterfly at the end and the n_split partial sums per lag are added in order by
// gsf_vario_struct_reduce: run-to-run deterministic.
constexpr int kVsThreads = 256;
constexpr int kVsRows = 64;      // A rows per tile
constexpr int kVsLags = 32;      // lags per CTA (4 per warp)
constexpr int kVsCols = 32;      // columns per tile

template <bool MASKED, bool CRESSIE>
__global__ void __launch_bounds__(kVsThreads) gsf_vario_struct(const double *f, const uint8_t *mask, int64_t n0,
                                                               int64_t n1, int n_split, double *part_v,
                                                               unsigned long long *part_c)
{
    constexpr int R = kVsRows, KB = kVsLags;
    __shared__ double sA[R][kVsCols];
    __shared__ double sB[R + KB][kVsCols];
    __shared__ uint8_t mA[R][kVsCols];        // 1 = excluded (masked)
    __shared__ uint8_t mB[R + KB][kVsCols];   // 1 = excluded (masked, or outside the field)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = 1 + (int64_t)blockIdx.x * KB;   // first lag of this CTA
    const int wl = 4 * warp;                           // this warp's first lag, relative to k0
    const int64_t n_rb = (n0 - k0 + R - 1) / R, n_cb = (n1 + kVsCols - 1) / kVsCols;
    const int64_t n_t = n_rb * n_cb;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned long long cn[4] = {0ull, 0ull, 0ull, 0ull};

    auto pair = [&](double a, uint8_t ma, double b, uint8_t mb, int j) {
        if (!(ma | mb)) {
            const double d = a - b;
            v[j] = __dadd_rn(v[j], CRESSIE ? __dsqrt_rn(fabs(d)) : __dmul_rn(d, d));
            if (MASKED) cn[j] += 1ull;
        }
    };

    for (int64_t t = blockIdx.y; t < n_t; t += n_split) {
        const int64_t i0 = (t / n_cb) * R, col = (t % n_cb) * kVsCols + lane;
        const bool col_ok = col < n1;
        __syncthreads();
        for (int r = warp; r < R + KB; r += kVsThreads / 32) {
            if (r < R) {
                const int64_t gi = i0 + r;
                const bool ok = col_ok && gi < n0;
                sA[r][lane] = ok ? f[gi * n1 + col] : 0.0;
                mA[r][lane] = MASKED && ok ? mask[gi * n1 + col] : 0;
            }
            const int64_t gi = i0 + k0 + r;
            const bool ok = col_ok && gi < n0;
            sB[r][lane] = ok ? f[gi * n1 + col] : 0.0;
            mB[r][lane] = !ok || (MASKED && mask[gi * n1 + col]) ? 1 : 0;
        }
        __syncthreads();
        double b0 = sB[wl][lane], b1 = sB[wl + 1][lane], b2 = sB[wl + 2][lane], b3;
        uint8_t q0 = mB[wl][lane], q1 = mB[wl + 1][lane], q2 = mB[wl + 2][lane], q3;
#pragma unroll 2
        for (int r = 0; r < R; r += 4) {
            double a;
            uint8_t ma;
            a = sA[r][lane]; ma = MASKED ? mA[r][lane] : 0; b3 = sB[r + wl + 3][lane]; q3 = mB[r + wl + 3][lane];
            pair(a, ma, b0, q0, 0); pair(a, ma, b1, q1, 1); pair(a, ma, b2, q2, 2); pair(a, ma, b3, q3, 3);
            a = sA[r + 1][lane]; ma = MASKED ? mA[r + 1][lane] : 0; b0 = sB[r + wl + 4][lane]; q0 = mB[r + wl + 4][lane];
            pair(a, ma, b1, q1, 0); pair(a, ma, b2, q2, 1); pair(a, ma, b3, q3, 2); pair(a, ma, b0, q0, 3);
            a = sA[r + 2][lane]; ma = MASKED ? mA[r + 2][lane] : 0; b1 = sB[r + wl + 5][lane]; q1 = mB[r + wl + 5][lane];
            pair(a, ma, b2, q2, 0); pair(a, ma, b3, q3, 1); pair(a, ma, b0, q0, 2); pair(a, ma, b1, q1, 3);
            a = sA[r + 3][lane]; ma = MASKED ? mA[r + 3][lane] : 0; b2 = sB[r + wl + 6][lane]; q2 = mB[r + wl + 6][lane];
            pair(a, ma, b3, q3, 0); pair(a, ma, b0, q0, 1); pair(a, ma, b1, q1, 2); pair(a, ma, b2, q2, 3);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double vj = v[j];
        unsigned long long cj = cn[j];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            vj += __shfl_xor_sync(0xffffffffu, vj, o);
            cj += __shfl_xor_sync(0xffffffffu, cj, o);
        }
        const int64_t k = k0 + wl + j;
        if (lane == 0 && k < n0) {
            part_v[(size_t)(k - 1) * n_split + blockIdx.y] = vj;
            part_c[(size_t)(k - 1) * n_split + blockIdx.y] = MASKED ? cj : 0ull;
        }
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
