// gsf_krige_kernels.cuh -- kriging field / error-variance kernels, sm_100a.
//
// Reference: krige::calculator_field_krige_and_variance / calculator_field_krige,
// /root/reference/src/krige.rs:24-118.  With K = krig_mat (C x C), V = krig_vecs (C x M), c = cond:
//     Y[i,p]   = sum_j K[j,i] V[j,p]            (krig_fac, :54 / :111)
//     field[p] = sum_i c_i    Y[i,p]            (:55 / :112)
//     error[p] = sum_i V[i,p] Y[i,p]            (:56)
//
// * field only: field = (K c)^T V -- O(C^2 + C M) instead of the reference's O(C^2 M): one small
//   mat-vec (gsf_krige_matvec) and one HBM-bound pass over V (gsf_krige_gemv).
// * with variance: Y = K^T V is an FP64 GEMM (2 C^2 M flops) that is never stored: each CTA owns 64
//   points, walks the condition rows in blocks of 32*RT with DMMA.8x8x4 accumulators and folds every
//   finished block of Y into the two running column sums (x c_i, x V[i,p]) in registers.
//   Operand tiles go through a 2-stage cp.async ring; row strides 36 / 68 doubles (4 mod 16) make
//   the "4 k-rows x 8 consecutive" fragment loads bank-conflict free (see gsf_grid_kernels.cuh).
//   Few points => the condition rows are split over gridDim.y; partial sums are combined in split
//   order by gsf_krige_reduce (deterministic).
// Inputs are zero-padded device copies: Cp = roundup(C, 32*RT), Mp = roundup(M, 64).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gsf_grid_kernels.cuh"

namespace gsf {

constexpr int kKrigeThreads = 128;
constexpr int kKrigeBP = 64;      // points per CTA
constexpr int kKrigeRT = 1;       // 8-row DMMA tiles per warp (2 was measured: 255 regs + spills, no net gain)
constexpr int kKrigeBI = 32 * kKrigeRT;   // condition rows per block
constexpr int kKrigeBK = 32;      // contraction block
constexpr int kKrigeSA = kKrigeBI + 4;   // smem row stride of the K tile (doubles)
constexpr int kKrigeSB = kKrigeBP + 4;   // smem row stride of the V tile
constexpr size_t kKrigeSmem = (size_t)2 * kKrigeBK * (kKrigeSA + kKrigeSB) * sizeof(double);

__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct KrigeArgs {
    const double *mat;     // [Cp][Cp] row-major, mat[j][i]
    const double *vecs;    // [Cp][ldv] row-major, vecs[j][p]
    const double *cond;    // [Cp]
    int64_t cp, ldv, mp;   // padded sizes; mp = padded points of this launch
    double *field, *error; // [splits][mp] partials (or the results when splits == 1)
    int splits;
};

__global__ void __launch_bounds__(kKrigeThreads) gsf_krige_gemm(KrigeArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                 // [2][BK][SA]
    double *sB = sA + 2 * kKrigeBK * kKrigeSA;                         // [2][BK][SB]
    __shared__ double s_red[2][4][kKrigeBP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * kKrigeBP;
    const int64_t n_iblocks = a.cp / kKrigeBI;
    const int64_t ib0 = n_iblocks * blockIdx.y / a.splits, ib1 = n_iblocks * (blockIdx.y + 1) / a.splits;
    const int64_t n_kblocks = a.cp / kKrigeBK;

    auto load_tiles = [&](int st, int64_t i0, int64_t kb) {
        const int64_t j0 = kb * kKrigeBK;
        double *dA = sA + st * kKrigeBK * kKrigeSA;
        double *dB = sB + st * kKrigeBK * kKrigeSB;
        // K tile: 32 rows (j) x BI doubles (i) in 16 B units, 4*RT per thread
        constexpr int kUnitsPerRow = kKrigeBI / 2;
#pragma unroll
        for (int q = 0; q < 4 * kKrigeRT; ++q) {
            const int e = tid + q * kKrigeThreads;
            const int r = e / kUnitsPerRow, c2 = e % kUnitsPerRow;
            cp_async16(dA + r * kKrigeSA + 2 * c2, a.mat + (j0 + r) * a.cp + i0 + 2 * c2);
        }
        // V tile: 32 rows (j) x 64 doubles (p) = 1024 x 16 B, 8 per thread
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int e = tid + q * kKrigeThreads;         // 0..1023
            const int r = e >> 5, c2 = e & 31;
            cp_async16(dB + r * kKrigeSB + 2 * c2, a.vecs + (j0 + r) * a.ldv + p0 + 2 * c2);
        }
        cp_async_commit();
    };

    double fld[8][2], err[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) fld[t][0] = fld[t][1] = err[t][0] = err[t][1] = 0.0;

    for (int64_t ib = ib0; ib < ib1; ++ib) {
        const int64_t i0 = ib * kKrigeBI;
        double acc[kKrigeRT][8][2];
#pragma unroll
        for (int rt = 0; rt < kKrigeRT; ++rt)
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[rt][t][0] = acc[rt][t][1] = 0.0;

        // 2-stage ring, ONE barrier per K block: the barrier at the top both publishes block kb
        // (after this thread's cp.async group landed) and guarantees everybody finished reading
        // the other stage (block kb-1), which the loads for kb+1 then overwrite during the DMMAs.
        load_tiles(0, i0, 0);
        for (int64_t kb = 0; kb < n_kblocks; ++kb) {
            const int st = (int)(kb & 1);
            cp_async_wait<0>();
            __syncthreads();
            if (kb + 1 < n_kblocks) load_tiles(st ^ 1, i0, kb + 1);
            const double *fa = sA + st * kKrigeBK * kKrigeSA + (lane & 3) * kKrigeSA + warp * 8 * kKrigeRT + (lane >> 2);
            const double *fb = sB + st * kKrigeBK * kKrigeSB + (lane & 3) * kKrigeSB + (lane >> 2);
#pragma unroll
            for (int k4 = 0; k4 < kKrigeBK / 4; ++k4) {
                double af[kKrigeRT];
#pragma unroll
                for (int rt = 0; rt < kKrigeRT; ++rt) af[rt] = fa[4 * k4 * kKrigeSA + 8 * rt];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const double bf = fb[4 * k4 * kKrigeSB + 8 * t];
#pragma unroll
                    for (int rt = 0; rt < kKrigeRT; ++rt) dmma884(acc[rt][t][0], acc[rt][t][1], af[rt], bf);
                }
            }
        }
        __syncthreads();   // all reads of the last stages done before the next row block reloads them
        // fold this block of Y into the running column sums
#pragma unroll
        for (int rt = 0; rt < kKrigeRT; ++rt) {
            const int64_t i = i0 + warp * 8 * kKrigeRT + 8 * rt + (lane >> 2);
            const double ci = a.cond[i];
            const double2 *vrow = reinterpret_cast<const double2 *>(a.vecs + i * a.ldv + p0);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double2 v = __ldg(vrow + 4 * t + (lane & 3));
                fld[t][0] = fma(ci, acc[rt][t][0], fld[t][0]);
                fld[t][1] = fma(ci, acc[rt][t][1], fld[t][1]);
                err[t][0] = fma(v.x, acc[rt][t][0], err[t][0]);
                err[t][1] = fma(v.y, acc[rt][t][1], err[t][1]);
            }
        }
    }

    // reduce over the 8 row lanes of the warp (fixed butterfly), then over the 4 warps
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double f = fld[t][e], r = err[t][e];
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) {
                f += __shfl_xor_sync(0xffffffffu, f, o);
                r += __shfl_xor_sync(0xffffffffu, r, o);
            }
            if ((lane >> 2) == 0) {
                s_red[0][warp][8 * t + 2 * lane + e] = f;
                s_red[1][warp][8 * t + 2 * lane + e] = r;
            }
        }
    __syncthreads();
    if (tid < kKrigeBP) {
        const double f = ((s_red[0][0][tid] + s_red[0][1][tid]) + s_red[0][2][tid]) + s_red[0][3][tid];
        const double r = ((s_red[1][0][tid] + s_red[1][1][tid]) + s_red[1][2][tid]) + s_red[1][3][tid];
        a.field[(int64_t)blockIdx.y * a.mp + p0 + tid] = f;
        a.error[(int64_t)blockIdx.y * a.mp + p0 + tid] = r;
    }
}

// out[p] = sum over splits (in split order) of part[s][p]
__global__ void gsf_krige_reduce(const double *part_f, const double *part_e, int splits, int64_t mp, int64_t m,
                                 double *field, double *error)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    double f = 0.0, e = 0.0;
    for (int s = 0; s < splits; ++s) {
        f += part_f[(int64_t)s * mp + p];
        e += part_e[(int64_t)s * mp + p];
    }
    field[p] = f;
    if (error) error[p] = e;
}

// w[j] = sum_i mat[j][i] * cond[i]   (one warp per row, fixed-order lane reduction)
__global__ void gsf_krige_matvec(const double *mat, const double *cond, int64_t cp, double *w)
{
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= cp) return;
    double s = 0.0;
    for (int64_t i = lane; i < cp; i += 32) s = fma(mat[j * cp + i], cond[i], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) w[j] = s;
}

// field[p] = sum_j w[j] * vecs[j][p]   (one thread per point, coalesced rows; HBM-bound)
__global__ void gsf_krige_gemv(const double *w, const double *vecs, int64_t c, int64_t ldv, int64_t m, double *field)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t j = 0;
    for (; j + 4 <= c; j += 4) {
        s0 = fma(w[j], __ldg(vecs + j * ldv + p), s0);
        s1 = fma(w[j + 1], __ldg(vecs + (j + 1) * ldv + p), s1);
        s2 = fma(w[j + 2], __ldg(vecs + (j + 2) * ldv + p), s2);
        s3 = fma(w[j + 3], __ldg(vecs + (j + 3) * ldv + p), s3);
    }
    for (; j < c; ++j) s0 = fma(w[j], __ldg(vecs + j * ldv + p), s0);
    field[p] = (s0 + s1) + (s2 + s3);
}

}  // namespace gsf
