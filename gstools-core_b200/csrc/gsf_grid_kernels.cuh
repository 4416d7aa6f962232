// gsf_grid_kernels.cuh -- structured-grid fast path, sm_100a.
//
// When the points are a rectilinear grid  x_j = (x0[j0], x1[j1], x2[j2])  flattened in C order
// (GSTools mesh_type="structured"), the phase separates per axis and
//     A_i cos(<k_i, x_j> - theta_i) = Re( E0[i,j0] * E1[i,j1] * F[i,j2] ),
//     E_a[i,j] = exp(i k_ia x_a[j]),   F[i,j] = A_i exp(i (k_i,last x_last[j] - theta_i)).
// The sum over modes is then a real GEMM with contraction length 2N:
//     out[r, c] = sum_{i} ( Gr[r,i] * Fr[i,c] - Gi[r,i] * Fi[i,c] ),   r = (j0,j1), c = (j2, comp)
// with G = E0*E1 formed on the fly (4 FP64 ops per (r, i), amortised over all columns).  That is
// 2 + 4/n_last FP64 FMAs per point*mode instead of 15 for the general kernel.
//
// The contraction runs on the FP64 tensor path (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05
// has no f64 kind).  Measured on B200: DMMA streams reach 18.5 T FMA/s (64 FMA/clk/SM), DFMA
// streams 17.1 T (register-file read limited) -- tools/micro/dmma_probe.cu.
//
//   gsf_grid_tables  : E_a and F tables via sincospi (O(N * sum n_a) work, negligible)
//   gsf_grid_gemm    : CTA = 4 warps, tile = 32 rows x (8*NT) columns, K-block of 16 modes;
//                      F tiles by 1-D TMA bulk copy into a 2-stage ring, G tile computed by the
//                      CTA into shared memory, each warp owns 8 rows x NT column tiles of DMMA
//                      accumulators.  Optional split over modes (grid.z) with a fixed-order
//                      fix-up by the last-arriving CTA (deterministic, no float atomics).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gsf_kernels.cuh"

namespace gsf {

#ifndef GSF_GRID_MIN_CTAS
#define GSF_GRID_MIN_CTAS 1
#endif
constexpr int kGridThreads = 128;   // 4 warps
constexpr int kGridRows = 32;       // rows per CTA (8 per warp)
constexpr int kGridBK = 16;         // modes per K block (32 contraction steps)
constexpr int kGridMaxNT = 16;      // column tiles (of 8) per CTA
constexpr int kGStride = 2 * kGridBK + 4;   // doubles per G row in smem (36: conflict-free LDS.64)
// doubles per F row in smem/global tiles: 8*NT + 4, i.e. 4 or 12 mod 16.  A 64-bit shared load is
// served per half-warp; the 16 lanes of a DMMA B fragment half read 4 k-rows x 4 columns, and
// with this stride the four rows start 8 banks apart (0/8/16/24): one wavefront per half-warp.
// (8 mod 16 puts rows 0/2 and 1/3 on the same banks: measured 2x the wavefronts, smem-bound.)
__host__ __device__ constexpr int grid_bnp(int nt) { return 8 * nt + 4; }
__host__ __device__ constexpr size_t grid_smem_bytes(int nt)
{
    return (size_t)(2 * (2 * kGridBK * grid_bnp(nt)) + 2 * (kGridRows * kGStride)) * sizeof(double) + 16;
}

// ------------------------------------------------------------------------------------------
// Table builder.  rec = pre-processed mode records of gsf_prep_modes (kh[D], -th, A[NC]).
//   E tables (axes 0 .. D-2): layout [n_a][Npad] complex (re, im), modes fastest.
//   F table (last axis): tiled for the GEMM kernel,
//       F[colblock][kappa = 2*i + {0: re, 1: -im}][BNp]  with column = j*NC + comp inside a block;
//     BNp = 8*NT + 4 doubles (padding zero) so the smem image is bank-conflict free as copied.
struct GridTableArgs {
    const double *rec; int rec_doubles; int dim; int nc;
    int64_t n_modes, n_modes_pad;
    const double *axis[3]; int64_t axis_n[3]; int64_t axis_s[3];
    double *E[2];               // E0, E1 (E1 unused for dim 2)
    double *F;
    int64_t n_cols;             // NC * n_last
    int cols_per_block;         // 8*NT
    int bnp;                    // grid_bnp(NT)
    int n_col_blocks;
    double scale;
    // Mode groups (tensor-structured modes, e.g. the Fourier method's lattice k = (kx_a, ky_b)): `mode_group`
    // consecutive modes share every wave-vector component except the last, so their F rows can be
    // summed BEFORE the GEMM:  sum_i Re(G_i F_i) = sum_i' Re(G_i' sum_il F_(i',il)).  n_modes then counts
    // the groups (the GEMM's contraction shrinks by the group size); 1 = every mode on its own.
    int mode_group;
};

__global__ void gsf_grid_tables(GridTableArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // mode (fastest)
    const int ax = blockIdx.z;
    if (i >= a.n_modes_pad) return;
    const bool last = ax == a.dim - 1;
    const bool live = i < a.n_modes;
    const int group = a.mode_group;
    const double *rec = a.rec + (live ? i * group : 0) * a.rec_doubles;   // first mode of the group
    if (last && group > 1) return;   // the F rows of mode groups come from gsf_grid_group_rows
    for (int64_t j = blockIdx.y; j < a.axis_n[ax]; j += gridDim.y) {   // axis index
        double c = 0.0, s = 0.0;
        if (live) {
            const double x = a.axis[ax][j * a.axis_s[ax]];
            const double t = last ? fma(rec[ax], x, rec[a.dim]) : __dmul_rn(rec[ax], x);
            sincospi(t, &s, &c);
        }
        if (!last) {
            double2 *E = reinterpret_cast<double2 *>(a.E[ax]);
            E[j * a.n_modes_pad + i] = make_double2(c, s);
            continue;
        }
        for (int comp = 0; comp < a.nc; ++comp) {
            const double amp = live ? __dmul_rn(rec[a.dim + 1 + comp], a.scale) : 0.0;
            const int64_t col = j * a.nc + comp;
            const int64_t blk = col / a.cols_per_block;
            const int64_t cin = col - blk * a.cols_per_block;
            double *base = a.F + (blk * (2 * a.n_modes_pad) + 2 * i) * a.bnp + cin;
            base[0] = __dmul_rn(amp, c);
            base[a.bnp] = -__dmul_rn(amp, s);
        }
    }
}

// F rows of mode groups (GridTableArgs::mode_group > 1): row i' = sum over the group's members of
// A_il * exp(i*pi*(kh_last,il * x_j - th_il)), each member with its own last wave number, phase and
// amplitude, accumulated in member order.  Thread = one point j of the last axis (coalesced writes of
// the tiled F rows); all threads of a block walk the same group, so the member records are uniform
// (broadcast) loads.  blockIdx.y strides over the groups.
__global__ void __launch_bounds__(128) gsf_grid_group_rows(GridTableArgs a)
{
    const int ax = a.dim - 1;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.axis_n[ax]) return;
    const int group = a.mode_group;
    const double x = a.axis[ax][j * a.axis_s[ax]];
    for (int64_t i = blockIdx.y; i < a.n_modes_pad; i += gridDim.y) {
        double re[3] = {0.0, 0.0, 0.0}, im[3] = {0.0, 0.0, 0.0};
        if (i < a.n_modes) {
            const double *rec = a.rec + i * group * a.rec_doubles;
            for (int il = 0; il < group; ++il) {
                const double *r = rec + (int64_t)il * a.rec_doubles;
                double c, s;
                sincospi(fma(__ldg(r + ax), x, __ldg(r + a.dim)), &s, &c);
                for (int comp = 0; comp < a.nc; ++comp) {
                    const double amp = __dmul_rn(__ldg(r + a.dim + 1 + comp), a.scale);
                    re[comp] = fma(amp, c, re[comp]);
                    im[comp] = fma(amp, s, im[comp]);
                }
            }
        }
        for (int comp = 0; comp < a.nc; ++comp) {
            const int64_t col = j * a.nc + comp;
            const int64_t blk = col / a.cols_per_block;
            const int64_t cin = col - blk * a.cols_per_block;
            double *base = a.F + (blk * (2 * a.n_modes_pad) + 2 * i) * a.bnp + cin;
            base[0] = re[comp];
            base[a.bnp] = -im[comp];
        }
    }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

struct GridGemmArgs {
    const double *E0, *E1;      // [n_a][Npad] complex
    const double *F;            // tiled, see above
    int64_t n_modes_pad;
    int64_t n_rows;             // rows of this launch
    int64_t row0;               // global row index of the launch's first row (for j0/j1 split)
    int64_t n1;                 // size of axis 1 (dim 3) -- row = j0*n1 + j1
    int64_t n_cols;             // valid columns overall
    int cols_per_block, bnp;
    int64_t n_last;             // points per row; column = j*nc + comp
    double *out;                // out[comp*os_comp + (row*n_last + j)*os_pt], rows relative to the launch
    int64_t os_comp, os_pt;
    double offset[3];           // per component, added once
    int nc;
    int k_splits;               // grid.z
    double *partial;            // [k_splits][n_rows][n_col_blocks*cols_per_block] when k_splits > 1
    unsigned int *tile_counter; // one per (row tile, col block), self-resetting
};

template <int D, int NT>
__global__ void __launch_bounds__(kGridThreads, GSF_GRID_MIN_CTAS) gsf_grid_gemm(GridGemmArgs a)
{
    constexpr int BN = 8 * NT;
    constexpr int BNP = grid_bnp(NT);
    constexpr int FSTAGE = 2 * kGridBK * BNP;            // doubles per F stage
    constexpr int GSTAGE = kGridRows * kGStride;         // doubles per G stage
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sF = reinterpret_cast<double *>(smem_raw);                  // [2][FSTAGE]
    double *sG = sF + 2 * FSTAGE;                                       // [2][GSTAGE]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sG + 2 * GSTAGE);      // [2]
    __shared__ unsigned int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row_tile = blockIdx.x;
    const int col_blk = blockIdx.y;
    const int split = blockIdx.z;
    const int64_t r_cta = row_tile * kGridRows;

    // K range of this split, in K blocks
    const int64_t n_kblocks = a.n_modes_pad / kGridBK;
    const int64_t kb0 = n_kblocks * split / a.k_splits;
    const int64_t kb1 = n_kblocks * (split + 1) / a.k_splits;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const double *Fblk = a.F + (int64_t)col_blk * (2 * a.n_modes_pad) * BNP;
    auto issue_f = [&](int64_t kb) {
        const int st = (int)((kb - kb0) & 1);
        const uint32_t bytes = FSTAGE * sizeof(double);
        mbar_expect_tx(&bar[st], bytes);
        bulk_g2s(sF + st * FSTAGE, Fblk + kb * FSTAGE, bytes, &bar[st]);
    };
    if (tid == 0) {
        if (kb0 < kb1) issue_f(kb0);
        if (kb0 + 1 < kb1) issue_f(kb0 + 1);
    }

    // G producer mapping: 32 rows x 16 modes = 512 complex per K block, 4 per thread:
    // thread -> mode m = tid % 16, rows (tid / 16) + 8*q, q = 0..3
    const int gm = tid & 15;
    const int gr = tid >> 4;
    int64_t e0_off[4], e1_off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int64_t r = r_cta + gr + 8 * q;
        if (r >= a.n_rows) r = a.n_rows - 1;
        r += a.row0;
        if (D == 3) {
            const int64_t j0 = r / a.n1, j1 = r - j0 * a.n1;
            e0_off[q] = j0 * a.n_modes_pad;
            e1_off[q] = j1 * a.n_modes_pad;
        } else {
            e0_off[q] = r * a.n_modes_pad;
            e1_off[q] = 0;
        }
    }
    const double2 *E0 = reinterpret_cast<const double2 *>(a.E0);
    const double2 *E1 = reinterpret_cast<const double2 *>(a.E1);

    double2 pe0[4], pe1[4];
    auto load_e = [&](int64_t kb) {
        const int64_t m = kb * kGridBK + gm;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pe0[q] = __ldg(E0 + e0_off[q] + m);
            if (D == 3) pe1[q] = __ldg(E1 + e1_off[q] + m);
        }
    };
    auto store_g = [&](int st) {
        double *g = sG + st * GSTAGE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double2 v = pe0[q];
            if (D == 3) {
                const double re = fma(pe0[q].x, pe1[q].x, -__dmul_rn(pe0[q].y, pe1[q].y));
                const double im = fma(pe0[q].x, pe1[q].y, __dmul_rn(pe0[q].y, pe1[q].x));
                v = make_double2(re, im);
            }
            *reinterpret_cast<double2 *>(g + (gr + 8 * q) * kGStride + 2 * gm) = v;
        }
    };

    double acc[NT][2];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;

    // Software pipeline, ONE __syncthreads per K block:
    //   iteration kb reads stage st = (kb-kb0)&1 (F landed by TMA two blocks ago, G written one
    //   block ago), writes G(kb+1) into the other stage from registers prefetched one block ago,
    //   and issues the E loads for kb+2.  The trailing barrier both publishes G(kb+1) and
    //   releases stage st for the TMA refill / the G(kb+2) stores.
    if (kb0 < kb1) {
        load_e(kb0);
        store_g(0);
        if (kb0 + 1 < kb1) load_e(kb0 + 1);
    }
    __syncthreads();
    for (int64_t kb = kb0; kb < kb1; ++kb) {
        const int st = (int)((kb - kb0) & 1);
        if (kb + 1 < kb1) {
            store_g(st ^ 1);                           // G(kb+1)
            if (kb + 2 < kb1) load_e(kb + 2);          // prefetch E for kb+2
        }
        mbar_wait(&bar[st], (uint32_t)(((kb - kb0) >> 1) & 1));

        const double *g = sG + st * GSTAGE + (warp * 8 + (lane >> 2)) * kGStride + (lane & 3);
        const double *f = sF + st * FSTAGE + (lane & 3) * BNP + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < 2 * kGridBK / 4; ++k4) {
            const double af = g[4 * k4];
#pragma unroll
            for (int t = 0; t < NT; ++t) dmma884(acc[t][0], acc[t][1], af, f[(4 * k4) * BNP + 8 * t]);
        }
        __syncthreads();
        if (tid == 0 && kb + 2 < kb1) issue_f(kb + 2);
    }

    // ---- epilogue.  D fragment: row = lane/4, cols = 2*(lane%4) + {0,1} of each 8-wide tile.
    const int64_t r_loc = r_cta + warp * 8 + (lane >> 2);
    const bool row_ok = r_loc < a.n_rows;
    const int c_in = 2 * (lane & 3);
    if (a.k_splits == 1) {
        if (row_ok) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t col = (int64_t)col_blk * BN + 8 * t + c_in + e;
                    if (col < a.n_cols) {
                        const int64_t j = a.nc == 1 ? col : col / a.nc;
                        const int comp = a.nc == 1 ? 0 : (int)(col - j * a.nc);
                        a.out[comp * a.os_comp + (r_loc * a.n_last + j) * a.os_pt] = acc[t][e] + a.offset[comp];
                    }
                }
            }
        }
        return;
    }

    // split over modes: write the partial, the last CTA of the tile sums them in split order
    const int64_t pcols = (int64_t)gridDim.y * BN;
    double *part = a.partial + ((int64_t)split * a.n_rows) * pcols;
    if (row_ok) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double2 v = make_double2(acc[t][0], acc[t][1]);
            *reinterpret_cast<double2 *>(part + r_loc * pcols + (int64_t)col_blk * BN + 8 * t + c_in) = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int *cnt = a.tile_counter + row_tile * gridDim.y + col_blk;
        const unsigned int prev = atomicAdd(cnt, 1u);
        s_last = prev == (unsigned int)(a.k_splits - 1);
        if (s_last) *cnt = 0;                          // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (row_ok) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            double s0 = 0.0, s1 = 0.0;
            for (int sp = 0; sp < a.k_splits; ++sp) {
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(
                    a.partial + ((int64_t)sp * a.n_rows + r_loc) * pcols + (int64_t)col_blk * BN + 8 * t + c_in));
                s0 += v.x;
                s1 += v.y;
            }
            const double r2[2] = {s0, s1};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t col = (int64_t)col_blk * BN + 8 * t + c_in + e;
                if (col < a.n_cols) {
                    const int64_t j = a.nc == 1 ? col : col / a.nc;
                    const int comp = a.nc == 1 ? 0 : (int)(col - j * a.nc);
                    a.out[comp * a.os_comp + (r_loc * a.n_last + j) * a.os_pt] = r2[e] + a.offset[comp];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-path peak: independent DMMA.8x8x4 accumulator chains, all SMs, full occupancy.
// The roofline denominator of gsf_grid_gemm (256 FMA per warp-level DMMA).
constexpr int kDmmaChains = 2;
constexpr int kDmmaUnroll = 32;
__global__ void __launch_bounds__(256) gsf_dmma_peak_kernel(double *sink, int iters, double a, double b)
{
    double c0[kDmmaChains], c1[kDmmaChains];
#pragma unroll
    for (int i = 0; i < kDmmaChains; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kDmmaUnroll; ++u)
#pragma unroll
            for (int i = 0; i < kDmmaChains; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kDmmaChains; ++i) s += c0[i] + c1[i];
    if (s == 123.456) sink[0] = s;
}

// ------------------------------------------------------------------------------------------
// Structured-grid detection for DEVICE-resident positions (the host-resident case is handled by
// detect_grid_host on the CPU).  Same logic: candidate axis lengths from the first index at which
// each slower coordinate changes, then an exact bitwise verification of every point.

// first[a] = min { j : pos[a, j] != pos[a, 0] bitwise }  (left at n_points if the row is constant)
__global__ void gsf_grid_first_change(const double *pos, int64_t ps0, int64_t ps1, int n_rows, int64_t n_points,
                                      unsigned long long *first)
{
    const int a = blockIdx.y;
    if (a >= n_rows) return;
    const double *row = pos + a * ps0;
    const long long b0 = __double_as_longlong(row[0]);
    unsigned long long best = (unsigned long long)n_points;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_points; j += (int64_t)gridDim.x * blockDim.x)
        if (__double_as_longlong(row[j * ps1]) != b0) { best = (unsigned long long)j; break; }
    // one global atomic per CTA: warp butterfly, then shared memory (a same-address atomic per thread
    // would serialise ~1e6 operations in L2)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    __shared__ unsigned long long s_best;
    if (threadIdx.x == 0) s_best = (unsigned long long)n_points;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && best < (unsigned long long)n_points) atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0 && s_best < (unsigned long long)n_points) atomicMin(first + a, s_best);
}

struct GridCheckArgs {
    const double *pos; int64_t ps0, ps1;
    int dim;
    int64_t n[3];
    int64_t n_points;
    int *flag;                  // set to 1 on any mismatch
    double *axes;               // out: the axis vectors, concatenated (axis 0, 1, 2)
};

// every point must equal (axis0[j0], axis1[j1], axis2[j2]) bit for bit, where axis a is read from
// pos itself (points whose other indices are 0); also writes the axis vectors to `axes`
__global__ void gsf_grid_check(GridCheckArgs a)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_points) return;
    int64_t idx[3] = {0, 0, 0}, rem = j;
    for (int d = a.dim - 1; d >= 0; --d) {
        idx[d] = rem % a.n[d];
        rem /= a.n[d];
    }
    int64_t stride = 1, axis_off = 0;
    for (int d = 0; d < a.dim; ++d) axis_off += a.n[d];
    bool bad = false;
    for (int d = a.dim - 1; d >= 0; --d) {
        axis_off -= a.n[d];
        const double want = a.pos[d * a.ps0 + (idx[d] * stride) * a.ps1];   // axis value: other indices 0
        const double have = a.pos[d * a.ps0 + j * a.ps1];
        bad |= __double_as_longlong(want) != __double_as_longlong(have);
        if (j == idx[d] * stride) a.axes[axis_off + idx[d]] = have;         // this point defines axis d
        stride *= a.n[d];
    }
    if (bad) atomicExch(a.flag, 1);
}

}  // namespace gsf
