// gsf_kernels.cuh -- sm_100a kernels of the randomization-method field summation.
//
// Math.  The reference evaluates, per point x_j and mode k_i (src/field.rs:57-60,145-152,242-245)
//     z1_i*cos(phi) + z2_i*sin(phi),   phi = <k_i, x_j>
// which is  A_i * cos(phi - theta_i)  with  A_i = hypot(z1_i, z2_i), theta_i = atan2(z2_i, z1_i).
// A pre-pass (gsf_prep_modes) converts every mode ONCE into half-turn units
//     kh = k/pi,  th = theta/pi,  amplitude(s) A (times spectrum_factor / projector p_a(k)),
// so the hot loop needs a single cos(pi*t), t = <kh, x> - th, per point*mode:
//     t   : D DFMA                      (init with -th)
//     r   : 3 DADD   n = rint(t) by the 1.5*2^52 trick, r = t - n exact, |r| <= 1/2
//     s   : 1 DMUL   s = r*r
//     u   : 1 DADD + 4 DFMA   u ~ sqrt2*cos(pi/2 r): monic minimax polynomial of degree 5 in s
//                             (cospi_poly.cuh; GSF_POLY_DEGREE=6 adds one DFMA, see below)
//     y   : 1 DFMA   cos(pi r) = u*u - 1
//     acc : NC DFMA  acc_a += A_a * (+-y), sign = parity of n (one integer IMAD on y's high word)
// = D + 10 + NC FP64-pipe instructions per point*mode (14 for the 3-D scalar field).
//
// Mapping.  One CTA = kThreads threads = a tile of P*kThreads/L points.  A group of L lanes shares
// one point set and splits the modes (lane sub handles modes sub, sub+L, ...); partial sums are
// combined with a fixed-order xor-shuffle tree.  L = 1 for large M (pure per-thread sums); L > 1
// only widens small problems so all 148 SMs have work.  Mode records stream through a two-stage
// shared-memory ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier); with L = 1 every
// LDS is a warp-wide broadcast.  Positions are read once per thread (coalesced, SoA rows).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "cospi_poly.cuh"

namespace gsf {

#ifndef GSF_THREADS
#define GSF_THREADS 128
#endif
#ifndef GSF_MODE_BLOCK
#define GSF_MODE_BLOCK 256
#endif
constexpr int kThreads = GSF_THREADS;   // threads per CTA (128: measured best of 64/128/256, tools/micro/tune_sum.cu)
#ifndef GSF_TAIL_P
#define GSF_TAIL_P 1
#endif
constexpr int kTailP = GSF_TAIL_P;  // points per thread of the short tiles that end a launch
constexpr int kModeBlock = GSF_MODE_BLOCK;   // mode records per shared-memory stage
constexpr int kStages = 2;
constexpr int kMaxTemplateDim = 8;   // = GSF_MAX_DIM: gsf_sum_kernel<D,...> is instantiated for D <= 8
constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52

enum Kind : int { kScalar = 0, kIncompr = 1, kFourier = 2 };

// Polynomial form of the hot loop (cospi_poly.cuh):
//   GSF_POLY_MONIC 0: u = U(s) by 6 DFMA, y = u*u - 1.  The first Horner step fma(U6, s, U5) has TWO
//      constant operands; a DFMA takes at most one from the constant bank / a uniform register, so
//      ptxas keeps both in registers: three register-file reads = 3 issue cycles instead of 2.
//   GSF_POLY_MONIC 1 (default): U = c*Q with Q monic: v = s + Q_{deg-1} (DADD), deg-1 DFMA,
//      w = v*v - 1/c^2; the factor c^2 is folded into the mode amplitudes by gsf_prep_modes.  No
//      instruction with three register sources in the polynomial.
//   GSF_POLY_DEGREE (monic form only): degree of U in s.
//      6: |cos error| <= 6.5e-16 (3 ulp at 1) -- far below the rounding of the phase itself
//         (|phi| * 1.1e-16) and 6 orders inside the 1e-9 sigma contract;
//      5 (default): <= 2.2e-13 absolute per term, measured <= 1e-12 sigma against the oracle on
//         every BASELINE config (tests/test_parity_gpu.py, profiles/poly_degree_r2.md) -- still three
//         orders inside the contract, one DFMA (2 issue cycles of ~31) cheaper per point*mode;
//      4: <= 1.9e-10 per term: too close to the contract, kept for the measurement only.
#ifndef GSF_POLY_MONIC
#define GSF_POLY_MONIC 1
#endif
#ifndef GSF_POLY_DEGREE
#define GSF_POLY_DEGREE 5
#endif
constexpr bool kPolyMonic = GSF_POLY_MONIC != 0;
// Every point x mode kernel is templated on the degree DEG.  Two are instantiated: kHiDeg for
// small (launch-latency-bound) problems, where the extra DFMA is free, and kFastDeg for the
// throughput-bound ones; the host picks ONE degree per call from the total work (choose_degree in
// gsfield.cu), so chunking and device sharding never change a result.
constexpr int kHiDeg = 6;
constexpr int kFastDeg = kPolyMonic ? GSF_POLY_DEGREE : 6;
static_assert(kFastDeg >= 4 && kFastDeg <= 6, "GSF_POLY_DEGREE must be 4, 5 or 6");
// FP64-pipe instructions of one cos(pi t): 3 DADD (reduction) + DMUL + (DADD + (deg-1) DFMA | 6 DFMA) + DFMA
__host__ __device__ constexpr int cos_slots(int deg) { return 3 + 1 + deg + 1; }
// amplitude factor the pre-pass folds into the records of the point x mode kernels
inline double amp_factor(int deg)
{
    if (!kPolyMonic) return 1.0;
    return deg == 4 ? GSF_Q4_S : deg == 5 ? GSF_Q5_S : GSF_Q6_S;
}
// the constants handed to the kernels as SumArgs::coef (c[0..deg-1]: Horner constants, c[6]: -1/c^2
// for the monic form, U6 otherwise)
inline void poly_constants(int deg, double *c)
{
    for (int i = 0; i < 8; ++i) c[i] = 0.0;
    if (!kPolyMonic) {
        c[0] = GSF_U0; c[1] = GSF_U1; c[2] = GSF_U2; c[3] = GSF_U3; c[4] = GSF_U4; c[5] = GSF_U5; c[6] = GSF_U6;
    } else if (deg == 4) {
        c[0] = GSF_Q4_0; c[1] = GSF_Q4_1; c[2] = GSF_Q4_2; c[3] = GSF_Q4_3; c[6] = -(GSF_Q4_E);
    } else if (deg == 5) {
        c[0] = GSF_Q5_0; c[1] = GSF_Q5_1; c[2] = GSF_Q5_2; c[3] = GSF_Q5_3; c[4] = GSF_Q5_4; c[6] = -(GSF_Q5_E);
    } else {
        c[0] = GSF_Q6_0; c[1] = GSF_Q6_1; c[2] = GSF_Q6_2; c[3] = GSF_Q6_3; c[4] = GSF_Q6_4; c[5] = GSF_Q6_5;
        c[6] = -(GSF_Q6_E);
    }
}

// doubles per pre-processed mode record: kh[D], th, A[NC], padded to an even count so every
// record (and every block of records) is a multiple of 16 bytes (TMA bulk-copy granularity).
__host__ __device__ constexpr int rec_doubles(int D, int NC) { return (D + 1 + NC + 1) & ~1; }

// ------------------------------------------------------------------------------------------
// Mode pre-processing: one thread per mode.  Inputs are the caller's (possibly strided) arrays.
struct PrepArgs {
    const double *k;  int64_t ks0, ks1;     // (D, N)
    const double *z1; int64_t z1s;
    const double *z2; int64_t z2s;
    const double *sf; int64_t sfs;          // nullptr unless fourier
    double *rec;                            // N * rec_doubles
    int64_t n_modes;
    int dim;
    int incompr;                            // NC = dim if set, else 1
    double scale;                           // folded into the amplitudes (fused post-scale, 8 f1)
};

// x * (1/pi) with 1/pi carried as a double-double: the result is the correctly rounded quotient
// in all but ~1e-16 of cases, so kh adds no error beyond one rounding of k/pi.
__device__ __forceinline__ double div_pi(double x)
{
    const double inv_pi_hi = 0x1.45f306dc9c883p-2;   // 0.31830988618379069
    const double inv_pi_lo = -0x1.6b01ec5417056p-56; // 1/pi - inv_pi_hi
    return fma(x, inv_pi_hi, __dmul_rn(x, inv_pi_lo));
}

// One mode -> one record.  `k_of(d)`, `z1`, `z2`, `sf` are the caller's raw values of this mode;
// shared by gsf_prep_modes (records in global memory) and gsf_small_kernel (records built per CTA in
// shared memory), so both produce bit-identical records.
template <class KOf>
__device__ __forceinline__ void prep_record(int D, bool incompr, double scale, double z1, double z2, bool has_sf,
                                            double sf, KOf k_of, double *rec)
{
    const int NC = incompr ? D : 1;
    const int R = rec_doubles(D, NC);
    double amp = hypot(z1, z2);
    const double th = div_pi(atan2(z2, z1));
    if (has_sf) amp = __dmul_rn(sf, amp);              // src/field.rs:243
    if (scale != 1.0) amp = __dmul_rn(amp, scale);

    double kk = 0.0;
    for (int d = 0; d < D; ++d) {
        const double kd = k_of(d);
        rec[d] = div_pi(kd);
        kk = __dadd_rn(kk, __dmul_rn(kd, kd));          // ShortVec::dot, src/short_vec.rs:31-33
    }
    rec[D] = -th;                                       // the dot-product chain starts from -theta
    if (!incompr) {
        rec[D + 1] = amp;
    } else {
        // projector, src/field.rs:138,148,151:  k_2 = k0/|k|^2 ; p0 = 1 - k0*k_2 ; pa = -(ka*k_2)
        const double k0 = k_of(0);
        const double k_2 = __ddiv_rn(k0, kk);           // NaN for k = 0, as in the reference
        rec[D + 1] = __dmul_rn(__dadd_rn(1.0, -__dmul_rn(k0, k_2)), amp);
        for (int d = 1; d < D; ++d) rec[D + 1 + d] = __dmul_rn(-__dmul_rn(k_of(d), k_2), amp);
    }
    for (int d = D + 1 + NC; d < R; ++d) rec[d] = 0.0;
}

__global__ void gsf_prep_modes(PrepArgs a)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_modes) return;
    const int NC = a.incompr ? a.dim : 1;
    prep_record(a.dim, a.incompr != 0, a.scale, a.z1[i * a.z1s], a.z2[i * a.z2s], a.sf != nullptr,
                a.sf ? a.sf[i * a.sfs] : 1.0, [&](int d) { return a.k[d * a.ks0 + i * a.ks1]; },
                a.rec + i * rec_doubles(a.dim, NC));
}

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS).
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// The seven polynomial constants travel as kernel parameters (SumArgs::coef, constant bank 0): as
// literals ptxas re-materialises them with UMOV pairs inside the mode loop, and every non-FP64
// instruction there costs an issue cycle the FP64 pipe could have used.
struct PolyCoef {
    double c[7];
};
// (-1)^rint(t) * cos(pi*r): cos(pi*t) for the reduced argument with the sign applied.
// cos_slots(DEG) FP64-pipe instructions + 1 integer multiply-add: adding parity<<31 to the high word
// modulo 2^32 is exactly an XOR of the sign bit.
template <int DEG>
__device__ __forceinline__ double cospi_signed(double t, const PolyCoef &c)
{
    const double tn = __dadd_rn(t, kMagic);     // low mantissa bits = rint(t)
    const double nf = __dadd_rn(tn, -kMagic);   // rint(t) as a double
    const double r = __dadd_rn(t, -nf);         // exact, |r| <= 1/2
    const double s = __dmul_rn(r, r);
    double u = kPolyMonic ? __dadd_rn(s, c.c[DEG - 1]) : fma(c.c[6], s, c.c[5]);
#pragma unroll
    for (int q = DEG - 2; q >= 0; --q) u = fma(u, s, c.c[q]);
    const double y = kPolyMonic ? fma(u, u, c.c[6]) : fma(u, u, -1.0);   // monic: c[6] = -1/c^2
    const uint32_t hi = static_cast<uint32_t>(__double2hiint(y)) +
                        (static_cast<uint32_t>(__double2loint(tn)) << 31);
    return __hiloint2double(static_cast<int>(hi), __double2loint(y));
}

struct SumArgs {
    const double *rec;            // pre-processed mode records
    int64_t n_modes;
    const double *pos;            // element strides: pos[a*ps0 + j*ps1]
    int64_t ps0, ps1;
    int64_t n_points;
    double *out;                  // out[a*os0 + j*os1]
    int64_t os0, os1;
    double offset[3];             // added once per output component (fused post-offset)
    int64_t n_big;                // number of leading tiles with P points per thread (rest: 1 point)
    double coef[8];               // GSF_U0..U6: kernel parameters live in constant bank 0, which
                                  // DFMA reads as a direct operand (no register, no RF read port)
    int dim;                      // read by gsf_sum_kernel_anyd only (the templates carry D)
};

// D: spatial dimension; NC: accumulators per point (1 scalar/fourier, D incompr);
// P: points per thread; L: lanes sharing one point (modes split over them).
// Source-level schedule knobs, chosen per (D, NC, P) by measurement on B200
// (tools/micro/tune_sum.cu): ptxas' instruction order decides how often the operand-reuse cache
// hits, and that is worth several percent on an FP64-bound loop.
//   style 0: one point after the other; style 1: all P points in lockstep, step by step.
#ifndef GSF_TUNE_STYLE
// (re-measured with the degree-5 polynomial, profiles/tune_sum_r2_deg5.txt: unroll 16 is worth
//  +0.6 .. 1.5 % on the 2-D / 3-D kernels, the 2-D scalar P = 3 kernel keeps 4, 1-D prefers style 0)
template <int D, int NC, int P>
__host__ __device__ constexpr int gsf_style() { return D == 1 ? 0 : 1; }
template <int D, int NC, int P>
__host__ __device__ constexpr int gsf_unroll()
{
    return (D == 2 && NC == 1 && P == 3) ? 4 : ((D == 1 || D > 3 || (NC == 3 && P == 4) || (D == 2 && NC == 1 && P == 1)) ? 8 : 16);
}
#else
template <int D, int NC, int P>
__host__ __device__ constexpr int gsf_style() { return GSF_TUNE_STYLE; }
template <int D, int NC, int P>
__host__ __device__ constexpr int gsf_unroll() { return GSF_TUNE_UNROLL; }
#endif

// Body for one CTA tile starting at point tile0 (P points per thread).
template <int D, int NC, int P, int L, int DEG>
__device__ __forceinline__ void sum_tile(const SumArgs &a, const int64_t tile0,
                                         double (*s_rec)[kModeBlock * rec_doubles(D, NC)], uint64_t *s_bar)
{
    constexpr int R = rec_doubles(D, NC);
    constexpr int kGroups = kThreads / L;            // point slots per CTA pass

    const int tid = threadIdx.x;
    const int sub = tid % L;                         // which slice of the modes
    const int grp = tid / L;                         // which point slot

    // ---- positions -> registers (coalesced: consecutive groups read consecutive points)
    double x[P][D];
    int64_t jpt[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int64_t j = tile0 + (int64_t)p * kGroups + grp;
        jpt[p] = j;
        const int64_t jc = j < a.n_points ? j : a.n_points - 1;   // clamp: tail lanes recompute
#pragma unroll
        for (int d = 0; d < D; ++d) x[p][d] = __ldg(a.pos + d * a.ps0 + jc * a.ps1);
    }

    double acc[P][NC];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[p][c] = sub == 0 ? a.offset[c < 3 ? c : 0] : 0.0;

    const PolyCoef coef = {{a.coef[0], a.coef[1], a.coef[2], a.coef[3], a.coef[4], a.coef[5], a.coef[6]}};

    // ---- mode ring: thread 0 is the producer
    const int64_t n_blocks = (a.n_modes + kModeBlock - 1) / kModeBlock;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&s_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int64_t b) {
        const int st = (int)(b % kStages);
        const int64_t m0 = b * kModeBlock;
        const int64_t cnt = (a.n_modes - m0) < kModeBlock ? (a.n_modes - m0) : kModeBlock;
        const uint32_t bytes = (uint32_t)(cnt * R * sizeof(double));
        mbar_expect_tx(&s_bar[st], bytes);
        bulk_g2s(&s_rec[st][0], a.rec + m0 * R, bytes, &s_bar[st]);
    };
    if (tid == 0) {
        for (int64_t b = 0; b < kStages && b < n_blocks; ++b) issue(b);
    }

    for (int64_t b = 0; b < n_blocks; ++b) {
        const int st = (int)(b % kStages);
        const uint32_t parity = (uint32_t)((b / kStages) & 1);
        mbar_wait(&s_bar[st], parity);
        const int64_t m0 = b * kModeBlock;
        const int cnt = (int)((a.n_modes - m0) < kModeBlock ? (a.n_modes - m0) : kModeBlock);
        const double *blk = &s_rec[st][0];

        constexpr int kUnroll = gsf_unroll<D, NC, P>();
#pragma unroll kUnroll
        for (int i = sub; i < cnt; i += L) {
            const double *m = blk + i * R;
            double kh[D], nth, amp[NC];
#pragma unroll
            for (int d = 0; d < D; ++d) kh[d] = m[d];
            nth = m[D];
#pragma unroll
            for (int c = 0; c < NC; ++c) amp[c] = m[D + 1 + c];
            if (gsf_style<D, NC, P>() == 0) {
                // one point after the other (ptxas interleaves the chains itself)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    double t = nth;
#pragma unroll
                    for (int d = 0; d < D; ++d) t = fma(kh[d], x[p][d], t);
                    const double y = cospi_signed<DEG>(t, coef);
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[p][c] = fma(amp[c], y, acc[p][c]);
                }
            } else {
            // The P points advance in lockstep, one step of the recipe at a time: consecutive
            // instructions then share the mode operand (kh[d], a coefficient, amp[c]) in the same
            // source slot, which ptxas serves from the operand-reuse cache -- a DFMA with three
            // fresh register sources costs 3 cycles instead of 2 on B200.
            double t[P], tn[P], sq[P], u[P];
#pragma unroll
            for (int p = 0; p < P; ++p) t[p] = fma(kh[0], x[p][0], nth);
#pragma unroll
            for (int d = 1; d < D; ++d)
#pragma unroll
                for (int p = 0; p < P; ++p) t[p] = fma(kh[d], x[p][d], t[p]);
#pragma unroll
            for (int p = 0; p < P; ++p) tn[p] = __dadd_rn(t[p], kMagic);       // low bits = rint(t)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double nf = __dadd_rn(tn[p], -kMagic);
                const double r = __dadd_rn(t[p], -nf);                          // exact, |r| <= 1/2
                sq[p] = __dmul_rn(r, r);
            }
#pragma unroll
            for (int p = 0; p < P; ++p)
                u[p] = kPolyMonic ? __dadd_rn(sq[p], coef.c[DEG - 1]) : fma(coef.c[6], sq[p], coef.c[5]);
#pragma unroll
            for (int q = DEG - 2; q >= 0; --q)
#pragma unroll
                for (int p = 0; p < P; ++p) u[p] = fma(u[p], sq[p], coef.c[q]);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double y = kPolyMonic ? fma(u[p], u[p], coef.c[6])         // cos(pi r) / c^2
                                            : fma(u[p], u[p], -1.0);             // cos(pi r)
                const uint32_t hi = static_cast<uint32_t>(__double2hiint(y)) +
                                    (static_cast<uint32_t>(__double2loint(tn[p])) << 31);
                u[p] = __hiloint2double(static_cast<int>(hi), __double2loint(y));   // (-1)^n cos(pi r)
            }
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p][c] = fma(amp[c], u[p], acc[p][c]);
            }
        }
        __syncthreads();                              // everyone is done with stage st
        if (tid == 0 && b + kStages < n_blocks) issue(b + kStages);
    }

    // ---- combine the L mode slices (fixed-order butterfly), then one store per point
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            double v = acc[p][c];
#pragma unroll
            for (int o = L / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[p][c] = v;
        }
        if (sub == 0 && jpt[p] < a.n_points) {
#pragma unroll
            for (int c = 0; c < NC; ++c) a.out[c * a.os0 + jpt[p] * a.os1] = acc[p][c];
        }
    }
}

// The kernel.  Tiles [0, n_big) carry P points per thread; the remaining tiles carry ONE point per
// thread (a third of the work for P = 3).  The block scheduler hands out CTAs in index order, so
// the short tiles run last: the machine drains in steps of the short tile, which removes most of
// the end-of-kernel imbalance (17 vs 18 long CTAs per SM at C2) and partial-occupancy tail.
template <int D, int NC, int P, int L, int DEG>
__global__ void __launch_bounds__(kThreads) gsf_sum_kernel(SumArgs a)
{
    constexpr int R = rec_doubles(D, NC);
    __shared__ __align__(128) double s_rec[kStages][kModeBlock * R];
    __shared__ __align__(8) uint64_t s_bar[kStages];
    constexpr int kTile = (kThreads / L) * P;        // points per long tile
    const int64_t b = blockIdx.x;
#ifdef GSF_NO_HYBRID_TAIL
    sum_tile<D, NC, P, L, DEG>(a, b * kTile, s_rec, s_bar);
#else
    if (P <= kTailP || L > 1 || b < a.n_big) {
        sum_tile<D, NC, P, L, DEG>(a, b * kTile, s_rec, s_bar);
    } else {
        sum_tile<D, NC, kTailP, (P <= kTailP ? L : 1), DEG>(a, a.n_big * kTile + (b - a.n_big) * (kThreads * kTailP),
                                                            s_rec, s_bar);
    }
#endif
}

// Scalar / Fourier field in ANY dimension (dim > 8; the reference's ndarray `dot` takes any length,
// src/field.rs:57).  Completeness path, not a tuned one: one point per thread, the dimension is a
// run-time loop, mode records and the point's coordinates come through L1 (`__ldg`; a warp reads one
// record as a broadcast and 32 consecutive points per coordinate row).  Same recipe, same record
// layout and the same mode order as gsf_sum_kernel<D,1,1,1>.
template <int DEG>
__global__ void __launch_bounds__(kThreads) gsf_sum_kernel_anyd(SumArgs a)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= a.n_points) return;
    const int D = a.dim;
    const int R = rec_doubles(D, 1);
    const PolyCoef coef = {{a.coef[0], a.coef[1], a.coef[2], a.coef[3], a.coef[4], a.coef[5], a.coef[6]}};
    const double *x = a.pos + j * a.ps1;
    double acc = a.offset[0];
    for (int64_t i = 0; i < a.n_modes; ++i) {
        const double *m = a.rec + i * R;
        double t = __ldg(m + D);
        for (int d = 0; d < D; ++d) t = fma(__ldg(m + d), __ldg(x + d * a.ps0), t);
        acc = fma(__ldg(m + D + 1), cospi_signed<DEG>(t, coef), acc);
    }
    a.out[j * a.os1] = acc;
}

// ------------------------------------------------------------------------------------------
// One-launch path for small problems (C1: 100 modes x 1e4 points, where a call is launch latency,
// not arithmetic).  The RAW modes travel inside the kernel parameters (rows k[0..D-1], z1, z2, sf of
// n_modes doubles each -- no upload, no pre-pass launch); every CTA builds the <= kModeBlock records
// in shared memory with prep_record and then runs the same recipe as gsf_sum_kernel<D,NC,1,L,kHiDeg>
// (one point per thread, `lanes` lanes per point, same mode order and butterfly => bit-identical
// results).  CAP = doubles of raw modes that fit: 480 keeps the launch inside the classic 4 KB
// parameter block, 1536 uses the large-parameter launch (CUDA 12.1+).
template <int CAP>
struct SmallArgs {
    SumArgs a;            // rec unused
    double scale;
    int lanes;            // L: power of two <= 32
    int has_sf;
    double raw[CAP];
};

template <int D, int NC, int CAP>
__global__ void __launch_bounds__(kThreads) gsf_small_kernel(const __grid_constant__ SmallArgs<CAP> s)
{
    constexpr int R = rec_doubles(D, NC);
    __shared__ __align__(16) double s_rec[kModeBlock * R];
    const int tid = threadIdx.x;
    const int N = (int)s.a.n_modes;
    for (int i = tid; i < N; i += kThreads)
        prep_record(D, NC > 1, s.scale, s.raw[D * N + i], s.raw[(D + 1) * N + i], s.has_sf != 0,
                    s.has_sf ? s.raw[(D + 2) * N + i] : 1.0, [&](int d) { return s.raw[d * N + i]; }, s_rec + i * R);
    const int L = s.lanes;
    const int sub = tid & (L - 1);
    const int grp = tid / L;
    const int64_t j = (int64_t)blockIdx.x * (kThreads / L) + grp;
    const int64_t jc = j < s.a.n_points ? j : s.a.n_points - 1;
    double x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = __ldg(s.a.pos + d * s.a.ps0 + jc * s.a.ps1);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = sub == 0 ? s.a.offset[c < 3 ? c : 0] : 0.0;
    const PolyCoef coef = {{s.a.coef[0], s.a.coef[1], s.a.coef[2], s.a.coef[3], s.a.coef[4], s.a.coef[5], s.a.coef[6]}};
    __syncthreads();
    for (int i = sub; i < N; i += L) {
        const double *m = s_rec + i * R;
        double t = m[D];
#pragma unroll
        for (int d = 0; d < D; ++d) t = fma(m[d], x[d], t);
        const double y = cospi_signed<kHiDeg>(t, coef);
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = fma(m[D + 1 + c], y, acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        double v = acc[c];
        for (int o = L >> 1; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[c] = v;
    }
    if (sub == 0 && j < s.a.n_points) {
#pragma unroll
        for (int c = 0; c < NC; ++c) s.a.out[c * s.a.os0 + j * s.a.os1] = acc[c];
    }
}

// ------------------------------------------------------------------------------------------
// DFMA peak micro-benchmark: kPeakChains independent dependent-FMA chains per thread, all SMs at
// full occupancy; each loop trip issues kPeakChains*kPeakUnroll DFMAs and nothing else of note.
// The chain is v = fma(v, a, v): only TWO distinct register sources.  A DFMA with three distinct
// register sources is register-file-read limited to one per 3 cycles on B200
// (tools/micro/dfma_patterns.cu: fma(v,a,b) 17.1 T/s, fma(v,a,v) 18.57 T/s = 64 FMA/clk/SM), so
// this pattern is the highest DFMA rate the part sustains -- the honest roofline denominator.
constexpr int kPeakChains = 8;
constexpr int kPeakUnroll = 32;
__global__ void __launch_bounds__(256) gsf_dfma_peak_kernel(double *sink, int iters, double a, double b)
{
    double v[kPeakChains];
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) v[c] = (double)(threadIdx.x + c) * 1e-3 + b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kPeakUnroll; ++u)
#pragma unroll
            for (int c = 0; c < kPeakChains; ++c) v[c] = fma(v[c], a, v[c]);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kPeakChains; ++c) s += v[c];
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

}  // namespace gsf
