/*
 * field_oracle.c -- CPU restatement of GSTools-Core's randomization-method field summation.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the timed CPU baseline.  Nothing
 * under gstools-core_b200/ (the product) may include, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) do.
 *
 * Parity status: PINNED.  The reference's Rust crate cannot be built in this image (no cargo/rustc),
 * so the oracle restates the algorithm and is pinned against the reference's own known-answer
 * tests: src/field.rs:359-382 (summator, bitwise), :334-357 (summator_fourier, bitwise),
 * :384-430 (summator_incompr, <= 6 ulp or <= f64::EPSILON abs).  See tests/test_oracle_golden.py.
 *
 * Arithmetic contract (what makes the bitwise match possible):
 *   - ndarray 0.15.6 1-D `dot` on short strided views (src/field.rs:57,242) and ShortVec::dot
 *     (src/short_vec.rs:31-33) are a left-to-right `sum = sum + a*b` starting from 0.0;
 *   - `Zip::fold` (src/field.rs:54-62, 237-246) visits modes in index order;
 *   - `f64::sin` / `f64::cos` are two separate libm calls (glibc here, as on the reference's
 *     Linux wheels);
 *   - no FMA contraction: build with  gcc -O2 -ffp-contract=off  and no -march=native.
 *
 * All arrays are addressed as base[i*stride0 + j*stride1] with ELEMENT strides, mirroring the
 * arbitrary-stride ndarray views the reference accepts (src/lib.rs:43-46).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GSO_OK 0
#define GSO_ERR_DIM 1
#define GSO_ERR_EMPTY 2
#define GSO_ERR_ALLOC 3

/* phase = <k_i, x_j>, sequential from 0.0 (ndarray dot / ShortVec::dot, src/short_vec.rs:31-33) */
static inline double dot_seq(int d, const double *k, int64_t ks0, const double *x, int64_t xs0)
{
    double s = 0.0;
    for (int a = 0; a < d; ++a)
        s = s + k[a * ks0] * x[a * xs0];
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* field::summator, src/field.rs:37-65.  out[j] = fold_i (sum + (z1_i*cos(ph) + z2_i*sin(ph))) */
static void summator_range(int d, int64_t N, const double *k, int64_t ks0, int64_t ks1,
                           const double *z1, int64_t z1s, const double *z2, int64_t z2s,
                           const double *pos, int64_t ps0, int64_t ps1, double *out,
                           int64_t j0, int64_t j1)
{
    for (int64_t j = j0; j < j1; ++j) {
        const double *x = pos + j * ps1;
        double sum = 0.0; /* fold identity, src/field.rs:55 */
        for (int64_t i = 0; i < N; ++i) {
            double phase = dot_seq(d, k + i * ks1, ks0, x, ps0);   /* :57 */
            double z12 = z1[i * z1s] * cos(phase) + z2[i * z2s] * sin(phase); /* :58 */
            sum = sum + z12;                                      /* :60 */
        }
        out[j] = sum;
    }
}

int gso_summator(int d, int64_t N, int64_t M, const double *k, int64_t ks0, int64_t ks1,
                 const double *z1, int64_t z1s, const double *z2, int64_t z2s,
                 const double *pos, int64_t ps0, int64_t ps1, double *out, int num_threads)
{
    if (d < 1) return GSO_ERR_DIM;
    (void)num_threads;
#ifdef _OPENMP
    if (num_threads > 1) {
        /* parallel over points like Zip::par_map_collect (src/field.rs:53); per-point mode order
         * is unchanged, so the result is identical for every thread count. */
        const int64_t blk = 256;
        const int64_t nblk = (M + blk - 1) / blk;
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
        for (int64_t b = 0; b < nblk; ++b) {
            int64_t j0 = b * blk, j1 = j0 + blk > M ? M : j0 + blk;
            summator_range(d, N, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, j0, j1);
        }
        return GSO_OK;
    }
#endif
    summator_range(d, N, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, 0, M);
    return GSO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* field::summator_fourier, src/field.rs:219-249.  z12 = sf_i * (z1_i*cos + z2_i*sin)  (:243)  */
static void fourier_range(int d, int64_t N, const double *sf, int64_t sfs, const double *k,
                          int64_t ks0, int64_t ks1, const double *z1, int64_t z1s,
                          const double *z2, int64_t z2s, const double *pos, int64_t ps0,
                          int64_t ps1, double *out, int64_t j0, int64_t j1)
{
    for (int64_t j = j0; j < j1; ++j) {
        const double *x = pos + j * ps1;
        double sum = 0.0; /* :241 */
        for (int64_t i = 0; i < N; ++i) {
            double phase = dot_seq(d, k + i * ks1, ks0, x, ps0); /* :242 */
            double z12 = sf[i * sfs] * (z1[i * z1s] * cos(phase) + z2[i * z2s] * sin(phase));
            sum = sum + z12; /* :245 */
        }
        out[j] = sum;
    }
}

int gso_summator_fourier(int d, int64_t N, int64_t M, const double *sf, int64_t sfs,
                         const double *k, int64_t ks0, int64_t ks1, const double *z1, int64_t z1s,
                         const double *z2, int64_t z2s, const double *pos, int64_t ps0,
                         int64_t ps1, double *out, int num_threads)
{
    if (d < 1) return GSO_ERR_DIM;
    (void)num_threads;
#ifdef _OPENMP
    if (num_threads > 1) {
        const int64_t blk = 256;
        const int64_t nblk = (M + blk - 1) / blk;
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
        for (int64_t b = 0; b < nblk; ++b) {
            int64_t j0 = b * blk, j1 = j0 + blk > M ? M : j0 + blk;
            fourier_range(d, N, sf, sfs, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, j0, j1);
        }
        return GSO_OK;
    }
#endif
    fourier_range(d, N, sf, sfs, k, ks0, ks1, z1, z1s, z2, z2s, pos, ps0, ps1, out, 0, M);
    return GSO_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* field::summator_incompr, src/field.rs:97-182.
 *
 * One mode applied to every point of an accumulator (the body of the rayon fold, :137-156).
 * acc is M records of d doubles (the Vec<ShortVec<N>> of :136; we drop ShortVec<3>'s padding
 * lane, which never takes part in the arithmetic). */
static void incompr_apply_mode(int d, const double *kv /* d contiguous */, double z1, double z2,
                               int64_t M, const double *pos, int64_t ps0, int64_t ps1,
                               double *acc)
{
    double kk = 0.0; /* cov_samples.dot(cov_samples), ShortVec::dot sequential */
    for (int a = 0; a < d; ++a)
        kk = kk + kv[a] * kv[a];
    const double k_2 = kv[0] / kk; /* :138  (NaN for a zero mode, as in the reference) */
    for (int64_t j = 0; j < M; ++j) {
        const double *x = pos + j * ps1;
        double phase = 0.0;
        for (int a = 0; a < d; ++a)
            phase = phase + kv[a] * x[a * ps0]; /* :145 */
        double z12 = z1 * cos(phase) + z2 * sin(phase); /* :146 */
        double *s = acc + j * d;
        s[0] += (1.0 - kv[0] * k_2) * z12; /* :148 */
        for (int a = 1; a < d; ++a)
            s[a] -= kv[a] * k_2 * z12; /* :151  ((k_a*k_2)*z12, left to right) */
    }
}

/* out is M records of d doubles == the reference's (d, M) Array2 in its F-ordered memory
 * (Array2::from_shape_vec((M, N), ..).reversed_axes(), src/field.rs:166-174).
 *
 * num_threads <= 1: one accumulator, modes applied in index order (what rayon does when the fold
 * is not split).  num_threads > 1: the reference's loop nest -- modes split into contiguous
 * ranges of >= 100 (with_min_len(100), :134), one M-long accumulator per range (:136), ranges
 * added left to right (:158-162).  Like the reference this changes the last ulps with the
 * thread count, which is why its own test allows 6 ulp (:428). */
int gso_summator_incompr(int d, int64_t N, int64_t M, const double *k, int64_t ks0, int64_t ks1,
                         const double *z1, int64_t z1s, const double *z2, int64_t z2s,
                         const double *pos, int64_t ps0, int64_t ps1, double *out,
                         int num_threads)
{
    if (d != 2 && d != 3) return GSO_ERR_DIM; /* :177-181 */
    if (N == 0) return GSO_ERR_EMPTY;         /* reduce_with(..).unwrap() on no items, :163 */

    int64_t nsplit = 1;
#ifdef _OPENMP
    if (num_threads > 1) {
        nsplit = N / 100;
        if (nsplit > num_threads) nsplit = num_threads;
        if (nsplit < 1) nsplit = 1;
    }
#endif
    if (nsplit == 1) {
        memset(out, 0, sizeof(double) * (size_t)M * d);
        double kv[3];
        for (int64_t i = 0; i < N; ++i) {
            for (int a = 0; a < d; ++a) kv[a] = k[a * ks0 + i * ks1]; /* ShortVec::from_array :115-118 */
            incompr_apply_mode(d, kv, z1[i * z1s], z2[i * z2s], M, pos, ps0, ps1, out);
        }
        return GSO_OK;
    }
#ifdef _OPENMP
    double *accs = (double *)calloc((size_t)nsplit * (size_t)M * d, sizeof(double));
    if (!accs) return GSO_ERR_ALLOC;
    /* The reference nests a second level of parallelism over the points inside every mode
     * (`pos.par_iter().zip(&mut summed_modes)`, :142), so all threads stay busy even when N/100 is
     * smaller than the pool.  Here: every mode range is also cut into point blocks, one task per
     * (range, block); inside a task the loop nest stays mode-outer / point-inner like the
     * reference's, and every point still receives its range's modes in index order -- the values
     * do not depend on the blocking. */
    int64_t pblocks = (num_threads + nsplit - 1) / nsplit;
    if (pblocks > M / 1024) pblocks = M / 1024;
    if (pblocks < 1) pblocks = 1;
#pragma omp parallel for collapse(2) schedule(static, 1) num_threads(num_threads)
    for (int64_t s = 0; s < nsplit; ++s) {
        for (int64_t b = 0; b < pblocks; ++b) {
            int64_t i0 = N * s / nsplit, i1 = N * (s + 1) / nsplit;
            int64_t j0 = M * b / pblocks, j1 = M * (b + 1) / pblocks;
            double *acc = accs + (size_t)s * M * d + (size_t)j0 * d;
            double kv[3];
            for (int64_t i = i0; i < i1; ++i) {
                for (int a = 0; a < d; ++a) kv[a] = k[a * ks0 + i * ks1];
                incompr_apply_mode(d, kv, z1[i * z1s], z2[i * z2s], j1 - j0, pos + j0 * ps1, ps0, ps1, acc);
            }
        }
    }
    memcpy(out, accs, sizeof(double) * (size_t)M * d);
    for (int64_t s = 1; s < nsplit; ++s) {
        const double *acc = accs + (size_t)s * M * d;
#pragma omp parallel for num_threads(num_threads)
        for (int64_t e = 0; e < M * d; ++e)
            out[e] += acc[e]; /* ShortVec::add, src/short_vec.rs:36-40 */
    }
    free(accs);
#endif
    return GSO_OK;
}

int gso_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* krige::calculator_field_krige[_and_variance], src/krige.rs:24-118  (SURVEY.md section 8 f4).
 *
 * For every target point p (column of krig_vecs):
 *     krig_fac_i = <column i of krig_mat, column p of krig_vecs>   (:54,:111; ndarray 1-D dot on
 *                  strided column views = sequential fold from 0.0, as for field.rs:57)
 *     field[p] = sum_i cond_i * krig_fac_i                         (:55,:112)
 *     error[p] = sum_i vecs[i,p] * krig_fac_i                      (:56)
 * The reference folds over i with rayon (`into_par_iter().fold().reduce()`), so its summation
 * order over i depends on the split; its tests allow 6 ulp (:203-226).  Here: index order.
 * error == NULL computes the field only. */
int gso_krige(int64_t C, int64_t M, const double *mat, int64_t ms0, int64_t ms1, const double *vecs,
              int64_t vs0, int64_t vs1, const double *cond, int64_t cs, double *field, double *error,
              int num_threads)
{
    if (C < 0 || M < 0) return GSO_ERR_DIM;
    (void)num_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(num_threads > 1 ? num_threads : 1)
#endif
    for (int64_t p = 0; p < M; ++p) {
        double f = 0.0, e = 0.0;
        for (int64_t i = 0; i < C; ++i) {
            double fac = 0.0;
            for (int64_t j = 0; j < C; ++j)
                fac = fac + mat[j * ms0 + i * ms1] * vecs[j * vs0 + p * vs1];
            f += cond[i * cs] * fac;
            e += vecs[i * vs0 + p * vs1] * fac;
        }
        field[p] = f;
        if (error) error[p] = e;
    }
    return GSO_OK;
}
