/*
 * variogram_oracle.c -- CPU restatement of GSTools-Core's empirical variogram estimators.
 *
 * TEST INFRASTRUCTURE ONLY (see field_oracle.c): the parity oracle for the variogram kernels of
 * gstools-core_b200/csrc/gsf_variogram_kernels.cuh.  Never linked into the product.
 *
 * Parity status: PINNED against the reference's own known-answer tests, src/variogram.rs:577-842
 * (transcribed in tests/golden/variogram_rs_kat.json, checked by tests/test_oracle_golden.py).
 *
 * The loop nests are the reference's: one pass over all point pairs PER BIN (the reference
 * parallelises over bins, src/variogram.rs:382-386 / 507-511), pairs visited i ascending, then
 * j = i+1.. ascending, every sum a sequential `acc += term`.  No FMA contraction (Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define GSO_OK 0
#define GSO_ERR_DIM 1
#define GSO_ERR_EMPTY 2

/* Estimator, src/variogram.rs:41-65.  est 'c' = Cressie, anything else = Matheron (:24-38). */
static inline double estimate(char est, double f_diff)
{
    if (est == 'c') return sqrt(fabs(f_diff));   /* :57-59 */
    return f_diff * f_diff;                      /* powi(2), :44-46 */
}

static inline double normalize(char est, double v, uint64_t c)
{
    const double cf = c == 0 ? 1.0 : (double)c;  /* :49, :62 */
    if (est == 'c') {
        const double m = 1. / cf * v;            /* :63: (1. / cf * *v).powi(4) */
        const double m2 = m * m;
        return 0.5 * (m2 * m2) / (0.457 + 0.494 / cf + 0.045 / (cf * cf));
    }
    return v / (2.0 * cf);                       /* :50 */
}

/* Euclid::dist, src/variogram.rs:92-103: acc += (l - r)^2 from 0.0, then sqrt */
static inline double dist_euclid(int d, const double *pos, int64_t ps0, int64_t ps1, int64_t i, int64_t j)
{
    double acc = 0.0;
    for (int a = 0; a < d; ++a) {
        const double t = pos[a * ps0 + i * ps1] - pos[a * ps0 + j * ps1];
        acc += t * t;
    }
    return sqrt(acc);
}

/* f64::to_radians = self * (PI / 180) */
static inline double to_radians(double x) { return x * (3.14159265358979323846264338327950288 / 180.0); }

/* Haversine::dist, src/variogram.rs:107-118 */
static inline double dist_haversine(const double *pos, int64_t ps0, int64_t ps1, int64_t i, int64_t j)
{
    const double lat_i = pos[i * ps1], lat_j = pos[j * ps1];
    const double lon_i = pos[ps0 + i * ps1], lon_j = pos[ps0 + j * ps1];
    const double diff_lat = to_radians(lat_i - lat_j);
    const double diff_lon = to_radians(lon_i - lon_j);
    const double s1 = sin(diff_lat / 2.0), s2 = sin(diff_lon / 2.0);
    const double arg = s1 * s1 + cos(to_radians(lat_i)) * cos(to_radians(lat_j)) * (s2 * s2);
    return 2.0 * atan2(sqrt(arg), sqrt(1.0 - arg));
}

/* variogram_structured / variogram_ma_structured, src/variogram.rs:136-240.
 * f is (n0, n1); mask (bytes, non-zero = masked) may be NULL; out has max(n0, 1) entries. */
int gso_variogram_structured(int64_t n0, int64_t n1, const double *f, int64_t fs0, int64_t fs1,
                             const uint8_t *mask, int64_t ms0, int64_t ms1, char est, double *out,
                             int num_threads)
{
    out[0] = 0.0;   /* variogram.push(0.0), :146 / :207 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads > 1 ? num_threads : 1)
#endif
    for (int64_t k = 1; k < n0; ++k) {
        double value = 0.0;
        uint64_t count = 0;
        for (int64_t i = 0; i < n0 - k; ++i)
            for (int64_t c = 0; c < n1; ++c) {
                if (mask && (mask[i * ms0 + c * ms1] || mask[(i + k) * ms0 + c * ms1])) continue;   /* :223-225 */
                value += estimate(est, f[i * fs0 + c * fs1] - f[(i + k) * fs0 + c * fs1]);
                count += 1;
            }
        out[k] = normalize(est, value, count);
    }
    (void)num_threads;
    return GSO_OK;
}

/* variogram_unstructured, src/variogram.rs:465-545.  f (nf, M), pos (d, M), bin_edges (nb + 1).
 * dist_type 'e' = Euclid, anything else = Haversine (:74-87). */
int gso_variogram_unstructured(int d, int64_t nf, int64_t M, int64_t nb, const double *f, int64_t fs0,
                               int64_t fs1, const double *edges, int64_t es, const double *pos,
                               int64_t ps0, int64_t ps1, char est, char dist_type, double *variogram,
                               uint64_t *counts, int num_threads)
{
    if (M < 1 || nb < 1) return GSO_ERR_EMPTY;            /* :472-483 and the `- 1` at :495 */
    if (dist_type != 'e' && d != 2) return GSO_ERR_DIM;   /* Haversine::check_dim, :120-122 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads > 1 ? num_threads : 1)
#endif
    for (int64_t b = 0; b < nb; ++b) {
        const double lo = edges[b * es], hi = edges[(b + 1) * es];
        double v = 0.0;
        uint64_t c = 0;
        for (int64_t i = 0; i < M - 1; ++i)
            for (int64_t j = i + 1; j < M; ++j) {
                const double dist = dist_type == 'e' ? dist_euclid(d, pos, ps0, ps1, i, j)
                                                     : dist_haversine(pos, ps0, ps1, i, j);
                if (dist < lo || dist >= hi) continue;    /* :518-520 */
                for (int64_t q = 0; q < nf; ++q) {
                    const double f_ij = f[q * fs0 + i * fs1] - f[q * fs0 + j * fs1];
                    if (isnan(f_ij)) continue;            /* :524-526 */
                    c += 1;
                    v += estimate(est, f_ij);
                }
            }
        variogram[b] = normalize(est, v, c);
        counts[b] = c;
    }
    (void)num_threads;
    return GSO_OK;
}

/* dir_test, src/variogram.rs:243-290 */
static int dir_test(int d, const double *dir, int64_t ds1, const double *pos, int64_t ps0, int64_t ps1,
                    int64_t i, int64_t j, double dist, double angles_tol, double bandwidth)
{
    double s_prod = 0.0;
    for (int a = 0; a < d; ++a) s_prod += (pos[a * ps0 + i * ps1] - pos[a * ps0 + j * ps1]) * dir[a * ds1];
    if (bandwidth > 0.0) {
        double b = 0.0;
        for (int a = 0; a < d; ++a) {
            const double t = (pos[a * ps0 + i * ps1] - pos[a * ps0 + j * ps1]) - s_prod * dir[a * ds1];
            b += t * t;
        }
        if (sqrt(b) >= bandwidth) return 0;
    }
    if (dist > 0.0) {
        const double angle = fabs(s_prod) / dist;
        if (angle < 1.0 && acos(angle) >= angles_tol) return 0;
    }
    return 1;
}

/* variogram_directional, src/variogram.rs:315-447.  direction (nd, d); outputs (nd, nb) row-major. */
int gso_variogram_directional(int d, int64_t nf, int64_t M, int64_t nb, int64_t nd, const double *f,
                              int64_t fs0, int64_t fs1, const double *edges, int64_t es, const double *pos,
                              int64_t ps0, int64_t ps1, const double *direction, int64_t ds0, int64_t ds1,
                              double angles_tol, double bandwidth, int separate_dirs, char est,
                              double *variogram, uint64_t *counts, int num_threads)
{
    if (M < 1 || nb < 1) return GSO_ERR_EMPTY;
    if (!(angles_tol > 0.0)) return GSO_ERR_DIM;          /* :343-346 */
    for (int64_t q = 0; q < nd * nb; ++q) {
        variogram[q] = 0.0;
        counts[q] = 0;
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads > 1 ? num_threads : 1)
#endif
    for (int64_t b = 0; b < nb; ++b) {
        const double lo = edges[b * es], hi = edges[(b + 1) * es];
        for (int64_t i = 0; i < M - 1; ++i)
            for (int64_t j = i + 1; j < M; ++j) {
                const double dist = dist_euclid(d, pos, ps0, ps1, i, j);   /* :396 */
                if (dist < lo || dist >= hi) continue;
                for (int64_t r = 0; r < nd; ++r) {
                    if (!dir_test(d, direction + r * ds0, ds1, pos, ps0, ps1, i, j, dist, angles_tol, bandwidth))
                        continue;
                    for (int64_t q = 0; q < nf; ++q) {
                        const double f_ij = f[q * fs0 + i * fs1] - f[q * fs0 + j * fs1];
                        if (isnan(f_ij)) continue;
                        counts[r * nb + b] += 1;
                        variogram[r * nb + b] += estimate(est, f_ij);
                    }
                    if (separate_dirs) break;             /* :424-426 */
                }
            }
        for (int64_t r = 0; r < nd; ++r)
            variogram[r * nb + b] = normalize(est, variogram[r * nb + b], counts[r * nb + b]);
    }
    (void)num_threads;
    return GSO_OK;
}
