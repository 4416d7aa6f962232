/* Plain-C client of the C ABI (include/gsfield.h): what the reference's Rust FFI would do.
 * Runs the reference's known-answer vectors (src/field.rs:258-431) through gsf_summate,
 * gsf_summate_incompr, gsf_summate_fourier and gsf_summate_ex; exit code 0 on success.
 *   gcc -std=c11 -I include tests/c_abi/kat_test.c -L gstools-core_b200/gstools_core -lgsfield -lm */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "gsfield.h"

static const double K[3][10] = {
    {-2.15, 1.04, 0.69, -1.09, -1.54, -2.32, -1.81, -2.78, 1.57, -3.44},
    {0.19, -1.24, -2.10, -2.86, -0.63, -0.51, -1.68, -0.07, 0.29, -0.007},
    {0.98, -2.83, -0.10, 3.23, 0.51, 0.13, -1.03, 1.53, -0.51, 2.82}};
static const double Z1[10] = {-1.93, 0.46, 0.66, 0.02, -0.10, 1.29, 0.93, -1.14, 1.81, 1.47};
static const double Z2[10] = {-0.26, 0.98, -1.30, 0.66, 0.57, -0.25, -0.31, -0.29, 0.69, 1.14};
static const double P[3][8] = {{0.00, 1.43, 2.86, 4.29, 5.71, 7.14, 9.57, 10.00},
                               {-5.00, -3.57, -2.14, -0.71, 0.71, 2.14, 3.57, 5.00},
                               {-6.00, -4.00, -2.00, 0.00, 2.00, 4.00, 6.00, 8.00}};
static const double SUM[8] = {0.3773130601113641, -4.298994445846448, 0.9285578931297425, 0.893013192171638,
                              -1.4956409956178418, -1.488542499264307, 0.19211668257573278, 2.3427520079106143};
static const double FOU[8] = {1.0666558330143816, -3.5855143411414883, -2.70208228699285, 9.808554698975039,
                              0.01634921830347258, -2.2356422006860663, 14.730786907708966, -2.851408419726332};
static const double INC[3][8] = {
    {0.7026540940472319, -1.9323916721330978, -0.4166102970790725, 0.27803989953742114, -2.0809691290114567,
     0.20148641078244162, 0.7758364517737109, 0.12811415623445488},
    {0.3498241912898348, -0.07775049450238455, -0.5970579726508763, 0.03011066817308309, -0.6406632397415202,
     0.4669548537557405, 0.908893008714896, -0.5120295866263118},
    {0.2838955719581232, -0.9042103150526011, -0.6494289973178196, -0.5654019280252776, -0.8386683161758316,
     -0.4648269322196026, -0.0656185245433833, 1.6593799470196355}};

static int check(const char *what, const double *got, const double *want, int n, double tol)
{
    double worst = 0.0;
    for (int i = 0; i < n; ++i) worst = fmax(worst, fabs(got[i] - want[i]));
    printf("%-28s max|d| = %.3g %s\n", what, worst, worst <= tol ? "ok" : "FAIL");
    return worst <= tol ? 0 : 1;
}

int main(void)
{
    int bad = 0, rc;
    double out[8], outi[24], want[24];
    if (gsf_abi_version() != GSF_ABI_VERSION) { printf("ABI version mismatch\n"); return 2; }
    if (gsf_device_count() < 1) { printf("no CUDA device: %d\n", gsf_summate(3, 10, 8, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, out, 0)); return 3; }

    rc = gsf_summate(3, 10, 8, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, out, 0);
    if (rc) { printf("gsf_summate: %s\n", gsf_last_error()); return 1; }
    bad += check("gsf_summate", out, SUM, 8, 1e-13);

    rc = gsf_summate_fourier(3, 10, 8, &K[0][0], 1, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, out, 0);
    if (rc) { printf("gsf_summate_fourier: %s\n", gsf_last_error()); return 1; }
    bad += check("gsf_summate_fourier", out, FOU, 8, 1e-13);

    /* (3, 8) result in Fortran order, as the reference returns it: out[a + 3*j] */
    rc = gsf_summate_incompr(3, 10, 8, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, outi, 1, 3, 0);
    if (rc) { printf("gsf_summate_incompr: %s\n", gsf_last_error()); return 1; }
    for (int a = 0; a < 3; ++a)
        for (int j = 0; j < 8; ++j) want[a + 3 * j] = INC[a][j];
    bad += check("gsf_summate_incompr (F)", outi, want, 24, 1e-13);
    /* and in C order: out[a*8 + j] */
    rc = gsf_summate_incompr(3, 10, 8, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, outi, 8, 1, 0);
    if (rc) { printf("gsf_summate_incompr: %s\n", gsf_last_error()); return 1; }
    bad += check("gsf_summate_incompr (C)", outi, &INC[0][0], 24, 1e-13);

    /* extended request: 2*sum + 1 */
    gsf_request r;
    memset(&r, 0, sizeof r);
    r.struct_size = (int32_t)sizeof r; r.kind = 0; r.dim = 3; r.n_modes = 10; r.n_points = 8;
    r.modes = &K[0][0]; r.modes_s0 = 10; r.modes_s1 = 1; r.z1 = Z1; r.z1_s = 1; r.z2 = Z2; r.z2_s = 1;
    r.pos = &P[0][0]; r.pos_s0 = 8; r.pos_s1 = 1; r.out = out; r.out_s1 = 1; r.scale = 2.0; r.offset[0] = 1.0;
    rc = gsf_summate_ex(&r);
    if (rc) { printf("gsf_summate_ex: %s\n", gsf_last_error()); return 1; }
    for (int j = 0; j < 8; ++j) want[j] = 2.0 * SUM[j] + 1.0;
    bad += check("gsf_summate_ex scale/offset", out, want, 8, 1e-13);

    /* error convention: status codes, never abort */
    rc = gsf_summate_incompr(1, 10, 8, &K[0][0], 10, 1, Z1, 1, Z2, 1, &P[0][0], 8, 1, outi, 1, 1, 0);
    printf("%-28s rc = %d (%s) %s\n", "incompr dim=1", rc, gsf_last_error(), rc == GSF_ERR_DIM ? "ok" : "FAIL");
    bad += rc != GSF_ERR_DIM;

    gsf_stats st;
    gsf_get_last_stats(&st);
    gsf_shutdown();
    printf(bad ? "FAILED\n" : "C ABI KAT OK\n");
    return bad ? 1 : 0;
}
