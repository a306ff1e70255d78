/*
 * c_driver.c -- a plain C program against include/bspatom.h and libbspatom.so only (no Python, no torch): what the
 * reference-side binding of INTEGRATION.md does from Fortran, written in C.  Hydrogen in a box on the shipped input's
 * linear grid (KIND_GRID = 0 of grid.f90:21-27: k-fold end knots, equidistant breakpoints), l = 0 and 1 through
 * bspatom_solve_batch, checked against E = -1/(2 n^2).
 * exit 0: ok;  77: no CUDA device (bspatom_create -> BSPATOM_ENODEVICE: the path has no CPU fallback);  1: failure.
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bspatom.h"

int main(void)
{
    /* layout of struct bsp_problem as this C compiler sees it: the Python (ctypes) and Fortran (BIND(C)) mirrors must agree */
    printf("c_driver: sizeof(bsp_problem) = %zu, offsets rt %zu pot_par %zu v_tab %zu l %zu ul_extra %zu nvec %zu sel_mode %zu sel_ecut_a %zu sel_ecut_b %zu\n",
           sizeof(bsp_problem), offsetof(bsp_problem, rt), offsetof(bsp_problem, pot_par), offsetof(bsp_problem, v_tab),
           offsetof(bsp_problem, l), offsetof(bsp_problem, ul_extra), offsetof(bsp_problem, nvec), offsetof(bsp_problem, sel_mode),
           offsetof(bsp_problem, sel_ecut_a), offsetof(bsp_problem, sel_ecut_b));
    bspatom_handle h = NULL;
    int rc = bspatom_create(&h, 0);
    if (rc == BSPATOM_ENODEVICE) { printf("c_driver: no CUDA device (rc=%d), nothing computed\n", rc); return 77; }
    if (rc) { printf("c_driver: bspatom_create rc=%d\n", rc); return 1; }
    enum { K = 7, NFUN = 120, NKP = NFUN + K };
    const double rb = 60.0;
    /* nfun functions after dropping the first and last B-spline (boundary conditions): nfun + 2 splines,
     * nfun + 2 - k + 1 intervals; the library gets the knots rt(1:nkp) of the retained functions exactly like
     * READ_INPUTS / GRID hand them over: rt(1:k-1) = ra, interior breakpoints, rt(nkp-k+2:nkp) = rb */
    double rt[NKP];
    const int nint = NFUN + 2 - K + 1;
    int i, m = 0;
    for (i = 0; i < K - 1; ++i) rt[m++] = 0.0;
    for (i = 1; i < nint; ++i) rt[m++] = rb * (double)i / (double)nint;
    for (i = 0; i < K - 1; ++i) rt[m++] = rb;
    if (m != NKP) { printf("c_driver: knot count %d != %d\n", m, NKP); return 1; }
    bsp_problem p[2];
    memset(p, 0, sizeof p);
    for (i = 0; i < 2; ++i) {
        p[i].k = K; p[i].nfun = NFUN; p[i].nkp = NKP; p[i].ka = K + 3; p[i].rt = rt;
        p[i].pot_kind = BSPATOM_POT_COULOMB; p[i].pot_par[0] = 1.0;
        p[i].l = i; p[i].nvec = 4;
    }
    double *E = (double *)bspatom_alloc_host(sizeof(double) * 2 * NFUN);
    double *C = (double *)bspatom_alloc_host(sizeof(double) * 2 * NFUN * 4);
    int info[2] = {-1, -1};
    rc = bspatom_solve_batch(h, 2, p, E, C, info);
    if (rc) { printf("c_driver: bspatom_solve_batch rc=%d: %s\n", rc, bspatom_last_error(h)); return 1; }
    int bad = (info[0] != 0) || (info[1] != 0);
    /* l = 0: n = 1, 2, 3;  l = 1: n = 2, 3 */
    const double want[5] = {-0.5, -0.125, -1.0 / 18.0, -0.125, -1.0 / 18.0};
    const double got[5] = {E[0], E[1], E[2], E[NFUN], E[NFUN + 1]};
    for (i = 0; i < 5; ++i) {
        printf("c_driver: E = %.12f (exact %.12f)\n", got[i], want[i]);
        if (fabs(got[i] - want[i]) > 1e-6) bad = 1;
    }
    double stats[24];
    bspatom_get_stats(h, stats, 24);
    printf("c_driver: info = %d %d, kernel launches = %.0f, version %d\n", info[0], info[1], stats[0], bspatom_version());
    bspatom_free_host(E); bspatom_free_host(C);
    bspatom_destroy(h);
    printf(bad ? "c_driver: FAILED\n" : "c_driver: ok\n");
    return bad;
}
