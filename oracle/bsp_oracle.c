/*
 * bsp_oracle.c -- CPU restatement of BspAtom's hot path (B-spline assembly).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (bspatom_b200/, the
 * C-ABI library) may link, import or call this file.  It is used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED by the reference's own tests: the reference ships no tests,
 * golden vectors or sample outputs.  This restatement
 * is pinned instead by (i) analytic hydrogen levels, (ii) LAPACK dsygv on the
 * identical matrices, (iii) a 40-digit mpmath spectrum of the shipped input
 * (tests/golden/), (iv) structural invariants.  No Fortran compiler exists in
 * the build image, so the reference itself cannot be compiled (oracle/_ref is
 * therefore absent).
 *
 * Every function cites the reference lines it follows.  Evaluation and operand
 * order follow the reference; compile with -O2 -ffp-contract=off.
 * All indices in the public interface are 1-based where the reference's are
 * (left, knot indices) and arrays are column-major like the Fortran ones.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAXK 32

/* ------------------------------------------------------------------------- *
 * READ_INPUTS size derivation            (ReadInputs.f90:39-69)
 * out[0..7] = ka, nbc1, nbc2, nkp, nointv, nfun, nintv_exp, nintv_lin
 * ------------------------------------------------------------------------- */
void bsp_sizes(int kind_grid, int k, int ka_in, int nfun0, int kind_bc1,
               int kind_bc2, double ra, double rb, double rmax, int *out)
{
    int ka = ka_in;
    if (ka == 0) ka = k + 3;                       /* ReadInputs.f90:39 */
    int nbc1 = k, nbc2 = k;                        /* :42-43 */
    if (kind_bc1 == 0) nbc1 = k - 1;               /* :44 */
    if (kind_bc2 == 0) nbc2 = k - 1;               /* :45 */
    int nfun = nfun0;
    int nkp = nfun + k;                            /* :47 */
    int nointv = nkp - nbc1 - nbc2 + 1;            /* :48 */
    double gsize = rb - ra;                        /* :50 */
    int nintv_exp = 0, nintv_lin = 0;
    if (kind_grid == 2) {                          /* :52-69 */
        double dx = gsize / (double)nointv;
        double rimax = (rmax - ra) / dx;
        int imax = (int)lround(rimax);             /* NINT: half away from 0 */
        nintv_exp = 3 * imax;
        nintv_lin = nointv - imax;
        nointv = nintv_exp + nintv_lin;
        nkp = nointv + nbc1 + nbc2 - 1;
        nfun = nkp - k;
    }
    out[0] = ka; out[1] = nbc1; out[2] = nbc2; out[3] = nkp;
    out[4] = nointv; out[5] = nfun; out[6] = nintv_exp; out[7] = nintv_lin;
}

/* ------------------------------------------------------------------------- *
 * gauleg (module version, 5 args)        (Modules.f90:112-153)
 * ------------------------------------------------------------------------- */
void bsp_gauleg(double x1, double x2, double *x, double *w, int n)
{
    const double PI = acos(-1.0);                  /* Modules.f90:9 */
    const double EPS1 = 2.220446049250313e-16 * 10;/* :129 */
    int m = (n + 1) / 2;
    double xm = 0.5 * (x2 + x1);
    double xl = 0.5 * (x2 - x1);
    /* pp is a routine-level variable in the reference: for odd n the middle
     * start value cos(pi/2) ~ 6e-17 already satisfies |z - z1| <= EPS1 with
     * z1 = 0, the Newton loop is skipped (:136) and w(middle) is formed with
     * the pp left over from the previous node (reference quirk, kept) */
    double pp = 0.0;
    for (int i = 1; i <= m; ++i) {
        double z = cos(PI * (i - .25) / (n + .5)); /* :134 */
        double z1 = 0.0;
        while (fabs(z - z1) > EPS1) {              /* :136 */
            double p1 = 1.0, p2 = 0.0, p3;
            for (int j = 1; j <= n; ++j) {
                p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            z1 = z;
            z = z1 - p1 / pp;
        }
        x[i - 1] = xm - xl * z;
        x[n - i] = xm + xl * z;
        w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp);
        w[n - i] = w[i - 1];
    }
}

/* ------------------------------------------------------------------------- *
 * GRID: knots rt(1:nkp), Aind(nfun,2)    (grid.f90:14-91)
 * rt must hold nkp+1 doubles? No: the KIND_GRID=2 loop writes index
 * nointv+nbc1 = nkp-nbc2+1 <= nkp, inside the array.
 * ------------------------------------------------------------------------- */
void bsp_grid(int kind_grid, int k, int nfun, int nkp, int nbc1, int nbc2,
              int nointv, int nintv_exp, int nintv_lin, double ra, double rb,
              double rmax, double *rt /* nkp */, double *aind /* nfun x 2 */)
{
#define RT(i) rt[(i) - 1]
    double gsize = rb - ra;
    for (int i = 1; i <= nbc1; ++i) RT(i) = ra;                 /* :16-18 */
    for (int i = nkp - nbc2 + 1; i <= nkp; ++i) RT(i) = rb;     /* :19-21 */
    if (kind_grid == 0) {                                       /* :27-29 */
        for (int i = nbc1 + 1; i <= nkp - nbc2; ++i)
            RT(i) = ra + (double)(i - nbc1) * gsize / (double)nointv;
    } else if (kind_grid == 1) {                                /* :35-42 */
        double delta = 0.01;
        double hin = log(gsize / delta) / (double)(nointv - 1);
        int j = 1;
        RT(nbc1 + 1) = delta;
        for (int i = nbc1 + 2; i <= nkp - nbc2; ++i) {
            RT(i) = RT(nbc1 + 1) * exp(hin * j);
            j = j + 1;
        }
    } else if (kind_grid == 2) {                                /* :49-61 */
        double delta = 0.01;
        double hin = log((rmax - ra) / delta) / (double)(nintv_exp - 1);
        int j = 1;
        RT(nbc1 + 1) = delta;
        for (int i = 2; i <= nintv_exp; ++i) {
            RT(i + nbc1) = delta * exp(hin * j);
            j = j + 1;
        }
        double dr = (rb - rmax) / (double)nintv_lin;
        for (int i = nintv_exp + 1; i <= nointv; ++i)
            RT(i + nbc1) = rmax + (double)(i - nintv_exp) * dr;
    }
    /* Aind                                                       :79-91 */
    for (int i = 1; i <= nfun; ++i) {
        double a1 = 0.0, a2 = 0.0;
        double dr = RT(i + k - 1) - RT(i);
        if (dr > 0.0) a1 = 1.0 / dr;
        dr = RT(i + k) - RT(i + 1);
        if (dr > 0.0) a2 = 1.0 / dr;
        aind[i - 1] = a1;              /* Aind(i,1) */
        aind[nfun + i - 1] = a2;       /* Aind(i,2) */
    }
#undef RT
}

/* ------------------------------------------------------------------------- *
 * interv                                  (interv.f90:86-116)
 * literal restatement: linear downward scan.
 * ------------------------------------------------------------------------- */
void bsp_interv(const double *xt, int lxt, double x, int *left, int *mflag)
{
#define XT(i) xt[(i) - 1]
    if (x > XT(lxt)) { *mflag = 1; *left = 1; return; }         /* :86-89 */
    else if (x < XT(1)) { *mflag = -1; *left = 1; return; }     /* :90-93 */
    else *mflag = 0;
    if (x == XT(lxt)) {                                         /* :100-105 */
        int l = lxt;
        for (;;) {
            if (XT(l) < XT(lxt)) { *left = l; return; }
            l = l - 1;
            if (l < 1) { *left = 1; return; }  /* all knots equal: the
                                                  reference would run off the
                                                  array; bounded here */
        }
    } else {                                                    /* :107-116 */
        int ilo = lxt - 1;
        *left = 0;     /* the reference leaves `left` undefined if not found */
        for (;;) {
            if (x < XT(ilo + 1) && x >= XT(ilo)) { *left = ilo; break; }
            ilo = ilo - 1;
            if (ilo == 0) break;
        }
    }
#undef XT
}

/* same result as bsp_interv for a NON-DECREASING knot vector, but the scan
 * starts at `hint` (any index >= the answer).  Used by the assembly loop to
 * avoid the O(nkp) rescan per quadrature point; tests check equality. */
static void interv_from(const double *xt, int lxt, double x, int hint,
                        int *left, int *mflag)
{
#define XT(i) xt[(i) - 1]
    if (x > XT(lxt)) { *mflag = 1; *left = 1; return; }
    else if (x < XT(1)) { *mflag = -1; *left = 1; return; }
    else *mflag = 0;
    if (x == XT(lxt)) { bsp_interv(xt, lxt, x, left, mflag); return; }
    int ilo = hint;
    if (ilo > lxt - 1) ilo = lxt - 1;
    /* make sure the hint is not below the answer */
    while (ilo < lxt - 1 && XT(ilo + 1) <= x) ++ilo;
    *left = 0;
    for (;;) {
        if (x < XT(ilo + 1) && x >= XT(ilo)) { *left = ilo; break; }
        ilo = ilo - 1;
        if (ilo == 0) break;
    }
#undef XT
}

/* ------------------------------------------------------------------------- *
 * BSPLVB (index==1 entry only, as the reference calls it)  (bsplvb.f90:10-52)
 * returns 0, or 1 for the reference's 'FATAL ERROR - BSPLVB' STOP (:30-34).
 * t has ndim entries; reads past ndim (reference quirk B-1) are treated as
 * the last knot value.
 * ------------------------------------------------------------------------- */
int bsp_bsplvb(int ndim, const double *t, int jhigh, double x, int left,
               double *biatx)
{
#define T(i) (((i) <= ndim) ? t[(i) - 1] : t[ndim - 1])
    double deltal[ORACLE_MAXK], deltar[ORACLE_MAXK];
    int j = 1;
    biatx[0] = 1.0;                                             /* :26 */
    if (jhigh <= j) return 0;                                   /* :27 */
    if (T(left + 1) <= T(left)) return 1;                       /* :30-34 */
    j = 1;
    for (;;) {                                                  /* :38-50 */
        deltar[j - 1] = T(left + j) - x;
        deltal[j - 1] = x - T(left + 1 - j);
        double saved = 0.0;
        for (int i = 1; i <= j; ++i) {
            double term = biatx[i - 1] / (deltar[i - 1] + deltal[j - i]);
            biatx[i - 1] = saved + deltar[i - 1] * term;
            saved = deltal[j - i] * term;
        }
        biatx[j] = saved;
        j = j + 1;
        if (jhigh <= j) break;
    }
    return 0;
#undef T
}

/* ------------------------------------------------------------------------- *
 * BSPALL                                  (Modules.f90:71-110)
 * hint<=0: literal interv; hint>0: interv_from(hint).
 * returns 0 / 1 (BSPLVB fatal).
 * ------------------------------------------------------------------------- */
static int bspall_impl(const double *rt, int nkp, int k, int nfun,
                       const double *aind, double r, int hint, int *left,
                       double *bsp, double *dbsp)
{
    double bsp1[ORACLE_MAXK], bspp[ORACLE_MAXK + 1];
    int mflag;
    for (int j = 0; j < k; ++j) bsp[j] = 0.0;                   /* :85 */
    for (int j = 0; j < k - 1; ++j) bsp1[j] = 0.0;              /* :86 */
    if (hint > 0) interv_from(rt, nkp, r, hint, left, &mflag);
    else bsp_interv(rt, nkp, r, left, &mflag);                  /* :87 */
    if (*left < 1) return 2;
    if (bsp_bsplvb(nkp, rt, k, r, *left, bsp)) return 1;        /* :88 */
    if (bsp_bsplvb(nkp, rt, k - 1, r, *left, bsp1)) return 1;   /* :89 */
    for (int j = 0; j <= k; ++j) bspp[j] = 0.0;                 /* :91 */
    for (int j = 1; j <= k - 1; ++j) bspp[j] = bsp1[j - 1];     /* :92-94 */
    for (int j = 1; j <= k; ++j) {                              /* :96-108 */
        int jp = j + (*left - k);
        double a1 = 0.0, a2 = 0.0;
        if (jp >= 1 && jp <= nfun) {
            a1 = aind[jp - 1];
            a2 = aind[nfun + jp - 1];
        }
        double b1 = bspp[j - 1];
        double b2 = bspp[j];
        dbsp[j - 1] = (double)(k - 1) * (a1 * b1 - a2 * b2);
    }
    return 0;
}

int bsp_bspall(const double *rt, int nkp, int k, int nfun, const double *aind,
               double r, int *left, double *bsp, double *dbsp)
{
    return bspall_impl(rt, nkp, k, nfun, aind, r, 0, left, bsp, dbsp);
}

/* ------------------------------------------------------------------------- *
 * SELPOT                                  (Modules.f90:263-295)
 * kind_pot 0 Coulomb, 1 Rogers (Ca+), 2 Simons-Fues: par = {Zatom, Ntot,
 *   Numn(1..3), alphan(1..3)}.
 * kind_pot 10 Yukawa  V = -Z exp(-lambda r)/r        par = {Z, lambda}
 * kind_pot 11 Tietz   V = -[1 + (Z-1)/(1+t r)^2]/r   par = {Z, t}
 *   (10, 11 are not in the reference: SURVEY.md 8(d) cfg3 defines them; the
 *   closest reference form is the Rogers branch.)
 * ------------------------------------------------------------------------- */
double bsp_selpot(int kind_pot, const double *par, double r)
{
    double vr = 0.0;
    if (kind_pot == 0) {
        vr = -par[0] / r;                                       /* :275 */
    } else if (kind_pot == 1) {                                 /* :279-285 */
        vr = 0.0;
        for (int i = 0; i < 3; ++i) {
            double ni = par[2 + i];
            vr = vr + ni * exp(-par[5 + i] * r);
        }
        vr = -1.0 * (par[0] - par[1] + vr) / r;
    } else if (kind_pot == 2) {
        vr = -par[0] / r;                                       /* :289 */
    } else if (kind_pot == 10) {
        vr = -par[0] * exp(-par[1] * r) / r;
    } else if (kind_pot == 11) {
        double d = 1.0 + par[1] * r;
        vr = -(1.0 + (par[0] - 1.0) / (d * d)) / r;
    }
    return vr;
}

/* Rogers alpha table                      (ReadInputs.f90:95-128) */
void bsp_rogers_params(double zatom, double *par /* 8 */)
{
    double numn[3] = {2, 8, 8};
    double aj[3][4] = {{0.8855, 0.2549, -0.0901, 0.0},
                       {0.3386, 1.1323, -0.4904, 0.0},
                       {0.1437, 0.9129, -0.6940, 0.2503}};
    int ntot = 0;
    par[0] = zatom;
    for (int i = 0; i < 3; ++i) {
        ntot = ntot + (int)numn[i];
        double xn = (double)(zatom - ntot);
        if (xn == 0.0) xn = 1.0;
        double suman = 0.0;
        for (int j = 0; j <= 3; ++j) suman = suman + aj[i][j] / pow(xn, j);
        par[5 + i] = (xn + 1.0) * suman;
        par[2 + i] = numn[i];
    }
    par[1] = (double)ntot;
}

/* ------------------------------------------------------------------------- *
 * MATRIX_SVT, scalar branch (KIND_PI <= 2)   (matrices.f90:68-183)
 *
 * Outputs (dense, column-major, N x N, caller-zeroed not required):
 *   S, V, T                         (:180-183)
 *   U(N,N,0:lmax)                   (:148-153,182)   may be NULL -> skipped
 *   R  = sumr  (length gauge rij)   (:144,160)
 *   Ri = sumc  (B_i (1/r) B_j)      (:141,162)
 *   D  = sumd  (B_i B_j')           (:142,163)   non-symmetric
 * bl: Bl(0:lmax) for KIND_POT==2, else NULL.
 * Only pairs with |ibra-jket| < k are visited; the reference visits all N^2
 * and gets exact zeros for the others (:71-72, empty ibet range).
 * fast != 0 uses the hinted interval search (non-decreasing knots only).
 * returns 0, or 1 if BSPLVB would STOP.
 * ------------------------------------------------------------------------- */
int bsp_matrix_svt(int nfun, int k, int ka, int nkp, const double *rt,
                   const double *aind, const double *xg, const double *wg,
                   int lmax, int kind_pot, const double *par, const double *bl,
                   int fast, double *S, double *V, double *T, double *U,
                   double *R, double *Ri, double *D)
{
    const double eps = 2.220446049250313e-16;                   /* Modules.f90:9 */
    const size_t nn = (size_t)nfun * (size_t)nfun;
    double bsp[ORACLE_MAXK], dbsp[ORACLE_MAXK];
    double *sumU = (double *)malloc(sizeof(double) * (size_t)(lmax + 1));
    memset(S, 0, nn * sizeof(double));
    memset(V, 0, nn * sizeof(double));
    memset(T, 0, nn * sizeof(double));
    if (U) memset(U, 0, nn * (size_t)(lmax + 1) * sizeof(double));
    if (R) memset(R, 0, nn * sizeof(double));
    if (Ri) memset(Ri, 0, nn * sizeof(double));
    if (D) memset(D, 0, nn * sizeof(double));
    int rc = 0;
    for (int ibra = 1; ibra <= nfun && !rc; ++ibra) {           /* :68 */
        int jlo = ibra - (k - 1) < 1 ? 1 : ibra - (k - 1);
        int jhi = ibra + (k - 1) > nfun ? nfun : ibra + (k - 1);
        for (int jket = jlo; jket <= jhi && !rc; ++jket) {      /* :69 */
            int ibetmin = ibra > jket ? ibra : jket;            /* :71 */
            int ibetmax = (ibra < jket ? ibra : jket) + k - 1;  /* :72 */
            double sumc = 0, sumd = 0, sumr = 0, sumS = 0, sumT = 0, sumV = 0;
            for (int lf = 0; lf <= lmax; ++lf) sumU[lf] = 0.0;
            for (int ibet = ibetmin; ibet <= ibetmax && !rc; ++ibet) { /* :89 */
                double f1 = (rt[ibet] + rt[ibet - 1]) / 2.0;    /* :91 */
                double f2 = (rt[ibet] - rt[ibet - 1]) / 2.0;    /* :92 */
                for (int igl = 0; igl < ka; ++igl) {            /* :94 */
                    double r = f1 + xg[igl] * f2;               /* :96 */
                    double dr = f2 * wg[igl];                   /* :97 */
                    int left;
                    rc = bspall_impl(rt, nkp, k, nfun, aind, r,
                                     fast ? ibet + 1 : 0, &left, bsp, dbsp);
                    if (rc) break;
                    if (r == 0.0) r = eps;                      /* :102 */
                    double vpot = bsp_selpot(kind_pot, par, r); /* :103 */
                    int ifun = ibra - (left - k);               /* :105 */
                    int jfun = jket - (left - k);               /* :106 */
                    if (ifun < 1 || ifun > k || jfun < 1 || jfun > k) {
                        rc = 3;  /* reference would index out of bounds */
                        break;
                    }
                    double fbra = bsp[ifun - 1], fket = bsp[jfun - 1];
                    double dfbra = dbsp[ifun - 1], dfket = dbsp[jfun - 1];
                    sumc = sumc + fbra * (1.0 / r) * fket * dr; /* :141 */
                    sumd = sumd + fbra * dfket * dr;            /* :142 */
                    sumr = sumr + fbra * r * fket * dr;         /* :144 */
                    sumS = sumS + fbra * fket * dr;             /* :145 */
                    sumV = sumV + fbra * vpot * fket * dr;      /* :146 */
                    sumT = sumT + dfbra * 0.5 * dfket * dr;     /* :147 */
                    if (U) {
                        for (int lf = 0; lf <= lmax; ++lf) {    /* :148-153 */
                            double vcent =
                                (double)(lf * (lf + 1)) / (2.0 * (r * r));
                            double vl = 0.0;
                            if (kind_pot == 2) vl = bl[lf] / (r * r);
                            sumU[lf] = sumU[lf] + fbra * (vcent + vl) * fket * dr;
                        }
                    }
                }
            }
            size_t ij = (size_t)(ibra - 1) + (size_t)(jket - 1) * (size_t)nfun;
            if (R) R[ij] = sumr;
            if (Ri) Ri[ij] = sumc;
            if (D) D[ij] = sumd;
            S[ij] = sumS;                                       /* :180 */
            V[ij] = sumV;                                       /* :181 */
            if (U)
                for (int lf = 0; lf <= lmax; ++lf)
                    U[ij + nn * (size_t)lf] = sumU[lf];         /* :182 */
            T[ij] = sumT;                                       /* :183 */
        }
    }
    free(sumU);
    return rc;
}

/* ------------------------------------------------------------------------- *
 * MATRIX_SVT, KIND_PI >= 3 branch        (matrices.f90:110-139, 164-175)
 * zIth : COMPLEX*16 zIth(nkp, ka, nlm, nm, ncomp) (re,im interleaved), the
 *        angular integrals on the radial quadrature grid (ZINT_TH, Ang_Ints.f90:544-600)
 * zA   : COMPLEX*16 zAij(nfun, nfun, nlm, nm, ncomp_out), dense, zero outside the band
 * Statement by statement: zsumc/zsumd (KIND_PI 3, 4: zsume/zsumf are never accumulated, so
 * components 3, 4 stay zero, :164-173) and zsumc..zsumf for KIND_PI >= 5 / >= 8.
 * Complex products are evaluated left to right, real * complex componentwise, like the
 * Fortran expression fbra * zfAr * fket * dr.
 * ------------------------------------------------------------------------- */
int bsp_matrix_zaij(int nfun, int k, int ka, int nkp, const double *rt, const double *aind, const double *xg,
                    const double *wg, int kind_pi, int nlm, int nm, int ncomp, const double *zIth, int ncomp_out,
                    double *zA)
{
    const double eps = 2.220446049250313e-16;
    double bsp[ORACLE_MAXK], dbsp[ORACLE_MAXK];
    const size_t nn = (size_t)nfun * (size_t)nfun, nblk = (size_t)nlm * (size_t)nm;
    memset(zA, 0, sizeof(double) * 2 * nn * nblk * (size_t)ncomp_out);
    double *zs = (double *)malloc(sizeof(double) * 2 * nblk * 4);   /* zsumc, zsumd, zsume, zsumf (il, jl) */
    int rc = 0;
    for (int ibra = 1; ibra <= nfun && !rc; ++ibra) {           /* :68 */
        int jlo = ibra - (k - 1) < 1 ? 1 : ibra - (k - 1);
        int jhi = ibra + (k - 1) > nfun ? nfun : ibra + (k - 1);
        for (int jket = jlo; jket <= jhi && !rc; ++jket) {      /* :69 */
            int ibetmin = ibra > jket ? ibra : jket;            /* :71 */
            int ibetmax = (ibra < jket ? ibra : jket) + k - 1;  /* :72 */
            memset(zs, 0, sizeof(double) * 2 * nblk * 4);       /* :82-87 */
            for (int ibet = ibetmin; ibet <= ibetmax && !rc; ++ibet) { /* :89 */
                double f1 = (rt[ibet] + rt[ibet - 1]) / 2.0;
                double f2 = (rt[ibet] - rt[ibet - 1]) / 2.0;
                for (int igl = 0; igl < ka; ++igl) {            /* :94 */
                    double r = f1 + xg[igl] * f2;
                    double dr = f2 * wg[igl];
                    int left;
                    rc = bspall_impl(rt, nkp, k, nfun, aind, r, ibet + 1, &left, bsp, dbsp);
                    if (rc) break;
                    if (r == 0.0) r = eps;                      /* :102 */
                    int ifun = ibra - (left - k), jfun = jket - (left - k);
                    if (ifun < 1 || ifun > k || jfun < 1 || jfun > k) { rc = 3; break; }
                    double fbra = bsp[ifun - 1], fket = bsp[jfun - 1], dfket = dbsp[jfun - 1];
                    for (int il = 0; il < nlm; ++il) {          /* :111 */
                        for (int jl = 0; jl < nm; ++jl) {       /* :112 */
                            size_t b = (size_t)il + (size_t)nlm * (size_t)jl;
#define ZITH(c, part) zIth[2 * ((size_t)(ibet - 1) + (size_t)nkp * ((size_t)igl + (size_t)ka * (b + nblk * (size_t)(c)))) + (part)]
                            if (kind_pi == 3 || kind_pi == 4) {
                                double are = ZITH(0, 0), aim = ZITH(0, 1);         /* zAlm, :118 */
                                double fre = are / r, fim = aim / r;               /* zfAr = zAlm / r, :119 */
                                zs[2 * (b + nblk * 0) + 0] += fbra * fre * fket * dr;   /* :120 */
                                zs[2 * (b + nblk * 0) + 1] += fbra * fim * fket * dr;
                                zs[2 * (b + nblk * 1) + 0] += fbra * are * dfket * dr;  /* :121 */
                                zs[2 * (b + nblk * 1) + 1] += fbra * aim * dfket * dr;
                            } else {
                                int nc = kind_pi >= 8 ? 4 : 2;                     /* :127-136 */
                                for (int c = 0; c < nc && c < ncomp; ++c) {
                                    zs[2 * (b + nblk * c) + 0] += fbra * ZITH(c, 0) * fket * dr;
                                    zs[2 * (b + nblk * c) + 1] += fbra * ZITH(c, 1) * fket * dr;
                                }
                            }
#undef ZITH
                        }
                    }
                }
            }
            size_t ij = (size_t)(ibra - 1) + (size_t)(jket - 1) * (size_t)nfun;
            for (size_t b = 0; b < nblk; ++b)                   /* :164-173 */
                for (int c = 0; c < ncomp_out && c < 4; ++c) {
                    zA[2 * (ij + nn * (b + nblk * (size_t)c)) + 0] = zs[2 * (b + nblk * c) + 0];
                    zA[2 * (ij + nn * (b + nblk * (size_t)c)) + 1] = zs[2 * (b + nblk * c) + 1];
                }
        }
    }
    free(zs);
    return rc;
}

/* ------------------------------------------------------------------------- *
 * TORMAT, matrix elements of r           (TorusFuns.f90:127-158)
 * rvecij(ni, li, nj, lj) = x^T Xij y with x = cinl(:, ni, li), y = cinl(:, nj, lj) through
 * DSVMV('U', ...) = DSYMV (upper triangle of Xij) + DDOT (Modules.f90:427-452).
 * cinl: (nfun, n1_max, 0:lmax) column-major; rvec: (n1_max, 0:lmax, n1_max, 0:lmax) column-major.
 * ------------------------------------------------------------------------- */
void bsp_tormat_rvec(int nfun, int n1_max, int lmax, const double *cinl, const double *Xij, double *rvec)
{
    double *v = (double *)malloc(sizeof(double) * (size_t)nfun);
    const size_t nl = (size_t)(lmax + 1), nn1 = (size_t)n1_max;
    for (size_t ni = 0; ni < nn1; ++ni)
        for (size_t li = 0; li < nl; ++li)
            for (size_t nj = 0; nj < nn1; ++nj)
                for (size_t lj = 0; lj < nl; ++lj) {
                    const double *x = cinl + (size_t)nfun * (ni + nn1 * li);
                    const double *y = cinl + (size_t)nfun * (nj + nn1 * lj);
                    /* v = Xij_sym(U) * y : DSYMV('U') reads A(i,j), i <= j, and mirrors it */
                    for (int i = 0; i < nfun; ++i) v[i] = 0.0;
                    for (int j = 0; j < nfun; ++j) {            /* reference BLAS column sweep */
                        double t1 = y[j], t2 = 0.0;
                        for (int i = 0; i < j; ++i) {
                            double a = Xij[(size_t)i + (size_t)j * (size_t)nfun];
                            v[i] += t1 * a;
                            t2 += a * y[i];
                        }
                        v[j] += t1 * Xij[(size_t)j + (size_t)j * (size_t)nfun] + t2;
                    }
                    double f = 0.0;
                    for (int i = 0; i < nfun; ++i) f += x[i] * v[i];   /* DDOT */
                    rvec[ni + nn1 * (li + nl * (nj + nn1 * lj))] = f;
                }
    free(v);
}

/* ------------------------------------------------------------------------- *
 * H_l = T + U_l + V                       (matrices.f90:244)
 * ------------------------------------------------------------------------- */
void bsp_hamiltonian(int nfun, const double *T, const double *Ul,
                     const double *V, double *H)
{
    size_t nn = (size_t)nfun * (size_t)nfun;
    for (size_t i = 0; i < nn; ++i) H[i] = T[i] + Ul[i] + V[i];
}

/* ------------------------------------------------------------------------- *
 * WRITE_WF wavefunction synthesis         (Bsp_Atom.f90:118-146)
 * psi(r_i), r_i = ra + i*(rb-ra)/npts, i = 0..npts  (npts = 10000 there)
 * ------------------------------------------------------------------------- */
int bsp_write_wf(int n, int k, int nkp, const double *rt, double ra, double rb,
                 const double *ci, int npts, double *r_out, double *psi_out)
{
    double bsp[ORACLE_MAXK];
    double dr = (rb - ra) / (double)npts;                       /* :120 */
    for (int i = 0; i <= npts; ++i) {
        double r = ra + (double)i * dr;                         /* :127 */
        for (int j = 0; j < k; ++j) bsp[j] = 0.0;
        int left, mflag;
        bsp_interv(rt, nkp, r, &left, &mflag);                  /* :130 */
        if (left < 1) return 2;
        if (bsp_bsplvb(nkp, rt, k, r, left, bsp)) return 1;     /* :131 */
        int jmin = left - k + 1;                                /* :133 */
        int jmax = jmin + k - 1;                                /* :134 */
        double sumf = 0.0;
        for (int j = jmin; j <= jmax; ++j) {                    /* :137-142 */
            int jfun = j - (left - k);
            double fr = 0.0;
            if (j >= 1 && j <= n) fr = ci[j - 1];
            if (jfun >= 1 && jfun <= k) sumf = sumf + fr * bsp[jfun - 1];
        }
        r_out[i] = r;
        psi_out[i] = sumf;
    }
    return 0;
}

/* ------------------------------------------------------------------------- *
 * TRANS_AMP dipole contraction, KIND_PI = 1,2   (PhotoIon.f90:90-105)
 *   v = A x   (DGEMV 'N')  ;  out(n) = DDOT(ci_fin(:,n), v)
 * The scalar prefactors An*c0 (:99,103) are applied by the caller.
 * A: N x N column-major; cfin: N x nfin column-major.
 * ------------------------------------------------------------------------- */
void bsp_dipole_dots(int n, const double *A, const double *x, int nfin,
                     const double *cfin, double *out)
{
    double *v = (double *)calloc((size_t)n, sizeof(double));
    /* reference DGEMV (netlib, TRANS='N', incx=incy=1): column sweep */
    for (int j = 0; j < n; ++j) {
        double temp = x[j];
        const double *a = A + (size_t)j * (size_t)n;
        for (int i = 0; i < n; ++i) v[i] = v[i] + temp * a[i];
    }
    for (int m = 0; m < nfin; ++m) {
        const double *u = cfin + (size_t)m * (size_t)n;
        double d = 0.0;
        for (int i = 0; i < n; ++i) d = d + u[i] * v[i];
        out[m] = d;
    }
    free(v);
}

/* ------------------------------------------------------------------------- *
 * Third comparator ("a third, independent comparator at N=1000"):
 * eigenvalues of the symmetric banded pencil H c = E S c by Sturm bisection in
 * EXTENDED precision (x87 long double, eps = 1.1e-19), straight on the band --
 * no reduction to standard or tridiagonal form, hence none of the eps*|E_max|
 * backward error DSYGV (matrices.f90:248: dpotrf/dsygst/dsytrd/dsteqr) carries.
 * nu(sigma) = number of negative pivots of the banded LDL^T of H - sigma S
 * (Sylvester's law of inertia, S positive definite) = number of eigenvalues
 * below sigma.  The result is the spectrum of the pencil *as stored in double*
 * to about 1e-18 |E_max|, the same definition as the 40-digit table of
 * tests/golden/make_golden.py, at sizes where mpmath takes hours.
 *
 * hb, sb: lower bands, hb[d*n + i] = H(i+d, i), d = 0..kd  (row-major (kd+1, n)).
 * idx[0..nidx): 0-based eigenvalue indices wanted; lo_hint/hi_hint (may be NULL):
 * a guess of a bracket per index (verified with two counts, widened if wrong).
 * ------------------------------------------------------------------------- */
static int band_count_ld(int n, int kd, const double *hb, const double *sb, long double sigma, long double pivmin)
{
    /* window w[r][c], r >= c: Schur complement of rows j..j+kd */
    long double w[ORACLE_MAXK][ORACLE_MAXK];
    int cnt = 0;
    for (int r = 0; r <= kd; ++r)
        for (int c = 0; c <= r; ++c) {
            /* entry (row r, col c), r - c = d */
            const int d = r - c;
            w[r][c] = (r < n) ? (long double)hb[(size_t)d * n + c] - sigma * (long double)sb[(size_t)d * n + c] : (r == c ? 1.0L : 0.0L);
        }
    for (int j = 0; j < n; ++j) {
        long double d = w[0][0];
        if (fabsl(d) < pivmin) d = -pivmin;
        if (d < 0.0L) ++cnt;
        long double l[ORACLE_MAXK], col0[ORACLE_MAXK];
        for (int i = 1; i <= kd; ++i) { col0[i] = w[i][0]; l[i] = col0[i] / d; }
        /* eliminate row/col j, shift the window up-left by one */
        for (int r = 1; r <= kd; ++r)
            for (int c = 1; c <= r; ++c) w[r - 1][c - 1] = w[r][c] - l[r] * col0[c];
        /* new last row: row j+kd+1 of the band, columns j+1 .. j+kd+1 */
        const int rn = j + kd + 1;
        for (int c = 0; c <= kd; ++c) {
            const int col = j + 1 + c, dd = rn - col;
            if (rn < n) w[kd][c] = (long double)hb[(size_t)dd * n + col] - sigma * (long double)sb[(size_t)dd * n + col];
            else w[kd][c] = (c == kd) ? 1.0L : 0.0L;
        }
    }
    return cnt;
}

int bsp_band_bisect_ld(int n, int kd, const double *hb, const double *sb, int nidx, const int *idx,
                       const double *lo_hint, const double *hi_hint, double *out)
{
    if (kd + 1 > ORACLE_MAXK) return -1;
    long double hmax = 0.0L, smax = 0.0L, s0 = 0.0L;
    for (int i = 0; i < n; ++i) {
        const long double h = fabsl((long double)hb[i]), s = (long double)sb[i];
        if (h > hmax) hmax = h;
        if (s > smax) smax = s;
        if (s > 0.0L && h / s > s0) s0 = h / s;
    }
    if (!(s0 > 0.0L)) s0 = 1.0L;
    /* global bounds by doubling */
    long double glo = -s0, ghi = s0;
    for (int t = 0; t < 80 && band_count_ld(n, kd, hb, sb, glo, 1e-4000L) > 0; ++t) glo *= 2.0L;
    for (int t = 0; t < 80 && band_count_ld(n, kd, hb, sb, ghi, 1e-4000L) < n; ++t) ghi *= 2.0L;
    for (int q = 0; q < nidx; ++q) {
        const int e = idx[q];
        if (e < 0 || e >= n) return -2;
        long double lo = glo, hi = ghi;
        if (lo_hint && hi_hint && lo_hint[q] < hi_hint[q]) {
            const long double a = lo_hint[q], b = hi_hint[q];
            if (band_count_ld(n, kd, hb, sb, a, 1e-4000L) <= e) lo = a;
            if (band_count_ld(n, kd, hb, sb, b, 1e-4000L) > e) hi = b;
        }
        /* invariant: nu(lo) <= e < nu(hi) */
        for (int it = 0; it < 200; ++it) {
            const long double mid = 0.5L * (lo + hi);
            if (!(mid > lo && mid < hi)) break;
            if (hi - lo <= 1e-19L * (fabsl(lo) + fabsl(hi))) break;
            const long double pm = 1e-4000L;
            if (band_count_ld(n, kd, hb, sb, mid, pm) <= e) lo = mid; else hi = mid;
        }
        out[q] = (double)(0.5L * (lo + hi));
    }
    return 0;
}
