/*
 * bsp_assembly.cuh -- banded B-spline matrix assembly on the GPU.
 *
 * Replaces MATRIX_SVT (matrices.f90:1-200, scalar branch KIND_PI <= 2):
 *   * the reference visits all N^2 (ibra,jket) pairs and, for each, re-runs
 *     interv + two bsplvb calls at every quadrature point of every common
 *     knot interval (k^2-fold redundancy, O(nkp) interval search each time);
 *   * here one half-warp (ka <= 16) or warp owns one knot interval: lane g
 *     evaluates the de Boor recursion (bsplvb.f90:38-50) once at quadrature
 *     point g in registers, `left` is the interval index (interv.f90 becomes a
 *     no-op), the k(k+1)/2 pair products of every matrix are reduced over the
 *     lanes with xor-shuffles (fixed tree -> deterministic), and the per-
 *     interval k x k blocks are summed in ascending interval order (the
 *     reference's doBetween order, matrices.f90:89) into full-band rows.
 *
 * Output layout ("full-band rows"): fb[i*FS + c] = A(i, i-B+c), B = k-1,
 * FS = 2B+2, rows n..nrows-1 padded (diag(H0) = 1, rest 0).
 */
#ifndef BSP_ASSEMBLY_CUH
#define BSP_ASSEMBLY_CUH

#include <cuda_runtime.h>
#include <math.h>
#include "bsp_core.h"

#define BSP_MAT_S 0
#define BSP_MAT_H0 1
#define BSP_MAT_Q 2
#define BSP_MAT_T 3
#define BSP_MAT_V 4
#define BSP_MAT_R 5
#define BSP_MAT_RINV 6
#define BSP_MAT_D 7
#define BSP_NMAT 8

#define BSP_ASM_TR 32      /* output rows per CTA */
#define BSP_ASM_THREADS 128

struct BspInstParams {
    int pot_kind;
    int has_vtab;
    double par[8];
};

struct BspAsmArgs {
    int n, nkp, ka, nrows, ninst;
    int want_pi;              /* 0: S,H0,Q only ; 1: also T,V,R,Rinv,D            */
    const double *rt;         /* [ninst][nkp]                                     */
    const double *xgwg;       /* [ninst][64]: xg[0..31], wg[0..31]                */
    const BspInstParams *par; /* [ninst]                                          */
    const double *vtab;       /* [ninst][(nkp-1)*ka] or NULL                      */
    double *fb[BSP_NMAT];     /* each [ninst][nrows][FS] (NULL if not wanted)     */
};

/* SELPOT (Modules.f90:263-295) + the cfg3 families of SURVEY.md 8(d) */
__device__ __forceinline__ double bsp_potential(int kind, const double *par, double r)
{
    double vr;
    if (kind == 0 || kind == 2) {
        vr = -par[0] / r;
    } else if (kind == 1) {
        vr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) vr = vr + par[2 + i] * exp(-par[5 + i] * r);
        vr = -1.0 * (par[0] - par[1] + vr) / r;
    } else if (kind == 10) {
        vr = -par[0] * exp(-par[1] * r) / r;
    } else if (kind == 11) {
        const double d = 1.0 + par[1] * r;
        vr = -(1.0 + (par[0] - 1.0) / (d * d)) / r;
    } else {
        vr = 0.0;
    }
    return vr;
}

/* values of the K non-zero B-splines of order K at x in interval `left`
 * (1-based) and their first derivatives: BSPALL (Modules.f90:85-108) with the
 * order K-1 values taken from the same recursion (bsplvb.f90:38-50). */
template <int K>
__device__ __forceinline__ void bsp_deboor(const double *__restrict__ rt, int nkp, int nfun, int left,
                                           double x, double (&bsp)[K], double (&dbsp)[K])
{
    auto T = [&](int i) -> double {   /* 1-based, clamped like the oracle */
        i = i < 1 ? 1 : (i > nkp ? nkp : i);
        return __ldg(rt + i - 1);
    };
    double deltar[K], deltal[K], bsp1[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { bsp[j] = 0.0; bsp1[j] = 0.0; }
    bsp[0] = 1.0;
    if (K == 2) bsp1[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
        deltar[j - 1] = T(left + j) - x;
        deltal[j - 1] = x - T(left + 1 - j);
        double saved = 0.0;
#pragma unroll
        for (int i = 1; i <= j; ++i) {
            const double term = bsp[i - 1] / (deltar[i - 1] + deltal[j - i]);
            bsp[i - 1] = saved + deltar[i - 1] * term;
            saved = deltal[j - i] * term;
        }
        bsp[j] = saved;
        if (j == K - 2) {
#pragma unroll
            for (int i = 0; i < K - 1; ++i) bsp1[i] = bsp[i];
        }
    }
    /* dbsp(j) = (k-1) (A1 bspp(j) - A2 bspp(j+1)), bspp(1)=0, bspp(j+1)=bsp1(j) */
#pragma unroll
    for (int j = 1; j <= K; ++j) {
        const int jp = j + (left - K);
        double a1 = 0.0, a2 = 0.0;
        if (jp >= 1 && jp <= nfun) {       /* Aind, grid.f90:82-91 */
            double d = T(jp + K - 1) - T(jp);
            if (d > 0.0) a1 = 1.0 / d;
            d = T(jp + K) - T(jp + 1);
            if (d > 0.0) a2 = 1.0 / d;
        }
        const double b1 = (j >= 2) ? bsp1[j - 2] : 0.0;
        const double b2 = (j <= K - 1) ? bsp1[j - 1] : 0.0;
        dbsp[j - 1] = (double)(K - 1) * (a1 * b1 - a2 * b2);
    }
}

template <int LPI>
__device__ __forceinline__ double bsp_seg_reduce(double v)
{
#pragma unroll
    for (int off = LPI / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

/*
 * grid = (ceil(nrows/TR), ninst), block = 128.
 * dynamic smem: nint_max * (nsym*PK + (want_pi ? K*K : 0)) doubles,
 * nint_max = TR + K - 1, PK = K(K+1)/2, nsym = want_pi ? 6 : 4 (S,T,V,Q[,R,Rinv]).
 */
template <int K, int LPI>
__global__ void __launch_bounds__(BSP_ASM_THREADS) bsp_assemble_kernel(BspAsmArgs a)
{
    constexpr int B = K - 1, FS = 2 * B + 2, PK = K * (K + 1) / 2;
    constexpr int IPW = 32 / LPI;
    extern __shared__ double sm[];
    const int inst = blockIdx.y;
    const int i0 = blockIdx.x * BSP_ASM_TR;
    const int n = a.n;
    const int rows = min(BSP_ASM_TR, a.nrows - i0);
    const int rows_real = max(0, min(BSP_ASM_TR, n - i0));
    const int nint = rows_real > 0 ? rows_real + K - 1 : 0;
    const int nsym = a.want_pi ? 6 : 4;
    const int per_int = nsym * PK + (a.want_pi ? K * K : 0);
    const double *rt = a.rt + (size_t)inst * a.nkp;
    const double *xg = a.xgwg + (size_t)inst * 64, *wg = xg + 32;
    const BspInstParams P = a.par[inst];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = BSP_ASM_THREADS / 32;

    /* ---- phase 1: per-interval local matrices -------------------------- */
    for (int q0 = warp * IPW; q0 < nint; q0 += nwarps * IPW) {
        const int q = q0 + lane / LPI;
        const int g = lane % LPI;
        const int m = i0 + 1 + q;            /* 1-based interval [rt(m), rt(m+1)] */
        bool active = (q < nint) && (g < a.ka) && (m <= a.nkp - 1);
        double ta = 0.0, tb = 0.0;
        if (active) { ta = __ldg(rt + m - 1); tb = __ldg(rt + m); }
        /* zero-width, negative or ulp-wide intervals contribute nothing
         * (reference quirk B-1) */
        if (!(tb - ta > 8.0 * BSP_EPS * fmax(fabs(ta), fabs(tb)))) active = false;
        double bsp[K], dbsp[K];
        double r = 1.0, dr = 0.0, vpot = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) { bsp[j] = 0.0; dbsp[j] = 0.0; }
        if (active) {
            const double f1 = (tb + ta) / 2.0;             /* matrices.f90:91 */
            const double f2 = (tb - ta) / 2.0;             /* :92 */
            r = f1 + xg[g] * f2;                           /* :96 */
            dr = f2 * wg[g];                               /* :97 */
            bsp_deboor<K>(rt, a.nkp, n, m, r, bsp, dbsp);  /* :100 */
            if (r == 0.0) r = BSP_EPS;                     /* :102 */
            if (P.has_vtab) vpot = __ldg(a.vtab + ((size_t)inst * (a.nkp - 1) + (m - 1)) * a.ka + g);
            else vpot = bsp_potential(P.pot_kind, P.par, r); /* :103 */
        }
        const double rinv = 1.0 / r;
        const double vcent = 1.0 / (2.0 * (r * r));        /* :149 with l(l+1) factored out */
        double *loc = sm + (size_t)(q < nint ? q : 0) * per_int;
        const bool writer = (g == 0) && (q < nint);
        int pk = 0;
#pragma unroll
        for (int ia = 0; ia < K; ++ia) {
#pragma unroll
            for (int ib = ia; ib < K; ++ib) {
                const double fa = bsp[ia], fk = bsp[ib];
                double vS = fa * fk * dr;                       /* :145 */
                double vT = dbsp[ia] * 0.5 * dbsp[ib] * dr;     /* :147 */
                double vV = fa * vpot * fk * dr;                /* :146 */
                double vQ = fa * vcent * fk * dr;               /* :152 */
                vS = bsp_seg_reduce<LPI>(vS);
                vT = bsp_seg_reduce<LPI>(vT);
                vV = bsp_seg_reduce<LPI>(vV);
                vQ = bsp_seg_reduce<LPI>(vQ);
                if (writer) {
                    loc[0 * PK + pk] = vS; loc[1 * PK + pk] = vT;
                    loc[2 * PK + pk] = vV; loc[3 * PK + pk] = vQ;
                }
                if (a.want_pi) {
                    double vR = fa * r * fk * dr;               /* :144 */
                    double vI = fa * rinv * fk * dr;            /* :141 */
                    vR = bsp_seg_reduce<LPI>(vR);
                    vI = bsp_seg_reduce<LPI>(vI);
                    if (writer) { loc[4 * PK + pk] = vR; loc[5 * PK + pk] = vI; }
                }
                ++pk;
            }
        }
        if (a.want_pi) {
#pragma unroll
            for (int ia = 0; ia < K; ++ia) {
#pragma unroll
                for (int ib = 0; ib < K; ++ib) {
                    double vD = bsp[ia] * dbsp[ib] * dr;        /* :142 */
                    vD = bsp_seg_reduce<LPI>(vD);
                    if (writer) loc[6 * PK + ia * K + ib] = vD;
                }
            }
        }
    }
    __syncthreads();

    /* ---- phase 2: sum the <= K interval blocks of every band entry ------ */
    const int nout = rows * (2 * B + 2);
    for (int o = threadIdx.x; o < nout; o += BSP_ASM_THREADS) {
        const int rr = o / FS, c = o % FS;
        const int i = i0 + rr, j = i - B + c;
        double s[BSP_NMAT];
#pragma unroll
        for (int mm = 0; mm < BSP_NMAT; ++mm) s[mm] = 0.0;
        if (i < n && c <= 2 * B && j >= 0 && j < n) {
            const int ibra = i + 1, jket = j + 1;
            const int bmin = max(ibra, jket), bmax = min(ibra, jket) + K - 1;   /* :71-72 */
            double sS = 0, sT = 0, sV = 0, sQ = 0, sR = 0, sI = 0, sD = 0;
            for (int ibet = bmin; ibet <= bmax; ++ibet) {                       /* :89 */
                const int q = ibet - (i0 + 1);
                const int ia = ibra - ibet + K - 1, ib = jket - ibet + K - 1;   /* :105-106 */
                const int lo = min(ia, ib), hi = max(ia, ib);
                const int pk = lo * K - (lo * (lo - 1)) / 2 + (hi - lo);
                const double *loc = sm + (size_t)q * per_int;
                sS += loc[0 * PK + pk]; sT += loc[1 * PK + pk];
                sV += loc[2 * PK + pk]; sQ += loc[3 * PK + pk];
                if (a.want_pi) {
                    sR += loc[4 * PK + pk]; sI += loc[5 * PK + pk];
                    sD += loc[6 * PK + ia * K + ib];
                }
            }
            s[BSP_MAT_S] = sS; s[BSP_MAT_H0] = sT + sV; s[BSP_MAT_Q] = sQ; s[BSP_MAT_T] = sT;
            s[BSP_MAT_V] = sV; s[BSP_MAT_R] = sR; s[BSP_MAT_RINV] = sI; s[BSP_MAT_D] = sD;
        } else if (i >= n && c == B) {
            s[BSP_MAT_H0] = 1.0;   /* padding row: decoupled, positive pivot */
        }
        const size_t off = ((size_t)inst * a.nrows + i) * FS + c;
#pragma unroll
        for (int mm = 0; mm < BSP_NMAT; ++mm)
            if (a.fb[mm]) a.fb[mm][off] = s[mm];
    }
}

#endif
