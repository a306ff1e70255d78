/*
 * bsp_assembly_z.cuh -- the KIND_PI >= 3 branch of MATRIX_SVT (/root/reference/src/matrices.f90:110-139, 164-175):
 * the complex band matrices of the structured-light interaction,
 *
 *   zAij(ibra, jket, il, jl, c) = sum_{ibet} sum_{igl}  fbra * W_c(ibet, igl, il, jl) * (fket | dfket) * dr ,
 *
 * from the angular integrals zIth(ibet, igl, il, jl, :) the host tabulates on the radial quadrature grid
 * (ZINT_TH, Ang_Ints.f90:544-600 -- the angular machinery stays in the Fortran host):
 *   KIND_PI = 3, 4 :  c = 1: W = zIth(..,1) / r, fket      (matrices.f90:118-120)
 *                     c = 2: W = zIth(..,1),     dfket     (:121)        [c = 3, 4 stay zero: zsume/zsumf are
 *                                                                          never accumulated in this branch, :170-173]
 *   KIND_PI >= 5   :  c = 1, 2: W = zIth(..,c), fket       (:127-130);   KIND_PI >= 8: also c = 3, 4 (:131-136)
 * Same operand order as the reference (((fbra * W) * fk) * dr, ibet outer, igl inner); one (il, jl) pair = one
 * "block" blk = il + nlm * jl, exactly the memory order of zIth(nkp, ka, nlm, nm, ncomp) and zAij(.., nlm, nm, ncomp).
 *
 * grid = (ceil(n / BSP_ZTR), nblk), 128 threads.  Phase 1: the B-spline values and derivatives of the tile's
 * knot intervals at every quadrature point go to shared memory (one thread per point, bsp_deboor as in the real
 * assembly); phase 2: one thread per (band entry, component) walks its <= k intervals.
 */
#ifndef BSP_ASSEMBLY_Z_CUH
#define BSP_ASSEMBLY_Z_CUH

#include "bsp_assembly.cuh"

#define BSP_ZTR 16
#define BSP_ZTERMS 4

struct BspZTerm {
    int src;     /* 0-based component of zIth, -1: the output component stays zero */
    int div_r;   /* W = zIth / r */
    int deriv;   /* ket factor: 0 = B_j, 1 = B_j' */
};

struct BspZArgs {
    int n, nkp, ka, nblk, nterm;
    const double *rt, *xgwg;    /* knots; xg[32], wg[32] */
    const double2 *zIth;        /* (nkp, ka, nblk, ncomp_in), ibet fastest */
    double2 *zA;                /* general band AB(2k-1, n) per (blk, term): zA[((term*nblk + blk)*n + j)*(2k-1) + (k-1 + i - j)] */
    BspZTerm term[BSP_ZTERMS];
};

template <int K>
__global__ void __launch_bounds__(128) bsp_assemble_zaij_kernel(BspZArgs a)
{
    constexpr int B = K - 1, LD = 2 * B + 1, PP = 2 * K + 2;   /* per point: bsp[K], dbsp[K], r, dr */
    extern __shared__ double zsm[];
    const int blk = blockIdx.y;
    const int i0 = blockIdx.x * BSP_ZTR;
    const int n = a.n, ka = a.ka;
    const int rows = min(BSP_ZTR, n - i0);
    const int nint = rows + K - 1;
    const double *xg = a.xgwg, *wg = a.xgwg + 32;

    for (int t = threadIdx.x; t < nint * ka; t += blockDim.x) {
        const int q = t / ka, g = t - q * ka;
        const int m = i0 + 1 + q;                     /* 1-based interval [rt(m), rt(m+1)] = ibet */
        double *pt = zsm + (size_t)t * PP;
        double bsp[K], dbsp[K], r = 1.0, dr = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) { bsp[j] = 0.0; dbsp[j] = 0.0; }
        bool active = (m <= a.nkp - 1);
        double ta = 0.0, tb = 0.0;
        if (active) { ta = a.rt[m - 1]; tb = a.rt[m]; }
        if (!(tb - ta > 8.0 * BSP_EPS * fmax(fabs(ta), fabs(tb)))) active = false;   /* empty interval (quirk B-1) */
        if (active) {
            const double f1 = (tb + ta) / 2.0, f2 = (tb - ta) / 2.0;   /* matrices.f90:91-92 */
            r = f1 + xg[g] * f2;                                       /* :96 */
            dr = f2 * wg[g];                                           /* :97 */
            bsp_deboor<K>(a.rt, a.nkp, n, m, r, bsp, dbsp);            /* :100 */
            if (r == 0.0) r = BSP_EPS;                                 /* :102 */
        }
#pragma unroll
        for (int j = 0; j < K; ++j) { pt[j] = bsp[j]; pt[K + j] = dbsp[j]; }
        pt[2 * K] = r;
        pt[2 * K + 1] = dr;
    }
    __syncthreads();

    const int nout = rows * LD * a.nterm;
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {
        const int tm = o / (rows * LD), rc = o - tm * rows * LD;
        const int rr = rc / LD, c = rc - rr * LD;
        const int i = i0 + rr, j = i - B + c;
        if (j < 0 || j >= n) continue;
        const BspZTerm T = a.term[tm];
        double are = 0.0, aim = 0.0;
        if (T.src >= 0) {
            const int ibra = i + 1, jket = j + 1;
            const int bmin = max(ibra, jket), bmax = min(min(ibra, jket) + K - 1, a.nkp - 1);   /* :71-72 */
            const double2 *W = a.zIth + (size_t)a.nkp * ka * ((size_t)blk + (size_t)a.nblk * T.src);
            for (int ibet = bmin; ibet <= bmax; ++ibet) {              /* :89 */
                const int q = ibet - (i0 + 1);
                const int ia = ibra - ibet + K - 1, ib = jket - ibet + K - 1;   /* :105-106, 0-based */
                for (int g = 0; g < ka; ++g) {                         /* :94 */
                    const double *pt = zsm + (size_t)(q * ka + g) * PP;
                    const double fbra = pt[ia];
                    const double fk = T.deriv ? pt[K + ib] : pt[ib];
                    const double r = pt[2 * K], dr = pt[2 * K + 1];
                    double2 w = W[(size_t)(ibet - 1) + (size_t)a.nkp * g];
                    if (T.div_r) { w.x = w.x / r; w.y = w.y / r; }     /* zfAr = zAlm / r, :119 */
                    are += ((fbra * w.x) * fk) * dr;
                    aim += ((fbra * w.y) * fk) * dr;
                }
            }
        }
        a.zA[(((size_t)tm * a.nblk + blk) * n + j) * LD + (B + i - j)] = make_double2(are, aim);
    }
}

#endif
