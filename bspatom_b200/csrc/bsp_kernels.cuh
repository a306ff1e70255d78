/*
 * bsp_kernels.cuh -- __global__ wrappers around the per-thread bodies of
 * bsp_core.h plus the small data-movement kernels (band combine, transpose).
 * Thread mapping everywhere: blockIdx.y = pencil, blockIdx.x*blockDim.x +
 * threadIdx.x = eigen index e, so a warp streams 32 consecutive e of one
 * pencil: every workspace access X/R/L[row][e] is a fully coalesced 256 B
 * line and every band-row read is a warp-uniform broadcast that stays in L1.
 */
#ifndef BSP_KERNELS_CUH
#define BSP_KERNELS_CUH

#include <cuda_runtime.h>
#include "bsp_core.h"

#ifndef BSP_EIG_THREADS
#define BSP_EIG_THREADS 128
#endif

/* resident blocks per SM the register allocator is asked to allow, by half bandwidth:
 * the pivot window is (B+1)(B+2)/2 doubles, so wide bands get fewer blocks instead of spills */
#ifndef BSP_MINB_ROUND
#define BSP_MINB_ROUND 4
#endif
#ifndef BSP_MINB_FACTOR
#define BSP_MINB_FACTOR 3
#endif
#ifndef BSP_MINB_BACK
#define BSP_MINB_BACK 4
#endif
constexpr int bsp_minb_128(int base, int B) { return B <= 6 ? base : (B == 7 ? (base > 3 ? 3 : base) : 2); }
/* `base` counts blocks of 128 threads; larger blocks keep the same number of resident warps */
constexpr int bsp_minb(int base, int B) { return (bsp_minb_128(base, B) * 128) / BSP_EIG_THREADS > 0 ? (bsp_minb_128(base, B) * 128) / BSP_EIG_THREADS : 1; }

template <int B>
__global__ void __launch_bounds__(BSP_NCAND) bsp_bounds_kernel(BspEigChunk g, double *cand_s, int *cand_c)
{
    bsp_bounds_candidate<B>(g, blockIdx.x, threadIdx.x, cand_s, cand_c);
}

__global__ void bsp_zero_words_kernel(int *w, int n)
{
    if ((int)threadIdx.x < n) w[threadIdx.x] = 0;
}

__global__ void bsp_bounds_pick_kernel(BspEigChunk g, const double *cand_s, const int *cand_c)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < g.npencil) bsp_bounds_pick(g, p, cand_s, cand_c);
}

/* true in exactly one thread of the grid: thread 0 of the block that finishes last */
__device__ __forceinline__ bool bsp_last_block(int *arrive)
{
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned t = atomicAdd((unsigned *)arrive, 1u);
        s_last = (t == total - 1);
        if (s_last) { *arrive = 0; __threadfence(); }
    }
    __syncthreads();
    return s_last && threadIdx.x == 0;
}

template <int B>
__global__ void __launch_bounds__(BSP_EIG_THREADS, bsp_minb(BSP_MINB_ROUND, B)) bsp_round_kernel(BspEigChunk g, int round, int max_rounds, int open_ok)
{
    if (g.counters[BSP_C_BRACKETED]) return;       /* written by the previous kernel: grid-uniform */
    bsp_multisection_round<B>(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, round);
    if (bsp_last_block(g.counters + BSP_C_ARRIVE)) bsp_round_ctl(g, round, max_rounds, open_ok);
}

__global__ void bsp_prepare_kernel(BspEigChunk g)
{
    bsp_refine_prepare(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

template <int B>
__global__ void __launch_bounds__(BSP_EIG_THREADS, bsp_minb(BSP_MINB_FACTOR, B)) bsp_factor_kernel(BspEigChunk g, int iter, int optional)
{
    if (optional && g.counters[BSP_C_REFINED]) return;
    bsp_factor_forward<B>(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, iter);
}

template <int B>
__global__ void __launch_bounds__(BSP_EIG_THREADS, bsp_minb(BSP_MINB_BACK, B)) bsp_back_kernel(BspEigChunk g, int corr_now, int corr_next, int optional)
{
    if (optional && g.counters[BSP_C_REFINED]) return;
    bsp_back_substitute<B>(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, corr_now, corr_next);
}

template <int B>
__global__ void __launch_bounds__(BSP_EIG_THREADS, bsp_minb(BSP_MINB_FACTOR, B)) bsp_factor_ckpt_kernel(BspEigChunk g, int iter, int optional)
{
    if (optional && g.counters[BSP_C_REFINED]) return;
    bsp_factor_checkpoint<B>(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, iter);
}

template <int B>
__global__ void __launch_bounds__(BSP_EIG_THREADS, bsp_minb(BSP_MINB_FACTOR, B)) bsp_back_rc_kernel(BspEigChunk g, int corr_now, int corr_next, int iter, int optional)
{
    if (optional && g.counters[BSP_C_REFINED]) return;
    bsp_back_recompute<B>(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, corr_now, corr_next, iter);
}

__global__ void bsp_check_kernel(BspEigChunk g, int iter)
{
    if (g.counters[BSP_C_REFINED]) return;
    bsp_check_converged(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, 1);
    if (bsp_last_block(g.counters + BSP_C_ARRIVE)) bsp_check_ctl(g, iter);
}

/* copy the chunk's control block into the run report and clear it for the next chunk of this stream */
__global__ void bsp_report_kernel(BspEigChunk g, int *report)
{
    if (threadIdx.x < BSP_C_WORDS) {
        report[threadIdx.x] = g.counters[threadIdx.x];
        g.counters[threadIdx.x] = 0;
    }
}

__global__ void bsp_finalize_kernel(BspEigChunk g, double *E, double *fac, int *bad, double res_tol)
{
    bsp_finalize_eigen(g, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, E, fac, bad, res_tol);
}

/* C[p] (n x nvec_p, column-major, column e = eigenvector e) = fac[e] * X[p][row][e]
 * 32x32 shared-memory transpose; coff[p] = offset of pencil p's block in C. */
__global__ void bsp_transpose_kernel(BspEigChunk g, const double *fac, double *C, const long long *coff)
{
    __shared__ double tile[32][33];
    const int p = blockIdx.z;
    const int nv = g.nvec[p];
    const int e0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (e0 >= nv) return;
    const double *X = g.X + (size_t)p * g.xrows * g.ldw;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int r = r0 + rr, e = e0 + threadIdx.x;
        double v = 0.0;
        if (r < g.n && e < nv) v = X[(size_t)r * g.ldw + e] * fac[(size_t)p * g.ldw + e];
        tile[rr][threadIdx.x] = v;
    }
    __syncthreads();
    double *Cp = C + coff[p];
    for (int ee = threadIdx.y; ee < 32; ee += blockDim.y) {
        const int e = e0 + ee, r = r0 + threadIdx.x;
        if (e < nv && r < g.n) Cp[(size_t)e * g.n + r] = tile[threadIdx.x][ee];
    }
}

/* fbH[p] = fbH0[inst] + cl[p] * fbQ[inst]   (Hij = Tij + Uij(:,:,l) + Vij, matrices.f90:244) */
__global__ void bsp_combine_kernel(double *fbH, const double *fbH0, const double *fbQ, const int *inst,
                                   const double *cl, size_t per_mat)
{
    const int p = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_mat) return;
    const size_t src = (size_t)inst[p] * per_mat + i;
    fbH[(size_t)p * per_mat + i] = fma(cl[p], fbQ[src], fbH0[src]);
}

/* S positive definite?  one thread per instance walks the LDL^T pivots of S with the register-window
 * recurrence of the Sturm count (H := S, sigma = 0).  pd_info as below. */
template <int B>
__global__ void bsp_pdcheck_fast_kernel(const double *fbS, int n, int npad, int nrows, int ninst, int *pd_info)
{
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= ninst) return;
    const double *S = fbS + (size_t)inst * nrows * (2 * B + 2);
    int first = -1;
    bsp_sturm_count<B>(S, S, npad, 0.0, 1e-300, &first);
    pd_info[inst] = (first >= 0 && first < n) ? first + 1 : 0;
}

/* S positive definite?  one thread per instance: banded Cholesky pivots.
 * pd_info[inst] = 0 or the 1-based index of the first non-positive pivot
 * (LAPACK dpotrf numbering; DSYGV reports N + that, matrices.f90:250). */
__global__ void bsp_pdcheck_kernel(const double *fbS, int n, int nrows, int B, int ninst, int *pd_info,
                                   double *Lout /* NULL or [ninst][n][B+1]: L(j+i,j) */)
{
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= ninst) return;
    const int FS = 2 * B + 2;
    const double *S = fbS + (size_t)inst * nrows * FS;
    /* window of the last B columns of L (scaled), generic B <= 15 */
    double w[16][16];
    for (int a = 0; a < 16; ++a) for (int b = 0; b < 16; ++b) w[a][b] = 0.0;
    int info = 0;
    for (int j = 0; j < n && !info; ++j) {
        /* column j of the Schur complement: a(j+i,j) - sum_{c<j} l(j+i,c) l(j,c) d(c) */
        double col[16];
        for (int i = 0; i <= B; ++i) {
            double v = (j + i < n) ? S[(size_t)(j + i) * FS + (B - i)] : 0.0;
            /* previous columns c = j-1 .. j-B are kept in w[(c)%16][*] as l(c+t,c)*sqrt(d) */
            for (int t = 1; t <= B - i; ++t) {
                const int c = j - t;
                if (c < 0) break;
                v -= w[c & 15][t + i] * w[c & 15][t];
            }
            col[i] = v;
        }
        if (!(col[0] > 0.0)) { info = j + 1; break; }
        const double d = sqrt(col[0]);
        w[j & 15][0] = d;
        for (int i = 1; i <= B; ++i) w[j & 15][i] = col[i] / d;
        for (int i = B + 1; i < 16; ++i) w[j & 15][i] = 0.0;
        if (Lout)
            for (int i = 0; i <= B; ++i) Lout[((size_t)inst * n + j) * (B + 1) + i] = w[j & 15][i];
    }
    pd_info[inst] = info;
}

#endif
