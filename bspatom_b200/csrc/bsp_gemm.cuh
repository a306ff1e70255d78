/*
 * bsp_gemm.cuh -- the only dense FP64 contraction of the path:
 *     D(nf, ni) = Cf^T * (A * Ci)
 * A banded (dipole operator between B-splines: rij, matrices.f90:160-163),
 * Cf/Ci eigenvector blocks.  Replaces the DGEMV + per-state DDOT loop of
 * TRANS_AMP (PhotoIon.f90:90-105) and its O(nbra*nket*N^2) generalisation
 * (PhotoIon.f90:188-250) by one banded product (HBM bound) and one TN GEMM on
 * the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no f64 kind).
 */
#ifndef BSP_GEMM_CUH
#define BSP_GEMM_CUH

#include <cuda_runtime.h>

/* Y(i, v) = sum_j A(i,j) Ci(j, v),  A in LAPACK general band storage
 * AB[(kd + i - j) + j*ld], ld = 2kd+1.  i fastest over threads: coalesced. */
__global__ void bsp_band_times_dense_kernel(int n, int kd, const double *__restrict__ AB, int nv,
                                            const double *__restrict__ Ci, double *__restrict__ Y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (i >= n || v >= nv) return;
    const int ld = 2 * kd + 1;
    const double *x = Ci + (size_t)v * n;
    double s = 0.0;
    const int j0 = max(0, i - kd), j1 = min(n - 1, i + kd);
    for (int j = j0; j <= j1; ++j) s = fma(__ldg(AB + (size_t)j * ld + (kd + i - j)), __ldg(x + j), s);
    Y[(size_t)v * n + i] = s;
}

__device__ __forceinline__ void bsp_dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

/*
 * D(M x N, column-major, ldd) = A^T B with A: K x M (column-major, lda),
 * B: K x N (column-major, ldb); i.e. D(m,n) = sum_k A(k,m) B(k,n): both
 * operands are contiguous along k ("TN").
 * CTA tile 64 x 64 x 16, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles.
 * Shared tiles are [64][16+4]: row stride 20 doubles makes the fragment loads
 * (8 rows x 4 k per warp) bank-conflict free.
 */
#define BSP_GT_M 64
#define BSP_GT_N 64
#define BSP_GT_K 16
#define BSP_GT_LD 20

__global__ void __launch_bounds__(128) bsp_dgemm_tn_kernel(int M, int N, int K, const double *__restrict__ A, int lda,
                                                           const double *__restrict__ Bm, int ldb,
                                                           double *__restrict__ D, int ldd)
{
    __shared__ double As[2][BSP_GT_M][BSP_GT_LD];
    __shared__ double Bs[2][BSP_GT_N][BSP_GT_LD];
    const int m0 = blockIdx.x * BSP_GT_M, n0 = blockIdx.y * BSP_GT_N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int g = lane >> 2, tg = lane & 3;

    /* global -> register staging: each thread moves 8 doubles of A and 8 of B
     * per k-tile: row r = tid/2 (0..63), k half = (tid&1)*8 */
    const int lr = tid >> 1, lk = (tid & 1) * 8;
    double ra[8], rb[8];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + lk + q;
            const int m = m0 + lr, nn = n0 + lr;
            ra[q] = (m < M && k < K) ? __ldg(A + (size_t)m * lda + k) : 0.0;
            rb[q] = (nn < N && k < K) ? __ldg(Bm + (size_t)nn * ldb + k) : 0.0;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            As[buf][lr][lk + q] = ra[q];
            Bs[buf][lr][lk + q] = rb[q];
        }
    };
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int nk = (K + BSP_GT_K - 1) / BSP_GT_K;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BSP_GT_K);
#pragma unroll
        for (int kk = 0; kk < BSP_GT_K; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[buf][wm + i * 8 + g][kk + tg];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[buf][wn + j * 8 + g][kk + tg];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) bsp_dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + wm + i * 8 + g;
            const int nn = n0 + wn + j * 8 + tg * 2;
            if (m < M) {
                if (nn < N) D[(size_t)nn * ldd + m] = acc[i][j][0];
                if (nn + 1 < N) D[(size_t)(nn + 1) * ldd + m] = acc[i][j][1];
            }
        }
}

#endif
