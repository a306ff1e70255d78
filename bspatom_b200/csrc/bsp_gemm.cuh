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
                                            const double *__restrict__ Ci, double *__restrict__ Y,
                                            long long strideC = -1, long long strideY = -1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (i >= n || v >= nv) return;
    const int ld = 2 * kd + 1;
    /* batch of vector blocks sharing the operator (default: blocks of n x nv back to back) */
    Ci += (size_t)blockIdx.z * (strideC < 0 ? (long long)n * nv : strideC);
    Y += (size_t)blockIdx.z * (strideY < 0 ? (long long)n * nv : strideY);
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

/* m16n8k8 f64: A 16x8 (row), B 8x8 (col), C 16x8.  Fragments (g = lane>>2, t = lane&3):
 *   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (t, g) b1 (t+4, g);
 *   c0,c1 (g, 2t..2t+1)  c2,c3 (g+8, 2t..2t+1) */
__device__ __forceinline__ void bsp_dmma_m16n8k8(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

/*
 * D(M x N, column-major, ldd) = A^T B with A: K x M (column-major, lda),
 * B: K x N (column-major, ldb); i.e. D(m,n) = sum_k A(k,m) B(k,n): both
 * operands are contiguous along k ("TN").
 * CTA tile 64 x 64 x 16, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA tiles.
 * Shared tiles are [64][16+4]: row stride 20 doubles makes the fragment loads
 * (8 rows x 4 k per warp) bank-conflict free.
 */
#ifndef BSP_DMMA_SHAPE
#define BSP_DMMA_SHAPE 1688 /* 884: m8n8k4 (sm_80 shape), 1688: m16n8k8 */
#endif
#define BSP_GT_M 64
#define BSP_GT_N 64
#define BSP_GT_K 16
#define BSP_GT_LD 20

__global__ void __launch_bounds__(128) bsp_dgemm_tn_kernel(int M, int N, int K, const double *__restrict__ A, int lda,
                                                           const double *__restrict__ Bm, int ldb,
                                                           double *__restrict__ D, int ldd,
                                                           long long strideA, long long strideB, long long strideD)
{
    A += (size_t)blockIdx.z * strideA;     /* batched: grid.z problems with constant strides */
    Bm += (size_t)blockIdx.z * strideB;
    D += (size_t)blockIdx.z * strideD;
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
#if BSP_DMMA_SHAPE == 1688
    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.0;
#else
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#endif

    const int nk = (K + BSP_GT_K - 1) / BSP_GT_K;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BSP_GT_K);
#if BSP_DMMA_SHAPE == 1688
#pragma unroll
        for (int kk = 0; kk < BSP_GT_K; kk += 8) {
            double af[2][4], bf[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                af[i][0] = As[buf][wm + i * 16 + g][kk + tg];
                af[i][1] = As[buf][wm + i * 16 + g + 8][kk + tg];
                af[i][2] = As[buf][wm + i * 16 + g][kk + tg + 4];
                af[i][3] = As[buf][wm + i * 16 + g + 8][kk + tg + 4];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bf[j][0] = Bs[buf][wn + j * 8 + g][kk + tg];
                bf[j][1] = Bs[buf][wn + j * 8 + g][kk + tg + 4];
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) bsp_dmma_m16n8k8(acc[i][j], af[i], bf[j]);
        }
#else
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
#endif
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#if BSP_DMMA_SHAPE == 1688
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int m = m0 + wm + i * 16 + g + hh * 8;
                const int nn = n0 + wn + j * 8 + tg * 2;
                if (m < M) {
                    if (nn < N) D[(size_t)nn * ldd + m] = acc[i][j][hh * 2 + 0];
                    if (nn + 1 < N) D[(size_t)(nn + 1) * ldd + m] = acc[i][j][hh * 2 + 1];
                }
            }
#else
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
#endif
}

/*
 * Large-tile variant for M, N >= 128: CTA tile 128 x 128 x 16, 8 warps (4 along M x 2 along N), warp tile
 * 32 x 64 = 4 x 8 DMMA.8x8x4 tiles (32 independent accumulator tiles per warp hide the DMMA latency),
 * double-buffered shared tiles [128][16+4] filled through registers with 16-byte global loads.
 * Arithmetic intensity per CTA 16 flop/B of L2 traffic (64 x 64 tiles: 8 flop/B).
 * Requires lda, ldb even and 16-byte aligned operands (checked by the launcher).
 */
#define BSP_G2_M 128
#define BSP_G2_N 128
#define BSP_G2_K 16
#define BSP_G2_LD 20
#define BSP_G2_SMEM (2 * (BSP_G2_M + BSP_G2_N) * BSP_G2_LD * (int)sizeof(double))

__global__ void __launch_bounds__(256, 1) bsp_dgemm_tn128_kernel(int M, int N, int K, const double *__restrict__ A, int lda,
                                                                 const double *__restrict__ Bm, int ldb,
                                                                 double *__restrict__ D, int ldd,
                                                                 long long strideA, long long strideB, long long strideD)
{
    extern __shared__ double sm2[];
    double (*As)[BSP_G2_M][BSP_G2_LD] = (double (*)[BSP_G2_M][BSP_G2_LD])sm2;
    double (*Bs)[BSP_G2_N][BSP_G2_LD] = (double (*)[BSP_G2_N][BSP_G2_LD])(sm2 + 2 * BSP_G2_M * BSP_G2_LD);
    A += (size_t)blockIdx.z * strideA;
    Bm += (size_t)blockIdx.z * strideB;
    D += (size_t)blockIdx.z * strideD;
    const int m0 = blockIdx.x * BSP_G2_M, n0 = blockIdx.y * BSP_G2_N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int g = lane >> 2, tg = lane & 3;
    const int lr = tid >> 1, lk = (tid & 1) * 8;     /* row 0..127, k half */
    double2 ra[4], rb[4];
    auto load_tiles = [&](int k0) {
        const int m = m0 + lr, nn = n0 + lr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + 2 * q;
            ra[q] = make_double2(0.0, 0.0);
            rb[q] = make_double2(0.0, 0.0);
            if (m < M) {
                if (k + 1 < K) ra[q] = __ldg(reinterpret_cast<const double2 *>(A + (size_t)m * lda + k));
                else if (k < K) ra[q].x = __ldg(A + (size_t)m * lda + k);
            }
            if (nn < N) {
                if (k + 1 < K) rb[q] = __ldg(reinterpret_cast<const double2 *>(Bm + (size_t)nn * ldb + k));
                else if (k < K) rb[q].x = __ldg(Bm + (size_t)nn * ldb + k);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<double2 *>(&As[buf][lr][lk + 2 * q]) = ra[q];
            *reinterpret_cast<double2 *>(&Bs[buf][lr][lk + 2 * q]) = rb[q];
        }
    };
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int nk = (K + BSP_G2_K - 1) / BSP_G2_K;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BSP_G2_K);
#pragma unroll
        for (int kk = 0; kk < BSP_G2_K; kk += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[buf][wm + i * 8 + g][kk + tg];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = Bs[buf][wn + j * 8 + g][kk + tg];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) bsp_dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int m = m0 + wm + i * 8 + g;
            const int nn = n0 + wn + j * 8 + tg * 2;
            if (m < M) {
                if (nn < N) D[(size_t)nn * ldd + m] = acc[i][j][0];
                if (nn + 1 < N) D[(size_t)(nn + 1) * ldd + m] = acc[i][j][1];
            }
        }
}

/* picks the tile variant; returns the CUDA status of the launch */
static inline cudaError_t bsp_launch_dgemm_tn(cudaStream_t st, int M, int N, int K, const double *A, int lda,
                                              const double *Bm, int ldb, double *D, int ldd, int batch,
                                              long long strideA, long long strideB, long long strideD)
{
    const bool aligned = ((lda | ldb) % 2 == 0) && ((strideA | strideB) % 2 == 0) &&
                         (((size_t)A | (size_t)Bm) % 16 == 0);
    const long long ctas128 = (long long)((M + BSP_G2_M - 1) / BSP_G2_M) * ((N + BSP_G2_N - 1) / BSP_G2_N) * batch;
    if (M >= 128 && N >= 128 && aligned && ctas128 >= 148) {   /* enough large tiles to fill the 148 SMs */
        cudaError_t e = cudaFuncSetAttribute(bsp_dgemm_tn128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BSP_G2_SMEM);
        if (e != cudaSuccess) return e;
        dim3 grid((M + BSP_G2_M - 1) / BSP_G2_M, (N + BSP_G2_N - 1) / BSP_G2_N, batch);
        bsp_dgemm_tn128_kernel<<<grid, 256, BSP_G2_SMEM, st>>>(M, N, K, A, lda, Bm, ldb, D, ldd, strideA, strideB, strideD);
    } else {
        dim3 grid((M + BSP_GT_M - 1) / BSP_GT_M, (N + BSP_GT_N - 1) / BSP_GT_N, batch);
        bsp_dgemm_tn_kernel<<<grid, 128, 0, st>>>(M, N, K, A, lda, Bm, ldb, D, ldd, strideA, strideB, strideD);
    }
    return cudaGetLastError();
}

#endif
