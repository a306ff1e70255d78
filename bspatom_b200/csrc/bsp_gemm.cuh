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

/* Y(i, v) = sum_j A(i,j) Ci(j, v),  A in LAPACK general band storage AB[(kd + i - j) + j*ld], ld = 2kd+1.
 * A block owns 128 rows x BSP_BTD_VECS vectors; thread = row.  The (128 + 2kd) x VECS tile of C goes to shared memory
 * with asynchronous 8-byte copies, all in flight at once; the thread's band row lives in registers (KD is a template
 * parameter; KD = 0 is the generic version that keeps it in shared memory).  HBM bound: reads C once, writes Y once.
 * History (cfg5 chain, 50 x 8 MB in, 50 x 8 MB out): one block per (128 rows, vector) 0.60 ms; 16 vectors per block
 * with x loaded element by element inside the FMA loop 0.60 ms (one or two misses in flight per warp); C tile through
 * cp.async but band row and x both read from shared memory per FMA 0.50 ms (ncu: LSU pipe 98 % busy, 26 LDS.64 per
 * output element); band row in registers: see profiles/README.md. */
#define BSP_BTD_VECS 16
template <int KD>
__global__ void __launch_bounds__(128) bsp_band_times_dense_kernel(int n, int kd, const double *__restrict__ AB, int nv,
                                                                   const double *__restrict__ Ci, double *__restrict__ Y,
                                                                   long long strideC, long long strideY)
{
    extern __shared__ double btd_sm[];      /* x [VECS][128 + 2kd] (| band [128][2kd+2] for KD = 0) */
    if (KD > 0) kd = KD;
    const int i0 = blockIdx.x * 128, i = i0 + threadIdx.x;
    const int v0 = blockIdx.y * BSP_BTD_VECS;
    const int ld = 2 * kd + 1, W = 128 + 2 * kd;
    double *xs = btd_sm;
    /* batch of vector blocks sharing the operator (default: blocks of n x nv back to back) */
    Ci += (size_t)blockIdx.z * (strideC < 0 ? (long long)n * nv : strideC);
    Y += (size_t)blockIdx.z * (strideY < 0 ? (long long)n * nv : strideY);
    const int nvb = min(BSP_BTD_VECS, nv - v0);
    for (int q = threadIdx.x; q < W * nvb; q += 128) {
        const int v = q / W, r = q - v * W, j = i0 - kd + r;
        if (j >= 0 && j < n)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(xs + q)),
                         "l"(Ci + (size_t)(v0 + v) * n + j) : "memory");
        else xs[q] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    /* row i of A: entries A(i, i-kd+d), d = 0..2kd (zero outside the matrix) */
    double a[2 * (KD > 0 ? KD : 1) + 1];
    double *band = btd_sm + BSP_BTD_VECS * W;
    if (KD > 0) {
#pragma unroll
        for (int d = 0; d < 2 * KD + 1; ++d) {
            const int j = i - KD + d;
            a[d] = (i < n && j >= 0 && j < n) ? __ldg(AB + (size_t)j * ld + (KD + i - j)) : 0.0;
        }
    } else {
        for (int d = 0; d < ld; ++d) {
            const int j = i - kd + d;
            band[d * 128 + threadIdx.x] = (i < n && j >= 0 && j < n) ? __ldg(AB + (size_t)j * ld + (kd + i - j)) : 0.0;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (i >= n) return;
    for (int v = 0; v < nvb; ++v) {
        const double *x = xs + v * W + threadIdx.x;
        double s = 0.0;
        if (KD > 0) {
#pragma unroll
            for (int d = 0; d < 2 * KD + 1; ++d) s = fma(a[d], x[d], s);
        } else {
            for (int d = 0; d < ld; ++d) s = fma(band[d * 128 + threadIdx.x], x[d], s);
        }
        Y[(size_t)(v0 + v) * n + i] = s;
    }
}

/* launch helper: batch blocks of vectors sharing the operator; the tiles of a block must fit shared memory */
static inline cudaError_t bsp_launch_band_times_dense(cudaStream_t st, int n, int kd, const double *AB, int nv, const double *Ci,
                                                      double *Y, int batch = 1, long long strideC = -1, long long strideY = -1)
{
    const bool fixed = kd >= 2 && kd <= 9;
    const size_t smem = ((size_t)BSP_BTD_VECS * (128 + 2 * kd) + (fixed ? 0 : (size_t)128 * (2 * kd + 1))) * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;          /* half bandwidth beyond ~90: not a banded operator */
    dim3 grid((n + 127) / 128, (nv + BSP_BTD_VECS - 1) / BSP_BTD_VECS, batch);
#define BSP_BTD_CASE(KD_)                                                                                                   \
    case KD_: bsp_band_times_dense_kernel<KD_><<<grid, 128, smem, st>>>(n, kd, AB, nv, Ci, Y, strideC, strideY); break;
    switch (fixed ? kd : 0) {
        BSP_BTD_CASE(2) BSP_BTD_CASE(3) BSP_BTD_CASE(4) BSP_BTD_CASE(5) BSP_BTD_CASE(6) BSP_BTD_CASE(7) BSP_BTD_CASE(8) BSP_BTD_CASE(9)
    default:
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(bsp_band_times_dense_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        bsp_band_times_dense_kernel<0><<<grid, 128, smem, st>>>(n, kd, AB, nv, Ci, Y, strideC, strideY);
    }
#undef BSP_BTD_CASE
    return cudaGetLastError();
}

/* m16n8k8 f64: A 16x8 (row), B 8x8 (col), C 16x8.  Fragments (g = lane>>2, t = lane&3):
 *   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (t, g) b1 (t+4, g);
 *   c0,c1 (g, 2t..2t+1)  c2,c3 (g+8, 2t..2t+1)
 * ptxas lowers it to four DMMA.8x8x4 (the only FP64 MMA shape of sm_100a: profiles/micro/dmma_peak.cu -- 37.1 TFLOP/s
 * issued as m16n8k8 against 29.7 as long runs of bare m8n8k4 with many accumulators). */
__device__ __forceinline__ void bsp_dmma_m16n8k8(double (&c)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

/*
 * D(M x N, column-major, ldd) = A^T B with A: K x M (column-major, lda), B: K x N (column-major, ldb); i.e.
 * D(m,n) = sum_k A(k,m) B(k,n): both operands are contiguous along k ("TN").  Batched over blockIdx.z with
 * constant strides.
 *
 * CTA tile BM x BN x 16, WARPS_M x WARPS_N warps, warp tile (BM/WARPS_M) x (BN/WARPS_N) in m16n8k8 MMAs.  Operand
 * tiles [row][16 + 4] (row stride 20 doubles: the fragment loads -- 8 rows x 4 k per half warp -- are bank-conflict
 * free) go global -> shared with cp.async (16 bytes, zero fill past M / N / K), STAGES tiles in flight: no
 * register staging, one barrier per k-tile.  Round 1 staged through registers with two buffers: 21.9 TFLOP/s on the
 * cfg5 chain, 11.8 on a single pair (cuBLAS on the box: 33.6 / 26.4; raw DMMA issue rate 37.1).
 */
#define BSP_GK 16
#define BSP_GLD 20

__device__ __forceinline__ void bsp_cp_async16(void *dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void bsp_cp_async8(void *dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool ALIGNED16>
__global__ void __launch_bounds__(32 * WARPS_M * WARPS_N) bsp_dgemm_tn_kernel(int M, int N, int K, const double *__restrict__ A, int lda,
                                                                               const double *__restrict__ Bm, int ldb,
                                                                               double *__restrict__ D, int ldd,
                                                                               long long strideA, long long strideB, long long strideD)
{
    constexpr int NT = 32 * WARPS_M * WARPS_N;
    constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
    constexpr int MI = WTM / 16, NI = WTN / 8;
    static_assert(WTM % 16 == 0 && WTN % 8 == 0, "warp tile in m16n8 MMAs");
    constexpr int STAGE_DOUBLES = (BM + BN) * BSP_GLD;
    extern __shared__ __align__(16) double gsm[];
    A += (size_t)blockIdx.z * strideA;
    Bm += (size_t)blockIdx.z * strideB;
    D += (size_t)blockIdx.z * strideD;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp % WARPS_M) * WTM, wn = (warp / WARPS_M) * WTN;
    const int g = lane >> 2, tg = lane & 3;

    /* one k-tile of both operands: rows 0..BM-1 of the A tile, then BN rows of the B tile */
    auto load_stage = [&](int slot, int kt) {
        double *sA = gsm + (size_t)slot * STAGE_DOUBLES;
        const int k0 = kt * BSP_GK;
        constexpr int CH = ALIGNED16 ? 2 : 1;                   /* doubles per asynchronous copy */
        constexpr int CPR = BSP_GK / CH;                        /* copies per row */
        for (int c = tid; c < (BM + BN) * CPR; c += NT) {
            const int row = c / CPR, kk = (c % CPR) * CH;
            const bool isA = row < BM;
            const int r = isA ? m0 + row : n0 + row - BM;
            const int lim = isA ? M : N;
            const double *src = isA ? A + (size_t)(r < lim ? r : 0) * lda : Bm + (size_t)(r < lim ? r : 0) * ldb;
            int nvalid = (r < lim) ? K - (k0 + kk) : 0;         /* doubles of this copy inside the operand */
            nvalid = nvalid < 0 ? 0 : (nvalid > CH ? CH : nvalid);
            const double *sp = src + (nvalid > 0 ? k0 + kk : 0);
            double *dst = sA + (size_t)row * BSP_GLD + kk;
            if (ALIGNED16) bsp_cp_async16(dst, sp, 8 * nvalid);
            else bsp_cp_async8(dst, sp, 8 * nvalid);
        }
    };
    double acc[MI][NI][4];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.0;

    const int nk = (K + BSP_GK - 1) / BSP_GK;
#pragma unroll
    for (int s_ = 0; s_ < STAGES - 1; ++s_) {
        if (s_ < nk) load_stage(s_, s_);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();     /* tile kt has landed for everybody; the slot of tile kt-1 is free */
        if (kt + STAGES - 1 < nk) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double *sA = gsm + (size_t)(kt % STAGES) * STAGE_DOUBLES;
        const double *sB = sA + (size_t)BM * BSP_GLD;
#pragma unroll
        for (int kk = 0; kk < BSP_GK; kk += 8) {
            double af[MI][4], bf[NI][2];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const double *pa = sA + (size_t)(wm + i * 16 + g) * BSP_GLD + kk + tg;
                af[i][0] = pa[0];
                af[i][1] = pa[8 * BSP_GLD];
                af[i][2] = pa[4];
                af[i][3] = pa[8 * BSP_GLD + 4];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const double *pb = sB + (size_t)(wn + j * 8 + g) * BSP_GLD + kk + tg;
                bf[j][0] = pb[0];
                bf[j][1] = pb[4];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) bsp_dmma_m16n8k8(acc[i][j], af[i], bf[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int m = m0 + wm + i * 16 + g + hh * 8;
                const int nn = n0 + wn + j * 8 + tg * 2;
                if (m < M) {
                    if (nn < N) D[(size_t)nn * ldd + m] = acc[i][j][hh * 2 + 0];
                    if (nn + 1 < N) D[(size_t)(nn + 1) * ldd + m] = acc[i][j][hh * 2 + 1];
                }
            }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES>
static inline cudaError_t bsp_launch_dgemm_variant(cudaStream_t st, bool aligned, int M, int N, int K, const double *A, int lda,
                                                   const double *Bm, int ldb, double *D, int ldd, int batch, long long strideA,
                                                   long long strideB, long long strideD)
{
    const int smem = STAGES * (BM + BN) * BSP_GLD * (int)sizeof(double);
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch);
    cudaError_t e;
    if (aligned) {
        auto kern = bsp_dgemm_tn_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, 32 * WARPS_M * WARPS_N, smem, st>>>(M, N, K, A, lda, Bm, ldb, D, ldd, strideA, strideB, strideD);
    } else {
        auto kern = bsp_dgemm_tn_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, 32 * WARPS_M * WARPS_N, smem, st>>>(M, N, K, A, lda, Bm, ldb, D, ldd, strideA, strideB, strideD);
    }
    return cudaGetLastError();
}

/* picks the tile variant by how many CTAs fill the 148 SMs; returns the CUDA status of the launch.
 * Variants (bspatom_set_option "gemm_variant" forces one for A/B runs; 0 = automatic):
 *   1: 128 x 128, 8 warps of 32 x 64        2: 128 x 64, 8 warps of 32 x 32, two CTAs per SM
 *   3: 64 x 64, 4 warps of 32 x 32          4: 128 x 128, 16 warps of 32 x 32 */
static int bsp_gemm_force = 0;
static inline cudaError_t bsp_launch_dgemm_tn(cudaStream_t st, int M, int N, int K, const double *A, int lda,
                                              const double *Bm, int ldb, double *D, int ldd, int batch,
                                              long long strideA, long long strideB, long long strideD)
{
    const bool aligned = ((lda | ldb) % 2 == 0) && ((strideA | strideB) % 2 == 0) && (((size_t)A | (size_t)Bm) % 16 == 0);
    auto ctas = [&](int bm, int bn) { return (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn) * batch; };
    int pick = bsp_gemm_force;
    /* measured on the cfg5 shapes (profiles/gemm_variants_r2.jsonl): two 128 x 64 CTAs per SM beat one 128 x 128 CTA
     * (24.6 against 21.2 / 22.8 TFLOP/s over the chain of 50); a single 1000^3 product is best on 64 x 64 tiles */
    if (!pick) pick = (M >= 128 && N >= 64 && ctas(128, 64) >= 2 * 148) ? 2 : 3;
    if (pick == 1) return bsp_launch_dgemm_variant<128, 128, 4, 2, 3>(st, aligned, M, N, K, A, lda, Bm, ldb, D, ldd, batch, strideA, strideB, strideD);
    if (pick == 2) return bsp_launch_dgemm_variant<128, 64, 4, 2, 3>(st, aligned, M, N, K, A, lda, Bm, ldb, D, ldd, batch, strideA, strideB, strideD);
    if (pick == 4) return bsp_launch_dgemm_variant<128, 128, 4, 4, 3>(st, aligned, M, N, K, A, lda, Bm, ldb, D, ldd, batch, strideA, strideB, strideD);
    return bsp_launch_dgemm_variant<64, 64, 2, 2, 4>(st, aligned, M, N, K, A, lda, Bm, ldb, D, ldd, batch, strideA, strideB, strideD);
}

#endif
