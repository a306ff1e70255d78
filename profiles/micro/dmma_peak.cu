// Raw FP64 MMA issue rate on sm_100a per instruction shape (registers only): which mma.sync f64 shape the
// contraction kernel should be built on, and what the FP64 pipe gives with plain DFMA.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/dmma_peak profiles/micro/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int ILP>
__global__ void __launch_bounds__(256) k_mma(double *out, int iters)
{
    double c[ILP][4];
    double a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; b[i] = 1.0 - threadIdx.x * 1e-9 + i; }
    for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (SHAPE == 884)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1684)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 1688)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else if (SHAPE == 16816) {
                double a2[4] = {a[0] + 1, a[1] + 1, a[2] + 1, a[3] + 1};
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[j][0]), "+d"(c[j][1]), "+d"(c[j][2]), "+d"(c[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a2[0]), "d"(a2[1]), "d"(a2[2]), "d"(a2[3]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            } else {   // plain DFMA: 4 independent chains per j
#pragma unroll
                for (int i = 0; i < 4; ++i) c[j][i] = fma(a[i], b[i], c[j][i]);
            }
        }
    }
    double s = 0;
    for (int j = 0; j < ILP; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int ILP>
void run(const char *name, double flops_per_instr_per_warp)
{
    double *d;
    const int blocks = 148 * 4, iters = 20000;
    cudaMalloc(&d, blocks * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mma<SHAPE, ILP><<<blocks, 256>>>(d, 100);
    cudaEventRecord(e0);
    k_mma<SHAPE, ILP><<<blocks, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = blocks * 8.0;
    const double tf = warps * iters * ILP * flops_per_instr_per_warp / (ms * 1e-3) / 1e12;
    printf("{\"shape\": \"%s\", \"ilp\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", name, ILP, ms, tf);
    cudaFree(d);
}

int main()
{
    run<884, 8>("mma.m8n8k4.f64", 2.0 * 8 * 8 * 4);
    run<884, 16>("mma.m8n8k4.f64", 2.0 * 8 * 8 * 4);
    run<1684, 8>("mma.m16n8k4.f64", 2.0 * 16 * 8 * 4);
    run<1688, 8>("mma.m16n8k8.f64", 2.0 * 16 * 8 * 8);
    run<1688, 16>("mma.m16n8k8.f64", 2.0 * 16 * 8 * 8);
    run<16816, 8>("mma.m16n8k16.f64", 2.0 * 16 * 8 * 16);
    run<0, 8>("dfma (4 chains x ilp)", 2.0 * 32 * 4);
    run<0, 16>("dfma (4 chains x ilp)", 2.0 * 32 * 4);
    return 0;
}
