// Micro-benchmark: FP64 throughput of DFMA vs mma.sync.m8n8k4.f64 (DMMA) on one GPU.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double *out, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
    double af = a + threadIdx.x * 1e-9, bf = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(af), "d"(bf));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            float ms;
            cudaEventRecord(e0);
            k_dfma<<<148 * 4, warps * 32>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 148 * 4 * warps * 32.0 * 16 * iters;
            if (rep) printf("DFMA warps/CTA=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
            cudaEventRecord(e0);
            k_dmma<<<148 * 4, warps * 32>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            fl = 2.0 * 148 * 4 * warps * 256.0 * 8 * iters;
            if (rep) printf("DMMA warps/CTA=%2d  %.3f ms  %.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
