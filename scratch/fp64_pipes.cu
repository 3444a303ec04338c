// Are the FP64 ALU pipe (DADD / DFMA) and the FP64 MMA path (mma.sync.m8n8k4.f64, SASS DMMA) separate units on B200?
// Three kernels with the same launch shape: DADD chains only, DMMA only, and both interleaved (the same amounts of each).
// If the mixed kernel takes max(t_add, t_mma) the units are separate; if it takes t_add + t_mma they are one.
#include <cstdio>
#include <cuda_runtime.h>
template <int NADD, int NMMA>
__global__ void k(double *out, int iters, double a, double b)
{
    double acc[16], c[8][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
    const double af = a + threadIdx.x * 1e-9, bf = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < NMMA; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(af), "d"(bf));
#pragma unroll
            for (int i = 0; i < NADD; ++i) acc[i] = acc[i] + b;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NADD, int NMMA>
float run(double *out, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<NADD, NMMA><<<148 * 4, 512>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    return ms;
}
int main()
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 4 * 512);
    const int iters = 10000;
    const float t_add = run<16, 0>(out, iters), t_mma = run<0, 8>(out, iters), t_both = run<16, 8>(out, iters);
    // per iteration and warp: 2 x 16 DADD (32 lanes each) and 2 x 8 DMMA (256 FMA each)
    const double adds = 2.0 * 16 * 32 * iters * (148.0 * 4 * 16), fmas = 2.0 * 8 * 256 * iters * (148.0 * 4 * 16);
    printf("DADD only  %.3f ms  %.2f T adds/s\n", t_add, adds / t_add * 1e-9);
    printf("DMMA only  %.3f ms  %.2f TFLOP/s\n", t_mma, 2 * fmas / t_mma * 1e-9);
    printf("both       %.3f ms  (sum %.3f, max %.3f)\n", t_both, t_add + t_mma, t_add > t_mma ? t_add : t_mma);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
