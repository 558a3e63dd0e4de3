// Micro-benchmark: FP64 FMA pipe vs FP64 tensor-core (mma.sync m8n8k4) issue rate per SM on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double *out, int iters, double x) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, 1e-9);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("  dfma: %d warps/SM: %.2f cycles per warp-DFMA per SM\n", blockDim.x / 32, double(t1 - t0) / (16.0 * iters * (blockDim.x / 32)));
}
__global__ void dmma_kernel(double *out, int iters, double x) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3;
    double a = x, b = 1.0 + 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("  dmma m8n8k4: %d warps/SM: %.2f cycles per warp-DMMA per SM (256 FMA each)\n", blockDim.x / 32, double(t1 - t0) / (4.0 * iters * (blockDim.x / 32)));
}
__global__ void ffma_kernel(float *out, int iters, float x) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, 1e-9f);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("  ffma: %d warps/SM: %.2f cycles per warp-FFMA per SM\n", blockDim.x / 32, double(t1 - t0) / (16.0 * iters * (blockDim.x / 32)));
}
int main() {
    double *d; cudaMalloc(&d, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 16, 32}) { dfma_kernel<<<148, w * 32>>>(d, 2000, 0.999); cudaDeviceSynchronize(); }
    for (int w : {4, 8, 16, 32}) { dmma_kernel<<<148, w * 32>>>(d, 2000, 0.999); cudaDeviceSynchronize(); }
    for (int w : {4, 8, 16, 32}) { ffma_kernel<<<148, w * 32>>>((float *) d, 2000, 0.999f); cudaDeviceSynchronize(); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
