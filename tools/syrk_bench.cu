// Micro-benchmark of the factor kernel's shared-memory SYRK stage in isolation (cycles per CTA).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kRowStride = 24;
template <int MODE>
__global__ void __launch_bounds__(256) syrk(double *out) {
    extern __shared__ __align__(16) double smem[];
    double (*rows)[kRowStride] = reinterpret_cast<double (*)[kRowStride]>(smem);
    const int tid = threadIdx.x;
    for (int i = tid; i < 256 * kRowStride; i += 256) smem[i] = 1e-3 * i;
    int slice, T;
    if (MODE == 2) { T = tid >> 4; slice = tid & 15; } else { slice = tid >> 4; T = tid & 15; }
    int bi = 0, bj = 0;
    {
        int t = T;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            int len = 5 - i;
            if (t >= 0 && t < len) { bi = i; bj = i + t; t = -1; }
            else if (t >= len) t -= len;
        }
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    __syncthreads();
    long long t0 = clock64();
    if (T < 15 || MODE == 1) {
#pragma unroll 4
        for (int kk = 0; kk < 16; kk++) {
            const double2 *row = reinterpret_cast<const double2 *>(rows[slice + 16 * kk]);
            double2 a01 = row[2 * bi], a23 = row[2 * bi + 1], b01 = row[2 * bj], b23 = row[2 * bj + 1];
            const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] += av[a] * bv[b];
        }
    } else {
#pragma unroll 4
        for (int kk = 0; kk < 16; kk++) {
            acc[0][0] += rows[slice + 16 * kk][20];
            acc[0][1] += rows[slice + 16 * kk][21];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    long long t2 = clock64();
    double s = 0;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) s += acc[a][b];
    out[blockIdx.x * 256 + tid] = s;
    if (tid == 0 && blockIdx.x == 0) printf("  mode %d: warp0 loop %lld cycles, to barrier %lld\n", MODE, t1 - t0, t2 - t0);
}
int main() {
    double *d; cudaMalloc(&d, 148 * 256 * sizeof(double));
    size_t sm = 256 * kRowStride * sizeof(double);
    cudaFuncSetAttribute(syrk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
    cudaFuncSetAttribute(syrk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
    cudaFuncSetAttribute(syrk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
    for (int r = 0; r < 2; r++) {
        syrk<0><<<40, 256, sm>>>(d); cudaDeviceSynchronize();
        syrk<1><<<40, 256, sm>>>(d); cudaDeviceSynchronize();
        syrk<2><<<40, 256, sm>>>(d); cudaDeviceSynchronize();
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
