// Micro-benchmark: host-visible latency of "launch a kernel, kernel stores a flag in pinned memory, host polls"
// for different parameter sizes / cluster attributes / store flavours. Prints median microseconds.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
struct Big { double x[1900]; };   // ~15 KB
struct Small { double x[8]; };
template <class PARAM, int MODE>
__global__ void k(const __grid_constant__ PARAM p, unsigned long long *flag, double *data, unsigned long long seq) {
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        if (MODE >= 1 && threadIdx.x < 212) for (int i = 0; i < 10; i++) data[i * 212 + threadIdx.x] = p.x[threadIdx.x & 7] + seq;
        if (MODE == 3) return;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (MODE == 2) *(volatile unsigned long long *) flag = seq;
            else asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
        }
    }
}
template <class PARAM, int MODE>
double run(const char *name, int cluster, dim3 grid, int threads, unsigned long long *flag, double *data, cudaStream_t s, bool sync) {
    PARAM p; memset(&p, 0, sizeof(p));
    std::vector<double> t, tl;
    static unsigned long long seq = 0;
    for (int it = 0; it < 300; it++) {
        ++seq;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = cluster > 0 ? 1 : 0;
        if (MODE == 3) for (int i = 0; i < 2120; i++) ((volatile unsigned long long *) data)[i] = 0xFFF8DEADBEEFCAFEull;
        auto t0 = std::chrono::steady_clock::now();
        cudaLaunchKernelEx(&cfg, k<PARAM, MODE>, p, flag, data, seq);
        auto t1 = std::chrono::steady_clock::now();
        if (sync) cudaStreamSynchronize(s);
        else if (MODE == 3) {
            for (int i = 0; i < 2120; i++)
                while (__atomic_load_n((unsigned long long *) data + i, __ATOMIC_ACQUIRE) == 0xFFF8DEADBEEFCAFEull) __builtin_ia32_pause();
        } else while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) != seq) __builtin_ia32_pause();
        auto t2 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(s);
        if (it >= 50) { t.push_back(std::chrono::duration<double, std::micro>(t2 - t0).count()); tl.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count()); }
    }
    std::sort(t.begin(), t.end()); std::sort(tl.begin(), tl.end());
    printf("  %-58s launch API %.2f us, launch->visible %.2f us\n", name, tl[tl.size() / 2], t[t.size() / 2]);
    return t[t.size() / 2];
}
int main() {
    unsigned long long *flag; double *data;
    cudaMallocHost(&flag, 64); cudaMallocHost(&data, 2120 * 8); *flag = 0;
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    run<Small, 0>("small params, 1 CTA, release flag, poll", 0, dim3(1), 128, flag, data, s, false);
    run<Small, 0>("small params, 1 CTA, release flag, streamSync", 0, dim3(1), 128, flag, data, s, true);
    run<Small, 2>("small params, 1 CTA, volatile flag, poll", 0, dim3(1), 128, flag, data, s, false);
    run<Big, 0>("15 KB params, 1 CTA, release flag, poll", 0, dim3(1), 128, flag, data, s, false);
    run<Big, 0>("15 KB params, 80 CTAs cluster 8, release flag, poll", 8, dim3(8, 10), 128, flag, data, s, false);
    run<Small, 0>("small params, 80 CTAs cluster 8, release flag, poll", 8, dim3(8, 10), 128, flag, data, s, false);
    run<Small, 0>("small params, 80 CTAs no cluster, release flag, poll", 0, dim3(8, 10), 128, flag, data, s, false);
    run<Big, 1>("15 KB params, 80 CTAs cluster 8, 17 KB data + release, poll", 8, dim3(8, 10), 256, flag, data, s, false);
    run<Big, 1>("15 KB params, 80 CTAs cluster 8, 17 KB data, streamSync", 8, dim3(8, 10), 256, flag, data, s, true);
    run<Small, 1>("small params, 80 CTAs cluster 8, 17 KB data + release, poll", 8, dim3(8, 10), 256, flag, data, s, false);
    run<Small, 3>("small params, 80 CTAs cluster 8, 17 KB data, poll every word (no flag)", 8, dim3(8, 10), 256, flag, data, s, false);
    run<Small, 1>("small params, 80 CTAs cluster 8, 17 KB data + release, poll", 8, dim3(8, 10), 256, flag, data, s, false);
    run<Small, 3>("small params, 80 CTAs cluster 8, 17 KB data, poll every word (no flag)", 8, dim3(8, 10), 256, flag, data, s, false);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
