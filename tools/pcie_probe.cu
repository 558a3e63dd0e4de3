// pcie_probe.cu — latency floor of a host <-> resident-kernel mailbox on this box (what bounds fflins_window_eval
// once the kernel launch is gone). Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pcie_probe tools/pcie_probe.cu
//   1. GPU -> pinned host read round trip (ld.volatile), ns (globaltimer) and cycles
//   2. ping-pong: host stores seq in pinned memory, a spinning kernel sees it and answers in another pinned word;
//      host-measured round trip (= pickup + reply), for 1 poller and for `n` pollers polling `k` 16-byte slots each
//   3. L2 round trip (ld.acquire.gpu), __threadfence, cluster barrier, DSMEM load — the hops of an on-device hand-off
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e = (x);                                                                    \
        if (e != cudaSuccess) {                                                                 \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                        \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void read_rtt(const unsigned long long *host, unsigned long long *out, int iters) {
    unsigned long long acc = 0, best = ~0ull, sum = 0;
    long long cbest = 1ll << 60;
    for (int i = 0; i < iters; i++) {
        unsigned long long t0 = gtime();
        long long c0          = clock64();
        unsigned long long v;
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(host + (i & 7)) : "memory");
        acc += v;
        long long c1          = clock64();
        unsigned long long t1 = gtime();
        best = min(best, t1 - t0);
        cbest = min(cbest, c1 - c0);
        sum += t1 - t0;
    }
    out[0] = best;
    out[1] = sum / iters;
    out[2] = (unsigned long long) cbest;
    out[3] = acc;
}

// nblocks x nthreads pollers; every thread polls `slots` 16-byte slots; block 0 thread 0 answers.
__global__ void pingpong(const ulonglong2 *mail, unsigned long long *reply, int slots, unsigned long long rounds,
                         unsigned long long timeout_ns) {
    __shared__ int s_ok;
    unsigned long long expect = 1;
    unsigned long long t_last = gtime();
    while (expect <= rounds) {
        bool ok = true;
        for (int s = 0; s < slots; s++) {
            unsigned long long w, t;
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];"
                         : "=l"(w), "=l"(t)
                         : "l"(mail + (threadIdx.x * slots + s) % 256)
                         : "memory");
            ok &= blockIdx.x == 0 ? t == expect : t >= expect;   // (only block 0 answers; the others are load)
            if (blockIdx.x != 0 && t >= expect) expect = t;
        }
        if (!__syncthreads_and(ok)) {
            if (threadIdx.x == 0) s_ok = gtime() - t_last > timeout_ns;
            __syncthreads();
            if (s_ok) return;
            continue;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(reply), "l"(expect) : "memory");
        expect++;
        t_last = gtime();
    }
}

__global__ void l2_hops(unsigned long long *g, unsigned long long *out, int iters) {
    // (a) ld.acquire.gpu round trip, (b) st + __threadfence, measured in cycles by one thread
    long long a = 0, b = 0;
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) {
        long long c0 = clock64();
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(g + 16 * (i & 15)) : "memory");
        acc += v;
        long long c1 = clock64();
        g[16 * (i & 15) + 1] = acc;
        __threadfence();
        long long c2 = clock64();
        a += c1 - c0;
        b += c2 - c1;
    }
    out[0] = a / iters;
    out[1] = b / iters;
    out[2] = acc;
}

__global__ void __cluster_dims__(8, 1, 1) cluster_hops(unsigned long long *out, int iters) {
    __shared__ double buf[256];
    cg::cluster_group cl = cg::this_cluster();
    buf[threadIdx.x]     = threadIdx.x;
    cl.sync();
    long long s = 0, d = 0;
    double acc  = 0;
    for (int i = 0; i < iters; i++) {
        long long c0 = clock64();
        cl.sync();
        long long c1 = clock64();
        const double *r = cl.map_shared_rank(buf, (cl.block_rank() + 1 + (i & 3)) & 7);
        acc += r[threadIdx.x];
        long long c2 = clock64();
        s += c1 - c0;
        d += c2 - c1;
    }
    cl.sync();
    if (cl.block_rank() == 0 && threadIdx.x == 0) {
        out[0] = s / iters;
        out[1] = d / iters;
        out[2] = (unsigned long long) acc;
    }
}

// every CTA polls its OWN region (distinct cache lines): do sysmem polls from several SMs still serialise?
__global__ void pingpong_own(const ulonglong2 *mail, unsigned long long *reply, int stride_slots, unsigned long long rounds,
                             unsigned long long timeout_ns) {
    __shared__ int s_ok;
    unsigned long long expect = 1;
    unsigned long long t_last = gtime();
    const ulonglong2 *mine = mail + (size_t) blockIdx.x * stride_slots + threadIdx.x;
    while (expect <= rounds) {
        unsigned long long w, t;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w), "=l"(t) : "l"(mine) : "memory");
        if (!__syncthreads_and(t == expect)) {
            if (threadIdx.x == 0) s_ok = gtime() - t_last > timeout_ns;
            __syncthreads();
            if (s_ok) return;
            continue;
        }
        // every CTA answers in its own word; the host waits for all of them
        if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(reply + 8 * blockIdx.x) = expect;
        expect++;
        t_last = gtime();
    }
}

// like ins_prep_kernel: thread i reads the 8 doubles of window entry i straight out of pinned host memory
__global__ void zc_read(const double *host, double *out, int w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w) return;
    double acc = host[i];
    for (int k = 0; k < 3; k++) acc += host[w + 3 * i + k];
    for (int k = 0; k < 4; k++) acc += host[4 * w + 4 * i + k];
    out[i] = acc;
}
// the resident kernel's dispatcher: one CTA polling 128 slots until *stop != 0 (or 2 s)
__global__ void poller(const ulonglong2 *mail, const int *stop, unsigned sleep_ns) {
    unsigned long long t0 = gtime();
    unsigned long long acc = 0;
    while (true) {
        unsigned long long w, t;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w), "=l"(t) : "l"(mail + threadIdx.x) : "memory");
        acc += t;
        int s = threadIdx.x == 0 ? *reinterpret_cast<const volatile int *>(stop) : 0;
        if (__syncthreads_or(s) || gtime() - t0 > 2000000000ull) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
    if (acc == 12345) printf("x");
}

int main() {
    unsigned long long *h = nullptr, *d_out = nullptr, *g = nullptr;
    CK(cudaMallocHost((void **) &h, 8192));
    CK(cudaMalloc((void **) &d_out, 256));
    CK(cudaMalloc((void **) &g, 4096));
    CK(cudaMemset(g, 0, 4096));
    for (int i = 0; i < 1024; i++) h[i] = 0;
    unsigned long long out[8];

    read_rtt<<<1, 1>>>(h, d_out, 2000);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d_out, 32, cudaMemcpyDeviceToHost));
    std::printf("gpu->host read RTT: min %llu ns, mean %llu ns, min %llu cycles\n", out[0], out[1], out[2]);

    l2_hops<<<1, 1>>>(g, d_out, 2000);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d_out, 24, cudaMemcpyDeviceToHost));
    std::printf("L2: ld.acquire.gpu %llu cycles, st + __threadfence %llu cycles\n", out[0], out[1]);

    cluster_hops<<<8, 128>>>(d_out, 2000);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d_out, 24, cudaMemcpyDeviceToHost));
    std::printf("cluster(8): barrier %llu cycles, DSMEM load %llu cycles\n", out[0], out[1]);

    {   // does a polling CTA slow down zero-copy reads of pinned memory by another kernel (the INS-window prologue)?
        const int w = 1000;
        double *hw = nullptr, *dout = nullptr;
        int *dstop = nullptr;
        CK(cudaMallocHost((void **) &hw, w * 64));
        for (int i = 0; i < w * 8; i++) hw[i] = i;
        CK(cudaMalloc((void **) &dout, w * 8));
        CK(cudaMalloc((void **) &dstop, 4));
        cudaStream_t s1, s2, s3;
        CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        const unsigned sleeps[] = {~0u, 0u, 500u, 2000u, 10000u};
        for (unsigned sl : sleeps) {
            CK(cudaMemset(dstop, 0, 4));
            if (sl != ~0u) poller<<<1, 128, 0, s2>>>(reinterpret_cast<ulonglong2 *>(h + 64), dstop, sl);
            float best = 1e9f, sum = 0;
            for (int r = 0; r < 30; r++) {
                cudaEventRecord(a, s1);
                zc_read<<<(w + 63) / 64, 64, 0, s1>>>(hw, dout, w);
                cudaEventRecord(b, s1);
                CK(cudaEventSynchronize(b));
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                best = std::min(best, ms);
                if (r >= 5) sum += ms;
            }
            int one = 1;
            CK(cudaMemcpyAsync(dstop, &one, 4, cudaMemcpyHostToDevice, s3));
            CK(cudaDeviceSynchronize());
            if (sl == ~0u) std::printf("zero-copy window read (64 KB), alone:            min %.1f us, mean %.1f us\n", best * 1e3, sum / 25 * 1e3);
            else std::printf("zero-copy window read, poller sleeping %5u ns:  min %.1f us, mean %.1f us\n", sl, best * 1e3, sum / 25 * 1e3);
        }
    }

    ulonglong2 *mail          = reinterpret_cast<ulonglong2 *>(h + 64);
    volatile unsigned long long *reply = h;
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const int cfgs[][3] = {{1, 1, 1}, {1, 128, 1}, {1, 128, 2}, {10, 32, 1}, {16, 32, 1}, {16, 128, 1}, {80, 32, 1}, {128, 32, 2}};
    for (auto &c : cfgs) {
        const int rounds = 3000;
        for (int i = 0; i < 256; i++) mail[i] = make_ulonglong2(0, 0);
        *reply = 0;
        pingpong<<<c[0], c[1], 0, s>>>(mail, h, c[2], rounds, 50000000ull);
        std::vector<double> us;
        for (int r = 1; r <= rounds; r++) {
            auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < 256; i++) {
                reinterpret_cast<volatile unsigned long long *>(mail)[2 * i] = r;
                __atomic_store_n(reinterpret_cast<unsigned long long *>(mail) + 2 * i + 1, (unsigned long long) r, __ATOMIC_RELEASE);
            }
            while (*reply != (unsigned long long) r) __builtin_ia32_pause();
            auto t1 = std::chrono::steady_clock::now();
            us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            for (volatile int k = 0; k < 300; k++) {}   // a little jitter between rounds
        }
        CK(cudaStreamSynchronize(s));
        std::sort(us.begin(), us.end());
        std::printf("ping-pong %3d CTAs x %3d threads x %d slots: p10 %.2f  p50 %.2f  p90 %.2f us (includes 256 slot stores)\n", c[0],
                    c[1], c[2], us[rounds / 10], us[rounds / 2], us[rounds * 9 / 10]);
    }
    {   // distinct regions per CTA, every CTA answers (no fence: a single 8-byte posted store)
        ulonglong2 *big = nullptr;
        unsigned long long *rep = nullptr;
        CK(cudaMallocHost((void **) &big, 128 * 64 * sizeof(ulonglong2)));
        CK(cudaMallocHost((void **) &rep, 128 * 64));
        const int cfg2[][2] = {{1, 32}, {10, 32}, {10, 64}, {16, 32}, {80, 32}};
        for (auto &c : cfg2) {
            const int rounds = 3000, nb = c[0], nt = c[1];
            for (int i = 0; i < 128 * 64; i++) big[i] = make_ulonglong2(0, 0);
            for (int i = 0; i < 128 * 8; i++) rep[i] = 0;
            pingpong_own<<<nb, nt, 0, s>>>(big, rep, 64, rounds, 50000000ull);
            std::vector<double> us;
            for (int r = 1; r <= rounds; r++) {
                auto t0 = std::chrono::steady_clock::now();
                for (int b = 0; b < nb; b++)
                    for (int i = 0; i < nt; i++) {
                        reinterpret_cast<volatile unsigned long long *>(big)[2 * (64 * b + i)] = r;
                        __atomic_store_n(reinterpret_cast<unsigned long long *>(big) + 2 * (64 * b + i) + 1, (unsigned long long) r,
                                         __ATOMIC_RELEASE);
                    }
                for (int b = 0; b < nb; b++)
                    while (*reinterpret_cast<volatile unsigned long long *>(rep + 8 * b) != (unsigned long long) r) __builtin_ia32_pause();
                auto t1 = std::chrono::steady_clock::now();
                us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
                for (volatile int k = 0; k < 300; k++) {}
            }
            CK(cudaStreamSynchronize(s));
            std::sort(us.begin(), us.end());
            std::printf("own-region ping-pong %3d CTAs x %3d threads (all answer, unfenced): p10 %.2f  p50 %.2f  p90 %.2f us\n", nb, nt,
                        us[rounds / 10], us[rounds / 2], us[rounds * 9 / 10]);
        }
    }
    return 0;
}
