// h2d_probe.cu — does a resident kernel that polls pinned host memory slow the host-to-device DMA down?
// Times 5 x 4.9 MB page-locked copies (one keyframe interval of raw AT128 records) back to back on a copy stream
//   (a) alone, (b) while P warps poll DISTINCT pinned 512-byte regions (one 16-byte load per lane, like the resident
//   evaluation kernel's per-pair mailboxes), (c) polling one 16-byte slot per warp, (d) polling with a pause between polls.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/h2d_probe tools/h2d_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); std::exit(1); } } while (0)

__global__ void poller(const ulonglong2 *mail, volatile int *stop, int lanes, int pause_ns, unsigned long long *polls) {
    const int lane = threadIdx.x & 31;
    const ulonglong2 *mine = mail + (size_t) blockIdx.x * 32;
    unsigned long long n = 0, acc = 0;
    while (!*stop) {
        if (lane < lanes) {
            ulonglong2 v;
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(mine + lane) : "memory");
            acc += v.x + v.y;
        }
        n++;
        if (pause_ns) __nanosleep(pause_ns);
    }
    if (threadIdx.x == 0) polls[blockIdx.x] = n + (acc & 1);
}

// A message pushed by a small host-to-device copy on its own stream, picked up by a kernel that polls DEVICE memory and
// answers with a posted store into pinned host memory: round trip on an idle link and while big copies saturate it —
// against the same round trip with the kernel polling the pinned host word itself.
__global__ void echo_kernel(const volatile unsigned long long *msg, volatile unsigned long long *ack, volatile int *stop) {
    unsigned long long last = 0;
    while (!*stop) {
        const unsigned long long v = *msg;
        if (v != last) {
            last = v;
            *ack = v;
        }
    }
}

static void message_probe(char *h, char *d, size_t bytes, cudaStream_t cs) {
    unsigned long long *h_msg, *h_ack, *d_msg;
    int *stop;
    CK(cudaMallocHost(&h_msg, 4096));
    CK(cudaMallocHost(&h_ack, 64));
    CK(cudaMallocHost(&stop, 64));
    CK(cudaMalloc(&d_msg, 4096));
    CK(cudaMemset(d_msg, 0, 4096));
    cudaStream_t ms, ks;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ms, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking));
    for (int via_copy = 0; via_copy < 2; via_copy++)
        for (int load = 0; load < 2; load++) {
            *stop  = 0;
            *h_ack = 0;
            *h_msg = 0;
            echo_kernel<<<1, 1, 0, ks>>>(via_copy ? d_msg : h_msg, h_ack, stop);
            if (load)
                for (int k = 0; k < 60; k++) CK(cudaMemcpyAsync(d + (k % 5) * bytes, h + (k % 5) * bytes, bytes, cudaMemcpyHostToDevice, cs));
            double sum = 0, worst = 0;
            const int reps = 200;
            for (int r = 1; r <= reps; r++) {
                const auto t0 = std::chrono::steady_clock::now();
                *h_msg = (unsigned long long) r;
                if (via_copy) CK(cudaMemcpyAsync(d_msg, h_msg, 1024, cudaMemcpyHostToDevice, ms));
                while (*(volatile unsigned long long *) h_ack != (unsigned long long) r) {}
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                sum += us;
                worst = us > worst ? us : worst;
            }
            *stop = 1;
            CK(cudaStreamSynchronize(ks));
            CK(cudaStreamSynchronize(cs));
            std::printf("message round trip, %-28s %-22s mean %.2f us, worst %.2f us\n", via_copy ? "1 KB copy + device poll:" : "kernel polls pinned host:",
                        load ? "link saturated by DMA" : "idle link", sum / reps, worst);
        }
}

int main() {
    const size_t bytes = 153600 * 32;
    const int copies = 5, reps = 40;
    char *h, *d;
    CK(cudaMallocHost(&h, bytes * copies));
    CK(cudaMalloc(&d, bytes * copies));
    ulonglong2 *mail;
    int *stop;
    unsigned long long *polls;
    CK(cudaMallocHost(&mail, 64 * 32 * 16));
    CK(cudaMallocHost(&stop, 64));
    CK(cudaMallocManaged(&polls, 64 * 8));
    cudaStream_t cs, ks;
    CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct Cfg { int warps, lanes, pause; const char *name; };
    const Cfg cfgs[] = {{0, 0, 0, "alone"}, {10, 32, 0, "10 warps x 512 B, back to back"}, {10, 1, 0, "10 warps x 16 B, back to back"},
                        {10, 32, 1000, "10 warps x 512 B, 1 us pause"}, {10, 32, 4000, "10 warps x 512 B, 4 us pause"},
                        {1, 32, 0, "1 warp x 512 B, back to back"}, {15, 32, 0, "15 warps x 512 B, back to back"}};
    for (const Cfg &c : cfgs) {
        *stop = 0;
        if (c.warps) poller<<<c.warps, 32, 0, ks>>>(mail, stop, c.lanes, c.pause, polls);
        float best = 1e9f, sum = 0;
        for (int r = 0; r < reps; r++) {
            CK(cudaEventRecord(e0, cs));
            for (int k = 0; k < copies; k++) CK(cudaMemcpyAsync(d + k * bytes, h + k * bytes, bytes, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(e1, cs));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
            sum += ms;
        }
        *stop = 1;
        CK(cudaStreamSynchronize(ks));
        std::printf("%-36s 5 x 4.9 MB: best %.1f us (%.1f GB/s), mean %.1f us (%.1f GB/s)%s\n", c.name, best * 1e3, bytes * copies / best / 1e6,
                    sum / reps * 1e3, bytes * copies / (sum / reps) / 1e6, "");
    }
    message_probe(h, d, bytes, cs);
    // the decimated points of one frame (every 80th 32-byte record, fetched as the 64-byte pair it belongs to) as a strided
    // copy on a second stream: completion time on an idle link and while big copies saturate the first stream
    {
        cudaStream_t s2;
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi));
        char *d2;
        CK(cudaMalloc(&d2, 1920 * 64));
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        for (int load = 0; load < 2; load++) {
            double sum = 0, worst = 0, wall = 0;
            const int reps = 40;
            for (int r = 0; r < reps; r++) {
                if (load)
                    for (int k = 0; k < 5; k++) CK(cudaMemcpyAsync(d + k * bytes, h + k * bytes, bytes, cudaMemcpyHostToDevice, cs));
                const auto t0 = std::chrono::steady_clock::now();
                CK(cudaEventRecord(a, s2));
                CK(cudaMemcpy2DAsync(d2, 64, h + 4 * bytes, 80 * 32, 64, 1920, cudaMemcpyHostToDevice, s2));
                CK(cudaEventRecord(b, s2));
                CK(cudaEventSynchronize(b));
                wall += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                sum += ms * 1e3;
                worst = ms * 1e3 > worst ? ms * 1e3 : worst;
                CK(cudaStreamSynchronize(cs));
            }
            std::printf("strided copy of 1920 x 64 B (pitch 2560 B) on a second stream, %-22s device %.1f us (worst %.1f), host wall %.1f us\n",
                        load ? "link saturated by DMA:" : "idle link:", sum / reps, worst, wall / reps);
        }
    }
    return 0;
}
