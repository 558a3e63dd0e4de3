// k_voxel.cu — K2 voxel-grid downsample: bbox -> voxel key -> stable LSD radix sort -> segmented
// centroid, restating pcl::VoxelGrid<PointXYZI>::applyFilter as FF-LINS uses it
// (call sites ff_lins/lidar/lidar_map.cc:71,209-210,260-261,311-312; arithmetic restated in
// SURVEY.md Appendix A.1). Voxel keys and output sets are exact; the centroids are bit-exact with the
// oracle's STABLE-ORDER model of PCL (fp32 sums in ascending original index): PCL itself sorts with an
// unstable sort, so its intra-voxel summation order is implementation-defined.
//
// Bit-exactness rules:
//   - keys: int(floorf(x * inv_leaf) - float(min_b)) with a separately rounded fp32 multiply;
//   - output order: ascending key;
//   - centroid: fp32 accumulators walked in ASCENDING ORIGINAL INDEX (the sort is stable and
//     carries the index as payload), one lane per voxel, then divided by float(count). A tree
//     reduction would change the rounding, so the walk is sequential by construction.
//
// B200 mapping: every stage is a streaming pass over 8-byte (key, index) pairs; the digit passes
// that the realised key width does not need exit immediately (the width is decided on the device
// from the bounding box, the host never synchronises inside the chain).
#include "kernels.h"

#include <atomic>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

namespace fflins {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS   = 16;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS; // 4096
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int SEG_THREADS = 256;
constexpr int SEG_ITEMS   = 4;
constexpr int SEG_TILE    = SEG_THREADS * SEG_ITEMS;

// meta layout (ints)
enum { M_BBOX = 0, M_OVERFLOW = 6, M_NBITS = 7, M_NOUT = 8, M_PARITY = 9 /* 9..13 */, M_SIZE = 16 };

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void vox_init_kernel(int *meta) {
    int t = threadIdx.x;
    if (t < 3) meta[M_BBOX + t] = f2ord(FLT_MAX);
    else if (t < 6) meta[M_BBOX + t] = f2ord(-FLT_MAX);
    else if (t < M_SIZE) meta[t] = 0;
}

__global__ void __launch_bounds__(256) vox_bbox_kernel(const float4 *__restrict__ in, int n, int *__restrict__ meta) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(in + i);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    // warp shuffle -> shared memory -> 6 atomics per CTA
    __shared__ float s_mn[3][8], s_mx[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            s_mn[a][wid] = mn[a];
            s_mx[a][wid] = mx[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int a = threadIdx.x % 3;
        if (threadIdx.x < 3) {
            float v = s_mn[a][0];
#pragma unroll
            for (int w = 1; w < 8; w++) v = fminf(v, s_mn[a][w]);
            atomicMin(meta + M_BBOX + a, f2ord(v));
        } else {
            float v = s_mx[a][0];
#pragma unroll
            for (int w = 1; w < 8; w++) v = fmaxf(v, s_mx[a][w]);
            atomicMax(meta + M_BBOX + 3 + a, f2ord(v));
        }
    }
}

struct Grid3 {
    int min_b[3];
    int mul1, mul2;
    bool overflow;
    int nbits;
};

__device__ __forceinline__ Grid3 make_grid(const int *__restrict__ meta, float inv) {
    Grid3 g;
    float mn[3], mx[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        mn[a] = ord2f(meta[M_BBOX + a]);
        mx[a] = ord2f(meta[M_BBOX + 3 + a]);
    }
    long long d[3];
    int div_b[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        d[a]       = (long long) (__fmul_rn(__fsub_rn(mx[a], mn[a]), inv)) + 1;
        g.min_b[a] = (int) floorf(__fmul_rn(mn[a], inv));
        int max_b  = (int) floorf(__fmul_rn(mx[a], inv));
        div_b[a]   = max_b - g.min_b[a] + 1;
    }
    g.overflow = (d[0] * d[1] * d[2]) > 2147483647LL;
    g.mul1     = div_b[0];
    g.mul2     = div_b[0] * div_b[1];
    long long cells = (long long) div_b[0] * div_b[1] * div_b[2];
    int nb = 0;
    while (nb < 32 && (1LL << nb) < cells) nb++;
    g.nbits = g.overflow ? 0 : (nb < 1 ? 1 : nb);
    return g;
}

__global__ void __launch_bounds__(256)
vox_keys_kernel(const float4 *__restrict__ in, int n, float inv, int *__restrict__ meta, uint32_t *__restrict__ keys,
                int32_t *__restrict__ keys_out) {
    Grid3 g = make_grid(meta, inv);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        meta[M_OVERFLOW] = g.overflow ? 1 : 0;
        meta[M_NBITS]    = g.nbits;
        meta[M_PARITY]   = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(in + i);
        int i0   = (int) (floorf(__fmul_rn(p.x, inv)) - (float) g.min_b[0]);
        int i1   = (int) (floorf(__fmul_rn(p.y, inv)) - (float) g.min_b[1]);
        int i2   = (int) (floorf(__fmul_rn(p.z, inv)) - (float) g.min_b[2]);
        int key  = i0 + i1 * g.mul1 + i2 * g.mul2;
        keys[i]  = (uint32_t) key;
        if (keys_out) keys_out[i] = key;
    }
}

// ---- stable LSD radix sort, 8-bit digits ---------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t *__restrict__ k0, const uint32_t *__restrict__ k1, const int *__restrict__ meta, int pass,
               int n, uint32_t *__restrict__ block_hist, int nb) {
    if (pass * 8 >= meta[M_NBITS]) return;
    const uint32_t *keys = meta[M_PARITY + pass] ? k1 : k0;
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int shift = pass * 8;
    int base        = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; r++) {
        int i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    block_hist[threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x];
}

// Exclusive scan of the digit-major per-tile histogram (256 * nb entries), one CTA. Each warp owns a
// contiguous chunk and reads it with coalesced 128-byte requests that are all in flight together
// (values stay in registers), scans 32 values per shuffle round with a running carry, the 32 warp
// totals are scanned through shared memory and the offsets written back in one pass.
constexpr int SCAN_ITEMS = 16;                       // 32-value rounds per warp and super-chunk
constexpr int SCAN_SUPER = 32 * 32 * SCAN_ITEMS;     // 16384 entries per super-chunk
// exclusive scan of block_hist[lo, hi) in place by one 1024-thread CTA, starting from `carry0`
__device__ __forceinline__ void scan_range(uint32_t *__restrict__ block_hist, int lo, int hi, uint32_t carry0) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = carry0;
    __syncthreads();
    for (int sbase = lo; sbase < hi; sbase += SCAN_SUPER) {
        const int wbase = sbase + wid * (32 * SCAN_ITEMS);
        uint32_t v[SCAN_ITEMS];
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; r++) {
            int i = wbase + r * 32 + lane;
            v[r]  = (i < hi) ? block_hist[i] : 0u;
        }
        uint32_t run = 0; // running total of this warp's earlier rounds
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; r++) {
            uint32_t incl = v[r];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            v[r]         = run + incl - v[r]; // exclusive within the warp chunk
            run += tot;
        }
        if (lane == 0) warp_sums[wid] = run;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            warp_sums[lane] = wi - w; // exclusive offset of each warp inside the super-chunk
        }
        __syncthreads();
        const uint32_t base = s_carry + warp_sums[wid];
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; r++) {
            int i = wbase + r * 32 + lane;
            if (i < hi) block_hist[i] = base + v[r];
        }
        __syncthreads();
        // carry for the next super-chunk: offset of the last warp + its total
        if (threadIdx.x == 1023) s_carry = base + run;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
rs_scan_kernel(int *__restrict__ meta, int pass, uint32_t *__restrict__ block_hist, int total) {
    const bool active = pass * 8 < meta[M_NBITS];
    if (threadIdx.x == 0) meta[M_PARITY + pass + 1] = meta[M_PARITY + pass] ^ (active ? 1 : 0);
    if (!active) return;
    scan_range(block_hist, 0, total, 0u);
}

// Large inputs (more than ~1 M points): the histogram table no longer fits one CTA's patience. Three launches:
// per-chunk sums, a scan of the (<= 1024) chunk sums, and the chunk scans seeded with them.
__global__ void __launch_bounds__(1024)
rs_scan_reduce_kernel(const int *__restrict__ meta, int pass, const uint32_t *__restrict__ block_hist, int total, int chunk,
                      uint32_t *__restrict__ partial) {
    if (pass * 8 >= meta[M_NBITS]) return;
    const int lo = blockIdx.x * chunk, hi = min(total, lo + chunk);
    uint32_t sum = 0;
    for (int i = lo + threadIdx.x; i < hi; i += 1024) sum += block_hist[i];
    __shared__ uint32_t ws[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = ws[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
__global__ void __launch_bounds__(1024)
rs_scan_partials_kernel(int *__restrict__ meta, int pass, uint32_t *__restrict__ partial, int g) {
    const bool active = pass * 8 < meta[M_NBITS];
    if (threadIdx.x == 0) meta[M_PARITY + pass + 1] = meta[M_PARITY + pass] ^ (active ? 1 : 0);
    if (!active) return;
    __shared__ uint32_t ws[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t v = threadIdx.x < g ? partial[threadIdx.x] : 0u;
    uint32_t incl    = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = ws[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        ws[lane] = wi - w;
    }
    __syncthreads();
    if (threadIdx.x < g) partial[threadIdx.x] = ws[wid] + incl - v;
}
__global__ void __launch_bounds__(1024)
rs_scan_apply_kernel(const int *__restrict__ meta, int pass, uint32_t *__restrict__ block_hist, int total, int chunk,
                     const uint32_t *__restrict__ partial) {
    if (pass * 8 >= meta[M_NBITS]) return;
    const int lo = blockIdx.x * chunk;
    scan_range(block_hist, lo, min(total, lo + chunk), partial[blockIdx.x]);
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(uint32_t *k0, uint32_t *k1, uint32_t *x0, uint32_t *x1, const int *__restrict__ meta, int pass, int n,
                  const uint32_t *__restrict__ block_hist, int nb) {
    if (pass * 8 >= meta[M_NBITS]) return;
    // ping-pong buffers: no __restrict__ here, the in/out roles are picked at run time
    const int parity       = meta[M_PARITY + pass];
    const uint32_t *keys   = parity ? k1 : k0;
    const uint32_t *idx    = parity ? x1 : x0;
    uint32_t *keys_out     = parity ? k0 : k1;
    uint32_t *idx_out      = parity ? x0 : x1;
    __shared__ uint32_t warp_hist[RS_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const int shift = pass * 8;
    const int chunk = blockIdx.x * RS_TILE + wid * (RS_TILE / RS_WARPS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
    // all loads of the tile first (independent, in flight together), ranking afterwards
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int i  = chunk + r * 32 + lane;
        key[r] = (i < n) ? keys[i] : 0xffffffffu;
        val[r] = (i < n) ? (pass == 0 ? (uint32_t) i : idx[i]) : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int i          = chunk + r * 32 + lane;
        bool valid     = i < n;
        uint32_t digit = (key[r] >> shift) & 255u;
        unsigned m     = __match_any_sync(0xffffffffu, valid ? digit : 256u + lane);
        uint32_t prev  = warp_hist[wid][digit];
        __syncwarp();
        uint32_t rk = __popc(m & lt);
        if (valid && rk == 0) warp_hist[wid][digit] = prev + __popc(m);
        __syncwarp();
        rank[r] = prev + rk;
    }
    __syncthreads();
    {
        // exclusive prefix over the warps of this block, seeded with the block's global offset
        int d        = threadIdx.x;
        uint32_t run = block_hist[d * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c      = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        int i = chunk + r * 32 + lane;
        if (i < n) {
            uint32_t digit = (key[r] >> shift) & 255u;
            uint32_t dst   = warp_hist[wid][digit] + rank[r];
            keys_out[dst]  = key[r];
            idx_out[dst]   = val[r];
        }
    }
}

// ---- segment heads + ordered centroid ---------------------------------------------------------------
__device__ __forceinline__ bool is_head(const uint32_t *__restrict__ keys, int i) { return i == 0 || keys[i] != keys[i - 1]; }

__global__ void __launch_bounds__(SEG_THREADS)
seg_count_kernel(const uint32_t *__restrict__ k0, const uint32_t *__restrict__ k1, const int *__restrict__ meta, int n,
                 uint32_t *__restrict__ block_heads) {
    const uint32_t *keys = meta[M_PARITY + 4] ? k1 : k0;
    int base             = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    int c                = 0;
    if (!meta[M_OVERFLOW]) {
#pragma unroll
        for (int j = 0; j < SEG_ITEMS; j++) {
            int i = base + j;
            if (i < n && is_head(keys, i)) c++;
        }
    }
    __shared__ int ws[SEG_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < SEG_THREADS / 32; w++) s += ws[w];
        block_heads[blockIdx.x] = s;
    }
}

// Second pass over the head flags: every head learns its output slot (CTA base from the per-CTA
// counts + in-CTA scan) and records where its segment starts.
__global__ void __launch_bounds__(SEG_THREADS)
seg_starts_kernel(const uint32_t *__restrict__ k0, const uint32_t *__restrict__ k1, int *__restrict__ meta, int n,
                  const uint32_t *__restrict__ block_heads, uint32_t *__restrict__ seg_start, int *__restrict__ d_nout) {
    const uint32_t *keys = meta[M_PARITY + 4] ? k1 : k0;
    if (meta[M_OVERFLOW]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            *d_nout      = n;
            meta[M_NOUT] = n;
        }
        return;
    }
    __shared__ int s_base;
    __shared__ int ws[SEG_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid == 0) {
        int s = 0;
        for (int b = lane; b < (int) blockIdx.x; b += 32) s += block_heads[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) s_base = s;
    }
    const int base = blockIdx.x * SEG_TILE + threadIdx.x * SEG_ITEMS;
    bool head[SEG_ITEMS];
    int c = 0;
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; j++) {
        int i   = base + j;
        head[j] = (i < n) && is_head(keys, i);
        c += head[j] ? 1 : 0;
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; w++) woff += ws[w];
    int slot = s_base + woff + incl - c;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SEG_THREADS - 1) {
        int total        = s_base + woff + incl;
        *d_nout          = total;
        meta[M_NOUT]     = total;
        seg_start[total] = (uint32_t) n; // sentinel
    }
#pragma unroll
    for (int j = 0; j < SEG_ITEMS; j++)
        if (head[j]) seg_start[slot++] = (uint32_t) (base + j);
}

// One warp per voxel: the lanes gather 32 points of the segment at a time (coalesced index read,
// 16-byte point gathers in flight together) into shared memory, then the fp32 sums are accumulated
// in ascending original index — sequentially, because fp32 addition is not associative and the
// result must equal PCL's (Appendix A.1 step 7). Every lane runs the same chain; lane 0 stores.
constexpr int CEN_THREADS = 256;
constexpr int CEN_ROUND   = 128; // points staged per warp round
__global__ void __launch_bounds__(CEN_THREADS)
seg_centroid_kernel(const float4 *__restrict__ in, const uint32_t *__restrict__ x0, const uint32_t *__restrict__ x1,
                    const int *__restrict__ meta, int n, const uint32_t *__restrict__ seg_start, float4 *__restrict__ out) {
    if (meta[M_OVERFLOW]) {
        // PCL: "leaf size too small": the input cloud is returned unchanged
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
        return;
    }
    __shared__ float4 stage[CEN_THREADS / 32][CEN_ROUND];
    const uint32_t *idx = meta[M_PARITY + 4] ? x1 : x0;
    const int n_seg     = meta[M_NOUT];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps = gridDim.x * (CEN_THREADS / 32);
    for (int s = blockIdx.x * (CEN_THREADS / 32) + wid; s < n_seg; s += warps) {
        const int start = (int) seg_start[s], end = (int) seg_start[s + 1];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (int base = start; base < end; base += CEN_ROUND) {
            // up to 4 gathers per lane in flight before the ordered accumulation
#pragma unroll
            for (int c = 0; c < CEN_ROUND / 32; c++) {
                int e = base + c * 32 + lane;
                if (e < end) stage[wid][c * 32 + lane] = __ldg(in + idx[e]);
            }
            __syncwarp();
            const int cnt = min(CEN_ROUND, end - base);
#pragma unroll 4
            for (int j = 0; j < cnt; j++) {
                float4 p = stage[wid][j];
                sx += p.x;
                sy += p.y;
                sz += p.z;
                si += p.w;
            }
            __syncwarp();
        }
        if (lane == 0) {
            float cnt = (float) (end - start);
            out[s]    = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        }
    }
}

// ---- large clouds: the whole filter in ONE cooperative launch -------------------------------------------
// The multi-kernel chain above costs 18 launches (3 per digit pass plus a single-CTA scan each) for a keyframe map of
// 768 k points that the GPU streams in a few microseconds, i.e. it is bound by launch gaps and by the serial scans.
// vox_coop_kernel runs the same algorithm as a persistent grid of one 512-thread CTA per SM (half the register file,
// 20 KB of shared memory: the solver-facing kernels of the compute stream still fit next to it) with grid barriers
// between the phases:   bbox | keys + digit histogram | (rank + scatter | digit histogram)* | head count | centroids.
//   * CTA c owns the contiguous chunk [c S, (c+1) S) of the (current order of the) items in every phase, so a stable
//     pass only needs, per digit, the counts of the chunks before it: a [G][bins] table that every CTA sums for itself
//     after the barrier (no separate scan kernel, no look-back chain: with <= 148 chunks the table is 300 KB);
//   * digits are 9 bits wide where that saves a pass (18-bit keys of a 60 m scene: two passes instead of three);
//   * (key, index) travel as one 8-byte item: one scattered store per point and pass;
//   * the ordered centroid walks the sorted items contiguously: each warp owns a slice, finds the voxel heads from the
//     keys (no segment table), gathers 64 points per round with the next round's gathers and the indices of the round
//     after that already in flight, and lets four lanes run the x / y / z / intensity chains in ascending original index.
// Bit-exactness is unchanged: stable LSD order, sequential fp32 sums, IEEE division.
constexpr int COOP_THREADS = 512;
constexpr int COOP_WARPS   = COOP_THREADS / 32;
constexpr int COOP_ROUNDS  = 8;                           // 32-item rounds per warp and sub-tile (held in registers)
constexpr int COOP_TILE    = COOP_THREADS * COOP_ROUNDS;  // 4096 items per sub-tile
constexpr int COOP_BINS    = 512;
constexpr int COOP_MAXG    = 320;                         // chunks (CTAs) the work tables are sized for (two per SM on 148 SMs)
constexpr int COOP_CEN     = 64;                          // points per centroid round and warp
constexpr int COOP_STAMPS  = 20;

struct CoopSmem {
    union {
        uint16_t warp_hist[COOP_WARPS][COOP_BINS];        // rank + scatter
        float4 stage[COOP_WARPS][COOP_CEN];               // centroid
    };
    uint32_t hist[COOP_BINS];
    uint32_t offset[COOP_BINS];
    uint32_t tot[COOP_BINS];
    float mn[3][COOP_WARPS], mx[3][COOP_WARPS];
    int bbox[6];
    int ws[32];
    int cta_base, cta_total;
};

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target) {
    // every CTA of the (co-resident, cooperative) grid arrives once per barrier; the counter only grows during a launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v, polls = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            // every wait is bounded: a peer that never arrives (a faulted launch left the counters behind) must end in an
            // error the host sees, not in a hung device (~2^27 L2 polls: seconds, a barrier takes microseconds)
            if (++polls > (1u << 27)) asm volatile("trap;");
        } while ((int) (v - target) < 0);
    }
    __syncthreads();
}

struct CoopArgs {
    const float4 *in;
    int n;
    float inv;
    int *meta;
    uint2 *kv[2];         // (key, original index) items, ping-pong
    void *table;          // [G][COOP_BINS] chunk digit counts: uint16_t while a chunk holds < 65536 items, else uint32_t
    uint32_t *heads;      // [G] chunk head counts, [COOP_MAXG + G][6] chunk bounding boxes behind them
    unsigned *bar;        // [0] grid barrier counter, [1] exit counter: both zero between launches (the last CTA to leave resets them)
    float4 *out;
    int *d_nout;
    int32_t *keys_out;
    int *report;          // optional pinned host int: voxel count (zero-copy read-back)
    unsigned long long *stamps; // optional: globaltimer at the phase boundaries (CTA 0; [COOP_STAMPS - 1]: last CTA to leave)
};

__device__ __forceinline__ unsigned long long coop_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void coop_exit(const CoopArgs &a) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.bar + 1, 1u);
        if (t == gridDim.x - 1) { // everybody is behind the last barrier
            a.bar[0] = 0;
            a.bar[1] = 0;
            if (a.stamps) a.stamps[COOP_STAMPS - 1] = coop_now();
        }
    }
}
__device__ __forceinline__ void coop_stamp(const CoopArgs &a, int k) {
    if (a.stamps && blockIdx.x == 0 && threadIdx.x == 0) a.stamps[k] = coop_now();
}

template <typename TabT>
__global__ void __launch_bounds__(COOP_THREADS, 2) vox_coop_kernel(const __grid_constant__ CoopArgs a) {
    __shared__ CoopSmem S;
    TabT *table = static_cast<TabT *>(a.table);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, c = blockIdx.x, n = a.n;
    const int chunk = (((n + G - 1) / G) + 127) & ~127;    // multiple of 128: unrolled warp rounds stay aligned
    const int lo = min(n, c * chunk), hi = min(n, lo + chunk);
    unsigned bar_target = 0;
    int *bbox_tab = (int *) (a.heads + COOP_MAXG);
    coop_stamp(a, 0);
    // ---- phase 0: bounding box of the chunk -> table, barrier, every CTA folds the G boxes -------------------------
    {
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int base = lo + wid * 128; base < hi; base += COOP_THREADS * 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 32 + lane;
                p[u] = i < hi ? __ldg(a.in + i) : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (base + u * 32 + lane < hi) {
                    mn[0] = fminf(mn[0], p[u].x); mx[0] = fmaxf(mx[0], p[u].x);
                    mn[1] = fminf(mn[1], p[u].y); mx[1] = fmaxf(mx[1], p[u].y);
                    mn[2] = fminf(mn[2], p[u].z); mx[2] = fmaxf(mx[2], p[u].z);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
                mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
            }
            if (lane == 0) { S.mn[k][wid] = mn[k]; S.mx[k][wid] = mx[k]; }
        }
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float v = lane < COOP_WARPS ? S.mn[k][lane] : FLT_MAX, u = lane < COOP_WARPS ? S.mx[k][lane] : -FLT_MAX;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
                }
                if (lane == 0) { bbox_tab[6 * c + k] = f2ord(v); bbox_tab[6 * c + 3 + k] = f2ord(u); }
            }
        }
        bar_target += G;
        grid_barrier(a.bar, bar_target);
        if (wid == 0) {
            int v[6];
#pragma unroll
            for (int k = 0; k < 6; k++) v[k] = k < 3 ? f2ord(FLT_MAX) : f2ord(-FLT_MAX);
            for (int g = lane; g < G; g += 32) {
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    int b = __ldcg(bbox_tab + 6 * g + k);
                    v[k]  = k < 3 ? min(v[k], b) : max(v[k], b);
                }
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                v[k] = k < 3 ? __reduce_min_sync(0xffffffffu, v[k]) : __reduce_max_sync(0xffffffffu, v[k]);
                if (lane == 0) S.bbox[k] = v[k];
            }
        }
        __syncthreads();
    }
    const Grid3 g = make_grid(S.bbox, a.inv);
    if (c == 0 && tid == 0) {
        a.meta[M_OVERFLOW] = g.overflow ? 1 : 0;
        a.meta[M_NBITS]    = g.nbits;
#pragma unroll
        for (int k = 0; k < 6; k++) a.meta[M_BBOX + k] = S.bbox[k];
    }
    coop_stamp(a, 1);
    if (g.overflow) { // PCL: "leaf size too small": the input cloud is returned unchanged
        for (int i = lo + tid; i < hi; i += COOP_THREADS) {
            a.out[i] = a.in[i];
            if (a.keys_out) a.keys_out[i] = 0;
        }
        if (c == 0 && tid == 0) {
            *a.d_nout      = n;
            a.meta[M_NOUT] = n;
            if (a.report) *a.report = n;
        }
        coop_exit(a);
        return;
    }
    const int passes = g.nbits <= 9 ? 1 : g.nbits <= 18 ? 2 : g.nbits <= 27 ? 3 : 4;
    const int dbits  = (g.nbits + passes - 1) / passes;
    const int bins   = 1 << dbits;
    const unsigned dmask = (unsigned) bins - 1u;
    const unsigned lt    = (1u << lane) - 1u;
    // ---- phase 1: keys + histogram of digit 0 -------------------------------------------------------------------------
    S.hist[tid] = 0;
    __syncthreads();
    for (int base = lo + wid * 128; base < hi; base += COOP_THREADS * 4) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * 32 + lane;
            p[u] = i < hi ? __ldg(a.in + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * 32 + lane;
            if (base + u * 32 < hi) { // warp-uniform
                unsigned d = bins + lane; // invalid lanes: distinct values, never counted
                if (i < hi) {
                    int i0k = (int) (floorf(__fmul_rn(p[u].x, a.inv)) - (float) g.min_b[0]);
                    int i1k = (int) (floorf(__fmul_rn(p[u].y, a.inv)) - (float) g.min_b[1]);
                    int i2k = (int) (floorf(__fmul_rn(p[u].z, a.inv)) - (float) g.min_b[2]);
                    int key = i0k + i1k * g.mul1 + i2k * g.mul2;
                    a.kv[0][i] = make_uint2((uint32_t) key, (uint32_t) i);
                    if (a.keys_out) a.keys_out[i] = key;
                    d = (uint32_t) key & dmask;
                }
                const unsigned m = __match_any_sync(0xffffffffu, d);
                if (i < hi && (m & lt) == 0) atomicAdd(&S.hist[d], (uint32_t) __popc(m));
            }
        }
    }
    __syncthreads();
    if (tid < bins) table[(size_t) c * COOP_BINS + tid] = (TabT) S.hist[tid];
    bar_target += G;
    grid_barrier(a.bar, bar_target);
    coop_stamp(a, 2);
    // ---- digit passes ---------------------------------------------------------------------------------------------------
    int cur = 0;
    for (int pass = 0; pass < passes; pass++) {
        const int shift  = pass * dbits;
        const uint2 *kin = a.kv[cur];
        uint2 *kout      = a.kv[cur ^ 1];
        // (a) where this chunk's items of every digit start: digit base + counts of the chunks before this one.
        //     Thread (q, r): E consecutive digits (one 16-byte load per table row), rows r, r + R, ...
        {
            constexpr int E = 16 / (int) sizeof(TabT), QN = COOP_BINS / E, R = COOP_THREADS / QN;
            const int q = tid % QN, r = tid / QN;
            uint32_t all[E], before[E];
#pragma unroll
            for (int k = 0; k < E; k++) all[k] = before[k] = 0;
            S.tot[tid]    = 0;
            S.offset[tid] = 0;
            __syncthreads();
            if (E * q < bins) {
#pragma unroll 4
                for (int gg = r; gg < G; gg += R) {
                    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(table + (size_t) gg * COOP_BINS) + q);
                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < E; k++) {
                        const uint32_t x = sizeof(TabT) == 2 ? ((w4[k >> 1] >> ((k & 1) * 16)) & 0xffffu) : w4[k & 3];
                        all[k] += x;
                        before[k] += gg < c ? x : 0u;
                    }
                }
#pragma unroll
                for (int k = 0; k < E; k++) {
                    if (all[k]) atomicAdd(&S.tot[E * q + k], all[k]);
                    if (before[k]) atomicAdd(&S.offset[E * q + k], before[k]);
                }
            }
            __syncthreads();
            // exclusive scan of the digit totals over the 512 bins (16 warps)
            const uint32_t tot = S.tot[tid];
            uint32_t incl      = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) S.ws[wid] = (int) incl;
            __syncthreads();
            uint32_t base = 0;
            for (int w = 0; w < wid; w++) base += (uint32_t) S.ws[w];
            S.offset[tid] += base + incl - tot;
            __syncthreads();
        }
        coop_stamp(a, 3 + 3 * pass);
        // (b) stable rank + scatter of the chunk, one sub-tile of COOP_TILE items at a time
        for (int t0 = lo; t0 < hi; t0 += COOP_TILE) {
            const int tn     = min(COOP_TILE, hi - t0);
            const int C      = ((tn + COOP_THREADS - 1) / COOP_THREADS) * 32; // items per warp (contiguous)
            const int rounds = C / 32;
            {
                uint4 z = make_uint4(0, 0, 0, 0);
                uint4 *wh = reinterpret_cast<uint4 *>(&S.warp_hist[0][0]);
#pragma unroll
                for (int i = 0; i < COOP_WARPS * COOP_BINS * 2 / 16 / COOP_THREADS; i++) wh[i * COOP_THREADS + tid] = z;
            }
            __syncthreads();
            uint2 it[COOP_ROUNDS];
            uint16_t rank[COOP_ROUNDS];
#pragma unroll
            for (int r = 0; r < COOP_ROUNDS; r++) {
                if (r < rounds) {
                    const int j = wid * C + r * 32 + lane;
                    it[r]       = j < tn ? __ldcg(kin + t0 + j) : make_uint2(0xffffffffu, 0u);
                }
            }
#pragma unroll
            for (int r = 0; r < COOP_ROUNDS; r++) {
                if (r < rounds) {
                    const int j      = wid * C + r * 32 + lane;
                    const bool valid = j < tn;
                    const uint32_t d = (it[r].x >> shift) & dmask;
                    const unsigned m = __match_any_sync(0xffffffffu, valid ? d : COOP_BINS + lane);
                    const uint16_t prev = S.warp_hist[wid][d];
                    __syncwarp();
                    const unsigned rk = __popc(m & lt);
                    if (valid && rk == 0) S.warp_hist[wid][d] = (uint16_t) (prev + __popc(m));
                    __syncwarp();
                    rank[r] = (uint16_t) (prev + rk);
                }
            }
            __syncthreads();
            {   // per digit: exclusive prefix over the warps
                uint32_t tot = 0;
#pragma unroll
                for (int w = 0; w < COOP_WARPS; w++) {
                    uint16_t cnt        = S.warp_hist[w][tid];
                    S.warp_hist[w][tid] = (uint16_t) tot;
                    tot += cnt;
                }
                S.tot[tid] = tot;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < COOP_ROUNDS; r++) {
                if (r < rounds) {
                    const int j = wid * C + r * 32 + lane;
                    if (j < tn) {
                        const uint32_t d = (it[r].x >> shift) & dmask;
                        kout[S.offset[d] + S.warp_hist[wid][d] + rank[r]] = it[r];
                    }
                }
            }
            __syncthreads();
            S.offset[tid] += S.tot[tid];
            __syncthreads();
        }
        cur ^= 1;
        bar_target += G;
        grid_barrier(a.bar, bar_target);
        coop_stamp(a, 4 + 3 * pass);
        if (pass + 1 < passes) {
            // (c) histogram of the next digit over this chunk of the new order
            const int nshift = (pass + 1) * dbits;
            const uint2 *kk  = a.kv[cur];
            S.hist[tid] = 0;
            __syncthreads();
            for (int base = lo + wid * 128; base < hi; base += COOP_THREADS * 4) {
                uint32_t key[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + u * 32 + lane;
                    key[u] = i < hi ? __ldcg(&kk[i].x) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + u * 32 + lane;
                    if (base + u * 32 < hi) {
                        const unsigned d = i < hi ? ((key[u] >> nshift) & dmask) : (unsigned) (bins + lane);
                        const unsigned m = __match_any_sync(0xffffffffu, d);
                        if (i < hi && (m & lt) == 0) atomicAdd(&S.hist[d], (uint32_t) __popc(m));
                    }
                }
            }
            __syncthreads();
            if (tid < bins) table[(size_t) c * COOP_BINS + tid] = (TabT) S.hist[tid];
            bar_target += G;
            grid_barrier(a.bar, bar_target);
            coop_stamp(a, 5 + 3 * pass);
        }
    }
    // ---- heads: every warp owns a slice of the sorted items ---------------------------------------------------------------
    const uint2 *skv = a.kv[cur];
    const int W     = G * COOP_WARPS;
    const int slice = (((n + W - 1) / W) + 31) & ~31;
    const int gw    = c * COOP_WARPS + wid;
    const int s_lo  = (int) min((long long) n, (long long) gw * slice), s_hi = min(n, s_lo + slice);
    int heads = 0;
    {
        uint32_t carry = s_lo > 0 && s_lo < s_hi ? __ldcg(&skv[s_lo - 1].x) : 0u; // key in front of the slice
        for (int base = s_lo; base < s_hi; base += 128) {
            uint32_t key[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 32 + lane;
                key[u] = i < s_hi ? __ldcg(&skv[i].x) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * 32 + lane;
                uint32_t kp = __shfl_up_sync(0xffffffffu, key[u], 1);
                if (lane == 0) kp = carry;
                const bool h = i < s_hi && (i == 0 || key[u] != kp);
                heads += __popc(__ballot_sync(0xffffffffu, h));
                carry = __shfl_sync(0xffffffffu, key[u], 31);
            }
        }
    }
    if (lane == 0) S.ws[wid] = heads;
    __syncthreads();
    if (wid == 0) {
        int w = lane < COOP_WARPS ? S.ws[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        S.ws[lane] = wi - w;
        if (lane == 31) a.heads[c] = (uint32_t) wi;
    }
    bar_target += G;
    grid_barrier(a.bar, bar_target);
    coop_stamp(a, 15);
    if (wid == 0) {
        int before = 0, all = 0;
        for (int gg = lane; gg < G; gg += 32) {
            int v = (int) __ldcg(a.heads + gg);
            all += v;
            before += gg < c ? v : 0;
        }
        before = __reduce_add_sync(0xffffffffu, before);
        all    = __reduce_add_sync(0xffffffffu, all);
        if (lane == 0) { S.cta_base = before; S.cta_total = all; }
    }
    __syncthreads();
    if (c == 0 && tid == 0) {
        *a.d_nout      = S.cta_total;
        a.meta[M_NOUT] = S.cta_total;
        if (a.report) *a.report = S.cta_total;
    }
    // ---- ordered centroids: the warp walks from the first head of its slice to the first head behind it ------------------
    if (heads > 0) {
        int slot    = S.cta_base + S.ws[wid];
        float sum   = 0.f;
        int count   = 0;
        bool open   = false, done = false;
        float *outf = reinterpret_cast<float *>(a.out);
        const float *st = reinterpret_cast<const float *>(&S.stage[wid][0]) + (lane & 3);
        auto load_items = [&](int i0, uint2 &x0, uint2 &x1) {
            x0 = (i0 + lane < n) ? __ldcg(skv + i0 + lane) : make_uint2(0u, 0u);
            x1 = (i0 + 32 + lane < n) ? __ldcg(skv + i0 + 32 + lane) : make_uint2(0u, 0u);
        };
        auto gather = [&](int i0, const uint2 &x0, const uint2 &x1, float4 &p0, float4 &p1) {
            p0 = (i0 + lane < n) ? __ldg(a.in + x0.y) : make_float4(0.f, 0.f, 0.f, 0.f);
            p1 = (i0 + 32 + lane < n) ? __ldg(a.in + x1.y) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int i0 = s_lo;
        uint2 a0, a1, b0, b1;
        float4 p0, p1;
        load_items(i0, a0, a1);
        load_items(i0 + COOP_CEN, b0, b1);
        uint32_t kprev_last = (i0 > 0) ? __ldcg(&skv[i0 - 1].x) : 0u;
        gather(i0, a0, a1, p0, p1);
        while (!done && i0 < n) {
            const int cnt = min(COOP_CEN, n - i0);
            // next round's gathers and the items of the round after that: in flight during this round's accumulation
            float4 q0, q1;
            uint2 c0, c1;
            gather(i0 + COOP_CEN, b0, b1, q0, q1);
            load_items(i0 + 2 * COOP_CEN, c0, c1);
            uint32_t kp0 = __shfl_up_sync(0xffffffffu, a0.x, 1), kp1 = __shfl_up_sync(0xffffffffu, a1.x, 1);
            const uint32_t k0_last = __shfl_sync(0xffffffffu, a0.x, 31);
            if (lane == 0) { kp0 = kprev_last; kp1 = k0_last; }
            const bool h0 = lane < cnt && (i0 + lane == 0 || a0.x != kp0);
            const bool h1 = 32 + lane < cnt && a1.x != kp1;
            const unsigned hm0 = __ballot_sync(0xffffffffu, h0), hm1 = __ballot_sync(0xffffffffu, h1);
            kprev_last = __shfl_sync(0xffffffffu, a1.x, 31);
            S.stage[wid][lane]      = p0;
            S.stage[wid][32 + lane] = p1;
            __syncwarp();
            // Heads of this round as a bit mask. `mine`: heads inside the slice; the first head behind the slice ends the
            // walk. The voxel carried in from earlier rounds and the one left open at the end of the round are summed by
            // the whole warp (lane & 3 = channel, adds back to back); the voxels that begin AND end inside the round are
            // summed concurrently, one lane each (what keeps one-point-per-voxel regions from serialising the warp).
            const unsigned long long hm = (unsigned long long) hm0 | ((unsigned long long) hm1 << 32);
            const int limit = max(0, min(cnt, s_hi - i0));
            const unsigned long long inmask = limit >= 64 ? ~0ull : ((1ull << limit) - 1ull);
            const unsigned long long mine = hm & inmask, foreign = hm & ~inmask;
            const int E = foreign ? __ffsll((long long) foreign) - 1 : cnt;   // items [0, E) belong to this warp's voxels
            auto add_range = [&](int j0, int j1) {
                int j = j0;
                for (; j + 8 <= j1; j += 8) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = st[4 * (j + k)];
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += v[k];
                }
                for (; j < j1; j++) sum += st[4 * j];
                count += j1 - j0;
            };
            if (open) add_range(0, mine ? __ffsll((long long) mine) - 1 : E);
            if (mine) {
                if (open) {
                    if (lane < 4) outf[4 * (size_t) slot + lane] = sum / (float) count;
                    slot++;
                }
                const int last = 63 - __clzll((long long) mine);
                const unsigned long long mid = mine & ~(1ull << last);
                if (mid) {
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int pos = 32 * half + lane;
                        if ((mid >> pos) & 1ull) {
                            const int nxt = pos + __ffsll((long long) (hm >> (pos + 1))); // a later head exists (at most `last`)
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int j = pos; j < nxt; j++) {
                                const float4 v = S.stage[wid][j];
                                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                            }
                            const float fc = (float) (nxt - pos);
                            a.out[slot + __popcll(mid & ((1ull << pos) - 1ull))] = make_float4(acc.x / fc, acc.y / fc, acc.z / fc, acc.w / fc);
                        }
                    }
                    slot += __popcll(mid);
                }
                open  = true;
                sum   = 0.f;
                count = 0;
                add_range(last, E);
            }
            if (foreign) {
                if (open && lane < 4) outf[4 * (size_t) slot + lane] = sum / (float) count;
                open = false;
                done = true;
            }
            __syncwarp();
            i0 += COOP_CEN;
            a0 = b0; a1 = b1; b0 = c0; b1 = c1; p0 = q0; p1 = q1;
        }
        if (open && lane < 4) outf[4 * (size_t) slot + lane] = sum / (float) count;
    }
    coop_stamp(a, 16);
    coop_exit(a);
}

// ---- small clouds (n <= SMALL_CAP): the whole filter in ONE CTA ---------------------------------------
// bbox -> keys -> stable LSD radix sort (8-bit digits, only the passes the key width needs) entirely in
// shared memory -> heads -> ordered centroid. Replaces 18 launch-latency-bound launches for the per-frame
// filter of ~2-3 k decimated points. (A first version sorted 64-bit composites with a bitonic network:
// 66 barrier rounds, ~3,000 instructions per warp, issue-bound on its one SM; the radix passes need ~1/4.)
constexpr int SMALL_CAP     = kVoxSmallCap;
constexpr int SMALL_THREADS = 1024;
constexpr int SMALL_ROUNDS  = SMALL_CAP / SMALL_THREADS; // 32-item rounds per warp

constexpr int SMALL_DIGIT = 9;                 // 18-bit keys (every per-frame cloud) sort in two passes
constexpr int SMALL_BINS  = 1 << SMALL_DIGIT;
struct SmallSmem {
    float4 pts[SMALL_CAP];
    uint32_t key[2][SMALL_CAP];
    uint16_t idx[2][SMALL_CAP];
    uint16_t warp_hist[32][SMALL_BINS];
    uint16_t half_tot[2][SMALL_BINS]; // digit totals of warps 0-15 and 16-31
    uint32_t digit_base[SMALL_BINS];
    float mn[3][32], mx[3][32];
    int bbox[6];
    int ws[32];
    int total;
};

__device__ __forceinline__ void
vox_small_body(const float4 *__restrict__ in, int n, float inv, int *__restrict__ meta, float4 *__restrict__ out,
               int *__restrict__ d_nout, int32_t *__restrict__ keys_out, bool has_tail, const VoxelTail &tail) {
    extern __shared__ __align__(16) unsigned char small_raw[];
    SmallSmem &S = *reinterpret_cast<SmallSmem *>(small_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // 1. stage the points, bounding box
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = tid; i < n; i += SMALL_THREADS) {
        float4 p = __ldg(in + i);
        S.pts[i] = p;
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) { S.mn[a][wid] = mn[a]; S.mx[a][wid] = mx[a]; }
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float v = S.mn[a][lane], u = S.mx[a][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
                u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
            }
            if (lane == 0) { S.bbox[a] = f2ord(v); S.bbox[3 + a] = f2ord(u); }
        }
    }
    __syncthreads();
    Grid3 g = make_grid(S.bbox, inv);
    if (tid == 0) {
        meta[M_OVERFLOW] = g.overflow ? 1 : 0;
        meta[M_NBITS]    = g.nbits;
    }
    if (g.overflow) { // PCL: "leaf size too small": the input cloud is returned unchanged
        for (int i = tid; i < n; i += SMALL_THREADS) {
            out[i] = S.pts[i];
            if (has_tail) tail.world[i] = project_rn(tail.T.R, tail.T.t, S.pts[i]);
            if (keys_out) keys_out[i] = 0;
        }
        if (tid == 0) {
            *d_nout = n;
            meta[M_NOUT] = n;
            if (has_tail) { tail.report[0] = tail.counters ? tail.counters[0] : 0; tail.report[1] = n; }
        }
        return;
    }
    // 2. keys
    for (int i = tid; i < n; i += SMALL_THREADS) {
        float4 p = S.pts[i];
        int i0   = (int) (floorf(__fmul_rn(p.x, inv)) - (float) g.min_b[0]);
        int i1   = (int) (floorf(__fmul_rn(p.y, inv)) - (float) g.min_b[1]);
        int i2   = (int) (floorf(__fmul_rn(p.z, inv)) - (float) g.min_b[2]);
        int key  = i0 + i1 * g.mul1 + i2 * g.mul2;
        if (keys_out) keys_out[i] = key;
        S.key[0][i] = (uint32_t) key;
        S.idx[0][i] = (uint16_t) i;
    }
    // 3. stable LSD radix sort in shared memory. Warp w owns the contiguous chunk [w*C, (w+1)*C).
    const int C       = ((n + SMALL_THREADS - 1) / SMALL_THREADS) * 32;
    const int rounds  = C / 32;
    const unsigned lt = (1u << lane) - 1u;
    int cur = 0;
    for (int shift = 0; shift < g.nbits; shift += SMALL_DIGIT) {
        for (int i = tid; i < 32 * SMALL_BINS; i += SMALL_THREADS) (&S.warp_hist[0][0])[i] = 0;
        __syncthreads();
        uint32_t key[SMALL_ROUNDS];
        uint16_t val[SMALL_ROUNDS], rank[SMALL_ROUNDS];
#pragma unroll
        for (int r = 0; r < SMALL_ROUNDS; r++) {
            if (r < rounds) {
                const int i      = wid * C + r * 32 + lane;
                const bool valid = i < n;
                key[r]           = valid ? S.key[cur][i] : 0xffffffffu;
                val[r]           = valid ? S.idx[cur][i] : (uint16_t) 0;
                const uint32_t d = (key[r] >> shift) & (SMALL_BINS - 1);
                const unsigned m = __match_any_sync(0xffffffffu, valid ? d : SMALL_BINS + lane);
                const uint16_t prev = S.warp_hist[wid][d];
                __syncwarp();
                const unsigned rk = __popc(m & lt);
                if (valid && rk == 0) S.warp_hist[wid][d] = (uint16_t) (prev + __popc(m));
                __syncwarp();
                rank[r] = (uint16_t) (prev + rk);
            }
        }
        __syncthreads();
        // per digit: exclusive prefix over the warps — thread (digit, half) walks 16 warps, the two halves run side by
        // side and the second half's offset (the first half's total) is added when the items are scattered
        {
            const int d = tid & (SMALL_BINS - 1), half = tid >> SMALL_DIGIT;
            uint32_t tot = 0;
#pragma unroll 8
            for (int w = 16 * half; w < 16 * half + 16; w++) {
                uint16_t c        = S.warp_hist[w][d];
                S.warp_hist[w][d] = (uint16_t) tot;
                tot += c;
            }
            S.half_tot[half][d] = (uint16_t) tot;
        }
        __syncthreads();
        uint32_t tot = 0;
        if (tid < SMALL_BINS) {
            tot = (uint32_t) S.half_tot[0][tid] + (uint32_t) S.half_tot[1][tid];
            // exclusive scan of the digit totals (16 warps)
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) S.ws[wid] = (int) incl;
            tot = incl - tot; // exclusive within the warp
        }
        __syncthreads();
        if (tid < SMALL_BINS) {
            uint32_t base = 0;
            for (int w = 0; w < wid; w++) base += (uint32_t) S.ws[w];
            S.digit_base[tid] = base + tot;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SMALL_ROUNDS; r++) {
            if (r < rounds) {
                const int i = wid * C + r * 32 + lane;
                if (i < n) {
                    const uint32_t d   = (key[r] >> shift) & (SMALL_BINS - 1);
                    const uint32_t dst = S.digit_base[d] + S.warp_hist[wid][d] + (wid >= 16 ? (uint32_t) S.half_tot[0][d] : 0u) + rank[r];
                    S.key[cur ^ 1][dst] = key[r];
                    S.idx[cur ^ 1][dst] = val[r];
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    const uint32_t *skey = S.key[cur];
    const uint16_t *sidx = S.idx[cur];
    // 4. heads -> slots (blocked: thread t owns items [t*per, (t+1)*per))
    const int per = (n + SMALL_THREADS - 1) / SMALL_THREADS;
    const int lo  = tid * per, hi = min(n, lo + per);
    int c = 0;
    for (int i = lo; i < hi; i++) c += (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads(); // S.ws is reused
    if (lane == 31) S.ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = S.ws[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        S.ws[lane] = wi - w;
        if (lane == 31) S.total = wi;
    }
    __syncthreads();
    int slot = S.ws[wid] + incl - c;
    // 5. ordered centroid: one thread per head, sequential fp32 adds in ascending index
    for (int i = lo; i < hi; i++) {
        const uint32_t key = skey[i];
        if (!(i == 0 || skey[i - 1] != key)) continue;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        int e = i;
        while (e < n && skey[e] == key) {
            float4 p = S.pts[sidx[e]];
            sx += p.x; sy += p.y; sz += p.z; si += p.w;
            e++;
        }
        float cnt = (float) (e - i);
        float4 c4 = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
        out[slot] = c4;
        // fused tail: world projection of the centroid (pointcloud.cc:38-59), exactly rounded
        if (has_tail) tail.world[slot] = project_rn(tail.T.R, tail.T.t, c4);
        slot++;
    }
    if (tid == 0) {
        *d_nout = S.total;
        meta[M_NOUT] = S.total;
        if (has_tail) { // zero-copy result read-back (pinned host memory)
            tail.report[0] = tail.counters ? tail.counters[0] : 0;
            tail.report[1] = S.total;
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS)
vox_small_kernel(const float4 *__restrict__ in, int n, float inv, int *__restrict__ meta, float4 *__restrict__ out,
                 int *__restrict__ d_nout, int32_t *__restrict__ keys_out, bool has_tail, const __grid_constant__ VoxelTail tail) {
    vox_small_body(in, n, inv, meta, out, d_nout, keys_out, has_tail, tail);
}

// one CTA per queued frame: decimated cloud -> voxel centroids -> world projection -> pinned count report
__global__ void __launch_bounds__(SMALL_THREADS) vox_frame_kernel(const __grid_constant__ FrameBatch B) {
    const FrameItem &it = B.it[blockIdx.x];
    if (it.ndec > SMALL_CAP) return; // multi-kernel path, launched by the host for this frame
    if (it.ndec <= 0) {
        if (threadIdx.x == 0) {
            *it.d_nout   = 0;
            it.report[0] = it.counters[0];
            it.report[1] = 0;
        }
        return;
    }
    VoxelTail tail;
    tail.T        = it.frame;
    tail.world    = it.world;
    tail.report   = it.report;
    tail.counters = it.counters;
    vox_small_body(it.out_dec, it.ndec, it.inv_leaf, it.meta, it.down, it.d_nout, nullptr, true, tail);
}

} // namespace

void launch_frame_voxel(cudaStream_t s, const FrameBatch &b) {
    if (b.count <= 0) return;
    static std::atomic<bool> attr_set[64]; // (zero-initialised; contexts on different threads may race here: the call is idempotent)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(vox_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SmallSmem));
        attr_set[dev & 63] = true;
    }
    vox_frame_kernel<<<b.count, SMALL_THREADS, sizeof(SmallSmem), s>>>(b);
}

size_t voxel_work_bytes(int cap, size_t *off) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    size_t nb  = (size_t) (cap + RS_TILE - 1) / RS_TILE + 1;
    size_t nb2 = (size_t) (cap + SEG_TILE - 1) / SEG_TILE + 1;
    size_t o   = 0;
    off[0] = o; o += al((size_t) cap * 4);
    off[1] = o; o += al((size_t) cap * 4);
    off[2] = o; o += al((size_t) cap * 4);
    off[3] = o; o += al((size_t) cap * 4);
    off[4] = o; o += al(std::max<size_t>(256 * nb * 4 + 4096, (size_t) COOP_MAXG * COOP_BINS * 4)); // + 1024 chunk sums of the multi-CTA scan; or the cooperative kernel's [G][bins] table
    off[5] = o; o += al(std::max<size_t>(nb2 * 4, (size_t) COOP_MAXG * 7 * 4)); // chunk head counts + chunk bounding boxes
    off[6] = o; o += al(M_SIZE * 4);
    off[7] = o; o += al(((size_t) cap + 1) * 4);
    return o;
}

int launch_voxel_downsample(cudaStream_t s, const VoxelWork &w, const float4 *in, int n, float leaf, float4 *out,
                            int *d_nout, int32_t *keys_out, int64_t *launches, Profiler *prof, const char *tag,
                            const VoxelTail *tail, int *report) {
    if (n <= 0) {
        cudaMemsetAsync(d_nout, 0, sizeof(int), s);
        return 0;
    }
    if (n > w.cap) return -1;
    const float inv = 1.0f / leaf;
    char name[64];
    auto nm = [&](const char *k) {
        snprintf(name, sizeof(name), "%s.%s", tag, k);
        return name;
    };
    if (n <= SMALL_CAP) {
        static std::atomic<bool> attr_set[64]; // (zero-initialised; contexts on different threads may race here: the call is idempotent)
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set[dev & 63]) {
            cudaFuncSetAttribute(vox_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SmallSmem));
            attr_set[dev & 63] = true;
        }
        ProfScope ps(prof, nm("one_cta"), s);
        VoxelTail none{};
        vox_small_kernel<<<1, SMALL_THREADS, sizeof(SmallSmem), s>>>(in, n, inv, w.meta, out, d_nout, keys_out, tail != nullptr,
                                                                     tail ? *tail : none);
        if (launches) *launches += 1;
        return tail ? 1 : 0;
    }
    if (w.fresh) { // barrier counters of a new work buffer
        cudaMemsetAsync(w.meta, 0, 256, s);
        w.fresh = false;
    }
    static const bool chain_only = getenv("FFLINS_VOXEL_CHAIN") != nullptr;
    static const bool coop_stamps = getenv("FFLINS_VOXEL_STAMPS") != nullptr;
    if (!chain_only) {
        static std::atomic<int> sm_count[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (!sm_count[dev & 63]) {
            int coop = 0, sms = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            sm_count[dev & 63] = coop ? sms : -1;
        }
        const int sm_n = sm_count[dev & 63].load();
        if (sm_n > 0 && w.keys[1] == w.keys[0] + w.cap && w.idx[1] == w.idx[0] + w.cap) {
            CoopArgs a{};
            a.in = in; a.n = n; a.inv = inv; a.meta = w.meta;
            // keys[0] | keys[1] and idx[0] | idx[1] are adjacent (the capacity is a power of two): two arrays of 8-byte items
            a.kv[0] = reinterpret_cast<uint2 *>(w.keys[0]); a.kv[1] = reinterpret_cast<uint2 *>(w.idx[0]);
            a.table = w.block_hist; a.heads = w.block_heads;
            a.bar = (unsigned *) (w.meta + 14);
            a.out = out; a.d_nout = d_nout; a.keys_out = keys_out; a.report = report;
            a.stamps = coop_stamps ? (unsigned long long *) (w.meta + 16) : nullptr;
            // one CTA per SM leaves half of every SM to the solver-facing kernels (a keyframe map); a launch that is large enough
            // to be bandwidth-bound takes both halves: twice the warps to hide the latency of its streaming phases
            const int per_sm = n >= (4 << 20) ? 2 : 1;
            int G = std::min(std::min(sm_n * per_sm, COOP_MAXG), (n + 2047) / 2048);
            void *args[] = {(void *) &a};
            static const bool force_wide = getenv("FFLINS_VOXEL_WIDE") != nullptr;   // tests: the uint32 instantiation at any size
            const bool narrow = !force_wide && (n + G - 1) / G + 128 < 65536; // chunk digit counts fit 16 bits
            ProfScope ps(prof, nm("coop"), s);
            cudaError_t e = cudaLaunchCooperativeKernel(narrow ? (const void *) vox_coop_kernel<uint16_t> : (const void *) vox_coop_kernel<uint32_t>,
                                                        dim3(G), dim3(COOP_THREADS), args, 0, s);
            if (e == cudaSuccess) {
                if (launches) *launches += 1;
                if (coop_stamps) { // debug: phase durations of this launch (synchronises the stream)
                    unsigned long long t[COOP_STAMPS] = {0};
                    cudaStreamSynchronize(s);
                    cudaMemcpy(t, w.meta + 16, sizeof(t), cudaMemcpyDeviceToHost);
                    fprintf(stderr, "vox_coop n=%d G=%d ns:", n, G);
                    unsigned long long prev = t[0];
                    for (int k = 1; k < COOP_STAMPS - 1; k++)
                        if (t[k] >= prev && t[k] != 0) {
                            fprintf(stderr, " [%d]+%llu", k, t[k] - prev);
                            prev = t[k];
                        }
                    fprintf(stderr, " cta0 %llu all %llu\n", prev - t[0], t[COOP_STAMPS - 1] - t[0]);
                    cudaMemsetAsync(w.meta + 16, 0, sizeof(t), s);
                }
                return report ? 2 : 0;
            }
            cudaGetLastError(); // fall back to the multi-kernel chain
        }
    }
    int g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        ProfScope ps(prof, nm("bbox_keys"), s);
        vox_init_kernel<<<1, 32, 0, s>>>(w.meta);
        vox_bbox_kernel<<<g, 256, 0, s>>>(in, n, w.meta);
        vox_keys_kernel<<<g, 256, 0, s>>>(in, n, inv, w.meta, w.keys[0], keys_out);
    }
    const int nb = (n + RS_TILE - 1) / RS_TILE;
    {
        ProfScope ps(prof, nm("radix_sort"), s);
        for (int pass = 0; pass < 4; pass++) {
            rs_hist_kernel<<<nb, RS_THREADS, 0, s>>>(w.keys[0], w.keys[1], w.meta, pass, n, w.block_hist, nb);
            const int total = 256 * nb;
            if (total <= 8 * SCAN_SUPER) {
                rs_scan_kernel<<<1, 1024, 0, s>>>(w.meta, pass, w.block_hist, total);
            } else {
                int g     = std::min(296, (total + SCAN_SUPER - 1) / SCAN_SUPER);
                int chunk = ((total + g - 1) / g + SCAN_SUPER - 1) / SCAN_SUPER * SCAN_SUPER;
                g         = (total + chunk - 1) / chunk;
                uint32_t *partial = w.block_hist + (size_t) 256 * ((size_t) (w.cap + RS_TILE - 1) / RS_TILE + 1);
                rs_scan_reduce_kernel<<<g, 1024, 0, s>>>(w.meta, pass, w.block_hist, total, chunk, partial);
                rs_scan_partials_kernel<<<1, 1024, 0, s>>>(w.meta, pass, partial, g);
                rs_scan_apply_kernel<<<g, 1024, 0, s>>>(w.meta, pass, w.block_hist, total, chunk, partial);
            }
            rs_scatter_kernel<<<nb, RS_THREADS, 0, s>>>(w.keys[0], w.keys[1], w.idx[0], w.idx[1], w.meta, pass, n,
                                                        w.block_hist, nb);
        }
    }
    const int nb2 = (n + SEG_TILE - 1) / SEG_TILE;
    {
        ProfScope ps(prof, nm("centroid"), s);
        seg_count_kernel<<<nb2, SEG_THREADS, 0, s>>>(w.keys[0], w.keys[1], w.meta, n, w.block_heads);
        seg_starts_kernel<<<nb2, SEG_THREADS, 0, s>>>(w.keys[0], w.keys[1], w.meta, n, w.block_heads, w.seg_start, d_nout);
        int gc = (n + 7) / 8;
        if (gc > 148 * 8) gc = 148 * 8;
        seg_centroid_kernel<<<gc, CEN_THREADS, 0, s>>>(in, w.idx[0], w.idx[1], w.meta, n, w.seg_start, out);
    }
    if (launches) *launches += 3 + (256 * nb <= 8 * SCAN_SUPER ? 12 : 20) + 3;
    return 0;
}

} // namespace fflins
