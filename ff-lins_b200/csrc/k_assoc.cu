#include <algorithm>
#include <atomic>
// k_assoc.cu — K4 keyframe-map hash insert and K5 frame-to-frame association:
// exact 5-NN on a device-resident voxel hash + 5-point plane fit + validity tests.
//
// Reference behaviour restated (not translated):
//   LidarMap::addNewKeyFrame -> ikd_Tree::build      ff_lins/lidar/lidar_map.cc:100-126
//   LidarMap::findFeaturesInMap                      ff_lins/lidar/lidar_map.cc:326-408
//   KD_TREE::nearestSearch / Search / calc_dist      ff_lins/lidar/ikd-Tree/ikd_Tree.cc:360-391,897-924,1427-1431
//   PointCloudCommon::planeEstimation                ff_lins/lidar/pointcloud.cc:61-93
//
// The kd-tree is NOT reproduced, only its result: the exact 5 nearest map points by fp32
// d2 = ((dx*dx + dy*dy) + dz*dz), cut-off d2 <= max_dist^2, ascending, ties broken by the lower
// map index. B200 design: one warp per (ref keyframe, query) over a two-level open-addressing hash
// per keyframe map: a fine table (key = packed voxel cell, value = map index) and a coarse occupancy
// table (one 64-bit mask per 4x4x4 block of cells). The warp sweeps Chebyshev stages of the search
// cube; occupied cells are gathered from the masks into a shared-memory queue and probed 32 at a
// time; each lane keeps a sorted top-5 of 64-bit (d2 bits << 32 | index) keys in registers and the
// warp merges them with shuffle min-reductions. The search stops after the first stage whose lower
// bound exceeds the current 5th distance (strictly, with a 1 mm guard), so the result is provably
// the exact 5-NN. All (ref, query) pairs of a keyframe go in ONE launch.
//
// This translation unit is compiled with -fmad=false: every fp32/fp64 product and sum rounds
// separately, like the reference's x86-64 build, so d2, the QR normals and the validity
// decisions are reproducible bit for bit.
#include "math_plane.cuh"

namespace fflins {

namespace {

typedef unsigned long long u64;
constexpr u64 kEmpty = 0xffffffffffffffffull;
constexpr u64 kInf   = 0xffffffffffffffffull;

__device__ __forceinline__ u64 pack_cell(int ix, int iy, int iz) {
    const int off = 1 << 20;
    u64 a = (u64) (unsigned) min(max(ix + off, 0), (1 << 21) - 1);
    u64 b = (u64) (unsigned) min(max(iy + off, 0), (1 << 21) - 1);
    u64 c = (u64) (unsigned) min(max(iz + off, 0), (1 << 21) - 1);
    return a | (b << 21) | (c << 42);
}
__device__ __forceinline__ unsigned hash_cell(u64 k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned) k;
}

__global__ void hash_clear_kernel(MapIndex X) {
    const int hcap = X.hmask + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hcap; i += gridDim.x * blockDim.x) {
        X.fine[i].key    = kEmpty;
        if (i == 0) X.fine[0].pad = 0ull; // "some cell holds more than one point" flag, see hash_insert_kernel
        X.coarse[i].key  = kEmpty;
        X.coarse[i].mask = 0ull;
        X.super[i].key   = kEmpty;
        X.super[i].mask  = 0ull;
    }
}

__device__ __forceinline__ void mask_insert(MaskSlot *__restrict__ tab, int hmask, u64 key, int bit) {
    unsigned h = hash_cell(key) & hmask;
    while (true) {
        u64 old = atomicCAS(&tab[h].key, kEmpty, key);
        if (old == kEmpty || old == key) {
            atomicOr(&tab[h].mask, 1ull << bit);
            return;
        }
        h = (h + 1) & hmask;
    }
}

// One fine slot per map point; points sharing a cell lie in the same probe run (not necessarily adjacent).
// m_dev (optional): the point count still sits in device memory (map built on the same stream just before); a
// table that would be loaded above 0.7 is left empty — the host applies the same test and rebuilds it.
__global__ void hash_insert_kernel(const float4 *__restrict__ map, int m, const int *__restrict__ m_dev, float inv_cell,
                                   MapIndex X) {
    if (m_dev) {
        m = *m_dev;
        if ((long long) m * 10 > (long long) (X.hmask + 1) * 7) return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        float4 p = __ldg(map + i);
        int ix = (int) floorf(p.x * inv_cell), iy = (int) floorf(p.y * inv_cell), iz = (int) floorf(p.z * inv_cell);
        u64 key    = pack_cell(ix, iy, iz);
        unsigned h = hash_cell(key) & X.hmask;
        while (true) {
            u64 old = atomicCAS(&X.fine[h].key, kEmpty, key);
            if (old == kEmpty) {
                X.fine[h].x   = p.x;
                X.fine[h].y   = p.y;
                X.fine[h].z   = p.z;
                X.fine[h].idx = i;
                break;
            }
            // A second point in the same cell: searches must then walk to the end of the probe run. With the search
            // cell equal to the voxel leaf (the default) a keyframe map never has one, and a probe stops at its hit.
            if (old == key) X.fine[0].pad = 1ull;
            h = (h + 1) & X.hmask;
        }
        // coarse occupancy: the 4x4x4 block of cells, bit of this cell
        mask_insert(X.coarse, X.hmask, pack_cell(ix >> 2, iy >> 2, iz >> 2), (ix & 3) | ((iy & 3) << 2) | ((iz & 3) << 4));
        // super occupancy: the 4x4x4 block of coarse blocks, bit of this coarse block
        mask_insert(X.super, X.hmask, pack_cell(ix >> 4, iy >> 4, iz >> 4),
                    ((ix >> 2) & 3) | (((iy >> 2) & 3) << 2) | (((iz >> 2) & 3) << 4));
    }
}

__device__ __forceinline__ bool ins5(u64 &a0, u64 &a1, u64 &a2, u64 &a3, u64 &a4, u64 v) {
    if (v < a4) {
        a4 = v;
        if (a4 < a3) { u64 t = a3; a3 = a4; a4 = t; }
        if (a3 < a2) { u64 t = a2; a2 = a3; a3 = t; }
        if (a2 < a1) { u64 t = a1; a1 = a2; a2 = t; }
        if (a1 < a0) { u64 t = a0; a0 = a1; a1 = t; }
        return true;
    }
    return false;
}

typedef MapIndex HashView;

// Probe one fine cell and push every map point stored under it. A slot is two independent 16-byte loads
// (key | x y, z idx | pad), so a hit needs ONE memory round trip; the next probe slot is requested before the
// current one is examined.
__device__ __forceinline__ void probe_cell(const HashView &H, bool multi, int fx, int fy, int fz, float qx, float qy, float qz,
                                           float max_d2, u64 &a0, u64 &a1, u64 &a2, u64 &a3, u64 &a4, bool &dirty) {
    const u64 key = pack_cell(fx, fy, fz);
    unsigned h    = hash_cell(key) & H.hmask;
    const ulonglong2 *tab = reinterpret_cast<const ulonglong2 *>(H.fine);
    ulonglong2 s0a = __ldg(tab + 2 * h), s0b = __ldg(tab + 2 * h + 1);
    while (true) {
        const unsigned hn = (h + 1) & H.hmask;
        ulonglong2 s1a = __ldg(tab + 2 * hn), s1b = __ldg(tab + 2 * hn + 1); // speculative next slot
        if (s0a.x == kEmpty) break;
        if (s0a.x == key) {
            float bx = __uint_as_float((unsigned) (s0a.y & 0xffffffffull)), by = __uint_as_float((unsigned) (s0a.y >> 32));
            float bz = __uint_as_float((unsigned) (s0b.x & 0xffffffffull));
            unsigned idx = (unsigned) (s0b.x >> 32);
            float ex = qx - bx, ey = qy - by, ez = qz - bz;
            float d2 = (ex * ex + ey * ey) + ez * ez; // ikd_Tree.cc:1427-1431 (no FMA: -fmad=false)
            if (d2 <= max_d2) dirty |= ins5(a0, a1, a2, a3, a4, ((u64) __float_as_uint(d2) << 32) | idx);
            if (!multi) break; // one point per cell: the hit ends the probe (otherwise walk to the end of the run)
        }
        s0a = s1a;
        s0b = s1b;
        h   = hn;
    }
}

__device__ __forceinline__ void warp_merge5(u64 a0, u64 a1, u64 a2, u64 a3, u64 a4, u64 best[5]) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
        u64 mn  = warp_min_u64(a0);
        best[k] = mn;
        if (a0 == mn && mn != kInf) { a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = kInf; }
    }
}

// 64-bit mask of the cells (sub = x | y<<2 | z<<4) of a 4x4x4 block whose Chebyshev distance to the query
// cell is <= r; (ax, ay, az) is the offset of the block's first cell from the query cell.
__device__ __forceinline__ unsigned axis_bits4(int a, int r) {
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) b |= (abs(a + i) <= r ? 1u : 0u) << i;
    return b;
}
// outer product of three 4-bit axis masks: bit (x | y<<2 | z<<4) = nx[x] & ny[y] & nz[z]
__device__ __forceinline__ u64 outer3(unsigned nx, unsigned ny, unsigned nz) {
    u64 X = (u64) nx * 0x1111111111111111ull;
    unsigned y16 = 0;
#pragma unroll
    for (int y = 0; y < 4; y++) y16 |= ((ny >> y) & 1u) ? (0xFu << (4 * y)) : 0u;
    u64 Y = (u64) y16 * 0x0001000100010001ull;
    u64 Z = 0ull;
#pragma unroll
    for (int z = 0; z < 4; z++) Z |= ((nz >> z) & 1u) ? (0xFFFFull << (16 * z)) : 0ull;
    return X & Y & Z;
}
__device__ __forceinline__ u64 cube_mask(int ax, int ay, int az, int r) {
    if (r < 0) return 0ull;
    return outer3(axis_bits4(ax, r), axis_bits4(ay, r), axis_bits4(az, r));
}
// The same for the 4x4x4 COARSE blocks (4 cells wide each) of a super block whose first cell sits at offset s from
// the query cell: bit i of the axis mask = the nearest (farthest) cell of sub-block i is within r cells on that axis.
// A block's Chebyshev range [mn, mx] is the per-axis maximum, so "mn <= r" and "mx <= r" are both outer products.
__device__ __forceinline__ unsigned axis_blocks_near(int s, int r) {
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = s + 4 * i;
        b |= (max(a, -(a + 3)) <= r ? 1u : 0u) << i;
    }
    return b;
}
__device__ __forceinline__ unsigned axis_blocks_far(int s, int r) {
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int a = s + 4 * i;
        b |= (max(abs(a), abs(a + 3)) <= r ? 1u : 0u) << i;
    }
    return b;
}
// coarse blocks of a super block that intersect the Chebyshev shell (lo, hi]: mn <= hi and mx > lo
__device__ __forceinline__ u64 shell_blocks(int sx, int sy, int sz, int lo, int hi) {
    u64 in = outer3(axis_blocks_near(sx, hi), axis_blocks_near(sy, hi), axis_blocks_near(sz, hi));
    if (lo >= 0) in &= ~outer3(axis_blocks_far(sx, lo), axis_blocks_far(sy, lo), axis_blocks_far(sz, lo));
    return in;
}

constexpr int kAssocWarps = 1;            // warps per CTA: one. A CTA keeps its slot until its SLOWEST warp is done and search times vary 20x
                                          // between queries: with 4 warps per CTA only 16 of the 28 possible warps per SM were active on average
constexpr int kLookupLanes = 16;           // lanes that fetch an occupancy mask per round (bounds the queues)
constexpr int kQueueCap   = kLookupLanes * 64; // occupied fine cells of 16 coarse blocks
constexpr int kCQueueCap  = kLookupLanes * 64; // occupied coarse blocks of 16 super blocks

// Warp-cooperative exact 5-NN on the two-level hash. best[] is warp-uniform on return (kInf = not found).
//
// The search cube is swept in Chebyshev stages (lo, hi] = (0,2], (2,4], (4,8], (8,r_max] of fine cells.
// In a stage every lane fetches the 64-bit occupancy mask of one 4x4x4 coarse block that intersects
// the stage (empty space costs one probe per 64 cells), the set bits inside the stage are written to
// a per-warp shared-memory queue, and the queue is drained 32 cells per round, one fine-hash probe
// per lane — so the probes of a round are in flight together and the load is balanced no matter how
// the points are distributed over the blocks. After a stage every unvisited map point is farther than
// hi*cell, so the search stops as soon as d5 < (hi*cell - 1 mm)^2: the result is the exact 5-NN.
__device__ __forceinline__ u64 lookup_mask(const MaskSlot *__restrict__ tab, int hmask, u64 key) {
    unsigned h = hash_cell(key) & hmask;
    const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(tab);
    while (true) {
        ulonglong2 s = __ldg(t + h); // key and mask in one 16-byte load
        if (s.x == kEmpty) return 0ull;
        if (s.x == key) return s.y;
        h = (h + 1) & hmask;
    }
}

// Chebyshev range [mn, mx] of the cells of a block of `size` cells per axis whose first cell sits at offset
// (ax, ay, az) from the query cell.
__device__ __forceinline__ void block_range(int ax, int ay, int az, int size, int &mn, int &mx) {
    const int e = size - 1;
    int mnx = max(0, max(ax, -(ax + e))), mny = max(0, max(ay, -(ay + e))), mnz = max(0, max(az, -(az + e)));
    int mxx = max(abs(ax), abs(ax + e)), mxy = max(abs(ay), abs(ay + e)), mxz = max(abs(az), abs(az + e));
    mn = max(mnx, max(mny, mnz));
    mx = max(mxx, max(mxy, mxz));
}

constexpr int kSparseFrom = 8; // stages reaching this far enumerate occupied coarse blocks through the super masks

// squared lower bound (in cells^2) of the distance from the query (anywhere inside its own cell) to a cell at
// per-axis offsets whose absolute values are at least gx, gy, gz cells
__device__ __forceinline__ float cells_lb2(int gx, int gy, int gz) {
    float x = (float) max(gx - 1, 0), y = (float) max(gy - 1, 0), z = (float) max(gz - 1, 0);
    return x * x + y * y + z * z;
}

// t -> (bx, by, bz) of an nx x ny x nz box without integer division: (t + 0.5) / d is never within 0.5 / d of an
// integer, so the truncated fp32 product is exact for the few hundred boxes a stage can have
__device__ __forceinline__ void split3(int t, int nx, int ny, int &bx, int &by, int &bz) {
    const int nxy      = nx * ny;
    const float inv_xy = __fdividef(1.0f, (float) nxy), inv_x = __fdividef(1.0f, (float) nx); // approximate: margin 0.5 / d
    bz      = (int) (((float) t + 0.5f) * inv_xy);
    int rem = t - bz * nxy;
    by      = (int) (((float) rem + 0.5f) * inv_x);
    bx      = rem - by * nx;
}

__device__ __forceinline__ void warp_knn5(const HashView &H, float qx, float qy, float qz, float inv_cell, float cell,
                                          int r_max, float max_d2, unsigned *__restrict__ queue,
                                          unsigned *__restrict__ cqueue, u64 best[5]) {
    const int lane = threadIdx.x & 31;
    const int cx = (int) floorf(qx * inv_cell), cy = (int) floorf(qy * inv_cell), cz = (int) floorf(qz * inv_cell);
    u64 a0 = kInf, a1 = kInf, a2 = kInf, a3 = kInf, a4 = kInf;
    const bool multi = __ldg(&H.fine[0].pad) != 0ull;
    bool dirty = false;                                  // this lane inserted since the last warp merge
    // pruning bound in cells^2: a cell (block) whose lower-bound distance exceeds the cut-off, or the current 5th
    // distance once five neighbours are known, cannot contribute (strict, with a 0.2 % guard for fp32 rounding)
    const float cell2 = cell * cell;
    float prune_c2    = max_d2 / cell2 * 1.002f;

    // first probe of the (<= 27) coarse masks of the near phase: issued now, consumed after the fast-exit test
    bool near_have = false;
    int nax = 0, nay = 0, naz = 0;
    u64 near_key = 0ull;
    ulonglong2 near_first = make_ulonglong2(kEmpty, 0ull);
    if (r_max >= 4) {
        const int lox = (cx - 4) >> 2, loy = (cy - 4) >> 2, loz = (cz - 4) >> 2;
        const int nx = ((cx + 4) >> 2) - lox + 1, ny = ((cy + 4) >> 2) - loy + 1, nz = ((cz + 4) >> 2) - loz + 1;
        if (lane < nx * ny * nz) {
            int bx, by, bz;
            split3(lane, nx, ny, bx, by, bz);
            nax = (bx + lox) * 4 - cx;
            nay = (by + loy) * 4 - cy;
            naz = (bz + loz) * 4 - cz;
            near_have  = true;
            near_key   = pack_cell(bx + lox, by + loy, bz + loz);
            near_first = __ldg(reinterpret_cast<const ulonglong2 *>(H.coarse) + (hash_cell(near_key) & H.hmask));
        }
    }

    // ---- fast exit: no occupied coarse block anywhere within r_max of the query (half of all (ref, query) pairs
    // of a turning sensor): one round of super-mask fetches instead of four empty stages
    {
        const int lox = (cx - r_max) >> 4, loy = (cy - r_max) >> 4, loz = (cz - r_max) >> 4;
        const int nx = ((cx + r_max) >> 4) - lox + 1, ny = ((cy + r_max) >> 4) - loy + 1, nz = ((cz + r_max) >> 4) - loz + 1;
        const int total = nx * ny * nz;
        bool any = false;
#pragma unroll 1
        for (int t0 = 0; t0 < total && !any; t0 += 32) {
            const int t = t0 + lane;
            bool mine   = false;
            if (t < total) {
                int bx, by, bz;
                split3(t, nx, ny, bx, by, bz);
                int sx = (bx + lox) * 16 - cx, sy = (by + loy) * 16 - cy, sz = (bz + loz) * 16 - cz;
                u64 m = lookup_mask(H.super, H.hmask, pack_cell(bx + lox, by + loy, bz + loz));
                mine  = (m & shell_blocks(sx, sy, sz, -1, r_max)) != 0ull;
            }
            any = __any_sync(0xffffffffu, mine);
        }
        if (!any) return;
    }

    int lo = -1, hi = 2; // stage (lo, hi]; the first stage includes the query's own cell (ch = 0)

    // One round: up to kLookupLanes lanes bring one coarse block each (first-cell offset ax, ay, az from the query
    // cell). Cells of the block inside the stage (and inside the pruning bound) go to the queue, the queue is
    // drained 32 at a time.
    // queue the kept cells of every lane's block and probe them 32 at a time
    auto drain = [&](u64 keep, int ax, int ay, int az) {
        const int cnt = __popcll(keep);
        int incl      = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        if (T == 0) return;
        int off = incl - cnt;
        {
            // cell (dx, dy, dz) = block offset + (sub & 3, (sub >> 2) & 3, sub >> 4), 10 bits per axis (biased by 512):
            // the three 2-bit fields of `sub` are spread to the three 10-bit fields and added to the block's code
            const unsigned base = (unsigned) (ax + 512) | ((unsigned) (ay + 512) << 10) | ((unsigned) (az + 512) << 20);
            unsigned half = (unsigned) keep;
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                while (half) {
                    const unsigned sub = (unsigned) (__ffs((int) half) - 1 + 32 * hh);
                    half &= half - 1;
                    queue[off++] = base + ((sub & 3u) | ((sub & 12u) << 8) | ((sub & 48u) << 16));
                }
                half = (unsigned) (keep >> 32);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int k = lane; k < T; k += 32) {
            unsigned e = queue[k];
            int dx = (int) (e & 1023u) - 512, dy = (int) ((e >> 10) & 1023u) - 512, dz = (int) (e >> 20) - 512;
            probe_cell(H, multi, cx + dx, cy + dy, cz + dz, qx, qy, qz, max_d2, a0, a1, a2, a3, a4, dirty);
        }
        __syncwarp();
    };
    // drop the cells of a block mask whose lower-bound distance exceeds the pruning bound
    auto prune_cells = [&](u64 keep, int ax, int ay, int az) {
        u64 m = keep;
        while (m) {
            int sub = __ffsll((long long) m) - 1;
            m &= m - 1;
            if (cells_lb2(abs(ax + (sub & 3)), abs(ay + ((sub >> 2) & 3)), abs(az + (sub >> 4))) > prune_c2) keep &= ~(1ull << sub);
        }
        return keep;
    };
    // after a stage: merge if anything was inserted, test the stop rule, tighten the pruning bound
    auto stage_done = [&]() {
        if (__any_sync(0xffffffffu, dirty)) { // the shuffle merge is skipped when the stage inserted nothing
            warp_merge5(a0, a1, a2, a3, a4, best);
            dirty = false;
        }
        if (best[4] != kInf) {
            float d5    = __uint_as_float((unsigned) (best[4] >> 32));
            float bound = (float) hi * cell - 1e-3f;
            if (bound > 0.f && d5 < bound * bound) return true;
            prune_c2 = fminf(prune_c2, d5 / cell2 * 1.002f);
        }
        return false;
    };

    auto round = [&](bool have, int ax, int ay, int az) {
        u64 keep = 0ull;
        if (have) {
            int gx = max(0, max(ax, -(ax + 3))), gy = max(0, max(ay, -(ay + 3))), gz = max(0, max(az, -(az + 3)));
            if (cells_lb2(gx, gy, gz) <= prune_c2) {
                u64 want = cube_mask(ax, ay, az, hi) & ~cube_mask(ax, ay, az, lo); // pure bit arithmetic
                if (want) keep = lookup_mask(H.coarse, H.hmask, pack_cell((ax + cx) >> 2, (ay + cy) >> 2, (az + cz) >> 2)) & want;
                if (hi >= 4) keep = prune_cells(keep, ax, ay, az); // per-cell pruning pays once the shell is wide
            }
        }
        drain(keep, ax, ay, az);
    };

    // ---- near phase: stages (-1,2] and (2,4] share ONE fetch of the (<= 27) coarse masks around the query; the
    // first probe of each mask was issued before the fast-exit test above, so both round trips overlap
    if (r_max >= 4) {
        u64 m27 = 0ull;
        if (near_have) {
            if (near_first.x == near_key) m27 = near_first.y;
            else if (near_first.x != kEmpty) m27 = lookup_mask(H.coarse, H.hmask, near_key);
        }
        hi = 2;
        drain(m27 & cube_mask(nax, nay, naz, 2), nax, nay, naz);
        if (stage_done()) return;
        lo = 2;
        hi = 4;
        drain(prune_cells(m27 & cube_mask(nax, nay, naz, 4) & ~cube_mask(nax, nay, naz, 2), nax, nay, naz), nax, nay, naz);
        if (stage_done()) return;
        lo = 4;
        hi = 8;
    }

#pragma unroll 1
    while (lo < r_max) {
        hi = min(hi, r_max);
        if (hi < kSparseFrom) {
            // near stages: dense enumeration of the (<= 27) coarse blocks around the query
            const int lox = (cx - hi) >> 2, loy = (cy - hi) >> 2, loz = (cz - hi) >> 2;
            const int nx = ((cx + hi) >> 2) - lox + 1, ny = ((cy + hi) >> 2) - loy + 1, nz = ((cz + hi) >> 2) - loz + 1;
            const int total = nx * ny * nz;
#pragma unroll 1
            for (int t0 = 0; t0 < total; t0 += kLookupLanes) {
                const int t = t0 + lane;
                int bx, by, bz;
                split3(t, nx, ny, bx, by, bz);
                round(lane < kLookupLanes && t < total, (bx + lox) * 4 - cx, (by + loy) * 4 - cy, (bz + loz) * 4 - cz);
            }
        } else {
            // wide stages: the super masks list the occupied coarse blocks, empty space costs one probe per
            // 16^3 cells. Pass 1: lanes fetch super masks and queue the occupied coarse blocks that intersect
            // the stage and the pruning bound. Pass 2: one coarse block per lane and round.
            const int lox = (cx - hi) >> 4, loy = (cy - hi) >> 4, loz = (cz - hi) >> 4;
            const int nx = ((cx + hi) >> 4) - lox + 1, ny = ((cy + hi) >> 4) - loy + 1, nz = ((cz + hi) >> 4) - loz + 1;
            const int total = nx * ny * nz;
#pragma unroll 1
            for (int t0 = 0; t0 < total; t0 += kLookupLanes) {
                const int t = t0 + lane;
                u64 keep    = 0ull;
                int sx = 0, sy = 0, sz = 0; // offset of the super block's first cell from the query cell
                if (lane < kLookupLanes && t < total) {
                    int bx, by, bz;
                    split3(t, nx, ny, bx, by, bz);
                    sx = (bx + lox) * 16 - cx;
                    sy = (by + loy) * 16 - cy;
                    sz = (bz + loz) * 16 - cz;
                    int mn, mx;
                    block_range(sx, sy, sz, 16, mn, mx);
                    if (mx > lo && mn <= hi) {
                        // the occupied coarse blocks that intersect the stage: pure bit arithmetic
                        keep = lookup_mask(H.super, H.hmask, pack_cell(bx + lox, by + loy, bz + loz)) & shell_blocks(sx, sy, sz, lo, hi);
                    }
                }
                const int cnt = __popcll(keep);
                int incl      = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int Tc = __shfl_sync(0xffffffffu, incl, 31); // coarse blocks of these 16 super blocks (<= 1024)
                if (Tc == 0) continue;
                int off = incl - cnt;
                while (keep) {
                    int sub = __ffsll((long long) keep) - 1;
                    keep &= keep - 1;
                    int ax = sx + 4 * (sub & 3), ay = sy + 4 * ((sub >> 2) & 3), az = sz + 4 * (sub >> 4);
                    cqueue[off++] = (unsigned) (ax + 512) | ((unsigned) (ay + 512) << 10) | ((unsigned) (az + 512) << 20);
                }
                __syncwarp();
#pragma unroll 1
                for (int k0 = 0; k0 < Tc; k0 += kLookupLanes) {
                    const int k = k0 + lane;
                    const bool have = lane < kLookupLanes && k < Tc;
                    unsigned e      = have ? cqueue[k] : 0u;
                    round(have, (int) (e & 1023u) - 512, (int) ((e >> 10) & 1023u) - 512, (int) (e >> 20) - 512);
                }
                __syncwarp();
            }
        }
        if (stage_done()) return;
        lo = hi;
        hi *= 2;
    }
}

// K5a: one warp per (ref, query): exact 5-NN only. Writes nn_idx / nn_d2 (ascending, -1 padded).
// Persistent grid of one-warp CTAs: every warp draws (ref, query) tasks from a counter until none is left. Search times
// vary 20x between queries (fast exits against old references, wide searches at the rim of a map); the newest references,
// which overlap the current frame most and cost 3x as much per query, are handed out FIRST so that the long searches
// start at once and the short ones fill the tail. (With a (query, ref) grid the block scheduler walked the references
// oldest first: SMs were active for 52 k of the kernel's 94 k cycles.)
__global__ void __launch_bounds__(kAssocWarps * 32)
associate_knn_kernel(const __grid_constant__ AssocParams P, const float4 *__restrict__ cur_world) {
    extern __shared__ unsigned smem_q[]; // per warp: fine-cell queue, then coarse-block queue
    unsigned *my_queue  = smem_q + (threadIdx.x >> 5) * (kQueueCap + kCQueueCap);
    unsigned *my_cqueue = my_queue + kQueueCap;
    const int lane  = threadIdx.x & 31;
    const int q     = P.q_dev ? min(*P.q_dev, P.q) : P.q;
    const int total = q * P.n_refs;
    int *counter    = P.ticket + 1;              // zero on entry; reset by the plane kernel's last CTA
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) break;
        const int ri = P.n_refs - 1 - t / q;     // newest reference first
        const int k  = t - (t / q) * q;
        const AssocRef &ref = P.ref[ri];
        const long long dbg_t0 = P.dbg_cycles ? clock64() : 0;
        // world -> ref lidar frame (lidar_map.cc:343-345,355-357): fp64 math, fp32 store
        float4 rp   = project_rn(ref.T_ref_world.R, ref.T_ref_world.t, __ldg(cur_world + k));
        u64 best[5] = {kInf, kInf, kInf, kInf, kInf};
        if (ref.m > 0) {
            const HashView &H = ref.index;
            warp_knn5(H, rp.x, rp.y, rp.z, P.inv_cell, P.cell, P.r_max, P.max_d2, my_queue, my_cqueue, best);
        }
        if (lane < 5) {
            u64 b = best[0];
#pragma unroll
            for (int j = 1; j < 5; j++) b = (lane == j) ? best[j] : b;
            bool have = b != kInf;
            ref.feat.nn_idx[5 * k + lane] = have ? (int) (unsigned) (b & 0xffffffffu) : -1;
            ref.feat.nn_d2[5 * k + lane]  = have ? __uint_as_float((unsigned) (b >> 32)) : 0.f;
        }
        if (P.dbg_cycles && lane == 0) P.dbg_cycles[ri * P.q + k] = clock64() - dbg_t0;
        __syncwarp();
    }
}

// K5b: one THREAD per (ref, query): plane fit + validity tests + feature record (lidar_map.cc:371-395).
// Split from the search so that the ~1,500-instruction fp64 QR runs once per lane instead of once per warp.
__global__ void __launch_bounds__(128)
associate_plane_kernel(const __grid_constant__ AssocParams P, const float4 *__restrict__ cur_world,
                       const float4 *__restrict__ cur_lidar) {
    const int ri = blockIdx.y;
    const int k  = blockIdx.x * blockDim.x + threadIdx.x;
    const AssocRef &ref = P.ref[ri];
    int8_t type     = -1;
    uint8_t covered = 0;
    if (k < (P.q_dev ? min(*P.q_dev, P.q) : P.q)) {
        float4 rp = project_rn(ref.T_ref_world.R, ref.T_ref_world.t, __ldg(cur_world + k));
        int idx[5];
#pragma unroll
        for (int j = 0; j < 5; j++) idx[j] = ref.feat.nn_idx[5 * k + j];
        double unit[3] = {0, 0, 0}, ninv = 0;
        if (idx[4] >= 0) { // five neighbours found (ascending, -1 padded)
            covered  = 1;
            float d5 = ref.feat.nn_d2[5 * k + 4];
            if (d5 < P.plane_d2_gate) {
                float4 pts[5];
#pragma unroll
                for (int j = 0; j < 5; j++) pts[j] = __ldg(ref.map + idx[j]);
                double dist[5];
                if (plane_estimation(pts, P.plane_thr, unit, ninv, dist)) {
                    double dis = fabs(((unit[0] * (double) rp.x + unit[1] * (double) rp.y) + unit[2] * (double) rp.z) + ninv);
                    float4 lp  = __ldg(cur_lidar + k);
                    float nrm  = sqrtf((lp.x * lp.x + lp.y * lp.y) + lp.z * lp.z);
                    double scalar = dis / (double) sqrtf(nrm); // lidar_map.cc:384 (fp32 sqrt of the fp32 norm)
                    if (scalar < P.scalar_thr && dis < P.sigma_gate * P.sigma) type = 0;
                }
            }
        }
        ref.feat.type[k]    = type;
        ref.feat.covered[k] = covered;
        ref.feat.outlier[k] = 0; // updateFeature(..., is_outlier=false) lidar_map.cc:394
        ref.feat.normal[3 * k + 0] = unit[0];
        ref.feat.normal[3 * k + 1] = unit[1];
        ref.feat.normal[3 * k + 2] = unit[2];
        ref.feat.d[k] = ninv;
    }
    int nf = __syncthreads_count(type == 0);
    int nc = __syncthreads_count(covered != 0);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        if (nf) atomicAdd(ref.counts + 0, nf);
        if (nc) atomicAdd(ref.counts + 1, nc);
        __threadfence();
        last = atomicAdd(P.ticket, 1) == (int) (gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < 2 * P.n_refs) P.host_counts[threadIdx.x] = atomicExch(P.ref[threadIdx.x >> 1].counts + (threadIdx.x & 1), 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        P.ticket[0] = 0;
        P.ticket[1] = 0;   // task counter of the search kernel
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.signal), "l"(P.seq) : "memory");
    }
}

__global__ void __launch_bounds__(kAssocWarps * 32)
knn5_kernel(MapIndex H, int m, const float4 *__restrict__ query, int q, float inv_cell, float cell, int r_max, float max_d2,
            int32_t *__restrict__ nn_idx, float *__restrict__ nn_d2, int32_t *__restrict__ nn_count) {
    extern __shared__ unsigned smem_q[];
    unsigned *my_queue  = smem_q + (threadIdx.x >> 5) * (kQueueCap + kCQueueCap);
    unsigned *my_cqueue = my_queue + kQueueCap;
    const int lane  = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gwarp >= q) return;
    float4 qp   = __ldg(query + gwarp);
    u64 best[5] = {kInf, kInf, kInf, kInf, kInf};
    if (m > 0) {
        warp_knn5(H, qp.x, qp.y, qp.z, inv_cell, cell, r_max, max_d2, my_queue, my_cqueue, best);
    }
    int found = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) found += (best[j] != kInf) ? 1 : 0;
    if (lane == 0) nn_count[gwarp] = found;
    if (lane < 5) {
        u64 b = best[0];
#pragma unroll
        for (int j = 1; j < 5; j++) b = (lane == j) ? best[j] : b;
        bool have = b != kInf;
        nn_idx[5 * gwarp + lane] = have ? (int) (unsigned) (b & 0xffffffffu) : -1;
        nn_d2[5 * gwarp + lane]  = have ? __uint_as_float((unsigned) (b >> 32)) : 0.f;
    }
}

__global__ void plane_estimation_kernel(const float4 *__restrict__ pts, int n, double thr, double *__restrict__ unit,
                                        double *__restrict__ ninv, double *__restrict__ dist, uint8_t *__restrict__ ok) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p[5];
#pragma unroll
    for (int k = 0; k < 5; k++) p[k] = pts[5 * i + k];
    double u[3], d, ds[5];
    bool r = plane_estimation(p, thr, u, d, ds);
    unit[3 * i] = u[0];
    unit[3 * i + 1] = u[1];
    unit[3 * i + 2] = u[2];
    ninv[i] = d;
#pragma unroll
    for (int k = 0; k < 5; k++) dist[5 * i + k] = ds[k];
    ok[i] = r ? 1 : 0;
}

constexpr size_t kKnnSmem = (size_t) kAssocWarps * (kQueueCap + kCQueueCap) * sizeof(unsigned);

void init_knn_smem() {
    static std::atomic<bool> done[64]; // (zero-initialised; the attribute call is idempotent)
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return;
    cudaFuncSetAttribute(associate_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kKnnSmem);
    cudaFuncSetAttribute(knn5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kKnnSmem);
    done[dev] = true;
}

} // namespace

void launch_hash_build(cudaStream_t s, const float4 *map, int m, float inv_cell, MapIndex index, const int *m_dev) {
    const int hcap = index.hmask + 1;
    int g = (hcap + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    hash_clear_kernel<<<g, 256, 0, s>>>(index);
    if (m > 0) {
        int gi = (m + 255) / 256;
        if (gi > 148 * 8) gi = 148 * 8;
        hash_insert_kernel<<<gi, 256, 0, s>>>(map, m, m_dev, inv_cell, index);
    }
}

void launch_associate(cudaStream_t s, const AssocParams &p, const float4 *cur_world, const float4 *cur_lidar) {
    long long warps = (long long) p.n_refs * p.q;
    if (warps <= 0) return;
    init_knn_smem();
    // one resident wave of one-warp CTAs (8 KB of queues each: 28 per SM), fed from a task counter
    static std::atomic<int> sms[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!sms[dev & 63]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev & 63] = n > 0 ? n : 148;
    }
    const long long wave = (long long) sms[dev & 63].load() * 28 * kAssocWarps;
    const int grid = (int) std::min<long long>((warps + kAssocWarps - 1) / kAssocWarps, wave / kAssocWarps);
    associate_knn_kernel<<<grid, kAssocWarps * 32, kKnnSmem, s>>>(p, cur_world);
    dim3 g2((p.q + 127) / 128, p.n_refs);
    associate_plane_kernel<<<g2, 128, 0, s>>>(p, cur_world, cur_lidar);
}

void launch_knn5(cudaStream_t s, MapIndex index, int m, const float4 *query, int q, float inv_cell, float cell, int r_max,
                 float max_d2, int32_t *nn_idx, float *nn_d2, int32_t *nn_count) {
    if (q <= 0) return;
    init_knn_smem();
    knn5_kernel<<<(q + kAssocWarps - 1) / kAssocWarps, kAssocWarps * 32, kKnnSmem, s>>>(index, m, query, q, inv_cell, cell, r_max,
                                                                                        max_d2, nn_idx, nn_d2, nn_count);
}

void launch_plane_estimation(cudaStream_t s, const float4 *pts, int n, double thr, double *unit, double *ninv,
                             double *dist, uint8_t *ok) {
    if (n <= 0) return;
    plane_estimation_kernel<<<(n + 127) / 128, 128, 0, s>>>(pts, n, thr, unit, ninv, dist, ok);
}

} // namespace fflins
