// k_factor.cu — K6 point-to-plane residual + analytic Jacobians (+ Huber corrector) fused with the
// block reduction into dense 19x19 J^T J / 19 J^T r blocks, and K7 residual-only chi-square culling.
//
// Reference behaviour restated (not translated):
//   LidarFactor::Evaluate                       ff_lins/factors/lidar_factor.cc:41-118
//   ResidualBlockInfo::Evaluate (loss corrector) ff_lins/factors/residual_block_info.cc:49-77
//   MarginalizationInfo::constructSingleEquation ff_lins/factors/marginalization_info.cc:181-210
//   Optimizer::lidarOutlierCullingByChi2         ff_lins/ff_lins/optimizer.cc:515-548
//
// All arithmetic is fp64 (the reference's normal equations are double). B200 design:
//  * everything that depends only on the (ref, cur) pair — three rotation matrices, the first-order time-delay
//    motion terms, the products A = Rbl^T Rt0^T, C = A Rt1, E = C Rbl — is computed once per CTA into shared
//    memory (three threads, two barriers); a thread then needs ~250 fp64 operations per residual instead of the
//    ~1000 of the per-residual formulation;
//  * the reduction is a SYRK out of shared memory on the FP64 tensor pipe (mma.sync.m8n8k4.f64 — tcgen05 has no
//    fp64 kind); warp partials are summed in warp order, the (up to 8) CTAs of a pair leave their 212 partial sums
//    in L2 and the last one to arrive (ticket) adds them in rank order: a fixed summation order, run-to-run
//    deterministic, no cluster barrier on the critical path;
//  * a solver iterates on these blocks 16 times per keyframe and each iteration used to be one kernel launch whose
//    launch -> host-visible latency (~20 us) dwarfed the ~4 us of work. The RESIDENT kernel below removes the launch:
//    one dispatcher CTA polls a pinned host mailbox (16-byte {word, sequence tag} slots, so a message is complete the
//    moment all its tags match — one PCIe round trip), republishes the message in device memory and 8 x 16 worker CTAs
//    pick it up from L2, evaluate and store the blocks straight into pinned host memory behind a system-scope
//    release. The kernel retires by itself after an idle period (default 1 ms), every wait in it is bounded, and
//    the launched kernel (same worker code, message by value) remains as the cold path and for profiling.
#include "math_factor.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fflins {

namespace {

constexpr int kCols      = 24;                 // 19 J | r | rho0 | valid | 1 | 0
constexpr int kColStride = kFactorBlock + 4;   // column-major tile, +4 doubles: conflict-free fragment loads
constexpr int kWarps     = kFactorBlock / 32;
constexpr int kSetupRing = 8;
constexpr unsigned long long kQuitSeq = ~0ull;

// message words (kernels.h)
enum { W_CMD = 0, W_NP = 1, W_TD = 2, W_THR = 3, W_SEQ = 4, W_OUT = 5, W_CNT = 6, W_SIG = 7, W_CUR = 8, W_EXT = 15, W_REF = kMsgHead };
static_assert(kMsgHead == 22, "message layout");

// D (8x8, fp64) += A (8x4, row) * B (4x8, col) on the FP64 tensor pipe.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// The CTA's view of one evaluation: header, pair constants and the pair's feature arrays, in shared memory.
struct PairShared {
    CurConst cc;
    RefConst rc;
    FactorSetupPair sp;
    const float *points;
    double *out_red;
    int *host_counts;
    unsigned long long *signal;
    unsigned long long seq;
    double td, thr, sigma, huber;
    unsigned cmd, flags;
    int n_pairs, q, stride4, last, count;
};

struct WorkBuffers {
    double *partials;            // [kMaxRefs][kFactorRanks][kRedSize]
    int *tickets;                // [kMaxRefs], zero on entry / exit
    int *counts;                 // [kMaxRefs], zero on entry / exit
    unsigned long long *acks;    // resident kernel only: one increment per CTA per message, once the message is consumed
};

constexpr int kTileDoubles   = kCols * kColStride + kRedSize + 4;
constexpr size_t kFactorSmem = (size_t) kTileDoubles * sizeof(double) + sizeof(PairShared);

__device__ __forceinline__ double w2d(unsigned long long w) { return __longlong_as_double((long long) w); }

// One evaluation (or chi-square pass) of pair `pair` by CTA `rank` of `ranks`. msg: kMsgWords words (kernel parameter
// space or global memory); setup: device memory.
__device__ void factor_work(const unsigned long long *__restrict__ msg, const FactorSetup *__restrict__ setup, int pair,
                            int rank, int ranks, double *smem, const WorkBuffers &wb, double *__restrict__ r_out,
                            double *__restrict__ J_out, uint8_t *__restrict__ valid_out) {
    double *cols   = smem;                          // [kCols][kColStride]
    double *s_red  = smem + kCols * kColStride;     // [kRedSize]
    PairShared &sh = *reinterpret_cast<PairShared *>(smem + kTileDoubles);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- header + pair constants (lidar_factor.cc:43-69): three threads in parallel, then the dependent products
    if (tid == 0) {
        const unsigned long long w0 = msg[W_CMD], w1 = msg[W_NP];
        sh.cmd     = (unsigned) w0;
        sh.flags   = (unsigned) (w0 >> 32);
        sh.n_pairs = (int) (unsigned) w1;
        sh.td      = w2d(msg[W_TD]);
        sh.thr     = w2d(msg[W_THR]);
        sh.seq     = msg[W_SEQ];
        sh.out_red     = reinterpret_cast<double *>(msg[W_OUT]);
        sh.host_counts = reinterpret_cast<int *>(msg[W_CNT]);
        sh.signal      = reinterpret_cast<unsigned long long *>(msg[W_SIG]);
        sh.count       = 0;
        double pc[7], ex[7];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            pc[i] = w2d(msg[W_CUR + i]);
            ex[i] = w2d(msg[W_EXT + i]);
        }
        make_cur_const_raw(pc, ex, sh.td, setup->v1, setup->w1, setup->td1, sh.cc);
    } else if (tid == 32) {
        sh.sp = setup->pair[pair];
        double pr[7];
#pragma unroll
        for (int i = 0; i < 7; i++) pr[i] = w2d(msg[W_REF + 7 * pair + i]);
        make_ref_const_a(pr, sh.sp.v0, sh.sp.w0, sh.sp.td0, w2d(msg[W_TD]), setup->v1, sh.rc);
    } else if (tid == 64) {
        sh.points  = setup->points;
        sh.stride4 = setup->stride4;
        sh.q       = setup->q;
        sh.sigma   = setup->sigma;
        sh.huber   = setup->huber_delta;
    }
    __syncthreads();
    if (tid == 0) {
        if (wb.acks) atomicAdd(wb.acks, 1ull);       // message and setup are consumed: the dispatcher may move on
        make_ref_const_b(sh.cc, sh.rc);
    }
    if (pair >= sh.n_pairs) return;
    const int q     = sh.q;
    const int tiles = (q + kFactorBlock - 1) / kFactorBlock;
    const int C     = tiles < ranks ? (tiles > 0 ? tiles : 1) : ranks;  // CTAs of this pair that hold work
    if (rank >= C) return;
    __syncthreads();
    const FactorSetupPair &pr = sh.sp;
    const float *__restrict__ points = sh.points;
    const int pstride = sh.stride4 ? 4 : 3;

    if (sh.cmd == kCmdChi2) {
        // ---- K7 (optimizer.cc:515-548): residual only, r^2 > threshold -> outlier
        int mine = 0;
        for (int tile = rank; tile < tiles; tile += C) {
            const int k = tile * kFactorBlock + tid;
            bool valid  = k < q;
            if (valid && pr.type) valid = pr.type[k] == 0 && pr.outlier[k] == 0;
            if (valid) {
                const float *pp = points + (size_t) k * pstride;
                V3 p = v3((double) pp[0], (double) pp[1], (double) pp[2]);
                V3 n = v3(pr.normal[3 * k], pr.normal[3 * k + 1], pr.normal[3 * k + 2]);
                double s = 1.0 / (pr.std ? pr.std[k] : sh.sigma);
                double r = eval_residual<false>(sh.cc, sh.rc, p, n, pr.d[k], s, nullptr);
                if (r * r > sh.thr) {       // optimizer.cc:536-541
                    if (pr.outlier) pr.outlier[k] = 1;
                    mine++;
                }
            }
        }
        if (mine) atomicAdd(&sh.count, mine);
        __syncthreads();
        if (tid == 0) {
            if (sh.count) atomicAdd(wb.counts + pair, sh.count);
            __threadfence();
            const int t = atomicAdd(wb.tickets + pair, 1);
            if (t == C - 1) {
                __threadfence();
                sh.host_counts[pair] = atomicExch(wb.counts + pair, 0);
                wb.tickets[pair]     = 0;
                st_release_sys(sh.signal + pair, sh.seq);
            }
        }
        return;
    }

    // ---- K6: residual + Jacobians, SYRK on the FP64 tensor pipe. The tile's rows v = [J(19) | r | rho0 | valid | 1 | 0]
    // are stored column-major in shared memory, G = V^T V (24 x 24) is split into 3 x 3 blocks of 8 x 8 and warp w
    // accumulates the 6 upper blocks over its 32 rows with mma.sync.m8n8k4.f64: the A fragment of column group g (lane l
    // holds V[k0 + l%4][8g + l/4]) is also the B fragment of the same group, so one k-step is 3 shared-memory loads and 6
    // DMMAs per lane. sum rho and the residual count come out of the same product as G[20][22] and G[21][22] (column 22
    // is all ones).
    double acc[6][2];                              // blocks (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
#pragma unroll
    for (int a = 0; a < 6; a++) acc[a][0] = acc[a][1] = 0.0;
    const bool reduce    = sh.out_red != nullptr;
    const unsigned flags = sh.flags;
    for (int tile = rank; tile < tiles; tile += C) {
        const int k = tile * kFactorBlock + tid;
        bool valid  = k < q;
        if (valid && pr.type) valid = (pr.type[k] == 0) && (pr.outlier[k] == 0); // optimizer.cc:468
        double J[19], r = 0.0, rho0 = 0.0;
#pragma unroll
        for (int i = 0; i < 19; i++) J[i] = 0.0;
        if (valid) {
            const float *pp = points + (size_t) k * pstride;
            V3 p = v3((double) pp[0], (double) pp[1], (double) pp[2]);
            V3 n = v3(pr.normal[3 * k], pr.normal[3 * k + 1], pr.normal[3 * k + 2]);
            double s = 1.0 / (pr.std ? pr.std[k] : sh.sigma); // sqrt_info (lidar_factor.cc:38)
            r = eval_residual<true>(sh.cc, sh.rc, p, n, pr.d[k], s, J);
            if (flags & 2u) { J[0] = J[1] = J[2] = J[3] = J[4] = J[5] = 0.0; }
            if (flags & 4u) { J[6] = J[7] = J[8] = J[9] = J[10] = J[11] = 0.0; }
            if (flags & 8u) { J[12] = J[13] = J[14] = J[15] = J[16] = J[17] = 0.0; }
            if (flags & 16u) J[18] = 0.0;
            if (flags & 1u) rho0 = huber_correct(sh.huber, r, J);
            else rho0 = r * r;
        }
        if (k < q) {
            size_t o = (size_t) pair * q + k;
            if (valid_out) valid_out[o] = valid ? 1 : 0;
            if (r_out) r_out[o] = r;
            if (J_out) {
                double *Jo = J_out + 22 * o;   // 1x7, 1x7, 1x7, 1x1 row-major; column 6 of each pose block is zero
#pragma unroll
                for (int b = 0; b < 3; b++) {
#pragma unroll
                    for (int i = 0; i < 6; i++) Jo[7 * b + i] = J[6 * b + i];
                    Jo[7 * b + 6] = 0.0;
                }
                Jo[21] = J[18];
            }
        }
        if (!reduce) continue;
#pragma unroll
        for (int i = 0; i < 19; i++) cols[i * kColStride + tid] = J[i];
        cols[19 * kColStride + tid] = r;
        cols[20 * kColStride + tid] = rho0;
        cols[21 * kColStride + tid] = valid ? 1.0 : 0.0;
        cols[22 * kColStride + tid] = 1.0;
        cols[23 * kColStride + tid] = 0.0;
        __syncwarp(); // warp w consumes exactly the 32 rows its own lanes produced
        {
            const double *base = cols + (lane >> 2) * kColStride + 32 * warp + (lane & 3);
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const double f0 = base[4 * ks], f1 = base[8 * kColStride + 4 * ks], f2 = base[16 * kColStride + 4 * ks];
                dmma884(acc[0], f0, f0);
                dmma884(acc[1], f0, f1);
                dmma884(acc[2], f0, f2);
                dmma884(acc[3], f1, f1);
                dmma884(acc[4], f1, f2);
                dmma884(acc[5], f2, f2);
            }
        }
        __syncwarp();
    }
    if (!reduce) return;
    // warp partials -> shared memory (each warp's rows are its own, but other warps' regions are overwritten:
    // wait for everyone): part[warp][block][row 0..7][col 0..7]
    __syncthreads();
    double *part = smem;
    {
        const int row = lane >> 2, col = 2 * (lane & 3);
#pragma unroll
        for (int a = 0; a < 6; a++)
            *reinterpret_cast<double2 *>(part + ((warp * 6 + a) * 8 + row) * 8 + col) = make_double2(acc[a][0], acc[a][1]);
    }
    __syncthreads();
    // thread e (and e + 128, e + 256 < 384) sums entry e = (block, row, col) over the warps in warp order and files it
    // under its place in the 212-vector [190 upper-triangle H | 19 b | sum rho | sum r'^2 | count]
    for (int e0 = tid; e0 < 6 * 64; e0 += kFactorBlock) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) sum += part[w * 6 * 64 + e0];
        const int blk = e0 >> 6, gi = blk < 3 ? 0 : (blk < 5 ? 1 : 2), gj = blk < 3 ? blk : (blk < 5 ? blk - 2 : 2);
        const int ra = 8 * gi + ((e0 >> 3) & 7), cb = 8 * gj + (e0 & 7);
        int e = -1;
        double sign = 1.0;
        if (ra < 19 && cb < 19) {
            if (ra <= cb) e = ra * 19 - (ra * (ra - 1)) / 2 + (cb - ra);          // upper triangle of H
        } else if (ra < 19 && cb == 19) {
            e    = 190 + ra;                                                     // b = -J^T r
            sign = -1.0;
        } else if (ra == 19 && cb == 19) {
            e = 210;                                                             // sum r'^2
        } else if (ra == 20 && cb == 22) {
            e = 209;                                                             // sum rho
        } else if (ra == 21 && cb == 22) {
            e = 211;                                                             // count
        }
        if (e >= 0) s_red[e] = sign * sum;   // s_red lies behind the tile buffer, `part` stays intact
    }
    __syncthreads();
    double *out = sh.out_red + (size_t) pair * kRedSize;
    if (C == 1) {
        for (int e = tid; e < kRedSize; e += kFactorBlock) out[e] = s_red[e];
    } else {
        // cross-CTA reduction through L2: every CTA leaves its 212 sums, the last one to take a ticket adds them in
        // rank order (threadfence-reduction pattern; fixed order = deterministic)
        double *mine = wb.partials + ((size_t) pair * kFactorRanks + rank) * kRedSize;
        for (int e = tid; e < kRedSize; e += kFactorBlock) mine[e] = s_red[e];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int t = atomicAdd(wb.tickets + pair, 1);
            sh.last     = t == C - 1;
            if (sh.last) __threadfence();
        }
        __syncthreads();
        if (!sh.last) return;
        const double *all = wb.partials + (size_t) pair * kFactorRanks * kRedSize;
        for (int e = tid; e < kRedSize; e += kFactorBlock) {
            double v[kFactorRanks];
#pragma unroll
            for (int rk = 0; rk < kFactorRanks; rk++) v[rk] = rk < C ? __ldcg(all + (size_t) rk * kRedSize + e) : 0.0;
            double total = 0.0;
#pragma unroll
            for (int rk = 0; rk < kFactorRanks; rk++) total += v[rk];
            out[e] = total;
        }
        if (tid == 0) wb.tickets[pair] = 0;
    }
    if (sh.signal) {
        // out_red is host memory: publish "pair done" behind the data so that the host can poll for it instead of
        // paying a stream synchronisation (the release at system scope orders the CTA's stores before the flag)
        __syncthreads();
        if (tid == 0) st_release_sys(sh.signal + pair, sh.seq);
    }
}

// Launched form: grid (kFactorRanks or fewer, n_pairs), message by value.
__global__ void __launch_bounds__(kFactorBlock)
factor_kernel(const __grid_constant__ FactorMsg M, const FactorSetup *__restrict__ setup, WorkBuffers wb,
              double *__restrict__ r_out, double *__restrict__ J_out, uint8_t *__restrict__ valid_out) {
    extern __shared__ __align__(16) double smem[];
    factor_work(M.w, setup, blockIdx.y, blockIdx.x, gridDim.x, smem, wb, r_out, J_out, valid_out);
}

// ---- resident form ------------------------------------------------------------------------------------------------
struct ResidentArgs {
    const ulonglong2 *h_slots;          // pinned mailbox: kMsgWords x {word, tag}
    const FactorSetup *h_setup;         // pinned ring [kSetupRing]
    unsigned long long *h_state;        // pinned: [0] last sequence number dispatched, [1] launch id (written at exit)
    FactorMsg *d_msg;
    FactorSetup *d_setup;
    unsigned long long *d_seq;          // dispatch counter (kQuitSeq = retire)
    unsigned long long start_seq, launch_id, idle_ns;
    WorkBuffers wb;
};

__device__ void resident_dispatch(const ResidentArgs &A, double *smem) {
    unsigned long long *s_msg = reinterpret_cast<unsigned long long *>(smem);
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    unsigned long long expect = A.start_seq + 1, acks_expected = 0, cached_gen = ~0ull;
    const unsigned long long n_workers = (unsigned long long) kFactorRanks * kMaxRefs;
    unsigned long long t_last = globaltimer_ns();
    while (true) {
        // one PCIe round trip per poll: every thread fetches its own {word, tag} slots; the message is complete when
        // all tags carry the expected sequence number (the host stores word before tag, 16-byte slots never tear)
        bool ok = true;
        for (int slot = tid; slot < kMsgWords; slot += kFactorBlock) {
            unsigned long long w, t;
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w), "=l"(t) : "l"(A.h_slots + slot) : "memory");
            ok &= t == expect;
            s_msg[slot] = w;
        }
        if (!__syncthreads_and(ok)) {
            if (tid == 0) s_flag = globaltimer_ns() - t_last > A.idle_ns;
            __syncthreads();
            if (s_flag) break;
            continue;
        }
        // the workers must be done reading the previous message (long since, normally)
        if (tid == 0) {
            int bad = 0;
            const unsigned long long t0 = globaltimer_ns();
            while (ld_acquire_gpu(A.wb.acks) < acks_expected)
                if (globaltimer_ns() - t0 > 4 * A.idle_ns + 1000000ull) {
                    bad = 1;
                    break;
                }
            s_flag = bad;
        }
        __syncthreads();
        if (s_flag) break;
        const unsigned cmd           = (unsigned) s_msg[W_CMD];
        const unsigned long long gen = s_msg[W_NP] >> 32;
        if (cmd == kCmdQuit) {
            expect++;   // (acknowledged as dispatched)
            break;
        }
        if (gen != cached_gen) { // new keyframe: fetch the setup from the host ring (one more round trip, once per keyframe)
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(A.h_setup + (gen % kSetupRing));
            unsigned long long *dst       = reinterpret_cast<unsigned long long *>(A.d_setup);
            for (int i = tid; i < (int) (sizeof(FactorSetup) / 8); i += kFactorBlock) {
                unsigned long long v;
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(src + i) : "memory");
                dst[i] = v;
            }
            cached_gen = gen;
        }
        for (int i = tid; i < kMsgWords; i += kFactorBlock) A.d_msg->w[i] = s_msg[i];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(A.d_seq, expect);
        acks_expected += n_workers;
        expect++;
        t_last = globaltimer_ns();
    }
    // retire: tell the workers, then the host (which falls back to a launch for anything not dispatched)
    if (tid == 0) {
        st_release_gpu(A.d_seq, kQuitSeq);
        A.h_state[0] = expect - 1;
        __threadfence_system();
        st_release_sys(A.h_state + 1, A.launch_id);
    }
}

__global__ void __launch_bounds__(kFactorBlock) resident_kernel(const __grid_constant__ ResidentArgs A) {
    extern __shared__ __align__(16) double smem[];
    if (blockIdx.y == kMaxRefs) {
        if (blockIdx.x == 0) resident_dispatch(A, smem);
        return;
    }
    __shared__ unsigned long long s_seq;
    unsigned long long last = A.start_seq;
    while (true) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned long long s;
            while ((s = ld_acquire_gpu(A.d_seq)) == last) {
                if (globaltimer_ns() - t0 > 8 * A.idle_ns + 20000000ull) { // the dispatcher is gone: never spin forever
                    s = kQuitSeq;
                    break;
                }
            }
            s_seq = s;
        }
        __syncthreads();
        const unsigned long long s = s_seq;
        if (s == kQuitSeq) return;
        factor_work(A.d_msg->w, A.d_setup, blockIdx.y, blockIdx.x, kFactorRanks, smem, A.wb, nullptr, nullptr, nullptr);
        last = s;
        __syncthreads();
    }
}

__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;
}

} // namespace

// -------------------------------------------------------------------------------------------------------------------
struct FactorRuntime {
    int device = 0;
    // setup ring (pinned + device), generation = number of distinct setups seen
    FactorSetup *h_setup = nullptr, *d_setup = nullptr;
    FactorSetup cached;
    bool have_cached = false;
    unsigned long long gen = 0;
    WorkBuffers wb{};
    // resident kernel
    cudaStream_t res_stream = nullptr;
    ulonglong2 *h_slots = nullptr;
    unsigned long long *h_state = nullptr;
    FactorMsg *d_msg = nullptr;
    FactorSetup *d_setup_res = nullptr;
    unsigned long long *d_seq = nullptr;
    unsigned long long mseq = 0, launch_id = 0, idle_ns = 1000000ull;
    bool running = false, enabled = true;
    int strikes = 0;
    unsigned long long evals_this_launch = 0;
    FactorStats st{};
};

void factor_runtime_destroy(FactorRuntime *rt);

FactorRuntime *factor_runtime_create(int device) {
    FactorRuntime *rt = new FactorRuntime();
    rt->device        = device;
    bool ok = cudaMallocHost((void **) &rt->h_setup, sizeof(FactorSetup) * kSetupRing) == cudaSuccess &&
              cudaMalloc((void **) &rt->d_setup, sizeof(FactorSetup) * kSetupRing) == cudaSuccess &&
              cudaMalloc((void **) &rt->wb.partials, sizeof(double) * kMaxRefs * kFactorRanks * kRedSize) == cudaSuccess &&
              cudaMalloc((void **) &rt->wb.tickets, sizeof(int) * 2 * kMaxRefs + 16) == cudaSuccess &&
              cudaMallocHost((void **) &rt->h_slots, sizeof(ulonglong2) * kMsgWords) == cudaSuccess &&
              cudaMallocHost((void **) &rt->h_state, 64) == cudaSuccess &&
              cudaMalloc((void **) &rt->d_msg, sizeof(FactorMsg)) == cudaSuccess &&
              cudaMalloc((void **) &rt->d_setup_res, sizeof(FactorSetup)) == cudaSuccess &&
              cudaMalloc((void **) &rt->d_seq, 64) == cudaSuccess &&
              cudaStreamCreateWithFlags(&rt->res_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        rt->wb.counts = rt->wb.tickets + kMaxRefs;
        rt->wb.acks   = nullptr;
        ok = cudaMemset(rt->wb.tickets, 0, sizeof(int) * 2 * kMaxRefs + 16) == cudaSuccess &&
             cudaMemset(rt->d_seq, 0, 64) == cudaSuccess;
        std::memset(rt->h_slots, 0, sizeof(ulonglong2) * kMsgWords);
        std::memset(rt->h_state, 0, 64);
        ok = ok && cudaFuncSetAttribute(factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kFactorSmem) == cudaSuccess &&
             cudaFuncSetAttribute(resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kFactorSmem) == cudaSuccess;
    }
    if (!ok) {
        cudaGetLastError();
        factor_runtime_destroy(rt);
        return nullptr;
    }
    if (getenv("FFLINS_NO_RESIDENT") || getenv("FFLINS_NO_POLL")) rt->enabled = false;
    if (const char *e = getenv("FFLINS_RESIDENT_IDLE_US")) {
        long v = atol(e);
        if (v >= 10 && v <= 1000000) rt->idle_ns = (unsigned long long) v * 1000ull;
    }
    return rt;
}

namespace {

using Clock = std::chrono::steady_clock;

void post_message(FactorRuntime *rt, const FactorMsg &m, unsigned long long tag) {
    // word before tag; x86-64 keeps the two stores in order and a 16-byte slot is fetched by the device in one piece
    volatile unsigned long long *p = reinterpret_cast<volatile unsigned long long *>(rt->h_slots);
    for (int i = 0; i < kMsgWords; i++) {
        p[2 * i] = m.w[i];
        __atomic_store_n(const_cast<unsigned long long *>(p + 2 * i + 1), tag, __ATOMIC_RELEASE);
    }
}

bool resident_exited(const FactorRuntime *rt) {
    return __atomic_load_n(rt->h_state + 1, __ATOMIC_ACQUIRE) == rt->launch_id;
}

cudaError_t resident_start(FactorRuntime *rt) {
    ResidentArgs A{};
    A.h_slots   = rt->h_slots;
    A.h_setup   = rt->h_setup;
    A.h_state   = rt->h_state;
    A.d_msg     = rt->d_msg;
    A.d_setup   = rt->d_setup_res;
    A.d_seq     = rt->d_seq;
    A.start_seq = rt->mseq;
    A.launch_id = ++rt->launch_id;
    A.idle_ns   = rt->idle_ns;
    A.wb        = rt->wb;
    A.wb.acks   = rt->d_seq + 1;
    unsigned long long init[2] = {rt->mseq, 0ull};
    cudaError_t e = cudaMemcpyAsync(rt->d_seq, init, sizeof(init), cudaMemcpyHostToDevice, rt->res_stream);
    if (e != cudaSuccess) return e;
    resident_kernel<<<dim3(kFactorRanks, kMaxRefs + 1), kFactorBlock, kFactorSmem, rt->res_stream>>>(A);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    rt->running           = true;
    rt->evals_this_launch = 0;
    rt->st.resident_launches++;
    return cudaSuccess;
}

// The resident kernel has retired (or is about to): account for it.
void resident_reap(FactorRuntime *rt) {
    cudaStreamSynchronize(rt->res_stream);
    rt->running = false;
    if (rt->evals_this_launch == 0) {
        if (++rt->strikes >= 3) rt->enabled = false; // e.g. under a profiler that serialises kernels: stop trying
    } else {
        rt->strikes = 0;
    }
}

// Post `m` to the resident kernel and wait for flags[0..n) = seq. Returns 0 done, 1 not handled (caller launches), < 0 error.
int resident_call(FactorRuntime *rt, FactorMsg &m, const unsigned long long *flags, int n, unsigned long long seq) {
    if (!rt->enabled) return 1;
    if (rt->running && resident_exited(rt)) resident_reap(rt);
    if (!rt->enabled) return 1;
    if (!rt->running) {
        cudaError_t e = resident_start(rt);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rt->enabled = false;
            return 1;
        }
    }
    const unsigned long long tag = ++rt->mseq;
    post_message(rt, m, tag);
    unsigned spins = 0;
    Clock::time_point t0;
    bool timing = false;
    for (int i = 0; i < n; i++) {
        while (__atomic_load_n(flags + i, __ATOMIC_ACQUIRE) != seq) {
            if ((++spins & 0x3ffu) == 0) {
                if (resident_exited(rt)) {
                    const unsigned long long done = rt->h_state[0];
                    if (done >= tag) rt->evals_this_launch++;
                    resident_reap(rt); // (joins the kernel: every store it made is visible now)
                    bool all = done >= tag;
                    for (int j = 0; j < n && all; j++) all = __atomic_load_n(flags + j, __ATOMIC_ACQUIRE) == seq;
                    if (all) {
                        rt->st.resident_evals++;
                        return 0;
                    }
                    rt->st.resident_fallbacks++;
                    return 1;
                }
                if (!timing) {
                    t0     = Clock::now();
                    timing = true;
                } else if (Clock::now() - t0 > std::chrono::seconds(2)) {
                    cudaError_t e = cudaStreamQuery(rt->res_stream);
                    rt->enabled   = false;
                    return e != cudaSuccess && e != cudaErrorNotReady ? -(int) e : -(int) cudaErrorLaunchTimeout;
                }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    rt->evals_this_launch++;
    rt->st.resident_evals++;
    return 0;
}

void fill_setup(const FactorParams &p, const float *points, int stride4, FactorSetup &s) {
    std::memset(&s, 0, sizeof(s));
    s.points  = points;
    s.stride4 = stride4;
    s.q       = p.q;
    std::memcpy(s.v1, p.v1, sizeof(s.v1));
    std::memcpy(s.w1, p.w1, sizeof(s.w1));
    s.td1         = p.td1;
    s.sigma       = p.sigma;
    s.huber_delta = p.huber_delta;
    for (int i = 0; i < p.n_pairs; i++) {
        FactorSetupPair &o  = s.pair[i];
        const FactorPair &f = p.pair[i];
        o.type    = f.type;
        o.outlier = f.outlier;
        o.normal  = f.normal;
        o.d       = f.d;
        o.std     = f.std;
        std::memcpy(o.v0, f.v0, sizeof(o.v0));
        std::memcpy(o.w0, f.w0, sizeof(o.w0));
        o.td0 = f.td0;
    }
}

// The setup of this evaluation in the ring (uploaded on `s` when it differs from the previous one).
cudaError_t push_setup(FactorRuntime *rt, cudaStream_t s, const FactorSetup &su) {
    if (rt->have_cached && std::memcmp(&rt->cached, &su, sizeof(su)) == 0) return cudaSuccess;
    rt->cached      = su;
    rt->have_cached = true;
    const unsigned long long g = ++rt->gen;
    FactorSetup *h             = rt->h_setup + (g % kSetupRing);
    *h                         = su;
    return cudaMemcpyAsync(rt->d_setup + (g % kSetupRing), h, sizeof(su), cudaMemcpyHostToDevice, s);
}

unsigned long long d2w(double d) {
    unsigned long long w;
    std::memcpy(&w, &d, 8);
    return w;
}

void fill_msg(const FactorRuntime *rt, const FactorParams &p, unsigned cmd, double thr, unsigned long long seq, double *out_red,
              int *host_counts, unsigned long long *signal, FactorMsg &m) {
    m.w[W_CMD] = (unsigned long long) cmd | ((unsigned long long) p.flags << 32);
    m.w[W_NP]  = (unsigned long long) (unsigned) p.n_pairs | (rt->gen << 32);
    m.w[W_TD]  = d2w(p.td);
    m.w[W_THR] = d2w(thr);
    m.w[W_SEQ] = seq;
    m.w[W_OUT] = (unsigned long long) (uintptr_t) out_red;
    m.w[W_CNT] = (unsigned long long) (uintptr_t) host_counts;
    m.w[W_SIG] = (unsigned long long) (uintptr_t) signal;
    for (int i = 0; i < 7; i++) {
        m.w[W_CUR + i] = d2w(p.pose_cur[i]);
        m.w[W_EXT + i] = d2w(p.ext[i]);
    }
    for (int j = 0; j < kMaxRefs; j++)
        for (int i = 0; i < 7; i++) m.w[W_REF + 7 * j + i] = j < p.n_pairs ? d2w(p.pair[j].pose_ref[i]) : 0ull;
}

int ranks_for(int q) {
    int tiles = (q + kFactorBlock - 1) / kFactorBlock;
    return tiles < 1 ? 1 : (tiles < kFactorRanks ? tiles : kFactorRanks);
}

} // namespace

void factor_runtime_quiesce(FactorRuntime *rt) {
    if (!rt || !rt->running) return;
    if (!resident_exited(rt)) {
        FactorMsg m{};
        m.w[W_CMD] = kCmdQuit;
        post_message(rt, m, ++rt->mseq);
    }
    cudaStreamSynchronize(rt->res_stream);
    rt->running = false;
}

void factor_runtime_destroy(FactorRuntime *rt) {
    if (!rt) return;
    factor_runtime_quiesce(rt);
    if (rt->res_stream) cudaStreamDestroy(rt->res_stream);
    if (rt->h_setup) cudaFreeHost(rt->h_setup);
    if (rt->h_slots) cudaFreeHost(rt->h_slots);
    if (rt->h_state) cudaFreeHost(rt->h_state);
    cudaFree(rt->d_setup);
    cudaFree(rt->wb.partials);
    cudaFree(rt->wb.tickets);
    cudaFree(rt->d_msg);
    cudaFree(rt->d_setup_res);
    cudaFree(rt->d_seq);
    delete rt;
}

void factor_runtime_stats(const FactorRuntime *rt, FactorStats *out) {
    if (rt && out) *out = rt->st;
}

int launch_factor_eval(FactorRuntime *rt, cudaStream_t s, const FactorParams &p, const float *points, int stride4, double *r,
                       double *J, uint8_t *valid, double *out_red, unsigned long long *signal, unsigned long long seq,
                       bool resident, int64_t *launches) {
    if (p.n_pairs <= 0 || p.q <= 0) return 0;
    FactorSetup su;
    fill_setup(p, points, stride4, su);
    cudaError_t e = push_setup(rt, s, su);
    if (e != cudaSuccess) return -(int) e;
    FactorMsg m;
    fill_msg(rt, p, kCmdEval, 0.0, seq, out_red, nullptr, signal, m);
    if (resident && signal && out_red && !r && !J && !valid) {
        int rc = resident_call(rt, m, signal, p.n_pairs, seq);
        if (rc <= 0) return rc;
    }
    factor_kernel<<<dim3(ranks_for(p.q), p.n_pairs), kFactorBlock, kFactorSmem, s>>>(m, rt->d_setup + (rt->gen % kSetupRing), rt->wb, r, J,
                                                                                      valid);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -(int) e;
    rt->st.launched_evals++;
    if (launches) (*launches)++;
    return 1;
}

int launch_chi2(FactorRuntime *rt, cudaStream_t s, const FactorParams &p, const float *points, int stride4, double threshold,
                int *host_counts, unsigned long long *signal, unsigned long long seq, bool resident, int64_t *launches) {
    if (p.n_pairs <= 0 || p.q <= 0) return 0;
    FactorSetup su;
    fill_setup(p, points, stride4, su);
    cudaError_t e = push_setup(rt, s, su);
    if (e != cudaSuccess) return -(int) e;
    FactorMsg m;
    fill_msg(rt, p, kCmdChi2, threshold, seq, nullptr, host_counts, signal, m);
    if (resident) {
        int rc = resident_call(rt, m, signal, p.n_pairs, seq);
        if (rc <= 0) return rc;
    }
    factor_kernel<<<dim3(ranks_for(p.q), p.n_pairs), kFactorBlock, kFactorSmem, s>>>(m, rt->d_setup + (rt->gen % kSetupRing), rt->wb,
                                                                                      nullptr, nullptr, nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return -(int) e;
    rt->st.launched_evals++;
    if (launches) (*launches)++;
    return 1;
}

double measure_fp64_tflops(cudaStream_t s) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr;
    if (cudaMalloc((void **) &d, 64) != cudaSuccess) return 0.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int iters = 1 << 14, blocks = sms * 8, threads = 256;
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, 256); // warm
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, s);
        fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, iters);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 8.0 * (double) iters * blocks * threads;
        if (ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    return best;
}

} // namespace fflins
