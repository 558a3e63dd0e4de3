// k_factor.cu — K6 point-to-plane residual + analytic Jacobians (+ Huber corrector) fused with the
// block reduction into dense 19x19 J^T J / 19 J^T r blocks, and K7 residual-only chi-square culling.
//
// Reference behaviour restated (not translated):
//   LidarFactor::Evaluate                       ff_lins/factors/lidar_factor.cc:41-118
//   ResidualBlockInfo::Evaluate (loss corrector) ff_lins/factors/residual_block_info.cc:49-77
//   MarginalizationInfo::constructSingleEquation ff_lins/factors/marginalization_info.cc:181-210
//   Optimizer::lidarOutlierCullingByChi2         ff_lins/ff_lins/optimizer.cc:515-548
//
// All arithmetic is fp64 (the reference's normal equations are double). B200 design: everything
// that depends only on the (ref, cur) pair — three rotation matrices, the first-order time-delay
// motion terms, the products A = Rbl^T Rt0^T, C = A Rt1, E = C Rbl — is computed once per CTA into
// shared memory; a thread then needs ~250 fp64 operations per residual instead of the ~1000 of the
// per-residual formulation. The reduction is a 256-row SYRK out of shared memory on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 — tcgen05 has no fp64 kind), warp partials and CTA partials are summed in a
// fixed order, so results are run-to-run deterministic.
#include "math_factor.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace fflins {

namespace {

constexpr int kCols      = 24;                 // 19 J | r | rho0 | valid | 1 | 0
constexpr int kColStride = kFactorBlock + 4;   // column-major tile, +4 doubles: conflict-free fragment loads
constexpr int kWarps     = kFactorBlock / 32;

// D (8x8, fp64) += A (8x4, row) * B (4x8, col) on the FP64 tensor pipe.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// Grid: (C, n_pairs) with the C CTAs of a pair forming ONE thread-block cluster (C = 1, 2, 4 or 8).
// CTA c handles residual tiles c, c + C, ... Reduction = a SYRK on the FP64 tensor pipe: the tile's rows
// v = [J(19) | r | rho0 | valid | 1 | 0] are stored column-major in shared memory, G = V^T V (24 x 24) is
// split into 3 x 3 blocks of 8 x 8 and warp w accumulates the 6 upper blocks over its 32 rows with
// mma.sync.m8n8k4.f64: the A fragment of column group g (lane l holds V[k0 + l%4][8g + l/4]) is also the B
// fragment of the same group, so one k-step is 3 shared-memory loads and 6 DMMAs per lane (the scalar
// version needed 4 x 16-byte loads per 16 FMAs and was shared-memory-bandwidth bound). sum rho and the
// residual count come out of the same product as G[20][22] and G[21][22] (column 22 is all ones).
// The 8 warp partials are summed in warp order, the CTAs then exchange their 212 sums through distributed
// shared memory and rank 0 adds them in rank order: no global round trip, no atomics, a fixed summation
// order (run-to-run deterministic).
template <int NP>
__global__ void __launch_bounds__(kFactorBlock)
factor_kernel(const __grid_constant__ FactorDev<NP> P, const float *__restrict__ points, int stride4,
              double *__restrict__ r_out, double *__restrict__ J_out, uint8_t *__restrict__ valid_out,
              double *__restrict__ out_red, unsigned long long *__restrict__ signal, unsigned long long seq) {
    extern __shared__ __align__(16) double smem[];
    double *cols  = smem;                          // [kCols][kColStride]
    double *s_red = smem + kCols * kColStride;     // [kRedSize]
    cg::cluster_group cluster = cg::this_cluster();
    const int pair       = blockIdx.y;
    const FactorDevPair &pr = P.pair[pair];
    const int tid           = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    double acc[6][2];                              // blocks (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
#pragma unroll
    for (int a = 0; a < 6; a++) acc[a][0] = acc[a][1] = 0.0;

    for (int tile = blockIdx.x; tile * kFactorBlock < P.q; tile += gridDim.x) {
        const int k = tile * kFactorBlock + threadIdx.x;
        bool valid  = k < P.q;
        if (valid && pr.type) valid = (pr.type[k] == 0) && (pr.outlier[k] == 0); // optimizer.cc:468
        double J[19], r = 0.0, rho0 = 0.0;
#pragma unroll
        for (int i = 0; i < 19; i++) J[i] = 0.0;
        if (valid) {
            const float *pp = points + (size_t) k * (stride4 ? 4 : 3);
            V3 p = v3((double) pp[0], (double) pp[1], (double) pp[2]);
            V3 n = v3(pr.normal[3 * k], pr.normal[3 * k + 1], pr.normal[3 * k + 2]);
            double s = 1.0 / (pr.std ? pr.std[k] : P.sigma); // sqrt_info (lidar_factor.cc:38)
            r = eval_residual<true>(P.cc, pr.rc, p, n, pr.d[k], s, J);
            if (P.flags & 2u) { J[0] = J[1] = J[2] = J[3] = J[4] = J[5] = 0.0; }
            if (P.flags & 4u) { J[6] = J[7] = J[8] = J[9] = J[10] = J[11] = 0.0; }
            if (P.flags & 8u) { J[12] = J[13] = J[14] = J[15] = J[16] = J[17] = 0.0; }
            if (P.flags & 16u) J[18] = 0.0;
            if (P.flags & 1u) rho0 = huber_correct(P.huber_delta, r, J);
            else rho0 = r * r;
        }
        if (k < P.q) {
            size_t o = (size_t) pair * P.q + k;
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
        if (out_red == nullptr) continue;
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
    if (out_red == nullptr) return;
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
    // thread e (and e + 256 < 384) sums entry e = (block, row, col) over the warps in warp order and files it
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
    cluster.sync();
    constexpr int kPer = (kRedSize + kFactorBlock - 1) / kFactorBlock;
    double total[kPer];
    if (cluster.block_rank() == 0) {
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const int e = tid + j * kFactorBlock;
            double v[8];
#pragma unroll
            for (unsigned rk = 0; rk < 8; rk++)   // independent remote loads in flight together, summed in rank order
                v[rk] = (e < kRedSize && rk < cluster.num_blocks()) ? cluster.map_shared_rank(s_red, rk)[e] : 0.0;
            total[j] = 0.0;
#pragma unroll
            for (unsigned rk = 0; rk < 8; rk++) total[j] += v[rk];
        }
    }
    cluster.sync(); // every CTA's shared memory stays alive until rank 0 has read it; the others are done now
    if (cluster.block_rank() != 0) return;
#pragma unroll
    for (int j = 0; j < kPer; j++) {
        const int e = tid + j * kFactorBlock;
        if (e < kRedSize) out_red[(size_t) pair * kRedSize + e] = total[j];
    }
    if (signal) {
        // out_red is host memory: publish "pair done" behind the data so that the host can poll for it instead of
        // paying a stream synchronisation (the release at system scope orders the CTA's stores before the flag)
        __syncthreads();
        if (tid == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(signal + pair), "l"(seq) : "memory");
    }
}

constexpr size_t kFactorSmem = (size_t) (kCols * kColStride + kRedSize + 4) * sizeof(double);

// K7. counters: [kMaxRefs] per-pair outlier counts + [1] finished-CTA ticket, all zero on entry and zero again
// on exit (the last CTA publishes the counts to pinned host memory, releases the flag and clears them).
template <int NP>
__global__ void __launch_bounds__(256)
chi2_kernel(const __grid_constant__ FactorDev<NP> P, const float *__restrict__ points, int stride4, double threshold,
            int *__restrict__ counters, int *__restrict__ host_counts, unsigned long long *__restrict__ signal,
            unsigned long long seq) {
    const int pair       = blockIdx.y;
    const FactorDevPair &pr = P.pair[pair];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool out    = false;
    if (k < P.q && pr.type[k] == 0 && pr.outlier[k] == 0) {
        const float *pp = points + (size_t) k * (stride4 ? 4 : 3);
        V3 p = v3((double) pp[0], (double) pp[1], (double) pp[2]);
        V3 n = v3(pr.normal[3 * k], pr.normal[3 * k + 1], pr.normal[3 * k + 2]);
        double s = 1.0 / (pr.std ? pr.std[k] : P.sigma);
        double r = eval_residual<false>(P.cc, pr.rc, p, n, pr.d[k], s, nullptr);
        if (r * r > threshold) {       // optimizer.cc:536-541
            pr.outlier[k] = 1;
            out = true;
        }
    }
    int c = __syncthreads_count(out ? 1 : 0);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        if (c) atomicAdd(counters + pair, c);
        __threadfence();
        last = atomicAdd(counters + kMaxRefs, 1) == (int) (gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < P.n_pairs) {
        host_counts[threadIdx.x] = atomicExch(counters + threadIdx.x, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        counters[kMaxRefs] = 0;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(signal), "l"(seq) : "memory");
    }
}

// host FactorParams -> the compact by-value kernel parameter block
template <int NP> void to_dev(const FactorParams &p, FactorDev<NP> &d) {
    d.n_pairs     = p.n_pairs;
    d.q           = p.q;
    d.flags       = p.flags;
    d.sigma       = p.sigma;
    d.huber_delta = p.huber_delta;
    make_cur_const(p, d.cc);
    for (int i = 0; i < p.n_pairs; i++) {
        FactorDevPair &o = d.pair[i];
        o.type    = p.pair[i].type;
        o.outlier = p.pair[i].outlier;
        o.normal  = p.pair[i].normal;
        o.d       = p.pair[i].d;
        o.std     = p.pair[i].std;
        make_ref_const(p, p.pair[i], d.cc, o.rc);
    }
}

template <int NP>
void factor_launch(cudaStream_t s, const FactorParams &p, const float *points, int stride4, double *r, double *J,
                   uint8_t *valid, double *out_red, unsigned long long *signal, unsigned long long seq) {
    static bool init[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!init[dev]) {
        cudaFuncSetAttribute(factor_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kFactorSmem);
        init[dev] = true;
    }
    FactorDev<NP> d;
    to_dev(p, d);
    int tiles = (p.q + kFactorBlock - 1) / kFactorBlock;
    int C     = 1;
    while (C < tiles && C < 8) C <<= 1; // cluster of 1, 2, 4 or 8 CTAs per pair
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim          = dim3(C, p.n_pairs);
    cfg.blockDim         = dim3(kFactorBlock);
    cfg.dynamicSmemBytes = kFactorSmem;
    cfg.stream           = s;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs    = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, factor_kernel<NP>, d, points, stride4, r, J, valid, out_red, signal, seq);
}

template <int NP>
void chi2_launch(cudaStream_t s, const FactorParams &p, const float *points, int stride4, double threshold, int *counters,
                 int *host_counts, unsigned long long *signal, unsigned long long seq) {
    FactorDev<NP> d;
    to_dev(p, d);
    dim3 grid((p.q + 255) / 256, p.n_pairs);
    chi2_kernel<NP><<<grid, 256, 0, s>>>(d, points, stride4, threshold, counters, host_counts, signal, seq);
}

} // namespace

void launch_factor_eval(cudaStream_t s, const FactorParams &p, const float *points, int stride4, double *r, double *J,
                        uint8_t *valid, double *out_red, unsigned long long *signal, unsigned long long seq) {
    if (p.n_pairs <= 0 || p.q <= 0) return;
    if (p.n_pairs <= 4) factor_launch<4>(s, p, points, stride4, r, J, valid, out_red, signal, seq);
    else if (p.n_pairs <= 8) factor_launch<8>(s, p, points, stride4, r, J, valid, out_red, signal, seq);
    else if (p.n_pairs <= 12) factor_launch<12>(s, p, points, stride4, r, J, valid, out_red, signal, seq);
    else factor_launch<16>(s, p, points, stride4, r, J, valid, out_red, signal, seq);
}

void launch_chi2(cudaStream_t s, const FactorParams &p, const float *points, int stride4, double threshold, int *counters,
                 int *host_counts, unsigned long long *signal, unsigned long long seq) {
    if (p.n_pairs <= 0 || p.q <= 0) return;
    if (p.n_pairs <= 4) chi2_launch<4>(s, p, points, stride4, threshold, counters, host_counts, signal, seq);
    else if (p.n_pairs <= 8) chi2_launch<8>(s, p, points, stride4, threshold, counters, host_counts, signal, seq);
    else if (p.n_pairs <= 12) chi2_launch<12>(s, p, points, stride4, threshold, counters, host_counts, signal, seq);
    else chi2_launch<16>(s, p, points, stride4, threshold, counters, host_counts, signal, seq);
}

} // namespace fflins
