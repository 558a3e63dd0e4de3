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
// per-residual formulation. The reduction is a 256-row "syrk" out of shared memory: thread e owns
// one of the 190 upper-triangle entries (or one of the 19 gradient entries / 3 scalars) and
// accumulates it over the CTA's residuals IN ROW ORDER, CTA partials are then summed in CTA order
// by the last CTA of the pair — a fixed summation order, so results are run-to-run deterministic.
// No tensor cores: a 19-wide rank-1 update is not a dense contraction worth a tcgen05 tile.
#include "math_factor.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace fflins {

namespace {

constexpr int kRowStride = 24; // 19 J + r + rho0 + valid + 2 pad: 192-byte rows, 16-byte aligned groups of 2

// Grid: (C, n_pairs) with the C CTAs of a pair forming ONE thread-block cluster (C = 1, 2, 4 or 8).
// CTA c handles residual tiles c, c + C, ... Reduction = a register-tiled SYRK out of shared memory:
// with v = [J(19) | r], G = sum v v^T is split into 5 x 5 blocks of 4 x 4; thread (slice s, block T) keeps
// the 16 entries of one upper block in registers and walks the 16 rows s, s+16, ... of the tile with four
// 16-byte shared-memory loads per row (0.25 loads per FMA instead of 2). The 16 slice partials are summed
// in slice order, the CTAs then exchange their 212 sums through distributed shared memory and rank 0 adds
// them in rank order: no global round trip, no atomics, a fixed summation order (bit-deterministic).
__global__ void __launch_bounds__(kFactorBlock)
factor_kernel(const __grid_constant__ FactorParams P, const float *__restrict__ points, int stride4,
              double *__restrict__ r_out, double *__restrict__ J_out, uint8_t *__restrict__ valid_out,
              double *__restrict__ out_red) {
    extern __shared__ __align__(16) double smem[];
    double (*rows)[kRowStride] = reinterpret_cast<double (*)[kRowStride]>(smem);          // [256][24]
    double *s_red              = smem + kFactorBlock * kRowStride;                         // [kRedSize]
    cg::cluster_group cluster = cg::this_cluster();
    const int pair       = blockIdx.y;
    const FactorPair &pr = P.pair[pair];
    const PairConst &pc  = pr.pc;
    const int tid        = threadIdx.x;
    const int slice = tid >> 4, T = tid & 15;     // 16 row slices x 16 blocks (15 upper 4x4 blocks of G + 1 scalar block)
    int bi = 0, bj = 0;                           // block coordinates (bi <= bj) of upper block T
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
            r = eval_residual<true>(pc, p, n, pr.d[k], s, J);
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
        for (int i = 0; i < 19; i++) rows[threadIdx.x][i] = J[i];
        rows[threadIdx.x][19] = r;
        rows[threadIdx.x][20] = rho0;
        rows[threadIdx.x][21] = valid ? 1.0 : 0.0;
        rows[threadIdx.x][22] = 0.0;
        rows[threadIdx.x][23] = 0.0;
        __syncthreads();
        if (T < 15) {
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
        } else { // scalar block: sum rho (entry 0) and the residual count (entry 1)
#pragma unroll 4
            for (int kk = 0; kk < 16; kk++) {
                acc[0][0] += rows[slice + 16 * kk][20];
                acc[0][1] += rows[slice + 16 * kk][21];
            }
        }
        __syncthreads();
    }
    if (out_red == nullptr) return;
    // slice partials -> shared memory (the row buffer is free now): part[slice][T][16]
    double *part = smem;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) part[(slice * 16 + T) * 16 + 4 * a + b] = acc[a][b];
    __syncthreads();
    {
        // thread (T, q) sums entry q of block T over the 16 slices in slice order and files it under its
        // place in the 212-vector [190 upper-triangle H | 19 b | sum rho | sum r'^2 | count]
        const int Tq = tid >> 4, q = tid & 15;
        double sum = 0.0;
#pragma unroll
        for (int s2 = 0; s2 < 16; s2++) sum += part[(s2 * 16 + Tq) * 16 + q];
        int ti = 0, tj = 0;
        {
            int t = Tq;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                int len = 5 - i;
                if (t >= 0 && t < len) { ti = i; tj = i + t; t = -1; }
                else if (t >= len) t -= len;
            }
        }
        int e = -1;
        double sign = 1.0;
        if (Tq < 15) {
            const int ra = 4 * ti + (q >> 2), cb = 4 * tj + (q & 3);
            if (ra < 19 && cb < 19) {
                if (ra <= cb) e = ra * 19 - (ra * (ra - 1)) / 2 + (cb - ra);          // upper triangle of H
            } else if (ra < 19 && cb == 19) {
                e    = 190 + ra;                                                     // b = -J^T r
                sign = -1.0;
            } else if (ra == 19 && cb == 19) {
                e = 210;                                                             // sum r'^2
            }
        } else if (q == 0) {
            e = 209;                                                                 // sum rho
        } else if (q == 1) {
            e = 211;                                                                 // count
        }
        __syncthreads(); // everyone has read `part` before s_red (which lies behind the row buffer) is written
        if (e >= 0) s_red[e] = sign * sum;
    }
    cluster.sync();
    if (cluster.block_rank() == 0 && tid < kRedSize) {
        double total = 0.0;
        for (unsigned rk = 0; rk < cluster.num_blocks(); rk++) total += cluster.map_shared_rank(s_red, rk)[tid];
        out_red[(size_t) pair * kRedSize + tid] = total;
    }
    cluster.sync(); // keep every CTA's shared memory alive until rank 0 has read it
}

constexpr size_t kFactorSmem = (size_t) (kFactorBlock * kRowStride + kRedSize + 4) * sizeof(double);

__global__ void __launch_bounds__(256)
chi2_kernel(const __grid_constant__ FactorParams P, const float *__restrict__ points, int stride4, double threshold,
            int *__restrict__ n_outliers) {
    const int pair       = blockIdx.y;
    const FactorPair &pr = P.pair[pair];
    const PairConst &pc  = pr.pc;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool out    = false;
    if (k < P.q && pr.type[k] == 0 && pr.outlier[k] == 0) {
        const float *pp = points + (size_t) k * (stride4 ? 4 : 3);
        V3 p = v3((double) pp[0], (double) pp[1], (double) pp[2]);
        V3 n = v3(pr.normal[3 * k], pr.normal[3 * k + 1], pr.normal[3 * k + 2]);
        double s = 1.0 / (pr.std ? pr.std[k] : P.sigma);
        double r = eval_residual<false>(pc, p, n, pr.d[k], s, nullptr);
        if (r * r > threshold) {       // optimizer.cc:536-541
            pr.outlier[k] = 1;
            out = true;
        }
    }
    int c = __syncthreads_count(out ? 1 : 0);
    if (threadIdx.x == 0 && c) atomicAdd(n_outliers + pair, c);
}

bool g_factor_init[64] = {false};

void init_factor() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_factor_init[dev]) return;
    cudaFuncSetAttribute(factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kFactorSmem);
    g_factor_init[dev] = true;
}

} // namespace

void launch_factor_eval(cudaStream_t s, FactorParams &p, const float *points, int stride4, double *r, double *J,
                        uint8_t *valid, double *out_red) {
    if (p.n_pairs <= 0 || p.q <= 0) return;
    init_factor();
    for (int i = 0; i < p.n_pairs; i++) make_pair_const(p, p.pair[i], p.pair[i].pc);
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
    cudaLaunchKernelEx(&cfg, factor_kernel, p, points, stride4, r, J, valid, out_red);
}

void launch_chi2(cudaStream_t s, FactorParams &p, const float *points, int stride4, double threshold, int *n_outliers) {
    if (p.n_pairs <= 0 || p.q <= 0) return;
    for (int i = 0; i < p.n_pairs; i++) make_pair_const(p, p.pair[i], p.pair[i].pc);
    dim3 grid((p.q + 255) / 256, p.n_pairs);
    chi2_kernel<<<grid, 256, 0, s>>>(p, points, stride4, threshold, n_outliers);
}

} // namespace fflins
