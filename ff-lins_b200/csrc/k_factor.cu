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

__constant__ unsigned char c_tri_a[190];
__constant__ unsigned char c_tri_b[190];

constexpr int kRow = 22; // 19 J + r + rho0 + valid

// Grid: (C, n_pairs) with the C CTAs of a pair forming ONE thread-block cluster (C = 1, 2, 4 or 8).
// CTA c handles residual tiles c, c + C, ...; thread e keeps its reduction entry in a register across
// tiles; the CTAs then exchange their 212 partial sums through distributed shared memory and rank 0
// adds them in rank order — no global round trip, no atomics, a fixed summation order.
__global__ void __launch_bounds__(kFactorBlock)
factor_kernel(const __grid_constant__ FactorParams P, const float *__restrict__ points, int stride4,
              double *__restrict__ r_out, double *__restrict__ J_out, uint8_t *__restrict__ valid_out,
              double *__restrict__ out_red) {
    __shared__ double rows[kFactorBlock][kRow + 1]; // +1: odd stride, conflict-free column reads
    __shared__ double s_red[kRedSize];
    cg::cluster_group cluster = cg::this_cluster();
    const int pair       = blockIdx.y;
    const FactorPair &pr = P.pair[pair];
    const PairConst &pc  = pr.pc;
    const int e          = threadIdx.x;
    int ca = 0, cb = -1;
    double sign = 1.0;
    if (e < 190) { ca = c_tri_a[e]; cb = c_tri_b[e]; }
    else if (e < 209) { ca = e - 190; cb = 19; sign = -1.0; }   // b = -J^T r
    else if (e == 209) { ca = 20; cb = -1; }                     // sum rho
    else if (e == 210) { ca = 19; cb = 19; }                     // sum r'^2
    else { ca = 21; cb = -1; }                                   // count
    // four interleaved partial sums shorten the dependent DFMA chain 4x; they are combined in a fixed order
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;

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
        __syncthreads();
        if (e < kRedSize) {
            if (cb >= 0) {
#pragma unroll 4
                for (int i = 0; i < kFactorBlock; i += 4) {
                    acc0 += rows[i][ca] * rows[i][cb];
                    acc1 += rows[i + 1][ca] * rows[i + 1][cb];
                    acc2 += rows[i + 2][ca] * rows[i + 2][cb];
                    acc3 += rows[i + 3][ca] * rows[i + 3][cb];
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < kFactorBlock; i += 4) {
                    acc0 += rows[i][ca];
                    acc1 += rows[i + 1][ca];
                    acc2 += rows[i + 2][ca];
                    acc3 += rows[i + 3][ca];
                }
            }
        }
        __syncthreads();
    }
    if (out_red == nullptr) return;
    if (e < kRedSize) s_red[e] = sign * ((acc0 + acc1) + (acc2 + acc3));
    cluster.sync();
    if (cluster.block_rank() == 0 && e < kRedSize) {
        double total = 0.0;
        for (unsigned rk = 0; rk < cluster.num_blocks(); rk++) total += cluster.map_shared_rank(s_red, rk)[e];
        out_red[(size_t) pair * kRedSize + e] = total;
    }
    cluster.sync(); // keep every CTA's shared memory alive until rank 0 has read it
}

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

bool g_tri_init[64] = {false};

void init_tri() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (g_tri_init[dev]) return;
    unsigned char a[190], b[190];
    int e = 0;
    for (int i = 0; i < 19; i++)
        for (int j = i; j < 19; j++) {
            a[e] = (unsigned char) i;
            b[e] = (unsigned char) j;
            e++;
        }
    cudaMemcpyToSymbol(c_tri_a, a, sizeof(a));
    cudaMemcpyToSymbol(c_tri_b, b, sizeof(b));
    g_tri_init[dev] = true;
}

} // namespace

void launch_factor_eval(cudaStream_t s, FactorParams &p, const float *points, int stride4, double *r, double *J,
                        uint8_t *valid, double *out_red) {
    if (p.n_pairs <= 0 || p.q <= 0) return;
    init_tri();
    for (int i = 0; i < p.n_pairs; i++) make_pair_const(p, p.pair[i], p.pair[i].pc);
    int tiles = (p.q + kFactorBlock - 1) / kFactorBlock;
    int C     = 1;
    while (C < tiles && C < 8) C <<= 1; // cluster of 1, 2, 4 or 8 CTAs per pair
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim          = dim3(C, p.n_pairs);
    cfg.blockDim         = dim3(kFactorBlock);
    cfg.dynamicSmemBytes = 0;
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
