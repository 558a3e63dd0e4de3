// math_solve.cuh — numerics of the device-side window solve (k_factor.cu: solve_kernel; entry point fflins_window_solve).
//
// Reference behaviour restated (not translated):
//   Optimizer::optimization       ff_lins/ff_lins/optimizer.cc:85-185   (LM, 5 iterations, chi2 cull, 10 iterations)
//   PoseManifold::Plus            ff_lins/factors/pose_manifold.cc:35-50
//   MarginalizationFactor::Evaluate (the form of the prior: e0 + J0 dx, dx = x (-) x0)   factors/marginalization_factor.cc:37-91
// The Levenberg-Marquardt loop is the one of ff-lins_b200/host/host.cc (lm_solve), which restates Ceres' documented
// trust-region algorithm: damping diag(J^T J) / radius, step accepted when the relative decrease exceeds 1e-3,
// radius *= 1 / max(1/3, 1 - (2 rho - 1)^3) on success, radius /= 2, 4, 8 ... on failure.
//
// The window state is [ref pose 0 .. n-1 | cur pose | extrinsic | td] (7 doubles per pose: t, q = x y z w), its tangent
// [6 | ... | 6 | 6 | 1], D = 6 n + 13. Every function is a sequence of parallel loops separated by barriers, written
// against a `Par` policy: ParSerial runs them on one host thread (tests/host_solve_check.cu compares that against an
// independent numpy LM before any GPU time is spent), the kernel's policy maps them onto the 128 threads of the solver CTA.
#pragma once
#include "kernels.h"

#include <cfloat>
#include <cmath>

namespace fflins {

constexpr int kSolveMaxDim  = 6 * kMaxPairs + 13;                         // 103
constexpr int kSolveMaxX    = 7 * kMaxPairs + 15;                         // 120
constexpr int kSolveMaxTri  = (kSolveMaxDim - 1) * kSolveMaxDim / 2;      // entries of the elimination table
constexpr int kRedStride    = 216;                                        // 212 sums of a pair, padded to whole 3-value units

__host__ __device__ inline int solve_dim(int n) { return 6 * n + 13; }
__host__ __device__ inline int solve_nx(int n) { return 7 * n + 15; }
__host__ __device__ inline int solve_ld(int D) { return D | 1; }          // odd: column walks touch distinct banks

// read-only global data (the prior): the non-coherent path tells the compiler that the shared-memory stores around it do not alias
__host__ __device__ inline double solve_ldg(const double *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct SolvePrior {              // 0.5 d^T H0 d - b0^T d + c0 with d = x (-) x0; H0 == nullptr: none
    const double *H0, *b0, *x0;
    double c0;
};

struct SolveWork {
    int n, D, ld;
    double *H, *A;               // D x ld: the linearisation at x (full symmetric) | scratch: damped copy, then the candidate's
    double *g, *gn;              // b = -J^T r at x | at the candidate
    double *dx, *rhs, *dinv, *delta;
    double *x, *xc;              // state | candidate
    unsigned char *fixed;        // D
    unsigned char *free_idx;     // the Df free entries of the tangent vector, ascending (constant blocks never enter the factorisation)
    int Df;
    const unsigned short *tri;   // (r << 8 | c), r >= c, row by row: the first m (m + 1) / 2 entries cover an m x m triangle
};

struct ParSerial {
    __host__ __device__ int tid() const { return 0; }
    __host__ __device__ int nthr() const { return 1; }
    __host__ __device__ void sync() const {}
    __host__ __device__ double sum(double v) const { return v; }
    __host__ __device__ double max(double v) const { return v; }
    __host__ __device__ int lanes() const { return 1; }       // threads of a warp
    __host__ __device__ void syncwarp() const {}
    __host__ __device__ int warp() const { return 0; }
    __host__ __device__ int nwarps() const { return 1; }
    __host__ __device__ int lane() const { return 0; }
    __host__ __device__ double warp_sum(double v) const { return v; }   // over the lanes of the caller's warp, in every lane
    __host__ __device__ void stamp(int) const {}                        // diagnostics: end of phase k (see SolveReport::phase_cycles)
    // U x = D^-1 y for the unit-diagonal upper factor left by solve_spd (rows of A above the diagonal, dinv): x_k =
    // y_k dinv_k, then y_r -= A[r][k] x_k for r < k, k = D-1 .. 0. Called by every thread; ends with a barrier.
    __host__ __device__ void backsub(int D, int ld, const double *A, const double *dinv, double *rhs, double *x_out) const {
        for (int k = D - 1; k >= 0; k--) {
            const double xk = rhs[k] * dinv[k];
            x_out[k]        = xk;
            for (int r = 0; r < k; r++) rhs[r] -= A[r * ld + k] * xk;
        }
    }
};

// (row, col) of a flat index that advances by a fixed stride, without a division per element
struct RowCol {
    int i, j, di, dj, D;
    __host__ __device__ RowCol(int start, int stride, int D_) : i(start / D_), j(start % D_), di(stride / D_), dj(stride % D_), D(D_) {}
    __host__ __device__ void next() {
        i += di;
        j += dj;
        if (j >= D) {
            j -= D;
            i++;
        }
    }
};

// Rotation::rotvec2quaternion (rotation.cc:61-65)
__host__ __device__ inline Q4 rotvec_to_q(V3 rv) {
    const double a = sqrt(rv.x * rv.x + rv.y * rv.y + rv.z * rv.z);
    if (a == 0.0) return Q4{0, 0, 0, 1};
    const double s = sin(a / 2) / a;
    return Q4{rv.x * s, rv.y * s, rv.z * s, cos(a / 2)};
}
// PoseManifold::Plus (pose_manifold.cc:35-50)
__host__ __device__ inline void pose_plus7(const double *x, const double *d, double *out) {
    for (int i = 0; i < 3; i++) out[i] = x[i] + d[i];
    const Q4 q = qmul(Q4{x[3], x[4], x[5], x[6]}, rotvec_to_q(v3(d[3], d[4], d[5])));
    const double nq = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    out[3] = q.x / nq;
    out[4] = q.y / nq;
    out[5] = q.z / nq;
    out[6] = q.w / nq;
}
// MarginalizationFactor's dx of a pose block (marginalization_factor.cc:54-64)
__host__ __device__ inline void pose_minus6(const double *x, const double *x0, double *d) {
    for (int i = 0; i < 3; i++) d[i] = x[i] - x0[i];
    const Q4 dq    = qmul(qinv(Q4{x0[3], x0[4], x0[5], x0[6]}), Q4{x[3], x[4], x[5], x[6]});
    const double s = dq.w < 0 ? -2.0 : 2.0;
    d[3] = s * dq.x;
    d[4] = s * dq.y;
    d[5] = s * dq.z;
}

// column of tangent entry a (0..18 = ref 6 | cur 6 | ext 6 | td) of pair p in the window's tangent vector
__host__ __device__ inline int solve_col(int n, int p, int a) { return a < 6 ? 6 * p + a : 6 * n + (a - 6); }
__host__ __device__ inline int tri19(int a, int c) { return a * 19 - (a * (a - 1)) / 2 + (c - a); }   // a <= c

__host__ __device__ inline void solve_fill_tri(int tid, int nthr, int D, unsigned short *tri) {
    for (int r = tid; r < D - 1; r += nthr)
        for (int c = 0; c <= r; c++) tri[r * (r + 1) / 2 + c] = (unsigned short) ((r << 8) | c);
}
__host__ __device__ inline void solve_fill_fixed(int tid, int nthr, int n, unsigned flags, unsigned fixed_refs, unsigned char *fixed) {
    const int D = solve_dim(n);
    for (int i = tid; i < D; i += nthr) {
        bool f;
        if (i < 6 * n) f = (flags & 2u) || ((fixed_refs >> (i / 6)) & 1u);
        else if (i < 6 * n + 6) f = (flags & 4u) != 0;
        else if (i < 6 * n + 12) f = (flags & 8u) != 0;
        else f = (flags & 16u) != 0;
        fixed[i] = f ? 1 : 0;
    }
}

// one thread, after solve_fill_fixed
__host__ __device__ inline int solve_fill_free(int D, const unsigned char *fixed, unsigned char *free_idx) {
    int nf = 0;
    for (int i = 0; i < D; i++)
        if (!fixed[i]) free_idx[nf++] = (unsigned char) i;
    return nf;
}

// The normal equations at x: prior + the pairs' reduced 19 x 19 blocks (red: n x kRedStride, the layout of kRedSize),
// in the order host.cc's linearize adds them (prior first, then the pairs in window order). Returns the cost.
#pragma nv_exec_check_disable
template <class Par>
__host__ __device__ double solve_assemble(const Par &par, const SolveWork &W, const SolvePrior &P, const double *x, const double *red,
                                          double *H, double *g) {
    const int n = W.n, D = W.D, ld = W.ld, tid = par.tid(), nt = par.nthr();
    const bool prior = P.H0 != nullptr;
    if (prior) {
        for (int k = tid; k < n + 3; k += nt) {
            if (k < n + 2) pose_minus6(x + 7 * k, P.x0 + 7 * k, W.delta + 6 * k);
            else W.delta[6 * n + 12] = x[7 * n + 14] - P.x0[7 * n + 14];
        }
    }
    par.sync();
    {
        RowCol rc(tid, nt, D);
        for (int e = tid; e < D * D; e += nt, rc.next()) H[rc.i * ld + rc.j] = prior ? solve_ldg(P.H0 + e) : 0.0;
    }
    double cpart = 0.0;
    for (int i = par.warp(); i < D; i += par.nwarps()) {   // one warp per row of H0, the lanes across it
        double gi = 0.0;
        if (prior) {
            double s = 0.0;
            for (int j = par.lane(); j < D; j += par.lanes()) s += solve_ldg(P.H0 + i * D + j) * W.delta[j];
            gi = P.b0[i] - par.warp_sum(s);
            cpart -= 0.5 * W.delta[i] * (P.b0[i] + gi);                      // 0.5 d^T H0 d - b0^T d
        }
        if (par.lane() == 0) g[i] = gi;
        else cpart = 0.0;
    }
    par.sync();
    double cost = (prior ? P.c0 : 0.0) + par.sum(cpart);
    for (int e = tid; e < 361 + 19; e += nt) {
        if (e < 361) {
            const int a = e / 19, c = e - 19 * a, t = a <= c ? tri19(a, c) : tri19(c, a);
            for (int p = 0; p < n; p++) H[solve_col(n, p, a) * ld + solve_col(n, p, c)] += red[p * kRedStride + t];
        } else {
            const int a = e - 361;
            for (int p = 0; p < n; p++) g[solve_col(n, p, a)] += red[p * kRedStride + 190 + a];
        }
    }
    double lpart = 0.0;
    for (int p = tid; p < n; p += nt) lpart += 0.5 * red[p * kRedStride + 209];
    par.sync();
    cost += par.sum(lpart);
    return cost;
}

// A x = rhs for a symmetric positive definite A (upper triangle used, overwritten): right-looking elimination
// A = L D L^T with the right-hand side carried along, then back substitution (Par::backsub: D dependent steps; on the
// device one warp with the right-hand side in registers and a shuffle per step). One barrier per column: the pivot row
// is final when its step starts and is not written during it. Returns false when a pivot is not positive.
// Measured on B200 (tools/solve_bench.py, D = 73, 512 threads): ~1,000 cycles per column, i.e. the dependent chain
// barrier -> pivot -> division -> one or two passes of the update; three attempts to shorten it lost (next pivot and its
// reciprocal carried in registers so that the division overlaps the update: +17 %; four entries per thread and pass with
// the operands fetched first: +27 %; single-precision reciprocal + two Newton steps instead of the division: +7 %).
#pragma nv_exec_check_disable
template <class Par>
__host__ __device__ bool solve_spd(const Par &par, const SolveWork &W, int D, double *A, double *rhs, double *x_out) {
    const int ld = W.ld, tid = par.tid(), nt = par.nthr();
    for (int k = 0; k < D; k++) {
        const double pivot = A[k * ld + k];
        if (!(pivot > 0.0) || !(pivot < DBL_MAX)) return false;   // the same value in every thread
        const double inv = 1.0 / pivot, gk = rhs[k] * inv;
        if (tid == 0) W.dinv[k] = inv;
        const int m = D - 1 - k, mt = m * (m + 1) / 2;
        const double *rowk = A + k * ld;
        for (int e = tid; e < mt + m; e += nt) {
            if (e < mt) {
                const unsigned rc = W.tri[e];
                const int row = D - 1 - (int) (rc >> 8), col = D - 1 - (int) (rc & 255u);   // k < row <= col
                A[row * ld + col] -= rowk[row] * (rowk[col] * inv);
            } else {
                const int i = k + 1 + (e - mt);
                rhs[i] -= rowk[i] * gk;
            }
        }
        par.sync();
    }
    par.stamp(6);
    par.backsub(D, ld, A, W.dinv, rhs, x_out);
    return true;
}

struct SolveStage {
    int iterations, successful, evaluations, failed;
    double cost_initial, cost_final;
};

// One ceres::Solve with LEVENBERG_MARQUARDT (host.cc lm_solve, restricted to what lives on the device: the LiDAR factors
// and a quadratic prior). eval(x, H, g) linearises at x and returns the cost; eval.failed() reports a broken hand-off.
#pragma nv_exec_check_disable
template <class Par, class Eval>
__host__ __device__ SolveStage solve_lm(const Par &par, SolveWork &W, const SolveOptions &O, int max_iters, Eval &eval) {
    const int n = W.n, D = W.D, ld = W.ld, tid = par.tid(), nt = par.nthr();
    SolveStage S{0, 0, 1, 0, 0.0, 0.0};
    double cost = eval(W.x, W.H, W.g);
    S.cost_initial = S.cost_final = cost;
    if (eval.failed()) {
        S.failed = 1;
        return S;
    }
    double radius = O.initial_radius, decrease = 2.0;
    for (int it = 0; it < max_iters; it++) {
        S.iterations++;
        double gm = 0.0;
        for (int i = tid; i < D; i += nt)
            if (!W.fixed[i]) gm = fmax(gm, fabs(W.g[i]));
        gm = par.max(gm);
        if (gm < O.gradient_tolerance) break;
        // the free rows and columns, compacted (constant blocks carry no information); damping diag(H) / radius
        const int Df = W.Df;
        {
            RowCol rc(tid, nt, Df > 0 ? Df : 1);
            for (int e = tid; e < Df * Df; e += nt, rc.next()) {
                const int i = W.free_idx[rc.i], j = W.free_idx[rc.j];
                double v    = W.H[i * ld + j];
                if (i == j) v += fmin(fmax(v, 1e-6), 1e32) / radius;
                W.A[rc.i * ld + rc.j] = v;
            }
        }
        for (int a = tid; a < Df; a += nt) W.rhs[a] = W.g[W.free_idx[a]];
        for (int i = tid; i < D; i += nt) W.dx[i] = 0.0;
        par.sync();
        par.stamp(2);
        const bool ok = solve_spd(par, W, Df, W.A, W.rhs, W.delta);
        if (ok)
            for (int a = tid; a < Df; a += nt) W.dx[W.free_idx[a]] = W.delta[a];
        par.sync();
        par.stamp(3);
        double mpart = 0.0, dpart = 0.0;
        if (ok) {
            for (int i = par.warp(); i < D; i += par.nwarps()) {   // one warp per row of H
                if (W.fixed[i]) continue;
                double hd = 0.0;
                for (int j = par.lane(); j < D; j += par.lanes())
                    if (!W.fixed[j]) hd += W.H[i * ld + j] * W.dx[j];
                hd = par.warp_sum(hd);
                if (par.lane() == 0) {
                    mpart += W.dx[i] * (W.g[i] - 0.5 * hd);   // model decrease dx^T (b - 0.5 H dx)
                    dpart += W.dx[i] * W.dx[i];
                }
            }
        }
        const double model = par.sum(mpart), dn = par.sum(dpart);
        par.stamp(4);
        if (!ok || !(model > 0.0)) {
            radius /= decrease;
            decrease *= 2.0;
            continue;
        }
        for (int k = tid; k < n + 3; k += nt) {
            if (k < n + 2) {
                if (W.fixed[6 * k]) {
                    for (int i = 0; i < 7; i++) W.xc[7 * k + i] = W.x[7 * k + i];
                } else {
                    pose_plus7(W.x + 7 * k, W.dx + 6 * k, W.xc + 7 * k);
                }
            } else {
                W.xc[7 * n + 14] = W.x[7 * n + 14] + W.dx[6 * n + 12];
            }
        }
        par.sync();
        par.stamp(5);
        const double cost_n = eval(W.xc, W.A, W.gn);
        S.evaluations++;
        if (eval.failed()) {
            S.failed = 1;
            break;
        }
        const double rho = (cost - cost_n) / model;
        if (rho > O.min_relative_decrease) {
            S.successful++;
            double xpart = 0.0;
            for (int i = tid; i < 7 * (n + 1); i += nt) xpart += W.xc[i] * W.xc[i];
            const double xn = par.sum(xpart), dcost = cost - cost_n;
            double *t;
            t = W.x;  W.x = W.xc;  W.xc = t;
            t = W.H;  W.H = W.A;   W.A = t;
            t = W.g;  W.g = W.gn;  W.gn = t;
            cost = cost_n;
            const double u = 2.0 * rho - 1.0;
            radius   = radius / fmax(1.0 / 3.0, 1.0 - u * u * u);
            decrease = 2.0;
            if (fabs(dcost) <= O.function_tolerance * cost) break;
            if (sqrt(dn) <= O.parameter_tolerance * (sqrt(xn) + O.parameter_tolerance)) break;
        } else {
            radius /= decrease;
            decrease *= 2.0;
        }
    }
    S.cost_final = cost;
    return S;
}

} // namespace fflins
