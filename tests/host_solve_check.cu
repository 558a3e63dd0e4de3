// host_solve_check.cu — the numerics of the device-side window solve (ff-lins_b200/csrc/math_solve.cuh: assembly,
// elimination, Levenberg-Marquardt control) executed on ONE host thread (ParSerial), with the per-pair blocks computed
// by the product's own factor algebra (math_factor.cuh) in plain loops. Built as a shared library and driven by
// tests/test_host_solve.py, which compares the result with an independent numpy LM (tests/lm_ref.py) over the
// oracle's blocks. Test infrastructure: nothing in the product links it.
#include "math_factor.cuh"
#include "math_solve.cuh"

#include <cstring>
#include <vector>

using namespace fflins;

namespace {

struct HostProblem {
    int n, q;
    const float *points;        // q x 4
    const int8_t *type;         // n x q
    uint8_t *outlier;           // n x q
    const double *normal;       // n x q x 3
    const double *d;            // n x q
    const double *motions;      // (n + 1) x 7: v, w, td of the refs, then of the current frame
    double sigma, huber;
    unsigned flags;
};

void pair_consts(const HostProblem &P, const double *x, int p, CurConst &cc, RefConst &rc) {
    const int n = P.n;
    const double *mc = P.motions + 7 * n, *mr = P.motions + 7 * p;
    const double td = x[7 * n + 14];
    make_cur_const_raw(x + 7 * n, x + 7 * n + 7, td, mc, mc + 3, mc[6], cc);
    make_ref_const_a(x + 7 * p, mr, mr + 3, mr[6], td, mc, rc);
    make_ref_const_b(cc, rc);
}

void eval_blocks(const HostProblem &P, const double *x, double *red) {
    for (int p = 0; p < P.n; p++) {
        CurConst cc;
        RefConst rc;
        pair_consts(P, x, p, cc, rc);
        double *R = red + (size_t) p * kRedStride;
        std::memset(R, 0, sizeof(double) * kRedStride);
        for (int k = 0; k < P.q; k++) {
            const size_t o = (size_t) p * P.q + k;
            if (P.type[o] != 0 || P.outlier[o] != 0) continue;
            double J[19];
            V3 pt = v3(P.points[4 * k], P.points[4 * k + 1], P.points[4 * k + 2]);
            V3 nn = v3(P.normal[3 * o], P.normal[3 * o + 1], P.normal[3 * o + 2]);
            double r = eval_residual<true>(cc, rc, pt, nn, P.d[o], 1.0 / P.sigma, J);
            if (P.flags & 2u) for (int i = 0; i < 6; i++) J[i] = 0;
            if (P.flags & 4u) for (int i = 6; i < 12; i++) J[i] = 0;
            if (P.flags & 8u) for (int i = 12; i < 18; i++) J[i] = 0;
            if (P.flags & 16u) J[18] = 0;
            double rho0 = (P.flags & 1u) ? huber_correct(P.huber, r, J) : r * r;
            int e = 0;
            for (int a = 0; a < 19; a++)
                for (int c = a; c < 19; c++) R[e++] += J[a] * J[c];
            for (int a = 0; a < 19; a++) R[190 + a] -= J[a] * r;
            R[209] += rho0;
            R[210] += r * r;
            R[211] += 1.0;
        }
    }
}

void chi2_pass(const HostProblem &P, const double *x, double thr, int *counts) {
    for (int p = 0; p < P.n; p++) {
        CurConst cc;
        RefConst rc;
        pair_consts(P, x, p, cc, rc);
        counts[p] = 0;
        for (int k = 0; k < P.q; k++) {
            const size_t o = (size_t) p * P.q + k;
            if (P.type[o] != 0 || P.outlier[o] != 0) continue;
            V3 pt = v3(P.points[4 * k], P.points[4 * k + 1], P.points[4 * k + 2]);
            V3 nn = v3(P.normal[3 * o], P.normal[3 * o + 1], P.normal[3 * o + 2]);
            double r = eval_residual<false>(cc, rc, pt, nn, P.d[o], 1.0 / P.sigma, nullptr);
            if (r * r > thr) {
                P.outlier[o] = 1;
                counts[p]++;
            }
        }
    }
}

struct HostEval {
    const HostProblem &P;
    SolveWork *W;
    SolvePrior prior;
    std::vector<double> red;
    double operator()(const double *x, double *H, double *g) {
        eval_blocks(P, x, red.data());
        return solve_assemble(ParSerial{}, *W, prior, x, red.data(), H, g);
    }
    bool failed() const { return false; }
};

} // namespace

extern "C" {

// opt: max_iters0, max_iters1, chi2_threshold, ftol, gtol, ptol, radius, flags, fixed_refs (as doubles)
// report: per stage iterations, successful, evaluations, cost_initial, cost_final (2 x 5), then n outlier counts
int solve_host_run(int n, int q, const float *points, const int8_t *type, uint8_t *outlier, const double *normal, const double *d,
                   const double *motions, double sigma, double huber, double *x, const double *opt, const double *H0,
                   const double *b0, const double *x0, double c0, double *report) {
    HostProblem P{n, q, points, type, outlier, normal, d, motions, sigma, huber, (unsigned) opt[7]};
    SolveOptions O{};
    O.max_iters[0]          = (int) opt[0];
    O.max_iters[1]          = (int) opt[1];
    O.chi2_threshold        = opt[2];
    O.function_tolerance    = opt[3];
    O.gradient_tolerance    = opt[4];
    O.parameter_tolerance   = opt[5];
    O.initial_radius        = opt[6];
    O.min_relative_decrease = 1e-3;
    O.flags                 = (unsigned) opt[7];
    O.fixed_refs            = (unsigned) opt[8];
    const int D = solve_dim(n), ld = solve_ld(D), NX = solve_nx(n);
    std::vector<double> H((size_t) D * ld), A((size_t) D * ld), g(ld), gn(ld), dx(ld), rhs(ld), dinv(ld), delta(ld), xs(x, x + NX), xc(NX);
    std::vector<unsigned short> tri((size_t) (D - 1) * D / 2 + 1);
    std::vector<unsigned char> fixed(D), free_idx(D);
    solve_fill_tri(0, 1, D, tri.data());
    solve_fill_fixed(0, 1, n, O.flags, O.fixed_refs, fixed.data());
    const int Df = solve_fill_free(D, fixed.data(), free_idx.data());
    SolveWork W{n, D, ld, H.data(), A.data(), g.data(), gn.data(), dx.data(), rhs.data(), dinv.data(), delta.data(), xs.data(), xc.data(),
                fixed.data(), free_idx.data(), Df, tri.data()};
    HostEval eval{P, &W, SolvePrior{H0, b0, x0, c0}, std::vector<double>((size_t) n * kRedStride)};
    SolveStage st[2];
    st[0] = solve_lm(ParSerial{}, W, O, O.max_iters[0], eval);
    std::vector<int> counts(n, 0);
    if (O.chi2_threshold > 0) chi2_pass(P, W.x, O.chi2_threshold, counts.data());
    st[1] = solve_lm(ParSerial{}, W, O, O.max_iters[1], eval);
    for (int s = 0; s < 2; s++) {
        report[5 * s + 0] = st[s].iterations;
        report[5 * s + 1] = st[s].successful;
        report[5 * s + 2] = st[s].evaluations;
        report[5 * s + 3] = st[s].cost_initial;
        report[5 * s + 4] = st[s].cost_final;
    }
    for (int p = 0; p < n; p++) report[10 + p] = counts[p];
    std::memcpy(x, W.x, sizeof(double) * NX);
    return 0;
}

// A x = b through solve_spd alone (n = number of pairs only sets the dimension D = 6 n + 13)
int solve_host_spd(int n, const double *A_in, const double *b, double *x) {
    const int D = solve_dim(n), ld = solve_ld(D);
    std::vector<double> A((size_t) D * ld), rhs(b, b + D), dinv(ld);
    for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++) A[(size_t) i * ld + j] = A_in[(size_t) i * D + j];
    std::vector<unsigned short> tri((size_t) (D - 1) * D / 2 + 1);
    solve_fill_tri(0, 1, D, tri.data());
    SolveWork W{};
    W.n = n; W.D = D; W.ld = ld; W.dinv = dinv.data(); W.tri = tri.data();
    rhs.resize(ld);
    return solve_spd(ParSerial{}, W, D, A.data(), rhs.data(), x) ? 0 : 1;
}

} // extern "C"
