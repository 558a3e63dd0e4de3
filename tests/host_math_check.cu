// host_math_check.cu — runs the product's __host__ __device__ math (ff-lins_b200/csrc/math_*.cuh)
// on the CPU against the oracle. Test infrastructure: built and run by tests/test_host_math.py.
// Prints one line per check: "<name> <max_rel_err>"; exit code 0 always (pytest parses).
#include "math_factor.cuh"
#include "math_plane.cuh"
#include "math_undistort.cuh"
#include "../oracle/oracle.h"

#include <cstdio>
#include <random>
#include <vector>

using namespace fflins;

static double relerr(const double *a, const double *b, int n) {
    double ma = 0, md = 0;
    for (int i = 0; i < n; i++) {
        ma = std::fmax(ma, std::fabs(b[i]));
        md = std::fmax(md, std::fabs(a[i] - b[i]));
    }
    return ma > 0 ? md / ma : md;
}

int main() {
    std::mt19937_64 rng(12345);
    std::normal_distribution<double> N(0, 1);
    std::uniform_real_distribution<double> U(-1, 1);
    // ---------------- factor ----------------
    double worst_r = 0, worst_J[4] = {0, 0, 0, 0}, worst_hub = 0;
    for (int trial = 0; trial < 2000; trial++) {
        FactorParams P{};
        P.n_pairs = 1;
        auto randpose = [&](double *p) {
            for (int i = 0; i < 3; i++) p[i] = 10 * U(rng);
            double q[4], n = 0;
            for (int i = 0; i < 4; i++) { q[i] = N(rng); n += q[i] * q[i]; }
            n = std::sqrt(n) * (1.0 + 1e-3 * U(rng)); // slightly non-unit like a Ceres iterate
            for (int i = 0; i < 4; i++) p[3 + i] = q[i] / n;
        };
        randpose(P.pair[0].pose_ref);
        randpose(P.pose_cur);
        randpose(P.ext);
        for (int i = 0; i < 3; i++) P.ext[i] *= 0.02;
        P.td = 0.02 * U(rng);
        double motion[14];
        for (int i = 0; i < 3; i++) {
            motion[i] = P.pair[0].v0[i] = 3 * N(rng);
            motion[3 + i] = P.pair[0].w0[i] = 0.3 * N(rng);
            motion[7 + i] = P.v1[i] = 3 * N(rng);
            motion[10 + i] = P.w1[i] = 0.3 * N(rng);
        }
        motion[6] = P.pair[0].td0 = 0.02 * U(rng);
        motion[13] = P.td1 = 0.02 * U(rng);
        float pf[3] = {(float) (20 * U(rng)), (float) (20 * U(rng)), (float) (5 * U(rng))};
        double n[3] = {N(rng), N(rng), N(rng)};
        double nn   = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int i = 0; i < 3; i++) n[i] /= nn;
        double d = 5 * U(rng), sigma = 0.1;
        PairConst pc;
        make_pair_const(P, P.pair[0], pc);
        double J[19];
        double r = eval_residual<true>(pc.cur, pc.ref, v3(pf[0], pf[1], pf[2]), v3(n[0], n[1], n[2]), d, 1.0 / sigma, J);
        const double *par[4] = {P.pair[0].pose_ref, P.pose_cur, P.ext, &P.td};
        double ro, J0[7], J1[7], J2[7], J3[1];
        double *jac[4] = {J0, J1, J2, J3};
        orc_lidar_factor_evaluate(pf, n, d, sigma, motion, par, &ro, jac);
        worst_r = std::fmax(worst_r, std::fabs(r - ro) / std::fmax(1.0, std::fabs(ro)));
        worst_J[0] = std::fmax(worst_J[0], relerr(J, J0, 6));
        worst_J[1] = std::fmax(worst_J[1], relerr(J + 6, J1, 6));
        worst_J[2] = std::fmax(worst_J[2], relerr(J + 12, J2, 6));
        worst_J[3] = std::fmax(worst_J[3], std::fabs(J[18] - J3[0]) / std::fmax(1.0, std::fabs(J3[0])));
        // huber
        double J22[22];
        for (int b = 0; b < 3; b++) { for (int i = 0; i < 7; i++) J22[7 * b + i] = jac[b][i]; }
        J22[21] = J3[0];
        double rh = ro;
        orc_huber_correct(1.0, &rh, J22);
        double rr = r;
        huber_correct(1.0, rr, J);
        double Jh[19];
        for (int b = 0; b < 3; b++) for (int i = 0; i < 6; i++) Jh[6 * b + i] = J22[7 * b + i];
        Jh[18] = J22[21];
        worst_hub = std::fmax(worst_hub, relerr(J, Jh, 19));
        worst_hub = std::fmax(worst_hub, std::fabs(rr - rh) / std::fmax(1.0, std::fabs(rh)));
    }
    printf("factor_r %.3e\nfactor_J_ref %.3e\nfactor_J_cur %.3e\nfactor_J_ext %.3e\nfactor_J_td %.3e\nhuber %.3e\n", worst_r,
           worst_J[0], worst_J[1], worst_J[2], worst_J[3], worst_hub);
    // ---------------- plane ----------------
    int plane_bit_mismatch = 0, plane_ok_mismatch = 0;
    double plane_worst = 0;
    for (int trial = 0; trial < 5000; trial++) {
        float pts[20];
        float4 p4[5];
        double nrm[3] = {N(rng), N(rng), N(rng)};
        double off = 3 * U(rng) + 4;
        int mode = trial % 4; // 0 planar, 1 noisy planar, 2 random, 3 nearly collinear
        for (int k = 0; k < 5; k++) {
            double a = 5 * U(rng), b = 5 * U(rng);
            double x, y, z;
            if (mode == 2) { x = 5 * U(rng); y = 5 * U(rng); z = 5 * U(rng); }
            else if (mode == 3) { x = a; y = 2 * a + 1e-4 * U(rng); z = -a + 1 + 1e-4 * U(rng); }
            else {
                x = a; y = b;
                z = (-off - nrm[0] * x - nrm[1] * y) / (std::fabs(nrm[2]) + 0.3);
                if (mode == 1) z += 0.03 * N(rng);
            }
            pts[4 * k] = (float) x; pts[4 * k + 1] = (float) y; pts[4 * k + 2] = (float) z; pts[4 * k + 3] = 0;
            p4[k] = make_float4(pts[4 * k], pts[4 * k + 1], pts[4 * k + 2], 0);
        }
        double u0[3], d0, ds0[5], u1[3], d1, ds1[5];
        bool ok0 = orc_plane_estimation(pts, 0.1, u0, &d0, ds0);
        bool ok1 = plane_estimation(p4, 0.1, u1, d1, ds1);
        if (ok0 != ok1) plane_ok_mismatch++;
        if (u0[0] != u1[0] || u0[1] != u1[1] || u0[2] != u1[2] || d0 != d1) plane_bit_mismatch++;
        if (mode != 3) {
            plane_worst = std::fmax(plane_worst, relerr(u1, u0, 3));
            plane_worst = std::fmax(plane_worst, std::fabs(d1 - d0) / std::fmax(1e-12, std::fabs(d0)));
        }
    }
    printf("plane_rel %.3e\nplane_bit_mismatch %d\nplane_ok_mismatch %d\n", plane_worst, plane_bit_mismatch,
           plane_ok_mismatch);
    // ---------------- undistort pose ----------------
    {
        const int W = 400;
        std::vector<double> t(W), p(3 * W), q(4 * W), tab(6 * W);
        for (int i = 0; i < W; i++) {
            t[i] = 345600.0 + i * 0.005;
            double a = 0.6 * i * 0.005;
            p[3 * i] = 10 * std::cos(a) + 1e-4 * N(rng);
            p[3 * i + 1] = 10 * std::sin(a) + 1e-4 * N(rng);
            p[3 * i + 2] = 1.5;
            double qq[4] = {0.01 * std::sin(3 * a), 0.02 * std::cos(2 * a), std::sin(a / 2), std::cos(a / 2)};
            double nn = 0;
            for (int k = 0; k < 4; k++) nn += qq[k] * qq[k];
            for (int k = 0; k < 4; k++) q[4 * i + k] = qq[k] / std::sqrt(nn);
        }
        if (true) { // a stationary stretch: identical consecutive quaternions (zero-rotation branch)
            for (int i = 100; i < 110; i++) for (int k = 0; k < 4; k++) q[4 * i + k] = q[4 * 100 + k];
        }
        Pose12 ext{};
        double Rbl[9] = {0.999, -0.04, 0.01, 0.04, 0.999, 0.02, -0.0108, -0.0196, 0.9997};
        for (int i = 0; i < 9; i++) ext.R[i] = Rbl[i];
        ext.t[0] = 0.05; ext.t[1] = -0.02; ext.t[2] = 0.1;
        // frame pose at a stamp near the end of the window: host evaluation must equal the oracle's
        double stamp = t[W - 7] + 0.0013;
        Pose12 frame;
        pose_from_window(t.data(), p.data(), q.data(), W, ext, stamp, frame);
        double Rf[9], tf[3];
        orc_pose_from_ins_window(t.data(), p.data(), q.data(), W, ext.R, ext.t, stamp, Rf, tf);
        int frame_pose_bits = 0;
        for (int i = 0; i < 9; i++) frame_pose_bits += (frame.R[i] != Rf[i]);
        for (int i = 0; i < 3; i++) frame_pose_bits += (frame.t[i] != tf[i]);
        std::vector<double> rows((size_t) W * kRowSize);
        for (int i = 0; i < W; i++) interval_row(t.data(), p.data(), q.data(), W, i, ext, frame, &rows[(size_t) i * kRowSize]);
        double worst = 0;
        int idx_mismatch = 0;
        double inv_dt = (W - 1) / (t[W - 1] - t[0]);
        for (int trial = 0; trial < 20000; trial++) {
            double time = t[0] + (t[W - 1] - t[0]) * (0.5 + 0.52 * U(rng));
            if (trial % 50 == 0) time = t[(trial / 50) % W]; // exactly on a sample
            int i0 = orc_ins_window_index(t.data(), W, time);
            int i1 = window_index(t.data(), W, time, t[0], t[W - 1], inv_dt);
            if (i0 != i1) idx_mismatch++;
            double R1[9], t1[3], R01o[9], t01o[3];
            orc_pose_from_ins_window(t.data(), p.data(), q.data(), W, ext.R, ext.t, time, R1, t1);
            orc_relative_pose(Rf, tf, R1, t1, R01o, t01o);
            // a point ~60 m away: the error is reported relative to the point's magnitude
            double px = 40.0 * U(rng), py = 40.0 * U(rng), pz = 8.0 * U(rng), out[3];
            row_apply(&rows[(size_t) i1 * kRowSize], time, px, py, pz, out);
            for (int k = 0; k < 3; k++) {
                double ref = R01o[3 * k] * px + R01o[3 * k + 1] * py + R01o[3 * k + 2] * pz + t01o[k];
                worst = std::fmax(worst, std::fabs(out[k] - ref) / 60.0);
            }
        }
        printf("frame_pose_bit_mismatch %d\n", frame_pose_bits);
        printf("undistort_pose_rel %.3e\nwindow_index_mismatch %d\n", worst, idx_mismatch);
    }
    return 0;
}
