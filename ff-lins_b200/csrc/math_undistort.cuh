// math_undistort.cuh — INS-window pose interpolation shared by K1 (k_frame.cu).
// __host__ __device__ so that tests can run the same code on the CPU (tests/host_math_check.cu).
#pragma once
#include "kernels.h"
#include <cmath>

namespace fflins {

__host__ __device__ __forceinline__ V3 quat_log(Q4 q) {
    // Rotation::quaternion2vector via Eigen AngleAxisd(q) (common/rotation.cc:67-70)
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    if (n < 2.220446049250313e-16) {
        double m = fmax(fabs(q.x), fmax(fabs(q.y), fabs(q.z)));
        if (m == 0.0) {
            n = 0.0;
        } else {
            double ax = q.x / m, ay = q.y / m, az = q.z / m;
            n = m * sqrt(ax * ax + ay * ay + az * az);
        }
    }
    if (n != 0.0) {
        double angle = 2.0 * atan2(n, fabs(q.w));
        if (q.w < 0.0) n = -n;
        return V3{angle * (q.x / n), angle * (q.y / n), angle * (q.z / n)};
    }
    return V3{0.0, 0.0, 0.0};
}

__host__ __device__ __forceinline__ Q4 quat_exp(V3 rv) {
    // Rotation::rotvec2quaternion (common/rotation.cc:61-65)
    double z     = rv.x * rv.x + rv.y * rv.y + rv.z * rv.z;
    double angle = sqrt(z);
    V3 vec       = rv;
    if (z > 0.0) vec = V3{rv.x / angle, rv.y / angle, rv.z / angle};
    double sn, cs;
    sincos(0.5 * angle, &sn, &cs);
    return Q4{sn * vec.x, sn * vec.y, sn * vec.z, cs};
}

// MISC::getInsWindowIndex (misc.cc:27-62). The C-ABI guarantees strictly increasing times and
// w <= 65536, under which the reference's capped bisection returns the unique bracket; a
// uniform-rate guess finds it in one probe for evenly sampled IMU data.
__host__ __device__ __forceinline__ int window_index(const double *__restrict__ t, int w, double time, double t_front,
                                            double t_back, double inv_dt) {
    if (t_front > time || t_back <= time) return 0;
    int g = (int) ((time - t_front) * inv_dt) + 1;
    g     = g < 1 ? 1 : (g > w - 1 ? w - 1 : g);
    if (t[g - 1] <= time && time < t[g]) return g;
    int sta = 0, end = w;
    for (int it = 0; it < 17; it++) {
        int mid       = (sta + end) >> 1;
        double first  = t[mid - 1];
        double second = t[mid];
        if (first <= time && time < second) return mid;
        if (first > time)
            end = mid;
        else if (second <= time)
            sta = mid;
    }
    return 0;
}

__host__ __device__ __forceinline__ Q4 ldq(const double *__restrict__ q) { return Q4{q[0], q[1], q[2], q[3]}; }
__host__ __device__ __forceinline__ V3 ld3(const double *__restrict__ p) { return V3{p[0], p[1], p[2]}; }

// MISC::stateToPoseWithExtrinsic (misc.cc:99-105)
__host__ __device__ __forceinline__ void state_to_pose(V3 p, Q4 q, const Pose12 &ext, M3 &R, V3 &t) {
    M3 Rq = quat_to_R(q);
    M3 Re;
#pragma unroll
    for (int i = 0; i < 9; i++) Re.m[i] = ext.R[i];
    t = p + mul(Rq, V3{ext.t[0], ext.t[1], ext.t[2]});
    R = mul(Rq, Re);
}

// MISC::statePoseInterpolation with the interval invariants taken from the table (misc.cc:82-97)
__host__ __device__ __forceinline__ void interp_pose(const double *__restrict__ ins_t, const double *__restrict__ ins_p,
                                            const double *__restrict__ ins_q, const double *__restrict__ tab,
                                            int idx, double time, const Pose12 &ext, M3 &R, V3 &t) {
    double t0 = ins_t[idx - 1], t1 = ins_t[idx];
    double sc = (time - t0) / (t1 - t0);
    V3 dp     = ld3(tab + 6 * idx);
    V3 rvec   = ld3(tab + 6 * idx + 3);
    Q4 dq     = quat_exp(rvec * sc);
    V3 p0     = ld3(ins_p + 3 * (idx - 1));
    Q4 q0     = ldq(ins_q + 4 * (idx - 1));
    V3 p      = p0 + dp * sc;
    Q4 q      = qmul(q0, qinv(dq));
    double nn = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q         = Q4{q.x / nn, q.y / nn, q.z / nn, q.w / nn};
    state_to_pose(p, q, ext, R, t);
}


// ---- per-interval closed form of the relative pose ---------------------------------------------------
// For a point time t inside IMU interval i (entries i-1, i), with s = (t - t0)/(t1 - t0),
// K = [rvec]x, theta = |rvec|, the reference's chain (misc.cc:82-105, pointcloud.cc:30-36) is
//   R(q)  = R(q0) exp(-s K),          p = p0 + s dp,
//   R01   = R0f^T R(q) R_bl,          t01 = R0f^T (p + R(q) t_bl - t0f).
// With exp(-sK) = I - s a(x) K + s^2 b(x) K^2, x = s theta, a = sin(x)/x, b = (1 - cos x)/x^2:
//   R01 = G0 - (s a) G1 + (s^2 b) G2,     t01 = c0 + s c1 - (s a) c2 + (s^2 b) c3,
// where G0 = A R_bl, G1 = A K R_bl, G2 = A K^2 R_bl, A = R0f^T R(q0), c0 = R0f^T (p0 - t0f) + A t_bl,
// c1 = R0f^T dp, c2 = A K t_bl, c3 = A K^2 t_bl depend only on the interval, so the undistorted point is
//   out = (G0 p + c0) - (s a)(G1 p + c2) + (s^2 b)(G2 p + c3) + s c1.
// Row layout (kRowSize = 44 doubles, 16-byte groups for 128-bit loads):
//   [0] t0  [1] 1/(t1-t0)  [2] theta^2  [3] pad
//   [4+4i ..]  G0 row i, c0[i]     [16+4i ..] G1 row i, c2[i]     [28+4i ..] G2 row i, c3[i]   (i = 0..2)
//   [40..42] c1   [43] pad
// Row 0 is the out-of-window fallback (misc.cc:76-79): the constant pose of window.back().
constexpr int kRowSize = 44;

__host__ __device__ inline void interval_row(const double *__restrict__ ins_t, const double *__restrict__ ins_p,
                                             const double *__restrict__ ins_q, int w, int i, const Pose12 &ext,
                                             const Pose12 &frame, double *__restrict__ row) {
    M3 R0f, Rbl;
    for (int k = 0; k < 9; k++) {
        R0f.m[k] = frame.R[k];
        Rbl.m[k] = ext.R[k];
    }
    V3 t0f = V3{frame.t[0], frame.t[1], frame.t[2]}, tbl = V3{ext.t[0], ext.t[1], ext.t[2]};
    for (int k = 0; k < kRowSize; k++) row[k] = 0.0;
    int i0   = (i == 0) ? (w - 1) : (i - 1);
    Q4 q0    = ldq(ins_q + 4 * i0);
    double n = sqrt(q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w);
    q0       = Q4{q0.x / n, q0.y / n, q0.z / n, q0.w / n};
    V3 p0    = ld3(ins_p + 3 * i0);
    M3 A     = mulTN(R0f, quat_to_R(q0));
    M3 G0    = mul(A, Rbl);
    V3 c0    = mulT(R0f, p0 - t0f) + mul(A, tbl);
    const double c0v[3] = {c0.x, c0.y, c0.z};
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) row[4 + 4 * r + c] = G0.m[3 * r + c];
        row[4 + 4 * r + 3] = c0v[r];
    }
    if (i == 0) return; // constant pose: G1 = G2 = c1..c3 = 0, s = 0
    V3 dp   = ld3(ins_p + 3 * i) - p0;
    Q4 dq   = qmul(qinv(ldq(ins_q + 4 * i)), ldq(ins_q + 4 * (i - 1)));
    V3 rv   = quat_log(dq);
    M3 K    = M3{{0.0, -rv.z, rv.y, rv.z, 0.0, -rv.x, -rv.y, rv.x, 0.0}};
    M3 AK   = mul(A, K);
    M3 AKK  = mul(AK, K);
    M3 G1   = mul(AK, Rbl), G2 = mul(AKK, Rbl);
    V3 c1   = mulT(R0f, dp), c2 = mul(AK, tbl), c3 = mul(AKK, tbl);
    const double c2v[3] = {c2.x, c2.y, c2.z}, c3v[3] = {c3.x, c3.y, c3.z};
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
            row[16 + 4 * r + c] = G1.m[3 * r + c];
            row[28 + 4 * r + c] = G2.m[3 * r + c];
        }
        row[16 + 4 * r + 3] = c2v[r];
        row[28 + 4 * r + 3] = c3v[r];
    }
    row[40] = c1.x;
    row[41] = c1.y;
    row[42] = c1.z;
    row[0]  = ins_t[i - 1];
    row[1]  = 1.0 / (ins_t[i] - ins_t[i - 1]);
    row[2]  = rv.x * rv.x + rv.y * rv.y + rv.z * rv.z;
}

// sin(x)/x and (1 - cos x)/x^2 from x^2
__host__ __device__ __forceinline__ void sinc_cosc(double x2, double &a, double &b) {
    if (x2 < 2.5e-3) { // |x| < 0.05: 6-term series, truncation error < 1e-25
        a = 1.0 + x2 * (-1.0 / 6 + x2 * (1.0 / 120 + x2 * (-1.0 / 5040 + x2 * (1.0 / 362880 + x2 * (-1.0 / 39916800)))));
        b = 0.5 + x2 * (-1.0 / 24 + x2 * (1.0 / 720 + x2 * (-1.0 / 40320 + x2 * (1.0 / 3628800 + x2 * (-1.0 / 479001600)))));
    } else {
        double x = sqrt(x2), sn, cs;
        sincos(x, &sn, &cs);
        a = sn / x;
        // 1 - cos x = 2 sin^2(x/2): no cancellation
        double sh = sin(0.5 * x);
        b = 2.0 * sh * sh / x2;
    }
}

// interpolation weights (s, s a, s^2 b) of a point time
__host__ __device__ __forceinline__ void row_weights(const double *__restrict__ row, double time, double &s, double &sa,
                                                     double &sb) {
    s = (time - row[0]) * row[1];
    double a, b;
    sinc_cosc(s * s * row[2], a, b);
    sa = s * a;
    sb = s * s * b;
}

// undistorted coordinates (fp64) of point (px, py, pz) taken at `time`, from its interval row
__host__ __device__ __forceinline__ void row_apply(const double *__restrict__ row, double time, double px, double py,
                                                   double pz, double out[3]) {
    double s, sa, sb;
    row_weights(row, time, s, sa, sb);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double A0 = row[4 + 4 * i] * px + row[5 + 4 * i] * py + row[6 + 4 * i] * pz + row[7 + 4 * i];
        double A1 = row[16 + 4 * i] * px + row[17 + 4 * i] * py + row[18 + 4 * i] * pz + row[19 + 4 * i];
        double A2 = row[28 + 4 * i] * px + row[29 + 4 * i] * py + row[30 + 4 * i] * pz + row[31 + 4 * i];
        out[i]    = A0 - sa * A1 + sb * A2 + s * row[40 + i];
    }
}

// MISC::getPoseFromInsWindow (misc.cc:64-80) for ONE time, written out in full (used on the host for the
// frame pose at `stamp`, lidar_map.cc:223)
__host__ __device__ inline bool pose_from_window(const double *ins_t, const double *ins_p, const double *ins_q, int w,
                                                 const Pose12 &ext, double time, Pose12 &out) {
    double inv_dt = (double) (w - 1) / (ins_t[w - 1] - ins_t[0]);
    int idx       = window_index(ins_t, w, time, ins_t[0], ins_t[w - 1], inv_dt);
    M3 R;
    V3 t;
    if (idx > 0) {
        double tab[12];
        V3 dp   = ld3(ins_p + 3 * idx) - ld3(ins_p + 3 * (idx - 1));
        Q4 dq   = qmul(qinv(ldq(ins_q + 4 * idx)), ldq(ins_q + 4 * (idx - 1)));
        V3 rvec = quat_log(dq);
        tab[6] = dp.x; tab[7] = dp.y; tab[8] = dp.z;
        tab[9] = rvec.x; tab[10] = rvec.y; tab[11] = rvec.z;
        interp_pose(ins_t + (idx - 1), ins_p + 3 * (idx - 1), ins_q + 4 * (idx - 1), tab, 1, time, ext, R, t);
    } else {
        state_to_pose(ld3(ins_p + 3 * (w - 1)), ldq(ins_q + 4 * (w - 1)), ext, R, t);
    }
    for (int k = 0; k < 9; k++) out.R[k] = R.m[k];
    out.t[0] = t.x;
    out.t[1] = t.y;
    out.t[2] = t.z;
    return idx > 0;
}

} // namespace fflins
