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

} // namespace fflins
