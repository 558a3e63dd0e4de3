// math_factor.cuh — fp64 point-to-plane factor algebra shared by the K6/K7 kernels (k_factor.cu).
// __host__ __device__ so that tests can run the very same code on the CPU against the oracle
// before any GPU time is spent (tests/host_math_check.cu).
#pragma once
#include "kernels.h"
#include <cfloat>
#include <cmath>

namespace fflins {

__host__ __device__ __forceinline__ M3 skew_plus_I(V3 v) { return M3{{1.0, -v.z, v.y, v.z, 1.0, -v.x, -v.y, v.x, 1.0}}; }

// the part that depends only on the current frame, the extrinsic and td: once per evaluation
__host__ __device__ inline void make_cur_const_raw(const double *pose_cur, const double *ext, double td, const double *v1p,
                                                   const double *w1p, double td1, CurConst &k) {
    V3 t1  = v3(pose_cur[0], pose_cur[1], pose_cur[2]);
    Q4 q1  = Q4{pose_cur[3], pose_cur[4], pose_cur[5], pose_cur[6]};
    k.tbl  = v3(ext[0], ext[1], ext[2]);
    Q4 qbl = Q4{ext[3], ext[4], ext[5], ext[6]};
    k.w1   = v3(w1p[0], w1p[1], w1p[2]);
    V3 v1  = v3(v1p[0], v1p[1], v1p[2]);
    double dt1 = td - td1;
    k.R1  = quat_to_R(q1);
    k.B1  = skew_plus_I(k.w1 * dt1);
    k.Rt1 = mul(k.R1, k.B1);
    k.tt1 = t1 + v1 * dt1;
    k.Rbl = quat_to_R(qbl);
}

// the per-reference part, first half: independent of the current frame
__host__ __device__ inline void make_ref_const_a(const double *pose_ref, const double *v0p, const double *w0p, double td0,
                                                 double td, const double *v1p, RefConst &c) {
    V3 t0  = v3(pose_ref[0], pose_ref[1], pose_ref[2]);
    Q4 q0  = Q4{pose_ref[3], pose_ref[4], pose_ref[5], pose_ref[6]};
    c.w0   = v3(w0p[0], w0p[1], w0p[2]);
    V3 v0  = v3(v0p[0], v0p[1], v0p[2]);
    V3 v1  = v3(v1p[0], v1p[1], v1p[2]);
    c.dv   = v1 - v0;
    double dt0 = td - td0;
    c.R0  = quat_to_R(q0);
    c.B0  = skew_plus_I(c.w0 * dt0);   // I + [w0 dt0]x   (lidar_factor.cc:59)
    c.Rt0 = mul(c.R0, c.B0);
    c.tt0 = t0 + v0 * dt0;
}
// ... second half: the products with the current-frame part
__host__ __device__ inline void make_ref_const_b(const CurConst &k, RefConst &c) {
    M3 A  = mul(transpose(k.Rbl), transpose(c.Rt0));
    M3 C  = mul(A, k.Rt1);             // `common` (lidar_factor.cc:102)
    M3 Rblt = transpose(k.Rbl);
#pragma unroll
    for (int i = 0; i < 9; i++) c.D.m[i] = C.m[i] - Rblt.m[i];
    c.E = mul(C, k.Rbl);
}

__host__ __device__ inline void make_cur_const(const FactorParams &P, CurConst &k) {
    make_cur_const_raw(P.pose_cur, P.ext, P.td, P.v1, P.w1, P.td1, k);
}
__host__ __device__ inline void make_ref_const(const FactorParams &P, const FactorPair &pr, const CurConst &k, RefConst &c) {
    make_ref_const_a(pr.pose_ref, pr.v0, pr.w0, pr.td0, P.td, P.v1, c);
    make_ref_const_b(k, c);
}

__host__ __device__ inline void make_pair_const(const FactorParams &P, const FactorPair &pr, PairConst &pc) {
    make_cur_const(P, pc.cur);
    make_ref_const(P, pr, pc.cur, pc.ref);
}

// r and the 19 tangent-space Jacobian entries [ref t, ref th, cur t, cur th, ext t, ext th, td].
template <bool JAC>
__host__ __device__ __forceinline__ double eval_residual(const CurConst &k, const RefConst &c, V3 p, V3 n, double d, double s, double *J) {
    V3 pb1 = mul(k.Rbl, p) + k.tbl;
    V3 pw  = mul(k.Rt1, pb1) + k.tt1;
    V3 u   = pw - c.tt0;
    V3 pb0 = mulT(c.Rt0, u);
    V3 pl0 = mulT(k.Rbl, pb0 - k.tbl);
    double r = s * (dot(n, pl0) + d);
    if (JAC) {
        V3 sn = n * s;
        V3 sc = mul(k.Rbl, sn);           // (s n^T Rbl^T)^T
        V3 g0 = mul(c.Rt0, sc);           // (sc^T Rt0^T)^T
        V3 w  = mulT(c.R0, u);            // R0^T (pw - tt0)
        V3 a  = mul(c.B0, sc);
        V3 jr = cross(a, w);              // sc^T B0^T [w]x
        V3 h  = mulT(k.R1, g0);
        V3 m  = mul(k.B1, pb1);
        V3 jc = cross(m, h);              // -g0^T R1 [B1 pb1]x
        V3 je = mulT(c.D, sn);            // s n^T (C - Rbl^T)
        V3 kk = mulT(c.E, sn);
        V3 jf = cross(sn, pl0) - cross(kk, p); // s n^T ([pl0]x - C Rbl [p]x)
        V3 z  = mul(k.R1, cross(k.w1, pb1)) + c.dv;
        double jt = dot(sc, cross(w, c.w0) + mulT(c.Rt0, z));
        J[0] = -g0.x; J[1] = -g0.y; J[2] = -g0.z;
        J[3] = jr.x;  J[4] = jr.y;  J[5] = jr.z;
        J[6] = g0.x;  J[7] = g0.y;  J[8] = g0.z;
        J[9] = jc.x;  J[10] = jc.y; J[11] = jc.z;
        J[12] = je.x; J[13] = je.y; J[14] = je.z;
        J[15] = jf.x; J[16] = jf.y; J[17] = jf.z;
        J[18] = jt;
    }
    return r;
}

// Ceres HuberLoss + corrector for a 1-dim residual (residual_block_info.cc:49-77). Returns rho(s).
__host__ __device__ __forceinline__ double huber_correct(double delta, double &r, double *J) {
    double sq = r * r, b = delta * delta;
    double rho0, rho1, rho2;
    if (sq > b) {
        double rr = sqrt(sq);
        rho0 = 2.0 * delta * rr - b;
        rho1 = fmax(DBL_MIN, delta / rr);
        rho2 = -rho1 / (2.0 * sq);
    } else {
        rho0 = sq; rho1 = 1.0; rho2 = 0.0;
    }
    double sqrt_rho1 = sqrt(rho1), scaling, alpha_sq;
    if (sq == 0.0 || rho2 <= 0.0) {
        scaling = sqrt_rho1;
        alpha_sq = 0.0;
    } else {
        double D     = 1.0 + 2.0 * sq * rho2 / rho1;
        double alpha = 1.0 - sqrt(D);
        scaling  = sqrt_rho1 / (1.0 - alpha);
        alpha_sq = alpha / sq;
    }
    if (J) {
#pragma unroll
        for (int i = 0; i < 19; i++) J[i] = sqrt_rho1 * (J[i] - alpha_sq * r * (r * J[i]));
    }
    r *= scaling;
    return rho0;
}

} // namespace fflins
