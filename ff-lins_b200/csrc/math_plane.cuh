// math_plane.cuh — 5-point plane fit (column-pivoted Householder QR) shared by K5 (k_assoc.cu).
// __host__ __device__ so that tests can run the same code on the CPU (tests/host_math_check.cu).
// Must be compiled with -fmad=false for bit-reproducible normals.
#pragma once
#include "kernels.h"
#include <cfloat>
#include <cmath>

namespace fflins {

// Eigen ColPivHouseholderQR<5x3>::compute + solve restated, register-resident:
// every index is static after unrolling, column swaps are predicated.
__host__ __device__ __forceinline__ void colpiv_qr_solve_5x3(double A[5][3], double x[3]) {
    double hc[3], upd[3], dir[3];
    int trans[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) s += A[i][k] * A[i][k];
        upd[k] = dir[k] = sqrt(s);
    }
    const double eps = 2.220446049250313e-16;
    double maxn      = fmax(upd[0], fmax(upd[1], upd[2]));
    double th        = maxn * eps / 5.0;
    const double threshold_helper = th * th;
    const double downdate_thr     = 1.4901161193847656e-08; // sqrt(eps)
    int nonzero = 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int big     = k;
        double bigv = upd[k];
#pragma unroll
        for (int j = k + 1; j < 3; j++)
            if (upd[j] > bigv) { bigv = upd[j]; big = j; }
        if (nonzero == 3 && bigv * bigv < threshold_helper * (double) (5 - k)) nonzero = k;
        trans[k] = big;
#pragma unroll
        for (int j = k + 1; j < 3; j++) {
            if (j == big) {
#pragma unroll
                for (int i = 0; i < 5; i++) { double t = A[i][k]; A[i][k] = A[i][j]; A[i][j] = t; }
                double t = upd[k]; upd[k] = upd[j]; upd[j] = t;
                t = dir[k]; dir[k] = dir[j]; dir[j] = t;
            }
        }
        double tail_sq = 0;
#pragma unroll
        for (int i = k + 1; i < 5; i++) tail_sq += A[i][k] * A[i][k];
        double c0 = A[k][k], beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau  = 0;
            beta = c0;
#pragma unroll
            for (int i = k + 1; i < 5; i++) A[i][k] = 0;
        } else {
            beta = sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
#pragma unroll
            for (int i = k + 1; i < 5; i++) A[i][k] = A[i][k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hc[k]   = tau;
        A[k][k] = beta;
        if (tau != 0) {
#pragma unroll
            for (int j = k + 1; j < 3; j++) {
                double tmp = 0;
#pragma unroll
                for (int i = k + 1; i < 5; i++) tmp += A[i][k] * A[i][j];
                tmp += A[k][j];
                A[k][j] -= tau * tmp;
#pragma unroll
                for (int i = k + 1; i < 5; i++) A[i][j] -= tau * A[i][k] * tmp;
            }
        }
#pragma unroll
        for (int j = k + 1; j < 3; j++) {
            if (upd[j] != 0) {
                double temp  = fabs(A[k][j]) / upd[j];
                temp         = (1.0 + temp) * (1.0 - temp);
                temp         = temp < 0 ? 0 : temp;
                double ratio = upd[j] / dir[j];
                double temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate_thr) {
                    double s = 0;
#pragma unroll
                    for (int i = k + 1; i < 5; i++) s += A[i][j] * A[i][j];
                    dir[j] = sqrt(s);
                    upd[j] = dir[j];
                } else {
                    upd[j] *= sqrt(temp);
                }
            }
        }
    }
    double c[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
    x[0] = x[1] = x[2] = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k < nonzero && hc[k] != 0) {
            double tmp = 0;
#pragma unroll
            for (int i = k + 1; i < 5; i++) tmp += A[i][k] * c[i];
            tmp += c[k];
            c[k] -= hc[k] * tmp;
#pragma unroll
            for (int i = k + 1; i < 5; i++) c[i] -= hc[k] * A[i][k] * tmp;
        }
    }
#pragma unroll
    for (int i = 2; i >= 0; i--) {
        if (i < nonzero) {
            double s = 0;
#pragma unroll
            for (int j = i + 1; j < 3; j++)
                if (j < nonzero) s += A[i][j] * c[j];
            c[i] = (c[i] - s) / A[i][i];
            x[i] = c[i];
        }
    }
#pragma unroll
    for (int k = 2; k >= 0; k--) {
#pragma unroll
        for (int j = k + 1; j < 3; j++)
            if (trans[k] == j) { double t = x[k]; x[k] = x[j]; x[j] = t; }
    }
}

// PointCloudCommon::planeEstimation (pointcloud.cc:61-93)
__host__ __device__ __forceinline__ bool plane_estimation(const float4 pts[5], double thr, double unit[3], double &ninv,
                                                 double dist[5]) {
    double A[5][3];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        A[k][0] = (double) pts[k].x;
        A[k][1] = (double) pts[k].y;
        A[k][2] = (double) pts[k].z;
    }
    double nrm[3];
    colpiv_qr_solve_5x3(A, nrm);
    double norm = sqrt((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
    ninv        = 1.0 / norm;
    unit[0]     = nrm[0] * ninv;
    unit[1]     = nrm[1] * ninv;
    unit[2]     = nrm[2] * ninv;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double dis = ok ? ((unit[0] * (double) pts[k].x + unit[1] * (double) pts[k].y) + unit[2] * (double) pts[k].z) + ninv
                        : 0.0;
        dist[k] = dis;
        if (fabs(dis) > thr) ok = false; // the reference returns at the first violation
    }
    return ok;
}

} // namespace fflins
