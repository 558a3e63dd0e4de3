// common.cuh — shared device/host helpers of the fflins B200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fflins {

// ---- tiny fixed-size fp64 algebra -----------------------------------------------------------
struct V3 {
    double x, y, z;
};
struct M3 {
    double m[9]; // row-major
};
struct Q4 {
    double x, y, z, w;
};
struct PoseD {
    M3 R;
    V3 t;
};

__host__ __device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// A v
__host__ __device__ __forceinline__ V3 mul(const M3 &A, V3 v) {
    return V3{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
              A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
// A^T v
__host__ __device__ __forceinline__ V3 mulT(const M3 &A, V3 v) {
    return V3{A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
              A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
__host__ __device__ __forceinline__ M3 mul(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
// A^T B
__host__ __device__ __forceinline__ M3 mulTN(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
    return C;
}
__host__ __device__ __forceinline__ M3 transpose(const M3 &A) {
    return M3{{A.m[0], A.m[3], A.m[6], A.m[1], A.m[4], A.m[7], A.m[2], A.m[5], A.m[8]}};
}
// Eigen Quaternion::toRotationMatrix (no normalisation), q = x,y,z,w
__host__ __device__ __forceinline__ M3 quat_to_R(Q4 q) {
    double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return M3{{1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz), tyz - twx, txz - twy,
               tyz + twx, 1.0 - (txx + tyy)}};
}
__host__ __device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__host__ __device__ __forceinline__ Q4 qinv(Q4 q) {
    double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    return Q4{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}

#ifdef __CUDACC__
// ---- exactly-rounded (no FMA contraction) rigid projection: fp64 math, fp32 store --------------
// PointCloudCommon::pointProjection (lidar/pointcloud.cc:53-59) with each product and sum rounded
// separately, ((r0*x + r1*y) + r2*z) + t, as the reference's no-FMA x86-64 build does.
__device__ __forceinline__ double dot3_rn(double a0, double a1, double a2, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, x), __dmul_rn(a1, y)), __dmul_rn(a2, z));
}
__device__ __forceinline__ float4 project_rn(const double *__restrict__ R, const double *__restrict__ t, float4 p) {
    double x = (double) p.x, y = (double) p.y, z = (double) p.z;
    float4 o;
    o.x = (float) __dadd_rn(dot3_rn(R[0], R[1], R[2], x, y, z), t[0]);
    o.y = (float) __dadd_rn(dot3_rn(R[3], R[4], R[5], x, y, z), t[1]);
    o.z = (float) __dadd_rn(dot3_rn(R[6], R[7], R[8], x, y, z), t[2]);
    o.w = p.w;
    return o;
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int lane_mask) {
    return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}
// lexicographic minimum of a 64-bit key over the warp with two hardware reductions (redux.sync): the high words
// first, then the low words of the lanes that hold the minimal high word
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 800
    // (only reached by tests/host_math_check.cu, which compiles this header for nvcc's default architecture)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long u = shfl_xor_u64(v, o);
        v = u < v ? u : v;
    }
    return v;
#else
    const unsigned hi = (unsigned) (v >> 32);
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = hi == mh ? (unsigned) v : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
    return ((unsigned long long) mh << 32) | ml;
#endif
}
#endif

// 12 doubles: R row-major then t — how poses travel to kernels
struct Pose12 {
    double R[9];
    double t[3];
};

} // namespace fflins
