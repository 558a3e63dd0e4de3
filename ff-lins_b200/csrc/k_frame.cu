// k_frame.cu — K1 INS-centric per-point undistortion (+ index decimation) and K3 rigid projection.
//
// Reference behaviour restated (not translated):
//   LidarMap::lidarFrameProcessing        ff_lins/lidar/lidar_map.cc:213-273
//   MISC::getInsWindowIndex / getPoseFromInsWindow / statePoseInterpolation / stateToPoseWithExtrinsic
//                                          ff_lins/common/misc.cc:27-105
//   PointCloudCommon::relativeProjection / pointCloudProjection / pointProjection
//                                          ff_lins/lidar/pointcloud.cc:30-59
//
// B200 design: the quaternion log of each IMU interval (atan2 + divisions) depends only on the
// interval, not on the point, so a W-thread prologue builds a 48-byte-per-interval table once per
// frame; the per-point kernel is then one table lookup (L1-resident, warp-broadcast because a
// warp's 32 consecutive points almost always share an interval), one sincos, two 3x3 products and
// an exactly-rounded fp64->fp32 projection. Points stream as float4 (16 B) + double (8 B) in,
// float4 out: 40 B/point of HBM traffic.
#include "math_undistort.cuh"

namespace fflins {

namespace {

__global__ void ins_prep_kernel(const double *__restrict__ ins_t, const double *__restrict__ ins_p,
                                const double *__restrict__ ins_q, int w, Pose12 ext, double stamp,
                                double *__restrict__ tab, Pose12 *__restrict__ frame_pose) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 1 && i < w) {
        V3 dp   = ld3(ins_p + 3 * i) - ld3(ins_p + 3 * (i - 1));
        Q4 dq   = qmul(qinv(ldq(ins_q + 4 * i)), ldq(ins_q + 4 * (i - 1)));
        V3 rvec = quat_log(dq);
        tab[6 * i + 0] = dp.x;
        tab[6 * i + 1] = dp.y;
        tab[6 * i + 2] = dp.z;
        tab[6 * i + 3] = rvec.x;
        tab[6 * i + 4] = rvec.y;
        tab[6 * i + 5] = rvec.z;
    }
    if (i == 0) {
        // window.back() pose: the out-of-window fallback (misc.cc:76-79)
        M3 R;
        V3 t;
        state_to_pose(ld3(ins_p + 3 * (w - 1)), ldq(ins_q + 4 * (w - 1)), ext, R, t);
#pragma unroll
        for (int k = 0; k < 9; k++) frame_pose[1].R[k] = R.m[k];
        frame_pose[1].t[0] = t.x;
        frame_pose[1].t[1] = t.y;
        frame_pose[1].t[2] = t.z;
        // frame pose at `stamp` (lidar_map.cc:223); needs its own interval invariants
        double inv_dt = (double) (w - 1) / (ins_t[w - 1] - ins_t[0]);
        int idx       = window_index(ins_t, w, stamp, ins_t[0], ins_t[w - 1], inv_dt);
        if (idx > 0) {
            V3 dp   = ld3(ins_p + 3 * idx) - ld3(ins_p + 3 * (idx - 1));
            Q4 dq   = qmul(qinv(ldq(ins_q + 4 * idx)), ldq(ins_q + 4 * (idx - 1)));
            V3 rvec = quat_log(dq);
            double loc[12];
            loc[6 * 1 + 0] = dp.x;
            loc[6 * 1 + 1] = dp.y;
            loc[6 * 1 + 2] = dp.z;
            loc[6 * 1 + 3] = rvec.x;
            loc[6 * 1 + 4] = rvec.y;
            loc[6 * 1 + 5] = rvec.z;
            // interp_pose indexes tab by idx: shift the base so that entry `idx` is loc[6..11]
            interp_pose(ins_t, ins_p, ins_q, loc - 6 * (idx - 1), idx, stamp, ext, R, t);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) frame_pose[0].R[k] = R.m[k];
        frame_pose[0].t[0] = t.x;
        frame_pose[0].t[1] = t.y;
        frame_pose[0].t[2] = t.z;
    }
}

struct RawAoS {   // PointXYZIT, ff_lins/lidar/lidar.h:30-38 (32 bytes)
    float4 xyz_pad;
    float intensity;
    float pad;
    double time;
};

template <bool AOS>
__global__ void __launch_bounds__(256)
undistort_kernel(const float4 *__restrict__ xyzi, const double *__restrict__ time, const RawAoS *__restrict__ aos,
                 int n, const double *__restrict__ ins_t, const double *__restrict__ ins_p,
                 const double *__restrict__ ins_q, int w, const double *__restrict__ tab,
                 const Pose12 *__restrict__ frame_pose, Pose12 ext, double td, int filter_num,
                 float4 *__restrict__ out_full, float4 *__restrict__ out_dec, int *__restrict__ n_fallback) {
    __shared__ Pose12 s_pose[2];
    if (threadIdx.x < 24) reinterpret_cast<double *>(s_pose)[threadIdx.x] = reinterpret_cast<const double *>(frame_pose)[threadIdx.x];
    __syncthreads();
    const double t_front = ins_t[0], t_back = ins_t[w - 1];
    const double inv_dt  = (double) (w - 1) / (t_back - t_front);
    M3 R0;
#pragma unroll
    for (int i = 0; i < 9; i++) R0.m[i] = s_pose[0].R[i];
    const V3 t0 = V3{s_pose[0].t[0], s_pose[0].t[1], s_pose[0].t[2]};

    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        float4 p;
        double tp;
        if (AOS) {
            const float4 *rec = reinterpret_cast<const float4 *>(aos + k);
            float4 a          = __ldg(rec);
            float4 b          = __ldg(rec + 1);
            p                 = make_float4(a.x, a.y, a.z, b.x);
            tp                = __hiloint2double(__float_as_int(b.w), __float_as_int(b.z));
        } else {
            p  = __ldg(xyzi + k);
            tp = __ldg(time + k);
        }
        double tt = tp + td;
        int idx   = window_index(ins_t, w, tt, t_front, t_back, inv_dt);
        M3 R1;
        V3 t1;
        if (idx > 0) {
            interp_pose(ins_t, ins_p, ins_q, tab, idx, tt, ext, R1, t1);
        } else {
#pragma unroll
            for (int i = 0; i < 9; i++) R1.m[i] = s_pose[1].R[i];
            t1 = V3{s_pose[1].t[0], s_pose[1].t[1], s_pose[1].t[2]};
            atomicAdd(n_fallback, 1);
        }
        // relativeProjection (pointcloud.cc:30-36)
        M3 R01 = mulTN(R0, R1);
        V3 t01 = mulT(R0, t1 - t0);
        double tv[3] = {t01.x, t01.y, t01.z};
        float4 o     = project_rn(R01.m, tv, p);
        out_full[k]  = o;
        // index decimation (lidar_map.cc:249-256)
        if (out_dec != nullptr && (k % filter_num) == 0) out_dec[k / filter_num] = o;
    }
}

__global__ void copy_decimate_kernel(const float4 *__restrict__ in, int n, int filter_num, float4 *__restrict__ out_full,
                                     float4 *__restrict__ out_dec) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        float4 p    = __ldg(in + k);
        out_full[k] = p;
        if ((k % filter_num) == 0) out_dec[k / filter_num] = p;
    }
}

__global__ void __launch_bounds__(256)
project_kernel(Pose12 T, const float4 *__restrict__ src, float4 *__restrict__ dst, int n, const int *__restrict__ n_dev) {
    if (n_dev) n = min(n, *n_dev);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = project_rn(T.R, T.t, src[k]);
}

__global__ void __launch_bounds__(256)
project_dev_kernel(const Pose12 *__restrict__ Tp, const float4 *__restrict__ src, float4 *__restrict__ dst, int n,
                   const int *__restrict__ n_dev) {
    __shared__ Pose12 T;
    if (threadIdx.x < 12) reinterpret_cast<double *>(&T)[threadIdx.x] = reinterpret_cast<const double *>(Tp)[threadIdx.x];
    __syncthreads();
    if (n_dev) n = min(n, *n_dev);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = project_rn(T.R, T.t, src[k]);
}

__global__ void __launch_bounds__(256) project_batch_kernel(const __grid_constant__ ProjectBatch B) {
    const int c = blockIdx.y;
    const int n = B.n[c];
    const float4 *__restrict__ src = B.src[c];
    float4 *__restrict__ dst       = B.dst[c];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = project_rn(B.T[c].R, B.T[c].t, src[k]);
}

inline int grid_for(int n, int block, int per_sm = 8) {
    int g = (n + block - 1) / block;
    int cap = 148 * per_sm;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

} // namespace

void launch_ins_prep(cudaStream_t s, const double *ins_t, const double *ins_p, const double *ins_q, int w, Pose12 ext,
                     double stamp, double *interval_tab, Pose12 *frame_pose) {
    ins_prep_kernel<<<(w + 127) / 128, 128, 0, s>>>(ins_t, ins_p, ins_q, w, ext, stamp, interval_tab, frame_pose);
}

void launch_undistort(cudaStream_t s, const float4 *xyzi, const double *time, const void *aos, int n,
                      const double *ins_t, const double *ins_p, const double *ins_q, int w,
                      const double *interval_tab, const Pose12 *frame_pose, Pose12 ext, double td, int filter_num,
                      float4 *out_full, float4 *out_dec, int *n_fallback) {
    if (n <= 0) return;
    int grid = grid_for(n, 256, 8);
    if (aos)
        undistort_kernel<true><<<grid, 256, 0, s>>>(nullptr, nullptr, static_cast<const RawAoS *>(aos), n, ins_t, ins_p,
                                                    ins_q, w, interval_tab, frame_pose, ext, td, filter_num, out_full,
                                                    out_dec, n_fallback);
    else
        undistort_kernel<false><<<grid, 256, 0, s>>>(xyzi, time, nullptr, n, ins_t, ins_p, ins_q, w, interval_tab,
                                                     frame_pose, ext, td, filter_num, out_full, out_dec, n_fallback);
}

void launch_copy_decimate(cudaStream_t s, const float4 *xyzi, int n, int filter_num, float4 *out_full, float4 *out_dec) {
    if (n <= 0) return;
    copy_decimate_kernel<<<grid_for(n, 256), 256, 0, s>>>(xyzi, n, filter_num, out_full, out_dec);
}

void launch_project(cudaStream_t s, Pose12 T, const float4 *src, float4 *dst, int n, const int *n_dev) {
    if (n <= 0) return;
    project_kernel<<<grid_for(n, 256), 256, 0, s>>>(T, src, dst, n, n_dev);
}

void launch_project_batch(cudaStream_t s, const ProjectBatch &b) {
    int nmax = 0;
    for (int i = 0; i < b.count; i++) nmax = b.n[i] > nmax ? b.n[i] : nmax;
    if (b.count <= 0 || nmax <= 0) return;
    dim3 grid(grid_for(nmax, 256), b.count);
    project_batch_kernel<<<grid, 256, 0, s>>>(b);
}

void launch_project_dev(cudaStream_t s, const Pose12 *T_dev, const float4 *src, float4 *dst, int n, const int *n_dev) {
    if (n <= 0) return;
    project_dev_kernel<<<grid_for(n, 256), 256, 0, s>>>(T_dev, src, dst, n, n_dev);
}

} // namespace fflins
