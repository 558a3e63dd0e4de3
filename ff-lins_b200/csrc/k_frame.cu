// k_frame.cu — K1 INS-centric per-point undistortion (+ index decimation) and K3 rigid projection.
//
// Reference behaviour restated (not translated):
//   LidarMap::lidarFrameProcessing        ff_lins/lidar/lidar_map.cc:213-273
//   MISC::getInsWindowIndex / getPoseFromInsWindow / statePoseInterpolation / stateToPoseWithExtrinsic
//                                          ff_lins/common/misc.cc:27-105
//   PointCloudCommon::relativeProjection / pointCloudProjection / pointProjection
//                                          ff_lins/lidar/pointcloud.cc:30-59
//
// B200 design: everything the reference recomputes per point but that depends only on the IMU
// interval — the quaternion log (atan2, divisions), the products with the frame pose and the
// extrinsic — is folded by a W-thread prologue into one 42-double row per interval (closed form in
// math_undistort.cuh). The per-point kernel is then: bracket lookup, one row read (L1-resident,
// warp-broadcast because a warp's 32 consecutive points almost always share an interval), two short
// polynomials, 27 FMAs and an exactly-rounded fp64->fp32 projection: ~100 fp64 operations instead of
// ~1500, which moves the kernel from the FP64 pipe to the HBM roofline. Points stream as float4
// (16 B) + double (8 B) in, float4 out: 40 B/point.
#include "math_undistort.cuh"

namespace fflins {

namespace {

// One thread per window entry: row i is the closed-form relative pose of interval [i-1, i]; row 0 the
// out-of-window fallback. The quaternion log (atan2 + divisions) happens here, once per interval.
// The window is read straight out of the caller's pinned staging slot (zero-copy over PCIe: 64 bytes per entry,
// no memcpy / event / cross-stream wait in front of the kernel); the time stamps the undistortion kernel needs
// for its interval lookup are copied to device memory on the way.
__global__ void ins_prep_kernel(const __grid_constant__ FrameBatch B) {
    const FrameItem &it = B.it[blockIdx.y];
    const int w = it.w;
    const double *__restrict__ ins_t = it.win, *__restrict__ ins_p = it.win + w, *__restrict__ ins_q = it.win + 4 * (size_t) w;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 4) it.counters[i] = 0; // per-frame counters (fallbacks, voxel count): saves a memset in the stream
    if (i >= w) return;
    it.t_dev[i] = ins_t[i];
    double row[kRowSize];
    interval_row(ins_t, ins_p, ins_q, w, i, it.ext, it.frame, row);
#pragma unroll
    for (int k = 0; k < kRowSize; k++) it.rows[(size_t) i * kRowSize + k] = row[k];
}

struct alignas(32) F8 {
    float4 a, b;
};
__device__ __forceinline__ F8 ld256_stream(const void *p) {
    F8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld128_stream_f64(const double *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(void *p, const F8 &v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z), "f"(v.a.w),
                 "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
}

struct RawAoS {   // PointXYZIT, ff_lins/lidar/lidar.h:30-38 (32 bytes)
    float4 xyz_pad;
    float intensity;
    float pad;
    double time;
};

// Two consecutive points per thread: they almost always share the IMU interval, so each 16-byte group of
// the 44-double row is fetched once (warp-broadcast out of L1) and applied to both points as it arrives.
// The first version issued 42 8-byte row loads per point and was bound by load-instruction issue, not HBM.
template <bool AOS>
__global__ void __launch_bounds__(128, 8) undistort_kernel(const __grid_constant__ FrameBatch B) {
    const FrameItem &it = B.it[blockIdx.y];
    const float4 *__restrict__ xyzi = it.xyzi;
    const double *__restrict__ time = it.time;
    const RawAoS *__restrict__ aos  = static_cast<const RawAoS *>(it.aos);
    const int n = it.n, w = it.w, filter_num = it.filter_num;
    const double *__restrict__ ins_t = it.t_dev;
    const double *__restrict__ rows  = it.rows;
    const double td = it.td;
    float4 *__restrict__ out_full = it.out_full;
    float4 *__restrict__ out_dec  = it.out_dec;
    int *__restrict__ n_fallback  = it.counters;
    if (n <= 0) return;
    const double t_front = it.t_front, t_back = it.t_back, inv_dt = it.inv_dt;
    const unsigned magic = it.dec_magic;
    // 32-byte accesses need 32-byte aligned arrays (device pools and page-locked buffers are; a caller's odd pointer is not)
    const bool wide = ((reinterpret_cast<uintptr_t>(out_full) | (AOS ? reinterpret_cast<uintptr_t>(aos) : (reinterpret_cast<uintptr_t>(xyzi) | (reinterpret_cast<uintptr_t>(time) << 1)))) & 31u) == 0;
    const int pairs      = (n + 1) >> 1;
    for (int pr = blockIdx.x * blockDim.x + threadIdx.x; pr < pairs; pr += gridDim.x * blockDim.x) {
        const int k0 = 2 * pr;
        float4 p[2];
        double tt[2];
        int idx[2];
        if (wide && k0 + 1 < n) {
            // one 256-bit request per 32-byte sector (streaming, no L1 allocation): the two points of the pair, or one raw record
            if (AOS) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const F8 r = ld256_stream(aos + k0 + j);
                    p[j]       = make_float4(r.a.x, r.a.y, r.a.z, r.b.x);
                    tt[j]      = __hiloint2double(__float_as_int(r.b.w), __float_as_int(r.b.z)) + td;
                }
            } else {
                const F8 r      = ld256_stream(xyzi + k0);
                const double2 t = ld128_stream_f64(time + k0);
                p[0] = r.a; p[1] = r.b;
                tt[0] = t.x + td; tt[1] = t.y + td;                   // lidar_map.cc:237
            }
        } else {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int k = min(k0 + j, n - 1);
                if (AOS) {
                    const float4 *rec = reinterpret_cast<const float4 *>(aos + k);
                    float4 a          = __ldg(rec);
                    float4 b          = __ldg(rec + 1);
                    p[j]              = make_float4(a.x, a.y, a.z, b.x);
                    tt[j]             = __hiloint2double(__float_as_int(b.w), __float_as_int(b.z)) + td;
                } else {
                    p[j]  = __ldg(xyzi + k);
                    tt[j] = __ldg(time + k) + td;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            idx[j] = window_index(ins_t, w, tt[j], t_front, t_back, inv_dt);
            if (idx[j] == 0 && k0 + j < n) atomicAdd(n_fallback, 1);
        }
        double o[2][3];
        if (idx[0] == idx[1]) {
            const double2 *rp = reinterpret_cast<const double2 *>(rows + (size_t) idx[0] * kRowSize);
            double2 h0 = __ldg(rp), h1 = __ldg(rp + 1);               // t0, inv_dt | theta^2, pad
            double s[2], sa[2], sb[2];
            const double hdr[3] = {h0.x, h0.y, h1.x};
#pragma unroll
            for (int j = 0; j < 2; j++) row_weights(hdr, tt[j], s[j], sa[j], sb[j]);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                double2 g0a = __ldg(rp + 2 + 2 * i), g0b = __ldg(rp + 3 + 2 * i);
                double2 g1a = __ldg(rp + 8 + 2 * i), g1b = __ldg(rp + 9 + 2 * i);
                double2 g2a = __ldg(rp + 14 + 2 * i), g2b = __ldg(rp + 15 + 2 * i);
                double c1   = __ldg(rows + (size_t) idx[0] * kRowSize + 40 + i);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    double px = (double) p[j].x, py = (double) p[j].y, pz = (double) p[j].z;
                    double A0 = g0a.x * px + g0a.y * py + g0b.x * pz + g0b.y;
                    double A1 = g1a.x * px + g1a.y * py + g1b.x * pz + g1b.y;
                    double A2 = g2a.x * px + g2a.y * py + g2b.x * pz + g2b.y;
                    o[j][i]   = A0 - sa[j] * A1 + sb[j] * A2 + s[j] * c1;
                }
            }
        } else { // interval boundary inside the pair: rare
#pragma unroll
            for (int j = 0; j < 2; j++)
                row_apply(rows + (size_t) idx[j] * kRowSize, tt[j], (double) p[j].x, (double) p[j].y, (double) p[j].z, o[j]);
        }
        // index decimation (lidar_map.cc:249-256): one multiply-high per pair instead of two divisions per point
        int q0 = magic ? (int) __umulhi((unsigned) k0, magic) : k0 / filter_num;
        int r0 = k0 - q0 * filter_num;
        float4 r[2];
#pragma unroll
        for (int j = 0; j < 2; j++) r[j] = make_float4((float) o[j][0], (float) o[j][1], (float) o[j][2], p[j].w);
        if (wide && k0 + 1 < n) {
            F8 v;
            v.a = r[0]; v.b = r[1];
            st256(out_full + k0, v);                                   // the pair fills its 32-byte sector in one store
        } else {
            out_full[k0] = r[0];
            if (k0 + 1 < n) out_full[k0 + 1] = r[1];
        }
        if (out_dec != nullptr) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (k0 + j < n && r0 == 0) out_dec[q0] = r[j];
                if (++r0 == filter_num) { // the next point starts the next decimation group
                    r0 = 0;
                    q0++;
                }
            }
        }
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
project_kernel(Pose12 T, const float4 *__restrict__ src, float4 *__restrict__ dst, int n, const int *__restrict__ n_dev,
               int *__restrict__ report, const int *__restrict__ counters) {
    if (n_dev) n = min(n, *n_dev);
    if (report && blockIdx.x == 0 && threadIdx.x == 0) { // zero-copy result read-back (pinned host memory)
        report[0] = counters ? counters[0] : 0;
        report[1] = n;
    }
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

void launch_frame_prologue(cudaStream_t s, const FrameBatch &b) {
    int wmax = 0;
    for (int i = 0; i < b.count; i++) wmax = b.it[i].w > wmax ? b.it[i].w : wmax;
    if (b.count <= 0 || wmax <= 0) return;
    ins_prep_kernel<<<dim3((wmax + 63) / 64, b.count), 64, 0, s>>>(b);
}

void launch_frame_undistort(cudaStream_t s, const FrameBatch &b) {
    int nmax = 0;
    for (int i = 0; i < b.count; i++) nmax = b.it[i].n > nmax ? b.it[i].n : nmax;
    if (b.count <= 0 || nmax <= 0) return;
    // eight 128-thread CTAs are resident per SM (64 registers): the grid is capped at exactly one resident wave and the
    // threads stride over the pairs (measured at 4 M points: 3.06 TB/s, against 2.83-2.98 TB/s with 2-16 waves)
    dim3 grid(grid_for((nmax + 1) / 2, 128, 8), b.count);
    if (b.it[0].aos) undistort_kernel<true><<<grid, 128, 0, s>>>(b);
    else undistort_kernel<false><<<grid, 128, 0, s>>>(b);
}

void launch_copy_decimate(cudaStream_t s, const float4 *xyzi, int n, int filter_num, float4 *out_full, float4 *out_dec) {
    if (n <= 0) return;
    copy_decimate_kernel<<<grid_for(n, 256), 256, 0, s>>>(xyzi, n, filter_num, out_full, out_dec);
}

void launch_project(cudaStream_t s, Pose12 T, const float4 *src, float4 *dst, int n, const int *n_dev, int *report,
                    const int *counters) {
    if (n <= 0 && !report) return;
    project_kernel<<<grid_for(n, 256), 256, 0, s>>>(T, src, dst, n, n_dev, report, counters);
}

void launch_project_batch(cudaStream_t s, const ProjectBatch &b) {
    int nmax = 0;
    for (int i = 0; i < b.count; i++) nmax = b.n[i] > nmax ? b.n[i] : nmax;
    if (b.count <= 0 || nmax <= 0) return;
    dim3 grid(grid_for(nmax, 256), b.count);
    project_batch_kernel<<<grid, 256, 0, s>>>(b);
}

} // namespace fflins
