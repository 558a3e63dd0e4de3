// fflins_types.h — layout-compatible stand-ins for the third-party types that appear in the
// reference's LiDAR interfaces, used ONLY when the real headers are absent (this image has no
// Eigen / PCL / Ceres). With the real libraries installed, define FFLINS_SHIM_USE_EIGEN /
// FFLINS_SHIM_USE_PCL / FFLINS_SHIM_USE_CERES before including lidar_shim.h and the shim binds
// to Eigen::Vector3d / Matrix3d, pcl::PointXYZI / PointCloud and ceres::SizedCostFunction instead.
//
// Mirrors: ff_lins/common/types.h:43-46 (Pose), ff_lins/lidar/lidar.h:30-38,88-95 (PointXYZIT,
// PointType, PointCloud, PointVector).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#if defined(FFLINS_SHIM_USE_EIGEN)
#include <Eigen/Geometry>
namespace fflins_shim {
using Vector3d = Eigen::Vector3d;
using Matrix3d = Eigen::Matrix3d;
} // namespace fflins_shim
#else
namespace fflins_shim {
struct Vector3d {
    double v[3] = {0, 0, 0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    const double *data() const { return v; }
    double *data() { return v; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    static Vector3d Zero() { return Vector3d(); }
};
// column-major storage, like Eigen::Matrix3d
struct Matrix3d {
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double &operator()(int r, int c) { return m[3 * c + r]; }
    double operator()(int r, int c) const { return m[3 * c + r]; }
    const double *data() const { return m; }
    double *data() { return m; }
    static Matrix3d Identity() { return Matrix3d(); }
};
} // namespace fflins_shim
#endif

using fflins_shim::Matrix3d;
using fflins_shim::Vector3d;

// ff_lins/common/types.h:43-46
typedef struct Pose {
    Matrix3d R;
    Vector3d t;
} Pose;

#if defined(FFLINS_SHIM_USE_PCL)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
namespace pcl {
// 32 bytes, 16-byte aligned, x y z pad | intensity pad pad pad: pcl::PointXYZI
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, data_pad = 1.f;
    float intensity = 0, pad1[3] = {0, 0, 0};
};
template <class PointT> class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void resize(size_t n) { points.resize(n); width = (uint32_t) n; }
    void reserve(size_t n) { points.reserve(n); }
    void clear() { points.clear(); width = 0; }
    void push_back(const PointT &p) { points.push_back(p); width = (uint32_t) points.size(); }
    PointT &operator[](size_t i) { return points[i]; }
    const PointT &operator[](size_t i) const { return points[i]; }
    PointCloud &operator+=(const PointCloud &o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        width = (uint32_t) points.size();
        return *this;
    }
};
} // namespace pcl
#endif

// ff_lins/lidar/lidar.h:30-38 — 32-byte raw point with per-point time
struct alignas(16) PointXYZIT {
    float x = 0, y = 0, z = 0, data_pad = 1.f;
    float intensity = 0;
    float pad1      = 0;
    double time     = 0;
};
static_assert(sizeof(PointXYZIT) == 32, "PointXYZIT must be 32 bytes (fflins_frame_process_aos)");
static_assert(sizeof(pcl::PointXYZI) == 32, "pcl::PointXYZI must be 32 bytes");

// ff_lins/lidar/lidar.h:88-95
typedef PointXYZIT PointTypeCustom;
typedef pcl::PointCloud<PointTypeCustom> PointCloudCustom;
typedef pcl::PointCloud<PointTypeCustom>::Ptr PointCloudCustomPtr;
typedef pcl::PointXYZI PointType;
typedef pcl::PointCloud<PointType> PointCloud;
typedef pcl::PointCloud<PointType>::Ptr PointCloudPtr;
typedef std::vector<PointType> PointVector;

// the slice of ff_lins/common/types.h:33-41 (IMU) and integration_state.h:35-52 (IntegrationState)
// that the LiDAR path reads: times, positions, quaternions (x,y,z,w)
struct IMU {
    double time = 0;
};
struct Quaterniond4 {
    double c[4] = {0, 0, 0, 1}; // x y z w, Eigen coeffs() order
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double w() const { return c[3]; }
};
struct IntegrationState {
    double time = 0;
    Vector3d p;
    Quaterniond4 q;
    Vector3d v;
};
