/* Stand-in for <pcl/point_types.h> for the build of the reference's own hot-path sources (oracle/Makefile `ffref`):
 * PCL is not installed. Field names, the 16-byte-aligned layout and getVector3fMap() as PCL documents them
 * (PCL_ADD_POINT4D: float data[4] overlaid with x, y, z; data[3] = 1). Test infrastructure only. */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_POINT_TYPES_H
#define FFLINS_ORACLE_REFSHIM_PCL_POINT_TYPES_H
#include "../mini_eigen.h"
#include <cstdint>
#include <memory>
#include <vector>
#undef EIGEN_ALIGN16
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
#define PCL_ADD_POINT4D                                                                                               \
    union EIGEN_ALIGN16 {                                                                                             \
        float data[4];                                                                                                \
        struct {                                                                                                      \
            float x, y, z;                                                                                            \
        };                                                                                                            \
    };                                                                                                                \
    Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(data); }                        \
    Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(data); }
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fields)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
namespace pcl {
/* the reference's ikd_Tree.cc instantiates its tree for these two as well (ikd_Tree.cc:1507-1510) */
struct EIGEN_ALIGN16 PointXYZ {
    PCL_ADD_POINT4D;
    PointXYZ() { data[0] = data[1] = data[2] = 0.f, data[3] = 1.f; }
};
struct EIGEN_ALIGN16 PointXYZINormal {
    PCL_ADD_POINT4D;
    float normal_x = 0, normal_y = 0, normal_z = 0, pad2_ = 0;
    float intensity = 0, curvature = 0, pad3_[2] = {0, 0};
    PointXYZINormal() { data[0] = data[1] = data[2] = 0.f, data[3] = 1.f; }
};
struct EIGEN_ALIGN16 PointXYZI {
    PCL_ADD_POINT4D;
    union {
        struct {
            float intensity;
        };
        float data_c[4];
    };
    PointXYZI() {
        data[0] = data[1] = data[2] = 0.f;
        data[3]                     = 1.f;
        data_c[0] = data_c[1] = data_c[2] = data_c[3] = 0.f;
    }
};
} // namespace pcl
#endif
