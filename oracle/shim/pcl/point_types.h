/* Stand-in for <pcl/point_types.h> (PCL is not installed in this image): only the three
 * point structs the reference's vendored ikd-Tree instantiates (ikd_Tree.cc:1507-1510),
 * with PCL's field names and 16-byte-aligned layouts. Test infrastructure only. */
#ifndef FFLINS_ORACLE_PCL_POINT_TYPES_SHIM_H
#define FFLINS_ORACLE_PCL_POINT_TYPES_SHIM_H
/* headers the real PCL header pulls in transitively and the kd-tree relies on */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unistd.h>
#include <vector>
/* Eigen::aligned_allocator is only an allocator (no arithmetic): std::allocator honours
 * alignas(16) since C++17. */
namespace Eigen {
template <class T> using aligned_allocator = std::allocator<T>;
}
namespace pcl {
struct alignas(16) PointXYZ {
    float x = 0, y = 0, z = 0, pad_ = 1.f;
};
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, pad_ = 1.f;
    float intensity = 0, pad1_[3] = {0, 0, 0};
};
struct alignas(16) PointXYZINormal {
    float x = 0, y = 0, z = 0, pad_ = 1.f;
    float normal_x = 0, normal_y = 0, normal_z = 0, pad2_ = 0;
    float intensity = 0, curvature = 0, pad3_[2] = {0, 0};
};
} // namespace pcl
#endif
