/* Stand-in for <pcl/point_cloud.h>: the container surface pointcloud.cc uses (points, size, resize, Ptr). */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_POINT_CLOUD_H
#define FFLINS_ORACLE_REFSHIM_PCL_POINT_CLOUD_H
#include "point_types.h"
/* headers the real PCL / Eigen headers pull in transitively and the reference relies on */
#include <deque>
#include <iostream>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <unordered_map>
namespace pcl {
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    size_t size() const { return points.size(); }
    void resize(size_t n) { points.resize(n); }
    void reserve(size_t n) { points.reserve(n); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); }
    void push_back(const P &p) { points.push_back(p); }
    PointCloud &operator+=(const PointCloud &o) {
        points.insert(points.end(), o.points.begin(), o.points.end());
        return *this;
    }
    P &operator[](size_t i) { return points[i]; }
    const P &operator[](size_t i) const { return points[i]; }
};
} // namespace pcl
#endif
