/* Stand-in for <pcl/filters/voxel_grid.h> (PCL is not installed): the VoxelGrid<PointXYZI> surface LidarMap uses
 * (setLeafSize, setInputCloud, filter), following PCL 1.12's voxel_grid.hpp as SURVEY.md Appendix A.1 restates it:
 * bounding box in fp32, key = (ijk - min_b) . (1, dx, dx dy) with ijk = floor(coordinate * inverse_leaf) in fp32, points
 * ordered by key, one centroid per key accumulated in fp32 over all fields then divided by the count, output in key
 * order; "leaf too small" (index overflow) returns the input. PCL sorts with an unstable sort, so its order of summation
 * inside a voxel is implementation-defined; this stand-in sums in ascending input index (std::stable_sort).
 * This is a SECOND restatement of PCL by this repository, not PCL: it lets the reference's LidarMap run end to end,
 * it does not pin the voxel filter. Test infrastructure only. */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_VOXEL_GRID_H
#define FFLINS_ORACLE_REFSHIM_PCL_VOXEL_GRID_H
#include "../point_cloud.h"
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
namespace pcl {
template <class P> class VoxelGrid {
    float inv_[3] = {0, 0, 0};
    typename PointCloud<P>::Ptr in_;
public:
    void setLeafSize(float lx, float ly, float lz) { inv_[0] = 1.0f / lx, inv_[1] = 1.0f / ly, inv_[2] = 1.0f / lz; }
    void setInputCloud(const typename PointCloud<P>::Ptr &c) { in_ = c; }
    void filter(PointCloud<P> &out) {
        PointCloud<P> res;          /* the output may alias the input's owner (mergeNonKeyframes) */
        const std::vector<P> &pts = in_->points;
        const size_t n = pts.size();
        if (n == 0) { out.points.clear(); return; }
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (const P &p : pts) {
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            for (int a = 0; a < 3; a++) mn[a] = std::min(mn[a], p.data[a]), mx[a] = std::max(mx[a], p.data[a]);
        }
        std::int64_t d[3];
        int minb[3], maxb[3];
        for (int a = 0; a < 3; a++) {
            d[a]    = static_cast<std::int64_t>((mx[a] - mn[a]) * inv_[a]) + 1;
            minb[a] = static_cast<int>(std::floor(mn[a] * inv_[a]));
            maxb[a] = static_cast<int>(std::floor(mx[a] * inv_[a]));
        }
        if (d[0] * d[1] * d[2] > static_cast<std::int64_t>(INT_MAX)) { res.points = pts; out = res; return; }
        const int div[3] = {maxb[0] - minb[0] + 1, maxb[1] - minb[1] + 1, maxb[2] - minb[2] + 1};
        const int mul[3] = {1, div[0], div[0] * div[1]};
        std::vector<std::pair<unsigned, unsigned>> idx;
        idx.reserve(n);
        for (size_t i = 0; i < n; i++) {
            const P &p = pts[i];
            if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
            int k = 0;
            for (int a = 0; a < 3; a++)
                k += static_cast<int>(std::floor(p.data[a] * inv_[a]) - static_cast<float>(minb[a])) * mul[a];
            idx.emplace_back(static_cast<unsigned>(k), static_cast<unsigned>(i));
        }
        std::stable_sort(idx.begin(), idx.end(),
                         [](const std::pair<unsigned, unsigned> &a, const std::pair<unsigned, unsigned> &b) { return a.first < b.first; });
        for (size_t s = 0; s < idx.size();) {
            size_t e = s + 1;
            while (e < idx.size() && idx[e].first == idx[s].first) e++;
            float acc[4] = {0, 0, 0, 0};
            for (size_t j = s; j < e; j++) {
                const P &p = pts[idx[j].second];
                acc[0] += p.x, acc[1] += p.y, acc[2] += p.z, acc[3] += p.intensity;
            }
            const float cnt = static_cast<float>(e - s);
            P c;
            c.x = acc[0] / cnt, c.y = acc[1] / cnt, c.z = acc[2] / cnt, c.intensity = acc[3] / cnt;
            res.points.push_back(c);
            s = e;
        }
        out = res;
    }
};
} // namespace pcl
#endif
