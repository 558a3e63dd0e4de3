/* C wrapper around the REFERENCE's own ikd-Tree (compiled from /root/reference, see
 * Makefile `ref`): exactly the calls FF-LINS makes — setDownsampleParam + build
 * (lidar_map.cc:113-114) and nearestSearch(p, 5, ..., 5.0) (lidar_map.cc:368).
 * Test infrastructure / CPU reference baseline only. */
#include "ikd_Tree.h"

#include <cstring>
#include <memory>

using Tree = ikd_Tree::KD_TREE<pcl::PointXYZI>;

extern "C" {

void *ikdref_build(const float *xyzi, int m, double downsample) {
    auto *tree = new Tree();
    tree->setDownsampleParam(downsample);
    Tree::PointVector pts(m);
    for (int i = 0; i < m; i++) {
        pts[i].x         = xyzi[4 * i];
        pts[i].y         = xyzi[4 * i + 1];
        pts[i].z         = xyzi[4 * i + 2];
        pts[i].intensity = xyzi[4 * i + 3];
    }
    tree->build(pts);
    return tree;
}

/* out_xyzi: k x 4 floats, out_d2: k floats; returns the number found. */
int ikdref_search(void *handle, const float *query_xyz, int k, double max_dist, float *out_xyzi, float *out_d2) {
    auto *tree = static_cast<Tree *>(handle);
    pcl::PointXYZI p;
    p.x = query_xyz[0];
    p.y = query_xyz[1];
    p.z = query_xyz[2];
    Tree::PointVector nearest;
    std::vector<float> d2;
    tree->nearestSearch(p, k, nearest, d2, max_dist);
    int n = (int) nearest.size();
    for (int i = 0; i < n; i++) {
        out_xyzi[4 * i]     = nearest[i].x;
        out_xyzi[4 * i + 1] = nearest[i].y;
        out_xyzi[4 * i + 2] = nearest[i].z;
        out_xyzi[4 * i + 3] = nearest[i].intensity;
        out_d2[i]           = d2[i];
    }
    return n;
}

void ikdref_free(void *handle) { delete static_cast<Tree *>(handle); }
}
