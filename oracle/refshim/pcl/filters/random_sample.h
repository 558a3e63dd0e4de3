/* Stand-in for <pcl/filters/random_sample.h>: only saveMapPointCloud (off the hot path) uses it. */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_RANDOM_SAMPLE_H
#define FFLINS_ORACLE_REFSHIM_PCL_RANDOM_SAMPLE_H
#include "../point_cloud.h"
namespace pcl {
template <class P> class RandomSample {
public:
    void setSample(unsigned) {}
    void setInputCloud(const typename PointCloud<P>::Ptr &) {}
    void filter(PointCloud<P> &out) { out.points.clear(); }
};
} // namespace pcl
#endif
