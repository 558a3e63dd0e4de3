/* Stand-in for <pcl_conversions/pcl_conversions.h>: pcl::fromROSMsg fills the typed cloud from the decoded arrays of the
 * stand-in message (the field that carries the per-point time is `time` for Velodyne, `t` for Ouster, `timestamp` for
 * Hesai points: lidar/lidar.h). */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_CONVERSIONS_H
#define FFLINS_ORACLE_REFSHIM_PCL_CONVERSIONS_H
#include "lidar/lidar.h"
#include <cfloat>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
inline void fflins_set_time(VelodynePoint &p, double t) { p.time = (float) t; }
inline void fflins_set_time(OusterPoint &p, double t) { p.t = (std::uint32_t) t; }
inline void fflins_set_time(HesaiPoint &p, double t) { p.timestamp = t; }
template <class P> void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<P> &cloud) {
    cloud.points.resize(msg.x.size());
    for (size_t k = 0; k < msg.x.size(); k++) {
        P &p        = cloud.points[k];
        p.x         = msg.x[k];
        p.y         = msg.y[k];
        p.z         = msg.z[k];
        p.intensity = msg.intensity[k];
        fflins_set_time(p, msg.time[k]);
    }
}
} // namespace pcl
#endif
