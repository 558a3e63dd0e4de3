/* Stand-in for <sensor_msgs/PointCloud2.h> (ROS is not installed): a message header with a stamp, and - instead of the
 * serialised byte blob - the typed point arrays pcl::fromROSMsg would decode (oracle/refshim/pcl_conversions). */
#ifndef FFLINS_ORACLE_REFSHIM_SENSOR_MSGS_POINTCLOUD2_H
#define FFLINS_ORACLE_REFSHIM_SENSOR_MSGS_POINTCLOUD2_H
#include <cstdint>
#include <memory>
#include <vector>
namespace ros {
struct Time {
    uint32_t sec = 0, nsec = 0;
    uint64_t toNSec() const { return (uint64_t) sec * 1000000000ull + (uint64_t) nsec; }
    double toSec() const { return (double) sec + 1e-9 * (double) nsec; }   /* ros::TimeBase::toSec */
};
} // namespace ros
namespace std_msgs {
struct Header {
    ros::Time stamp;
};
} // namespace std_msgs
namespace sensor_msgs {
struct PointCloud2 {
    std_msgs::Header header;
    /* decoded fields, one entry per point: x, y, z, intensity, and the sensor's time field in its own unit */
    std::vector<float> x, y, z, intensity;
    std::vector<double> time;
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
} // namespace sensor_msgs
#endif
