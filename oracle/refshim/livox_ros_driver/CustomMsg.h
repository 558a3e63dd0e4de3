/* Stand-in for <livox_ros_driver/CustomMsg.h>. */
#ifndef FFLINS_ORACLE_REFSHIM_LIVOX_CUSTOMMSG_H
#define FFLINS_ORACLE_REFSHIM_LIVOX_CUSTOMMSG_H
#include "CustomPoint.h"
#include <sensor_msgs/PointCloud2.h>
namespace livox_ros_driver {
struct CustomMsg {
    std_msgs::Header header;
    std::vector<CustomPoint> points;
};
typedef std::shared_ptr<const CustomMsg> CustomMsgConstPtr;
} // namespace livox_ros_driver
#endif
