/* Stand-in for <livox_ros_driver/CustomPoint.h>: the message's fields (livox_ros_driver CustomPoint.msg). */
#ifndef FFLINS_ORACLE_REFSHIM_LIVOX_CUSTOMPOINT_H
#define FFLINS_ORACLE_REFSHIM_LIVOX_CUSTOMPOINT_H
#include <cstdint>
namespace livox_ros_driver {
struct CustomPoint {
    uint32_t offset_time;
    float x, y, z;
    uint8_t reflectivity, tag, line;
};
} // namespace livox_ros_driver
#endif
