/* Stand-in for <pcl/io/pcd_io.h>: the writer is never reached (is_save_pointcloud = false). */
#ifndef FFLINS_ORACLE_REFSHIM_PCL_PCD_IO_H
#define FFLINS_ORACLE_REFSHIM_PCL_PCD_IO_H
#include <string>
namespace pcl {
class PCDWriter {
public:
    template <class C> int writeBinaryCompressed(const std::string &, const C &) { return 0; }
};
} // namespace pcl
#endif
