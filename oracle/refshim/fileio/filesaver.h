/* Stand-in for the reference's fileio/filesaver.h (its base classes pull in abseil): never used with
 * is_save_pointcloud = false. */
#ifndef FFLINS_ORACLE_REFSHIM_FILESAVER_H
#define FFLINS_ORACLE_REFSHIM_FILESAVER_H
#include <memory>
#include <string>
#include <vector>
class FileSaver {
public:
    typedef std::shared_ptr<FileSaver> Ptr;
    enum { TEXT = 0, BINARY = 1 };
    FileSaver(const std::string &, int, int = TEXT) {}
    void dump(const std::vector<double> &) {}
};
#endif
