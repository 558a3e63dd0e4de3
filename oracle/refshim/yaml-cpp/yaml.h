/* Stand-in for <yaml-cpp/yaml.h> (not installed): LidarMap's constructor reads eight scalars by
 * config["section"]["key"].as<T>() (lidar_map.cc:38-56); the values come from a table the test wrapper fills
 * (ffref_yaml_set) instead of a file. */
#ifndef FFLINS_ORACLE_REFSHIM_YAML_H
#define FFLINS_ORACLE_REFSHIM_YAML_H
#include <map>
#include <string>
#include <vector>
namespace YAML {
inline std::map<std::string, double> &fflins_table() {
    static std::map<std::string, double> t;
    return t;
}
class Node {
    std::string path_;
public:
    Node() {}
    explicit Node(std::string p) : path_(std::move(p)) {}
    Node operator[](const char *key) const { return Node(path_.empty() ? std::string(key) : path_ + "." + key); }
    template <class T> T as() const { return static_cast<T>(fflins_table().at(path_)); }
};
inline Node LoadFile(const std::string &) { return Node(); }
} // namespace YAML
#endif
