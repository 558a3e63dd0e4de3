/* Stand-in for the reference's common/logging.h (glog + abseil are not installed): the log sinks swallow their
 * arguments. Found before the reference's header because -Ioracle/refshim precedes -I$(REF)/ff_lins. */
#ifndef FFLINS_ORACLE_REFSHIM_LOGGING_H
#define FFLINS_ORACLE_REFSHIM_LOGGING_H
#include <string>
struct FflinsNullLog {
    template <class T> FflinsNullLog &operator<<(const T &) { return *this; }
};
#define LOGI (FflinsNullLog())
#define LOGW (FflinsNullLog())
#define LOGE (FflinsNullLog())
#define LOGF (FflinsNullLog())
namespace absl {
template <class... A> std::string StrFormat(const char *, const A &...) { return std::string(); }
} // namespace absl
class Logging {
public:
    static std::string doubleData(double v) { return std::to_string(v); }
};
#endif
