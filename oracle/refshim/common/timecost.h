/* Stand-in for the reference's common/timecost.h (abseil clock): only feeds a statistics figure. */
#ifndef FFLINS_ORACLE_REFSHIM_TIMECOST_H
#define FFLINS_ORACLE_REFSHIM_TIMECOST_H
#include <chrono>
class TimeCost {
    std::chrono::steady_clock::time_point t0_ = std::chrono::steady_clock::now();
public:
    double costInMillisecond() const {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count();
    }
};
#endif
