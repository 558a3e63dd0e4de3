/* Stand-in for <tbb/tbb.h>: parallel_for over a blocked_range runs the body once over the whole range (the bodies
 * are independent per index, so the results do not depend on the partition). */
#ifndef FFLINS_ORACLE_REFSHIM_TBB_H
#define FFLINS_ORACLE_REFSHIM_TBB_H
#include <cstddef>
namespace tbb {
template <class T> class blocked_range {
    T b_, e_;
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
};
/* task_group::run executes the task at once (LidarMap uses it to allocate / release kd-trees off the critical path) */
class task_group {
public:
    template <class F> void run(const F &f) { f(); }
    void wait() {}
};
template <class R, class F> void parallel_for(const R &range, const F &body) { body(range); }
} // namespace tbb
#endif
