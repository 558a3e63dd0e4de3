/* Stand-in for <tbb/tbb.h>: parallel_for over a blocked_range runs the body once over the whole range (the bodies
 * are independent per index, so the results do not depend on the partition). */
#ifndef FFLINS_ORACLE_REFSHIM_TBB_H
#define FFLINS_ORACLE_REFSHIM_TBB_H
#include <cstddef>
#include <vector>
namespace tbb {
template <class T> class blocked_range {
    T b_, e_;
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
};
/* marginalization_info.cc:139-140 builds per-thread partial sums when there are more factors than threads; the stand-in
 * reports "unlimited" concurrency so that the serial branch (the same accumulation, in factor order) is the one taken */
namespace info {
inline int default_concurrency() { return 1 << 30; }
} // namespace info
template <class T> using concurrent_vector = std::vector<T>;
/* task_group::run executes the task at once (LidarMap uses it to allocate / release kd-trees off the critical path) */
class task_group {
public:
    template <class F> void run(const F &f) { f(); }
    void wait() {}
};
/* parallel_for: serial by default (the parity builds). With -DFFLINS_REFSHIM_OPENMP -fopenmp (the CPU-baseline build,
 * oracle/_ref/libff_ref_omp.so) the range is cut into chunks that run as OpenMP tasks; a parallel_for reached from
 * inside another one (updateLidarFeaturesInMap -> findFeaturesInMap) adds its chunks to the same team, as TBB's
 * work-stealing scheduler would. The bodies are independent per index, so results do not depend on the partition. */
#ifdef FFLINS_REFSHIM_OPENMP
} // namespace tbb
#include <omp.h>
namespace tbb {
template <class R, class F> void parallel_for(const R &range, const F &body) {
    const auto b = range.begin(), e = range.end();
    if (!(b < e)) return;
    const size_t n = (size_t) (e - b);
    const size_t threads = (size_t) omp_get_max_threads();
    size_t grain = n / (8 * threads);
    if (grain < 16) grain = n < 64 ? 1 : 16;
    const size_t chunks = (n + grain - 1) / grain;
    auto run = [&]() {
#pragma omp taskloop grainsize(1) default(shared)
        for (size_t c = 0; c < chunks; c++) {
            const size_t lo = c * grain, hi = lo + grain < n ? lo + grain : n;
            body(R(b + lo, b + hi));
        }
    };
    if (omp_in_parallel()) {
        run();
    } else {
#pragma omp parallel
#pragma omp single
        run();
    }
}
#else
template <class R, class F> void parallel_for(const R &range, const F &body) { body(range); }
#endif
} // namespace tbb
#endif
