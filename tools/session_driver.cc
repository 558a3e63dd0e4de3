// session_driver.cc — the LiDAR call sequence of LINS::runFusion (ff_lins/ff_lins/ff_lins.cc:236-283, 349-367) and of
// Optimizer::optimization / updateParametersFromOptimizer (ff_lins/ff_lins/optimizer.cc:85-185, 203-218) as a NATIVE loop over
// the C-ABI of include/fflins.h — the same sequence as ff-lins_b200/fflins/pipeline.py (Session), which stays the readable
// statement of it and the one the tests drive. The reference's caller is C++ (one fusion thread); timing the path through
// ~40 ctypes calls per keyframe measured the interpreter as much as the library (≈2.5 us per call), so bench.py hands the
// steady-state loop to this driver. Nothing here computes: every call below is an entry point of libfflins_b200.so.
//
//   per frame      fflins_frame_process_dev / _aos (out = NULL: queued)
//   per keyframe   fflins_keyframe_merge (out = NULL) -> fflins_frame_release x (K - 1) -> fflins_keyframe_add ->
//                  fflins_frame_set_motion -> fflins_associate -> n0 x fflins_window_eval -> fflins_window_chi2 ->
//                  n1 x fflins_window_eval -> fflins_reproject_window -> fflins_keyframe_remove_oldest (window full)
//   drv_set_solve(1): the n0 + chi2 + n1 sequence is ONE fflins_window_solve (the Levenberg-Marquardt loop on the device)
#include "../include/fflins.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>

extern "C" {

typedef struct drv_frame {
    const void *pts;       // mode 0: device float4 xyzi; mode 1: host 32-byte PointXYZIT records (page-locked)
    const double *time;    // mode 0: device per-point times; mode 1: unused
    const double *ins_t, *ins_p, *ins_q;   // host INS window
    int32_t n, w;
    double stamp, td;
    double v[3], omega[3];
    double pose7[7];       // IMU pose block of the frame (solver state when it becomes a keyframe)
} drv_frame;

struct KeyState {
    double pose7[7];
    fflins_pose pose;
};

typedef struct drv_session {
    fflins_ctx *ctx;
    fflins_pose pose_bl;
    double ext7[7];
    int K, n0, n1, window;
    uint32_t flags;
    double chi2_threshold;
    uint64_t next_id;
    long long cursor;                      // frames consumed so far (index into the caller's frame list, modulo its length)
    std::vector<uint64_t> nonkeys;
    std::deque<uint64_t> keys;
    std::unordered_map<uint64_t, KeyState> key;
    fflins_assoc_stats assoc;
    bool have_assoc;
    int32_t chi2_outliers[16];
    double last_td;
    int last_pairs;                        // pairs of the last window evaluation (the oldest keyframe may have left since)
    std::vector<double> H, b, cost;
    int32_t n_res[16];
    int use_solve;                         // 1: the keyframe's solve runs on the device (fflins_window_solve)
    double prior_w[4];                     // weights of the diagonal prior: pose t, pose theta, td, (spare)
    std::vector<double> H0, b0, x0;
    fflins_solve_report report;
    int prefetch;                          // host-buffer mode announces frames of the next interval one interval ahead (1: newest, 2: all)
    const struct drv_frame *run_frames;    // the frame list, its length and the mode of the drv_run in progress
    int run_n, run_mode;
    char err[256];
} drv_session;

drv_session *drv_create(fflins_ctx *ctx, const fflins_pose *pose_bl, const double *ext7, int frames_per_key, int n_eval0, int n_eval1,
                        uint32_t flags, int window, double chi2_threshold) {
    drv_session *s = new drv_session();
    s->ctx     = ctx;
    s->pose_bl = *pose_bl;
    std::memcpy(s->ext7, ext7, sizeof(s->ext7));
    s->K = frames_per_key;
    s->n0 = n_eval0;
    s->n1 = n_eval1;
    s->window = window;
    s->flags  = flags;
    s->chi2_threshold = chi2_threshold;
    s->next_id = 0;
    s->cursor  = 0;
    s->have_assoc = false;
    s->last_td = 0.0;
    s->last_pairs = 0;
    s->H.resize(361 * 16);
    s->b.resize(19 * 16);
    s->cost.resize(2 * 16);
    s->err[0] = 0;
    s->use_solve = 0;
    s->prefetch = 0;
    s->run_frames = nullptr;
    s->run_n = s->run_mode = 0;
    s->prior_w[0] = 1.0 / (0.05 * 0.05);
    s->prior_w[1] = 1.0 / (0.01 * 0.01);
    s->prior_w[2] = 1.0 / (1e-3 * 1e-3);
    s->prior_w[3] = 0;
    std::memset(&s->report, 0, sizeof(s->report));
    return s;
}
void drv_set_solve(drv_session *s, int on) { s->use_solve = on; }
// host-buffer mode: announce the newest frame of the NEXT keyframe interval (fflins_host_prefetch) as soon as the current
// interval's copies are queued - what a sensor callback does that hands frames over while the fusion thread is solving
void drv_set_prefetch(drv_session *s, int on) { s->prefetch = on; }
int drv_last_solve(const drv_session *s, int32_t *iters2, int32_t *evals2) {
    iters2[0] = s->report.iterations[0];
    iters2[1] = s->report.iterations[1];
    evals2[0] = s->report.evaluations[0];
    evals2[1] = s->report.evaluations[1];
    return s->report.status;
}
void drv_destroy(drv_session *s) { delete s; }
const char *drv_error(const drv_session *s) { return s->err; }

// DRV_TRACE=1: wall-clock time spent inside each entry point (printed by drv_trace_dump)
static const bool g_trace = std::getenv("DRV_TRACE") != nullptr;
struct TraceRow {
    const char *name;
    double us;
    long long calls;
};
static TraceRow g_rows[32];
static int g_nrows = 0;
static inline void trace_add(const char *name, double us) {
    for (int i = 0; i < g_nrows; i++)
        if (g_rows[i].name == name) {
            g_rows[i].us += us;
            g_rows[i].calls++;
            return;
        }
    if (g_nrows < 32) g_rows[g_nrows++] = TraceRow{name, us, 1};
}
void drv_trace_dump(int reset) {
    for (int i = 0; i < g_nrows; i++)
        std::fprintf(stderr, "drv_trace %-64.64s calls=%8lld avg=%9.2f us total=%10.2f ms\n", g_rows[i].name, g_rows[i].calls,
                     g_rows[i].us / (double) g_rows[i].calls, g_rows[i].us * 1e-3);
    if (reset) g_nrows = 0;
}

#define DRV_CK(call)                                                                                       \
    do {                                                                                                   \
        std::chrono::steady_clock::time_point t0_;                                                         \
        if (g_trace) t0_ = std::chrono::steady_clock::now();                                               \
        int rc_ = (call);                                                                                  \
        if (g_trace) trace_add(#call, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0_).count()); \
        if (rc_ != FFLINS_OK) {                                                                            \
            std::snprintf(s->err, sizeof(s->err), "%s -> %d: %s", #call, rc_, fflins_last_error(s->ctx)); \
            return rc_;                                                                                    \
        }                                                                                                  \
    } while (0)

static int window_args(drv_session *s, uint64_t *ids, double *pose_refs, const double **pose_cur) {
    const int n = (int) s->keys.size() - 1;
    for (int i = 0; i < n; i++) {
        ids[i] = s->keys[i];
        std::memcpy(pose_refs + 7 * i, s->key[s->keys[i]].pose7, 7 * sizeof(double));
    }
    *pose_cur = s->key[s->keys.back()].pose7;
    return n;
}

static int keyframe_work(drv_session *s, uint64_t fid, const drv_frame &fr) {
    DRV_CK(fflins_keyframe_merge(s->ctx, fid, s->nonkeys.data(), (int32_t) s->nonkeys.size(), nullptr));
    if (s->prefetch && s->run_mode == 1 && s->run_frames) { // (the cursor already points behind this keyframe)
        // prefetch = 1: the newest frame of the next interval (the one its association waits for); 2: all its frames,
        // newest first
        const int ahead = s->prefetch >= 2 ? s->K : 1;
        for (int a = 0; a < ahead; a++) {
            const drv_frame &nx = s->run_frames[(s->cursor + s->K - 1 - a) % s->run_n];
            DRV_CK(fflins_host_prefetch(s->ctx, nx.pts, nx.n));
        }
    }
    for (uint64_t nk : s->nonkeys) DRV_CK(fflins_frame_release(s->ctx, nk));
    s->nonkeys.clear();
    DRV_CK(fflins_keyframe_add(s->ctx, fid));
    DRV_CK(fflins_frame_set_motion(s->ctx, fid, fr.v, fr.omega, fr.td));
    KeyState &ks = s->key[fid];
    std::memcpy(ks.pose7, fr.pose7, sizeof(ks.pose7));
    DRV_CK(fflins_frame_get_pose(s->ctx, fid, &ks.pose));   // evaluated on the host at enqueue time: no synchronisation
    s->keys.push_back(fid);
    if (s->keys.size() >= 2) {
        DRV_CK(fflins_associate(s->ctx, &s->assoc));
        s->have_assoc = true;
        uint64_t ids[16];
        double pose_refs[7 * 16];
        const double *pose_cur;
        const int n = window_args(s, ids, pose_refs, &pose_cur);
        if (s->use_solve) {
            // the same schedule as ONE call: n0 LM iterations, the chi2 cull, n1 LM iterations on the device, every pose of
            // the window, the extrinsic and td free, tied to where they start by a diagonal prior (what the inertial
            // factors and the marginalization prior do in the full problem). Tolerances off: the iteration limits are
            // the work, as in the round-trip sequence below. The solution is dropped (the sequence replays the same poses).
            const int D = 6 * n + 13, NX = 7 * n + 15;
            s->H0.assign((size_t) D * D, 0.0);
            s->b0.assign(D, 0.0);
            s->x0.resize(NX);
            for (int i = 0; i < D; i++) s->H0[(size_t) i * D + i] = i == D - 1 ? s->prior_w[2] : ((i % 6) < 3 ? s->prior_w[0] : s->prior_w[1]);
            double refs[7 * 16], cur7[7], ext7[7], td = fr.td;
            std::memcpy(refs, pose_refs, sizeof(double) * 7 * n);
            std::memcpy(cur7, pose_cur, sizeof(cur7));
            std::memcpy(ext7, s->ext7, sizeof(ext7));
            std::memcpy(s->x0.data(), refs, sizeof(double) * 7 * n);
            std::memcpy(s->x0.data() + 7 * n, cur7, sizeof(cur7));
            std::memcpy(s->x0.data() + 7 * n + 7, ext7, sizeof(ext7));
            s->x0[7 * n + 14] = td;
            fflins_solve_options opt;
            fflins_solve_options_default(&opt);
            opt.max_iterations[0]   = s->n0;
            opt.max_iterations[1]   = s->n1;
            opt.chi2_threshold      = s->chi2_threshold;
            opt.function_tolerance  = 0;
            opt.gradient_tolerance  = 0;
            opt.parameter_tolerance = 0;
            opt.flags               = s->flags;
            fflins_solve_prior prior{s->H0.data(), s->b0.data(), s->x0.data(), 0.0};
            DRV_CK(fflins_window_solve(s->ctx, n, ids, refs, cur7, ext7, &td, &opt, &prior, &s->report));
            for (int i = 0; i < n; i++) s->n_res[i] = s->report.n_res[i];
        } else {
            for (int it = 0; it < s->n0; it++)
                DRV_CK(fflins_window_eval(s->ctx, n, ids, pose_refs, pose_cur, s->ext7, fr.td, s->flags, s->H.data(), s->b.data(), s->cost.data(), s->n_res));
            DRV_CK(fflins_window_chi2(s->ctx, n, ids, pose_refs, pose_cur, s->ext7, fr.td, s->chi2_threshold, s->chi2_outliers));
            for (int it = 0; it < s->n1; it++)
                DRV_CK(fflins_window_eval(s->ctx, n, ids, pose_refs, pose_cur, s->ext7, fr.td, s->flags, s->H.data(), s->b.data(), s->cost.data(), s->n_res));
        }
        s->last_td = fr.td;
        s->last_pairs = n;
        // updateParametersFromOptimizer: re-project every keyframe of the window (optimizer.cc:203-218)
        uint64_t all[16];
        fflins_pose poses[16];
        const int m = (int) s->keys.size();
        for (int i = 0; i < m; i++) {
            all[i]   = s->keys[i];
            poses[i] = s->key[s->keys[i]].pose;
        }
        DRV_CK(fflins_reproject_window(s->ctx, m, all, poses));
    }
    if ((int) s->keys.size() > s->window) {
        uint64_t rid = 0;
        DRV_CK(fflins_keyframe_remove_oldest(s->ctx, &rid));
        s->key.erase(rid);
        s->keys.pop_front();
    }
    return FFLINS_OK;
}

// `steps` keyframe intervals (K frames each), cycling over the caller's frames. mode 0: points resident in device memory
// (fflins_frame_process_dev), mode 1: raw 32-byte records in page-locked host memory (fflins_frame_process_aos).
int drv_run(drv_session *s, int mode, const drv_frame *frames, int n_frames, int steps) {
    s->run_frames = frames;
    s->run_n      = n_frames;
    s->run_mode   = mode;
    for (int st = 0; st < steps; st++) {
        for (int k = 0; k < s->K; k++) {
            const drv_frame &fr = frames[s->cursor % n_frames];
            s->cursor++;
            const uint64_t fid = s->next_id++;
            if (mode == 0)
                DRV_CK(fflins_frame_process_dev(s->ctx, fid, (const float *) fr.pts, fr.time, fr.n, fr.ins_t, fr.ins_p, fr.ins_q, fr.w,
                                                &s->pose_bl, fr.stamp, fr.td, nullptr));
            else
                DRV_CK(fflins_frame_process_aos(s->ctx, fid, fr.pts, fr.n, fr.ins_t, fr.ins_p, fr.ins_q, fr.w, &s->pose_bl, fr.stamp, fr.td,
                                                nullptr));
            if ((int) s->nonkeys.size() + 1 >= s->K) {
                int rc = keyframe_work(s, fid, fr);
                if (rc != FFLINS_OK) return rc;
            } else {
                s->nonkeys.push_back(fid);
            }
        }
    }
    return FFLINS_OK;
}

// what bench.py reports about the last keyframe
int drv_last_assoc(const drv_session *s, int32_t *n_refs, int32_t *n_queries) {
    if (!s->have_assoc) return 1;
    *n_refs    = s->assoc.n_refs;
    *n_queries = s->assoc.n_queries;
    return 0;
}
// residual counts of the last window evaluation, one per (ref, cur) pair; returns the pair count
int drv_last_nres(const drv_session *s, int32_t *out16) {
    const int n = s->last_pairs;
    for (int i = 0; i < n && i < 16; i++) out16[i] = s->n_res[i];
    return n;
}
// blocks of the last window evaluation (H 361, b 19, cost 2 per pair); returns the pair count
int drv_last_blocks(const drv_session *s, double *H, double *b, double *cost) {
    const int n = s->last_pairs;
    std::memcpy(H, s->H.data(), sizeof(double) * 361 * n);
    std::memcpy(b, s->b.data(), sizeof(double) * 19 * n);
    std::memcpy(cost, s->cost.data(), sizeof(double) * 2 * n);
    return n;
}
int drv_keys(const drv_session *s, uint64_t *ids, int cap) {
    int n = 0;
    for (uint64_t k : s->keys)
        if (n < cap) ids[n++] = k;
    return n;
}
// wall-clock round trip of one solver-facing evaluation of the current window, in microseconds
double drv_eval_roundtrip_us(drv_session *s, int warm, int reps) {
    if (s->keys.size() < 2) return -1.0;
    uint64_t ids[16];
    double pose_refs[7 * 16];
    const double *pose_cur;
    const int n = window_args(s, ids, pose_refs, &pose_cur);
    for (int i = 0; i < warm; i++)
        if (fflins_window_eval(s->ctx, n, ids, pose_refs, pose_cur, s->ext7, s->last_td, s->flags, s->H.data(), s->b.data(), s->cost.data(), s->n_res) != FFLINS_OK)
            return -1.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++)
        if (fflins_window_eval(s->ctx, n, ids, pose_refs, pose_cur, s->ext7, s->last_td, s->flags, s->H.data(), s->b.data(), s->cost.data(), s->n_res) != FFLINS_OK)
            return -1.0;
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

} // extern "C"
