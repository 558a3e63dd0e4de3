// context.cu — the C-ABI of include/fflins.h: device-resident session state (frames, keyframe
// maps + hashes, features) and the stream-ordered launch sequences behind every entry point.
// Host arithmetic here is limited to 3x3 pose algebra the reference also does on the host.
// There is no CPU fallback: a missing CUDA device fails fflins_create.
#include "../../include/fflins.h"
#include "kernels.h"
#include "math_undistort.cuh"

#include <chrono>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

using namespace fflins;

namespace {

struct Cloud {
    float4 *d = nullptr;
    int n     = 0;
    int cap   = 0;
};

struct Frame {
    uint64_t id = 0;
    Cloud c[5];
    fflins_pose pose{};
    double v[3] = {0, 0, 0}, w[3] = {0, 0, 0}, td = 0;
    int n_raw = 0, n_fallback = 0;
    bool is_key = false;
    bool hash_pending = false; // map still being built on the map stream: the hash is inserted at first use
    bool hash_early   = false; // ... unless the map job itself inserted it (table sized from the previous map)
    int pending_slot = -1; // >= 0: counts of the last frame_process still sit in pinned slot (asynchronous mode)
    const int *d_ndown = nullptr;        // device copy of the voxel count of its last chain (batch scratch) ...
    unsigned long long d_ndown_gen = 0;  // ... valid while no later launch has reused the scratch (ctx->scratch_gen)
    bool bg = false;       // its kernel chain runs on the background stream behind its own host-to-device copy (see flush_frames)
    // keyframe hash (K4)
    MapIndex index; // fine / coarse / super hash tables (hcap slots each)
    int hcap                  = 0;
    // features stored on this (ref) frame, indexed by the cur frame's points
    FeatureSoA feat;
    int feat_cap = 0, feat_q = 0;
    int *feat_counts = nullptr;
};

struct Pool {
    std::map<size_t, std::vector<void *>> free_lists;
    std::unordered_map<void *, size_t> sizes;
    size_t total = 0;
    static size_t round(size_t b) {
        size_t r = 256;
        while (r < b) r <<= 1;
        return r;
    }
    void *alloc(size_t bytes) {
        size_t r  = round(bytes);
        auto &lst = free_lists[r];
        if (!lst.empty()) {
            void *p = lst.back();
            lst.pop_back();
            return p;
        }
        void *p = nullptr;
        if (cudaMalloc(&p, r) != cudaSuccess) return nullptr;
        sizes[p] = r;
        total += r;
        return p;
    }
    void release(void *p) {
        if (!p) return;
        free_lists[sizes[p]].push_back(p);
    }
    void destroy() {
        for (auto &kv : sizes) cudaFree(kv.first);
        sizes.clear();
        free_lists.clear();
    }
};

struct ProfRec {
    int name;
    cudaEvent_t a, b;
};

constexpr int kWinRing   = 2 * kFrameBatch;  // pinned INS-window slots
constexpr int kStageRing = 2 * kFrameBatch;  // device staging slots for host-resident points (a batch + the frames announced ahead)
constexpr int kMaxAnnounced = kFrameBatch;   // raw frames announced ahead of their processing (fflins_host_prefetch)
constexpr int kGuardRing = 8;
constexpr unsigned long long kQueued = ~0ull;
constexpr int kItemScratch = 32;             // ints per batch item: [0..3] counters, [16..31] voxel meta

struct QueuedFrame {
    Frame *f;
    FrameItem item;      // scratch pointers are filled in at flush time
    int wslot;           // pinned window slot
    int stage_slot;      // device staging slot (-1: device-resident points)
    // pinned host points: the copies are issued at flush time (newest frame first); n_copy = 0: nothing deferred
    int n_copy;
    const void *h_src[2];
    char *d_dst[2];
    size_t bytes[2];
};

// Pageable caller buffers: cudaMemcpyAsync stages them through the driver's bounce buffer on the calling thread at
// 6-8 GB/s (0.6 ms for a 4.9 MB AT128 frame — the whole cost of the drop-in path, which cannot ask the caller for pinned
// memory). The context copies them into its own page-locked ring with a few helper threads instead and hands the DMA
// engine a pinned source.
struct CopyPool {
    static constexpr int kWorkers = 3;
    std::thread th[kWorkers];
    std::mutex mu;
    std::condition_variable cv, done_cv;
    const char *src = nullptr;
    char *dst = nullptr;
    size_t bytes = 0;
    unsigned long long job = 0;
    int pending = 0;
    bool quit = false, started = false;

    static void slice(int part, int parts, size_t bytes, size_t *lo, size_t *hi) {
        const size_t chunk = ((bytes + parts - 1) / parts + 4095) & ~(size_t) 4095;
        *lo = std::min(bytes, chunk * part);
        *hi = std::min(bytes, chunk * (part + 1));
    }
    void start() {
        if (started) return;
        started = true;
        for (int w = 0; w < kWorkers; w++)
            th[w] = std::thread([this, w] {
                unsigned long long seen = 0;
                std::unique_lock<std::mutex> lk(mu);
                while (true) {
                    cv.wait(lk, [&] { return quit || job != seen; });
                    if (quit) return;
                    seen = job;
                    size_t lo, hi;
                    slice(w + 1, kWorkers + 1, bytes, &lo, &hi);
                    const char *s = src;
                    char *d       = dst;
                    lk.unlock();
                    if (hi > lo) std::memcpy(d + lo, s + lo, hi - lo);
                    lk.lock();
                    if (--pending == 0) done_cv.notify_one();
                }
            });
    }
    void copy(void *d, const void *s, size_t n) {
        if (n < (1u << 20) || !started) {   // small: not worth a wake-up
            std::memcpy(d, s, n);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            src     = (const char *) s;
            dst     = (char *) d;
            bytes   = n;
            pending = kWorkers;
            job++;
        }
        cv.notify_all();
        size_t lo, hi;
        slice(0, kWorkers + 1, n, &lo, &hi);
        std::memcpy((char *) d + lo, (const char *) s + lo, hi - lo);
        std::unique_lock<std::mutex> lk(mu);
        done_cv.wait(lk, [&] { return pending == 0; });
    }
    void stop() {
        if (!started) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        for (auto &t : th)
            if (t.joinable()) t.join();
        started = false;
    }
};

} // namespace

struct fflins_ctx : public Profiler {
    fflins_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::unordered_map<uint64_t, Frame *> frames;
    std::deque<uint64_t> keyframes;
    Pool pool;
    std::string err;
    std::atomic<int64_t> launches{0};
    float cell_scale = 1.0f;

    // staging (device)
    // raw-input staging: a ring of device buffers filled by a dedicated copy stream, so that the H2D copy
    // of frame k+1 overlaps the kernels of frame k (asynchronous mode)
    void *d_stage = nullptr;      // kStageRing x stage_cap bytes
    size_t stage_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t bg_stream = nullptr;   // kernels of the background frames of a split batch (the copy stream only copies)
    int stage_next = 0;
    // raw frames announced ahead of their fflins_frame_process_aos call (fflins_host_prefetch): their copies are already on
    // the copy stream, in staging slot `slot` (guard kQueued until the frame is queued or the announcement is dropped)
    struct Announced {
        const void *host;
        size_t bytes;
        int slot;
        char *dev;
    };
    std::vector<Announced> announced;
    cudaEvent_t ev_copied = nullptr;  // copy stream: the host-to-device copies the compute stream waits for are done
    cudaEvent_t ev_bg_done = nullptr; // background stream: the background part of the last split batch is done
    cudaEvent_t ev_bg_copied = nullptr; // copy stream: the copies of that part have arrived
    cudaEvent_t ev_join_c = nullptr;
    cudaEvent_t ev_tl[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // FFLINS_COPY_TIMELINE=1 (debug)
    bool tl_armed = false;
    cudaEvent_t ev_flush = nullptr;   // compute stream: everything queued before a split batch was flushed
    cudaEvent_t ev_join_a = nullptr, ev_join_b = nullptr; // timer: map / copy stream joins
    bool bg_active = false;           // ... and nobody has waited for it yet
    bool bg_staged = false;           // ... and it contains host-to-device copies of point data
    int merge_reserve = 1;            // frames per keyframe seen so far: every frame's full cloud is allocated with room for them
    bool copy_owes_bg_wait = false;   // a background part ran since the copy stream last waited for one (stage_window slots)
    unsigned long long last_main_seq = 0;
    unsigned long long scratch_gen = 0; // bumped whenever the per-item batch scratch is handed out again
    // page-locked bounce ring for pageable caller buffers (filled by CopyPool, drained by the copy stream)
    CopyPool copy_pool;
    unsigned char *h_bounce = nullptr;
    size_t bounce_cap = 0;
    int bounce_next = 0;
    cudaEvent_t bounce_ev[kStageRing] = {};
    // Frame queue: asynchronous fflins_frame_process* calls only stage their inputs; the per-frame kernel chain
    // is launched for all queued frames at once (flush_frames) by the first call that needs a result.
    std::vector<QueuedFrame> fq;
    // batch guards: batch_ev[seq % kGuardRing] is recorded once batch `seq` has consumed its inputs; the pinned
    // window slots and device staging slots remember the batch that reads them (kQueued = not launched yet)
    cudaEvent_t batch_ev[kGuardRing] = {};
    unsigned long long batch_seq = 0;
    unsigned long long win_guard[kWinRing] = {}, stage_guard[kStageRing] = {};
    int *d_fscratch = nullptr;        // per batch item: 4 counters + M_SIZE voxel meta ints
    int ins_cap = 0;
    double *d_tab = nullptr;          // kFrameBatch x ins_cap x (kRowSize + 1) doubles
    int *d_counters = nullptr;      // [64]
    VoxelWork vox;
    void *vox_base = nullptr;
    // keyframe-map construction (merge projection + map voxel filter) runs on its own stream, concurrently
    // with association and factor evaluation of the same keyframe, which do not depend on the new map
    cudaStream_t map_stream = nullptr;
    VoxelWork vox_map;
    void *vox_map_base = nullptr;
    int *d_map_count = nullptr, *h_map_count = nullptr;
    int last_map_n = 0; // size of the last keyframe map built on the map stream (sizes the next early index)
    cudaEvent_t ev_main_done = nullptr, ev_map_done = nullptr;
    bool map_job_active = false;
    uint64_t map_job_key = 0;
    std::vector<Frame *> deferred_frames; // released while the map job may still read them
    // The launches of a map job (3 since the cooperative voxel filter, ~20 before) are issued by a helper thread (the reference, too, hands its per-keyframe
    // tree work to a tbb::task_group, lidar_map.cc:121-125,137), so the fusion thread goes straight on to the
    // association. The helper only issues CUDA calls with arguments prepared by the caller.
    std::thread map_worker;
    std::mutex map_mu;
    std::condition_variable map_cv;
    std::function<void()> map_task;   // pending job (empty = none)
    bool map_issued = true;           // the helper has issued every launch of the last job
    bool map_quit   = false;
    std::vector<void *> deferred_blocks;
    // factor scratch
    FactorRuntime *frt = nullptr;   // K6 / K7 runtime: setup ring, partial sums, resident evaluation kernel
    double *d_red = nullptr;
    int *d_assoc_counts = nullptr; // [kMaxRefs][2]: PLANE / covered counts of one association launch
    // The resident evaluation kernel is not ordered on the compute stream: it may only be used while nothing the
    // evaluation reads (features, outlier flags, the current keyframe's cloud) is still being produced on a stream.
    // Set by the calls that WAIT for such work (association, launched evaluations), cleared by the calls that queue it.
    bool eval_inputs_settled = false;
    // pinned host scratch
    unsigned char *h_pin = nullptr;
    size_t h_pin_cap = 0;
    // pinned completion flags released by the kernels of the synchronous solver-facing calls
    // (window_eval): the host polls them instead of synchronising the stream
    unsigned long long *h_signal = nullptr; // [kMaxRefs] window_eval flags | -, associate flag | [8..] chi2 counts | [16..] associate counts | [48..] chi2 flags
    unsigned long long signal_seq = 0;
    // asynchronous frames: per-frame count slots (pinned, written by the last kernel of the frame chain)
    int *h_slots = nullptr;       // kSlots x 2 ints
    int next_slot = 0;
    std::vector<uint64_t> pending; // frame ids whose counts are not read back yet
    // ring of pinned INS-window staging areas (the CPU packs window k+1 while copy k may still be queued)
    unsigned char *h_win = nullptr;
    size_t h_win_slot = 0;
    int win_next = 0;

    // profiler
    bool prof_on = false;
    std::vector<std::string> prof_names;
    std::vector<double> prof_ms;
    std::vector<int64_t> prof_cnt;
    std::vector<ProfRec> prof_open;
    std::vector<cudaEvent_t> ev_free;
    int prof_cur = -1;
    cudaEvent_t prof_a = nullptr;
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;

    void begin(const char *name, cudaStream_t s) override {
        if (!prof_on) return;
        int id = -1;
        for (size_t i = 0; i < prof_names.size(); i++)
            if (prof_names[i] == name) id = (int) i;
        if (id < 0) {
            id = (int) prof_names.size();
            prof_names.push_back(name);
            prof_ms.push_back(0);
            prof_cnt.push_back(0);
        }
        prof_cur = id;
        prof_a   = get_event();
        cudaEventRecord(prof_a, s);
    }
    void end(cudaStream_t s) override {
        if (!prof_on || prof_cur < 0) return;
        cudaEvent_t b = get_event();
        cudaEventRecord(b, s);
        prof_open.push_back({prof_cur, prof_a, b});
        prof_cur = -1;
    }
    cudaEvent_t get_event() {
        if (!ev_free.empty()) {
            cudaEvent_t e = ev_free.back();
            ev_free.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void prof_collect() {
        if (map_stream) cudaStreamSynchronize(map_stream);
        cudaStreamSynchronize(stream);
        for (auto &r : prof_open) {
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            prof_ms[r.name] += ms;
            prof_cnt[r.name] += 1;
            ev_free.push_back(r.a);
            ev_free.push_back(r.b);
        }
        prof_open.clear();
    }
    Profiler *prof() { return prof_on ? this : nullptr; }
};

namespace {

constexpr int kSlots = 64;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return FFLINS_E_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

#define REQUIRE(cond, msg)               \
    do {                                 \
        if (!(cond)) {                   \
            ctx->err = msg;              \
            return FFLINS_E_INVALID;     \
        }                                \
    } while (0)

int fail(fflins_ctx *ctx, int code, const std::string &msg) {
    ctx->err = msg;
    return code;
}

Pose12 to12(const fflins_pose &p) {
    Pose12 o;
    std::memcpy(o.R, p.R, sizeof(o.R));
    std::memcpy(o.t, p.t, sizeof(o.t));
    return o;
}

// 3x3 helpers, evaluated ((a0*b0 + a1*b1) + a2*b2) like the oracle; nvcc's host pass targets
// baseline x86-64, which has no FMA, so products and sums round separately.
void mat_tn(const double *A, const double *B, double *C) { // A^T B
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = (A[i] * B[j] + A[3 + i] * B[3 + j]) + A[6 + i] * B[6 + j];
}
void mat_tv(const double *A, const double *v, double *o) { // A^T v
    for (int i = 0; i < 3; i++) o[i] = (A[i] * v[0] + A[3 + i] * v[1]) + A[6 + i] * v[2];
}

Frame *find_frame(fflins_ctx *ctx, uint64_t id) {
    auto it = ctx->frames.find(id);
    return it == ctx->frames.end() ? nullptr : it->second;
}
Frame *get_frame(fflins_ctx *ctx, uint64_t id) {
    Frame *f = find_frame(ctx, id);
    if (!f) {
        f     = new Frame();
        f->id = id;
        for (int i = 0; i < 9; i++) f->pose.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        ctx->frames[id] = f;
    }
    return f;
}

// Read back the counts of every frame processed asynchronously (one synchronisation for all of them).
int flush_frames(fflins_ctx *ctx);

// The compute stream joins the background part of a split batch (flush_frames): later work on it may touch those frames.
int join_bg(fflins_ctx *ctx) {
    if (!ctx->bg_active) return FFLINS_OK;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_bg_done, 0));
    ctx->bg_active = false;
    for (auto &kv : ctx->frames) kv.second->bg = false;
    for (Frame *f : ctx->deferred_frames) f->bg = false;
    return FFLINS_OK;
}

// Read back the counts of every frame processed asynchronously (one synchronisation for all of them).
// all = false: only the frames whose chain ran on the compute stream — the background frames of a split batch stay
// pending, so that the caller (association, factor evaluation of the newest keyframe) does not wait for their copies.
int resolve_pending(fflins_ctx *ctx, bool all = true) {
    {
        int rc = flush_frames(ctx);
        if (rc) return rc;
        if (all && (rc = join_bg(ctx))) return rc;
    }
    if (ctx->pending.empty()) return FFLINS_OK;
    if (!all) { // nothing to adopt while only background frames are pending: no synchronisation either
        bool any = false;
        for (uint64_t id : ctx->pending) {
            Frame *f = find_frame(ctx, id);
            any |= f && f->pending_slot >= 0 && !f->bg;
        }
        if (!any) return FFLINS_OK;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> keep;
    for (uint64_t id : ctx->pending) {
        Frame *f = find_frame(ctx, id);
        if (!f || f->pending_slot < 0) continue;
        if (f->bg) {
            keep.push_back(id);
            continue;
        }
        const int *cnt                 = ctx->h_slots + 2 * f->pending_slot;
        f->n_fallback                  = cnt[0];
        f->c[FFLINS_CLOUD_LIDAR].n     = cnt[1];
        f->c[FFLINS_CLOUD_WORLD].n     = cnt[1];
        f->pending_slot                = -1;
    }
    ctx->pending.swap(keep);
    return FFLINS_OK;
}
void free_frame(fflins_ctx *ctx, Frame *f);

// Wait for the keyframe-map job on the map stream (if any), adopt its voxel count and free what had to
// outlive it. Called by everything that reads a map cloud, its size, or frees buffers the job may touch.
int resolve_map(fflins_ctx *ctx) {
    if (!ctx->map_job_active) return FFLINS_OK;
    {
        // the helper thread must have recorded the completion event before we can wait on it
        std::unique_lock<std::mutex> lk(ctx->map_mu);
        ctx->map_cv.wait(lk, [ctx] { return ctx->map_issued; });
    }
    CK(cudaEventSynchronize(ctx->ev_map_done));
    ctx->map_job_active = false;
    if (Frame *key = find_frame(ctx, ctx->map_job_key)) {
        const int nm = ctx->h_map_count[0];
        key->c[FFLINS_CLOUD_MAP].n = nm;
        ctx->last_map_n            = nm;
        if (key->hash_early) { // the map job built the index too: valid unless the size guess was too small
            key->hash_early   = false;
            key->hash_pending = (long long) nm * 10 > (long long) key->hcap * 7;
        }
    }
    {
        bool any_bg = false; // (a frame released while an unrelated job was running may still be busy on the background stream)
        for (Frame *f : ctx->deferred_frames) any_bg |= f->bg;
        if (any_bg) {
            int rcj = join_bg(ctx);
            if (rcj) return rcj;
        }
    }
    for (Frame *f : ctx->deferred_frames) free_frame(ctx, f);
    ctx->deferred_frames.clear();
    for (void *b : ctx->deferred_blocks) ctx->pool.release(b);
    ctx->deferred_blocks.clear();
    return FFLINS_OK;
}

// reserve (>= n): capacity to allocate when the cloud has to grow anyway
int ensure_cloud(fflins_ctx *ctx, Cloud &c, int n, bool keep = false, long long reserve = 0) {
    if (n <= c.cap) return FFLINS_OK;
    size_t want = Pool::round((size_t) std::max<long long>(std::max(n, 256), std::min<long long>(reserve, 1ll << 27)) * sizeof(float4));
    float4 *p   = (float4 *) ctx->pool.alloc(want);
    if (!p) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    if (keep && c.d && c.n > 0)
        cudaMemcpyAsync(p, c.d, (size_t) c.n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream);
    // single-stream context: handing the old block back to the pool right away is safe, every later
    // use of it is ordered after the copy above
    if (c.d) ctx->pool.release(c.d);
    c.d   = p;
    c.cap = (int) (want / sizeof(float4));
    return FFLINS_OK;
}

void free_frame(fflins_ctx *ctx, Frame *f) {
    for (auto &c : f->c) ctx->pool.release(c.d);
    ctx->pool.release(f->index.fine);
    ctx->pool.release(f->index.coarse);
    ctx->pool.release(f->index.super);
    ctx->pool.release(f->feat.type);
    ctx->pool.release(f->feat.covered);
    ctx->pool.release(f->feat.outlier);
    ctx->pool.release(f->feat.nn_idx);
    ctx->pool.release(f->feat.nn_d2);
    ctx->pool.release(f->feat.normal);
    ctx->pool.release(f->feat.d);
    ctx->pool.release(f->feat_counts);
    delete f;
}

int flush_frames(fflins_ctx *ctx);

int ensure_stage(fflins_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->stage_cap) return FFLINS_OK;
    int rc;
    if ((rc = flush_frames(ctx))) return rc;
    if (ctx->d_stage) {
        CK(cudaStreamSynchronize(ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->bg_stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->pool.release(ctx->d_stage);
    size_t cap   = Pool::round(bytes);
    ctx->d_stage = ctx->pool.alloc(cap * kStageRing);
    if (!ctx->d_stage) {
        ctx->stage_cap = 0;
        return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    }
    ctx->stage_cap = cap;
    for (auto &g : ctx->stage_guard) g = 0;
    ctx->announced.clear();   // (the ring moved: an announced frame is copied again when it is processed)
    return FFLINS_OK;
}

// Next slot of the device staging ring: the copy stream first waits until the batch that read the slot's previous
// content has consumed it (a slot still waiting in the frame queue forces the queue out first).
int stage_acquire(fflins_ctx *ctx, size_t bytes, char **ptr, int *slot) {
    int rc;
    if ((rc = ensure_stage(ctx, bytes))) return rc;
    const int k = ctx->stage_next;
    for (size_t a = 0; a < ctx->announced.size(); a++)
        if (ctx->announced[a].slot == k) { // the ring came round to an announced frame that was never processed: dropped
            ctx->stage_guard[k] = 0;       // (its copy is on the copy stream like every later writer of the slot: FIFO)
            ctx->announced.erase(ctx->announced.begin() + a);
            break;
        }
    if (ctx->stage_guard[k] == kQueued && (rc = flush_frames(ctx))) return rc;
    ctx->stage_next = (k + 1) % kStageRing;
    // (copies are FIFO on the copy stream, so waiting here also covers a copy that is issued later, at flush time)
    if (ctx->stage_guard[k]) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->batch_ev[ctx->stage_guard[k] % kGuardRing], 0));
    ctx->stage_guard[k] = kQueued;
    *ptr  = (char *) ctx->d_stage + (size_t) k * ctx->stage_cap;
    *slot = k;
    return FFLINS_OK;
}

int ensure_ins(fflins_ctx *ctx, int w) {
    if (w <= ctx->ins_cap) return FFLINS_OK;
    int rc;
    if ((rc = flush_frames(ctx))) return rc;
    ctx->pool.release(ctx->d_tab);
    int cap    = std::max(w, 2048);
    // interval rows (K1 prologue) + time stamps, one set per batch item
    ctx->d_tab = (double *) ctx->pool.alloc((size_t) kFrameBatch * cap * (kRowSize + 1 + 8) * sizeof(double));
    if (!ctx->d_tab) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    ctx->ins_cap = cap;
    return FFLINS_OK;
}

int ensure_voxel_work(fflins_ctx *ctx, VoxelWork &vw, void *&base, int n) {
    if (n <= vw.cap) return FFLINS_OK;
    ctx->pool.release(base);
    int cap = (int) (Pool::round((size_t) std::max(n, 4096)));
    size_t off[8];
    size_t bytes = voxel_work_bytes(cap, off);
    base         = ctx->pool.alloc(bytes);
    if (!base) {
        vw.cap = 0;
        return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    }
    char *b        = (char *) base;
    vw.keys[0]     = (uint32_t *) (b + off[0]);
    vw.keys[1]     = (uint32_t *) (b + off[1]);
    vw.idx[0]      = (uint32_t *) (b + off[2]);
    vw.idx[1]      = (uint32_t *) (b + off[3]);
    vw.block_hist  = (uint32_t *) (b + off[4]);
    vw.block_heads = (uint32_t *) (b + off[5]);
    vw.meta        = (int *) (b + off[6]);
    vw.seg_start   = (uint32_t *) (b + off[7]);
    vw.cap         = cap;
    vw.fresh       = true;
    return FFLINS_OK;
}
int ensure_voxel(fflins_ctx *ctx, int n) { return ensure_voxel_work(ctx, ctx->vox, ctx->vox_base, n); }

unsigned char *pin(fflins_ctx *ctx, size_t bytes) {
    if (bytes > ctx->h_pin_cap) {
        if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
        size_t cap = std::max<size_t>(bytes, 1 << 20);
        if (cudaMallocHost((void **) &ctx->h_pin, cap) != cudaSuccess) {
            ctx->h_pin     = nullptr;
            ctx->h_pin_cap = 0;
            return nullptr;
        }
        ctx->h_pin_cap = cap;
    }
    return ctx->h_pin;
}

// Waits until the kernel(s) on `s` have released flags[0..n) = seq (pinned host memory, see factor_kernel).
// Polling costs ~1 us after the device's store lands; a stream synchronisation costs 5-10 us. The stream is
// queried every few thousand polls so that a failed launch or a sticky device error ends the wait.
cudaError_t wait_signal(const unsigned long long *flags, int n, unsigned long long seq, cudaStream_t s) {
    for (int i = 0; i < n; i++) {
        unsigned spins = 0;
        while (__atomic_load_n(flags + i, __ATOMIC_ACQUIRE) != seq) {
            if ((++spins & 0xfffu) == 0) {
                cudaError_t e = cudaStreamQuery(s);
                if (e == cudaSuccess) break;                 // stream drained: the stores are visible
                if (e != cudaErrorNotReady) return e;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    return cudaSuccess;
}

void fill_info(Frame *f, fflins_frame_info *out) {
    if (!out) return;
    out->id         = f->id;
    out->n_raw      = f->n_raw;
    out->n_dec      = f->c[FFLINS_CLOUD_LIDAR_FULL].n;
    out->n_down     = f->c[FFLINS_CLOUD_LIDAR].n;
    out->n_map      = f->c[FFLINS_CLOUD_MAP].n;
    out->n_fallback = f->n_fallback;
    out->reserved   = 0;
    out->pose       = f->pose;
}

// Shared tail of the two lidarFrameProcessing overloads: voxel filter of the decimated cloud,
// world projection, result read-back (one synchronisation per frame).
// FFLINS_TRACE_HOST=1 (debug): where the host time of an entry point goes; printed every 1,000 calls
struct HostTrace {
    const char *what;
    const char *lap_name[6];
    bool on = getenv("FFLINS_TRACE_HOST") != nullptr;
    double us[6] = {0, 0, 0, 0, 0, 0};
    long long calls = 0;
    std::chrono::steady_clock::time_point t;
    void start() { if (on) t = std::chrono::steady_clock::now(); }
    void lap(int k) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        us[k] += std::chrono::duration<double, std::micro>(n - t).count();
        t = n;
    }
    void done() {
        if (!on || ++calls % 1000) return;
        std::fprintf(stderr, "[fflins] %s host time (us):", what);
        for (int k = 0; k < 6 && lap_name[k]; k++) std::fprintf(stderr, " %s %.2f |", lap_name[k], us[k] / calls);
        std::fprintf(stderr, "\n");
    }
};
HostTrace g_merge_trace{"keyframe_merge", {"flush_frames", "resolve_map", "prepare", "hand over", nullptr, nullptr}};
HostTrace g_enqueue_trace{"frame_process (queued)", {"check + lookup", "pack window", "stage", "pose", "buffers + queue", nullptr}};

int check_window(fflins_ctx *ctx, const double *ins_t, int w) {
    REQUIRE(w >= 2 && w <= 65536, "INS window must hold 2..65536 entries");
    for (int i = 1; i < w; i++) REQUIRE(ins_t[i] > ins_t[i - 1], "INS window times must be strictly increasing");
    return FFLINS_OK;
}

// t | p | q packed into a pinned slot (64 B per entry) that the prologue kernel reads in place. The slots form a
// ring: a slot is repacked only after the batch that read it has run.
int pack_window(fflins_ctx *ctx, const double *ins_t, const double *ins_p, const double *ins_q, int w, double **hp_out,
                int *slot_out) {
    const size_t need = (size_t) w * 64;
    int rc;
    if (need > ctx->h_win_slot) {
        if ((rc = flush_frames(ctx))) return rc;
        CK(cudaStreamSynchronize(ctx->stream)); // prologue kernels read the ring in place
        if (ctx->h_win) cudaFreeHost(ctx->h_win);
        ctx->h_win_slot = std::max<size_t>(need, 4096 * 64);
        if (cudaMallocHost((void **) &ctx->h_win, ctx->h_win_slot * kWinRing) != cudaSuccess) {
            ctx->h_win      = nullptr;
            ctx->h_win_slot = 0;
            return fail(ctx, FFLINS_E_NOMEM, "pinned allocation failed");
        }
        for (auto &g : ctx->win_guard) g = 0;
    }
    const int k = ctx->win_next;
    if (ctx->win_guard[k] == kQueued && (rc = flush_frames(ctx))) return rc;
    ctx->win_next = (k + 1) % kWinRing;
    if (ctx->win_guard[k]) CK(cudaEventSynchronize(ctx->batch_ev[ctx->win_guard[k] % kGuardRing]));
    ctx->win_guard[k] = kQueued;
    double *hp = (double *) (ctx->h_win + (size_t) k * ctx->h_win_slot);
    std::memcpy(hp, ins_t, (size_t) w * 8);
    std::memcpy(hp + w, ins_p, (size_t) w * 24);
    std::memcpy(hp + 4 * (size_t) w, ins_q, (size_t) w * 32);
    *hp_out   = hp;
    *slot_out = k;
    return FFLINS_OK;
}

// A new batch id; its event is recorded on the compute stream by the caller once the inputs are consumed.
unsigned long long next_batch(fflins_ctx *ctx) {
    const unsigned long long seq = ++ctx->batch_seq;
    if (!ctx->batch_ev[seq % kGuardRing]) cudaEventCreateWithFlags(&ctx->batch_ev[seq % kGuardRing], cudaEventDisableTiming);
    return seq;
}

// Bookkeeping shared by every frame entry point: clouds sized, counts pending in a pinned slot.
int frame_begin(fflins_ctx *ctx, Frame *f, int n, int ndec, int **report) {
    int rc;
    // any frame may become a keyframe, whose full cloud then takes the points of the frames merged into it: with room for
    // them from the start the merge neither reallocates nor copies the keyframe's own points (HBM is not the constraint)
    if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_MAP_FULL], n, false, (long long) n * ctx->merge_reserve))) return rc;
    if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_LIDAR_FULL], ndec))) return rc;
    if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_LIDAR], ndec))) return rc;
    if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_WORLD], ndec))) return rc;
    if ((int) ctx->pending.size() >= kSlots - 1 && (rc = resolve_pending(ctx))) return rc;
    const int slot = ctx->next_slot;
    ctx->next_slot = (ctx->next_slot + 1) % kSlots;
    *report        = ctx->h_slots + 2 * slot;
    (*report)[0] = (*report)[1] = -1; // never adopt the slot's previous user's counts
    f->n_raw                        = n;
    f->n_fallback                   = 0;
    f->c[FFLINS_CLOUD_MAP_FULL].n   = n;
    f->c[FFLINS_CLOUD_LIDAR_FULL].n = ndec;
    f->c[FFLINS_CLOUD_LIDAR].n      = 0;
    f->c[FFLINS_CLOUD_WORLD].n      = 0;
    f->c[FFLINS_CLOUD_MAP].n        = 0;
    f->pending_slot                 = slot;
    ctx->pending.push_back(f->id);
    return FFLINS_OK;
}

// Voxel filter + world projection + count report of the batch: one CTA per frame, or the multi-kernel filter
// for a decimated cloud too large for a single CTA.
int frame_voxel_tail(fflins_ctx *ctx, cudaStream_t s, const FrameBatch &B) {
    bool any_small = false;
    for (int i = 0; i < B.count; i++) any_small |= B.it[i].ndec <= kVoxSmallCap;
    if (any_small) {
        ProfScope ps(ctx->prof(), "voxel_frame.one_cta", s);
        launch_frame_voxel(s, B);
        ctx->launches++;
    }
    for (int i = 0; i < B.count; i++) {
        const FrameItem &it = B.it[i];
        if (it.ndec <= kVoxSmallCap) continue;
        int rc;
        if ((rc = ensure_voxel(ctx, it.ndec))) return rc;
        int64_t nl = 0;
        launch_voxel_downsample(s, ctx->vox, it.out_dec, it.ndec, (float) ctx->cfg.downsample_size, it.down, it.d_nout, nullptr,
                                &nl, ctx->prof(), "voxel_frame");
        ctx->launches += nl;
        ProfScope ps(ctx->prof(), "project_world", s);
        launch_project(s, it.frame, it.down, it.world, it.ndec, it.d_nout, it.report, it.counters);
        ctx->launches++;
    }
    return FFLINS_OK;
}

// Per-item scratch of batch slot i.
void item_scratch(fflins_ctx *ctx, int i, FrameItem &it) {
    it.rows     = ctx->d_tab + (size_t) i * ctx->ins_cap * (kRowSize + 1 + 8);
    it.t_dev    = it.rows + (size_t) ctx->ins_cap * kRowSize;
    it.counters = ctx->d_fscratch + i * kItemScratch;
    it.d_nout   = it.counters + 1;
    it.meta     = it.counters + 16;
}

// Prologue, undistortion and voxel tail of one batch on stream s; *seq_out = the batch id whose event guards the
// batch's window and staging slots.
int launch_chain(fflins_ctx *ctx, cudaStream_t s, const FrameBatch &B, unsigned long long *seq_out) {
    {
        ProfScope ps(ctx->prof(), "ins_prep", s);
        launch_frame_prologue(s, B);
        ctx->launches++;
    }
    {
        ProfScope ps(ctx->prof(), "undistort", s);
        launch_frame_undistort(s, B);
        ctx->launches++;
    }
    const unsigned long long seq = next_batch(ctx);
    CK(cudaEventRecord(ctx->batch_ev[seq % kGuardRing], s)); // window slots and staging slots may be refilled after this
    *seq_out = seq;
    return FFLINS_OK;
}

int issue_copies(fflins_ctx *ctx, const QueuedFrame &q) {
    for (int c = 0; c < q.n_copy; c++)
        CK(cudaMemcpyAsync(q.d_dst[c], q.h_src[c], q.bytes[c], cudaMemcpyHostToDevice, ctx->copy_stream));
    return FFLINS_OK;
}

bool staged_any(const fflins_ctx *ctx) {
    for (const QueuedFrame &q : ctx->fq)
        if (q.stage_slot >= 0) return true;
    return false;
}

// Frames whose points cross PCIe on the copy stream take their INS window along (64 bytes per entry, in front of the points):
// the prologue's zero-copy read of the pinned window costs ~8 us on an idle link, but 100+ us while the background copies
// of the batch saturate it (measured: newest frame copied -> its chain done 150-170 us instead of ~45 us).
int stage_window(fflins_ctx *ctx, const QueuedFrame &q, FrameItem &it) {
    if (q.stage_slot < 0 || it.w <= 0) return FFLINS_OK;
    double *dst = it.t_dev + ctx->ins_cap;   // behind the item's time stamps (item_scratch)
    CK(cudaMemcpyAsync(dst, it.win, (size_t) it.w * 64, cudaMemcpyHostToDevice, ctx->copy_stream));
    it.win = dst;
    return FFLINS_OK;
}

// Launch the kernel chain of every queued frame: prologue, undistortion, voxel tail — three launches per batch.
//
// Batches of two or more frames are SPLIT: the newest frame (the one the caller is about to use: association and
// factor evaluation only ever need the current keyframe) is copied first and processed on the compute stream; the
// older frames of the batch — which only feed the keyframe-map merge — are copied behind it (pinned host buffers: their
// copies are issued here, newest first) and processed on the background stream, so that their PCIe and kernel time overlaps
// the solver instead of preceding it.
int flush_frames(fflins_ctx *ctx) {
    if (ctx->fq.empty()) return FFLINS_OK;
    cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
    const int count = (int) ctx->fq.size();
    const int L     = count - 1;
    int rc;
    if ((rc = join_bg(ctx))) return rc; // scratch slots and stream order: one background part at a time
    ctx->scratch_gen++;
    // the staged windows land in the per-item scratch (stage_window): the copy stream must not overwrite a slot whose
    // previous prologue may still be reading it (compute stream: slot of the newest frame; background stream: the others)
    if (staged_any(ctx)) {
        if (ctx->last_main_seq) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->batch_ev[ctx->last_main_seq % kGuardRing], 0));
        // (not `had_bg`: join_bg may have been called since — it orders the COMPUTE stream behind that part, not this one)
        if (ctx->copy_owes_bg_wait && !(count >= 2 && !ctx->prof_on)) {
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_bg_done, 0));
            ctx->copy_owes_bg_wait = false;
        }
    }
    const bool split = count >= 2 && !ctx->prof_on;
    if (split) CK(cudaEventRecord(ctx->ev_flush, s));
    bool staged = false;
    for (const QueuedFrame &q : ctx->fq) staged |= q.stage_slot >= 0;
    FrameBatch B{};
    if (!split) {
        B.count = count;
        for (int i = 0; i < count; i++) {
            B.it[i] = ctx->fq[i].item;
            item_scratch(ctx, i, B.it[i]);
            if ((rc = stage_window(ctx, ctx->fq[i], B.it[i]))) return rc;
            if ((rc = issue_copies(ctx, ctx->fq[i]))) return rc;
        }
        if (staged) { // the points of these frames travel on the copy stream
            CK(cudaEventRecord(ctx->ev_copied, cs));
            CK(cudaStreamWaitEvent(s, ctx->ev_copied, 0));
        }
        unsigned long long seq;
        if ((rc = launch_chain(ctx, s, B, &seq))) return rc;
        for (int i = 0; i < count; i++) {
            const QueuedFrame &q = ctx->fq[i];
            ctx->win_guard[q.wslot] = seq;
            if (q.stage_slot >= 0) ctx->stage_guard[q.stage_slot] = seq;
            q.f->d_ndown       = B.it[i].d_nout;
            q.f->d_ndown_gen = ctx->scratch_gen;
        }
        ctx->last_main_seq = seq;
        ctx->fq.clear();
        rc = frame_voxel_tail(ctx, s, B);
        CK(cudaGetLastError());
        return rc;
    }
    // newest frame: its copy goes first, the compute stream runs its chain
    static const bool copy_tl = getenv("FFLINS_COPY_TIMELINE") != nullptr;
    if (copy_tl) {
        if (!ctx->ev_tl[0])
            for (auto &e : ctx->ev_tl) CK(cudaEventCreate(&e));
        if (ctx->tl_armed) { // the previous batch: where its copies and kernels sat in time (synchronises: debug only)
            CK(cudaEventSynchronize(ctx->ev_tl[4]));
            CK(cudaEventSynchronize(ctx->ev_tl[5]));
            float a = 0, b = 0, c2 = 0, d = 0, e = 0;
            cudaEventElapsedTime(&a, ctx->ev_tl[0], ctx->ev_tl[1]);
            cudaEventElapsedTime(&b, ctx->ev_tl[1], ctx->ev_tl[2]);
            cudaEventElapsedTime(&c2, ctx->ev_tl[2], ctx->ev_tl[3]);
            cudaEventElapsedTime(&d, ctx->ev_tl[3], ctx->ev_tl[4]);
            cudaEventElapsedTime(&e, ctx->ev_tl[1], ctx->ev_tl[5]);
            fprintf(stderr, "[fflins] copy timeline (us): wait+newest copy %.1f | %d background copies %.1f | -> bg kernels start %.1f | bg kernels %.1f | newest copied -> its chain done %.1f\n",
                    a * 1e3, L, b * 1e3, c2 * 1e3, d * 1e3, e * 1e3);
        }
        CK(cudaEventRecord(ctx->ev_tl[0], cs));
    }
    const QueuedFrame &ql = ctx->fq[L];
    B.count = 1;
    B.it[0] = ql.item;
    item_scratch(ctx, L, B.it[0]);
    if ((rc = stage_window(ctx, ql, B.it[0]))) return rc;
    if ((rc = issue_copies(ctx, ql))) return rc;
    if (copy_tl) CK(cudaEventRecord(ctx->ev_tl[1], cs));
    if (ql.stage_slot >= 0) {
        CK(cudaEventRecord(ctx->ev_copied, cs));
        CK(cudaStreamWaitEvent(s, ctx->ev_copied, 0));
    }
    unsigned long long seq;
    if ((rc = launch_chain(ctx, s, B, &seq))) return rc;
    ctx->win_guard[ql.wslot] = seq;
    if (ql.stage_slot >= 0) ctx->stage_guard[ql.stage_slot] = seq;
    ql.f->d_ndown       = B.it[0].d_nout;
    ql.f->d_ndown_gen = ctx->scratch_gen;
    if ((rc = frame_voxel_tail(ctx, s, B))) return rc;
    if (copy_tl) CK(cudaEventRecord(ctx->ev_tl[5], s));
    // the rest: their copies in FIFO order on the copy stream, their kernels on the background stream behind the last of
    // those copies. The copy stream ONLY copies: were the kernels queued on it as well, the newest frame of the NEXT batch
    // could not start crossing PCIe until they had run (~60 us of kernels per batch in front of that copy). Scratch slots
    // 0..L-1 were last used by the previous background part (same stream).
    // Pool blocks handed to these frames may have been released by earlier compute-stream work (re-projection, voxel
    // tail, ...): the kernels start behind everything the compute stream had queued when the batch was flushed.
    cudaStream_t bs = ctx->bg_stream;
    CK(cudaStreamWaitEvent(bs, ctx->ev_flush, 0));
    if (ctx->copy_owes_bg_wait && staged) { // (see above: scratch slots 0..L-1 of the previous part)
        CK(cudaStreamWaitEvent(cs, ctx->ev_bg_done, 0));
        ctx->copy_owes_bg_wait = false;
    }
    ctx->last_main_seq = seq;
    FrameBatch G{};
    G.count = L;
    for (int i = 0; i < L; i++) {
        G.it[i] = ctx->fq[i].item;
        item_scratch(ctx, i, G.it[i]);
        if ((rc = stage_window(ctx, ctx->fq[i], G.it[i]))) return rc;
        if ((rc = issue_copies(ctx, ctx->fq[i]))) return rc;
        ctx->fq[i].f->bg = true;
    }
    // (pageable sources were copied at call time, also on the copy stream: the event covers them too)
    if (staged) {   // (device-resident frames: nothing crosses the copy stream)
        CK(cudaEventRecord(ctx->ev_bg_copied, cs));
        CK(cudaStreamWaitEvent(bs, ctx->ev_bg_copied, 0));
    }
    if (copy_tl) CK(cudaEventRecord(ctx->ev_tl[2], cs));
    if (copy_tl) CK(cudaEventRecord(ctx->ev_tl[3], bs));
    unsigned long long seq_bg;
    if ((rc = launch_chain(ctx, bs, G, &seq_bg))) return rc;
    for (int i = 0; i < L; i++) {
        ctx->win_guard[ctx->fq[i].wslot] = seq_bg;
        if (ctx->fq[i].stage_slot >= 0) ctx->stage_guard[ctx->fq[i].stage_slot] = seq_bg;
    }
    ctx->fq.clear();
    if ((rc = frame_voxel_tail(ctx, bs, G))) return rc;
    CK(cudaEventRecord(ctx->ev_bg_done, bs));
    if (copy_tl) {
        CK(cudaEventRecord(ctx->ev_tl[4], bs));
        ctx->tl_armed = true;
    }
    ctx->bg_active = true;
    ctx->bg_staged = staged;
    ctx->copy_owes_bg_wait = true;
    CK(cudaGetLastError());
    return FFLINS_OK;
}

// Common body of fflins_frame_process / _aos / _dev: stage the inputs, queue the frame, and (synchronous
// call, or a full queue) launch.
int enqueue_frame(fflins_ctx *ctx, uint64_t id, const float *xyzi, const double *time, const void *aos, bool on_device,
                  int n, const double *ins_t, const double *ins_p, const double *ins_q, int w, const fflins_pose *pose_b_l,
                  double stamp, double td, fflins_frame_info *out) {
    int rc;
    g_enqueue_trace.start();
    if ((rc = check_window(ctx, ins_t, w))) return rc;
    ctx->eval_inputs_settled = false;
    Frame *f = get_frame(ctx, id);
    // one input kind per batch; a frame is queued at most once
    bool must_flush = (int) ctx->fq.size() >= kFrameBatch;
    for (const QueuedFrame &q : ctx->fq) must_flush |= q.f == f || (q.item.aos != nullptr) != (aos != nullptr);
    if (must_flush && (rc = flush_frames(ctx))) return rc;
    if (f->bg && (rc = join_bg(ctx))) return rc; // re-processing a frame whose previous chain is still on the background stream
    if ((rc = ensure_ins(ctx, w))) return rc;
    const int F    = ctx->cfg.point_filter_num;
    const int ndec = (n + F - 1) / F;
    QueuedFrame q{};
    q.f          = f;
    q.stage_slot = -1;
    FrameItem &it = q.item;
    double *hp;
    g_enqueue_trace.lap(0);
    if ((rc = pack_window(ctx, ins_t, ins_p, ins_q, w, &hp, &q.wslot))) return rc;
    g_enqueue_trace.lap(1);
    it.win = hp;
    it.w   = w;
    if (on_device || n == 0) {
        it.xyzi = (const float4 *) xyzi;
        it.time = time;
        it.aos  = aos;
    } else { // host-resident points: every host-to-device copy goes to the copy stream
        size_t bytes = aos ? (size_t) n * 32 : (size_t) n * 24;
        char *pts;
        int ann = -1;
        for (size_t a = 0; aos && a < ctx->announced.size(); a++)
            if (ctx->announced[a].host == aos && ctx->announced[a].bytes == bytes) ann = (int) a;
        if (ann >= 0) {
            // announced ahead (fflins_host_prefetch): the records are already in, or on their way into, a staging slot
            q.stage_slot = ctx->announced[ann].slot;
            q.n_copy     = 0;
            it.aos       = ctx->announced[ann].dev;
            ctx->announced.erase(ctx->announced.begin() + ann);
        } else {
        if ((rc = stage_acquire(ctx, bytes + 256, &pts, &q.stage_slot))) return rc;
        if (aos) {
            q.n_copy   = 1;
            q.h_src[0] = aos;
            q.d_dst[0] = pts;
            q.bytes[0] = (size_t) n * 32;
            it.aos     = pts;
        } else {
            q.n_copy   = 2;
            q.h_src[0] = xyzi;
            q.d_dst[0] = pts;
            q.bytes[0] = (size_t) n * 16;
            q.h_src[1] = time;
            q.d_dst[1] = pts + (size_t) n * 16;
            q.bytes[1] = (size_t) n * 8;
            it.xyzi    = (const float4 *) pts;
            it.time    = (const double *) (pts + (size_t) n * 16);
        }
        // Pinned (page-locked / registered) buffers are copied when the batch is launched, newest frame first: an
        // asynchronous copy from pinned memory already obliges the caller to keep the buffer until the frame's results
        // are consumed. Pageable buffers are copied now (the runtime stages them before cudaMemcpyAsync returns).
        bool pinned = !out;
        for (int c = 0; c < q.n_copy && pinned; c++) {
            cudaPointerAttributes at{};
            pinned = cudaPointerGetAttributes(&at, q.h_src[c]) == cudaSuccess && at.type == cudaMemoryTypeHost;
        }
        cudaGetLastError();
        if (!pinned) {
            // pageable source: into the page-locked bounce ring with the copy pool (the caller's buffer is free again when
            // the call returns), then an ordinary pinned copy
            size_t total = 0;
            for (int c = 0; c < q.n_copy; c++) total += (q.bytes[c] + 255) & ~(size_t) 255;
            static const bool no_bounce = getenv("FFLINS_NO_BOUNCE") != nullptr;
            if (!no_bounce && total >= (1u << 20)) {
                if (total > ctx->bounce_cap) {
                    CK(cudaStreamSynchronize(ctx->copy_stream));
                    if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
                    ctx->bounce_cap = Pool::round(total);
                    if (cudaMallocHost((void **) &ctx->h_bounce, ctx->bounce_cap * kStageRing) != cudaSuccess) {
                        ctx->h_bounce   = nullptr;
                        ctx->bounce_cap = 0;
                        return fail(ctx, FFLINS_E_NOMEM, "pinned allocation failed");
                    }
                    ctx->copy_pool.start();
                }
                const int k = ctx->bounce_next;
                ctx->bounce_next = (k + 1) % kStageRing;
                if (!ctx->bounce_ev[k]) CK(cudaEventCreateWithFlags(&ctx->bounce_ev[k], cudaEventDisableTiming));
                else CK(cudaEventSynchronize(ctx->bounce_ev[k]));   // the slot's previous copy has left it
                unsigned char *slot = ctx->h_bounce + (size_t) k * ctx->bounce_cap;
                size_t off = 0;
                for (int c = 0; c < q.n_copy; c++) {
                    ctx->copy_pool.copy(slot + off, q.h_src[c], q.bytes[c]);
                    q.h_src[c] = slot + off;
                    off += (q.bytes[c] + 255) & ~(size_t) 255;
                }
                if ((rc = issue_copies(ctx, q))) return rc;
                CK(cudaEventRecord(ctx->bounce_ev[k], ctx->copy_stream));
            } else if ((rc = issue_copies(ctx, q))) {
                return rc;
            }
            q.n_copy = 0;
        }
        }
    }
    g_enqueue_trace.lap(2);
    it.ext = to12(*pose_b_l);
    pose_from_window(ins_t, ins_p, ins_q, w, it.ext, stamp, it.frame);
    std::memcpy(f->pose.R, it.frame.R, sizeof(it.frame.R));
    std::memcpy(f->pose.t, it.frame.t, sizeof(it.frame.t));
    f->td = td;
    g_enqueue_trace.lap(3);
    if ((rc = frame_begin(ctx, f, n, ndec, &it.report))) return rc;
    it.n          = n;
    it.td         = td;
    it.filter_num = F;
    // k / F by multiply-high: exact for k < 2^32 / F (ceil(2^32 / F) * F - 2^32 < F)
    it.dec_magic  = (F > 1 && (unsigned long long) n * (unsigned long long) F < (1ull << 32)) ? (unsigned) (((1ull << 32) + F - 1) / F) : 0u;
    it.t_front    = ins_t[0];
    it.t_back     = ins_t[w - 1];
    it.inv_dt     = (double) (w - 1) / (ins_t[w - 1] - ins_t[0]);
    it.out_full   = f->c[FFLINS_CLOUD_MAP_FULL].d;
    it.out_dec    = f->c[FFLINS_CLOUD_LIDAR_FULL].d;
    it.ndec       = ndec;
    it.inv_leaf   = 1.0f / (float) ctx->cfg.downsample_size;
    it.down       = f->c[FFLINS_CLOUD_LIDAR].d;
    it.world      = f->c[FFLINS_CLOUD_WORLD].d;
    ctx->fq.push_back(q);
    g_enqueue_trace.lap(4);
    g_enqueue_trace.done();
    if (!out) return FFLINS_OK; // asynchronous: launched together with the next frames, read back on demand
    if ((rc = resolve_pending(ctx))) return rc;
    fill_info(f, out);
    return FFLINS_OK;
}

int next_pow2(int v) {
    int r = 1024;
    while (r < v) r <<= 1;
    return r;
}

int ensure_features(fflins_ctx *ctx, Frame *f, int q) {
    if (q <= f->feat_cap && f->feat.type) return FFLINS_OK;
    int cap = std::max(q, 1024);
    Pool &P = ctx->pool;
    P.release(f->feat.type);
    P.release(f->feat.covered);
    P.release(f->feat.outlier);
    P.release(f->feat.nn_idx);
    P.release(f->feat.nn_d2);
    P.release(f->feat.normal);
    P.release(f->feat.d);
    f->feat.type    = (int8_t *) P.alloc(cap);
    f->feat.covered = (uint8_t *) P.alloc(cap);
    f->feat.outlier = (uint8_t *) P.alloc(cap);
    f->feat.nn_idx  = (int32_t *) P.alloc((size_t) cap * 20);
    f->feat.nn_d2   = (float *) P.alloc((size_t) cap * 20);
    f->feat.normal  = (double *) P.alloc((size_t) cap * 24);
    f->feat.d       = (double *) P.alloc((size_t) cap * 8);
    if (!f->feat_counts) f->feat_counts = (int *) P.alloc(256);
    if (!f->feat.type || !f->feat.covered || !f->feat.outlier || !f->feat.nn_idx || !f->feat.nn_d2 || !f->feat.normal ||
        !f->feat.d || !f->feat_counts)
        return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    f->feat_cap = cap;
    return FFLINS_OK;
}

float search_cell(const fflins_ctx *ctx) { return (float) ctx->cfg.downsample_size * ctx->cell_scale; }

// The kNN queue entries pack cell / block offsets into 10-bit fields biased by 512 (k_assoc.cu): offsets reach
// -(r_max + 15), so the search radius in cells is bounded.
constexpr int kMaxSearchCells = 511 - 16;
int search_r_max(double max_dist, float cell) { return (int) std::ceil(max_dist / cell) + 1; }

int build_hash(fflins_ctx *ctx, Frame *f) {
    const Cloud &m = f->c[FFLINS_CLOUD_MAP];
    int hcap       = next_pow2(2 * std::max(m.n, 1));
    if (hcap > f->hcap) {
        ctx->pool.release(f->index.fine);
        ctx->pool.release(f->index.coarse);
        ctx->pool.release(f->index.super);
        f->index.fine   = (FineSlot *) ctx->pool.alloc((size_t) hcap * sizeof(FineSlot));
        f->index.coarse = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
        f->index.super  = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
        if (!f->index.fine || !f->index.coarse || !f->index.super) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    }
    f->hcap        = hcap;
    f->index.hmask = hcap - 1;
    ProfScope ps(ctx->prof(), "hash_build", ctx->stream);
    launch_hash_build(ctx->stream, m.d, m.n, 1.0f / search_cell(ctx), f->index);
    ctx->launches += (m.n > 0) ? 2 : 1;
    return FFLINS_OK;
}

void fill_assoc_common(const fflins_ctx *ctx, AssocParams &P) {
    float cell      = search_cell(ctx);
    P.inv_cell      = 1.0f / cell;
    P.cell          = cell;
    double max_dist = ctx->cfg.max_search_square_distance; // passed as max_dist and squared (ikd_Tree.cc:902)
    P.r_max         = search_r_max(max_dist, cell);
    P.max_d2        = (float) (max_dist * max_dist);
    P.plane_d2_gate = (float) ctx->cfg.max_search_square_distance;
    P.plane_thr     = ctx->cfg.plane_estimation_threshold;
    P.sigma         = ctx->cfg.point_to_plane_std;
    P.scalar_thr    = ctx->cfg.plane_scalar_threshold;
    P.sigma_gate    = ctx->cfg.plane_sigma_gate;
}

// world -> ref transform as formed at lidar_map.cc:343-345: R^T and (-R^T) t
Pose12 ref_world(const fflins_pose &pose) {
    Pose12 T;
    double neg[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            T.R[3 * i + j] = pose.R[3 * j + i];
            neg[3 * i + j] = -T.R[3 * i + j];
        }
    for (int i = 0; i < 3; i++)
        T.t[i] = (neg[3 * i] * pose.t[0] + neg[3 * i + 1] * pose.t[1]) + neg[3 * i + 2] * pose.t[2];
    return T;
}

int run_associate(fflins_ctx *ctx, const std::vector<Frame *> &refs, Frame *cur, int32_t *nf, int32_t *nc) {
    cudaStream_t s = ctx->stream;
    bool need_bg = false, late_q = false;
    ctx->eval_inputs_settled = false; // features are about to be rewritten on the stream
    {
        int rcp = flush_frames(ctx);
        if (rcp) return rcp;
        need_bg = cur->bg; // only the frames involved have to be complete
        for (Frame *r : refs) need_bg |= r->bg;
        // The current frame's chain has just been launched on this stream and its voxel count is still on the device:
        // launch the search for the upper bound (the decimated size) and let the kernels read the count, instead of
        // synchronising the stream in the middle of the association's critical path. Counts are adopted afterwards.
        static const bool no_poll = getenv("FFLINS_NO_POLL") != nullptr, debug_knn = getenv("FFLINS_DEBUG_KNN") != nullptr;
        // (no reference = no kernel to wait for: the counts are then adopted the ordinary way, behind a synchronisation)
        late_q = !refs.empty() && !need_bg && cur->pending_slot >= 0 && cur->d_ndown && cur->d_ndown_gen == ctx->scratch_gen &&
                 cur->c[FFLINS_CLOUD_LIDAR_FULL].n > 0 && !ctx->prof_on && !no_poll && !debug_knn;
        if (!late_q && (rcp = resolve_pending(ctx, need_bg))) return rcp;
    }
    const int q    = late_q ? cur->c[FFLINS_CLOUD_LIDAR_FULL].n : cur->c[FFLINS_CLOUD_LIDAR].n;
    AssocParams P{};
    fill_assoc_common(ctx, P);
    P.n_refs = (int) refs.size();
    P.q      = q;
    P.q_dev  = late_q ? cur->d_ndown : nullptr;
    int rc;
    for (size_t i = 0; i < refs.size(); i++) {
        Frame *r = refs[i];
        if (r->hash_pending || r->hash_early) { // map built on the map stream
            if (ctx->map_job_active && ctx->map_job_key == r->id && (rc = resolve_map(ctx))) return rc;
            if (r->hash_pending) { // no index yet (or its size guess failed): insert it now, at first use as a reference
                if ((rc = build_hash(ctx, r))) return rc;
                r->hash_pending = false;
            }
        }
        if ((rc = ensure_features(ctx, r, q))) return rc;
        if (!r->index.fine && r->c[FFLINS_CLOUD_MAP].n > 0) return fail(ctx, FFLINS_E_INVALID, "reference frame has no map index (fflins_keyframe_add not called)");
        AssocRef &a   = P.ref[i];
        a.map         = r->c[FFLINS_CLOUD_MAP].d;
        a.index       = r->index;
        a.m           = r->c[FFLINS_CLOUD_MAP].n;
        a.T_ref_world = ref_world(r->pose);
        a.feat        = r->feat;
        a.counts      = ctx->d_assoc_counts + 2 * i;
        r->feat_q     = q;
    }
    long long *dbg = nullptr;
    static const bool debug_knn_stats = getenv("FFLINS_DEBUG_KNN") != nullptr;
    if (debug_knn_stats) {
        dbg = (long long *) ctx->pool.alloc((size_t) P.n_refs * q * 8 + 8);
        P.dbg_cycles = dbg;
    }
    // counts and completion flag come back through pinned memory (published by the last CTA of the plane kernel)
    unsigned long long *flag = ctx->h_signal + kMaxRefs + 1;
    int *h = (int *) (ctx->h_signal + kMaxRefs + 16);
    P.ticket      = ctx->d_assoc_counts + 2 * kMaxRefs;
    P.host_counts = h;
    P.signal      = flag;
    P.seq         = ++ctx->signal_seq;
    const bool launched = q > 0 && !refs.empty();
    {
        ProfScope ps(ctx->prof(), "associate", s);
        launch_associate(s, P, cur->c[FFLINS_CLOUD_WORLD].d, cur->c[FFLINS_CLOUD_LIDAR].d);
        if (launched) ctx->launches += 2;
    }
    CK(cudaGetLastError());
    if (launched) {
        static const bool no_poll = getenv("FFLINS_NO_POLL") != nullptr;
        if (ctx->prof_on || no_poll || dbg) CK(cudaStreamSynchronize(s));
        else CK(wait_signal(flag, 1, P.seq, s));
        ctx->eval_inputs_settled = true; // the flag is released behind the whole chain: frame chain, search, plane fit
    }
    if (dbg) {
        std::vector<long long> cyc((size_t) P.n_refs * q);
        cudaMemcpy(cyc.data(), dbg, cyc.size() * 8, cudaMemcpyDeviceToHost);
        ctx->pool.release(dbg);
        for (int r = 0; r < P.n_refs; r++) {
            std::vector<long long> v(cyc.begin() + (size_t) r * q, cyc.begin() + (size_t) (r + 1) * q);
            std::sort(v.begin(), v.end());
            long long sum = 0;
            for (auto x : v) sum += x;
            if (!v.empty())
                std::fprintf(stderr, "knn ref %d: cycles/warp min %lld p50 %lld p90 %lld p99 %lld max %lld mean %lld\n", r, v.front(),
                             v[v.size() / 2], v[v.size() * 9 / 10], v[v.size() * 99 / 100], v.back(), sum / (long long) v.size());
        }
    }
    for (size_t i = 0; i < refs.size(); i++) {
        nf[i] = launched ? h[2 * i] : 0;
        nc[i] = launched ? h[2 * i + 1] : 0;
    }
    if (late_q) {
        // the completion flag was released behind the whole chain: the pinned count slots are final, adopt them
        // without a stream synchronisation
        std::vector<uint64_t> keep;
        for (uint64_t id : ctx->pending) {
            Frame *f = find_frame(ctx, id);
            if (!f || f->pending_slot < 0) continue;
            if (f->bg) {
                keep.push_back(id);
                continue;
            }
            const int *cnt             = ctx->h_slots + 2 * f->pending_slot;
            f->n_fallback              = cnt[0];
            f->c[FFLINS_CLOUD_LIDAR].n = cnt[1];
            f->c[FFLINS_CLOUD_WORLD].n = cnt[1];
            f->pending_slot            = -1;
        }
        ctx->pending.swap(keep);
        for (Frame *r : refs) r->feat_q = cur->c[FFLINS_CLOUD_LIDAR].n;
    }
    return FFLINS_OK;
}

void unpack_red(const double *red, double *H, double *b, double *cost, int32_t *n_res) {
    if (H) {
        int e = 0;
        for (int i = 0; i < 19; i++)
            for (int j = i; j < 19; j++) {
                H[19 * i + j] = red[e];
                H[19 * j + i] = red[e];
                e++;
            }
    }
    if (b)
        for (int i = 0; i < 19; i++) b[i] = red[190 + i];
    if (cost) {
        cost[0] = 0.5 * red[209];
        cost[1] = red[210];
    }
    if (n_res) *n_res = (int32_t) std::llround(red[211]);
}

int fill_factor_pairs(fflins_ctx *ctx, FactorParams &P, int n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                      Frame *cur, const double *pose_cur, const double *ext, double td, uint32_t flags) {
    REQUIRE(n_pairs >= 0 && n_pairs <= kMaxPairs, "too many pairs (a window holds at most 15 reference keyframes)");
    {
        int rcp = flush_frames(ctx);
        if (rcp) return rcp;
        if ((rcp = resolve_pending(ctx, cur->bg))) return rcp;
    }
    P.n_pairs = n_pairs;
    P.q       = cur->c[FFLINS_CLOUD_LIDAR].n;
    // copies of the batch's background frames may still be crossing PCIe: the resident kernel polls with one warp only
    static const int force_quiet = getenv("FFLINS_QUIET") ? atoi(getenv("FFLINS_QUIET")) : -1;   // experiments: 0 never, 1 always
    const bool quiet = force_quiet >= 0 ? force_quiet != 0 : (ctx->bg_active && ctx->bg_staged);
    P.flags   = (flags & ~kFlagQuietNext) | (quiet ? kFlagQuietNext : 0u);
    std::memcpy(P.pose_cur, pose_cur, 7 * sizeof(double));
    std::memcpy(P.ext, ext, 7 * sizeof(double));
    P.td = td;
    std::memcpy(P.v1, cur->v, sizeof(P.v1));
    std::memcpy(P.w1, cur->w, sizeof(P.w1));
    P.td1         = cur->td;
    P.sigma       = ctx->cfg.point_to_plane_std;
    P.huber_delta = ctx->cfg.huber_delta;
    for (int i = 0; i < n_pairs; i++) {
        Frame *r = find_frame(ctx, ref_ids[i]);
        if (!r) return fail(ctx, FFLINS_E_NOTFOUND, "unknown ref frame");
        if (!r->feat.type || r->feat_q != P.q)
            return fail(ctx, FFLINS_E_INVALID, "ref frame holds no features for the current frame (associate first)");
        FactorPair &p = P.pair[i];
        p.type        = r->feat.type;
        p.outlier     = r->feat.outlier;
        p.normal      = r->feat.normal;
        p.d           = r->feat.d;
        p.std         = nullptr;
        std::memcpy(p.pose_ref, pose_refs + 7 * i, 7 * sizeof(double));
        std::memcpy(p.v0, r->v, sizeof(p.v0));
        std::memcpy(p.w0, r->w, sizeof(p.w0));
        p.td0 = r->td;
    }
    return FFLINS_OK;
}

} // namespace

extern "C" {

const char *fflins_version(void) { return "fflins-b200 0.1 (sm_100a)"; }

void fflins_config_default(fflins_config *c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->point_to_plane_std           = 0.1;
    c->window_size                  = 10;
    c->point_filter_num             = 10;
    c->downsample_size              = 0.5;
    c->frame_rate                   = 10;
    c->plane_estimation_threshold   = 0.1;
    c->minimum_keyframe_translation = 0.4;
    c->max_search_square_distance   = 5.0;
    c->max_keyframe_interval        = 0.5 * 0.95;
    c->min_keyframe_rotation        = 10.0 * M_PI / 180.0;
    c->plane_scalar_threshold       = 0.1;
    c->plane_sigma_gate             = 4.5;
    c->huber_delta                  = 1.0;
    c->max_points_per_frame         = 0;
    c->max_merge_frames             = 0;
}

int fflins_create(const fflins_config *cfg, int device, fflins_ctx **out) {
    if (!cfg || !out) return FFLINS_E_INVALID;
    *out      = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return FFLINS_E_NODEVICE;
    if (cfg->point_filter_num < 1 || cfg->downsample_size <= 0 || cfg->point_to_plane_std <= 0 ||
        cfg->window_size < 1 || cfg->window_size >= kMaxRefs)
        return FFLINS_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return FFLINS_E_NODEVICE;
    fflins_ctx *ctx = new fflins_ctx();
    ctx->cfg        = *cfg;
    ctx->device     = device;
    if (const char *e = getenv("FFLINS_CELL_SCALE")) {
        float v = (float) atof(e);
        if (v >= 0.5f && v <= 8.f) ctx->cell_scale = v;
    }
    if (!(cfg->max_search_square_distance > 0) || search_r_max(cfg->max_search_square_distance, search_cell(ctx)) > kMaxSearchCells) {
        delete ctx; // leaf too small for the search radius: the packed queue fields of the kNN would wrap
        return FFLINS_E_INVALID;
    }
    // the solver-facing calls (association, factor evaluation) are latency-critical and run on the main stream;
    // the keyframe-map rebuild on the map stream is throughput work that must not delay their CTAs
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->bg_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->map_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_main_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_bg_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_bg_copied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join_c, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_flush, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join_a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join_b, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_map_done, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return FFLINS_E_CUDA;
    }
    ctx->d_counters   = (int *) ctx->pool.alloc(64 * sizeof(int));
    ctx->d_fscratch   = (int *) ctx->pool.alloc(kFrameBatch * kItemScratch * sizeof(int));
    ctx->d_red        = (double *) ctx->pool.alloc((size_t) kMaxRefs * kRedSize * sizeof(double));
    ctx->d_map_count  = (int *) ctx->pool.alloc(64);
    ctx->d_assoc_counts = (int *) ctx->pool.alloc((kMaxRefs * 2 + 2) * sizeof(int));   // counts, ticket, task counter
    if (!ctx->d_counters || !ctx->d_fscratch || !ctx->d_red || !ctx->d_map_count || !ctx->d_assoc_counts) {
        fflins_destroy(ctx);
        return FFLINS_E_NOMEM;
    }
    int n = cfg->max_points_per_frame > 0 ? cfg->max_points_per_frame : 262144;
    int k = cfg->max_merge_frames > 0 ? cfg->max_merge_frames : 8;
    if (ensure_voxel(ctx, n) != FFLINS_OK || ensure_voxel_work(ctx, ctx->vox_map, ctx->vox_map_base, n * k) != FFLINS_OK || ensure_stage(ctx, (size_t) n * 32) != FFLINS_OK ||
        ensure_ins(ctx, 4096) != FFLINS_OK || !pin(ctx, 1 << 20)) {
        fflins_destroy(ctx);
        return FFLINS_E_NOMEM;
    }
    if (cudaMallocHost((void **) &ctx->h_slots, kSlots * 2 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost((void **) &ctx->h_signal, (5 * kMaxRefs + 16) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMallocHost((void **) &ctx->h_map_count, 64) != cudaSuccess) {
        ctx->h_slots = nullptr;
        fflins_destroy(ctx);
        return FFLINS_E_NOMEM;
    }
    std::memset(ctx->h_signal, 0, (5 * kMaxRefs + 16) * sizeof(unsigned long long));
    ctx->frt = factor_runtime_create(device);
    if (!ctx->frt) {
        fflins_destroy(ctx);
        return FFLINS_E_NOMEM;
    }
    cudaMemsetAsync(ctx->d_assoc_counts, 0, (kMaxRefs * 2 + 2) * sizeof(int), ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        fflins_destroy(ctx);
        return FFLINS_E_CUDA;
    }
    if (!getenv("FFLINS_NO_HELPER_THREAD")) {
        ctx->map_worker = std::thread([ctx, device] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(ctx->map_mu);
            while (true) {
                ctx->map_cv.wait(lk, [ctx] { return ctx->map_quit || (bool) ctx->map_task; });
                if (ctx->map_task) {
                    std::function<void()> task = std::move(ctx->map_task);
                    ctx->map_task              = nullptr;
                    lk.unlock();
                    task();
                    lk.lock();
                    ctx->map_issued = true;
                    ctx->map_cv.notify_all();
                } else if (ctx->map_quit) {
                    return;
                }
            }
        });
    }
    *out = ctx;
    return FFLINS_OK;
}

void fflins_destroy(fflins_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->map_worker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(ctx->map_mu);
            ctx->map_quit = true;
            ctx->map_cv.notify_all();
        }
        ctx->map_worker.join();
    }
    ctx->copy_pool.stop();
    factor_runtime_destroy(ctx->frt); // retires the resident evaluation kernel first
    ctx->frt = nullptr;
    if (ctx->map_stream) cudaStreamSynchronize(ctx->map_stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->bg_stream) cudaStreamSynchronize(ctx->bg_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (Frame *f : ctx->deferred_frames) delete f;
    ctx->deferred_frames.clear();
    for (auto &kv : ctx->frames) {
        Frame *f = kv.second;
        delete f;
    }
    ctx->frames.clear();
    ctx->pool.destroy();
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
    for (auto e : ctx->bounce_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
    if (ctx->h_map_count) cudaFreeHost(ctx->h_map_count);
    if (ctx->h_signal) cudaFreeHost(ctx->h_signal);
    if (ctx->ev_main_done) cudaEventDestroy(ctx->ev_main_done);
    if (ctx->ev_map_done) cudaEventDestroy(ctx->ev_map_done);
    if (ctx->map_stream) cudaStreamDestroy(ctx->map_stream);
    if (ctx->h_win) cudaFreeHost(ctx->h_win);
    for (auto e : ctx->batch_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
    if (ctx->ev_bg_done) cudaEventDestroy(ctx->ev_bg_done);
    if (ctx->ev_bg_copied) cudaEventDestroy(ctx->ev_bg_copied);
    if (ctx->ev_join_c) cudaEventDestroy(ctx->ev_join_c);
    if (ctx->ev_flush) cudaEventDestroy(ctx->ev_flush);
    if (ctx->ev_join_a) cudaEventDestroy(ctx->ev_join_a);
    if (ctx->ev_join_b) cudaEventDestroy(ctx->ev_join_b);
    for (auto &r : ctx->prof_open) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : ctx->ev_free) cudaEventDestroy(e);
    if (ctx->timer_a) cudaEventDestroy(ctx->timer_a);
    if (ctx->timer_b) cudaEventDestroy(ctx->timer_b);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->bg_stream) cudaStreamDestroy(ctx->bg_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *fflins_last_error(const fflins_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fflins_synchronize(fflins_ctx *ctx) {
    if (!ctx) return FFLINS_E_INVALID;
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    {
        int rcf = flush_frames(ctx);
        if (rcf) return rcf;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return resolve_pending(ctx);
}

int fflins_host_register(void *ptr, uint64_t bytes) {
    if (!ptr || !bytes) return FFLINS_E_INVALID;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? FFLINS_OK : FFLINS_E_CUDA;
}
int fflins_host_unregister(void *ptr) {
    if (!ptr) return FFLINS_E_INVALID;
    return cudaHostUnregister(ptr) == cudaSuccess ? FFLINS_OK : FFLINS_E_CUDA;
}

int fflins_host_buffer_done(fflins_ctx *ctx, const void *ptr) {
    if (!ctx) return FFLINS_E_INVALID;
    bool queued = false;
    for (const QueuedFrame &q : ctx->fq)
        for (int c = 0; c < q.n_copy; c++) queued |= q.h_src[c] == ptr;
    if (queued) {
        int rc = flush_frames(ctx);   // its copy has not even been issued yet
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(ctx->copy_stream));   // every host-to-device copy of point data runs on the copy stream
    return FFLINS_OK;
}

int fflins_host_prefetch(fflins_ctx *ctx, const void *points32, int32_t n) {
    if (!ctx || !points32 || n <= 0) return FFLINS_E_INVALID;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, points32) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) return FFLINS_OK;   // pageable buffers are staged at call time anyway: nothing to start early
    for (const QueuedFrame &q : ctx->fq)
        for (int c = 0; c < q.n_copy; c++)
            if (q.h_src[c] == points32) return FFLINS_OK;   // already queued
    const size_t bytes = (size_t) n * 32;
    for (size_t a = 0; a < ctx->announced.size(); a++)
        if (ctx->announced[a].host == points32) { // announced again (same buffer): the older announcement is dropped
            ctx->stage_guard[ctx->announced[a].slot] = 0;
            ctx->announced.erase(ctx->announced.begin() + a);
            break;
        }
    if ((int) ctx->announced.size() >= kMaxAnnounced) { // at most half the staging ring is held by announcements
        ctx->stage_guard[ctx->announced.front().slot] = 0;
        ctx->announced.erase(ctx->announced.begin());
    }
    char *pts;
    int slot, rc;
    if ((rc = stage_acquire(ctx, bytes + 256, &pts, &slot))) return rc;
    CK(cudaMemcpyAsync(pts, points32, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->announced.push_back({points32, bytes, slot, pts});
    return FFLINS_OK;
}

int fflins_frame_process(fflins_ctx *ctx, uint64_t frame_id, const float *xyzi, const double *time, int32_t n,
                         const double *ins_t, const double *ins_p, const double *ins_q, int32_t w,
                         const fflins_pose *pose_b_l, double stamp, double td, fflins_frame_info *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || (xyzi && time)) && ins_t && ins_p && ins_q && pose_b_l, "null argument");
    return enqueue_frame(ctx, frame_id, xyzi, time, nullptr, false, n, ins_t, ins_p, ins_q, w, pose_b_l, stamp, td, out);
}

int fflins_frame_process_aos(fflins_ctx *ctx, uint64_t frame_id, const void *points32, int32_t n, const double *ins_t,
                             const double *ins_p, const double *ins_q, int32_t w, const fflins_pose *pose_b_l,
                             double stamp, double td, fflins_frame_info *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || points32) && ins_t && ins_p && ins_q && pose_b_l, "null argument");
    return enqueue_frame(ctx, frame_id, nullptr, nullptr, points32, false, n, ins_t, ins_p, ins_q, w, pose_b_l, stamp, td, out);
}

int fflins_frame_process_dev(fflins_ctx *ctx, uint64_t frame_id, const float *d_xyzi, const double *d_time, int32_t n,
                             const double *ins_t, const double *ins_p, const double *ins_q, int32_t w,
                             const fflins_pose *pose_b_l, double stamp, double td, fflins_frame_info *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || (d_xyzi && d_time)) && ins_t && ins_p && ins_q && pose_b_l, "null argument");
    return enqueue_frame(ctx, frame_id, d_xyzi, d_time, nullptr, true, n, ins_t, ins_p, ins_q, w, pose_b_l, stamp, td, out);
}

int fflins_frame_process_static(fflins_ctx *ctx, uint64_t frame_id, const float *xyzi, int32_t n, const double state_p[3],
                                const double state_q[4], const fflins_pose *pose_b_l, fflins_frame_info *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || xyzi) && state_p && state_q && pose_b_l, "null argument");
    int rc;
    if ((rc = flush_frames(ctx))) return rc;
    ctx->eval_inputs_settled = false;
    cudaStream_t s = ctx->stream;
    Frame *f       = get_frame(ctx, frame_id);
    const int F    = ctx->cfg.point_filter_num;
    const int ndec = (n + F - 1) / F;
    FrameBatch B{};
    B.count       = 1;
    FrameItem &it = B.it[0];
    if ((rc = ensure_ins(ctx, 2))) return rc;
    ctx->scratch_gen++;
    item_scratch(ctx, 0, it);
    // MISC::stateToPoseWithExtrinsic (misc.cc:99-105), host fp64
    {
        Q4 q{state_q[0], state_q[1], state_q[2], state_q[3]};
        M3 R = quat_to_R(q);
        for (int i = 0; i < 3; i++)
            f->pose.t[i] = state_p[i] + ((R.m[3 * i] * pose_b_l->t[0] + R.m[3 * i + 1] * pose_b_l->t[1]) + R.m[3 * i + 2] * pose_b_l->t[2]);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                f->pose.R[3 * i + j] = (R.m[3 * i] * pose_b_l->R[j] + R.m[3 * i + 1] * pose_b_l->R[3 + j]) + R.m[3 * i + 2] * pose_b_l->R[6 + j];
    }
    if ((rc = frame_begin(ctx, f, n, ndec, &it.report))) return rc;
    it.frame    = to12(f->pose);
    it.out_dec  = f->c[FFLINS_CLOUD_LIDAR_FULL].d;
    it.ndec     = ndec;
    it.inv_leaf = 1.0f / (float) ctx->cfg.downsample_size;
    it.down     = f->c[FFLINS_CLOUD_LIDAR].d;
    it.world    = f->c[FFLINS_CLOUD_WORLD].d;
    CK(cudaMemsetAsync(it.counters, 0, 4 * sizeof(int), s));
    if (n > 0) {
        char *stage;
        int slot;
        if ((rc = stage_acquire(ctx, (size_t) n * 16, &stage, &slot))) return rc;
        CK(cudaMemcpyAsync(stage, xyzi, (size_t) n * 16, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
        CK(cudaStreamWaitEvent(s, ctx->ev_copied, 0));
        {
            ProfScope ps(ctx->prof(), "copy_decimate", s);
            launch_copy_decimate(s, (const float4 *) stage, n, F, f->c[FFLINS_CLOUD_MAP_FULL].d,
                                 f->c[FFLINS_CLOUD_LIDAR_FULL].d);
            ctx->launches++;
        }
        const unsigned long long seq = next_batch(ctx);
        CK(cudaEventRecord(ctx->batch_ev[seq % kGuardRing], s));
        ctx->stage_guard[slot] = seq;
    }
    if ((rc = frame_voxel_tail(ctx, s, B))) return rc;
    if (!out) return FFLINS_OK;
    if ((rc = resolve_pending(ctx))) return rc;
    fill_info(f, out);
    return FFLINS_OK;
}

int fflins_frame_info_get(fflins_ctx *ctx, uint64_t frame_id, fflins_frame_info *out) {
    if (!ctx || !out) return FFLINS_E_INVALID;
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    {
        int rcp = resolve_pending(ctx);
        if (rcp) return rcp;
    }
    Frame *f = find_frame(ctx, frame_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    fill_info(f, out);
    return FFLINS_OK;
}

int fflins_frame_download(fflins_ctx *ctx, uint64_t frame_id, int which, float *xyzi, int32_t cap, int32_t *n) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(which >= 0 && which < 5, "bad cloud selector");
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    {
        int rcp = resolve_pending(ctx);
        if (rcp) return rcp;
    }
    Frame *f = find_frame(ctx, frame_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    const Cloud &c = f->c[which];
    if (n) *n = c.n;
    if (!xyzi) return FFLINS_OK;
    if (cap < c.n) return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
    if (c.n > 0) {
        CK(cudaMemcpyAsync(xyzi, c.d, (size_t) c.n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return FFLINS_OK;
}

int fflins_frame_upload(fflins_ctx *ctx, uint64_t frame_id, int which, const float *xyzi, int32_t n) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(which >= 0 && which < 5 && n >= 0 && (n == 0 || xyzi), "bad argument");
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    {
        int rcp = resolve_pending(ctx);
        if (rcp) return rcp;
    }
    Frame *f = get_frame(ctx, frame_id);
    int rc;
    if ((rc = ensure_cloud(ctx, f->c[which], n))) return rc;
    if (n > 0) {
        CK(cudaMemcpyAsync(f->c[which].d, xyzi, (size_t) n * 16, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    f->c[which].n = n;
    if (which == FFLINS_CLOUD_MAP_FULL) f->n_raw = n;
    return FFLINS_OK;
}

int fflins_frame_set_pose(fflins_ctx *ctx, uint64_t frame_id, const fflins_pose *pose) {
    if (!ctx || !pose) return FFLINS_E_INVALID;
    Frame *f = get_frame(ctx, frame_id);
    f->pose  = *pose;
    return FFLINS_OK;
}

int fflins_frame_get_pose(fflins_ctx *ctx, uint64_t frame_id, fflins_pose *pose) {
    if (!ctx || !pose) return FFLINS_E_INVALID;
    Frame *f = find_frame(ctx, frame_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    *pose = f->pose;
    return FFLINS_OK;
}

int fflins_frame_set_motion(fflins_ctx *ctx, uint64_t frame_id, const double v[3], const double omega[3], double td) {
    if (!ctx || !v || !omega) return FFLINS_E_INVALID;
    Frame *f = get_frame(ctx, frame_id);
    std::memcpy(f->v, v, sizeof(f->v));
    std::memcpy(f->w, omega, sizeof(f->w));
    f->td = td;
    return FFLINS_OK;
}

int fflins_frame_release(fflins_ctx *ctx, uint64_t frame_id) {
    if (!ctx) return FFLINS_E_INVALID;
    {
        int rcf = flush_frames(ctx); // a queued frame must be launched before its buffers go back to the pool
        if (rcf) return rcf;
    }
    auto it = ctx->frames.find(frame_id);
    if (it == ctx->frames.end()) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    for (auto k : ctx->keyframes)
        if (k == frame_id) return fail(ctx, FFLINS_E_INVALID, "frame is a keyframe: remove it from the window first");
    // no synchronisation: the blocks go back to the pool and any reuse is ordered on the context's stream;
    // while a map job is in flight on the map stream the release is deferred until the job is done
    if (ctx->map_job_active) {
        if (ctx->map_job_key == frame_id) {
            int rcm = resolve_map(ctx);
            if (rcm) return rcm;
            free_frame(ctx, it->second);
        } else {
            ctx->deferred_frames.push_back(it->second);
        }
    } else {
        // its chain may still be running on the copy stream: order the compute stream (where the blocks are reused) behind it
        if (it->second->bg) {
            int rcj = join_bg(ctx);
            if (rcj) return rcj;
        }
        free_frame(ctx, it->second);
    }
    ctx->frames.erase(it);
    return FFLINS_OK;
}

int fflins_reproject(fflins_ctx *ctx, uint64_t frame_id, const fflins_pose *pose) {
    if (!ctx || !pose) return FFLINS_E_INVALID;
    {
        int rcp = resolve_pending(ctx);
        if (rcp) return rcp;
    }
    Frame *f = find_frame(ctx, frame_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    f->pose = *pose;
    int n   = f->c[FFLINS_CLOUD_LIDAR].n;
    int rc;
    if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_WORLD], n))) return rc;
    {
        ProfScope ps(ctx->prof(), "reproject", ctx->stream);
        launch_project(ctx->stream, to12(*pose), f->c[FFLINS_CLOUD_LIDAR].d, f->c[FFLINS_CLOUD_WORLD].d, n, nullptr);
        if (n > 0) ctx->launches++;
    }
    f->c[FFLINS_CLOUD_WORLD].n = n;
    return FFLINS_OK;
}

int fflins_reproject_window(fflins_ctx *ctx, int32_t n, const uint64_t *frame_ids, const fflins_pose *poses) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && n <= kMaxRefs && (n == 0 || (frame_ids && poses)), "bad argument");
    {
        int rcp = resolve_pending(ctx);
        if (rcp) return rcp;
    }
    ProjectBatch B{};
    B.count = n;
    int rc;
    for (int i = 0; i < n; i++) {
        Frame *f = find_frame(ctx, frame_ids[i]);
        if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
        f->pose = poses[i];
        int m   = f->c[FFLINS_CLOUD_LIDAR].n;
        if ((rc = ensure_cloud(ctx, f->c[FFLINS_CLOUD_WORLD], m))) return rc;
        f->c[FFLINS_CLOUD_WORLD].n = m;
        B.n[i]   = m;
        B.src[i] = f->c[FFLINS_CLOUD_LIDAR].d;
        B.dst[i] = f->c[FFLINS_CLOUD_WORLD].d;
        B.T[i]   = to12(poses[i]);
    }
    ProfScope ps(ctx->prof(), "reproject", ctx->stream);
    launch_project_batch(ctx->stream, B);
    ctx->launches++;
    return FFLINS_OK;
}

int fflins_keyframe_selection(fflins_ctx *ctx, uint64_t frame_id, double frame_stamp, double last_key_stamp,
                              int *is_keyframe, double stats[3]) {
    if (!ctx || !is_keyframe) return FFLINS_E_INVALID;
    Frame *f = find_frame(ctx, frame_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    REQUIRE(!ctx->keyframes.empty(), "no keyframe in the window");
    Frame *k = find_frame(ctx, ctx->keyframes.back());
    // lidar_map.cc:77-98: interval, translation, rotation angle of R_f^T R_k
    double interval = frame_stamp - last_key_stamp;
    double dt[3]    = {f->pose.t[0] - k->pose.t[0], f->pose.t[1] - k->pose.t[1], f->pose.t[2] - k->pose.t[2]};
    double trans    = std::sqrt(dt[0] * dt[0] + dt[1] * dt[1] + dt[2] * dt[2]);
    double Rr[9];
    mat_tn(f->pose.R, k->pose.R, Rr);
    // AngleAxisd(R): via quaternion; angle = 2 atan2(|vec|, |w|)
    double tr = Rr[0] + Rr[4] + Rr[8];
    double vx = Rr[7] - Rr[5], vy = Rr[2] - Rr[6], vz = Rr[3] - Rr[1];
    double sn = 0.5 * std::sqrt(vx * vx + vy * vy + vz * vz), cs = 0.5 * (tr - 1.0);
    double rot = std::fabs(std::atan2(sn, cs));
    if (stats) {
        stats[0] = interval;
        stats[1] = trans;
        stats[2] = rot;
    }
    *is_keyframe = (interval > ctx->cfg.max_keyframe_interval) || (trans > ctx->cfg.minimum_keyframe_translation) ||
                   (rot > ctx->cfg.min_keyframe_rotation);
    return FFLINS_OK;
}

int fflins_voxel_downsample(fflins_ctx *ctx, const float *xyzi, int32_t n, double leaf, float *out_xyzi, int32_t cap,
                            int32_t *n_out, int32_t *keys) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || xyzi) && n_out && leaf > 0, "bad argument");
    if (n == 0) {
        *n_out = 0;
        return FFLINS_OK;
    }
    int rc;
    if ((rc = ensure_voxel(ctx, n))) return rc;
    cudaStream_t s = ctx->stream;
    float4 *d_in   = (float4 *) ctx->pool.alloc((size_t) n * 16);
    float4 *d_out  = (float4 *) ctx->pool.alloc((size_t) n * 16);
    int32_t *d_keys = keys ? (int32_t *) ctx->pool.alloc((size_t) n * 4) : nullptr;
    auto cleanup = [&]() {
        ctx->pool.release(d_in);
        ctx->pool.release(d_out);
        ctx->pool.release(d_keys);
    };
    if (!d_in || !d_out || (keys && !d_keys)) {
        cleanup();
        return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    }
    cudaMemcpyAsync(d_in, xyzi, (size_t) n * 16, cudaMemcpyHostToDevice, s);
    int64_t nl = 0;
    launch_voxel_downsample(s, ctx->vox, d_in, n, (float) leaf, d_out, ctx->d_counters + 2, d_keys, &nl, ctx->prof(), "voxel");
    ctx->launches += nl;
    int m = 0;
    cudaMemcpyAsync(&m, ctx->d_counters + 2, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (keys) cudaMemcpyAsync(keys, d_keys, (size_t) n * 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        cleanup();
        return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    }
    *n_out = m;
    if (out_xyzi) {
        if (cap < m) {
            cleanup();
            return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
        }
        cudaMemcpyAsync(out_xyzi, d_out, (size_t) m * 16, cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    cleanup();
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    return FFLINS_OK;
}

int fflins_cloud_projection(fflins_ctx *ctx, const fflins_pose *pose, const float *src, int32_t n, float *dst) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(pose && n >= 0 && (n == 0 || (src && dst)), "bad argument");
    if (n == 0) return FFLINS_OK;
    cudaStream_t s = ctx->stream;
    float4 *d      = (float4 *) ctx->pool.alloc((size_t) n * 16);
    if (!d) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    cudaMemcpyAsync(d, src, (size_t) n * 16, cudaMemcpyHostToDevice, s);
    launch_project(s, to12(*pose), d, d, n, nullptr);
    ctx->launches++;
    cudaMemcpyAsync(dst, d, (size_t) n * 16, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    ctx->pool.release(d);
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    return FFLINS_OK;
}

int fflins_plane_estimation(fflins_ctx *ctx, const float *pts, int32_t n, double threshold, double *unit_normal,
                            double *norm_inverse, double *distance, uint8_t *is_plane) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || (pts && unit_normal && norm_inverse && distance && is_plane)), "bad argument");
    if (n == 0) return FFLINS_OK;
    cudaStream_t s = ctx->stream;
    size_t bytes   = (size_t) n * (80 + 24 + 8 + 40 + 8);
    char *d        = (char *) ctx->pool.alloc(bytes);
    if (!d) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    float4 *d_pts  = (float4 *) d;
    double *d_unit = (double *) (d + (size_t) n * 80);
    double *d_ninv = d_unit + 3 * (size_t) n;
    double *d_dist = d_ninv + n;
    uint8_t *d_ok  = (uint8_t *) (d_dist + 5 * (size_t) n);
    cudaMemcpyAsync(d_pts, pts, (size_t) n * 80, cudaMemcpyHostToDevice, s);
    launch_plane_estimation(s, d_pts, n, threshold, d_unit, d_ninv, d_dist, d_ok);
    ctx->launches++;
    cudaMemcpyAsync(unit_normal, d_unit, (size_t) n * 24, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(norm_inverse, d_ninv, (size_t) n * 8, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(distance, d_dist, (size_t) n * 40, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(is_plane, d_ok, (size_t) n, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    ctx->pool.release(d);
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    return FFLINS_OK;
}

int fflins_knn5(fflins_ctx *ctx, const float *map_xyzi, int32_t m, const float *query_xyzi, int32_t q, double max_dist,
                int32_t *nn_idx, float *nn_d2, int32_t *nn_count) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(m >= 0 && q >= 0 && (m == 0 || map_xyzi) && (q == 0 || (query_xyzi && nn_idx && nn_d2 && nn_count)) && max_dist > 0,
            "bad argument");
    if (q == 0) return FFLINS_OK;
    REQUIRE(search_r_max(max_dist, search_cell(ctx)) <= kMaxSearchCells, "max_dist too large for the search cell (more than 495 cells)");
    cudaStream_t s = ctx->stream;
    int hcap       = next_pow2(2 * std::max(m, 1));
    float4 *d_map  = (float4 *) ctx->pool.alloc((size_t) std::max(m, 1) * 16);
    float4 *d_q    = (float4 *) ctx->pool.alloc((size_t) q * 16);
    MapIndex X;
    X.fine   = (FineSlot *) ctx->pool.alloc((size_t) hcap * sizeof(FineSlot));
    X.coarse = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
    X.super  = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
    X.hmask  = hcap - 1;
    int32_t *d_idx = (int32_t *) ctx->pool.alloc((size_t) q * 20);
    float *d_d2    = (float *) ctx->pool.alloc((size_t) q * 20);
    int32_t *d_cnt = (int32_t *) ctx->pool.alloc((size_t) q * 4);
    auto cleanup = [&]() {
        ctx->pool.release(d_map);
        ctx->pool.release(d_q);
        ctx->pool.release(X.fine);
        ctx->pool.release(X.coarse);
        ctx->pool.release(X.super);
        ctx->pool.release(d_idx);
        ctx->pool.release(d_d2);
        ctx->pool.release(d_cnt);
    };
    if (!d_map || !d_q || !X.fine || !X.coarse || !X.super || !d_idx || !d_d2 || !d_cnt) {
        cleanup();
        return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    }
    if (m > 0) cudaMemcpyAsync(d_map, map_xyzi, (size_t) m * 16, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_q, query_xyzi, (size_t) q * 16, cudaMemcpyHostToDevice, s);
    float cell = search_cell(ctx);
    launch_hash_build(s, d_map, m, 1.0f / cell, X);
    launch_knn5(s, X, m, d_q, q, 1.0f / cell, cell, (int) std::ceil(max_dist / cell) + 1,
                (float) (max_dist * max_dist), d_idx, d_d2, d_cnt);
    ctx->launches += 3;
    cudaMemcpyAsync(nn_idx, d_idx, (size_t) q * 20, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(nn_d2, d_d2, (size_t) q * 20, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(nn_count, d_cnt, (size_t) q * 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cleanup();
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    return FFLINS_OK;
}

int fflins_keyframe_merge(fflins_ctx *ctx, uint64_t key_id, const uint64_t *nonkey_ids, int32_t n, fflins_frame_info *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || nonkey_ids), "bad argument");
    int rc;
    g_merge_trace.start();
    if ((rc = flush_frames(ctx))) return rc; // the frames being merged may still be queued
    g_merge_trace.lap(0);
    if ((rc = resolve_map(ctx))) return rc;  // one map job in flight at a time
    g_merge_trace.lap(1);
    Frame *key = find_frame(ctx, key_id);
    if (!key) return fail(ctx, FFLINS_E_NOTFOUND, "unknown keyframe");
    cudaStream_t ms = ctx->map_stream;
    int total       = key->c[FFLINS_CLOUD_MAP_FULL].n;
    std::vector<Frame *> nk(n);
    for (int i = 0; i < n; i++) {
        nk[i] = find_frame(ctx, nonkey_ids[i]);
        if (!nk[i]) return fail(ctx, FFLINS_E_NOTFOUND, "unknown non-keyframe");
        total += nk[i]->c[FFLINS_CLOUD_MAP_FULL].n;
    }
    if ((rc = ensure_voxel_work(ctx, ctx->vox_map, ctx->vox_map_base, total))) return rc;
    ctx->merge_reserve = std::min(std::max(ctx->merge_reserve, n + 1), (int) kMaxRefs);
    // the map stream starts once everything enqueued so far on the compute stream (the undistortion of the
    // frames being merged) is done
    CK(cudaEventRecord(ctx->ev_main_done, ctx->stream));
    CK(cudaStreamWaitEvent(ms, ctx->ev_main_done, 0));
    if (ctx->bg_active) CK(cudaStreamWaitEvent(ms, ctx->ev_bg_done, 0)); // ... and the background part of the batch
    // grow the keyframe's full cloud to K*N on the map stream; blocks the job may still read are freed later
    Cloud &full = key->c[FFLINS_CLOUD_MAP_FULL];
    if (total > full.cap) {
        size_t want = Pool::round((size_t) std::max(total, 256) * sizeof(float4));
        float4 *p   = (float4 *) ctx->pool.alloc(want);
        if (!p) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
        if (full.d && full.n > 0) CK(cudaMemcpyAsync(p, full.d, (size_t) full.n * sizeof(float4), cudaMemcpyDeviceToDevice, ms));
        if (full.d) ctx->deferred_blocks.push_back(full.d);
        full.d   = p;
        full.cap = (int) (want / sizeof(float4));
    }
    Cloud &map = key->c[FFLINS_CLOUD_MAP];
    if (total > map.cap) {
        size_t want = Pool::round((size_t) std::max(total, 256) * sizeof(float4));
        float4 *p   = (float4 *) ctx->pool.alloc(want);
        if (!p) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
        if (map.d) ctx->deferred_blocks.push_back(map.d);
        map.d   = p;
        map.cap = (int) (want / sizeof(float4));
    }
    int off = full.n;
    std::vector<ProjectBatch> batches;
    for (int i0 = 0; i0 < n; i0 += kMaxRefs) {
        // relative pose non-keyframe -> keyframe (lidar_map.cc:198-199); the projected clouds are written
        // straight behind the keyframe's own points (the reference projects in place, then appends: :201-205),
        // up to 16 clouds per launch
        ProjectBatch B{};
        B.count = std::min(n - i0, (int) kMaxRefs);
        for (int j = 0; j < B.count; j++) {
            Frame *g = nk[i0 + j];
            mat_tn(key->pose.R, g->pose.R, B.T[j].R);
            double dt[3] = {g->pose.t[0] - key->pose.t[0], g->pose.t[1] - key->pose.t[1], g->pose.t[2] - key->pose.t[2]};
            mat_tv(key->pose.R, dt, B.T[j].t);
            B.n[j]   = g->c[FFLINS_CLOUD_MAP_FULL].n;
            B.src[j] = g->c[FFLINS_CLOUD_MAP_FULL].d;
            B.dst[j] = full.d + off;
            off += B.n[j];
        }
        batches.push_back(B);
    }
    full.n = total;
    // the keyframe's search index is inserted by the map job as well (off the association's critical path): its
    // tables are sized from the previous keyframe map, the real count is checked when the job is adopted
    bool early = false;
    MapIndex early_index{};
    if (!out && ctx->last_map_n > 0) {
        int hcap = next_pow2(2 * ctx->last_map_n);
        if (hcap > key->hcap) {
            if (key->index.fine) ctx->deferred_blocks.push_back(key->index.fine);
            if (key->index.coarse) ctx->deferred_blocks.push_back(key->index.coarse);
            if (key->index.super) ctx->deferred_blocks.push_back(key->index.super);
            key->index.fine   = (FineSlot *) ctx->pool.alloc((size_t) hcap * sizeof(FineSlot));
            key->index.coarse = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
            key->index.super  = (MaskSlot *) ctx->pool.alloc((size_t) hcap * sizeof(MaskSlot));
            if (!key->index.fine || !key->index.coarse || !key->index.super) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
            key->hcap = hcap;
        }
        key->index.hmask = key->hcap - 1;
        early            = true;
        early_index      = key->index;
    }
    key->hash_early = early;
    // everything below only issues CUDA work on the map stream with the arguments prepared above
    const float4 *in_ptr = full.d;
    float4 *out_ptr      = map.d;
    const bool profiling = ctx->prof_on;
    const float inv_cell = 1.0f / search_cell(ctx);
    auto issue = [ctx, ms, batches, in_ptr, out_ptr, total, profiling, early, early_index, inv_cell]() {
        int64_t nl = 0;
        for (const ProjectBatch &B : batches) {
            ProfScope ps(profiling ? ctx : nullptr, "merge_project", ms);
            launch_project_batch(ms, B);
            nl++;
        }
        // the cooperative filter stores the voxel count straight into the pinned word (rc 2); other paths copy it
        const int vrc = launch_voxel_downsample(ms, ctx->vox_map, in_ptr, total, (float) ctx->cfg.downsample_size, out_ptr,
                                                ctx->d_map_count, nullptr, &nl, profiling ? ctx : nullptr, "voxel_map", nullptr,
                                                ctx->h_map_count);
        if (vrc != 2) cudaMemcpyAsync(ctx->h_map_count, ctx->d_map_count, sizeof(int), cudaMemcpyDeviceToHost, ms);
        if (early) {
            ProfScope ps(profiling ? ctx : nullptr, "hash_build", ms);
            launch_hash_build(ms, out_ptr, total, inv_cell, early_index, ctx->d_map_count);
            nl += 2;
        }
        cudaEventRecord(ctx->ev_map_done, ms);
        ctx->launches += nl;
    };
    g_merge_trace.lap(2);
    if (out || profiling || !ctx->map_worker.joinable()) {
        issue(); // synchronous call, profiling pass (the event list is not thread-safe) or no helper thread
    } else {
        std::lock_guard<std::mutex> lk(ctx->map_mu);
        ctx->map_task   = issue;
        ctx->map_issued = false;
        ctx->map_cv.notify_all();
    }
    g_merge_trace.lap(3);
    g_merge_trace.done();
    ctx->map_job_active = true;
    ctx->map_job_key    = key_id;
    map.n               = 0;
    if (!out) return FFLINS_OK; // asynchronous: the map is adopted by the first call that needs it
    if ((rc = resolve_map(ctx))) return rc;
    if ((rc = resolve_pending(ctx))) return rc;
    fill_info(key, out);
    return FFLINS_OK;
}

int fflins_keyframe_add(fflins_ctx *ctx, uint64_t key_id) {
    if (!ctx) return FFLINS_E_INVALID;
    Frame *f = find_frame(ctx, key_id);
    if (!f) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    REQUIRE((int) ctx->keyframes.size() < kMaxRefs, "keyframe window is full");
    int rc;
    if (ctx->map_job_active && ctx->map_job_key == key_id) {
        // its map is still being built: the map job inserts the index itself (hash_early) or it happens at first
        // use as a reference
        f->hash_pending = !f->hash_early;
    } else if ((rc = build_hash(ctx, f))) {
        return rc;
    }
    f->is_key = true;
    ctx->keyframes.push_back(key_id);
    return FFLINS_OK;
}

static int remove_key(fflins_ctx *ctx, bool oldest, uint64_t *removed_id) {
    REQUIRE(!ctx->keyframes.empty(), "no keyframe in the window");
    {
        int rcf = flush_frames(ctx);
        if (rcf) return rcf;
        uint64_t victim = oldest ? ctx->keyframes.front() : ctx->keyframes.back();
        if (ctx->map_job_active && ctx->map_job_key == victim) {
            int rcm = resolve_map(ctx);
            if (rcm) return rcm;
        }
    }
    uint64_t id = oldest ? ctx->keyframes.front() : ctx->keyframes.back();
    if (oldest)
        ctx->keyframes.pop_front();
    else
        ctx->keyframes.pop_back();
    if (removed_id) *removed_id = id;
    auto it = ctx->frames.find(id);
    if (it != ctx->frames.end()) {
        if (it->second->bg) {
            int rcj = join_bg(ctx);
            if (rcj) return rcj;
        }
        free_frame(ctx, it->second);
        ctx->frames.erase(it);
    }
    return FFLINS_OK;
}

int fflins_keyframe_remove_oldest(fflins_ctx *ctx, uint64_t *removed_id) {
    if (!ctx) return FFLINS_E_INVALID;
    return remove_key(ctx, true, removed_id);
}
int fflins_keyframe_remove_newest(fflins_ctx *ctx, uint64_t *removed_id) {
    if (!ctx) return FFLINS_E_INVALID;
    return remove_key(ctx, false, removed_id);
}
int fflins_keyframe_count(fflins_ctx *ctx, int32_t *n) {
    if (!ctx || !n) return FFLINS_E_INVALID;
    *n = (int32_t) ctx->keyframes.size();
    return FFLINS_OK;
}
int fflins_keyframe_ids(fflins_ctx *ctx, uint64_t *ids, int32_t cap, int32_t *n) {
    if (!ctx || !n) return FFLINS_E_INVALID;
    *n = (int32_t) ctx->keyframes.size();
    if (ids) {
        if (cap < *n) return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
        for (size_t i = 0; i < ctx->keyframes.size(); i++) ids[i] = ctx->keyframes[i];
    }
    return FFLINS_OK;
}

int fflins_associate(fflins_ctx *ctx, fflins_assoc_stats *out) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(ctx->keyframes.size() >= 1, "no keyframe in the window");
    Frame *cur = find_frame(ctx, ctx->keyframes.back());
    std::vector<Frame *> refs;
    for (size_t i = 0; i + 1 < ctx->keyframes.size(); i++) refs.push_back(find_frame(ctx, ctx->keyframes[i]));
    int32_t nf[kMaxRefs] = {0}, nc[kMaxRefs] = {0};
    int rc = run_associate(ctx, refs, cur, nf, nc);
    if (rc) return rc;
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->n_refs    = (int32_t) refs.size();
        out->n_queries = cur->c[FFLINS_CLOUD_LIDAR].n;
        for (size_t i = 0; i < refs.size(); i++) {
            out->num_features[i] = nf[i];
            out->num_covered[i]  = nc[i];
            out->ref_ids[i]      = refs[i]->id;
        }
    }
    return FFLINS_OK;
}

int fflins_associate_pair(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, int32_t *num_features, int32_t *num_covered) {
    if (!ctx) return FFLINS_E_INVALID;
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    Frame *ref = find_frame(ctx, ref_id), *cur = find_frame(ctx, cur_id);
    if (!ref || !cur) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    int rc;
    if (!ref->index.fine || ref->hcap < next_pow2(2 * std::max(ref->c[FFLINS_CLOUD_MAP].n, 1)))
        if ((rc = build_hash(ctx, ref))) return rc;
    int32_t nf[kMaxRefs] = {0}, nc[kMaxRefs] = {0};
    std::vector<Frame *> refs{ref};
    rc = run_associate(ctx, refs, cur, nf, nc);
    if (rc) return rc;
    if (num_features) *num_features = nf[0];
    if (num_covered) *num_covered = nc[0];
    return FFLINS_OK;
}

int fflins_features_download(fflins_ctx *ctx, uint64_t ref_id, int32_t cap, int32_t *q, int8_t *type, uint8_t *covered,
                             uint8_t *outlier, int32_t *nn_idx, float *nn_d2, float *nn_points, double *unit_normal,
                             double *norm_inverse) {
    if (!ctx) return FFLINS_E_INVALID;
    {
        int rcm = resolve_map(ctx);
        if (rcm) return rcm;
    }
    Frame *r = find_frame(ctx, ref_id);
    if (!r) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    int n = r->feat_q;
    if (q) *q = n;
    if (!(type || covered || outlier || nn_idx || nn_d2 || nn_points || unit_normal || norm_inverse)) return FFLINS_OK;
    if (cap < n) return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
    if (n == 0) return FFLINS_OK;
    cudaStream_t s = ctx->stream;
    std::vector<int32_t> idx_tmp;
    if (nn_points && !nn_idx) {
        idx_tmp.resize((size_t) 5 * n);
        nn_idx = idx_tmp.data();
    }
    if (type) CK(cudaMemcpyAsync(type, r->feat.type, n, cudaMemcpyDeviceToHost, s));
    if (covered) CK(cudaMemcpyAsync(covered, r->feat.covered, n, cudaMemcpyDeviceToHost, s));
    if (outlier) CK(cudaMemcpyAsync(outlier, r->feat.outlier, n, cudaMemcpyDeviceToHost, s));
    if (nn_idx) CK(cudaMemcpyAsync(nn_idx, r->feat.nn_idx, (size_t) n * 20, cudaMemcpyDeviceToHost, s));
    if (nn_d2) CK(cudaMemcpyAsync(nn_d2, r->feat.nn_d2, (size_t) n * 20, cudaMemcpyDeviceToHost, s));
    if (unit_normal) CK(cudaMemcpyAsync(unit_normal, r->feat.normal, (size_t) n * 24, cudaMemcpyDeviceToHost, s));
    if (norm_inverse) CK(cudaMemcpyAsync(norm_inverse, r->feat.d, (size_t) n * 8, cudaMemcpyDeviceToHost, s));
    std::vector<float> map;
    if (nn_points) {
        map.resize((size_t) 4 * std::max(r->c[FFLINS_CLOUD_MAP].n, 1));
        if (r->c[FFLINS_CLOUD_MAP].n > 0)
            CK(cudaMemcpyAsync(map.data(), r->c[FFLINS_CLOUD_MAP].d, (size_t) r->c[FFLINS_CLOUD_MAP].n * 16, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if (nn_points) {
        for (size_t i = 0; i < (size_t) 5 * n; i++) {
            int id = nn_idx[i];
            if (id >= 0)
                std::memcpy(nn_points + 4 * i, map.data() + 4 * (size_t) id, 16);
            else
                std::memset(nn_points + 4 * i, 0, 16);
        }
    }
    return FFLINS_OK;
}

int fflins_features_set_outlier(fflins_ctx *ctx, uint64_t ref_id, const uint8_t *outlier, int32_t q) {
    if (!ctx || !outlier) return FFLINS_E_INVALID;
    Frame *r = find_frame(ctx, ref_id);
    if (!r) return fail(ctx, FFLINS_E_NOTFOUND, "unknown frame");
    REQUIRE(q == r->feat_q, "feature count mismatch");
    if (q > 0) {
        CK(cudaMemcpyAsync(r->feat.outlier, outlier, q, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return FFLINS_OK;
}

int fflins_factor_eval_batch(fflins_ctx *ctx, int32_t n, const float *point, const double *unit_normal,
                             const double *norm_inverse, const double *std, const double motion[14],
                             const double pose_ref[7], const double pose_cur[7], const double ext[7], double td,
                             uint32_t flags, double *r, double *J) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n >= 0 && (n == 0 || (point && unit_normal && norm_inverse && std)) && motion && pose_ref && pose_cur && ext,
            "bad argument");
    if (n == 0) return FFLINS_OK;
    cudaStream_t s = ctx->stream;
    size_t bytes   = (size_t) n * (12 + 24 + 8 + 8 + 8 + 176) + 1024;
    char *d        = (char *) ctx->pool.alloc(bytes);
    if (!d) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
    double *d_n   = (double *) d;
    double *d_d   = d_n + 3 * (size_t) n;
    double *d_std = d_d + n;
    double *d_r   = d_std + n;
    double *d_J   = d_r + n;
    float *d_p    = (float *) (d_J + 22 * (size_t) n);
    cudaMemcpyAsync(d_n, unit_normal, (size_t) n * 24, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_d, norm_inverse, (size_t) n * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_std, std, (size_t) n * 8, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_p, point, (size_t) n * 12, cudaMemcpyHostToDevice, s);
    FactorParams P{};
    P.n_pairs = 1;
    P.q       = n;
    P.flags   = flags & ~kFlagQuietNext;
    std::memcpy(P.pose_cur, pose_cur, 56);
    std::memcpy(P.ext, ext, 56);
    P.td = td;
    std::memcpy(P.pair[0].pose_ref, pose_ref, 56);
    std::memcpy(P.pair[0].v0, motion, 24);
    std::memcpy(P.pair[0].w0, motion + 3, 24);
    P.pair[0].td0 = motion[6];
    std::memcpy(P.v1, motion + 7, 24);
    std::memcpy(P.w1, motion + 10, 24);
    P.td1            = motion[13];
    P.sigma          = ctx->cfg.point_to_plane_std;
    P.huber_delta    = ctx->cfg.huber_delta;
    P.pair[0].type   = nullptr;
    P.pair[0].outlier = nullptr;
    P.pair[0].normal = d_n;
    P.pair[0].d      = d_d;
    P.pair[0].std    = d_std;
    {
        int64_t nl = 0;
        int rcl    = launch_factor_eval(ctx->frt, s, P, d_p, 0, r ? d_r : nullptr, J ? d_J : nullptr, nullptr, false, nullptr, false, nullptr, &nl);
        ctx->launches += nl;
        if (rcl < 0) {
            ctx->pool.release(d);
            return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString((cudaError_t) -rcl));
        }
    }
    if (r) cudaMemcpyAsync(r, d_r, (size_t) n * 8, cudaMemcpyDeviceToHost, s);
    if (J) cudaMemcpyAsync(J, d_J, (size_t) n * 176, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    ctx->pool.release(d);
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    return FFLINS_OK;
}

int fflins_factor_eval(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, const double pose_ref[7], const double pose_cur[7],
                       const double ext[7], double td, uint32_t flags, int32_t cap, int32_t *n_res, uint8_t *valid,
                       double *r, double *J, double *H, double *b, double *cost) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(pose_ref && pose_cur && ext, "bad argument");
    Frame *cur = find_frame(ctx, cur_id);
    if (!cur) return fail(ctx, FFLINS_E_NOTFOUND, "unknown cur frame");
    FactorParams P{};
    int rc;
    if ((rc = fill_factor_pairs(ctx, P, 1, &ref_id, pose_ref, cur, pose_cur, ext, td, flags))) return rc;
    const int q = P.q;
    if ((valid || r || J) && cap < q) return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
    cudaStream_t s = ctx->stream;
    if (q == 0) {
        if (n_res) *n_res = 0;
        if (H) std::memset(H, 0, 361 * 8);
        if (b) std::memset(b, 0, 19 * 8);
        if (cost) cost[0] = cost[1] = 0;
        return FFLINS_OK;
    }
    char *d = nullptr;
    double *d_r = nullptr, *d_J = nullptr;
    uint8_t *d_valid = nullptr;
    if (valid || r || J) {
        d = (char *) ctx->pool.alloc((size_t) q * (8 + 176 + 1) + 256);
        if (!d) return fail(ctx, FFLINS_E_NOMEM, "device allocation failed");
        d_r     = (double *) d;
        d_J     = d_r + q;
        d_valid = (uint8_t *) (d_J + 22 * (size_t) q);
    }
    {
        ProfScope ps(ctx->prof(), "factor_eval", s);
        int64_t nl = 0;
        int rcl    = launch_factor_eval(ctx->frt, s, P, (const float *) cur->c[FFLINS_CLOUD_LIDAR].d, 1, r ? d_r : nullptr,
                                        J ? d_J : nullptr, valid ? d_valid : nullptr, true, ctx->d_red, false, nullptr, &nl);
        ctx->launches += nl;
        if (rcl < 0) {
            ctx->pool.release(d);
            return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString((cudaError_t) -rcl));
        }
    }
    unsigned char *h = pin(ctx, kRedSize * 8);
    if (!h) {
        ctx->pool.release(d);
        return fail(ctx, FFLINS_E_NOMEM, "pinned allocation failed");
    }
    cudaMemcpyAsync(h, ctx->d_red, kRedSize * 8, cudaMemcpyDeviceToHost, s);
    if (r) cudaMemcpyAsync(r, d_r, (size_t) q * 8, cudaMemcpyDeviceToHost, s);
    if (J) cudaMemcpyAsync(J, d_J, (size_t) q * 176, cudaMemcpyDeviceToHost, s);
    if (valid) cudaMemcpyAsync(valid, d_valid, (size_t) q, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    ctx->pool.release(d);
    if (e != cudaSuccess) return fail(ctx, FFLINS_E_CUDA, cudaGetErrorString(e));
    unpack_red((const double *) h, H, b, cost, n_res);
    return FFLINS_OK;
}

int fflins_window_eval(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                       const double pose_cur[7], const double ext[7], double td, uint32_t flags, double *H, double *b,
                       double *cost, int32_t *n_res) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n_pairs >= 0 && (n_pairs == 0 || (ref_ids && pose_refs)) && pose_cur && ext, "bad argument");
    REQUIRE(!ctx->keyframes.empty(), "no keyframe in the window");
    if (n_pairs == 0) return FFLINS_OK;
    Frame *cur = find_frame(ctx, ctx->keyframes.back());
    FactorParams P{};
    int rc;
    if ((rc = fill_factor_pairs(ctx, P, n_pairs, ref_ids, pose_refs, cur, pose_cur, ext, td, flags))) return rc;
    if (P.q == 0) {
        for (int i = 0; i < n_pairs; i++) unpack_red(std::vector<double>(kRedSize, 0.0).data(), H ? H + 361 * i : nullptr,
                                                     b ? b + 19 * i : nullptr, cost ? cost + 2 * i : nullptr, n_res ? n_res + i : nullptr);
        return FFLINS_OK;
    }
    cudaStream_t s = ctx->stream;
    static const bool no_poll = getenv("FFLINS_NO_POLL") != nullptr;
    FactorResults res{};
    res.H     = H;
    res.b     = b;
    res.cost  = cost;
    res.n_res = n_res;
    int how;
    {
        // the reduced blocks are stored by the kernel straight into pinned host memory (zero-copy): no device-to-host
        // memcpy between the evaluation and its results. Normally the evaluation is POSTED to the resident kernel (no
        // launch at all); the launched kernel is the cold path and the one the per-stage profiler times.
        ProfScope ps(ctx->prof(), "factor_eval", s);
        int64_t nl = 0;
        how = launch_factor_eval(ctx->frt, s, P, (const float *) cur->c[FFLINS_CLOUD_LIDAR].d, 1, nullptr, nullptr, nullptr, true,
                                 nullptr, ctx->eval_inputs_settled && !ctx->prof_on && !no_poll, &res, &nl);
        ctx->launches += nl;
    }
    if (how < 0) return fail(ctx, FFLINS_E_CUDA, std::string("factor evaluation: ") + cudaGetErrorString((cudaError_t) -how));
    if (how == 1) {
        if (ctx->prof_on || no_poll) CK(cudaStreamSynchronize(s));   // the profiler's closing event needs the stream drained
        else CK(wait_signal(res.signal, n_pairs, res.seq, s));
        ctx->eval_inputs_settled = true; // the compute stream has drained up to here
    }
    if (res.unpacked) return FFLINS_OK;   // the resident path has already filed its result units under H / b / cost / n_res
    const double *red = res.red;
    // the lines were just written by the device: get all the misses in flight before the dependent reads
    for (size_t o = 0; o < (size_t) n_pairs * kRedSize * 8; o += 64) __builtin_prefetch((const char *) red + o, 0, 0);
    for (int i = 0; i < n_pairs; i++)
        unpack_red(red + (size_t) i * kRedSize, H ? H + 361 * i : nullptr, b ? b + 19 * i : nullptr,
                   cost ? cost + 2 * i : nullptr, n_res ? n_res + i : nullptr);
    return FFLINS_OK;
}

int fflins_window_chi2(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                       const double pose_cur[7], const double ext[7], double td, double threshold, int32_t *n_outliers) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n_pairs >= 0 && (n_pairs == 0 || (ref_ids && pose_refs)) && pose_cur && ext, "bad argument");
    REQUIRE(!ctx->keyframes.empty(), "no keyframe in the window");
    if (n_pairs == 0) return FFLINS_OK;
    Frame *cur = find_frame(ctx, ctx->keyframes.back());
    FactorParams P{};
    int rc;
    if ((rc = fill_factor_pairs(ctx, P, n_pairs, ref_ids, pose_refs, cur, pose_cur, ext, td, 0))) return rc;
    if (P.q == 0) {
        if (n_outliers) std::memset(n_outliers, 0, n_pairs * sizeof(int));
        return FFLINS_OK;
    }
    cudaStream_t s = ctx->stream;
    // counts and completion flags come back through pinned memory (the last CTA of each pair publishes them): no memset,
    // no copy, no stream synchronisation — and normally no launch either (resident kernel)
    static const bool no_poll = getenv("FFLINS_NO_POLL") != nullptr;
    FactorResults res{};
    int how;
    {
        ProfScope ps(ctx->prof(), "chi2", s);
        int64_t nl = 0;
        how = launch_chi2(ctx->frt, s, P, (const float *) cur->c[FFLINS_CLOUD_LIDAR].d, 1, threshold,
                          ctx->eval_inputs_settled && !ctx->prof_on && !no_poll, &res, &nl);
        ctx->launches += nl;
    }
    if (how < 0) return fail(ctx, FFLINS_E_CUDA, std::string("chi2 cull: ") + cudaGetErrorString((cudaError_t) -how));
    if (how == 1) {
        if (ctx->prof_on || no_poll) CK(cudaStreamSynchronize(s));
        else CK(wait_signal(res.signal, n_pairs, res.seq, s));
        ctx->eval_inputs_settled = true;
    }
    if (n_outliers) std::memcpy(n_outliers, res.counts, n_pairs * sizeof(int));
    return FFLINS_OK;
}

void fflins_solve_options_default(fflins_solve_options *o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->max_iterations[0]   = 5;
    o->max_iterations[1]   = 10;
    o->chi2_threshold      = 3.841;
    o->function_tolerance  = 1e-6;
    o->gradient_tolerance  = 1e-10;
    o->parameter_tolerance = 1e-8;
    o->initial_radius      = 1e4;
    o->flags               = FFLINS_HUBER;
    o->fixed_refs          = 0;
}

int fflins_window_solve(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, double *pose_refs, double pose_cur[7],
                        double ext[7], double *td, const fflins_solve_options *opt, const fflins_solve_prior *prior,
                        fflins_solve_report *report) {
    if (!ctx) return FFLINS_E_INVALID;
    REQUIRE(n_pairs >= 1 && ref_ids && pose_refs && pose_cur && ext && td && opt, "bad argument");
    REQUIRE(opt->max_iterations[0] >= 0 && opt->max_iterations[1] >= 0 && opt->max_iterations[0] + opt->max_iterations[1] <= 1000 &&
                opt->initial_radius > 0,
            "bad solve options");
    REQUIRE(!prior || (prior->H0 && prior->b0 && prior->x0), "incomplete prior");
    REQUIRE(!ctx->keyframes.empty(), "no keyframe in the window");
    Frame *cur = find_frame(ctx, ctx->keyframes.back());
    FactorParams P{};
    int rc;
    if ((rc = fill_factor_pairs(ctx, P, n_pairs, ref_ids, pose_refs, cur, pose_cur, ext, *td, opt->flags))) return rc;
    if (report) std::memset(report, 0, sizeof(*report));
    if (P.q == 0) return FFLINS_OK;   // no residual: the state stays where it is
    SolveOptions O{};
    O.max_iters[0]          = opt->max_iterations[0];
    O.max_iters[1]          = opt->max_iterations[1];
    O.chi2_threshold        = opt->chi2_threshold;
    O.function_tolerance    = opt->function_tolerance;
    O.gradient_tolerance    = opt->gradient_tolerance;
    O.parameter_tolerance   = opt->parameter_tolerance;
    O.initial_radius        = opt->initial_radius;
    O.min_relative_decrease = 1e-3;
    O.flags                 = opt->flags & 0x1fu;
    O.fixed_refs            = opt->fixed_refs;
    std::vector<double> x(7 * (size_t) n_pairs + 15);
    SolveReport R{};
    cudaStream_t s = ctx->stream;
    int how;
    {
        ProfScope ps(ctx->prof(), "window_solve", s);
        int64_t nl = 0;
        how = launch_window_solve(ctx->frt, s, P, (const float *) cur->c[FFLINS_CLOUD_LIDAR].d, 1, O, prior ? prior->H0 : nullptr,
                                  prior ? prior->b0 : nullptr, prior ? prior->x0 : nullptr, prior ? prior->c0 : 0.0, x.data(), &R, &nl);
        ctx->launches += nl;
    }
    if (how < 0) return fail(ctx, FFLINS_E_CUDA, std::string("window solve: ") + cudaGetErrorString((cudaError_t) -how));
    ctx->eval_inputs_settled = true;   // the kernel ran behind everything queued on the compute stream
    if (report) {
        report->status = R.status;
        for (int k = 0; k < 2; k++) {
            report->iterations[k]   = R.iterations[k];
            report->successful[k]   = R.successful[k];
            report->evaluations[k]  = R.evaluations[k];
            report->cost_initial[k] = R.cost_initial[k];
            report->cost_final[k]   = R.cost_final[k];
        }
        for (int i = 0; i < n_pairs; i++) {
            report->n_outliers[i] = R.n_outliers[i];
            report->n_res[i]      = R.n_res[i];
        }
        std::memcpy(report->phase_cycles, R.phase_cycles, sizeof(report->phase_cycles));
    }
    if (R.status != 0) return fail(ctx, FFLINS_E_CUDA, "window solve: a hand-off inside the kernel timed out");
    std::memcpy(pose_refs, x.data(), sizeof(double) * 7 * (size_t) n_pairs);
    std::memcpy(pose_cur, x.data() + 7 * (size_t) n_pairs, 56);
    std::memcpy(ext, x.data() + 7 * (size_t) n_pairs + 7, 56);
    *td = x[7 * (size_t) n_pairs + 14];
    return FFLINS_OK;
}

int fflins_profile_enable(fflins_ctx *ctx, int on) {
    if (!ctx) return FFLINS_E_INVALID;
    if (!on && ctx->prof_on) ctx->prof_collect();
    ctx->prof_on = on != 0;
    return FFLINS_OK;
}

int fflins_profile_reset(fflins_ctx *ctx) {
    if (!ctx) return FFLINS_E_INVALID;
    ctx->prof_collect();
    std::fill(ctx->prof_ms.begin(), ctx->prof_ms.end(), 0.0);
    std::fill(ctx->prof_cnt.begin(), ctx->prof_cnt.end(), 0);
    return FFLINS_OK;
}

int fflins_profile_get(fflins_ctx *ctx, char *names, int32_t names_cap, double *ms, int64_t *launches, int32_t cap,
                       int32_t *n) {
    if (!ctx || !n) return FFLINS_E_INVALID;
    ctx->prof_collect();
    *n = (int32_t) ctx->prof_names.size();
    if (!names && !ms && !launches) return FFLINS_OK;
    if (cap < *n) return fail(ctx, FFLINS_E_CAPACITY, "output buffer too small");
    std::string all;
    for (size_t i = 0; i < ctx->prof_names.size(); i++) {
        all += ctx->prof_names[i];
        all += '\n';
        if (ms) ms[i] = ctx->prof_ms[i];
        if (launches) launches[i] = ctx->prof_cnt[i];
    }
    if (names) {
        if ((int32_t) all.size() + 1 > names_cap) return fail(ctx, FFLINS_E_CAPACITY, "name buffer too small");
        std::memcpy(names, all.c_str(), all.size() + 1);
    }
    return FFLINS_OK;
}

int fflins_timer_start(fflins_ctx *ctx) {
    if (!ctx) return FFLINS_E_INVALID;
    if (!ctx->timer_a) {
        CK(cudaEventCreate(&ctx->timer_a));
        CK(cudaEventCreate(&ctx->timer_b));
    }
    CK(cudaEventRecord(ctx->timer_a, ctx->stream));
    return FFLINS_OK;
}

int fflins_timer_stop(fflins_ctx *ctx, double *ms) {
    if (!ctx || !ms || !ctx->timer_a) return FFLINS_E_INVALID;
    {
        int rcf = flush_frames(ctx); // queued frames belong to the timed region
        if (rcf) return rcf;
    }
    // the stop event is recorded behind EVERYTHING the context has in flight: the keyframe-map job on the map stream
    // (issued by the helper thread) and the background frames / copies on the copy stream are joined first
    if (ctx->map_job_active) {
        std::unique_lock<std::mutex> lk(ctx->map_mu);
        ctx->map_cv.wait(lk, [ctx] { return ctx->map_issued; });
    }
    CK(cudaEventRecord(ctx->ev_join_a, ctx->map_stream));
    CK(cudaEventRecord(ctx->ev_join_b, ctx->copy_stream));
    CK(cudaEventRecord(ctx->ev_join_c, ctx->bg_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join_a, 0));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join_b, 0));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join_c, 0));
    CK(cudaEventRecord(ctx->timer_b, ctx->stream));
    CK(cudaEventSynchronize(ctx->timer_b));
    float f = 0;
    CK(cudaEventElapsedTime(&f, ctx->timer_a, ctx->timer_b));
    *ms = f;
    return FFLINS_OK;
}

int64_t fflins_launch_count(const fflins_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int fflins_eval_stats_get(const fflins_ctx *ctx, fflins_eval_stats *out) {
    if (!ctx || !out) return FFLINS_E_INVALID;
    FactorStats st{};
    factor_runtime_stats(ctx->frt, &st);
    out->resident_launches  = st.resident_launches;
    out->resident_evals     = st.resident_evals;
    out->launched_evals     = st.launched_evals;
    out->resident_fallbacks = st.resident_fallbacks;
    return FFLINS_OK;
}

int fflins_measure_fp64(fflins_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return FFLINS_E_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    *tflops = measure_fp64_tflops(ctx->stream);
    CK(cudaGetLastError());
    return FFLINS_OK;
}

} // extern "C"
