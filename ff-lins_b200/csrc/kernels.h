// kernels.h — launchers of the hand-written sm_100a kernels (definitions in k_*.cu).
#pragma once
#include "common.cuh"

namespace fflins {

constexpr int kMaxRefs = 16;

// Optional per-kernel CUDA-event timing hook (implemented by the context, off by default).
struct Profiler {
    virtual void begin(const char *name, cudaStream_t s) = 0;
    virtual void end(cudaStream_t s)                     = 0;
    virtual ~Profiler() {}
};
struct ProfScope {
    Profiler *p;
    cudaStream_t s;
    ProfScope(Profiler *p_, const char *name, cudaStream_t s_) : p(p_), s(s_) {
        if (p) p->begin(name, s);
    }
    ~ProfScope() {
        if (p) p->end(s);
    }
};

// ---- K1 undistort (lidar_map.cc:223-256, misc.cc:27-105, pointcloud.cc:30-59) + per-frame K2/K3 -----
// The per-frame chain (interval rows -> undistortion + decimation -> single-CTA voxel filter + world
// projection) is launched for up to kFrameBatch queued frames at once: blockIdx.y (blockIdx.x for the voxel
// filter) selects the frame. A lone 150k-point frame keeps ~1/10 of the GPU busy for ~40 us of dependent
// launches; frames queued by asynchronous fflins_frame_process* calls share those three launches.
// rows: kRowSize doubles per window entry, the closed-form relative pose of IMU interval [i-1, i] w.r.t. the
// frame pose (math_undistort.cuh); row 0 = out-of-window fallback (misc.cc:76-79).
constexpr int kFrameBatch  = 8;
constexpr int kVoxSmallCap = 4096; // largest decimated cloud the single-CTA voxel filter takes
struct FrameItem {
    // prologue: the packed window [t(w) | p(3w) | q(4w)] may be pinned host memory (read zero-copy)
    const double *win;
    int w;
    Pose12 ext, frame;
    double *rows;        // [w][kRowSize]
    double *t_dev;       // [w] device copy of the time stamps for the interval lookup
    int *counters;       // [4] zeroed by the prologue; [0] = out-of-window fallbacks
    // undistortion: aos != nullptr: raw 32-byte PointXYZIT records; else xyzi + time arrays
    const float4 *xyzi;
    const double *time;
    const void *aos;
    int n;
    double td;
    int filter_num;
    unsigned dec_magic;  // ceil(2^32 / filter_num): k / filter_num = umulhi(k, dec_magic), exact while n * filter_num < 2^32 (0: divide)
    double t_front, t_back, inv_dt; // window span and (w - 1) / span, evaluated once on the host
    float4 *out_full, *out_dec;
    // voxel filter of the decimated cloud + world projection + count report (pinned, 2 ints)
    int ndec;
    float inv_leaf;
    float4 *down, *world;
    int *report;
    int *meta;           // [M_SIZE] scratch
    int *d_nout;
};
struct FrameBatch {
    int count;
    FrameItem it[kFrameBatch];
};
void launch_frame_prologue(cudaStream_t s, const FrameBatch &b);
void launch_frame_undistort(cudaStream_t s, const FrameBatch &b); // every item of one input kind (aos or not)
// items with ndec > kVoxSmallCap are skipped (the caller runs the multi-kernel filter for them)
void launch_frame_voxel(cudaStream_t s, const FrameBatch &b);
// static overload (lidar_map.cc:275-324): copy + decimate
void launch_copy_decimate(cudaStream_t s, const float4 *xyzi, int n, int filter_num, float4 *out_full,
                          float4 *out_dec);

// ---- K3 rigid projection (pointcloud.cc:38-59). n_dev (optional) overrides n with a device count.
// report (optional, host-mapped pinned memory, 2 ints): block 0 stores counters[0] and the point count
// there, so the caller needs no device-to-host copy after the frame chain.
void launch_project(cudaStream_t s, Pose12 T, const float4 *src, float4 *dst, int n, const int *n_dev,
                    int *report = nullptr, const int *counters = nullptr);
// up to kMaxRefs clouds in one launch (blockIdx.y selects the cloud)
struct ProjectBatch {
    int count;
    int n[kMaxRefs];
    const float4 *src[kMaxRefs];
    float4 *dst[kMaxRefs];
    Pose12 T[kMaxRefs];
};
void launch_project_batch(cudaStream_t s, const ProjectBatch &b);

// ---- K2 voxel downsample (pcl::VoxelGrid) --------------------------------------
struct VoxelWork {
    uint32_t *keys[2]   = {nullptr, nullptr};
    uint32_t *idx[2]    = {nullptr, nullptr};
    uint32_t *block_hist = nullptr; // 256 * nblocks
    uint32_t *block_heads = nullptr;
    uint32_t *seg_start  = nullptr; // cap + 1
    int      *meta      = nullptr;  // see k_voxel.cu
    int       cap       = 0;        // points capacity
    mutable bool fresh  = true;     // meta (barrier counters of the cooperative kernel) not yet zeroed
};
size_t voxel_work_bytes(int cap, size_t *offsets /*8*/);
// out must hold n points; *d_nout (device int) receives the voxel count, d_flag receives 1 on the
// PCL "leaf too small" overflow (output = input). keys_out optional (n).
// Optional fused tail (single-CTA path only, n <= 4096): world[i] = T * out[i] and report[0..1] =
// {counters[0], voxel count} stored to pinned host memory. Returns 1 when the tail was fused, 0 when the
// caller still has to project, 2 when the voxel count was stored to `report` (optional pinned host int; large-cloud
// cooperative kernel only), < 0 on error.
struct VoxelTail {
    Pose12 T;
    float4 *world;
    int *report;
    const int *counters;
};
int launch_voxel_downsample(cudaStream_t s, const VoxelWork &w, const float4 *in, int n, float leaf, float4 *out,
                            int *d_nout, int32_t *keys_out, int64_t *launches, Profiler *prof, const char *tag,
                            const VoxelTail *tail = nullptr, int *report = nullptr);

// ---- K4 keyframe hash insert (replaces ikd_Tree::build, lidar_map.cc:100-118) -------------------------
// Two-level (plus one) open-addressing index of a keyframe map, all tables with hmask + 1 slots:
//   fine   : one 32-byte slot per map point {packed cell key, x, y, z, map index} - a probe that hits brings the
//            point along in the same memory round trip (points sharing a cell sit in consecutive probe slots);
//   coarse : {packed 4x4x4-cell block key, 64-bit occupancy mask of its cells};
//   super  : {packed 4x4x4-coarse-block key, 64-bit occupancy mask of its coarse blocks}.
struct alignas(32) FineSlot {
    unsigned long long key;
    float x, y, z;
    int idx;
    unsigned long long pad;
};
struct alignas(16) MaskSlot {
    unsigned long long key;
    unsigned long long mask;
};
struct MapIndex {
    FineSlot *fine   = nullptr;
    MaskSlot *coarse = nullptr;
    MaskSlot *super  = nullptr;
    int hmask        = 0;
};
// m_dev != nullptr: the count is read from device memory (m is then only an upper bound used to size the grid)
void launch_hash_build(cudaStream_t s, const float4 *map, int m, float inv_cell, MapIndex index,
                       const int *m_dev = nullptr);

// ---- K5 association (lidar_map.cc:326-408, pointcloud.cc:61-93, ikd_Tree.cc:897-924) -------------------
struct FeatureSoA {
    int8_t  *type    = nullptr; // FEATURE_NONE / PLANE
    uint8_t *covered = nullptr;
    uint8_t *outlier = nullptr;
    int32_t *nn_idx  = nullptr; // 5 per query
    float   *nn_d2   = nullptr; // 5 per query
    double  *normal  = nullptr; // 3 per query
    double  *d       = nullptr; // norm_inverse
};
struct AssocRef {
    const float4 *map;
    MapIndex index;
    int m;
    Pose12 T_ref_world;
    FeatureSoA feat;
    int *counts; // [2]: num_features, num_covered
};
struct AssocParams {
    int n_refs;
    int q;
    float inv_cell;
    float cell;
    int r_max;
    float max_d2;          // (max_search_square_distance)^2, ikd_Tree.cc:902
    float plane_d2_gate;   // max_search_square_distance, lidar_map.cc:374
    double plane_thr, sigma, scalar_thr, sigma_gate;
    long long *dbg_cycles; // debug: per-(ref, query) warp cycles (nullptr = off)
    const int *q_dev;      // optional: the query count still sits in device memory (q is then an upper bound)
    // completion: the last CTA of the plane kernel copies the per-ref counts (AssocRef::counts, zero on entry, zero
    // again on exit) to host_counts[2 * ref + {0, 1}] (pinned) and releases *signal = seq behind them
    int *ticket;           // two device ints, zero on entry / exit: [0] completion ticket of the plane kernel, [1] task counter of the search
    int *host_counts;
    unsigned long long *signal;
    unsigned long long seq;
    AssocRef ref[kMaxRefs];
};
void launch_associate(cudaStream_t s, const AssocParams &p, const float4 *cur_world, const float4 *cur_lidar);
// brute-force-free stand-alone 5-NN against a hash (fflins_knn5)
void launch_knn5(cudaStream_t s, MapIndex index, int m, const float4 *query, int q, float inv_cell, float cell, int r_max,
                 float max_d2, int32_t *nn_idx, float *nn_d2, int32_t *nn_count);
void launch_plane_estimation(cudaStream_t s, const float4 *pts, int n, double thr, double *unit, double *ninv,
                             double *dist, uint8_t *ok);

// ---- K6 / K7 factors (lidar_factor.cc:41-118, residual_block_info.cc:49-77, marginalization_info.cc:181-210)
// everything of LidarFactor::Evaluate that depends only on the (ref, cur) pair (math_factor.cuh)
struct CurConst {           // ... and only on the current frame / extrinsic (shared by every pair of a launch)
    M3 R1, B1, Rt1, Rbl;
    V3 tt1, tbl, w1;
};
struct RefConst {           // ... and on the reference frame too
    M3 R0, B0, Rt0, D, E;
    V3 tt0, w0, dv;
};
struct PairConst {
    CurConst cur;
    RefConst ref;
};
struct FactorPair {
    const int8_t *type;     // nullptr => every residual valid (stateless batch)
    uint8_t *outlier;       // read by K6, written by K7
    const double *normal;
    const double *d;
    const double *std;      // nullptr => sigma below
    double pose_ref[7];
    double v0[3], w0[3], td0;
};
struct FactorParams {
    int n_pairs;
    int q;                  // features per pair
    uint32_t flags;
    double pose_cur[7], ext[7], td;
    double v1[3], w1[3], td1;
    double sigma, huber_delta;
    FactorPair pair[kMaxRefs];
};
// What the kernels receive. The data of an evaluation is split by how often it changes:
//   FactorSetup  once per keyframe (feature arrays, frozen per-frame motion): lives in device memory, uploaded when it
//                changes (setup generation);
//   FactorMsg    once per evaluation (what a solver iteration changes: pose blocks, extrinsic, td, flags): ~1 KB of
//                64-bit words, passed by value to a launched kernel or through the pinned mailbox of the RESIDENT kernel.
// Everything of LidarFactor::Evaluate that depends only on the (ref, cur) pair (CurConst / RefConst) is evaluated on
// the device by the first threads of each CTA (pair_const_to_smem).
struct FactorSetupPair {
    const int8_t *type;     // nullptr => every residual valid (stateless batch)
    uint8_t *outlier;       // read by K6, written by K7
    const double *normal;
    const double *d;
    const double *std;      // nullptr => sigma
    double v0[3], w0[3], td0;
};
struct FactorSetup {
    const float *points;    // cur-frame points: float4 (stride4 = 1) or packed float3
    int stride4;
    int q;                  // features per pair
    double v1[3], w1[3], td1;
    double sigma, huber_delta;
    FactorSetupPair pair[kMaxRefs];
};
enum { kCmdEval = 1, kCmdChi2 = 2, kCmdQuit = 3 };
// internal flag bit (never an API flag): while host-to-device copies are in flight only ONE warp of the resident kernel polls
// host memory for the next message (k_factor.cu, quiet mode)
constexpr uint32_t kFlagQuietNext = 1u << 31;
// word 0: cmd | flags << 32     word 1: n_pairs | setup generation << 32     word 2: td     word 3: chi2 threshold
// word 4: completion value (signal[pair] = seq)   5..11 pose_cur   12..18 ext   19 + 7 i .. pose_ref[i]
// A window holds at most kMaxRefs keyframes, i.e. kMaxRefs - 1 references: 19 + 7 * 15 = 124 words, so that the
// 128 threads of the resident kernel's dispatcher fetch a whole message with ONE 16-byte load each.
constexpr int kMaxPairs = kMaxRefs - 1;
constexpr int kMsgHead  = 19;
constexpr int kMsgWords = kMsgHead + 7 * kMaxPairs;
static_assert(kMsgWords <= 128, "a message must fit one poll of the dispatcher CTA");
struct FactorMsg {
    unsigned long long w[kMsgWords];
};
constexpr int kFactorBlock = 128;
constexpr int kFactorRanks = 8;   // CTAs per pair (each takes every 8th 128-residual tile)
constexpr int kRedSize     = 212; // 190 upper-tri H + 19 b + cost0 + cost1 + count

// Per-context state of K6 / K7: setup ring, cross-CTA partial sums, and the resident evaluation kernel with its
// pinned mailbox. Created / destroyed by the context (k_factor.cu).
struct FactorRuntime;
FactorRuntime *factor_runtime_create(int device);
void factor_runtime_destroy(FactorRuntime *rt);
struct FactorStats {
    long long resident_launches, resident_evals, launched_evals, resident_fallbacks;
};
void factor_runtime_stats(const FactorRuntime *rt, FactorStats *out);
// Stop the resident kernel (if it is running) and wait for it: called before anything that synchronises the device.
void factor_runtime_quiesce(FactorRuntime *rt);

// points: xyz of cur-frame points as float4 (stride4=1) or packed float3 (stride4=0).
// r/J/valid (optional, device) are per (pair, feature): r[pair*q+k], J[22*(pair*q+k)].
// Where the reduced results of the solver-facing calls arrive: pinned host memory owned by the runtime, written by the
// kernels (zero-copy) with signal[pair] = seq released at system scope behind each pair's data, so the host polls
// instead of synchronising a stream.
struct FactorResults {
    const double *red;                 // [kMaxRefs][kRedSize]
    const int *counts;                 // [kMaxRefs] chi2
    const unsigned long long *signal;  // [kMaxRefs]
    unsigned long long seq;            // value the flags take for THIS call
    // optional: where the caller wants the blocks (per pair: H 19x19 row-major, b 19, cost 2, n_res). A call served by
    // the resident kernel unpacks its tagged result units straight into them (unpacked = true) instead of into `red`.
    double *H, *b, *cost;
    int32_t *n_res;
    bool unpacked;
};
// want_red: reduce into the runtime's pinned result block (FactorResults) — or, when d_red is given, into device memory
// (no flags). resident = true (only with want_red and no per-feature outputs): the evaluation is POSTED to the resident
// kernel's mailbox instead of being launched (the kernel is started on demand and retires by itself when idle) and the
// call returns after the results have ARRIVED. Returns 0 = done, 1 = launched on `s` (the caller waits for res->signal or
// the stream), < 0 = error (cudaError_t negated).
int launch_factor_eval(FactorRuntime *rt, cudaStream_t s, const FactorParams &p, const float *points, int stride4, double *r,
                       double *J, uint8_t *valid, bool want_red, double *d_red, bool resident, FactorResults *res,
                       int64_t *launches);
// K7: outlier flags are set in p.pair[i].outlier; per-pair counts land in res->counts. Same return convention.
int launch_chi2(FactorRuntime *rt, cudaStream_t s, const FactorParams &p, const float *points, int stride4, double threshold,
                bool resident, FactorResults *res, int64_t *launches);
// ---- device-side window solve (k_factor.cu: solve_kernel, math_solve.cuh): Optimizer::optimization (optimizer.cc:85-185)
// over the LiDAR factors of the window and a quadratic prior, one launch per keyframe.
struct SolveOptions {
    int max_iters[2];            // optimizer.cc:87-88
    double chi2_threshold;       // <= 0: no cull between the two solves
    double function_tolerance, gradient_tolerance, parameter_tolerance, initial_radius, min_relative_decrease;
    unsigned flags;              // FFLINS_HUBER / FFLINS_CONST_* as in fflins_window_eval
    unsigned fixed_refs;         // bit p: reference pose p is held constant
};
struct SolveReport {
    int status;                  // 0 ok, 1 a hand-off inside the kernel timed out
    int iterations[2], successful[2], evaluations[2], failed[2];
    double cost_initial[2], cost_final[2];
    int steps;                   // broadcasts of the solver CTA (evaluations + chi2 + retire)
    int n_outliers[kMaxRefs], n_res[kMaxRefs];
    // diagnostics: SM cycles of the solver CTA per phase, summed over the solve: 0 waiting for the evaluators (broadcast ->
    // all sums gathered), 1 assembly, 2 damped copy, 3 elimination + back substitution, 4 model decrease, 5 Plus
    double phase_cycles[8];
};
// State layout x: n reference poses, the current pose, the extrinsic (7 doubles each: t, q = x y z w), td; the prior
// (host pointers, H0 == nullptr: none) is 0.5 d^T H0 d - b0^T d + c0 over the tangent [6 n | 6 | 6 | 1], d = x (-) x0.
// Synchronous: returns after the result has arrived in pinned memory. 0 ok, < 0 cudaError_t negated.
int launch_window_solve(FactorRuntime *rt, cudaStream_t s, const FactorParams &p, const float *points, int stride4,
                        const SolveOptions &opt, const double *H0, const double *b0, const double *x0, double c0, double *x_out,
                        SolveReport *rep, int64_t *launches);
// fp64 FMA throughput of the device (the roofline of K6): TFLOP/s of a register-resident DFMA loop on every SM.
double measure_fp64_tflops(cudaStream_t s);

} // namespace fflins
