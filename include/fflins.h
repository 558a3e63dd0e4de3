/*
 * fflins.h — C-ABI of the B200-native FF-LINS LiDAR hot path.
 *
 * Every entry point replaces one interface of the reference (paths relative to the
 * upstream FF-LINS tree); the C++ shim in ff-lins_b200/shim/ binds them behind the
 * reference's own class names. Rules of the boundary:
 *   - plain pointers and sizes, POD structs, int status (0 ok, <0 FFLINS_E_*); never throws;
 *   - one context per session per GPU; calls on one context are serialised by the caller
 *     (the reference has a single fusion thread, ff_lins/ff_lins/ff_lins.cc:166);
 *   - host buffers are caller-owned, read/written only during the call;
 *   - NO CPU fallback: without a CUDA device fflins_create fails with FFLINS_E_NODEVICE.
 *
 * Conventions
 *   point cloud   float xyzi[4*n]           (x, y, z, intensity) — the 16 live bytes of pcl::PointXYZI
 *   raw AoS       32-byte PointXYZIT records (ff_lins/lidar/lidar.h:30-38): x,y,z,pad | intensity,pad | double time
 *   pose          fflins_pose {R row-major 3x3, t}   (common/types.h:43-46 `Pose`)
 *   pose block    double[7] = tx,ty,tz,qx,qy,qz,qw   (factors/lidar_factor.cc:43-50)
 *   quaternions   x,y,z,w (Eigen coeffs order)
 *   H block       19x19 row-major, order [ref t, ref th, cur t, cur th, ext t, ext th, td]
 */
#ifndef FFLINS_H
#define FFLINS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fflins_ctx fflins_ctx;

enum {
    FFLINS_OK          = 0,
    FFLINS_E_INVALID   = -1, /* bad argument */
    FFLINS_E_CUDA      = -2, /* CUDA runtime error, see fflins_last_error */
    FFLINS_E_NOMEM     = -3,
    FFLINS_E_NOTFOUND  = -4, /* unknown frame id */
    FFLINS_E_CAPACITY  = -5, /* caller buffer too small */
    FFLINS_E_NODEVICE  = -6  /* no CUDA device: there is no CPU path */
};

/* Which cloud of a frame (lidar/lidar_frame.h:179-186). */
enum {
    FFLINS_CLOUD_LIDAR_FULL = 0, /* pointcloud_lidar_full_: undistorted + index-decimated   */
    FFLINS_CLOUD_LIDAR      = 1, /* pointcloud_lidar_     : voxel centroids, lidar frame    */
    FFLINS_CLOUD_WORLD      = 2, /* pointcloud_world_     : same points, world frame        */
    FFLINS_CLOUD_MAP_FULL   = 3, /* pointcloud_map_full_  : all undistorted points (merged) */
    FFLINS_CLOUD_MAP        = 4  /* pointcloud_map_       : voxel centroids of MAP_FULL     */
};

/* Feature type (lidar/lidar_feature.h:40-44). */
enum { FFLINS_FEATURE_NONE = -1, FFLINS_FEATURE_PLANE = 0, FFLINS_FEATURE_CORNER = 1 };

/* Flags of the factor entry points. */
enum {
    FFLINS_HUBER      = 1u << 0, /* apply HuberLoss(1.0) corrector per residual (residual_block_info.cc:49-77) */
    FFLINS_CONST_REF  = 1u << 1, /* parameter block held constant: its Jacobian columns are zero            */
    FFLINS_CONST_CUR  = 1u << 2, /*   (what Ceres signals with jacobians[i]==nullptr,                       */
    FFLINS_CONST_EXT  = 1u << 3, /*    optimizer.cc:418-435)                                                */
    FFLINS_CONST_TD   = 1u << 4
};

/* Mirrors the yaml keys read by LidarMap::LidarMap (lidar/lidar_map.cc:45-55) plus the
 * hard-coded constants of lidar_map.h:87-94, lidar_map.cc:385, optimizer.cc:444,516. */
typedef struct fflins_config {
    double  point_to_plane_std;           /* optimizer.optimize_point_to_plane_std           */
    int32_t window_size;                  /* optimizer.optimize_window_size                  */
    int32_t point_filter_num;             /* lidar.point_filter_num                          */
    double  downsample_size;              /* lidar.downsample_size (voxel leaf, m)           */
    double  frame_rate;                   /* lidar.frame_rate                                */
    double  plane_estimation_threshold;   /* lidar.plane_estimation_threshold                */
    double  minimum_keyframe_translation; /* lidar.minimum_keyframe_translation              */
    double  max_search_square_distance;   /* lidar_map.h:89  = 5.0 (squared again by ikd-Tree, ikd_Tree.cc:902) */
    double  max_keyframe_interval;        /* lidar_map.h:94  = 0.475 s                       */
    double  min_keyframe_rotation;        /* lidar_map.h:92  = 10 deg, in rad                */
    double  plane_scalar_threshold;       /* lidar_map.cc:385 = 0.1                          */
    double  plane_sigma_gate;             /* lidar_map.cc:385 = 4.5                          */
    double  huber_delta;                  /* optimizer.cc:444 = 1.0                          */
    int32_t max_points_per_frame;         /* capacity hint, 0 -> 262144                      */
    int32_t max_merge_frames;             /* frames merged into one keyframe, 0 -> 8         */
} fflins_config;

typedef struct fflins_pose {
    double R[9]; /* row-major */
    double t[3];
} fflins_pose;

typedef struct fflins_frame_info {
    uint64_t    id;
    int32_t     n_raw;      /* points given                                             */
    int32_t     n_dec;      /* pointcloud_lidar_full_ size (index decimation)           */
    int32_t     n_down;     /* pointcloud_lidar_/world_ size (voxel centroids) = Q      */
    int32_t     n_map;      /* pointcloud_map_ size (0 until merged)                    */
    int32_t     n_fallback; /* points whose time fell outside the INS window (misc.cc:76-79) */
    int32_t     reserved;
    fflins_pose pose;       /* lidar pose at `stamp`                                    */
} fflins_frame_info;

typedef struct fflins_assoc_stats {
    int32_t n_refs;
    int32_t n_queries;           /* Q of the current keyframe                          */
    int32_t num_features[16];    /* per ref, PLANE count (lidar_map.cc:401)            */
    int32_t num_covered[16];     /* per ref, 5-neighbours-found count (:402)           */
    uint64_t ref_ids[16];
} fflins_assoc_stats;

/* ---- context ------------------------------------------------------------------------- */
void        fflins_config_default(fflins_config *cfg);                 /* values of config/ff_lins_robot.yaml */
int         fflins_create(const fflins_config *cfg, int device, fflins_ctx **out); /* LidarMap::LidarMap lidar_map.cc:42 */
void        fflins_destroy(fflins_ctx *ctx);
const char *fflins_last_error(const fflins_ctx *ctx);
const char *fflins_version(void);
int         fflins_synchronize(fflins_ctx *ctx);
/* Page-lock a caller buffer so the H2D/D2H copies inside the calls run at PCIe rate. */
int         fflins_host_register(void *ptr, uint64_t bytes);
int         fflins_host_unregister(void *ptr);
/* Returns once no queued or in-flight host-to-device copy of this context still reads the caller's point buffer `ptr`
 * (asynchronous frames): after it the buffer may be freed, overwritten or un-registered. Cheap when the copy is long done. */
int         fflins_host_buffer_done(fflins_ctx *ctx, const void *ptr);
/* Announce a raw frame (n 32-byte PointXYZIT records in page-locked memory) before its fflins_frame_process_aos call: the
 * host-to-device copy starts now, behind the copies already queued, and the later call with the same pointer and count
 * finds the records on the device. The counterpart of LINS::addNewLidarFrame (ff_lins.cc): the sensor callback hands a
 * frame over while the fusion thread is still solving the previous keyframe. Up to 8 frames may be announced (beyond
 * that, and for a buffer announced twice, the oldest announcement is dropped; one that is never processed is dropped when
 * its staging slot is needed); pageable buffers are ignored. The buffer must stay valid and
 * unchanged until the frame has been processed (fflins_host_buffer_done). */
int         fflins_host_prefetch(fflins_ctx *ctx, const void *points32, int32_t n);

/* ---- per frame: LidarMap::lidarFrameProcessing (lidar_map.cc:213-273) ------------------ */
/* `out` may be NULL = ASYNCHRONOUS mode: the call only QUEUES the frame (INS window packed, frame pose
 * evaluated on the host, buffers sized) and returns. Up to 8 queued frames share one launch of each kernel of
 * the per-frame chain; they are launched, and their counts read back, by the next call that needs a result
 * (fflins_frame_info_get, fflins_keyframe_merge, fflins_associate, downloads, fflins_synchronize, ...).
 * The frame pose is always available at once (fflins_frame_get_pose). Buffer lifetime in asynchronous mode:
 * page-locked (fflins_host_register / cudaMallocHost) point buffers and device point buffers (_dev) must stay
 * untouched until fflins_synchronize, fflins_frame_info_get or a download of THAT frame returns — their
 * copies are issued when the batch is launched, newest frame first, and the older frames of the batch are
 * processed on a background stream behind their own copies so that association and factor evaluation of the newest
 * frame never wait for them; pageable buffers are copied before the call returns.
 * SoA input. ins_* is the INS window (deque<pair<IMU,IntegrationState>>): strictly
 * increasing times, positions, quaternions; 2 <= w <= 65536 (misc.cc:27-62). stamp is
 * frame->stamp() (= end time + td, ff_lins.cc:352-353), td is frame->timeDelay(). */
int fflins_frame_process(fflins_ctx *ctx, uint64_t frame_id, const float *xyzi, const double *time, int32_t n,
                         const double *ins_t, const double *ins_p, const double *ins_q, int32_t w,
                         const fflins_pose *pose_b_l, double stamp, double td, fflins_frame_info *out);
/* Same, input as the reference's raw 32-byte PointXYZIT records (zero host repacking). */
int fflins_frame_process_aos(fflins_ctx *ctx, uint64_t frame_id, const void *points32, int32_t n,
                             const double *ins_t, const double *ins_p, const double *ins_q, int32_t w,
                             const fflins_pose *pose_b_l, double stamp, double td, fflins_frame_info *out);
/* Same, point data already resident in device memory (d_xyzi, d_time are device pointers; the small
 * INS window stays a host buffer — it is produced by the CPU-side INS mechanisation). */
int fflins_frame_process_dev(fflins_ctx *ctx, uint64_t frame_id, const float *d_xyzi, const double *d_time, int32_t n,
                             const double *ins_t, const double *ins_p, const double *ins_q, int32_t w,
                             const fflins_pose *pose_b_l, double stamp, double td, fflins_frame_info *out);
/* Static overload used during initialisation (lidar_map.cc:275-324): no undistortion. */
int fflins_frame_process_static(fflins_ctx *ctx, uint64_t frame_id, const float *xyzi, int32_t n,
                                const double state_p[3], const double state_q[4], const fflins_pose *pose_b_l,
                                fflins_frame_info *out);

int fflins_frame_info_get(fflins_ctx *ctx, uint64_t frame_id, fflins_frame_info *out);
int fflins_frame_download(fflins_ctx *ctx, uint64_t frame_id, int which, float *xyzi, int32_t cap, int32_t *n);
/* Replace one cloud of a frame from host data (creates the frame if unknown). */
int fflins_frame_upload(fflins_ctx *ctx, uint64_t frame_id, int which, const float *xyzi, int32_t n);
int fflins_frame_set_pose(fflins_ctx *ctx, uint64_t frame_id, const fflins_pose *pose);    /* LidarFrame::setPose */
int fflins_frame_get_pose(fflins_ctx *ctx, uint64_t frame_id, fflins_pose *pose);
/* LidarFrame::setVelocity/setAngularVelocity/setTimeDelay (ff_lins.cc:474-489). */
int fflins_frame_set_motion(fflins_ctx *ctx, uint64_t frame_id, const double v[3], const double omega[3], double td);
int fflins_frame_release(fflins_ctx *ctx, uint64_t frame_id);
/* setPose + pointCloudProjection(lidar -> world) (optimizer.cc:203-218). */
int fflins_reproject(fflins_ctx *ctx, uint64_t frame_id, const fflins_pose *pose);
/* The whole loop of Optimizer::updateParametersFromOptimizer (optimizer.cc:203-218) in one launch:
 * setPose + lidar->world projection for n (<= 16) frames. */
int fflins_reproject_window(fflins_ctx *ctx, int32_t n, const uint64_t *frame_ids, const fflins_pose *poses);
/* LidarMap::keyframeSelection (lidar_map.cc:77-98): host arithmetic on stored poses. */
int fflins_keyframe_selection(fflins_ctx *ctx, uint64_t frame_id, double frame_stamp, double last_key_stamp,
                              int *is_keyframe, double stats[3]);

/* ---- stage-level entry points (PCL VoxelGrid / PointCloudCommon replacements) ----------- */
/* pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(leaf) (lidar_map.cc:71,209,260).
 * keys (optional, n entries) receives each input point's voxel key. */
int fflins_voxel_downsample(fflins_ctx *ctx, const float *xyzi, int32_t n, double leaf, float *out_xyzi, int32_t cap,
                            int32_t *n_out, int32_t *keys);
/* PointCloudCommon::pointCloudProjection (pointcloud.cc:38-51); dst may alias src. */
int fflins_cloud_projection(fflins_ctx *ctx, const fflins_pose *pose, const float *src, int32_t n, float *dst);
/* PointCloudCommon::planeEstimation (pointcloud.cc:61-93), batched: pts is n x 5 x xyzi. */
int fflins_plane_estimation(fflins_ctx *ctx, const float *pts, int32_t n, double threshold, double *unit_normal,
                            double *norm_inverse, double *distance, uint8_t *is_plane);
/* KD_TREE::nearestSearch semantics (ikd_Tree.cc:360-391, 897-924) against an ad-hoc map:
 * exact 5-NN by fp32 d2, cut-off d2 <= max_dist^2, ties by lower map index, ascending. */
int fflins_knn5(fflins_ctx *ctx, const float *map_xyzi, int32_t m, const float *query_xyzi, int32_t q,
                double max_dist, int32_t *nn_idx, float *nn_d2, int32_t *nn_count);

/* ---- keyframe map maintenance ---------------------------------------------------------- */
/* LidarMap::mergeNonKeyframes (lidar_map.cc:190-211). `out` may be NULL = asynchronous: projection and map
 * voxel filter run on the context's map stream, concurrently with the association and factor evaluation of
 * the same keyframe (which do not read the new map); the map and its voxel hash are adopted by the first
 * call that needs them (the next fflins_associate, a download, fflins_frame_info_get, ...). */
int fflins_keyframe_merge(fflins_ctx *ctx, uint64_t key_id, const uint64_t *nonkey_ids, int32_t n,
                          fflins_frame_info *out);
/* LidarMap::addNewKeyFrame (lidar_map.cc:100-126): the keyframe joins the window; its map gets the device search index
 * that replaces the ikd-Tree (inserted here, or already by the asynchronous map job of fflins_keyframe_merge). */
int fflins_keyframe_add(fflins_ctx *ctx, uint64_t key_id);
int fflins_keyframe_remove_oldest(fflins_ctx *ctx, uint64_t *removed_id); /* lidar_map.cc:128-173 */
int fflins_keyframe_remove_newest(fflins_ctx *ctx, uint64_t *removed_id); /* lidar_map.cc:175-188 */
int fflins_keyframe_count(fflins_ctx *ctx, int32_t *n);
int fflins_keyframe_ids(fflins_ctx *ctx, uint64_t *ids, int32_t cap, int32_t *n);

/* ---- association: LidarMap::updateLidarFeaturesInMap / findFeaturesInMap ---------------- */
/* cur = newest keyframe, refs = all older keyframes (lidar_map.cc:428-459). Always synchronous: the counts are
 * returned when the features of every (ref, cur) pair are complete in device memory. */
int fflins_associate(fflins_ctx *ctx, fflins_assoc_stats *out);
/* One (ref, cur) pair (lidar_map.cc:326-408). */
int fflins_associate_pair(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, int32_t *num_features,
                          int32_t *num_covered);
/* Features stored on the ref frame, indexed by cur-frame point (lidar_feature.h:35-114).
 * Any output may be NULL. nn_points is q x 5 x xyzi (LidarFeature::nearestPoints). */
int fflins_features_download(fflins_ctx *ctx, uint64_t ref_id, int32_t cap, int32_t *q, int8_t *type,
                             uint8_t *covered, uint8_t *outlier, int32_t *nn_idx, float *nn_d2, float *nn_points,
                             double *unit_normal, double *norm_inverse);
int fflins_features_set_outlier(fflins_ctx *ctx, uint64_t ref_id, const uint8_t *outlier, int32_t q);

/* ---- factors ---------------------------------------------------------------------------- */
/* Stateless batch of LidarFactor::Evaluate (factors/lidar_factor.cc:41-118).
 * point xyz[3n] (f32), unit_normal[3n], norm_inverse[n], std[n]; motion = v0[3],w0[3],td0,v1[3],w1[3],td1 (14).
 * r[n] and J[22n] (row-major 1x7,1x7,1x7,1x1 per residual) may be NULL. */
int fflins_factor_eval_batch(fflins_ctx *ctx, int32_t n, const float *point, const double *unit_normal,
                             const double *norm_inverse, const double *std, const double motion[14],
                             const double pose_ref[7], const double pose_cur[7], const double ext[7], double td,
                             uint32_t flags, double *r, double *J);
/* All valid residuals (PLANE, not outlier; optimizer.cc:468) of one (ref, cur) pair from the
 * device-resident features: optional per-feature r[Q], J[22Q], valid[Q]; optional reduced
 * H[361] = sum J^T J, b[19] = -sum J^T r (marginalization_info.cc:181-210), cost[2] =
 * {0.5*sum rho(r^2), sum r'^2}. */
int fflins_factor_eval(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, const double pose_ref[7],
                       const double pose_cur[7], const double ext[7], double td, uint32_t flags, int32_t cap,
                       int32_t *n_res, uint8_t *valid, double *r, double *J, double *H, double *b, double *cost);
/* Every (ref, newest keyframe) pair of the window in ONE launch: H[361*n], b[19*n], cost[2*n], n_res[n]. */
int fflins_window_eval(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                       const double pose_cur[7], const double ext[7], double td, uint32_t flags, double *H,
                       double *b, double *cost, int32_t *n_res);
/* Optimizer::lidarOutlierCullingByChi2 (optimizer.cc:515-548): r^2 > threshold -> outlier. */
int fflins_window_chi2(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                       const double pose_cur[7], const double ext[7], double td, double threshold,
                       int32_t *n_outliers);

/* ---- the window solve on the device ----------------------------------------------------- */
/* Optimizer::optimization (optimizer.cc:85-185) over the LiDAR factors of the window: Levenberg-Marquardt with
 * max_iterations[0] iterations, the chi2 cull (optimizer.cc:515-548), max_iterations[1] iterations, as ONE kernel
 * launch: evaluation (lidar_factor.cc:41-118), J^T J reduction (marginalization_info.cc:181-210), assembly of the
 * window's normal equations, factorisation, PoseManifold::Plus (pose_manifold.cc:35-50) and the accept / reject
 * decision all happen on the device; only the final state returns. fflins_window_eval remains the entry point for a
 * host solver (Ceres) that wants the blocks of every iteration.
 * State: n reference poses, the current pose, the extrinsic (7 doubles: t, q = x y z w), td. Tangent vector
 * [ref 0 .. ref n-1 | cur | ext | td], D = 6 n + 13. Everything that is not a LiDAR factor enters as a quadratic
 * prior 0.5 d^T H0 d - b0^T d + c0 with d = x (-) x0 as in MarginalizationFactor::Evaluate
 * (marginalization_factor.cc:37-91; H0 = J0^T J0, b0 = -J0^T e0, c0 = 0.5 e0^T e0). */
typedef struct fflins_solve_options {
    int32_t  max_iterations[2];    /* 5, 10 (optimizer.cc:87-88)                                        */
    double   chi2_threshold;       /* 3.841 (optimizer.cc:516); <= 0: no cull between the two solves    */
    double   function_tolerance;   /* 1e-6  \                                                           */
    double   gradient_tolerance;   /* 1e-10  > Ceres defaults; 0 disables the early exit                */
    double   parameter_tolerance;  /* 1e-8  /                                                           */
    double   initial_radius;       /* 1e4                                                               */
    uint32_t flags;                /* FFLINS_HUBER | FFLINS_CONST_*                                      */
    uint32_t fixed_refs;           /* bit i: reference pose i is held constant                          */
} fflins_solve_options;
void fflins_solve_options_default(fflins_solve_options *o);
typedef struct fflins_solve_prior {
    const double *H0;              /* D x D, symmetric, row-major */
    const double *b0;              /* D                           */
    const double *x0;              /* 7 n + 15: linearisation point in the state layout */
    double        c0;
} fflins_solve_prior;
typedef struct fflins_solve_report {
    int32_t status;                /* 0 ok */
    int32_t iterations[2], successful[2], evaluations[2];
    double  cost_initial[2], cost_final[2];
    int32_t n_outliers[16], n_res[16];
    double  phase_cycles[8];       /* diagnostics: SM cycles of the solver CTA summed over the solve: 0 waiting for the
                                    * evaluations, 1 assembly, 2 damped copy, 3 factorisation + substitution, 4 model
                                    * decrease, 5 Plus */
} fflins_solve_report;
/* pose_refs [7 n], pose_cur, ext, td: in = initial state, out = solution. prior and report may be NULL. */
int fflins_window_solve(fflins_ctx *ctx, int32_t n_pairs, const uint64_t *ref_ids, double *pose_refs, double pose_cur[7],
                        double ext[7], double *td, const fflins_solve_options *opt, const fflins_solve_prior *prior,
                        fflins_solve_report *report);

/* ---- measurement ------------------------------------------------------------------------- */
/* Per-kernel CUDA-event timing on the context's stream (off by default). */
int fflins_profile_enable(fflins_ctx *ctx, int on);
int fflins_profile_reset(fflins_ctx *ctx);
/* names: '\n'-separated kernel names; ms/launches: arrays of `cap`. */
int fflins_profile_get(fflins_ctx *ctx, char *names, int32_t names_cap, double *ms, int64_t *launches, int32_t cap,
                       int32_t *n);
/* CUDA-event stopwatch on the context's stream: start records an event, stop records a second
 * one, waits for it and returns the elapsed device time in milliseconds. */
int fflins_timer_start(fflins_ctx *ctx);
int fflins_timer_stop(fflins_ctx *ctx, double *ms);
/* Total kernels launched by this context so far. */
int64_t fflins_launch_count(const fflins_ctx *ctx);
/* How fflins_window_eval / fflins_window_chi2 were served: by a message to the RESIDENT evaluation kernel (no launch;
 * the kernel is started on demand, polls a pinned mailbox and retires after FFLINS_RESIDENT_IDLE_US, default 1000) or
 * by an ordinary kernel launch (first use, profiling, FFLINS_NO_RESIDENT=1, or the resident kernel just retired). */
typedef struct fflins_eval_stats {
    int64_t resident_launches;  /* times the resident kernel was (re)started              */
    int64_t resident_evals;     /* evaluations / chi2 passes served through the mailbox   */
    int64_t launched_evals;     /* ... served by a kernel launch                          */
    int64_t resident_fallbacks; /* posted, but the kernel had retired: launched instead   */
} fflins_eval_stats;
int fflins_eval_stats_get(const fflins_ctx *ctx, fflins_eval_stats *out);
/* fp64 FMA throughput of the device in TFLOP/s (register-resident DFMA loop on every SM): the roofline of K6. */
int fflins_measure_fp64(fflins_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* FFLINS_H */
