/*
 * fflins_host.h — C-ABI of the HOST side that sits either side of the LiDAR hot path:
 * INS mechanization and the INS window that feeds fflins_frame_process, IMU preintegration and the time-node
 * bookkeeping that produces each keyframe's velocity / angular rate, the sliding-window solver (IMU factors + the
 * 19x19 LiDAR blocks of fflins_window_eval + marginalization prior) that stands in for Ceres, and the converters /
 * writers of the reference's input and output formats. Plain CPU code (the reference runs all of this on the CPU
 * as well): libfflins_host.so has no CUDA dependency. Paths are relative to the upstream FF-LINS tree.
 *
 * Conventions as in fflins.h: quaternions x,y,z,w; pose block [tx,ty,tz,qx,qy,qz,qw]; mix block [v(3), bg(3), ba(3)].
 */
#ifndef FFLINS_HOST_H
#define FFLINS_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* common/types.h:66-72 (odovel unused by the LiDAR-inertial configuration) */
typedef struct fflins_imu {
    double time, dt;
    double dtheta[3], dvel[3];
    double odovel;
} fflins_imu;

/* common/integration_state.h:36-51, the part the normal preintegration uses */
typedef struct fflins_nav_state {
    double time;
    double p[3], q[4], v[3], bg[3], ba[3];
} fflins_nav_state;

/* common/integration_state.h IntegrationConfiguration: gravity in the navigation frame; earth rotation off (the
 * reference's LiDAR-inertial configs run `iswithearth: false`) */
typedef struct fflins_ins_config {
    double gravity[3];
} fflins_ins_config;

/* common/integration_state.h:60-75 IntegrationParameters */
typedef struct fflins_imu_noise {
    double acc_vrw, gyr_arw, gyr_bias_std, acc_bias_std, corr_time, gravity;
} fflins_imu_noise;

/* ---- INS mechanization (common/misc.cc:138-181) ------------------------------------------------------------------ */
void fflins_ins_mechanization(const fflins_ins_config *cfg, const fflins_imu *imu_pre, const fflins_imu *imu_cur,
                              fflins_nav_state *state);
/* MISC::isNeedInterpolation / imuInterpolation (misc.cc:240-284) */
int  fflins_imu_need_interpolation(const fflins_imu *imu0, const fflins_imu *imu1, double mid);
void fflins_imu_interpolation(const fflins_imu *imu01, fflins_imu *imu00, fflins_imu *imu11, double mid);

/* ---- the INS window: deque<pair<IMU, IntegrationState>> of LINS (ff_lins.h:114-115) ------------------------------ */
typedef struct fflins_ins_window fflins_ins_window;
fflins_ins_window *fflins_insw_create(const fflins_ins_config *cfg);
void fflins_insw_destroy(fflins_ins_window *w);
/* start the window at imu0 with a known state (the initial alignment of LINS::linsInitialization) */
void fflins_insw_reset(fflins_ins_window *w, const fflins_imu *imu0, const fflins_nav_state *state0);
/* LINS::doInsMechanization (ff_lins.cc:423-456): mechanize one more IMU epoch onto the back */
void fflins_insw_add_imu(fflins_ins_window *w, const fflins_imu *imu);
int  fflins_insw_size(const fflins_ins_window *w);
int  fflins_insw_back(const fflins_ins_window *w, fflins_imu *imu, fflins_nav_state *state);
/* the (t, p, q) arrays fflins_frame_process takes; returns the number of entries written (<= cap) */
int  fflins_insw_export(const fflins_ins_window *w, double *t, double *p, double *q, int cap);
/* MISC::redoInsMechanization (misc.cc:183-238): re-mechanize from an optimised state to the back, drop stale epochs */
int  fflins_insw_redo(fflins_ins_window *w, const fflins_nav_state *updated, int reserved_ins_num);
/* MISC::getImuSeriesFromTo (misc.cc:286-340); returns the count written or -1 */
int  fflins_insw_imu_series(const fflins_ins_window *w, double start, double end, fflins_imu *series, int cap);

/* ---- IMU preintegration, "normal" variant (preintegration_base.cc:39-84, preintegration_normal.cc:36-330) ------- */
typedef struct fflins_preint fflins_preint;
fflins_preint *fflins_preint_create(const fflins_imu_noise *noise, const fflins_imu *imu0, const fflins_nav_state *state);
void fflins_preint_destroy(fflins_preint *p);
void fflins_preint_add_imu(fflins_preint *p, const fflins_imu *imu);              /* addNewImu */
void fflins_preint_reintegrate(fflins_preint *p, const fflins_nav_state *state);  /* reintegration */
void fflins_preint_current_state(const fflins_preint *p, fflins_nav_state *out);
void fflins_preint_delta_state(const fflins_preint *p, fflins_nav_state *out);    /* delta p, q, v + linearisation biases */
double fflins_preint_delta_time(const fflins_preint *p);
/* evaluate + residualJacobian{Pose0,Mix0,Pose1,Mix1}: residual[15]; J row-major 15x7, 15x9, 15x7, 15x9 (any may be NULL).
 * cov / jac (optional, 225 each): the propagated covariance and bias Jacobian. */
void fflins_preint_evaluate(fflins_preint *p, const double pose0[7], const double mix0[9], const double pose1[7],
                            const double mix1[9], double residual[15], double *J_pose0, double *J_mix0, double *J_pose1,
                            double *J_mix1);
void fflins_preint_matrices(const fflins_preint *p, double *cov225, double *jac225);
/* LINS::addNewLidarFrameTimeNode (ff_lins.cc:474-489): mean angular rate over the last half LiDAR frame, bias removed */
void fflins_preint_frame_omega(const fflins_preint *p, double lidar_frame_dt, double imu_dt, const double bg[3],
                               double omega[3]);

/* ---- sliding-window solver (ff_lins/optimizer.cc:85-185, 221-435) ------------------------------------------------ */
/* The LiDAR share of the problem is evaluated by the caller: these are exactly the signatures of fflins_window_eval and
 * fflins_window_chi2 (include/fflins.h) with the context passed as `user`. */
typedef int (*fflins_lidar_eval_fn)(void *user, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                                    const double pose_cur[7], const double ext[7], double td, uint32_t flags, double *H,
                                    double *b, double *cost, int32_t *n_res);
typedef int (*fflins_lidar_chi2_fn)(void *user, int32_t n_pairs, const uint64_t *ref_ids, const double *pose_refs,
                                    const double pose_cur[7], const double ext[7], double td, double threshold,
                                    int32_t *n_outliers);
typedef struct fflins_solver fflins_solver;
typedef struct fflins_solver_options {
    int32_t first_iterations, second_iterations;   /* optimizer.cc:87-88: 5 and 10      */
    double  chi2_threshold;                        /* optimizer.cc:516: 3.841           */
    int32_t estimate_extrinsic, estimate_td;       /* optimizer.cc:418-435 (only once the window is full) */
    int32_t window_size;                           /* optimize_window_size              */
    double  min_speed_estimate_td;                 /* optimizer.h MINIMUM_SPEED_ESTIMATE_LIDAR_TD */
} fflins_solver_options;
typedef struct fflins_solver_summary {
    int32_t iterations[2], successful[2], outliers;
    double  cost_initial, cost_final;
} fflins_solver_summary;
void fflins_solver_options_default(fflins_solver_options *o);
fflins_solver *fflins_solver_create(const fflins_solver_options *o);
void fflins_solver_destroy(fflins_solver *s);
/* extrinsic_b_l_[8] of the optimizer: pose block + time delay */
void fflins_solver_set_extrinsic(fflins_solver *s, const double ext[7], double td);
void fflins_solver_get_extrinsic(const fflins_solver *s, double ext[7], double *td);
/* first node + the prior of Optimizer::marginalization's is_use_prior_ branch (ImuPosePriorFactor / ImuMixPriorFactor) */
void fflins_solver_set_first_node(fflins_solver *s, double time, const fflins_nav_state *state, const double pose_std[6],
                                  const double mix_std[9]);
/* LINS::addNewTimeNode (ff_lins.cc:491-526): the preintegration from the previous node to `time` is built by the caller
 * (fflins_insw_imu_series + fflins_preint_*); the solver takes ownership of it and appends the node. lidar_id: the
 * keyframe living at this node (its id in the fflins context), or UINT64_MAX. */
int  fflins_solver_add_node(fflins_solver *s, double time, fflins_preint *preint, uint64_t lidar_id);
/* attach a keyframe to an existing node (LINS::addNewLidarFrameTimeNode with is_add_node = false), or overwrite a node's
 * state (initial guess) */
int  fflins_solver_set_node_lidar(fflins_solver *s, int index, uint64_t lidar_id);
int  fflins_solver_set_node_state(fflins_solver *s, int index, const fflins_nav_state *state);
int  fflins_solver_num_nodes(const fflins_solver *s);
int  fflins_solver_get_node(const fflins_solver *s, int index, fflins_nav_state *state, uint64_t *lidar_id);
/* Optimizer::optimization: first solve, chi-square cull, second solve; re-integration of the preintegrations while the
 * window is not full (optimizer.cc:169-172). */
int  fflins_solver_optimize(fflins_solver *s, fflins_lidar_eval_fn eval, fflins_lidar_chi2_fn chi2, void *user,
                            fflins_solver_summary *out);
/* Optimizer::marginalization (optimizer.cc:221-404): the oldest node is eliminated; its IMU factor, the previous prior,
 * the initial priors and the LiDAR factors referenced to it are folded into the new prior. */
int  fflins_solver_marginalize(fflins_solver *s, fflins_lidar_eval_fn eval, void *user);
/* dense prior of the current window (testing): rows, cols, J0 (row-major), e0 */
int  fflins_solver_prior(const fflins_solver *s, int *rows, int *cols, double *J0, double *e0, int cap);

/* ---- converters and writers (ROS/lidar/lidar_converter.cc:36-232, ff_lins.cc:815-852, filesaver.cc:52-67) -------- */
/* One raw record per point -> the reference's 32-byte PointXYZIT records (lidar/lidar.h:30-38). Every converter drops
 * points outside [min_range, max_range] (squared-range test) and returns the number of records written.
 * Livox CustomMsg points: offset_time ns from the message stamp; tag filter (tag & 0x30) in {0x10, 0x00}; line < scan_line;
 * every `filter`-th point is NOT applied here (the reference decimates after undistortion). */
typedef struct fflins_livox_point {
    uint32_t offset_time;
    float x, y, z;
    uint8_t reflectivity, tag, line;
} fflins_livox_point;
int fflins_convert_livox(const fflins_livox_point *pts, int n, uint64_t timebase_ns, int scan_line, double min_range,
                         double max_range, int to_gps_time, void *out32, int cap, double *start, double *end);
/* Velodyne / Ouster / Hesai style: x, y, z, intensity, per-point time. time_mode 0: seconds relative to the message stamp
 * (Velodyne `time`), 1: nanoseconds relative to the stamp (Ouster `t`), 2: absolute seconds (Hesai `timestamp`). */
int fflins_convert_generic(const float *xyzi, const double *time, int n, double stamp, int time_mode, double min_range,
                           double max_range, int to_gps_time, void *out32, int cap, double *start, double *end);
/* trajectory.csv line (ff_lins.cc:815-852 writeNavResult): "%-15.9lf" time, position, quaternion x y z w; returns bytes */
int fflins_format_trajectory(const fflins_nav_state *s, char *buf, int cap);
/* imu_error line of writeNavResult: time, gyro bias deg/h (3), accelerometer bias mGal (3), odometer scale */
int fflins_format_imu_error(const fflins_nav_state *s, char *buf, int cap);
/* lidar_statistic.txt line from LidarMap::statisticParameters (lidar_map.cc:484-497) */
int fflins_format_statistics(const double *values, int n, char *buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* FFLINS_HOST_H */
