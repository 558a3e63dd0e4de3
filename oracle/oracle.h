/*
 * oracle.h — C interface of the CPU oracle. TEST INFRASTRUCTURE ONLY.
 *
 * The oracle restates, on the CPU and with zero dependencies, the arithmetic of the
 * FF-LINS LiDAR hot path (citations beside each function in oracle.cc). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it; the product library (ff-lins_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: the reference ships no tests or golden vectors for this path, and its third-party
 * libraries (Eigen, PCL, Ceres, TBB, glog, yaml-cpp) are absent here. Pinned against the reference's OWN
 * SOURCE FILES compiled where they lie under /root/reference into oracle/_ref/ (oracle/Makefile `ref`):
 *   (1) kNN: the vendored ikd-Tree (libikd_ref.so), bit-identical neighbours and distances;
 *   (2) lidar_factor.cc, residual_block_info.cc, marginalization_info.cc, marginalization_factor.cc, pose_manifold.cc, preintegration_base.cc, preintegration_normal.cc, lidar_converter.cc, misc.cc, rotation.cc, pointcloud.cc and lidar_map.cc + ikd-Tree (libff_ref.so, built against
 *       the stand-in headers of oracle/refshim/): the undistortion loop, frame poses, projections, merge, every
 *       output of LidarFactor::Evaluate and of the Huber corrector (ResidualBlockInfo::Evaluate), feature types, the five nearest points and the counters of
 *       updateLidarFeaturesInMap are bit-identical to this oracle; plane parameters agree to 1e-11 (the 5 x 3
 *       pivoted-QR solve is written independently on either side); tests/test_reference_sources.py,
 *       committed outputs tests/golden/ffref_*.npz.
 * Still "parity unpinned": what lives INSIDE the absent libraries - Eigen's own evaluation order of small products
 * and its QR / quaternion kernels (modelled in refshim/mini_eigen.h), PCL's VoxelGrid (restated, twice, from its
 * published algorithm: summation order inside a voxel is implementation-defined in PCL), ceres::HuberLoss' rho values.
 * Those are cross-checked against LAPACK / scipy / finite differences / closed forms (tests/test_oracle.py).
 */
#ifndef FFLINS_ORACLE_H
#define FFLINS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* pose: R row-major [9], t [3] */
int  orc_ins_window_index(const double *ins_t, int w, double time);
int  orc_pose_from_ins_window(const double *ins_t, const double *ins_p, const double *ins_q, int w,
                              const double *R_bl, const double *t_bl, double time, double *R, double *t);
void orc_state_to_pose(const double *p, const double *q, const double *R_bl, const double *t_bl, double *R,
                       double *t);
/* U5 per-point loop. Returns number of fallbacks. */
int  orc_undistort(const float *xyzi, const double *time, int n, const double *ins_t, const double *ins_p,
                   const double *ins_q, int w, const double *R_bl, const double *t_bl, double stamp, double td,
                   float *out_xyzi, double *pose_R, double *pose_t, int threads);
int  orc_decimate(const float *xyzi, int n, int filter_num, float *out_xyzi);
/* V1. keys (n, optional), order (n, optional: input indices sorted by (key,index)). Returns count or -1 on overflow. */
int  orc_voxel_filter(const float *xyzi, int n, double leaf, float *out_xyzi, int32_t *keys, int32_t *order);
void orc_project(const double *R, const double *t, const float *src, int n, float *dst, int threads);
void orc_relative_pose(const double *R0, const double *t0, const double *R1, const double *t1, double *R01,
                       double *t01);
/* A1 brute force. Returns count found (<=5). */
int  orc_knn5(const float *map_xyzi, int m, const float *query_xyz, double max_dist, int32_t *idx, float *d2);
/* A2. pts: 5 x xyzi. Returns is_plane. */
int  orc_plane_estimation(const float *pts, double threshold, double *unit_normal, double *norm_inverse,
                          double *distance);
/* A3 for one ref: outputs per query. knn_mode 0 = brute force, 1 = exact kd-tree (same results). */
void orc_associate(const double *ref_R, const double *ref_t, const float *cur_world, const float *cur_lidar, int q,
                   const float *map_xyzi, int m, double max_sq, double plane_thr, double sigma, double scalar_thr,
                   double sigma_gate, int knn_mode, int threads, int8_t *type, uint8_t *covered, int32_t *nn_idx,
                   float *nn_d2, double *normal, double *d, int32_t *num_features, int32_t *num_covered);
/* F1 single residual, Ceres calling convention. parameters = {pose_ref,pose_cur,ext,td}. */
int  orc_lidar_factor_evaluate(const float *point_xyz, const double *n, double d, double std, const double *motion,
                               const double *const *parameters, double *residuals, double **jacobians);
/* F4 corrector on (r, J[22]) in place; returns rho[0]. */
double orc_huber_correct(double delta, double *r, double *J22);
/* F1(+F4)(+F5) over a batch: valid[n] mask; r[n], J[22n] optional; H[361], b[19], cost[2] optional. */
void orc_factor_batch(int n, const float *point_xyz, const double *normal, const double *d, const double *std,
                      const uint8_t *valid, const double *motion, const double *pose_ref, const double *pose_cur,
                      const double *ext, double td, uint32_t flags, double huber_delta, double *r, double *J,
                      double *H, double *b, double *cost, int threads);
int  orc_chi2(int n, const float *point_xyz, const double *normal, const double *d, const double *std,
              const uint8_t *valid, const double *motion, const double *pose_ref, const double *pose_cur,
              const double *ext, double td, double threshold, uint8_t *outlier, int threads);
/* ---- CPU-baseline forms of the per-keyframe work (bench.py's reference arm): same arithmetic, the reference's own
 * life cycle and parallel structure. */
void *orc_tree_build(const float *map_xyzi, int m);     /* once per keyframe (lidar_map.cc:100-118) */
void orc_tree_free(void *tree);
void orc_associate_window(int n_refs, void *const *trees, const float *const *maps, const int *m, const double *ref_R,
                          const double *ref_t, const float *cur_world, const float *cur_lidar, int q, double max_sq,
                          double plane_thr, double sigma, double scalar_thr, double sigma_gate, int threads, int8_t *type,
                          uint8_t *covered, int32_t *nn_idx, float *nn_d2, double *normal, double *d, int32_t *num_features,
                          int32_t *num_covered);
void orc_window_eval(int n_pairs, int q, const float *point_xyz, const double *normal, const double *d, double std,
                     const uint8_t *valid, const double *motions, const double *pose_refs, const double *pose_cur,
                     const double *ext, double td, uint32_t flags, double huber_delta, double *H, double *b, double *cost,
                     int threads);
void orc_window_chi2(int n_pairs, int q, const float *point_xyz, const double *normal, const double *d, double std,
                     const uint8_t *valid, const double *motions, const double *pose_refs, const double *pose_cur,
                     const double *ext, double td, double threshold, uint8_t *outlier, int32_t *counts, int threads);
int  orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
