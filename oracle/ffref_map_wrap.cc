/* C wrapper around the REFERENCE's own lidar::LidarMap (lidar/lidar_map.cc, compiled where it lies together with the
 * reference's vendored ikd-Tree, pointcloud.cc, misc.cc and rotation.cc: oracle/Makefile `ref` ->
 * oracle/_ref/libff_ref.so). It drives the class in the order ff_lins.cc uses it - lidarFrameProcessing per frame;
 * mergeNonKeyframes + addNewKeyFrame + updateLidarFeaturesInMap per keyframe; removeOldest / removeNewest - and hands the
 * clouds and LidarFeatures back as plain arrays. Library dependencies are stand-ins (oracle/refshim/): Eigen primitives,
 * PCL point / cloud containers and a VoxelGrid restated from PCL's algorithm, yaml-cpp as a value table, TBB run
 * serially. Test infrastructure only. */
#include "lidar/lidar_map.h"

#include "common/misc.h"
#include "factors/lidar_factor.h"

#include <yaml-cpp/yaml.h>

#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
struct FactorRef {
    std::unique_ptr<LidarFactor> factor;
    lidar::LidarFeature::Ptr feature;
    int pair;
};
struct Session {
    std::unique_ptr<lidar::LidarMap> map;
    std::vector<lidar::LidarFrame::Ptr> pending; /* frames since the last keyframe, oldest first */
    std::vector<FactorRef> factors;               /* the residual blocks of the current solve (addLidarFactors) */
};
typedef std::deque<std::pair<IMU, IntegrationState>> InsWindow;

Pose make_pose(const double *R, const double *t) {
    Pose p;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) p.R(r, c) = R[3 * r + c];
    p.t = Vector3d(t[0], t[1], t[2]);
    return p;
}
int copy_cloud(const PointCloudPtr &c, float *out, int cap) {
    if (!c) return 0;
    int n = (int) c->size();
    if (out)
        for (int k = 0; k < n && k < cap; k++) {
            out[4 * k] = c->points[k].x, out[4 * k + 1] = c->points[k].y, out[4 * k + 2] = c->points[k].z;
            out[4 * k + 3] = c->points[k].intensity;
        }
    return n;
}
} // namespace

extern "C" {

void *ffref_map_create(double point_to_plane_std, int window_size, double downsample_size, double frame_rate,
                       double plane_threshold, double min_keyframe_translation, int point_filter_num) {
    auto &t                                    = YAML::fflins_table();
    t["optimizer.optimize_point_to_plane_std"] = point_to_plane_std;
    t["optimizer.optimize_window_size"]        = window_size;
    t["lidar.downsample_size"]                 = downsample_size;
    t["lidar.frame_rate"]                      = frame_rate;
    t["lidar.plane_estimation_threshold"]      = plane_threshold;
    t["lidar.minimum_keyframe_translation"]    = min_keyframe_translation;
    t["lidar.point_filter_num"]                = point_filter_num;
    t["lidar.is_save_pointcloud"]              = 0;
    auto *s                                    = new Session;
    s->map.reset(new lidar::LidarMap("", ""));
    return s;
}
void ffref_map_destroy(void *h) { delete static_cast<Session *>(h); }

/* worker threads of the OpenMP build (launchers such as torchrun export OMP_NUM_THREADS=1); returns what is in effect */
int ffref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void) n;
    return 1;
#endif
}

/* LidarMap::lidarFrameProcessing(ins_window, pose_b_l, frame) (lidar_map.cc:213-273) on a new LidarFrame; the frame joins
 * the pending list. Returns the number of points of the frame's down-sampled cloud. */
int ffref_map_frame(void *h, const float *xyzi, const double *time, int n, const double *ins_t, const double *ins_p,
                    const double *ins_q, int w, const double *R_bl, const double *t_bl, double stamp, double td) {
    auto *s = static_cast<Session *>(h);
    PointCloudCustomPtr raw(new PointCloudCustom);
    raw->resize(n);
    for (int k = 0; k < n; k++) {
        auto &p     = raw->points[k];
        p.x         = xyzi[4 * k];
        p.y         = xyzi[4 * k + 1];
        p.z         = xyzi[4 * k + 2];
        p.data[3]   = 1.0f;
        p.intensity = xyzi[4 * k + 3];
        p.time      = time[k];
    }
    InsWindow win;
    for (int k = 0; k < w; k++) {
        IMU imu;
        imu.time = ins_t[k];
        imu.dt   = 0;
        IntegrationState st;
        st.time = ins_t[k];
        st.p    = Vector3d(ins_p[3 * k], ins_p[3 * k + 1], ins_p[3 * k + 2]);
        st.q    = Quaterniond(ins_q[4 * k + 3], ins_q[4 * k], ins_q[4 * k + 1], ins_q[4 * k + 2]);
        win.emplace_back(imu, st);
    }
    auto frame = lidar::LidarFrame::createFrame(stamp, stamp, raw);
    frame->setTimeDelay(td);
    s->map->lidarFrameProcessing(win, make_pose(R_bl, t_bl), frame);
    s->pending.push_back(frame);
    return (int) frame->pointCloudLidar()->size();
}

/* the newest pending frame becomes a keyframe: mergeNonKeyframes(older pending frames, it) (lidar_map.cc:190-211),
 * addNewKeyFrame (:100-124) and, from the second keyframe on, updateLidarFeaturesInMap (:426-459). Returns the window
 * size. */
int ffref_map_keyframe(void *h) {
    auto *s  = static_cast<Session *>(h);
    auto key = s->pending.back();
    s->pending.pop_back();
    key->setKeyFrame();
    s->map->mergeNonKeyframes(s->pending, key);
    s->pending.clear();
    s->map->addNewKeyFrame(key);
    if (s->map->keyframes().size() > 1) s->map->updateLidarFeaturesInMap();
    return (int) s->map->keyframes().size();
}
void ffref_map_update_features(void *h) { static_cast<Session *>(h)->map->updateLidarFeaturesInMap(); }
void ffref_map_remove_oldest(void *h) { static_cast<Session *>(h)->map->removeOldestLidarFrame(); }
void ffref_map_remove_newest(void *h) { static_cast<Session *>(h)->map->removeNewestLidarFrame(); }
int ffref_map_window(void *h) { return (int) static_cast<Session *>(h)->map->keyframes().size(); }

void ffref_map_set_pose(void *h, int key, const double *R, const double *t) {
    static_cast<Session *>(h)->map->keyframes()[key]->setPose(make_pose(R, t));
}
void ffref_map_get_pose(void *h, int key, double *R, double *t) {
    const Pose &p = static_cast<Session *>(h)->map->keyframes()[key]->pose();
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[3 * r + c] = p.R(r, c);
    for (int i = 0; i < 3; i++) t[i] = p.t[i];
}

/* LidarMap::keyframeSelection (lidar_map.cc:75-98) for a candidate frame with the given pose and stamp against the newest
 * keyframe; stats = {interval, translation, rotation [rad]} as the reference records them (statisticParameters). */
int ffref_map_keyframe_selection(void *h, const double *R, const double *t, double stamp, double *stats) {
    auto *s = static_cast<Session *>(h);
    PointCloudCustomPtr raw(new PointCloudCustom);
    auto frame = lidar::LidarFrame::createFrame(stamp, stamp, raw);
    frame->setPose(make_pose(R, t));
    bool is_key = s->map->keyframeSelection(frame);
    const auto &sp = s->map->statisticParameters();
    stats[0] = sp[4];
    stats[1] = sp[5];
    stats[2] = sp[6] / R2D;
    return is_key ? 1 : 0;
}

/* which: 0 lidar (down-sampled), 1 world, 2 map, 3 map full, 4 lidar full (decimated). Returns the point count. */
int ffref_map_cloud(void *h, int key, int which, float *out, int cap) {
    auto f = static_cast<Session *>(h)->map->keyframes()[key];
    switch (which) {
        case 0: return copy_cloud(f->pointCloudLidar(), out, cap);
        case 1: return copy_cloud(f->pointCloudWorld(), out, cap);
        case 2: return copy_cloud(f->pointCloudMap(), out, cap);
        case 3: return copy_cloud(f->pointCloudMapFull(), out, cap);
        default: return copy_cloud(f->pointCloudLidarFull(), out, cap);
    }
}

/* LidarFeatures stored on reference keyframe `key` for the newest keyframe's points: type (LidarFeature::FEATURE_*),
 * unit normal, norm inverse, the nearest points found (up to 5), std and outlier flag. Returns the feature count
 * (call with type == NULL to size the arrays). */
int ffref_map_features(void *h, int key, int32_t *type, double *normal, double *norm_inverse, float *nearest_xyzi,
                       int32_t *n_nearest, double *point_std, uint8_t *outlier) {
    auto f   = static_cast<Session *>(h)->map->keyframes()[key];
    auto &ft = f->features();
    if (!type) return (int) ft.size();
    for (size_t k = 0; k < ft.size(); k++) {
        type[k]         = (int32_t) ft[k]->featureType();
        norm_inverse[k] = ft[k]->normInverse();
        for (int i = 0; i < 3; i++) normal[3 * k + i] = ft[k]->unitNormalVector()[i];
        const auto &np = ft[k]->nearestPoints();
        n_nearest[k]   = (int32_t) np.size();
        for (size_t j = 0; j < np.size() && j < 5; j++) {
            float *o = nearest_xyzi + 20 * k + 4 * j;
            o[0] = np[j].x, o[1] = np[j].y, o[2] = np[j].z, o[3] = np[j].intensity;
        }
        point_std[k] = ft[k]->pointStd();
        outlier[k]   = ft[k]->isOutlier() ? 1 : 0;
    }
    return (int) ft.size();
}
void ffref_map_counts(void *h, int32_t *num_features, int32_t *num_covered) {
    auto cur      = static_cast<Session *>(h)->map->keyframes().back();
    *num_features = cur->numFeatures();
    *num_covered  = cur->numCovered();
}

/* ---- the solver-facing part, as Optimizer drives it (ff_lins/optimizer.cc) ------------------------------------------- */

/* LidarFrame::setVelocity / setAngularVelocity (ff_lins.cc sets them when a keyframe's time node is created) */
void ffref_map_set_motion(void *h, int key, const double *v, const double *omega) {
    auto f = static_cast<Session *>(h)->map->keyframes()[key];
    f->setVelocity(Vector3d(v[0], v[1], v[2]));
    f->setAngularVelocity(Vector3d(omega[0], omega[1], omega[2]));
}

/* Optimizer::addLidarFactors (optimizer.cc:437-478): one LidarFactor per PLANE, non-outlier feature of every
 * (reference, newest) pair, built from the frames' velocity / angular velocity / time delay. Returns their number. */
int ffref_map_add_factors(void *h) {
    auto *s = static_cast<Session *>(h);
    s->factors.clear();
    const auto &keys = s->map->keyframes();
    auto cur         = keys.back();
    const auto &points_lidar = cur->pointCloudLidar()->points;
    for (size_t i = 0; i + 1 < keys.size(); i++) {
        auto ref             = keys[i];
        const auto &features = ref->features();
        for (size_t k = 0; k < features.size(); k++) {
            const auto &feature = features[k];
            if (!feature || (feature->featureType() != lidar::LidarFeature::FEATURE_PLANE) || feature->isOutlier()) continue;
            FactorRef fr;
            fr.factor.reset(new LidarFactor(points_lidar[k], feature->unitNormalVector(), feature->normInverse(),
                                            feature->pointStd(), ref->velocity(), ref->angularVelocity(), ref->timeDelay(),
                                            cur->velocity(), cur->angularVelocity(), cur->timeDelay()));
            fr.feature = feature;
            fr.pair    = (int) i;
            s->factors.push_back(std::move(fr));
        }
    }
    return (int) s->factors.size();
}

/* What one solver iteration asks of the LiDAR residual blocks: LidarFactor::Evaluate (residual + the four Jacobian
 * blocks) of every block, the robust-loss correction Ceres applies (HuberLoss(huber_delta) through the corrector of
 * residual_block_info.cc:49-77, restated inline; huber_delta <= 0: none) and the accumulation of J^T J, J^T r and the
 * cost per pair over the 19 tangent columns (6 + 6 + 6 + 1: the quaternion's fourth column is zero). Parallel over the
 * blocks like Ceres' evaluator (num_threads). pose_refs: n x 7; H: n x 19 x 19; b: n x 19; cost: n. */
void ffref_map_evaluate_factors(void *h, const double *pose_refs, const double *pose_cur, const double *ext, double td,
                                double huber_delta, double *H, double *b, double *cost) {
    auto *s       = static_cast<Session *>(h);
    const int n   = (int) s->map->keyframes().size() - 1;
    const size_t m = s->factors.size();
    static const int col[19] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 21};
    ceres::HuberLoss loss(huber_delta > 0 ? huber_delta : 1.0);
    std::memset(H, 0, sizeof(double) * n * 361);
    std::memset(b, 0, sizeof(double) * n * 19);
    std::memset(cost, 0, sizeof(double) * n);
#pragma omp parallel
    {
        std::vector<double> Hl((size_t) n * 361, 0.0), bl((size_t) n * 19, 0.0), cl(n, 0.0);
#pragma omp for schedule(static)
        for (size_t k = 0; k < m; k++) {
            const FactorRef &fr        = s->factors[k];
            const double *parameters[4] = {pose_refs + 7 * fr.pair, pose_cur, ext, &td};
            double r, J[22];
            double *jac[4] = {J, J + 7, J + 14, J + 21};
            fr.factor->Evaluate(parameters, &r, jac);
            double rho0 = r * r;
            if (huber_delta > 0) {
                double rho[3];
                const double sq_norm = r * r;
                loss.Evaluate(sq_norm, rho);
                rho0                   = rho[0];
                const double sqrt_rho1 = std::sqrt(rho[1]);
                double residual_scaling, alpha_sq_norm;
                if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
                    residual_scaling = sqrt_rho1;
                    alpha_sq_norm    = 0.0;
                } else {
                    const double D     = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
                    const double alpha = 1.0 - std::sqrt(D);
                    residual_scaling   = sqrt_rho1 / (1 - alpha);
                    alpha_sq_norm      = alpha / sq_norm;
                }
                for (int i = 0; i < 22; i++) J[i] = sqrt_rho1 * (J[i] - alpha_sq_norm * r * (r * J[i]));
                r *= residual_scaling;
            }
            double v[19];
            for (int i = 0; i < 19; i++) v[i] = J[col[i]];
            double *Hp = Hl.data() + (size_t) fr.pair * 361, *bp = bl.data() + (size_t) fr.pair * 19;
            for (int i = 0; i < 19; i++) {
                for (int j = i; j < 19; j++) Hp[19 * i + j] += v[i] * v[j];
                bp[i] -= v[i] * r;
            }
            cl[fr.pair] += 0.5 * rho0;
        }
#pragma omp critical
        {
            for (size_t i = 0; i < Hl.size(); i++) H[i] += Hl[i];
            for (size_t i = 0; i < bl.size(); i++) b[i] += bl[i];
            for (int i = 0; i < n; i++) cost[i] += cl[i];
        }
    }
    for (int p = 0; p < n; p++)
        for (int i = 0; i < 19; i++)
            for (int j = 0; j < i; j++) H[p * 361 + 19 * i + j] = H[p * 361 + 19 * j + i];
}

/* Optimizer::lidarOutlierCullingByChi2 (optimizer.cc:515-548): residual-only Evaluate of every block, setOutlier() above
 * the threshold (tbb::parallel_for in the reference). Returns the number of new outliers; per_pair (n) optional. */
int ffref_map_chi2(void *h, const double *pose_refs, const double *pose_cur, const double *ext, double td, double threshold,
                   int32_t *per_pair) {
    auto *s        = static_cast<Session *>(h);
    const size_t m = s->factors.size();
    const int n    = (int) s->map->keyframes().size() - 1;
    std::vector<uint8_t> is_outlier(m, 0);
    auto culling_function = [&](const tbb::blocked_range<size_t> &range) {
        for (size_t k = range.begin(); k != range.end(); k++) {
            const FactorRef &fr         = s->factors[k];
            const double *parameters[4] = {pose_refs + 7 * fr.pair, pose_cur, ext, &td};
            double residual;
            fr.factor->Evaluate(parameters, &residual, nullptr);
            double cost = residual * residual;
            if (cost > threshold) {
                fr.feature->setOutlier();
                is_outlier[k] = 1;
            }
        }
    };
    tbb::parallel_for(tbb::blocked_range<size_t>(0, m), culling_function);
    int total = 0;
    if (per_pair) std::memset(per_pair, 0, sizeof(int32_t) * n);
    for (size_t k = 0; k < m; k++)
        if (is_outlier[k]) {
            total++;
            if (per_pair) per_pair[s->factors[k].pair]++;
        }
    return total;
}

/* the world clouds of the window after the solver moved the poses (ff_lins.cc updateParametersFromOptimizer:
 * pointCloudProjection(pose, lidar, world) per keyframe) */
void ffref_map_reproject(void *h) {
    for (auto &f : static_cast<Session *>(h)->map->keyframes())
        lidar::PointCloudCommon::pointCloudProjection(f->pose(), f->pointCloudLidar(), f->pointCloudWorld());
}
}
