/* C wrapper around the REFERENCE's own lidar::LidarMap (lidar/lidar_map.cc, compiled where it lies together with the
 * reference's vendored ikd-Tree, pointcloud.cc, misc.cc and rotation.cc: oracle/Makefile `ref` ->
 * oracle/_ref/libff_ref.so). It drives the class in the order ff_lins.cc uses it - lidarFrameProcessing per frame;
 * mergeNonKeyframes + addNewKeyFrame + updateLidarFeaturesInMap per keyframe; removeOldest / removeNewest - and hands the
 * clouds and LidarFeatures back as plain arrays. Library dependencies are stand-ins (oracle/refshim/): Eigen primitives,
 * PCL point / cloud containers and a VoxelGrid restated from PCL's algorithm, yaml-cpp as a value table, TBB run
 * serially. Test infrastructure only. */
#include "lidar/lidar_map.h"

#include "common/misc.h"

#include <yaml-cpp/yaml.h>

#include <cstring>

namespace {
struct Session {
    std::unique_ptr<lidar::LidarMap> map;
    std::vector<lidar::LidarFrame::Ptr> pending; /* frames since the last keyframe, oldest first */
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
}
