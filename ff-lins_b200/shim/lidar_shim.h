// lidar_shim.h — header-only C++ mirror of the reference's LiDAR interfaces over the C-ABI
// (include/fflins.h). Same class names, method names, argument meaning and return conventions as
//   lidar::LidarFeature      ff_lins/lidar/lidar_feature.h:35-114
//   lidar::LidarFrame        ff_lins/lidar/lidar_frame.h:36-203
//   lidar::PointCloudCommon  ff_lins/lidar/pointcloud.h:31-50
//   lidar::LidarMap          ff_lins/lidar/lidar_map.h:40-118
//   LidarFactor              ff_lins/factors/lidar_factor.h:31-51
// so that LINS (ff_lins/ff_lins/ff_lins.cc) and Optimizer (ff_lins/ff_lins/optimizer.cc) compile
// against it unchanged. All point data lives on the GPU; host clouds / features are materialised
// lazily when an accessor asks for them. No CPU fallback: every operation is a C-ABI call.
#pragma once
#include "../../include/fflins.h"
#include "fflins_types.h"

#include <cstdio>
#include <deque>
#include <fstream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>

namespace fflins_shim {

// process-wide default session: set by the first LidarMap, used by the entry points whose reference
// signatures carry no handle (LidarFactor's constructor, PointCloudCommon's statics)
inline fflins_ctx *&default_ctx() {
    static fflins_ctx *c = nullptr;
    return c;
}

// Ceres evaluates residual blocks from several threads (optimizer.cc:103, num_threads = 8); calls on one context are
// serialised (include/fflins.h), so every adapter takes this lock around its C-ABI call.
inline std::mutex &ctx_mutex() {
    static std::mutex m;
    return m;
}

inline fflins_pose toAbi(const Pose &p) {
    fflins_pose o;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o.R[3 * r + c] = p.R(r, c);
    for (int i = 0; i < 3; i++) o.t[i] = p.t[i];
    return o;
}
inline Pose fromAbi(const fflins_pose &p) {
    Pose o;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) o.R(r, c) = p.R[3 * r + c];
    for (int i = 0; i < 3; i++) o.t[i] = p.t[i];
    return o;
}
inline void check(fflins_ctx *ctx, int rc, const char *what) {
    if (rc != FFLINS_OK) {
        std::fprintf(stderr, "[fflins] %s failed (%d): %s\n", what, rc, ctx ? fflins_last_error(ctx) : "no context");
        throw std::runtime_error(std::string("fflins: ") + what);
    }
}
// the keys LidarMap::LidarMap reads (lidar_map.cc:45-55); flat "key: value" scan of the yaml file
inline fflins_config configFromYaml(const std::string &path) {
    fflins_config c;
    fflins_config_default(&c);
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        auto colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string key = line.substr(0, colon), val = line.substr(colon + 1);
        key.erase(0, key.find_first_not_of(" \t"));
        key.erase(key.find_last_not_of(" \t") + 1);
        std::istringstream vs(val);
        double d;
        if (!(vs >> d)) continue;
        if (key == "optimize_point_to_plane_std") c.point_to_plane_std = d;
        else if (key == "optimize_window_size") c.window_size = (int32_t) d;
        else if (key == "downsample_size") c.downsample_size = d;
        else if (key == "frame_rate") c.frame_rate = d;
        else if (key == "plane_estimation_threshold") c.plane_estimation_threshold = d;
        else if (key == "minimum_keyframe_translation") c.minimum_keyframe_translation = d;
        else if (key == "point_filter_num") c.point_filter_num = (int32_t) d;
    }
    return c;
}

} // namespace fflins_shim

namespace lidar {

class LidarFeature {
public:
    typedef std::shared_ptr<LidarFeature> Ptr;
    enum FeatureType { FEATURE_NONE = -1, FEATURE_PLANE = 0, FEATURE_CORNER = 1 };

    LidarFeature() = default;
    LidarFeature(PointVector nearest_points, Vector3d unit_normal_vector, double norm_inverse, FeatureType type)
        : nearest_points_(std::move(nearest_points)), unit_normal_vector_(unit_normal_vector), norm_inverse_(norm_inverse),
          type_(type) {}

    FeatureType featureType() { return type_; }
    void updateFeature(PointVector nearest_points, Vector3d unit_normal_vector, double norm_inverse, FeatureType type,
                       bool is_outlier) {
        nearest_points_     = std::move(nearest_points);
        unit_normal_vector_ = unit_normal_vector;
        norm_inverse_       = norm_inverse;
        type_               = type;
        is_outlier_         = is_outlier;
    }
    const PointVector &nearestPoints() { return nearest_points_; }
    const Vector3d &unitNormalVector() { return unit_normal_vector_; }
    double normInverse() const { return norm_inverse_; }
    // setOutlier marks the host mirror; LidarMap::pushOutliers() sends the flags to the device
    void setOutlier() {
        is_outlier_ = true;
        outlier_counts_++;
    }
    bool isOutlier() const { return is_outlier_; }
    int outlierCounts() const { return outlier_counts_; }
    void setPointStd(double std) { point_std_ = std; }
    double pointStd() { return point_std_; }

private:
    PointVector nearest_points_;
    Vector3d unit_normal_vector_;
    double norm_inverse_ = 0;
    double point_std_    = 0;
    bool is_outlier_{false};
    int outlier_counts_{0};
    FeatureType type_{FEATURE_NONE};
};

class LidarFrame {
public:
    typedef std::shared_ptr<LidarFrame> Ptr;
    typedef unsigned long ulong_t;

    LidarFrame() = delete;
    LidarFrame(ulong_t id, double start_time, double end_time, PointCloudCustomPtr pointcloud)
        : raw_pointcloud_(std::move(pointcloud)), stamp_(end_time), start_time_(start_time), end_time_(end_time), id_(id) {}

    static LidarFrame::Ptr createFrame(double start_time, double end_time, PointCloudCustomPtr pointcloud) {
        static ulong_t factory_id = 0;
        return std::make_shared<LidarFrame>(factory_id++, start_time, end_time, pointcloud);
    }

    ulong_t id() { return id_; }
    const Pose &pose() { return pose_; }
    // setPose also tells the device (optimizer.cc:212-213); the world cloud is re-projected by
    // LidarMap::reprojectWorld / PointCloudCommon::pointCloudProjection as in the reference
    void setPose(Pose pose) {
        pose_ = std::move(pose);
        if (ctx_) {
            fflins_pose p = fflins_shim::toAbi(pose_);
            fflins_frame_set_pose(ctx_, id_, &p);
        }
    }
    const Vector3d &velocity() { return velocity_; }
    void setVelocity(Vector3d velocity) {
        velocity_ = velocity;
        pushMotion();
    }
    const Vector3d &angularVelocity() { return omega_; }
    void setAngularVelocity(Vector3d omega) {
        omega_ = omega;
        pushMotion();
    }
    double timeDelay() const { return td_; }
    void setTimeDelay(double td) {
        td_ = td;
        pushMotion();
    }
    double stamp() const { return stamp_; }
    void setStamp(double stamp) { stamp_ = stamp; }
    double startTime() const { return start_time_ + td_; }
    double endTime() const { return end_time_ + td_; }

    PointCloudCustomPtr rawPointCloud() const { return raw_pointcloud_; }
    PointCloudPtr pointCloudLidar() { return cloud(FFLINS_CLOUD_LIDAR); }
    PointCloudPtr pointCloudWorld() { return cloud(FFLINS_CLOUD_WORLD); }
    PointCloudPtr pointCloudLidarFull() { return cloud(FFLINS_CLOUD_LIDAR_FULL); }
    PointCloudPtr pointCloudMapFull() { return cloud(FFLINS_CLOUD_MAP_FULL); }
    PointCloudPtr pointCloudMap() { return cloud(FFLINS_CLOUD_MAP); }

    // features stored on this (reference) frame, indexed by the current keyframe's points
    std::vector<LidarFeature::Ptr> &features() {
        if (features_dirty_ && ctx_) {
            int32_t q = 0;
            fflins_shim::check(ctx_, fflins_features_download(ctx_, id_, 0, &q, 0, 0, 0, 0, 0, 0, 0, 0), "features size");
            std::vector<int8_t> type(q);
            std::vector<uint8_t> outlier(q);
            std::vector<float> pts((size_t) 20 * q);
            std::vector<double> normal((size_t) 3 * q), d(q);
            if (q > 0)
                fflins_shim::check(ctx_,
                                   fflins_features_download(ctx_, id_, q, &q, type.data(), nullptr, outlier.data(), nullptr,
                                                            nullptr, pts.data(), normal.data(), d.data()),
                                   "features download");
            features_.resize(q);
            for (int k = 0; k < q; k++) {
                PointVector nn(5);
                for (int j = 0; j < 5; j++) {
                    const float *p  = &pts[(size_t) 20 * k + 4 * j];
                    nn[j].x         = p[0];
                    nn[j].y         = p[1];
                    nn[j].z         = p[2];
                    nn[j].intensity = p[3];
                }
                auto f = std::make_shared<LidarFeature>();
                f->updateFeature(std::move(nn), Vector3d(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]), d[k],
                                 (LidarFeature::FeatureType) type[k], outlier[k] != 0);
                f->setPointStd(point_std_);
                features_[k] = f;
            }
            features_dirty_ = false;
        }
        return features_;
    }

    void setFrameInMap() { is_in_map_ = true; }
    bool isFrameInMap() { return is_in_map_; }
    void setKeyFrame() { is_keyframe_ = true; }
    bool isKeyFrame() const { return is_keyframe_; }
    int numFeatures() const { return num_features_; }
    void setNumFeatures(int num_features) { num_features_ = num_features; }
    int numCovered() const { return num_covered_; }
    void setNumCovered(int num_covered) { num_covered_ = num_covered; }

    // ---- binding to the device session (used by LidarMap) ----
    // `alive` is the owning LidarMap's life token: a frame may outlive its map (LINS keeps frames in its own containers)
    void bind(fflins_ctx *ctx, double point_std, std::shared_ptr<bool> alive = nullptr) {
        ctx_       = ctx;
        point_std_ = point_std;
        if (ctx) {   // (un-binding keeps the session the raw buffer was queued on: the destructor still has to wait for its copy)
            sess_  = ctx;
            alive_ = std::move(alive);
        }
    }
    void invalidate(bool clouds, bool features) {
        if (clouds)
            for (auto &c : cached_) c.reset();
        if (features) features_dirty_ = true;
    }
    void adoptPose(const fflins_pose &p) { pose_ = fflins_shim::fromAbi(p); }
    fflins_ctx *context() const { return ctx_; }
    // the raw cloud's storage is page-locked once, so that the frame can be QUEUED (out = NULL) and its copy run at PCIe
    // rate behind the caller's back; unlocked again when the frame goes away
    bool pinRaw() {
        if (!raw_pinned_ && raw_pointcloud_ && !raw_pointcloud_->points.empty())
            raw_pinned_ = fflins_host_register(raw_pointcloud_->points.data(), raw_pointcloud_->points.size() * sizeof(PointTypeCustom)) == FFLINS_OK;
        return raw_pinned_;
    }
    ~LidarFrame() {
        if (raw_pinned_) {
            if (sess_ && alive_ && *alive_) fflins_host_buffer_done(sess_, raw_pointcloud_->points.data());   // a queued copy may still read it
            fflins_host_unregister(raw_pointcloud_->points.data());
        }
    }

private:
    void pushMotion() {
        if (ctx_) fflins_frame_set_motion(ctx_, id_, velocity_.data(), omega_.data(), td_);
    }
    PointCloudPtr cloud(int which) {
        if (!cached_[which]) {
            auto pc = PointCloudPtr(new PointCloud);
            if (ctx_) {
                int32_t n = 0;
                fflins_shim::check(ctx_, fflins_frame_download(ctx_, id_, which, nullptr, 0, &n), "cloud size");
                std::vector<float> buf((size_t) 4 * n);
                if (n > 0) fflins_shim::check(ctx_, fflins_frame_download(ctx_, id_, which, buf.data(), n, &n), "cloud download");
                pc->resize(n);
                for (int i = 0; i < n; i++) {
                    pc->points[i].x         = buf[4 * i];
                    pc->points[i].y         = buf[4 * i + 1];
                    pc->points[i].z         = buf[4 * i + 2];
                    pc->points[i].intensity = buf[4 * i + 3];
                }
            }
            cached_[which] = pc;
        }
        return cached_[which];
    }

    PointCloudCustomPtr raw_pointcloud_;
    PointCloudPtr cached_[5];
    std::vector<LidarFeature::Ptr> features_;
    bool features_dirty_ = false;
    int num_features_ = 0, num_covered_ = 0;
    Pose pose_;
    Vector3d velocity_, omega_;
    double stamp_;
    double start_time_, end_time_;
    double td_{0};
    double point_std_ = 0.1;
    bool is_in_map_{false};
    bool is_keyframe_{false};
    bool raw_pinned_{false};
    ulong_t id_;
    fflins_ctx *ctx_ = nullptr, *sess_ = nullptr;
    std::shared_ptr<bool> alive_;
};

class PointCloudCommon {
public:
    static constexpr int PLANE_ESTIMATION_POINT_SZIE = 5;

    // PointCloudCommon::pointCloudProjection (pointcloud.cc:38-51); dst may alias src
    static void pointCloudProjection(fflins_ctx *ctx, const Pose &pose, PointCloudPtr src, PointCloudPtr dst) {
        size_t n = src->size();
        std::vector<float> buf(4 * n);
        for (size_t i = 0; i < n; i++) {
            buf[4 * i]     = src->points[i].x;
            buf[4 * i + 1] = src->points[i].y;
            buf[4 * i + 2] = src->points[i].z;
            buf[4 * i + 3] = src->points[i].intensity;
        }
        fflins_pose p = fflins_shim::toAbi(pose);
        fflins_shim::check(ctx, fflins_cloud_projection(ctx, &p, buf.data(), (int32_t) n, buf.data()), "pointCloudProjection");
        if (dst != src) *dst = *src;
        for (size_t i = 0; i < n; i++) {
            dst->points[i].x = buf[4 * i];
            dst->points[i].y = buf[4 * i + 1];
            dst->points[i].z = buf[4 * i + 2];
        }
    }
    // the reference's exact signatures (no handle): default session
    static void pointCloudProjection(const Pose &pose, PointCloudPtr src, PointCloudPtr dst) {
        pointCloudProjection(fflins_shim::default_ctx(), pose, src, dst);
    }
    static bool planeEstimation(const PointVector &points, double threshold, Vector3d &unit_normal_vector,
                                double &norm_inverse, std::vector<double> &distance) {
        return planeEstimation(fflins_shim::default_ctx(), points, threshold, unit_normal_vector, norm_inverse, distance);
    }
    // PointCloudCommon::planeEstimation (pointcloud.cc:61-93)
    static bool planeEstimation(fflins_ctx *ctx, const PointVector &points, double threshold, Vector3d &unit_normal_vector,
                                double &norm_inverse, std::vector<double> &distance) {
        if (points.size() < (size_t) PLANE_ESTIMATION_POINT_SZIE) return false;
        float pts[20];
        for (int k = 0; k < 5; k++) {
            pts[4 * k]     = points[k].x;
            pts[4 * k + 1] = points[k].y;
            pts[4 * k + 2] = points[k].z;
            pts[4 * k + 3] = points[k].intensity;
        }
        double unit[3], dist[5];
        uint8_t ok = 0;
        fflins_shim::check(ctx, fflins_plane_estimation(ctx, pts, 1, threshold, unit, &norm_inverse, dist, &ok), "planeEstimation");
        unit_normal_vector = Vector3d(unit[0], unit[1], unit[2]);
        distance.assign(dist, dist + 5);
        return ok != 0;
    }
};

class LidarMap {
public:
    typedef std::shared_ptr<LidarMap> Ptr;

    LidarMap() = delete;
    // LidarMap(configfile, outputpath) (lidar_map.cc:42): the yaml keys are read by configFromYaml
    LidarMap(const std::string &configfile, const std::string & /*outputpath*/, int device = 0)
        : LidarMap(fflins_shim::configFromYaml(configfile), device) {}
    explicit LidarMap(const fflins_config &cfg, int device = 0) : cfg_(cfg) {
        int rc = fflins_create(&cfg_, device, &ctx_);
        if (rc != FFLINS_OK) throw std::runtime_error("fflins_create failed: no CUDA device (there is no CPU path)");
        if (!fflins_shim::default_ctx()) fflins_shim::default_ctx() = ctx_;
    }
    ~LidarMap() {
        *alive_ = false;   // frames that outlive the map must not touch the session any more
        if (fflins_shim::default_ctx() == ctx_) fflins_shim::default_ctx() = nullptr;
        fflins_destroy(ctx_);
    }
    LidarMap(const LidarMap &) = delete;
    LidarMap &operator=(const LidarMap &) = delete;

    // lidar_map.cc:213-273
    bool lidarFrameProcessing(const std::deque<std::pair<IMU, IntegrationState>> &ins_window, const Pose &pose_b_l,
                              LidarFrame::Ptr frame) {
        const size_t w = ins_window.size();
        ins_t_.resize(w);
        ins_p_.resize(3 * w);
        ins_q_.resize(4 * w);
        for (size_t i = 0; i < w; i++) {
            ins_t_[i] = ins_window[i].first.time;
            for (int k = 0; k < 3; k++) ins_p_[3 * i + k] = ins_window[i].second.p[k];
            ins_q_[4 * i]     = ins_window[i].second.q.x();
            ins_q_[4 * i + 1] = ins_window[i].second.q.y();
            ins_q_[4 * i + 2] = ins_window[i].second.q.z();
            ins_q_[4 * i + 3] = ins_window[i].second.q.w();
        }
        fflins_pose ext = fflins_shim::toAbi(pose_b_l);
        fflins_frame_info info;
        auto raw = frame->rawPointCloud();
        frame->bind(ctx_, cfg_.point_to_plane_std, alive_);
        // Asynchronous by default: the frame is only QUEUED (out = NULL). The raw records are pageable (a std::vector): the
        // library copies them into its page-locked ring before the call returns (so the cloud may be freed at once), the
        // kernel chain is issued with the next frames of the keyframe interval, and the counts are read back by whatever
        // needs them first (an accessor, mergeNonKeyframes, updateLidarFeaturesInMap). The frame pose, which is all LINS
        // needs at once (keyframeSelection), is evaluated on the host inside the call. (Page-locking every frame's vector
        // in place — LidarFrame::pinRaw — was measured at ~3 ms per frame for cudaHostRegister + Unregister: not used.)
        const bool async = async_frames_;
        int rc = fflins_frame_process_aos(ctx_, frame->id(), raw->points.data(), (int32_t) raw->size(), ins_t_.data(),
                                          ins_p_.data(), ins_q_.data(), (int32_t) w, &ext, frame->stamp(),
                                          frame->timeDelay(), async ? nullptr : &info);
        if (rc == FFLINS_OK && async) rc = fflins_frame_get_pose(ctx_, frame->id(), &info.pose);
        if (rc != FFLINS_OK) {
            std::fprintf(stderr, "[fflins] lidarFrameProcessing: %s\n", fflins_last_error(ctx_));
            return false;
        }
        frame->adoptPose(info.pose);
        frame->invalidate(true, false);
        return true;
    }
    // false: every lidarFrameProcessing call returns with the frame complete (pageable staged copy, one synchronisation)
    void setAsyncFrames(bool on) { async_frames_ = on; }
    // static overload, lidar_map.cc:275-324
    bool lidarFrameProcessing(const IntegrationState &state, const Pose &pose_b_l, LidarFrame::Ptr frame) {
        auto raw = frame->rawPointCloud();
        std::vector<float> xyzi(4 * raw->size());
        for (size_t i = 0; i < raw->size(); i++) {
            xyzi[4 * i]     = raw->points[i].x;
            xyzi[4 * i + 1] = raw->points[i].y;
            xyzi[4 * i + 2] = raw->points[i].z;
            xyzi[4 * i + 3] = raw->points[i].intensity;
        }
        double q[4]     = {state.q.x(), state.q.y(), state.q.z(), state.q.w()};
        fflins_pose ext = fflins_shim::toAbi(pose_b_l);
        fflins_frame_info info;
        frame->bind(ctx_, cfg_.point_to_plane_std, alive_);
        int rc = fflins_frame_process_static(ctx_, frame->id(), xyzi.data(), (int32_t) raw->size(), state.p.data(), q, &ext, &info);
        if (rc != FFLINS_OK) return false;
        frame->adoptPose(info.pose);
        frame->invalidate(true, false);
        return true;
    }

    // lidar_map.cc:428-459
    bool updateLidarFeaturesInMap() {
        fflins_assoc_stats st;
        if (fflins_associate(ctx_, &st) != FFLINS_OK) return false;
        auto cur = keyframes_.back();
        double sum = 0;
        for (int i = 0; i < st.n_refs; i++) {
            keyframes_[i]->invalidate(false, true);
            sum += st.num_features[i];
        }
        if (st.n_refs > 0) { // last-writer-wins in the reference (:404-405); the newest ref is what :447-448 reads
            cur->setNumFeatures(st.num_features[st.n_refs - 1]);
            cur->setNumCovered(st.num_covered[st.n_refs - 1]);
        }
        stat_newest_features_  = cur->numFeatures();
        stat_average_features_ = (sum + cur->numFeatures()) / (double) keyframes_.size();
        return true;
    }

    // lidar_map.cc:77-98
    bool keyframeSelection(LidarFrame::Ptr frame) {
        int is_key = 0;
        double st[3];
        fflins_shim::check(ctx_, fflins_keyframe_selection(ctx_, frame->id(), frame->stamp(), keyframes_.back()->stamp(), &is_key, st),
                           "keyframeSelection");
        stat_keyframe_interval_ = st[0];
        stat_keyframe_translation_ = st[1];
        stat_keyframe_rotation_ = st[2];
        if (is_key) frame->setKeyFrame();
        return frame->isKeyFrame();
    }
    // lidar_map.cc:190-211; the non-keyframes' device clouds are released (ff_lins.cc:245 clears them)
    void mergeNonKeyframes(const std::vector<LidarFrame::Ptr> &nonkeyframes, LidarFrame::Ptr keyframe) {
        std::vector<uint64_t> ids;
        for (auto &f : nonkeyframes) ids.push_back(f->id());
        fflins_frame_info info;
        fflins_shim::check(ctx_, fflins_keyframe_merge(ctx_, keyframe->id(), ids.data(), (int32_t) ids.size(), async_frames_ ? nullptr : &info),
                           "mergeNonKeyframes");
        keyframe->invalidate(true, false);
        // frames retired at the previous keyframe go now: their queued copies are long done, so un-pinning them costs nothing
        retired_.clear();
        for (auto &f : nonkeyframes) {
            if (keep_nonkey_clouds_) f->pointCloudMapFull(); // keep a host copy alive if a caller still wants it (viewer)
            fflins_frame_release(ctx_, f->id());
            f->bind(nullptr, cfg_.point_to_plane_std);
            retired_.push_back(f);
        }
    }
    // lidar_map.cc:100-126
    void addNewKeyFrame(LidarFrame::Ptr frame) {
        fflins_shim::check(ctx_, fflins_keyframe_add(ctx_, frame->id()), "addNewKeyFrame");
        keyframes_.push_back(frame);
    }
    // lidar_map.cc:128-173 / 175-188
    void removeOldestLidarFrame() {
        marginalized_frame_ = keyframes_.front();
        if (keep_nonkey_clouds_) marginalized_frame_->pointCloudMapFull();
        keyframes_.pop_front();
        uint64_t id;
        fflins_keyframe_remove_oldest(ctx_, &id);
        marginalized_frame_->bind(nullptr, cfg_.point_to_plane_std);
    }
    void removeNewestLidarFrame() {
        auto f = keyframes_.back();
        keyframes_.pop_back();
        uint64_t id;
        fflins_keyframe_remove_newest(ctx_, &id);
        f->bind(nullptr, cfg_.point_to_plane_std);
    }
    const std::deque<LidarFrame::Ptr> &keyframes() { return keyframes_; }
    bool isWindowFull() { return keyframes_.size() > (size_t) cfg_.window_size; }
    bool isWindowHalfFull() { return keyframes_.size() > (size_t) cfg_.window_size / 2; }

    // lidar_map.cc:484-497 (the timecost column is filled by the caller's own stopwatch)
    const std::vector<double> &statisticParameters() {
        statistic_parameters_ = {keyframes_.back()->stamp(), stat_average_features_, (double) stat_newest_features_,
                                 (double) keyframes_.back()->pointCloudLidarFull()->size(), stat_keyframe_interval_,
                                 stat_keyframe_translation_, stat_keyframe_rotation_ * 180.0 / M_PI, 0.0};
        return statistic_parameters_;
    }

    // ---- additions that make the batched device path reachable from the reference's call sites ----
    fflins_ctx *context() { return ctx_; }
    // Optimizer::updateParametersFromOptimizer (optimizer.cc:203-218): one launch for the whole window
    void setKeepReleasedClouds(bool on) { keep_nonkey_clouds_ = on; }
    void reprojectWorld(const std::vector<Pose> &poses) {
        std::vector<uint64_t> ids;
        std::vector<fflins_pose> ps;
        for (size_t i = 0; i < keyframes_.size() && i < poses.size(); i++) {
            ids.push_back(keyframes_[i]->id());
            ps.push_back(fflins_shim::toAbi(poses[i]));
            keyframes_[i]->adoptPose(ps.back());
            keyframes_[i]->invalidate(true, false);
        }
        fflins_shim::check(ctx_, fflins_reproject_window(ctx_, (int32_t) ids.size(), ids.data(), ps.data()), "reprojectWorld");
    }
    // host-side LidarFeature::setOutlier flags of one ref frame -> device (optimizer.cc:537-541)
    void pushOutliers(LidarFrame::Ptr ref) {
        auto &fs = ref->features();
        std::vector<uint8_t> mask(fs.size());
        for (size_t k = 0; k < fs.size(); k++) mask[k] = fs[k]->isOutlier();
        fflins_features_set_outlier(ctx_, ref->id(), mask.data(), (int32_t) mask.size());
    }

private:
    fflins_config cfg_;
    fflins_ctx *ctx_ = nullptr;
    std::shared_ptr<bool> alive_ = std::make_shared<bool>(true);
    std::deque<LidarFrame::Ptr> keyframes_;
    std::vector<LidarFrame::Ptr> retired_;
    LidarFrame::Ptr marginalized_frame_;
    std::vector<double> ins_t_, ins_p_, ins_q_;
    std::vector<double> statistic_parameters_;
    bool async_frames_ = true;
    bool keep_nonkey_clouds_ = false;   // the reference keeps them for the viewer / PCD dump only (lidar_map.cc:140-170)
    double stat_average_features_ = 0, stat_keyframe_interval_ = 0, stat_keyframe_translation_ = 0, stat_keyframe_rotation_ = 0;
    int stat_newest_features_ = 0;
};

} // namespace lidar

// ---- LidarFactor (ff_lins/factors/lidar_factor.h:31-51) ---------------------------------------------
#if defined(FFLINS_SHIM_USE_CERES)
#include <ceres/ceres.h>
typedef ceres::SizedCostFunction<1, 7, 7, 7, 1> LidarFactorBase;
#else
namespace fflins_shim {
// the slice of ceres::SizedCostFunction<1,7,7,7,1> the factor uses
class LidarFactorBase {
public:
    virtual ~LidarFactorBase() {}
    virtual bool Evaluate(const double *const *parameters, double *residuals, double **jacobians) const = 0;
    int num_residuals() const { return 1; }
    std::vector<int> parameter_block_sizes() const { return {7, 7, 7, 1}; }
};
} // namespace fflins_shim
typedef fflins_shim::LidarFactorBase LidarFactorBase;
#endif

// Same constructor and Evaluate signature as the reference. One residual per call costs a device
// round trip; it exists so existing call sites (optimizer.cc:472-479,533) keep compiling. The
// throughput path is LidarWindowEvaluator below (one launch per Ceres iteration for ALL residuals).
class LidarFactor : public LidarFactorBase {
public:
    // the reference's exact constructor (lidar_factor.h:34-35): default session
    LidarFactor(const PointType &point_lidar, Vector3d unit_normal_vector, double norm_inverse, double std, Vector3d vel0,
                Vector3d omega0, double td0, Vector3d vel1, Vector3d omega1, double td1)
        : LidarFactor(fflins_shim::default_ctx(), point_lidar, unit_normal_vector, norm_inverse, std, vel0, omega0, td0, vel1,
                      omega1, td1) {}
    LidarFactor(fflins_ctx *ctx, const PointType &point_lidar, Vector3d unit_normal_vector, double norm_inverse, double std,
                Vector3d vel0, Vector3d omega0, double td0, Vector3d vel1, Vector3d omega1, double td1)
        : ctx_(ctx), norm_inverse_(norm_inverse), std_(std) {
        point_[0] = point_lidar.x;
        point_[1] = point_lidar.y;
        point_[2] = point_lidar.z;
        for (int i = 0; i < 3; i++) {
            normal_[i]      = unit_normal_vector[i];
            motion_[i]      = vel0[i];
            motion_[3 + i]  = omega0[i];
            motion_[7 + i]  = vel1[i];
            motion_[10 + i] = omega1[i];
        }
        motion_[6]  = td0;
        motion_[13] = td1;
    }
    // parameters = {pose_ref[7], pose_cur[7], extrinsic[7], td[1]}; jacobians / jacobians[i] may be null
    bool Evaluate(const double *const *parameters, double *residuals, double **jacobians) const override {
        double J[22];
        std::lock_guard<std::mutex> lk(fflins_shim::ctx_mutex());
        int rc = fflins_factor_eval_batch(ctx_, 1, point_, normal_, &norm_inverse_, &std_, motion_, parameters[0], parameters[1],
                                          parameters[2], parameters[3][0], 0, residuals, jacobians ? J : nullptr);
        if (rc != FFLINS_OK) return false;
        if (jacobians) {
            const int off[4] = {0, 7, 14, 21}, len[4] = {7, 7, 7, 1};
            for (int b = 0; b < 4; b++)
                if (jacobians[b]) std::memcpy(jacobians[b], J + off[b], len[b] * sizeof(double));
        }
        return true;
    }

private:
    fflins_ctx *ctx_;
    float point_[3];
    double normal_[3];
    double norm_inverse_, std_;
    double motion_[14];
};

// All residuals of all (ref, newest keyframe) pairs in one launch: robustified H = sum J^T J (19x19),
// b = -sum J^T r, cost per pair — what MarginalizationInfo::constructEquation accumulates
// (marginalization_info.cc:133-210) and what a square-root-compressed Ceres factor consumes
// (INTEGRATION.md §3).
class LidarWindowEvaluator {
public:
    explicit LidarWindowEvaluator(fflins_ctx *ctx) : ctx_(ctx) {}
    struct Blocks {
        std::vector<double> H, b, cost;
        std::vector<int32_t> n_res;
    };
    bool evaluate(const std::vector<uint64_t> &ref_ids, const std::vector<double> &pose_refs, const double pose_cur[7],
                  const double ext[7], double td, uint32_t flags, Blocks &out) {
        size_t n = ref_ids.size();
        out.H.assign(361 * n, 0);
        out.b.assign(19 * n, 0);
        out.cost.assign(2 * n, 0);
        out.n_res.assign(n, 0);
        std::lock_guard<std::mutex> lk(fflins_shim::ctx_mutex());
        return fflins_window_eval(ctx_, (int32_t) n, ref_ids.data(), pose_refs.data(), pose_cur, ext, td, flags, out.H.data(),
                                  out.b.data(), out.cost.data(), out.n_res.data()) == FFLINS_OK;
    }
    bool chi2(const std::vector<uint64_t> &ref_ids, const std::vector<double> &pose_refs, const double pose_cur[7],
              const double ext[7], double td, double threshold, std::vector<int32_t> &n_outliers) {
        n_outliers.assign(ref_ids.size(), 0);
        std::lock_guard<std::mutex> lk(fflins_shim::ctx_mutex());
        return fflins_window_chi2(ctx_, (int32_t) ref_ids.size(), ref_ids.data(), pose_refs.data(), pose_cur, ext, td, threshold,
                                  n_outliers.data()) == FFLINS_OK;
    }

private:
    fflins_ctx *ctx_;
};

// ---- Ceres-shaped adapters for one (ref, cur) pair --------------------------------------------------
// Both keep LidarFactor's parameter-block signature {pose_ref[7], pose_cur[7], extrinsic[7], td[1]} and are added with
// loss = nullptr: HuberLoss(1.0) is applied per residual inside the kernel, exactly as ResidualBlockInfo::Evaluate does
// (residual_block_info.cc:49-77) — Ceres itself would apply a loss per residual BLOCK, which is not what the reference's
// one-block-per-point construction means.
#if defined(FFLINS_SHIM_USE_CERES)
typedef ceres::CostFunction LidarPairFactorBase;
#else
namespace fflins_shim {
class LidarPairFactorBase {   // the slice of ceres::CostFunction a dynamically sized factor uses
public:
    virtual ~LidarPairFactorBase() {}
    virtual bool Evaluate(const double *const *parameters, double *residuals, double **jacobians) const = 0;
    int num_residuals() const { return num_residuals_; }
    const std::vector<int32_t> &parameter_block_sizes() const { return sizes_; }

protected:
    void set_num_residuals(int n) { num_residuals_ = n; }
    std::vector<int32_t> *mutable_parameter_block_sizes() { return &sizes_; }

private:
    int num_residuals_ = 0;
    std::vector<int32_t> sizes_;
};
} // namespace fflins_shim
typedef fflins_shim::LidarPairFactorBase LidarPairFactorBase;
#endif

// (i) batched factor: num_residuals = P (the pair's valid residuals, fixed at construction: PLANE and not outlier,
// optimizer.cc:468), one kernel launch per Evaluate for all of them.
class LidarBatchedFactor : public LidarPairFactorBase {
public:
    LidarBatchedFactor(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, uint32_t flags = FFLINS_HUBER)
        : ctx_(ctx), ref_(ref_id), cur_(cur_id), flags_(flags) {
        int32_t q = 0;
        fflins_shim::check(ctx_, fflins_features_download(ctx_, ref_, 0, &q, 0, 0, 0, 0, 0, 0, 0, 0), "features size");
        std::vector<int8_t> type(q);
        std::vector<uint8_t> outlier(q);
        if (q > 0)
            fflins_shim::check(ctx_, fflins_features_download(ctx_, ref_, q, &q, type.data(), nullptr, outlier.data(), nullptr, nullptr,
                                                              nullptr, nullptr, nullptr), "features");
        for (int k = 0; k < q; k++)
            if (type[k] == FFLINS_FEATURE_PLANE && !outlier[k]) rows_.push_back(k);
        q_ = q;
        set_num_residuals((int) rows_.size());
        *mutable_parameter_block_sizes() = {7, 7, 7, 1};
    }
    bool Evaluate(const double *const *parameters, double *residuals, double **jacobians) const override {
        std::vector<double> r(q_), J(jacobians ? (size_t) 22 * q_ : 0);
        std::vector<uint8_t> valid(q_);
        int32_t n_res = 0;
        {
            std::lock_guard<std::mutex> lk(fflins_shim::ctx_mutex());
            if (fflins_factor_eval(ctx_, ref_, cur_, parameters[0], parameters[1], parameters[2], parameters[3][0], flags_, q_, &n_res,
                                   valid.data(), r.data(), jacobians ? J.data() : nullptr, nullptr, nullptr, nullptr) != FFLINS_OK)
                return false;
        }
        const int P = (int) rows_.size();
        const int off[4] = {0, 7, 14, 21}, len[4] = {7, 7, 7, 1};
        for (int i = 0; i < P; i++) {
            const int k  = rows_[i];
            residuals[i] = valid[k] ? r[k] : 0.0;   // a residual culled after construction contributes nothing
            if (!jacobians) continue;
            for (int b = 0; b < 4; b++)
                if (jacobians[b])
                    for (int c = 0; c < len[b]; c++) jacobians[b][(size_t) i * len[b] + c] = valid[k] ? J[(size_t) 22 * k + off[b] + c] : 0.0;
        }
        return true;
    }

private:
    fflins_ctx *ctx_;
    uint64_t ref_, cur_;
    uint32_t flags_;
    int32_t q_ = 0;
    std::vector<int> rows_;
};

// (ii) square-root-compressed factor: the kernel returns the pair's robustified H = sum J^T J (19 x 19), b = -sum J^T r and
// c = sum r^2; with H = V S V^T the factor hands Ceres J' = S^{1/2} V^T (19 rows) and r' = -S^{-1/2} V^T b, plus one
// Jacobian-free residual sqrt(c - |r'|^2): cost, gradient and Gauss-Newton Hessian are those of the P original residuals
// (the algebra of MarginalizationInfo::linearization, marginalization_info.cc:93-107). 20 residuals instead of P.
class LidarSqrtFactor : public LidarPairFactorBase {
public:
    LidarSqrtFactor(fflins_ctx *ctx, uint64_t ref_id, uint64_t cur_id, uint32_t flags = FFLINS_HUBER)
        : ctx_(ctx), ref_(ref_id), cur_(cur_id), flags_(flags) {
        set_num_residuals(20);
        *mutable_parameter_block_sizes() = {7, 7, 7, 1};
    }
    bool Evaluate(const double *const *parameters, double *residuals, double **jacobians) const override {
        double H[361], b[19], cost[2];
        int32_t n_res = 0;
        {
            std::lock_guard<std::mutex> lk(fflins_shim::ctx_mutex());
            if (fflins_factor_eval(ctx_, ref_, cur_, parameters[0], parameters[1], parameters[2], parameters[3][0], flags_, 0, &n_res,
                                   nullptr, nullptr, nullptr, H, b, cost) != FFLINS_OK)
                return false;
        }
        double w[19], V[361];
        symEig19(H, w, V);
        double wmax = 0;
        for (int k = 0; k < 19; k++) wmax = std::fmax(wmax, w[k]);
        double rr = 0;
        for (int k = 0; k < 19; k++) {
            // keep every direction whose Jacobian is above ~1e-10 of the largest (a weakly observed block such as td must keep
            // its gradient); true null directions (ref t = -cur t) sit at ~1e-16 of the largest eigenvalue
            const bool live = w[k] > 1e-20 * wmax && w[k] > 0;
            const double sq = live ? std::sqrt(w[k]) : 0.0;
            double vb = 0;
            for (int j = 0; j < 19; j++) vb += V[19 * j + k] * b[j];
            residuals[k] = live ? -vb / sq : 0.0;
            rr += residuals[k] * residuals[k];
            if (jacobians) {   // tangent column j of the 19-vector -> block / column of the 7,7,7,1 layout (column 6 = 0)
                for (int j = 0; j < 19; j++) {
                    const int blk = j < 18 ? j / 6 : 3, col = j < 18 ? j % 6 : 0, len = blk < 3 ? 7 : 1;
                    if (jacobians[blk]) jacobians[blk][(size_t) k * len + col] = sq * V[19 * j + k];
                }
                for (int blk = 0; blk < 3; blk++)
                    if (jacobians[blk]) jacobians[blk][(size_t) k * 7 + 6] = 0.0;
            }
        }
        residuals[19] = std::sqrt(std::fmax(cost[1] - rr, 0.0));
        if (jacobians)
            for (int blk = 0; blk < 4; blk++)
                if (jacobians[blk])
                    for (int c = 0; c < (blk < 3 ? 7 : 1); c++) jacobians[blk][(size_t) 19 * (blk < 3 ? 7 : 1) + c] = 0.0;
        return true;
    }

private:
    // cyclic Jacobi: H = V diag(w) V^T (row-major 19 x 19)
    static void symEig19(const double *Hin, double *w, double *V) {
        const int n = 19;
        double A[361];
        std::memcpy(A, Hin, sizeof(A));
        for (int i = 0; i < n * n; i++) V[i] = (i % (n + 1) == 0) ? 1.0 : 0.0;
        for (int sweep = 0; sweep < 60; sweep++) {
            double off = 0, diag = 0;
            for (int i = 0; i < n; i++) {
                diag += A[n * i + i] * A[n * i + i];
                for (int j = i + 1; j < n; j++) off += A[n * i + j] * A[n * i + j];
            }
            if (off <= 1e-32 * (diag + 1e-300)) break;
            for (int p = 0; p < n; p++)
                for (int q = p + 1; q < n; q++) {
                    const double apq = A[n * p + q];
                    if (std::fabs(apq) < 1e-300) continue;
                    const double theta = (A[n * q + q] - A[n * p + p]) / (2 * apq);
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                    const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                    for (int k = 0; k < n; k++) {
                        const double akp = A[n * k + p], akq = A[n * k + q];
                        A[n * k + p] = c * akp - s * akq;
                        A[n * k + q] = s * akp + c * akq;
                    }
                    for (int k = 0; k < n; k++) {
                        const double apk = A[n * p + k], aqk = A[n * q + k];
                        A[n * p + k] = c * apk - s * aqk;
                        A[n * q + k] = s * apk + c * aqk;
                    }
                    for (int k = 0; k < n; k++) {
                        const double vkp = V[n * k + p], vkq = V[n * k + q];
                        V[n * k + p] = c * vkp - s * vkq;
                        V[n * k + q] = s * vkp + c * vkq;
                    }
                }
        }
        for (int i = 0; i < n; i++) w[i] = A[n * i + i];
    }
    fflins_ctx *ctx_;
    uint64_t ref_, cur_;
    uint32_t flags_;
};
