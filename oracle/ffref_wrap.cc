/* C wrapper around the REFERENCE's own hot-path sources, compiled where they lie under /root/reference by
 * oracle/Makefile `ffref` into oracle/_ref/libff_ref.so: factors/lidar_factor.cc (LidarFactor::Evaluate),
 * common/misc.cc (getInsWindowIndex, getPoseFromInsWindow, statePoseInterpolation, stateToPoseWithExtrinsic,
 * insMechanization), common/rotation.cc and lidar/pointcloud.cc (relativeProjection, pointProjection,
 * pointCloudProjection, planeEstimation). Eigen / PCL / Ceres / TBB / glog are not installed here: the sources are
 * compiled against the stand-in headers of oracle/refshim/ (see mini_eigen.h for what that does and does not pin).
 * The only logic restated in this file is the per-point loop of LidarMap::lidarFrameProcessing (lidar_map.cc:222-243),
 * whose translation unit needs the whole map / kd-tree / VoxelGrid stack. Argument layouts are those of oracle.h so
 * that tests can call both sides with the same arrays. Test infrastructure only. */
#include "common/misc.h"
#include "common/rotation.h"
#include "factors/lidar_factor.h"
#include "factors/marginalization_factor.h"
#include "factors/marginalization_info.h"
#include "factors/pose_manifold.h"
#include "factors/residual_block_info.h"
#include "preintegration/preintegration_factor.h"
#include "preintegration/preintegration_normal.h"
#include "lidar/pointcloud.h"

#include <cstring>

#include "lidar_converter.h"

namespace {
typedef std::deque<std::pair<IMU, IntegrationState>> InsWindow;

InsWindow make_window(const double *ins_t, const double *ins_p, const double *ins_q, int w) {
    InsWindow win;
    for (int k = 0; k < w; k++) {
        IMU imu;
        imu.time = ins_t[k];
        imu.dt   = 0;
        IntegrationState st;
        st.time = ins_t[k];
        st.p    = Vector3d(ins_p[3 * k], ins_p[3 * k + 1], ins_p[3 * k + 2]);
        st.q    = Quaterniond(ins_q[4 * k + 3], ins_q[4 * k], ins_q[4 * k + 1], ins_q[4 * k + 2]); /* arrays are xyzw */
        win.emplace_back(imu, st);
    }
    return win;
}
Pose make_pose(const double *R, const double *t) {
    Pose p;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) p.R(r, c) = R[3 * r + c];
    p.t = Vector3d(t[0], t[1], t[2]);
    return p;
}
void store_pose(const Pose &p, double *R, double *t) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[3 * r + c] = p.R(r, c);
    for (int i = 0; i < 3; i++) t[i] = p.t[i];
}
PointType make_point(const float *xyzi) {
    PointType p;
    p.x = xyzi[0], p.y = xyzi[1], p.z = xyzi[2], p.intensity = xyzi[3];
    return p;
}
} // namespace

extern "C" {

int ffref_ins_window_index(const double *ins_t, int w, double time) {
    static const double zero3[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1};
    InsWindow win;
    for (int k = 0; k < w; k++) {
        InsWindow one = make_window(ins_t + k, zero3, q, 1);
        win.push_back(one.front());
    }
    return (int) MISC::getInsWindowIndex(win, time);
}

int ffref_pose_from_ins_window(const double *ins_t, const double *ins_p, const double *ins_q, int w, const double *R_bl,
                               const double *t_bl, double time, double *R, double *t) {
    InsWindow win = make_window(ins_t, ins_p, ins_q, w);
    Pose pose;
    bool ok = MISC::getPoseFromInsWindow(win, make_pose(R_bl, t_bl), time, pose);
    store_pose(pose, R, t);
    return ok ? 1 : 0;
}

void ffref_state_to_pose(const double *p, const double *q, const double *R_bl, const double *t_bl, double *R, double *t) {
    IntegrationState st;
    st.p = Vector3d(p[0], p[1], p[2]);
    st.q = Quaterniond(q[3], q[0], q[1], q[2]);
    store_pose(MISC::stateToPoseWithExtrinsic(st, make_pose(R_bl, t_bl)), R, t);
}

/* the per-point loop of LidarMap::lidarFrameProcessing (lidar_map.cc:222-243) over the reference's own
 * MISC::getPoseFromInsWindow and PointCloudCommon::relativeProjection; returns the number of fallbacks */
int ffref_undistort(const float *xyzi, const double *time, int n, const double *ins_t, const double *ins_p,
                    const double *ins_q, int w, const double *R_bl, const double *t_bl, double stamp, double td,
                    float *out_xyzi, double *pose_R, double *pose_t) {
    InsWindow win = make_window(ins_t, ins_p, ins_q, w);
    Pose pose_b_l = make_pose(R_bl, t_bl), lidar_pose;
    MISC::getPoseFromInsWindow(win, pose_b_l, stamp, lidar_pose);
    store_pose(lidar_pose, pose_R, pose_t);
    int fallbacks = 0;
    for (int k = 0; k < n; k++) {
        PointType point = make_point(xyzi + 4 * k);
        Pose curr_pose;
        if (!MISC::getPoseFromInsWindow(win, pose_b_l, time[k] + td, curr_pose)) fallbacks++;
        PointType out       = lidar::PointCloudCommon::relativeProjection(lidar_pose, curr_pose, point);
        out_xyzi[4 * k]     = out.x;
        out_xyzi[4 * k + 1] = out.y;
        out_xyzi[4 * k + 2] = out.z;
        out_xyzi[4 * k + 3] = out.intensity;
    }
    return fallbacks;
}

void ffref_project(const double *R, const double *t, const float *src, int n, float *dst) {
    PointCloudPtr a(new PointCloud), b(new PointCloud);
    a->resize(n);
    for (int k = 0; k < n; k++) a->points[k] = make_point(src + 4 * k);
    lidar::PointCloudCommon::pointCloudProjection(make_pose(R, t), a, b);
    for (int k = 0; k < n; k++) {
        dst[4 * k]     = b->points[k].x;
        dst[4 * k + 1] = b->points[k].y;
        dst[4 * k + 2] = b->points[k].z;
        dst[4 * k + 3] = b->points[k].intensity;
    }
}

int ffref_plane_estimation(const float *pts, double threshold, double *unit_normal, double *norm_inverse, double *distance) {
    PointVector points;
    for (int k = 0; k < 5; k++) points.push_back(make_point(pts + 4 * k));
    Vector3d n;
    std::vector<double> dist;
    bool ok = lidar::PointCloudCommon::planeEstimation(points, threshold, n, *norm_inverse, dist);
    for (int i = 0; i < 3; i++) unit_normal[i] = n[i];
    for (size_t i = 0; i < 5; i++) distance[i] = i < dist.size() ? dist[i] : 0.0;
    return ok ? 1 : 0;
}

/* motion = {v0, w0, td0, v1, w1, td1}; parameters = {pose_ref, pose_cur, ext, td} (7, 7, 7, 1) */
int ffref_lidar_factor_evaluate(const float *point_xyz, const double *n, double d, double std, const double *motion,
                                const double *const *parameters, double *residuals, double **jacobians) {
    PointType p;
    p.x = point_xyz[0], p.y = point_xyz[1], p.z = point_xyz[2];
    LidarFactor f(p, Vector3d(n[0], n[1], n[2]), d, std, Vector3d(motion[0], motion[1], motion[2]),
                  Vector3d(motion[3], motion[4], motion[5]), motion[6], Vector3d(motion[7], motion[8], motion[9]),
                  Vector3d(motion[10], motion[11], motion[12]), motion[13]);
    return f.Evaluate(parameters, residuals, jacobians) ? 1 : 0;
}

/* ResidualBlockInfo::Evaluate (residual_block_info.cc:33-80) over a LidarFactor with ceres::HuberLoss(delta) (delta <= 0:
 * no loss function): corrected residual and the four corrected Jacobian blocks [7, 7, 7, 1] -> J22 */
void ffref_residual_block_evaluate(const float *point_xyz, const double *n, double d, double std, const double *motion,
                                   const double *const *parameters, double huber_delta, double *residual, double *J22) {
    PointType p;
    p.x = point_xyz[0], p.y = point_xyz[1], p.z = point_xyz[2];
    auto factor = std::make_shared<LidarFactor>(p, Vector3d(n[0], n[1], n[2]), d, std, Vector3d(motion[0], motion[1], motion[2]),
                                                Vector3d(motion[3], motion[4], motion[5]), motion[6],
                                                Vector3d(motion[7], motion[8], motion[9]),
                                                Vector3d(motion[10], motion[11], motion[12]), motion[13]);
    std::shared_ptr<ceres::LossFunction> loss;
    if (huber_delta > 0) loss = std::make_shared<ceres::HuberLoss>(huber_delta);
    std::vector<double *> blocks;
    for (int i = 0; i < 4; i++) blocks.push_back(const_cast<double *>(parameters[i]));
    ResidualBlockInfo info(factor, loss, blocks, std::vector<int>{});
    info.Evaluate();
    residual[0] = info.residuals()[0];
    int o       = 0;
    for (int b = 0; b < 4; b++) {
        const auto &J = info.jacobians()[b];
        for (int j = 0; j < J.cols(); j++) J22[o++] = J(0, j);
    }
}

/* MarginalizationInfo (factors/marginalization_info.cc:26-260) over the LiDAR residual blocks of ONE (oldest reference,
 * newest) pair, as Optimizer::marginalization (optimizer.cc:221-404) builds them for the LiDAR share: one
 * ResidualBlockInfo(LidarFactor, HuberLoss(delta), {pose_ref, pose_cur, ext, td}, marginalize {0}) per valid feature;
 * marginalization() = evaluate, constructEquation (J^T J / -J^T e over local sizes), schurElimination, linearization.
 * Outputs the prior on the remaining blocks in the fixed order [pose_cur 6 | ext 6 | td 1] (the reference's own order
 * follows an unordered_map): J0 (13 x 13, row-major, columns permuted to that order) and e0 (13). Returns 0 when the
 * reference reports the marginalization as invalid. */
int ffref_marginalize_lidar_pair(int n, const float *point_xyz, const double *normal, const double *d, double std,
                                 const double *motion, const double *pose_ref, const double *pose_cur, const double *ext,
                                 double td, double huber_delta, double *J0, double *e0, const double *eval_cur,
                                 const double *eval_ext, double eval_td, double *eval_residuals, double *eval_jacobian) {
    double blocks[4][7];
    std::memcpy(blocks[0], pose_ref, 7 * sizeof(double));
    std::memcpy(blocks[1], pose_cur, 7 * sizeof(double));
    std::memcpy(blocks[2], ext, 7 * sizeof(double));
    blocks[3][0] = td;
    std::unordered_map<long, long> ids;
    std::unordered_map<long, double *> address;
    for (long i = 0; i < 4; i++) {
        ids[reinterpret_cast<long>(blocks[i])] = i;
        address[i]                            = blocks[i];
    }
    auto info = std::make_shared<MarginalizationInfo>();
    info->updateParamtersIds(ids);
    std::shared_ptr<ceres::LossFunction> loss;
    if (huber_delta > 0) loss = std::make_shared<ceres::HuberLoss>(huber_delta);
    for (int k = 0; k < n; k++) {
        PointType p;
        p.x = point_xyz[3 * k], p.y = point_xyz[3 * k + 1], p.z = point_xyz[3 * k + 2];
        auto factor = std::make_shared<LidarFactor>(
            p, Vector3d(normal[3 * k], normal[3 * k + 1], normal[3 * k + 2]), d[k], std, Vector3d(motion[0], motion[1], motion[2]),
            Vector3d(motion[3], motion[4], motion[5]), motion[6], Vector3d(motion[7], motion[8], motion[9]),
            Vector3d(motion[10], motion[11], motion[12]), motion[13]);
        info->addResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
            factor, loss, std::vector<double *>{blocks[0], blocks[1], blocks[2], blocks[3]}, std::vector<int>{0}));
    }
    if (!info->marginalization()) return 0;
    std::vector<double *> kept = info->getParamterBlocks(address);
    const int m = info->marginalizedSize(), rdim = info->remainedSize();
    if (m != 6 || rdim != 13 || kept.size() != 3) return 0;
    int col_of[13]; /* reference column -> canonical column */
    for (size_t i = 0; i < kept.size(); i++) {
        const int at    = info->remainedBlockIndex()[i] - m;
        const int local = MarginalizationInfo::localSize(info->remainedBlockSize()[i]);
        const int base  = kept[i] == blocks[1] ? 0 : kept[i] == blocks[2] ? 6 : 12;
        for (int j = 0; j < local; j++) col_of[at + j] = base + j;
    }
    const auto &J = info->linearizedJacobians();
    const auto &e = info->linearizedResiduals();
    for (int i = 0; i < 13; i++) {
        e0[i] = e[i];
        for (int j = 0; j < 13; j++) J0[13 * i + col_of[j]] = J(i, j);
    }
    if (eval_cur) {
        /* MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91) of that prior at displaced parameter blocks:
         * residuals e0 + J0 dx (13) and the Jacobian with respect to the TANGENT of [pose_cur | ext | td] (13 x 13,
         * row-major: the leftCols(local) of every block, canonical column order) */
        MarginalizationFactor prior(info);
        double cur7[7], ext7[7], tdv[1] = {eval_td};
        std::memcpy(cur7, eval_cur, sizeof(cur7));
        std::memcpy(ext7, eval_ext, sizeof(ext7));
        std::vector<const double *> params;
        std::vector<std::vector<double>> jac(kept.size());
        std::vector<double *> jp;
        for (size_t i = 0; i < kept.size(); i++) {
            params.push_back(kept[i] == blocks[1] ? cur7 : kept[i] == blocks[2] ? ext7 : tdv);
            jac[i].assign((size_t) 13 * info->remainedBlockSize()[i], 0.0);
            jp.push_back(jac[i].data());
        }
        prior.Evaluate(params.data(), eval_residuals, jp.data());
        for (size_t i = 0; i < kept.size(); i++) {
            const int size  = info->remainedBlockSize()[i];
            const int local = MarginalizationInfo::localSize(size);
            const int base  = kept[i] == blocks[1] ? 0 : kept[i] == blocks[2] ? 6 : 12;
            for (int r = 0; r < 13; r++)
                for (int j = 0; j < local; j++) eval_jacobian[13 * r + base + j] = jac[i][(size_t) r * size + j];
        }
    }
    return 1;
}

/* ---- ROS/lidar/lidar_converter.cc (LidarConverter): Livox CustomMsg and Velodyne / Ouster / Hesai PointCloud2 to the
 * reference's PointXYZIT records. Messages are the stand-ins of oracle/refshim (decoded arrays instead of a byte blob).
 * out32: 32-byte records {x, y, z, pad, intensity, pad, time}. kind: 2 Velodyne, 3 Ouster, 4 Hesai (lidar/lidar.h LidarType). */
namespace {
int store_cloud(const PointCloudCustomPtr &cloud, void *out32, int cap) {
    int n = 0;
    for (const auto &p : cloud->points) {
        if (n >= cap) break;
        char *o = static_cast<char *>(out32) + 32 * (size_t) n++;
        std::memset(o, 0, 32);
        std::memcpy(o, &p.x, 4), std::memcpy(o + 4, &p.y, 4), std::memcpy(o + 8, &p.z, 4);
        std::memcpy(o + 16, &p.intensity, 4);
        std::memcpy(o + 24, &p.time, 8);
    }
    return (int) cloud->size();
}
} // namespace
int ffref_convert_livox(const void *pts19, int n, uint32_t sec, uint32_t nsec, int scan_line, double min_range, double max_range,
                        int to_gps_time, void *out32, int cap, double *start, double *end) {
    auto msg = std::make_shared<livox_ros_driver::CustomMsg>();
    msg->header.stamp.sec = sec, msg->header.stamp.nsec = nsec;
    msg->points.resize(n);
    for (int k = 0; k < n; k++) {   /* packed rows: uint32 offset, 3 floats, 3 bytes */
        const char *r = static_cast<const char *>(pts19) + 19 * (size_t) k;
        auto &p       = msg->points[k];
        std::memcpy(&p.offset_time, r, 4), std::memcpy(&p.x, r + 4, 4), std::memcpy(&p.y, r + 8, 4), std::memcpy(&p.z, r + 12, 4);
        p.reflectivity = (uint8_t) r[16], p.tag = (uint8_t) r[17], p.line = (uint8_t) r[18];
    }
    LidarConverter conv(scan_line, min_range, max_range);
    PointCloudCustomPtr cloud(new PointCloudCustom);
    conv.livoxPointCloudConvertion(msg, cloud, *start, *end, to_gps_time != 0);
    return store_cloud(cloud, out32, cap);
}
int ffref_convert_generic(int kind, const float *xyzi, const double *time, int n, uint32_t sec, uint32_t nsec, double min_range,
                          double max_range, int to_gps_time, void *out32, int cap, double *start, double *end) {
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->header.stamp.sec = sec, msg->header.stamp.nsec = nsec;
    for (int k = 0; k < n; k++) {
        msg->x.push_back(xyzi[4 * k]), msg->y.push_back(xyzi[4 * k + 1]), msg->z.push_back(xyzi[4 * k + 2]);
        msg->intensity.push_back(xyzi[4 * k + 3]);
        msg->time.push_back(time[k]);
    }
    LidarConverter conv(128, min_range, max_range);
    PointCloudCustomPtr cloud(new PointCloudCustom);
    sensor_msgs::PointCloud2ConstPtr cmsg = msg;
    if (kind == 2) conv.velodynePointCloudConvertion(cmsg, cloud, *start, *end, to_gps_time != 0);
    else if (kind == 3) conv.ousterPointCloudConvertion(cmsg, cloud, *start, *end, to_gps_time != 0);
    else conv.hesaiPointCloudConvertion(cmsg, cloud, *start, *end, to_gps_time != 0);
    return store_cloud(cloud, out32, cap);
}

/* ---- the INS window as LINS keeps it (ff_lins.cc: insMechanization per IMU epoch, push_back) with the reference's own
 * MISC::getImuSeriesFromTo (misc.cc:282-336) and MISC::redoInsMechanization (misc.cc:186-236), iswithearth = false.
 * imu rows = {time, dt, dtheta[3], dvel[3]}; states = {p, q xyzw, v, bg, ba}. Outputs: the IMU series between two times
 * (interpolated at both ends) and, after re-mechanizing from `updated` with `reserved` epochs kept, the window's epochs. ---- */
int ffref_ins_window_scenario(const double *gravity, const double *imu, int n, const double *state16, double time0,
                              double series_start, double series_end, double *series_out, int series_cap, int *series_ok,
                              const double *updated16, double updated_time, int reserved, double *win_time, double *win_state16,
                              int win_cap) {
    IntegrationConfiguration cfg{};
    cfg.iswithearth = false;
    cfg.gravity     = Vector3d(gravity[0], gravity[1], gravity[2]);
    auto ld_state   = [](const double *s, double t) {
        IntegrationState st;
        st.time = t;
        st.p    = Vector3d(s[0], s[1], s[2]);
        st.q    = Quaterniond(s[6], s[3], s[4], s[5]);
        st.v    = Vector3d(s[7], s[8], s[9]);
        st.bg   = Vector3d(s[10], s[11], s[12]);
        st.ba   = Vector3d(s[13], s[14], s[15]);
        return st;
    };
    auto ld = [](const double *a) {
        IMU m;
        m.time   = a[0];
        m.dt     = a[1];
        m.dtheta = Vector3d(a[2], a[3], a[4]);
        m.dvel   = Vector3d(a[5], a[6], a[7]);
        m.odovel = 0;
        return m;
    };
    std::deque<std::pair<IMU, IntegrationState>> window;
    IntegrationState st = ld_state(state16, time0);
    window.emplace_back(ld(imu), st);
    for (int k = 1; k < n; k++) {
        MISC::insMechanization(cfg, ld(imu + 8 * (k - 1)), ld(imu + 8 * k), st);
        window.emplace_back(ld(imu + 8 * k), st);
    }
    std::vector<IMU> series;
    *series_ok = MISC::getImuSeriesFromTo(window, series_start, series_end, series) ? 1 : 0;
    int ns     = 0;
    for (const IMU &m : series) {
        if (ns >= series_cap) break;
        double *o = series_out + 8 * ns++;
        o[0] = m.time, o[1] = m.dt;
        for (int i = 0; i < 3; i++) o[2 + i] = m.dtheta[i], o[5 + i] = m.dvel[i];
    }
    MISC::redoInsMechanization(cfg, ld_state(updated16, updated_time), (size_t) reserved, window);
    int nw = 0;
    for (const auto &e : window) {
        if (nw >= win_cap) break;
        win_time[nw] = e.second.time;
        double *o    = win_state16 + 16 * nw++;
        for (int i = 0; i < 3; i++) o[i] = e.second.p[i], o[7 + i] = e.second.v[i], o[10 + i] = e.second.bg[i], o[13 + i] = e.second.ba[i];
        o[3] = e.second.q.x(), o[4] = e.second.q.y(), o[5] = e.second.q.z(), o[6] = e.second.q.w();
    }
    (void) ns;
    return (int) series.size() * 100000 + (int) window.size();   /* series count, window size */
}

/* ---- IMU preintegration: PreintegrationNormal (preintegration/preintegration_base.cc, preintegration_normal.cc) driven
 * through PreintegrationFactor::Evaluate (preintegration_factor.h:45-70), the cost function the optimizer adds per time node.
 * imu = {time, dt, dtheta[3], dvel[3]}; state16 = {p, q xyzw, v, bg, ba}; noise = {acc_vrw, gyr_arw, gyr_bias_std,
 * acc_bias_std, corr_time, gravity}. ---- */
namespace {
struct PreintProbe : PreintegrationNormal {
    using PreintegrationNormal::PreintegrationNormal;
    const Eigen::MatrixXd &cov() const { return covariance_; }
    const Eigen::MatrixXd &jac() const { return jacobian_; }
};
IMU ld_imu(const double *a) {
    IMU m;
    m.time   = a[0];
    m.dt     = a[1];
    m.dtheta = Vector3d(a[2], a[3], a[4]);
    m.dvel   = Vector3d(a[5], a[6], a[7]);
    m.odovel = 0;
    return m;
}
void st_state(const IntegrationState &st, double *o) {
    for (int i = 0; i < 3; i++) o[i] = st.p[i], o[7 + i] = st.v[i], o[10 + i] = st.bg[i], o[13 + i] = st.ba[i];
    o[3] = st.q.x(), o[4] = st.q.y(), o[5] = st.q.z(), o[6] = st.q.w();
}
} // namespace

void *ffref_preint_create(const double *noise, const double *imu0, const double *state16, double time) {
    auto par          = std::make_shared<IntegrationParameters>();
    par->acc_vrw      = noise[0];
    par->gyr_arw      = noise[1];
    par->gyr_bias_std = noise[2];
    par->acc_bias_std = noise[3];
    par->corr_time    = noise[4];
    par->gravity      = noise[5];
    IntegrationState st;
    st.time = time;
    st.p    = Vector3d(state16[0], state16[1], state16[2]);
    st.q    = Quaterniond(state16[6], state16[3], state16[4], state16[5]);
    st.v    = Vector3d(state16[7], state16[8], state16[9]);
    st.bg   = Vector3d(state16[10], state16[11], state16[12]);
    st.ba   = Vector3d(state16[13], state16[14], state16[15]);
    return new std::shared_ptr<PreintProbe>(std::make_shared<PreintProbe>(par, ld_imu(imu0), st));
}
void ffref_preint_destroy(void *h) { delete static_cast<std::shared_ptr<PreintProbe> *>(h); }
void ffref_preint_add(void *h, const double *imu) { (*static_cast<std::shared_ptr<PreintProbe> *>(h))->addNewImu(ld_imu(imu)); }
void ffref_preint_states(void *h, double *current16, double *delta16, double *delta_time) {
    auto &p = *static_cast<std::shared_ptr<PreintProbe> *>(h);
    st_state(p->currentState(), current16);
    st_state(p->deltaState(), delta16);
    *delta_time = p->deltaTime();
}
void ffref_preint_matrices(void *h, double *cov, double *jac) {
    auto &p = *static_cast<std::shared_ptr<PreintProbe> *>(h);
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) cov[15 * i + j] = p->cov()(i, j), jac[15 * i + j] = p->jac()(i, j);
}
/* residuals 15; Jacobians row-major 15 x 7, 15 x 9, 15 x 7, 15 x 9 (NULL: not requested) */
void ffref_preint_evaluate(void *h, const double *pose0, const double *mix0, const double *pose1, const double *mix1,
                           double *residuals, double *J0, double *J1, double *J2, double *J3) {
    auto &p = *static_cast<std::shared_ptr<PreintProbe> *>(h);
    PreintegrationFactor factor(p);
    const double *parameters[4] = {pose0, mix0, pose1, mix1};
    double *jacobians[4]        = {J0, J1, J2, J3};
    factor.Evaluate(parameters, residuals, (J0 || J1 || J2 || J3) ? jacobians : nullptr);
}

/* PoseManifold (factors/pose_manifold.cc:35-83): Plus, Minus and the two Jacobians (7 x 6 and 6 x 7, row-major) */
void ffref_pose_plus(const double *x, const double *delta, double *x_plus_delta) { PoseManifold().Plus(x, delta, x_plus_delta); }
void ffref_pose_minus(const double *y, const double *x, double *y_minus_x) { PoseManifold().Minus(y, x, y_minus_x); }
void ffref_pose_jacobians(const double *x, double *plus_7x6, double *minus_6x7) {
    PoseManifold m;
    m.PlusJacobian(x, plus_7x6);
    m.MinusJacobian(x, minus_6x7);
}

/* MISC::insMechanization (misc.cc:138-184), iswithearth = false: state = {p, q xyzw, v, bg, ba} (16 doubles) updated in
 * place; imu = {time, dt, dtheta[3], dvel[3]} */
void ffref_ins_mechanization(const double *gravity, const double *imu_pre, const double *imu_cur, double *state16,
                             double *state_time) {
    IntegrationConfiguration cfg{};
    cfg.iswithearth = false;
    cfg.gravity     = Vector3d(gravity[0], gravity[1], gravity[2]);
    auto ld         = [](const double *a) {
        IMU m;
        m.time   = a[0];
        m.dt     = a[1];
        m.dtheta = Vector3d(a[2], a[3], a[4]);
        m.dvel   = Vector3d(a[5], a[6], a[7]);
        m.odovel = 0;
        return m;
    };
    IntegrationState st;
    st.time = *state_time;
    st.p    = Vector3d(state16[0], state16[1], state16[2]);
    st.q    = Quaterniond(state16[6], state16[3], state16[4], state16[5]);
    st.v    = Vector3d(state16[7], state16[8], state16[9]);
    st.bg   = Vector3d(state16[10], state16[11], state16[12]);
    st.ba   = Vector3d(state16[13], state16[14], state16[15]);
    MISC::insMechanization(cfg, ld(imu_pre), ld(imu_cur), st);
    *state_time = st.time;
    for (int i = 0; i < 3; i++) state16[i] = st.p[i], state16[7 + i] = st.v[i];
    state16[3] = st.q.x(), state16[4] = st.q.y(), state16[5] = st.q.z(), state16[6] = st.q.w();
}
}
