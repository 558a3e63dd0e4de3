// shim_bench.cc — the reference's own call sequence (LINS::runFusion, ff_lins.cc:236-283, 349-367; Optimizer, optimizer.cc:
// 85-185) driven through the C++ shim (ff-lins_b200/shim/lidar_shim.h): lidar::LidarMap::lidarFrameProcessing on pageable
// PointCloudCustom buffers, mergeNonKeyframes, addNewKeyFrame, updateLidarFeaturesInMap, 5 + 10 window evaluations with a
// chi-square cull between them (the Ceres solves' LiDAR share), re-projection of the window, removeOldestLidarFrame.
// This is the DROP-IN figure: what a maintainer gets by swapping lidar_map.h for the shim, nothing tuned around it.
//   shim_bench <frames.bin> <steps> <warmup> [sync]
// frames.bin is written by bench.py (write_shim_frames): header, then per frame the scalars, the INS window and the raw
// 32-byte PointXYZIT records.
#include "../ff-lins_b200/shim/lidar_shim.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct FrameData {
    double stamp, td, v[3], omega[3], pose7[7], start, end;
    std::deque<std::pair<IMU, IntegrationState>> window;
    PointCloudCustomPtr cloud;
};

static bool rd(FILE *f, void *p, size_t n) { return std::fread(p, 1, n, f) == n; }

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: shim_bench frames.bin steps warmup [sync]\n");
        return 2;
    }
    const int steps = std::atoi(argv[2]), warmup = std::atoi(argv[3]);
    const bool sync = argc > 4 && std::string(argv[4]) == "sync";
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[6];
    double Rbl[9], tbl[3], ext7[7];
    if (!rd(f, hdr, sizeof(hdr)) || hdr[0] != 0x46464c53 || !rd(f, Rbl, sizeof(Rbl)) || !rd(f, tbl, sizeof(tbl)) || !rd(f, ext7, sizeof(ext7)))
        return 4;
    const int n_frames = hdr[1], n_points = hdr[2], W = hdr[3], filter_num = hdr[4], per_key = hdr[5];
    std::vector<FrameData> frames(n_frames);
    std::vector<double> it(W), ip(3 * (size_t) W), iq(4 * (size_t) W);
    for (auto &fr : frames) {
        double sc[19];
        if (!rd(f, sc, sizeof(sc)) || !rd(f, it.data(), 8 * (size_t) W) || !rd(f, ip.data(), 24 * (size_t) W) || !rd(f, iq.data(), 32 * (size_t) W))
            return 5;
        fr.stamp = sc[0];
        fr.td    = sc[1];
        for (int i = 0; i < 3; i++) fr.v[i] = sc[2 + i], fr.omega[i] = sc[5 + i];
        for (int i = 0; i < 7; i++) fr.pose7[i] = sc[8 + i];
        fr.start = sc[15];
        fr.end   = sc[16];
        for (int k = 0; k < W; k++) {
            IMU imu;
            imu.time = it[k];
            IntegrationState st;
            st.time = it[k];
            st.p    = Vector3d(ip[3 * k], ip[3 * k + 1], ip[3 * k + 2]);
            for (int c = 0; c < 4; c++) st.q.c[c] = iq[4 * k + c];
            fr.window.emplace_back(imu, st);
        }
        fr.cloud = PointCloudCustomPtr(new PointCloudCustom);
        fr.cloud->resize(n_points);
        if (!rd(f, fr.cloud->points.data(), 32 * (size_t) n_points)) return 6;
    }
    std::fclose(f);

    fflins_config cfg;
    fflins_config_default(&cfg);
    cfg.point_filter_num     = filter_num;
    cfg.max_points_per_frame = n_points > 262144 ? n_points : 262144;
    lidar::LidarMap map(cfg);
    map.setAsyncFrames(!sync);
    Pose ext;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) ext.R(r, c) = Rbl[3 * r + c];
        ext.t[r] = tbl[r];
    }
    LidarWindowEvaluator ev(map.context());
    std::vector<lidar::LidarFrame::Ptr> nonkeys;
    std::vector<std::vector<double>> key_pose7;   // parallel to map.keyframes()
    size_t cursor = 0;
    LidarWindowEvaluator::Blocks blk;
    std::vector<int32_t> outl;
    long features = 0;

    auto step = [&]() -> bool {
        for (int j = 0; j < per_key; j++) {
            FrameData &fd = frames[cursor++ % frames.size()];
            auto frame    = lidar::LidarFrame::createFrame(fd.start, fd.end, fd.cloud);
            frame->setTimeDelay(fd.td);
            frame->setStamp(fd.stamp);
            if (!map.lidarFrameProcessing(fd.window, ext, frame)) return false;
            if (j + 1 < per_key) {
                nonkeys.push_back(frame);
                continue;
            }
            map.mergeNonKeyframes(nonkeys, frame);
            nonkeys.clear();
            map.addNewKeyFrame(frame);
            frame->setVelocity(Vector3d(fd.v[0], fd.v[1], fd.v[2]));
            frame->setAngularVelocity(Vector3d(fd.omega[0], fd.omega[1], fd.omega[2]));
            key_pose7.emplace_back(fd.pose7, fd.pose7 + 7);
            const auto &keys = map.keyframes();
            if (keys.size() >= 2) {
                if (!map.updateLidarFeaturesInMap()) return false;
                features += keys.back()->numFeatures();
                std::vector<uint64_t> ids;
                std::vector<double> poses;
                for (size_t i = 0; i + 1 < keys.size(); i++) {
                    ids.push_back(keys[i]->id());
                    poses.insert(poses.end(), key_pose7[i].begin(), key_pose7[i].end());
                }
                const double *pc = key_pose7.back().data();
                for (int it = 0; it < 5; it++)
                    if (!ev.evaluate(ids, poses, pc, ext7, fd.td, FFLINS_HUBER, blk)) return false;
                if (!ev.chi2(ids, poses, pc, ext7, fd.td, 3.841, outl)) return false;
                for (int it = 0; it < 10; it++)
                    if (!ev.evaluate(ids, poses, pc, ext7, fd.td, FFLINS_HUBER, blk)) return false;
                std::vector<Pose> world(keys.size());
                for (size_t i = 0; i < keys.size(); i++) world[i] = keys[i]->pose();
                map.reprojectWorld(world);
            }
            if (map.isWindowFull()) {
                map.removeOldestLidarFrame();
                key_pose7.erase(key_pose7.begin());
            }
        }
        return true;
    };
    for (int i = 0; i < warmup; i++)
        if (!step()) return 7;
    fflins_synchronize(map.context());
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; i++)
        if (!step()) return 8;
    fflins_synchronize(map.context());
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fflins_eval_stats es{};
    fflins_eval_stats_get(map.context(), &es);
    std::printf("{\"shim_frames_per_s\": %.2f, \"ms_per_step\": %.4f, \"steps\": %d, \"frames_per_step\": %d, \"points_per_frame\": %d, "
                "\"mode\": \"%s\", \"keyframes\": %zu, \"resident_evals\": %lld, \"launched_evals\": %lld, \"mean_plane_features\": %.0f}\n",
                per_key * steps / sec, sec * 1e3 / steps, steps, per_key, n_points, sync ? "synchronous frames" : "queued frames",
                map.keyframes().size(), (long long) es.resident_evals, (long long) es.launched_evals,
                (double) features / std::max(1, steps + warmup));
    return 0;
}
