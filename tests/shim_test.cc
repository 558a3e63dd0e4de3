// shim_test.cc — drives the C++ shim (ff-lins_b200/shim/lidar_shim.h) the way LINS::runFusion and
// Optimizer do: frames -> keyframe merge -> association -> LidarFactor::Evaluate / window blocks,
// and checks single-residual factors against the oracle. Needs a GPU to RUN; compiling it is part
// of the CPU test-suite (tests/test_boundary.py).
#include "../ff-lins_b200/shim/lidar_shim.h"
#include "../oracle/oracle.h"

#include <cmath>
#include <cstdio>
#include <random>

static PointCloudCustomPtr make_scan(int n, double t0, std::mt19937 &rng) {
    // points on the floor z = -1.5 and on a wall x = 8, seen from a slowly moving sensor
    std::uniform_real_distribution<double> U(-1, 1);
    auto pc = PointCloudCustomPtr(new PointCloudCustom);
    pc->resize(n);
    for (int i = 0; i < n; i++) {
        auto &p = pc->points[i];
        if (i % 2) {
            p.x = (float) (8.0 + 0.01 * U(rng));
            p.y = (float) (6 * U(rng));
            p.z = (float) (2 * U(rng));
        } else {
            p.x = (float) (3 + 5 * (U(rng) + 1));
            p.y = (float) (6 * U(rng));
            p.z = (float) (-1.5 + 0.01 * U(rng));
        }
        p.intensity = (float) i;
        p.time      = t0 + 0.1 * i / n;
    }
    return pc;
}

int main() {
    fflins_config cfg;
    fflins_config_default(&cfg);
    cfg.point_filter_num = 4;
    lidar::LidarMap map(cfg);
    std::mt19937 rng(7);
    Pose ext;
    const double T0 = 1000.0;
    std::vector<lidar::LidarFrame::Ptr> nonkeys;
    int n_keys = 0;
    for (int f = 0; f < 9; f++) {
        double start = T0 + 0.1 * f, end = start + 0.1;
        // INS window: straight motion along y at 1 m/s, identity attitude, 200 Hz
        std::deque<std::pair<IMU, IntegrationState>> win;
        for (int k = 0; k < 300; k++) {
            IMU imu;
            imu.time = end + 0.05 - 0.005 * (299 - k);
            IntegrationState st;
            st.time = imu.time;
            st.p    = Vector3d(0, 1.0 * (imu.time - T0), 0);
            win.emplace_back(imu, st);
        }
        auto frame = lidar::LidarFrame::createFrame(start, end, make_scan(4000, start, rng));
        frame->setTimeDelay(0.0);
        frame->setStamp(end);
        if (!map.lidarFrameProcessing(win, ext, frame)) return 2;
        if (frame->pointCloudLidar()->size() == 0 || frame->pointCloudMapFull()->size() != 4000) return 3;
        bool is_key = (f % 3 == 2);
        if (!is_key) {
            nonkeys.push_back(frame);
            continue;
        }
        map.mergeNonKeyframes(nonkeys, frame);
        nonkeys.clear();
        map.addNewKeyFrame(frame);
        frame->setVelocity(Vector3d(0, 1, 0));
        frame->setAngularVelocity(Vector3d(0, 0, 0.01));
        n_keys++;
        if (map.keyframes().size() >= 2 && !map.updateLidarFeaturesInMap()) return 4;
    }
    auto cur = map.keyframes().back();
    auto ref = map.keyframes().front();
    auto &feats = ref->features();
    int planes = 0, checked = 0;
    double worst = 0;
    double pose_ref[7] = {0, 0.3, 0, 0, 0, 0, 1}, pose_cur[7] = {0, 0.9, 0, 0, 0, 0, 1}, extp[7] = {0, 0, 0, 0, 0, 0, 1}, td = 0.001;
    const double *par[4] = {pose_ref, pose_cur, extp, &td};
    for (size_t k = 0; k < feats.size(); k++) {
        if (feats[k]->featureType() != lidar::LidarFeature::FEATURE_PLANE) continue;
        planes++;
        if (checked >= 16) continue;
        checked++;
        const auto &pt = cur->pointCloudLidar()->points[k];
        LidarFactor factor(pt, feats[k]->unitNormalVector(), feats[k]->normInverse(), feats[k]->pointStd(), ref->velocity(),
                           ref->angularVelocity(), ref->timeDelay(), cur->velocity(), cur->angularVelocity(), cur->timeDelay());
        double r, J0[7], J1[7], J2[7], J3[1];
        double *jac[4] = {J0, J1, nullptr, J3}; // a null block, like a constant Ceres parameter
        if (!factor.Evaluate(par, &r, jac)) return 5;
        float p3[3]      = {pt.x, pt.y, pt.z};
        double motion[14] = {0, 1, 0, 0, 0, 0.01, 0, 0, 1, 0, 0, 0, 0.01, 0};
        double ro, K0[7], K1[7], K2[7], K3[1];
        double *kj[4] = {K0, K1, K2, K3};
        orc_lidar_factor_evaluate(p3, feats[k]->unitNormalVector().data(), feats[k]->normInverse(), feats[k]->pointStd(), motion,
                                  par, &ro, kj);
        worst = std::fmax(worst, std::fabs(r - ro));
        for (int i = 0; i < 7; i++) worst = std::fmax(worst, std::fmax(std::fabs(J0[i] - K0[i]), std::fabs(J1[i] - K1[i])));
        worst = std::fmax(worst, std::fabs(J3[0] - K3[0]));
    }
    LidarWindowEvaluator ev(map.context());
    std::vector<uint64_t> ids;
    std::vector<double> poses;
    for (size_t i = 0; i + 1 < map.keyframes().size(); i++) {
        ids.push_back(map.keyframes()[i]->id());
        poses.insert(poses.end(), pose_ref, pose_ref + 7);
    }
    LidarWindowEvaluator::Blocks blk;
    if (!ev.evaluate(ids, poses, pose_cur, extp, td, FFLINS_HUBER, blk)) return 6;
    // ---- the two Ceres-shaped pair adapters against the per-residual sum -----------------------------
    // reference: every valid residual through LidarFactor::Evaluate, Huber corrector per residual on the host
    // (residual_block_info.cc:49-77, here the oracle's restatement), summed into cost / gradient / J^T J over the 19 tangent
    // columns [ref 6 | cur 6 | ext 6 | td]
    double Hs[361] = {0}, gs[19] = {0}, cs = 0;
    int P = 0;
    for (size_t k = 0; k < feats.size(); k++) {
        if (feats[k]->featureType() != lidar::LidarFeature::FEATURE_PLANE || feats[k]->isOutlier()) continue;
        const auto &pt = cur->pointCloudLidar()->points[k];
        LidarFactor factor(pt, feats[k]->unitNormalVector(), feats[k]->normInverse(), feats[k]->pointStd(), ref->velocity(),
                           ref->angularVelocity(), ref->timeDelay(), cur->velocity(), cur->angularVelocity(), cur->timeDelay());
        double r, J[22];
        double *jac[4] = {J, J + 7, J + 14, J + 21};
        if (!factor.Evaluate(par, &r, jac)) return 8;
        orc_huber_correct(1.0, &r, J);
        double jl[19];
        for (int a = 0; a < 18; a++) jl[a] = J[7 * (a / 6) + a % 6];
        jl[18] = J[21];
        for (int a = 0; a < 19; a++) {
            gs[a] += jl[a] * r;
            for (int c = 0; c < 19; c++) Hs[19 * a + c] += jl[a] * jl[c];
        }
        cs += r * r;
        P++;
    }
    auto sums = [](const LidarPairFactorBase &f, const double *const *par, double *H, double *g, double *c) -> bool {
        const int m = f.num_residuals();
        std::vector<double> r(m), J0(7 * m), J1(7 * m), J2(7 * m), J3(m);
        double *jac[4] = {J0.data(), J1.data(), J2.data(), J3.data()};
        if (!f.Evaluate(par, r.data(), jac)) return false;
        std::fill(H, H + 361, 0.0);
        std::fill(g, g + 19, 0.0);
        *c = 0;
        for (int i = 0; i < m; i++) {
            double jl[19];
            for (int a = 0; a < 6; a++) {
                jl[a]      = J0[7 * i + a];
                jl[6 + a]  = J1[7 * i + a];
                jl[12 + a] = J2[7 * i + a];
            }
            jl[18] = J3[i];
            if (J0[7 * i + 6] != 0 || J1[7 * i + 6] != 0 || J2[7 * i + 6] != 0) return false;   // column 6 must be zero
            for (int a = 0; a < 19; a++) {
                g[a] += jl[a] * r[i];
                for (int cc = 0; cc < 19; cc++) H[19 * a + cc] += jl[a] * jl[cc];
            }
            *c += r[i] * r[i];
        }
        return true;
    };
    auto rel = [](const double *a, const double *b, int n) {
        double num = 0, den = 0;
        for (int i = 0; i < n; i++) {
            num = std::fmax(num, std::fabs(a[i] - b[i]));
            den = std::fmax(den, std::fabs(b[i]));
        }
        return den > 0 ? num / den : num;
    };
    double Hb[361], gb[19], cb, Hq[361], gq[19], cq;
    LidarBatchedFactor batched(map.context(), ref->id(), cur->id());
    LidarSqrtFactor sqrtf(map.context(), ref->id(), cur->id());
    if (batched.num_residuals() != P || sqrtf.num_residuals() != 20) return 9;
    if (!sums(batched, par, Hb, gb, &cb) || !sums(sqrtf, par, Hq, gq, &cq)) return 10;
    const double eb = std::fmax(std::fmax(rel(Hb, Hs, 361), rel(gb, gs, 19)), std::fabs(cb - cs) / cs);
    const double eq = std::fmax(std::fmax(rel(Hq, Hs, 361), rel(gq, gs, 19)), std::fabs(cq - cs) / cs);
    std::printf("sqrt factor: H rel %.3e  g rel %.3e  cost rel %.3e | batched: H %.3e g %.3e cost %.3e\n", rel(Hq, Hs, 361), rel(gq, gs, 19),
                std::fabs(cq - cs) / cs, rel(Hb, Hs, 361), rel(gb, gs, 19), std::fabs(cb - cs) / cs);
    std::printf("SHIM_OK keyframes=%d planes=%d residuals=%d factor_max_abs_err=%.3e H00=%.6e adapters: P=%d batched_rel=%.2e sqrt_rel=%.2e\n",
                n_keys, planes, blk.n_res[0], worst, blk.H[0], P, eb, eq);
    return (planes > 20 && worst < 1e-9 && blk.n_res[0] > 0 && P > 20 && eb < 1e-9 && eq < 1e-9) ? 0 : 7;
}
