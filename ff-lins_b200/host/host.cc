// host.cc — libfflins_host.so: the CPU side either side of the LiDAR hot path (include/fflins_host.h).
//
// Reference behaviour restated (not translated; no Eigen / Ceres here, plain arrays and a small dense toolkit):
//   MISC::insMechanization / redoInsMechanization / imuInterpolation / getImuSeriesFromTo   ff_lins/common/misc.cc:138-340
//   PreintegrationBase / PreintegrationNormal                ff_lins/preintegration/preintegration_{base,normal}.cc
//   LINS::addNewTimeNode / addNewLidarFrameTimeNode          ff_lins/ff_lins/ff_lins.cc:458-526
//   Optimizer::optimization / marginalization                ff_lins/ff_lins/optimizer.cc:85-185, 221-435
//   MarginalizationInfo / MarginalizationFactor              ff_lins/factors/marginalization_{info,factor}.cc
//   PoseManifold                                             ff_lins/factors/pose_manifold.cc:35-59
//   ImuPosePriorFactor / ImuMixPriorFactor                   ff_lins/preintegration/imu_{pose,mix}_prior_factor.h
//   LidarConverter                                           ROS/lidar/lidar_converter.cc:36-232
//   FileSaver text format                                    ff_lins/fileio/filesaver.cc:52-67
// The LiDAR factors never run here: the solver receives their per-pair 19x19 blocks from the caller's callback
// (fflins_window_eval on the GPU). Ceres' trust-region Levenberg-Marquardt is restated from its documented algorithm
// (Ceres is not in this image): damping diag(J^T J) / radius, step accepted when the relative decrease exceeds 1e-3,
// radius *= 1 / max(1/3, 1 - (2 rho - 1)^3) on success and halved (with a doubling factor) on failure.
#include "../../include/fflins_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <vector>

namespace {

// ---- small fixed-size algebra (quaternions x, y, z, w) -------------------------------------------------------------
struct V3 {
    double x = 0, y = 0, z = 0;
};
inline V3 v3(const double *a) { return {a[0], a[1], a[2]}; }
inline void st3(double *a, V3 v) {
    a[0] = v.x;
    a[1] = v.y;
    a[2] = v.z;
}
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return a * s; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
struct Q {
    double x = 0, y = 0, z = 0, w = 1;
};
inline Q qld(const double *a) { return {a[0], a[1], a[2], a[3]}; }
inline void qst(double *a, Q q) {
    a[0] = q.x;
    a[1] = q.y;
    a[2] = q.z;
    a[3] = q.w;
}
inline Q qmul(Q a, Q b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q qinv(Q q) {   // Eigen Quaternion::inverse: conjugate / squared norm
    double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
inline Q qnormalized(Q q) {
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
struct M3 {
    double m[9];
};
inline M3 q2R(Q q) {   // Eigen toRotationMatrix (no normalisation)
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
}
inline V3 mul(const M3 &A, V3 v) {
    return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
            A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline V3 qrot(Q q, V3 v) { return mul(q2R(q), v); }   // Eigen q * v
inline M3 skew(V3 v) { return {{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }
// Rotation::rotvec2quaternion (rotation.cc:61-65): Quaterniond(AngleAxisd(|v|, v.normalized()))
inline Q rotvec2q(V3 rv) {
    double a = norm(rv);
    if (a == 0.0) return {0, 0, 0, 1};
    double s = std::sin(a / 2) / a;
    return {rv.x * s, rv.y * s, rv.z * s, std::cos(a / 2)};
}
// bottom-right 3x3 of Rotation::quaternionleft / quaternionright (rotation.cc:88-106)
inline M3 qleft3(Q q) {
    M3 s = skew({q.x, q.y, q.z});
    for (int i = 0; i < 9; i++) s.m[i] += (i % 4 == 0) ? q.w : 0.0;
    return s;
}
inline M3 qright3(Q q) {
    M3 s = skew({q.x, q.y, q.z});
    for (int i = 0; i < 9; i++) s.m[i] = ((i % 4 == 0) ? q.w : 0.0) - s.m[i];
    return s;
}
inline M3 mm(const M3 &A, const M3 &B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}

// ---- dense row-major matrices (the solver's normal equations are <= ~180 x 180) -------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t) r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t) i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t) i * c + j]; }
    static Mat identity(int n) {
        Mat I(n, n);
        for (int i = 0; i < n; i++) I(i, i) = 1;
        return I;
    }
};
Mat operator*(const Mat &A, const Mat &B) {
    Mat C(A.r, B.c);
    for (int i = 0; i < A.r; i++)
        for (int k = 0; k < A.c; k++) {
            const double aik = A(i, k);
            if (aik == 0.0) continue;
            for (int j = 0; j < B.c; j++) C(i, j) += aik * B(k, j);
        }
    return C;
}
Mat transpose(const Mat &A) {
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; i++)
        for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
    return T;
}
void set_block(Mat &A, int r0, int c0, const M3 &B, double s = 1.0) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A(r0 + i, c0 + j) = s * B.m[3 * i + j];
}
M3 get_block(const Mat &A, int r0, int c0) {
    M3 B;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) B.m[3 * i + j] = A(r0 + i, c0 + j);
    return B;
}
// Cholesky A = L L^T (lower); false if not positive definite
bool cholesky(const Mat &A, Mat &L) {
    const int n = A.r;
    L = Mat(n, n);
    for (int j = 0; j < n; j++) {
        double d = A(j, j);
        for (int k = 0; k < j; k++) d -= L(j, k) * L(j, k);
        if (!(d > 0.0)) return false;
        const double ljj = std::sqrt(d);
        L(j, j) = ljj;
        for (int i = j + 1; i < n; i++) {
            double s = A(i, j);
            for (int k = 0; k < j; k++) s -= L(i, k) * L(j, k);
            L(i, j) = s / ljj;
        }
    }
    return true;
}
void chol_solve(const Mat &L, const double *b, double *x) {
    const int n = L.r;
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L(i, k) * y[k];
        y[i] = s / L(i, i);
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = y[i];
        for (int k = i + 1; k < n; k++) s -= L(k, i) * x[k];
        x[i] = s / L(i, i);
    }
}
Mat spd_inverse(const Mat &A) {
    Mat L;
    const int n = A.r;
    Mat inv(n, n);
    if (!cholesky(A, L)) return inv;
    std::vector<double> e(n), x(n);
    for (int j = 0; j < n; j++) {
        std::fill(e.begin(), e.end(), 0.0);
        e[j] = 1.0;
        chol_solve(L, e.data(), x.data());
        for (int i = 0; i < n; i++) inv(i, j) = x[i];
    }
    return inv;
}
// symmetric eigen-decomposition by cyclic Jacobi rotations: A = V diag(w) V^T (what SelfAdjointEigenSolver provides)
void sym_eig(Mat A, std::vector<double> &w, Mat &V) {
    const int n = A.r;
    V = Mat::identity(n);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; i++) {
            diag += A(i, i) * A(i, i);
            for (int j = i + 1; j < n; j++) off += A(i, j) * A(i, j);
        }
        if (off <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A(p, q);
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (A(q, q) - A(p, p)) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A(k, p), akq = A(k, q);
                    A(k, p) = c * akp - s * akq;
                    A(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A(p, k), aqk = A(q, k);
                    A(p, k) = c * apk - s * aqk;
                    A(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; i++) w[i] = A(i, i);
}

constexpr double kMinTimeInterval = 0.0001;   // MISC::MINIMUM_TIME_INTERVAL (misc.h)
constexpr double kMargEps         = 1e-8;     // MarginalizationInfo::EPS

// ---- INS -----------------------------------------------------------------------------------------------------------
void mechanize(const fflins_ins_config &cfg, const fflins_imu &pre, const fflins_imu &cur, fflins_nav_state &s) {
    // bias compensation (misc.cc:141-146)
    const V3 bg = v3(s.bg), ba = v3(s.ba);
    const V3 cth = v3(cur.dtheta) - cur.dt * bg, cdv = v3(cur.dvel) - cur.dt * ba;
    const V3 pth = v3(pre.dtheta) - pre.dt * bg, pdv = v3(pre.dvel) - pre.dt * ba;
    const double dt = cur.dt;
    s.time = cur.time;
    // two-sample coning / sculling (:151-154)
    const V3 dvfb   = cdv + 0.5 * cross(cth, cdv) + (1.0 / 12.0) * (cross(pth, cdv) + cross(pdv, cth));
    const V3 dtheta = cth + (1.0 / 12.0) * cross(pth, cth);
    Q q = qld(s.q);
    const V3 dvel = qrot(q, dvfb) + v3(cfg.gravity) * dt;   // (:171)
    q = qnormalized(qmul(q, rotvec2q(dtheta)));              // (:174-175)
    qst(s.q, q);
    const V3 v = v3(s.v);
    st3(s.p, v3(s.p) + dt * v + 0.5 * dt * dvel);            // mean velocity of the two epochs (:179)
    st3(s.v, v + dvel);
}

int need_interpolation(const fflins_imu &imu0, const fflins_imu &imu1, double mid) {   // misc.cc:240-264
    if (imu0.time < mid && imu1.time > mid) {
        if (mid - imu0.time < kMinTimeInterval) return -1;
        if (imu1.time - mid < kMinTimeInterval) return 1;
        return 2;
    }
    return 0;
}
void interpolate(const fflins_imu &imu01, fflins_imu &imu00, fflins_imu &imu11, double mid) {   // misc.cc:266-284
    const fflins_imu buff = imu01;
    const double scale    = (buff.time - mid) / buff.dt;
    imu00.time = mid;
    imu00.dt   = buff.dt - (buff.time - mid);
    imu11.time = buff.time;
    imu11.dt   = buff.time - mid;
    for (int k = 0; k < 3; k++) {
        imu00.dtheta[k] = buff.dtheta[k] * (1 - scale);
        imu00.dvel[k]   = buff.dvel[k] * (1 - scale);
        imu11.dtheta[k] = buff.dtheta[k] * scale;
        imu11.dvel[k]   = buff.dvel[k] * scale;
    }
    imu00.odovel = buff.odovel * (1 - scale);
    imu11.odovel = buff.odovel * scale;
}

} // namespace

struct fflins_ins_window {
    fflins_ins_config cfg;
    std::deque<std::pair<fflins_imu, fflins_nav_state>> w;
    // MISC::getInsWindowIndex (misc.cc:27-62): first index with time[idx-1] <= t < time[idx]; 0 outside the window
    size_t index(double t) const {
        if (w.size() < 2 || t < w.front().first.time || t >= w.back().first.time) return 0;
        size_t lo = 0, hi = w.size() - 1;   // invariant: time[lo] <= t < time[hi]
        while (hi - lo > 1) {
            size_t mid = (lo + hi) / 2;
            if (w[mid].first.time <= t) lo = mid;
            else hi = mid;
        }
        return hi;
    }
};

namespace {

// ---- preintegration ------------------------------------------------------------------------------------------------
constexpr int NS = 15, NN = 12;   // state dp dv dq dbg dba; noise nw na nbg nba

} // namespace

struct fflins_preint {
    fflins_imu_noise par;
    fflins_nav_state cur, delta;       // current_state_, delta_state_ (delta.bg / ba = linearisation biases)
    std::vector<fflins_imu> buf;
    double delta_time = 0, start_time = 0, end_time = 0;
    Mat jac, cov, noise, sqrt_info;
    V3 gravity;
    Q corrected_q;
    V3 corrected_p, corrected_v;

    void reset_state(const fflins_nav_state &s) {   // preintegration_normal.cc:302-312
        delta_time = 0;
        delta      = fflins_nav_state{};
        delta.q[3] = 1;
        std::memcpy(delta.bg, s.bg, 24);
        std::memcpy(delta.ba, s.ba, 24);
        jac = Mat::identity(NS);
        cov = Mat(NS, NS);
    }
    fflins_imu compensate(const fflins_imu &imu) const {   // compensationBias
        fflins_imu o = imu;
        for (int k = 0; k < 3; k++) {
            o.dtheta[k] -= o.dt * delta.bg[k];
            o.dvel[k] -= o.dt * delta.ba[k];
        }
        return o;
    }
    void integrate(const fflins_imu &pre, const fflins_imu &imu) {   // preintegration_base.cc:39-72
        const double dt = imu.dt;
        delta_time += dt;
        end_time = imu.time;
        cur.time = imu.time;
        const V3 cth = v3(imu.dtheta), cdv = v3(imu.dvel), pth = v3(pre.dtheta), pdv = v3(pre.dvel);
        const V3 dvfb = cdv + 0.5 * cross(cth, cdv) + (1.0 / 12.0) * (cross(pth, cdv) + cross(pdv, cth));
        V3 dvel = qrot(qld(cur.q), dvfb) + gravity * dt;
        st3(cur.p, v3(cur.p) + dt * v3(cur.v) + 0.5 * dt * dvel);
        st3(cur.v, v3(cur.v) + dvel);
        const V3 dtheta = cth + (1.0 / 12.0) * cross(pth, cth);
        qst(cur.q, qnormalized(qmul(qld(cur.q), rotvec2q(dtheta))));
        dvel = qrot(qld(delta.q), dvfb);
        st3(delta.p, v3(delta.p) + dt * v3(delta.v) + 0.5 * dt * dvel);
        st3(delta.v, v3(delta.v) + dvel);
        qst(delta.q, qnormalized(qmul(qld(delta.q), rotvec2q(dtheta))));
    }
    void propagate(const fflins_imu &imu) {   // updateJacobianAndCovariance, preintegration_normal.cc:262-300
        const double dt = imu.dt;
        Mat phi(NS, NS);
        const M3 I = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
        const M3 Rd = q2R(qld(delta.q));
        set_block(phi, 0, 0, I);
        set_block(phi, 0, 3, I, dt);
        set_block(phi, 3, 3, I);
        set_block(phi, 3, 6, mm(Rd, skew(v3(imu.dvel))), -1.0);
        set_block(phi, 3, 12, Rd, -dt);
        {
            M3 a = skew(v3(imu.dtheta));
            for (int i = 0; i < 9; i++) a.m[i] = I.m[i] - a.m[i];
            set_block(phi, 6, 6, a);
        }
        set_block(phi, 6, 9, I, -dt);
        set_block(phi, 9, 9, I, 1 - dt / par.corr_time);
        set_block(phi, 12, 12, I, 1 - dt / par.corr_time);
        jac = phi * jac;
        Mat gt(NS, NN);
        set_block(gt, 3, 3, Rd);
        set_block(gt, 6, 0, I);
        set_block(gt, 9, 6, I);
        set_block(gt, 12, 9, I);
        const Mat gq  = gt * noise * transpose(gt);
        const Mat a   = phi * gq, b = gq * transpose(phi);
        Mat pcp       = phi * cov * transpose(phi);
        for (size_t i = 0; i < pcp.a.size(); i++) pcp.a[i] += 0.5 * dt * (a.a[i] + b.a[i]);
        cov = pcp;
    }
    void process(size_t index) {   // integrationProcess
        const fflins_imu pre = compensate(buf[index - 1]), imu = compensate(buf[index]);
        integrate(pre, imu);
        propagate(imu);
    }
};

namespace {

// evaluate + the four Jacobians (preintegration_normal.cc:36-140). J* row-major 15 x {7, 9, 7, 9}, any may be null.
void preint_evaluate(fflins_preint &P, const double *pose0, const double *mix0, const double *pose1, const double *mix1,
                     double *residual, double *Jp0, double *Jm0, double *Jp1, double *Jm1) {
    // sqrt_information = LLT(covariance^-1).matrixL().transpose()
    {
        Mat L;
        const Mat inv = spd_inverse(P.cov);
        if (!cholesky(inv, L)) L = Mat::identity(NS);
        P.sqrt_info = transpose(L);
    }
    const M3 dp_dbg = get_block(P.jac, 0, 9), dp_dba = get_block(P.jac, 0, 12), dv_dbg = get_block(P.jac, 3, 9),
             dv_dba = get_block(P.jac, 3, 12), dq_dbg = get_block(P.jac, 6, 9);
    const V3 p0 = v3(pose0), p1 = v3(pose1), v0 = v3(mix0), v1 = v3(mix1);
    const Q q0 = qld(pose0 + 3), q1 = qld(pose1 + 3);
    const V3 bg0 = v3(mix0 + 3), ba0 = v3(mix0 + 6), bg1 = v3(mix1 + 3), ba1 = v3(mix1 + 6);
    const V3 dbg = bg0 - v3(P.delta.bg), dba = ba0 - v3(P.delta.ba);
    const double T = P.delta_time;
    P.corrected_p = v3(P.delta.p) + mul(dp_dba, dba) + mul(dp_dbg, dbg);
    P.corrected_v = v3(P.delta.v) + mul(dv_dba, dba) + mul(dv_dbg, dbg);
    P.corrected_q = qmul(qld(P.delta.q), rotvec2q(mul(dq_dbg, dbg)));
    const Q q0i   = qinv(q0);
    const V3 g    = P.gravity;
    const V3 a_p  = qrot(q0i, p1 - p0 - v0 * T - 0.5 * T * T * g);
    const V3 a_v  = qrot(q0i, v1 - v0 - g * T);
    double r[NS];
    st3(r + 0, a_p - P.corrected_p);
    st3(r + 3, a_v - P.corrected_v);
    {
        const Q e = qmul(qmul(qinv(P.corrected_q), q0i), q1);
        r[6] = 2 * e.x;
        r[7] = 2 * e.y;
        r[8] = 2 * e.z;
    }
    st3(r + 9, bg1 - bg0);
    st3(r + 12, ba1 - ba0);
    const Mat &W = P.sqrt_info;
    if (residual)
        for (int i = 0; i < NS; i++) {
            double s = 0;
            for (int k = 0; k < NS; k++) s += W(i, k) * r[k];
            residual[i] = s;
        }
    auto emit = [&](const Mat &J, double *out) {
        const Mat WJ = W * J;
        std::memcpy(out, WJ.a.data(), sizeof(double) * WJ.a.size());
    };
    const M3 R0i = q2R(q0i);
    if (Jp0) {
        Mat J(NS, 7);
        set_block(J, 0, 0, R0i, -1.0);
        set_block(J, 0, 3, skew(a_p));
        set_block(J, 3, 3, skew(a_v));
        {   // bottom-right 3x3 of the 4x4 product quaternionleft(a) * quaternionright(b): L3(a) R3(b) - vec(a) vec(b)^T
            const Q a = qmul(qinv(q1), q0), b = P.corrected_q;
            M3 LR = mm(qleft3(a), qright3(b));
            const double va[3] = {a.x, a.y, a.z}, vb[3] = {b.x, b.y, b.z};
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) LR.m[3 * i + j] -= va[i] * vb[j];
            set_block(J, 6, 3, LR, -1.0);
        }
        emit(J, Jp0);
    }
    if (Jp1) {
        Mat J(NS, 7);
        set_block(J, 0, 0, R0i);
        set_block(J, 6, 3, qleft3(qmul(qmul(qinv(P.corrected_q), q0i), q1)));
        emit(J, Jp1);
    }
    if (Jm0) {
        Mat J(NS, 9);
        const M3 I = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
        set_block(J, 0, 0, R0i, -T);
        set_block(J, 0, 3, dp_dbg, -1.0);
        set_block(J, 0, 6, dp_dba, -1.0);
        set_block(J, 3, 0, R0i, -1.0);
        set_block(J, 3, 3, dv_dbg, -1.0);
        set_block(J, 3, 6, dv_dba, -1.0);
        set_block(J, 6, 3, mm(qleft3(qmul(qmul(qinv(q1), q0), qld(P.delta.q))), dq_dbg), -1.0);
        set_block(J, 9, 3, I, -1.0);
        set_block(J, 12, 6, I, -1.0);
        emit(J, Jm0);
    }
    if (Jm1) {
        Mat J(NS, 9);
        const M3 I = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};
        set_block(J, 3, 0, R0i);
        set_block(J, 9, 3, I);
        set_block(J, 12, 6, I);
        emit(J, Jm1);
    }
}

// PoseManifold::Plus (pose_manifold.cc:35-50)
void pose_plus(const double *x, const double *d, double *out) {
    for (int i = 0; i < 3; i++) out[i] = x[i] + d[i];
    qst(out + 3, qnormalized(qmul(qld(x + 3), rotvec2q(v3(d + 3)))));
}
// MarginalizationFactor's dx for a pose block (marginalization_factor.cc:54-64)
void pose_minus(const double *x, const double *x0, double *d) {
    for (int i = 0; i < 3; i++) d[i] = x[i] - x0[i];
    const Q dq = qmul(qinv(qld(x0 + 3)), qld(x + 3));
    const double s = dq.w < 0 ? -2.0 : 2.0;
    d[3] = s * dq.x;
    d[4] = s * dq.y;
    d[5] = s * dq.z;
}

} // namespace

// ---- the sliding-window solver -------------------------------------------------------------------------------------
// Local (tangent) vector of the window: [node 0: pose 6, mix 9 | node 1 ... | ext 6 | td 1].
struct fflins_solver {
    struct Node {
        double time;
        double pose[7], mix[9];
        uint64_t lidar_id;
        std::unique_ptr<fflins_preint> preint;   // from the previous node to this one (null for node 0)
        uint64_t uid;                             // stable identity across marginalizations
    };
    fflins_solver_options opt;
    std::deque<Node> nodes;
    double ext[7] = {0, 0, 0, 0, 0, 0, 1}, td = 0;
    uint64_t next_uid = 1;
    // initial priors (is_use_prior_): ImuPosePriorFactor + ImuMixPriorFactor on node 0
    bool use_prior = false;
    double prior_pose[7], prior_pose_std[6], prior_mix[9], prior_mix_std[9];
    // marginalization prior: e = e0 + J0 dx over `pcols` local columns: per kept block its uid (0 = ext, 1 = td; pose / mix
    // flagged), linearisation point, local size
    struct PBlock {
        uint64_t uid;
        int kind;   // 0 node pose, 1 node mix, 2 ext, 3 td
        int size;   // local
        double x0[9];
    };
    bool have_marg = false;
    std::vector<PBlock> pblocks;
    Mat J0;
    std::vector<double> e0;

    int n_local() const { return 15 * (int) nodes.size() + 7; }
    int off_pose(int k) const { return 15 * k; }
    int off_mix(int k) const { return 15 * k + 6; }
    int off_ext() const { return 15 * (int) nodes.size(); }
    int off_td() const { return 15 * (int) nodes.size() + 6; }
    bool window_full() const {
        int lidar = 0;
        for (auto &n : nodes) lidar += n.lidar_id != UINT64_MAX;
        return lidar > opt.window_size;   // LidarMap::isWindowFull: size > window
    }
};

namespace {

struct Problem {   // one linearisation
    Mat H;
    std::vector<double> b;
    double cost = 0;
};

struct LidarLink {
    std::vector<uint64_t> ref_ids;
    std::vector<int> ref_nodes;
    int cur_node = -1;
};

LidarLink lidar_link(const fflins_solver &S) {
    LidarLink L;
    for (int k = (int) S.nodes.size() - 1; k >= 0; k--)
        if (S.nodes[k].lidar_id != UINT64_MAX) {
            L.cur_node = k;
            break;
        }
    if (L.cur_node < 0) return L;
    for (int k = 0; k < L.cur_node; k++)
        if (S.nodes[k].lidar_id != UINT64_MAX) {
            L.ref_ids.push_back(S.nodes[k].lidar_id);
            L.ref_nodes.push_back(k);
        }
    return L;
}

void add_dense(Problem &P, const double *J, int m, const int *cols, int nc, const double *r) {
    // H += J^T J, b -= J^T r, cost += 0.5 r^T r ; J is m x nc (row-major), cols map to the window's local vector
    for (int a = 0; a < nc; a++) {
        if (cols[a] < 0) continue;
        double g = 0;
        for (int i = 0; i < m; i++) g += J[i * nc + a] * r[i];
        P.b[cols[a]] -= g;
        for (int c = a; c < nc; c++) {
            if (cols[c] < 0) continue;
            double s = 0;
            for (int i = 0; i < m; i++) s += J[i * nc + a] * J[i * nc + c];
            P.H(cols[a], cols[c]) += s;
            if (cols[a] != cols[c]) P.H(cols[c], cols[a]) += s;
        }
    }
    for (int i = 0; i < m; i++) P.cost += 0.5 * r[i] * r[i];
}

// IMU factor between node k and k + 1 (PreintegrationFactor): 15 residuals, local Jacobian 15 x 30
void add_imu_factor(const fflins_solver &S, int k, Problem &P) {
    fflins_preint &pre = *S.nodes[k + 1].preint;
    double r[15], Jp0[105], Jm0[135], Jp1[105], Jm1[135];
    preint_evaluate(pre, S.nodes[k].pose, S.nodes[k].mix, S.nodes[k + 1].pose, S.nodes[k + 1].mix, r, Jp0, Jm0, Jp1, Jm1);
    double J[15 * 30];
    int cols[30];
    for (int i = 0; i < 15; i++) {
        for (int j = 0; j < 6; j++) J[i * 30 + j] = Jp0[i * 7 + j];          // PlusJacobian = [I6; 0]: first 6 columns
        for (int j = 0; j < 9; j++) J[i * 30 + 6 + j] = Jm0[i * 9 + j];
        for (int j = 0; j < 6; j++) J[i * 30 + 15 + j] = Jp1[i * 7 + j];
        for (int j = 0; j < 9; j++) J[i * 30 + 21 + j] = Jm1[i * 9 + j];
    }
    for (int j = 0; j < 15; j++) {
        cols[j]      = S.off_pose(k) + j;
        cols[15 + j] = S.off_pose(k + 1) + j;
    }
    add_dense(P, J, 15, cols, 30, r);
}

// ImuPosePriorFactor (imu_pose_prior_factor.h:42-68) + ImuMixPriorFactor on node 0
void add_initial_priors(const fflins_solver &S, Problem &P) {
    const auto &n = S.nodes[0];
    double r[6], J[36] = {0};
    int cols[6];
    for (int k = 0; k < 3; k++) r[k] = (n.pose[k] - S.prior_pose[k]) / S.prior_pose_std[k];
    const Q e = qmul(qinv(qld(n.pose + 3)), qld(S.prior_pose + 3));
    const double ev[3] = {e.x, e.y, e.z};
    const M3 R = qright3(e);
    for (int k = 0; k < 3; k++) r[3 + k] = 2 * ev[k] / S.prior_pose_std[3 + k];
    for (int k = 0; k < 3; k++) J[6 * k + k] = 1.0 / S.prior_pose_std[k];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) J[6 * (3 + i) + 3 + j] = -R.m[3 * i + j] / S.prior_pose_std[3 + i];
    for (int j = 0; j < 6; j++) cols[j] = S.off_pose(0) + j;
    add_dense(P, J, 6, cols, 6, r);
    double rm[9], Jm[81] = {0};
    int cm[9];
    for (int k = 0; k < 9; k++) {
        rm[k]         = (n.mix[k] - S.prior_mix[k]) / S.prior_mix_std[k];
        Jm[9 * k + k] = 1.0 / S.prior_mix_std[k];
        cm[k]         = S.off_mix(0) + k;
    }
    add_dense(P, Jm, 9, cm, 9, rm);
}

// where a prior block lives in the current window (-1: gone)
int prior_block_offset(const fflins_solver &S, const fflins_solver::PBlock &pb, const double **x) {
    if (pb.kind == 2) {
        *x = S.ext;
        return S.off_ext();
    }
    if (pb.kind == 3) {
        *x = &S.td;
        return S.off_td();
    }
    for (int k = 0; k < (int) S.nodes.size(); k++)
        if (S.nodes[k].uid == pb.uid) {
            *x = pb.kind == 0 ? S.nodes[k].pose : S.nodes[k].mix;
            return pb.kind == 0 ? S.off_pose(k) : S.off_mix(k);
        }
    return -1;
}

// MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91): e = e0 + J0 dx
void add_marg_prior(const fflins_solver &S, Problem &P) {
    const int m = S.J0.r, nc = S.J0.c;
    std::vector<double> dx(nc, 0.0), r(S.e0);
    std::vector<int> cols(nc, -1);
    int c0 = 0;
    for (const auto &pb : S.pblocks) {
        const double *x = nullptr;
        const int off   = prior_block_offset(S, pb, &x);
        if (off >= 0) {
            if (pb.kind == 0 || pb.kind == 2) pose_minus(x, pb.x0, dx.data() + c0);
            else
                for (int i = 0; i < pb.size; i++) dx[c0 + i] = x[i] - pb.x0[i];
            for (int i = 0; i < pb.size; i++) cols[c0 + i] = off + i;
        }
        c0 += pb.size;
    }
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int j = 0; j < nc; j++) s += S.J0(i, j) * dx[j];
        r[i] += s;
    }
    add_dense(P, S.J0.a.data(), m, cols.data(), nc, r.data());
}

// the LiDAR share: per-pair robustified blocks from the caller (GPU), scattered into the window's system
int add_lidar(const fflins_solver &S, const LidarLink &L, fflins_lidar_eval_fn eval, void *user, uint32_t flags, Problem &P,
              const std::vector<int> *only_pairs = nullptr) {
    const int n = (int) L.ref_ids.size();
    if (n == 0 || !eval) return 0;
    std::vector<double> pose_refs(7 * (size_t) n), H(361 * (size_t) n), b(19 * (size_t) n), cost(2 * (size_t) n);
    std::vector<int32_t> nres(n);
    for (int i = 0; i < n; i++) std::memcpy(pose_refs.data() + 7 * i, S.nodes[L.ref_nodes[i]].pose, 56);
    int rc = eval(user, n, L.ref_ids.data(), pose_refs.data(), S.nodes[L.cur_node].pose, S.ext, S.td, flags, H.data(), b.data(),
                  cost.data(), nres.data());
    if (rc != 0) return rc;
    for (int i = 0; i < n; i++) {
        if (only_pairs && std::find(only_pairs->begin(), only_pairs->end(), i) == only_pairs->end()) continue;
        int cols[19];
        for (int j = 0; j < 6; j++) {
            cols[j]      = S.off_pose(L.ref_nodes[i]) + j;
            cols[6 + j]  = S.off_pose(L.cur_node) + j;
            cols[12 + j] = S.off_ext() + j;
        }
        cols[18] = S.off_td();
        const double *Hi = H.data() + 361 * (size_t) i, *bi = b.data() + 19 * (size_t) i;
        for (int a = 0; a < 19; a++) {
            P.b[cols[a]] += bi[a];   // b = -J^T r already
            for (int c = 0; c < 19; c++) P.H(cols[a], cols[c]) += Hi[19 * a + c];
        }
        P.cost += cost[2 * i];
    }
    return 0;
}

uint32_t lidar_flags(const fflins_solver &S, bool ext_const, bool td_const) {
    (void) S;
    return 1u /* HUBER */ | (ext_const ? 8u : 0u) | (td_const ? 16u : 0u);
}

int linearize(const fflins_solver &S, const LidarLink &L, fflins_lidar_eval_fn eval, void *user, bool ext_const, bool td_const,
              Problem &P) {
    const int n = S.n_local();
    P.H = Mat(n, n);
    P.b.assign(n, 0.0);
    P.cost = 0;
    for (int k = 0; k + 1 < (int) S.nodes.size(); k++) add_imu_factor(S, k, P);
    if (S.use_prior) add_initial_priors(S, P);
    if (S.have_marg) add_marg_prior(S, P);
    return add_lidar(S, L, eval, user, lidar_flags(S, ext_const, td_const), P);
}

struct Snapshot {
    std::vector<double> poses, mixes;
    double ext[7], td;
};
Snapshot snapshot(const fflins_solver &S) {
    Snapshot s;
    for (auto &n : S.nodes) {
        s.poses.insert(s.poses.end(), n.pose, n.pose + 7);
        s.mixes.insert(s.mixes.end(), n.mix, n.mix + 9);
    }
    std::memcpy(s.ext, S.ext, 56);
    s.td = S.td;
    return s;
}
void restore(fflins_solver &S, const Snapshot &s) {
    for (size_t k = 0; k < S.nodes.size(); k++) {
        std::memcpy(S.nodes[k].pose, s.poses.data() + 7 * k, 56);
        std::memcpy(S.nodes[k].mix, s.mixes.data() + 9 * k, 72);
    }
    std::memcpy(S.ext, s.ext, 56);
    S.td = s.td;
}
void apply_step(fflins_solver &S, const std::vector<double> &dx, bool ext_const, bool td_const) {
    for (size_t k = 0; k < S.nodes.size(); k++) {
        double out[7];
        pose_plus(S.nodes[k].pose, dx.data() + S.off_pose((int) k), out);
        std::memcpy(S.nodes[k].pose, out, 56);
        for (int i = 0; i < 9; i++) S.nodes[k].mix[i] += dx[S.off_mix((int) k) + i];
    }
    if (!ext_const) {
        double out[7];
        pose_plus(S.ext, dx.data() + S.off_ext(), out);
        std::memcpy(S.ext, out, 56);
    }
    if (!td_const) S.td += dx[S.off_td()];
}

// One ceres::Solve with LEVENBERG_MARQUARDT (restated from the documented algorithm, see the file header).
int lm_solve(fflins_solver &S, const LidarLink &L, fflins_lidar_eval_fn eval, void *user, bool ext_const, bool td_const,
             int max_iters, int *iters, int *successful, double *cost0, double *cost1) {
    Problem P;
    int rc = linearize(S, L, eval, user, ext_const, td_const, P);
    if (rc) return rc;
    *cost0 = P.cost;
    double radius = 1e4, decrease = 2.0;
    const int n = S.n_local();
    *iters = *successful = 0;
    for (int it = 0; it < max_iters; it++) {
        (*iters)++;
        // columns of constant blocks carry no information: pin them
        std::vector<char> fixed(n, 0);
        if (ext_const)
            for (int j = 0; j < 6; j++) fixed[S.off_ext() + j] = 1;
        if (td_const) fixed[S.off_td()] = 1;
        Mat A = P.H;
        std::vector<double> g = P.b;
        double gmax = 0;
        for (int i = 0; i < n; i++) {
            if (fixed[i]) {
                for (int j = 0; j < n; j++) A(i, j) = A(j, i) = 0;
                A(i, i) = 1;
                g[i]    = 0;
            }
            gmax = std::max(gmax, std::fabs(g[i]));
        }
        if (gmax < 1e-10) break;   // gradient tolerance
        for (int i = 0; i < n; i++)
            if (!fixed[i]) A(i, i) += std::min(std::max(P.H(i, i), 1e-6), 1e32) / radius;
        Mat Lc;
        std::vector<double> dx(n, 0.0);
        bool ok = cholesky(A, Lc);
        if (ok) chol_solve(Lc, g.data(), dx.data());
        double model = 0;   // model decrease = dx^T (b - 0.5 H dx)
        if (ok) {
            for (int i = 0; i < n; i++) {
                if (fixed[i]) continue;
                double hd = 0;
                for (int j = 0; j < n; j++)
                    if (!fixed[j]) hd += P.H(i, j) * dx[j];
                model += dx[i] * (P.b[i] - 0.5 * hd);
            }
        }
        if (!ok || !(model > 0)) {
            radius /= decrease;
            decrease *= 2;
            continue;
        }
        const Snapshot snap = snapshot(S);
        apply_step(S, dx, ext_const, td_const);
        Problem Pn;
        rc = linearize(S, L, eval, user, ext_const, td_const, Pn);
        if (rc) {
            restore(S, snap);
            return rc;
        }
        const double rho = (P.cost - Pn.cost) / model;
        if (rho > 1e-3) {   // min_relative_decrease
            (*successful)++;
            const double dcost = P.cost - Pn.cost;
            double xn = 0, dn = 0;
            for (int i = 0; i < n; i++) dn += dx[i] * dx[i];
            for (auto &nd : S.nodes)
                for (int i = 0; i < 7; i++) xn += nd.pose[i] * nd.pose[i];
            P = std::move(Pn);
            radius   = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2 * rho - 1, 3));
            decrease = 2.0;
            if (std::fabs(dcost) <= 1e-6 * P.cost) break;                      // function tolerance
            if (std::sqrt(dn) <= 1e-8 * (std::sqrt(xn) + 1e-8)) break;         // parameter tolerance
        } else {
            restore(S, snap);
            radius /= decrease;
            decrease *= 2;
        }
    }
    *cost1 = P.cost;
    return 0;
}

} // namespace

extern "C" {

void fflins_ins_mechanization(const fflins_ins_config *cfg, const fflins_imu *imu_pre, const fflins_imu *imu_cur,
                              fflins_nav_state *state) {
    mechanize(*cfg, *imu_pre, *imu_cur, *state);
}
int fflins_imu_need_interpolation(const fflins_imu *imu0, const fflins_imu *imu1, double mid) { return need_interpolation(*imu0, *imu1, mid); }
void fflins_imu_interpolation(const fflins_imu *imu01, fflins_imu *imu00, fflins_imu *imu11, double mid) {
    interpolate(*imu01, *imu00, *imu11, mid);
}

fflins_ins_window *fflins_insw_create(const fflins_ins_config *cfg) {
    auto *w = new fflins_ins_window();
    w->cfg  = *cfg;
    return w;
}
void fflins_insw_destroy(fflins_ins_window *w) { delete w; }
void fflins_insw_reset(fflins_ins_window *w, const fflins_imu *imu0, const fflins_nav_state *state0) {
    w->w.clear();
    fflins_nav_state s = *state0;
    s.time             = imu0->time;
    w->w.emplace_back(*imu0, s);
}
void fflins_insw_add_imu(fflins_ins_window *w, const fflins_imu *imu) {
    if (w->w.empty()) return;
    fflins_nav_state s = w->w.back().second;
    mechanize(w->cfg, w->w.back().first, *imu, s);
    w->w.emplace_back(*imu, s);
}
int fflins_insw_size(const fflins_ins_window *w) { return (int) w->w.size(); }
int fflins_insw_back(const fflins_ins_window *w, fflins_imu *imu, fflins_nav_state *state) {
    if (w->w.empty()) return -1;
    if (imu) *imu = w->w.back().first;
    if (state) *state = w->w.back().second;
    return 0;
}
int fflins_insw_export(const fflins_ins_window *w, double *t, double *p, double *q, int cap) {
    int n = std::min((int) w->w.size(), cap);
    int first = (int) w->w.size() - n;   // the newest `cap` entries
    for (int i = 0; i < n; i++) {
        const auto &e = w->w[first + i];
        t[i] = e.first.time;
        std::memcpy(p + 3 * i, e.second.p, 24);
        std::memcpy(q + 4 * i, e.second.q, 32);
    }
    return n;
}
int fflins_insw_redo(fflins_ins_window *w, const fflins_nav_state *updated, int reserved_ins_num) {
    fflins_nav_state state = *updated;
    const size_t index     = w->index(state.time);
    if (index == 0) return -1;
    fflins_imu imu0 = w->w[index - 1].first, imu1 = w->w[index].first;
    const int need = need_interpolation(imu0, imu1, state.time);
    if (need == -1) {
        mechanize(w->cfg, imu0, imu1, state);
        w->w[index].second = state;
    } else if (need == 1) {
        state.time         = imu1.time;
        w->w[index].second = state;
    } else if (need == 2) {
        interpolate(imu1, imu0, imu1, state.time);
        mechanize(w->cfg, imu0, imu1, state);
        w->w[index].second = state;
    }
    for (size_t k = index + 1; k < w->w.size(); k++) {   // only the IMU epochs are updated (misc.cc:221-227)
        imu0 = imu1;
        imu1 = w->w[k].first;
        mechanize(w->cfg, imu0, imu1, state);
        w->w[k].second = state;
    }
    if ((int) index >= reserved_ins_num)
        for (size_t k = 0; k < index - (size_t) reserved_ins_num; k++) w->w.pop_front();
    return 0;
}
int fflins_insw_imu_series(const fflins_ins_window *w, double start, double end, fflins_imu *out, int cap) {
    const size_t is = w->index(start), ie = w->index(end);
    if (is == 0 && ie == 0) return -1;
    std::vector<fflins_imu> s;
    fflins_imu imu0 = w->w[is - 1].first, imu1 = w->w[is].first, imu{};
    int need = need_interpolation(imu0, imu1, start);
    if (need == -1) {
        s.push_back(imu0);
        s.push_back(imu1);
    } else if (need == 1) {
        s.push_back(imu1);
    } else if (need == 2) {
        interpolate(imu1, imu, imu1, start);
        s.push_back(imu);
        s.push_back(imu1);
    }
    for (size_t k = is + 1; k + 1 < ie; k++) s.push_back(w->w[k].first);
    imu0 = w->w[ie - 1].first;
    imu1 = w->w[ie].first;
    need = need_interpolation(imu0, imu1, end);
    if (need == -1) {
        s.push_back(imu0);
    } else if (need == 1) {
        s.push_back(imu0);
        s.push_back(imu1);
    } else if (need == 2) {
        s.push_back(imu0);
        interpolate(imu1, imu, imu1, end);
        s.push_back(imu);
    }
    if (s.empty()) return -1;
    s.back().time = end;   // explicit (misc.cc:337)
    if ((int) s.size() > cap) return -1;
    std::memcpy(out, s.data(), sizeof(fflins_imu) * s.size());
    return (int) s.size();
}

fflins_preint *fflins_preint_create(const fflins_imu_noise *noise, const fflins_imu *imu0, const fflins_nav_state *state) {
    auto *p       = new fflins_preint();
    p->par        = *noise;
    p->cur        = *state;
    p->start_time = p->end_time = imu0->time;
    p->buf.push_back(*imu0);
    p->gravity = {0, 0, noise->gravity};
    p->reset_state(p->cur);
    p->noise = Mat::identity(NN);   // setNoiseMatrix (preintegration_normal.cc:314-322)
    for (int k = 0; k < 3; k++) {
        p->noise(k, k)         = noise->gyr_arw * noise->gyr_arw;
        p->noise(3 + k, 3 + k) = noise->acc_vrw * noise->acc_vrw;
        p->noise(6 + k, 6 + k) = 2 * noise->gyr_bias_std * noise->gyr_bias_std / noise->corr_time;
        p->noise(9 + k, 9 + k) = 2 * noise->acc_bias_std * noise->acc_bias_std / noise->corr_time;
    }
    return p;
}
void fflins_preint_destroy(fflins_preint *p) { delete p; }
void fflins_preint_add_imu(fflins_preint *p, const fflins_imu *imu) {
    p->buf.push_back(*imu);
    p->process(p->buf.size() - 1);
}
void fflins_preint_reintegrate(fflins_preint *p, const fflins_nav_state *state) {
    p->cur = *state;
    p->reset_state(p->cur);
    for (size_t k = 1; k < p->buf.size(); k++) p->process(k);
}
void fflins_preint_current_state(const fflins_preint *p, fflins_nav_state *out) { *out = p->cur; }
void fflins_preint_delta_state(const fflins_preint *p, fflins_nav_state *out) { *out = p->delta; }
double fflins_preint_delta_time(const fflins_preint *p) { return p->delta_time; }
void fflins_preint_evaluate(fflins_preint *p, const double pose0[7], const double mix0[9], const double pose1[7],
                            const double mix1[9], double residual[15], double *J_pose0, double *J_mix0, double *J_pose1,
                            double *J_mix1) {
    preint_evaluate(*p, pose0, mix0, pose1, mix1, residual, J_pose0, J_mix0, J_pose1, J_mix1);
}
void fflins_preint_matrices(const fflins_preint *p, double *cov225, double *jac225) {
    if (cov225) std::memcpy(cov225, p->cov.a.data(), 225 * 8);
    if (jac225) std::memcpy(jac225, p->jac.a.data(), 225 * 8);
}
void fflins_preint_frame_omega(const fflins_preint *p, double lidar_frame_dt, double imu_dt, const double bg[3], double omega[3]) {
    // ff_lins.cc:474-489: mean rate over the last half LiDAR frame of the preintegration's IMU buffer
    double sum[3] = {0, 0, 0}, dt = 0;
    double start  = (p->delta_time - lidar_frame_dt * 0.5) / imu_dt;
    size_t k0     = start > 0 ? (size_t) start : 0;
    for (size_t k = k0; k < p->buf.size(); k++) {
        for (int i = 0; i < 3; i++) sum[i] += p->buf[k].dtheta[i];
        dt += p->buf[k].dt;
    }
    for (int i = 0; i < 3; i++) omega[i] = (dt > 0 ? sum[i] / dt : 0.0) - bg[i];
}

// ---- solver --------------------------------------------------------------------------------------------------------
void fflins_solver_options_default(fflins_solver_options *o) {
    o->first_iterations      = 5;
    o->second_iterations     = 10;
    o->chi2_threshold        = 3.841;
    o->estimate_extrinsic    = 1;
    o->estimate_td           = 1;
    o->window_size           = 10;
    o->min_speed_estimate_td = 1.0;
}
fflins_solver *fflins_solver_create(const fflins_solver_options *o) {
    auto *s = new fflins_solver();
    s->opt  = *o;
    return s;
}
void fflins_solver_destroy(fflins_solver *s) { delete s; }
void fflins_solver_set_extrinsic(fflins_solver *s, const double ext[7], double td) {
    std::memcpy(s->ext, ext, 56);
    s->td = td;
}
void fflins_solver_get_extrinsic(const fflins_solver *s, double ext[7], double *td) {
    std::memcpy(ext, s->ext, 56);
    if (td) *td = s->td;
}
static void state_to_blocks(const fflins_nav_state &st, double *pose, double *mix) {   // PreintegrationBase::stateToData
    std::memcpy(pose, st.p, 24);
    std::memcpy(pose + 3, st.q, 32);
    std::memcpy(mix, st.v, 24);
    std::memcpy(mix + 3, st.bg, 24);
    std::memcpy(mix + 6, st.ba, 24);
}
static void blocks_to_state(double time, const double *pose, const double *mix, fflins_nav_state &st) {   // stateFromData
    st      = fflins_nav_state{};
    st.time = time;
    std::memcpy(st.p, pose, 24);
    qst(st.q, qnormalized(qld(pose + 3)));
    std::memcpy(st.v, mix, 24);
    std::memcpy(st.bg, mix + 3, 24);
    std::memcpy(st.ba, mix + 6, 24);
}
void fflins_solver_set_first_node(fflins_solver *s, double time, const fflins_nav_state *state, const double pose_std[6],
                                  const double mix_std[9]) {
    s->nodes.clear();
    fflins_solver::Node n;
    n.time     = time;
    n.lidar_id = UINT64_MAX;
    n.uid      = s->next_uid++;
    state_to_blocks(*state, n.pose, n.mix);
    s->nodes.push_back(std::move(n));
    s->use_prior = pose_std && mix_std;
    if (s->use_prior) {
        std::memcpy(s->prior_pose, s->nodes[0].pose, 56);
        std::memcpy(s->prior_mix, s->nodes[0].mix, 72);
        std::memcpy(s->prior_pose_std, pose_std, 48);
        std::memcpy(s->prior_mix_std, mix_std, 72);
    }
    s->have_marg = false;
}
int fflins_solver_add_node(fflins_solver *s, double time, fflins_preint *preint, uint64_t lidar_id) {
    if (s->nodes.empty() || !preint) return -1;
    fflins_solver::Node n;
    n.time     = time;
    n.lidar_id = lidar_id;
    n.uid      = s->next_uid++;
    fflins_nav_state st = preint->cur;   // addNewTimeNode: current state of the preintegration (ff_lins.cc:516-518)
    st.time             = time;
    state_to_blocks(st, n.pose, n.mix);
    n.preint.reset(preint);
    s->nodes.push_back(std::move(n));
    return (int) s->nodes.size() - 1;
}
int fflins_solver_set_node_lidar(fflins_solver *s, int index, uint64_t lidar_id) {
    if (index < 0 || index >= (int) s->nodes.size()) return -1;
    s->nodes[index].lidar_id = lidar_id;
    return 0;
}
int fflins_solver_set_node_state(fflins_solver *s, int index, const fflins_nav_state *state) {
    if (index < 0 || index >= (int) s->nodes.size() || !state) return -1;
    state_to_blocks(*state, s->nodes[index].pose, s->nodes[index].mix);
    return 0;
}
int fflins_solver_num_nodes(const fflins_solver *s) { return (int) s->nodes.size(); }
int fflins_solver_get_node(const fflins_solver *s, int index, fflins_nav_state *state, uint64_t *lidar_id) {
    if (index < 0 || index >= (int) s->nodes.size()) return -1;
    if (state) blocks_to_state(s->nodes[index].time, s->nodes[index].pose, s->nodes[index].mix, *state);
    if (lidar_id) *lidar_id = s->nodes[index].lidar_id;
    return 0;
}

int fflins_solver_optimize(fflins_solver *s, fflins_lidar_eval_fn eval, fflins_lidar_chi2_fn chi2, void *user,
                           fflins_solver_summary *out) {
    if (s->nodes.size() < 2) return -1;
    const LidarLink L = lidar_link(*s);
    const bool full   = s->window_full();
    // optimizer.cc:406-435: extrinsic and td stay constant until the window is full; td also below the speed gate
    bool ext_const = !s->opt.estimate_extrinsic || !full, td_const = !s->opt.estimate_td || !full;
    if (!td_const && L.cur_node >= 0) {
        const double *v = s->nodes[L.cur_node].mix;
        if (std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) < s->opt.min_speed_estimate_td) td_const = true;
    }
    fflins_solver_summary sum{};
    double c0 = 0, c1 = 0;
    int rc = lm_solve(*s, L, eval, user, ext_const, td_const, s->opt.first_iterations, &sum.iterations[0], &sum.successful[0], &c0, &c1);
    if (rc) return rc;
    sum.cost_initial = c0;
    if (chi2 && !L.ref_ids.empty()) {   // lidarOutlierCullingByChi2 (optimizer.cc:515-548)
        const int n = (int) L.ref_ids.size();
        std::vector<double> pose_refs(7 * (size_t) n);
        std::vector<int32_t> cnt(n, 0);
        for (int i = 0; i < n; i++) std::memcpy(pose_refs.data() + 7 * i, s->nodes[L.ref_nodes[i]].pose, 56);
        rc = chi2(user, n, L.ref_ids.data(), pose_refs.data(), s->nodes[L.cur_node].pose, s->ext, s->td, s->opt.chi2_threshold, cnt.data());
        if (rc) return rc;
        for (int c : cnt) sum.outliers += c;
    }
    rc = lm_solve(*s, L, eval, user, ext_const, td_const, s->opt.second_iterations, &sum.iterations[1], &sum.successful[1], &c0, &c1);
    if (rc) return rc;
    sum.cost_final = c1;
    if (!full) {   // doReintegration (optimizer.cc:169-172): re-linearise the preintegrations at the new biases
        for (size_t k = 1; k < s->nodes.size(); k++) {
            fflins_nav_state st;
            blocks_to_state(s->nodes[k - 1].time, s->nodes[k - 1].pose, s->nodes[k - 1].mix, st);
            fflins_preint_reintegrate(s->nodes[k].preint.get(), &st);
        }
    }
    if (out) *out = sum;
    return 0;
}

int fflins_solver_marginalize(fflins_solver *s, fflins_lidar_eval_fn eval, void *user) {
    if (s->nodes.size() < 2) return -1;
    const LidarLink L = lidar_link(*s);
    // H0 / b0 over the whole local vector from the factors that touch node 0 (optimizer.cc:262-360): the previous prior,
    // the IMU factor 0 -> 1, the initial priors, and the LiDAR factors whose reference is node 0 (Huber(1.0), marg index {0})
    Problem P;
    const int n = s->n_local();
    P.H = Mat(n, n);
    P.b.assign(n, 0.0);
    if (s->have_marg) add_marg_prior(*s, P);
    add_imu_factor(*s, 0, P);
    if (s->use_prior) {
        add_initial_priors(*s, P);
        s->use_prior = false;
    }
    if (L.cur_node > 0 && !L.ref_nodes.empty() && L.ref_nodes[0] == 0) {
        std::vector<int> only{0};
        int rc = add_lidar(*s, L, eval, user, 1u, P, &only);
        if (rc) return rc;
    }
    // Schur complement of node 0 (15 leading columns), eigen-decomposition with the EPS cut (marginalization_info.cc:93-131)
    const int m = 15, r = n - m;
    Mat Hmm(m, m);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Hmm(i, j) = 0.5 * (P.H(i, j) + P.H(j, i));
    std::vector<double> w;
    Mat V;
    sym_eig(Hmm, w, V);
    Mat Hmm_inv(m, m);
    for (int k = 0; k < m; k++) {
        if (!(w[k] > kMargEps)) continue;
        for (int i = 0; i < m; i++)
            for (int j = 0; j < m; j++) Hmm_inv(i, j) += V(i, k) * V(j, k) / w[k];
    }
    Mat Hrm(r, m), Hmr(m, r), Hrr(r, r);
    std::vector<double> bm(m), br(r);
    for (int i = 0; i < m; i++) bm[i] = P.b[i];
    for (int i = 0; i < r; i++) {
        br[i] = P.b[m + i];
        for (int j = 0; j < m; j++) {
            Hrm(i, j) = P.H(m + i, j);
            Hmr(j, i) = P.H(j, m + i);
        }
        for (int j = 0; j < r; j++) Hrr(i, j) = P.H(m + i, m + j);
    }
    const Mat T = Hrm * Hmm_inv;
    const Mat Hp_corr = T * Hmr;
    Mat Hp(r, r);
    std::vector<double> bp(r);
    for (int i = 0; i < r; i++) {
        double sb = 0;
        for (int j = 0; j < m; j++) sb += T(i, j) * bm[j];
        bp[i] = br[i] - sb;
        for (int j = 0; j < r; j++) Hp(i, j) = Hrr(i, j) - Hp_corr(i, j);
    }
    // only the blocks the marginal actually constrains are kept (zero rows / columns carry nothing)
    std::vector<fflins_solver::PBlock> blocks;
    std::vector<int> keep;   // columns of Hp
    auto touched = [&](int off, int size) {
        for (int i = 0; i < size; i++)
            for (int j = 0; j < r; j++)
                if (Hp(off + i, j) != 0.0) return true;
        return false;
    };
    auto push = [&](int off_window, int size, int kind, uint64_t uid, const double *x, int xlen) {
        const int off = off_window - m;
        if (!touched(off, size)) return;
        fflins_solver::PBlock pb{};
        pb.uid  = uid;
        pb.kind = kind;
        pb.size = size;
        std::memcpy(pb.x0, x, sizeof(double) * xlen);
        blocks.push_back(pb);
        for (int i = 0; i < size; i++) keep.push_back(off + i);
    };
    for (int k = 1; k < (int) s->nodes.size(); k++) {
        push(s->off_pose(k), 6, 0, s->nodes[k].uid, s->nodes[k].pose, 7);
        push(s->off_mix(k), 9, 1, s->nodes[k].uid, s->nodes[k].mix, 9);
    }
    push(s->off_ext(), 6, 2, 0, s->ext, 7);
    push(s->off_td(), 1, 3, 0, &s->td, 1);
    const int kc = (int) keep.size();
    Mat Hk(kc, kc);
    std::vector<double> bk(kc);
    for (int i = 0; i < kc; i++) {
        bk[i] = bp[keep[i]];
        for (int j = 0; j < kc; j++) Hk(i, j) = 0.5 * (Hp(keep[i], keep[j]) + Hp(keep[j], keep[i]));
    }
    sym_eig(Hk, w, V);
    // J0 = S^{1/2} V^T, e0 = -S^{-1/2} V^T bp over the eigen-directions above EPS (linearization)
    std::vector<int> dirs;
    for (int k = 0; k < kc; k++)
        if (w[k] > kMargEps) dirs.push_back(k);
    s->J0 = Mat((int) dirs.size(), kc);
    s->e0.assign(dirs.size(), 0.0);
    for (size_t a = 0; a < dirs.size(); a++) {
        const int k = dirs[a];
        const double sq = std::sqrt(w[k]);
        double vb = 0;
        for (int j = 0; j < kc; j++) {
            s->J0((int) a, j) = sq * V(j, k);
            vb += V(j, k) * bk[j];
        }
        s->e0[a] = -vb / sq;
    }
    s->pblocks   = blocks;
    s->have_marg = !dirs.empty();
    s->nodes.pop_front();
    if (!s->nodes.empty()) s->nodes.front().preint.reset();
    return 0;
}

int fflins_solver_prior(const fflins_solver *s, int *rows, int *cols, double *J0, double *e0, int cap) {
    if (!s->have_marg) {
        if (rows) *rows = 0;
        if (cols) *cols = 0;
        return 0;
    }
    if (rows) *rows = s->J0.r;
    if (cols) *cols = s->J0.c;
    if (J0 || e0) {
        if (cap < s->J0.r * s->J0.c) return -1;
        if (J0) std::memcpy(J0, s->J0.a.data(), sizeof(double) * s->J0.a.size());
        if (e0) std::memcpy(e0, s->e0.data(), sizeof(double) * s->e0.size());
    }
    return 0;
}

// ---- converters and writers ----------------------------------------------------------------------------------------
namespace {
struct Rec32 {   // PointXYZIT (lidar/lidar.h:30-38)
    float x, y, z, pad0;
    float intensity, pad1;
    double time;
};
static_assert(sizeof(Rec32) == 32, "record layout");
inline double to_gps(double unixs, int on) {   // GpsTime::unix2gps (common/gpstime.h:38-43), second of week
    if (!on) return unixs;
    const double seconds = unixs + 18 /* GPS_LEAP_SECOND */ - 315964800;
    const double week    = std::floor(seconds / 604800);
    return seconds - week * 604800;
}
} // namespace

int fflins_convert_livox(const fflins_livox_point *pts, int n, uint64_t timebase_ns, int scan_line, double min_range,
                         double max_range, int to_gps_time, void *out32, int cap, double *start, double *end) {
    Rec32 *out = static_cast<Rec32 *>(out32);
    const double lo = min_range * min_range, hi = max_range * max_range;
    double s = DBL_MAX, e = 0;
    int m = 0;
    for (int k = 0; k < n; k++) {
        const auto &raw = pts[k];
        if (!((raw.line < scan_line) && ((raw.tag & 0x30) == 0x00 || (raw.tag & 0x30) == 0x10))) continue;   // echo 0 / 1
        const double t = to_gps((double) (timebase_ns + raw.offset_time) * 1.0e-9, to_gps_time);
        s = std::min(s, t);
        e = std::max(e, t);
        const float d2f = raw.x * raw.x + raw.y * raw.y + raw.z * raw.z;   // squaredNorm of the float vector
        const double d2 = d2f;
        if (d2 > lo && d2 < hi) {
            if (m >= cap) return -1;
            out[m++] = Rec32{raw.x, raw.y, raw.z, 1.0f, (float) raw.reflectivity, 0.0f, t};
        }
    }
    if (start) *start = s;
    if (end) *end = e;
    return m;
}

int fflins_convert_generic(const float *xyzi, const double *time, int n, double stamp, int time_mode, double min_range,
                           double max_range, int to_gps_time, void *out32, int cap, double *start, double *end) {
    Rec32 *out = static_cast<Rec32 *>(out32);
    const double lo = min_range * min_range, hi = max_range * max_range;
    const double base = to_gps(stamp, to_gps_time);
    double s = DBL_MAX, e = 0;
    int m = 0;
    for (int k = 0; k < n; k++) {
        const float *p = xyzi + 4 * (size_t) k;
        double t;
        if (time_mode == 0) t = base + time[k];               // Velodyne: seconds after the stamp
        else if (time_mode == 1) t = base + time[k] * 1.0e-9; // Ouster: nanoseconds after the stamp
        else t = to_gps(time[k], to_gps_time);                // Hesai: absolute
        s = std::min(s, t);
        e = std::max(e, t);
        const float d2f = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        const double d2 = d2f;
        if (d2 > lo && d2 < hi) {
            if (m >= cap) return -1;
            out[m++] = Rec32{p[0], p[1], p[2], 1.0f, p[3], 0.0f, t};
        }
    }
    if (start) *start = s;
    if (end) *end = e;
    return m;
}

static int format_line(const double *v, int n, char *buf, int cap) {   // FileSaver::dump_ TEXT (filesaver.cc:52-63)
    int off = 0;
    for (int k = 0; k < n; k++) {
        int w = std::snprintf(buf + off, cap > off ? (size_t) (cap - off) : 0, "%-15.9lf ", v[k]);
        if (w < 0) return -1;
        off += w;
    }
    if (off + 1 >= cap) return -1;
    buf[off++] = '\n';
    buf[off]   = 0;
    return off;
}
int fflins_format_trajectory(const fflins_nav_state *s, char *buf, int cap) {
    const double v[8] = {s->time, s->p[0], s->p[1], s->p[2], s->q[0], s->q[1], s->q[2], s->q[3]};
    return format_line(v, 8, buf, cap);
}
int fflins_format_imu_error(const fflins_nav_state *s, char *buf, int cap) {   // writeNavResult: deg/h and mGal
    const double R2D = 180.0 / M_PI;
    const double v[8] = {s->time, s->bg[0] * R2D * 3600, s->bg[1] * R2D * 3600, s->bg[2] * R2D * 3600,
                         s->ba[0] * 1e5, s->ba[1] * 1e5, s->ba[2] * 1e5, 0.0};
    return format_line(v, 8, buf, cap);
}
int fflins_format_statistics(const double *values, int n, char *buf, int cap) { return format_line(values, n, buf, cap); }

} // extern "C"
