/*
 * oracle.cc — CPU restatement of the FF-LINS LiDAR hot path. TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may load it and for the parity status: pinned against the reference's own
 * lidar_factor.cc / misc.cc / rotation.cc / pointcloud.cc / lidar_map.cc / ikd-Tree compiled into oracle/_ref/;
 * "parity unpinned" only for what lives inside Eigen / PCL / Ceres themselves).
 *
 * Written from the reference's behaviour, not its text: each function cites the upstream
 * file:line it follows (paths relative to the FF-LINS tree). Build with
 *   g++ -O2 -ffp-contract=off -fopenmp      (oracle/Makefile)
 * -ffp-contract=off mirrors the reference's no-FMA x86-64 build (CMakeLists.txt:7-10): every
 * product and sum rounds separately, in the left-to-right order written here.
 *
 * Decisions where the reference is implementation-defined:
 *   - kNN ties: (d2 fp32, map index) lexicographic;
 *   - VoxelGrid intra-voxel order: ascending original index, sequential fp32 adds;
 *   - small fixed-size products: row-dot evaluated ((a0*b0 + a1*b1) + a2*b2).
 */
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 {
    double x, y, z;
};
struct M3 {
    double m[9];
}; // row-major
struct Quat {
    double x, y, z, w;
};

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 mv(const M3 &A, V3 v) {
    return {(A.m[0] * v.x + A.m[1] * v.y) + A.m[2] * v.z, (A.m[3] * v.x + A.m[4] * v.y) + A.m[5] * v.z,
            (A.m[6] * v.x + A.m[7] * v.y) + A.m[8] * v.z};
}
inline M3 mm(const M3 &A, const M3 &B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = (A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j]) + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
inline M3 tr(const M3 &A) { return {{A.m[0], A.m[3], A.m[6], A.m[1], A.m[4], A.m[7], A.m[2], A.m[5], A.m[8]}}; }
inline M3 ident() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 madd(const M3 &A, const M3 &B) {
    M3 C;
    for (int i = 0; i < 9; i++) C.m[i] = A.m[i] + B.m[i];
    return C;
}
inline M3 msub(const M3 &A, const M3 &B) {
    M3 C;
    for (int i = 0; i < 9; i++) C.m[i] = A.m[i] - B.m[i];
    return C;
}
/* row-vector times matrix: (v^T A) */
inline V3 vm(V3 v, const M3 &A) {
    return {(v.x * A.m[0] + v.y * A.m[3]) + v.z * A.m[6], (v.x * A.m[1] + v.y * A.m[4]) + v.z * A.m[7],
            (v.x * A.m[2] + v.y * A.m[5]) + v.z * A.m[8]};
}
/* common/rotation.cc:82-86 */
inline M3 skew(V3 v) { return {{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }

/* Eigen Quaternion::toRotationMatrix, no normalisation. */
inline M3 quat_to_R(Quat q) {
    double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 R;
    R.m[0] = 1.0 - (tyy + tzz);
    R.m[1] = txy - twz;
    R.m[2] = txz + twy;
    R.m[3] = txy + twz;
    R.m[4] = 1.0 - (txx + tzz);
    R.m[5] = tyz - twx;
    R.m[6] = txz - twy;
    R.m[7] = tyz + twx;
    R.m[8] = 1.0 - (txx + tyy);
    return R;
}
/* Hamilton product, Eigen's generic formula. */
inline Quat qmul(Quat a, Quat b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
inline double qsqnorm(Quat q) { return ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w; }
/* Eigen Quaternion::inverse: conjugate / squaredNorm. */
inline Quat qinv(Quat q) {
    double n2 = qsqnorm(q);
    return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}
inline Quat qnormalized(Quat q) {
    double n = std::sqrt(qsqnorm(q));
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}

/* Rotation::quaternion2vector (common/rotation.cc:67-70) through Eigen AngleAxisd(q). */
inline V3 quat_to_rotvec(Quat q) {
    double n = std::sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);
    if (n < DBL_EPSILON) { /* Eigen falls back to stableNorm for tiny vectors */
        double m = std::max(std::fabs(q.x), std::max(std::fabs(q.y), std::fabs(q.z)));
        if (m == 0.0) {
            n = 0.0;
        } else {
            double ax = q.x / m, ay = q.y / m, az = q.z / m;
            n = m * std::sqrt((ax * ax + ay * ay) + az * az);
        }
    }
    if (n != 0.0) {
        double angle = 2.0 * std::atan2(n, std::fabs(q.w));
        if (q.w < 0.0) n = -n;
        V3 axis = {q.x / n, q.y / n, q.z / n};
        return {angle * axis.x, angle * axis.y, angle * axis.z};
    }
    return {0.0, 0.0, 0.0}; /* angle 0 * axis (1,0,0) */
}
/* Rotation::rotvec2quaternion (common/rotation.cc:61-65). */
inline Quat rotvec_to_quat(V3 rv) {
    double z     = (rv.x * rv.x + rv.y * rv.y) + rv.z * rv.z;
    double angle = std::sqrt(z);
    V3 vec       = rv;
    if (z > 0.0) { /* Eigen normalized(): unchanged when squaredNorm is 0 */
        double s = std::sqrt(z);
        vec      = {rv.x / s, rv.y / s, rv.z / s};
    }
    double ha = 0.5 * angle;
    double s = std::sin(ha), c = std::cos(ha);
    return {s * vec.x, s * vec.y, s * vec.z, c};
}

struct Pose {
    M3 R;
    V3 t;
};

/* MISC::stateToPoseWithExtrinsic (common/misc.cc:99-105). */
inline Pose state_to_pose(V3 p, Quat q, const Pose &ext) {
    Pose out;
    M3 R  = quat_to_R(q);
    out.t = add(p, mv(R, ext.t));
    out.R = mm(R, ext.R);
    return out;
}

/* MISC::getInsWindowIndex (common/misc.cc:27-62), literal including the 17-iteration cap. */
int ins_window_index(const double *t, int w, double time) {
    if (w <= 0 || t[0] > time || t[w - 1] <= time) return 0;
    size_t index = 0, sta = 0, end = (size_t) w;
    int counts = 0;
    while (true) {
        size_t mid    = (sta + end) / 2;
        double first  = t[mid - 1];
        double second = t[mid];
        if (first <= time && time < second) {
            index = mid;
            break;
        } else if (first > time) {
            end = mid;
        } else if (second <= time) {
            sta = mid;
        }
        if (counts++ > 15) break;
    }
    return (int) index;
}

/* MISC::statePoseInterpolation (common/misc.cc:82-97). */
inline void state_interp(double t0, V3 p0, Quat q0, double t1, V3 p1, Quat q1, double mid, V3 &p, Quat &q) {
    V3 dp    = sub(p1, p0);
    Quat dq  = qmul(qinv(q1), q0);
    V3 rvec  = quat_to_rotvec(dq);
    double s = (mid - t0) / (t1 - t0);
    rvec     = scale(rvec, s);
    dq       = rotvec_to_quat(rvec);
    p        = add(p0, scale(dp, s));
    q        = qnormalized(qmul(q0, qinv(dq)));
}

inline V3 ld3(const double *a) { return {a[0], a[1], a[2]}; }
inline Quat ldq(const double *a) { return {a[0], a[1], a[2], a[3]}; }
inline M3 ldm(const double *a) {
    M3 R;
    std::memcpy(R.m, a, sizeof(R.m));
    return R;
}

/* MISC::getPoseFromInsWindow (common/misc.cc:64-80). */
bool pose_from_window(const double *it, const double *ip, const double *iq, int w, const Pose &ext, double time,
                      Pose &pose) {
    int index = ins_window_index(it, w, time);
    if (index > 0) {
        V3 p;
        Quat q;
        state_interp(it[index - 1], ld3(ip + 3 * (index - 1)), ldq(iq + 4 * (index - 1)), it[index],
                     ld3(ip + 3 * index), ldq(iq + 4 * index), time, p, q);
        pose = state_to_pose(p, q, ext);
        return true;
    }
    pose = state_to_pose(ld3(ip + 3 * (w - 1)), ldq(iq + 4 * (w - 1)), ext);
    return false;
}

/* PointCloudCommon::pointProjection (lidar/pointcloud.cc:53-59): fp64 math, fp32 store. */
inline void project_point(const Pose &T, const float *src, float *dst) {
    V3 p  = {(double) src[0], (double) src[1], (double) src[2]};
    V3 r  = add(mv(T.R, p), T.t);
    float o0 = (float) r.x, o1 = (float) r.y, o2 = (float) r.z, o3 = src[3];
    dst[0] = o0;
    dst[1] = o1;
    dst[2] = o2;
    dst[3] = o3;
}
/* PointCloudCommon::relativeProjection pose part (lidar/pointcloud.cc:30-36). */
inline Pose relative_pose(const Pose &p0, const Pose &p1) {
    Pose r;
    M3 R0t = tr(p0.R);
    r.R    = mm(R0t, p1.R);
    r.t    = mv(R0t, sub(p1.t, p0.t));
    return r;
}

/* ---- exact static kd-tree used only as the *timed* CPU association baseline; results are
 * identical to brute force ((d2, index) order), verified in tests. Plays the role of
 * ikd_Tree::build / nearestSearch (ikd_Tree.cc:346-391). */
struct KdTree {
    struct Node {
        int lo, hi;      /* index range into perm (leaf) */
        int left, right; /* children or -1 */
        int axis;
        float split;
        float bmin[3], bmax[3];
    };
    const float *pts = nullptr;
    std::vector<int> perm;
    std::vector<Node> nodes;

    int build_rec(int lo, int hi) {
        Node nd;
        nd.lo = lo;
        nd.hi = hi;
        nd.left = nd.right = -1;
        for (int a = 0; a < 3; a++) {
            nd.bmin[a] = FLT_MAX;
            nd.bmax[a] = -FLT_MAX;
        }
        for (int i = lo; i < hi; i++)
            for (int a = 0; a < 3; a++) {
                float v    = pts[4 * perm[i] + a];
                nd.bmin[a] = std::min(nd.bmin[a], v);
                nd.bmax[a] = std::max(nd.bmax[a], v);
            }
        int id = (int) nodes.size();
        nodes.push_back(nd);
        if (hi - lo > 8) {
            int axis = 0;
            float ext = -1.f;
            for (int a = 0; a < 3; a++) {
                float e = nd.bmax[a] - nd.bmin[a];
                if (e > ext) {
                    ext  = e;
                    axis = a;
                }
            }
            int mid = (lo + hi) / 2;
            std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                             [&](int a, int b) { return pts[4 * a + axis] < pts[4 * b + axis]; });
            int l            = build_rec(lo, mid);
            int r            = build_rec(mid, hi);
            nodes[id].left   = l;
            nodes[id].right  = r;
            nodes[id].axis   = axis;
            nodes[id].split  = pts[4 * perm[mid] + axis];
        }
        return id;
    }
    void build(const float *p, int m) {
        pts = p;
        perm.resize(m);
        std::iota(perm.begin(), perm.end(), 0);
        nodes.clear();
        nodes.reserve(m / 4 + 8);
        if (m > 0) build_rec(0, m);
    }
    static inline float box_d2(const Node &n, const float *q) {
        float d = 0.f;
        for (int a = 0; a < 3; a++) {
            if (q[a] < n.bmin[a]) d += (q[a] - n.bmin[a]) * (q[a] - n.bmin[a]);
            if (q[a] > n.bmax[a]) d += (q[a] - n.bmax[a]) * (q[a] - n.bmax[a]);
        }
        return d;
    }
    struct Top5 {
        float d2[5];
        int idx[5];
        int n = 0;
        inline bool better(float d, int i, int k) const { return d < d2[k] || (d == d2[k] && i < idx[k]); }
        inline void push(float d, int i) {
            if (n == 5 && !better(d, i, 4)) return;
            int k = (n < 5) ? n++ : 4;
            while (k > 0 && better(d, i, k - 1)) {
                d2[k]  = d2[k - 1];
                idx[k] = idx[k - 1];
                k--;
            }
            d2[k]  = d;
            idx[k] = i;
        }
    };
    void search_rec(int id, const float *q, float maxd2, Top5 &top) const {
        const Node &n = nodes[id];
        float bd      = box_d2(n, q);
        /* conservative pruning: box distance is a lower bound only up to rounding, keep 1e-4 slack */
        float lim = (top.n == 5) ? top.d2[4] : maxd2;
        if (bd * 0.9999f > lim) return;
        if (n.left < 0) {
            for (int i = n.lo; i < n.hi; i++) {
                int pi    = perm[i];
                float dx = q[0] - pts[4 * pi], dy = q[1] - pts[4 * pi + 1], dz = q[2] - pts[4 * pi + 2];
                float d  = dx * dx + dy * dy + dz * dz;
                if (d <= maxd2) top.push(d, pi);
            }
            return;
        }
        int first = n.left, second = n.right;
        if (box_d2(nodes[n.right], q) < box_d2(nodes[n.left], q)) std::swap(first, second);
        search_rec(first, q, maxd2, top);
        search_rec(second, q, maxd2, top);
    }
};

/* KD_TREE::nearestSearch semantics (ikd_Tree.cc:897-924 accept rule `dist <= max_dist^2`,
 * :1427-1431 fp32 left-to-right d2, results ascending :385-389). Truth = brute force. */
int knn5_brute(const float *map, int m, const float *q, double max_dist, int32_t *idx, float *d2) {
    double max_sqr = max_dist * max_dist;
    KdTree::Top5 top;
    for (int i = 0; i < m; i++) {
        const float *b = map + 4 * i;
        float d        = (q[0] - b[0]) * (q[0] - b[0]) + (q[1] - b[1]) * (q[1] - b[1]) + (q[2] - b[2]) * (q[2] - b[2]);
        if ((double) d <= max_sqr) top.push(d, i);
    }
    for (int k = 0; k < top.n; k++) {
        idx[k] = top.idx[k];
        d2[k]  = top.d2[k];
    }
    return top.n;
}

/* Eigen ColPivHouseholderQR<Matrix<double,5,3>>::compute + solve,
 * restated: column pivoting on the largest remaining column norm with the LAPACK
 * down-dating rule, Householder reflections, solve on nonzero_pivots columns. */
void colpiv_qr_solve_5x3(double A[5][3], const double bvec[5], double xout[3]) {
    const int rows = 5, cols = 3;
    double hc[3];
    int trans[3];
    double updated[3], direct[3];
    for (int k = 0; k < cols; k++) {
        double s = 0;
        for (int i = 0; i < rows; i++) s += A[i][k] * A[i][k];
        updated[k] = direct[k] = std::sqrt(s);
    }
    double maxn = std::max(updated[0], std::max(updated[1], updated[2]));
    double th   = maxn * DBL_EPSILON / (double) rows;
    double threshold_helper  = th * th;
    double downdate_thr      = std::sqrt(DBL_EPSILON);
    int nonzero_pivots       = cols;
    for (int k = 0; k < cols; k++) {
        int big = k;
        for (int j = k + 1; j < cols; j++)
            if (updated[j] > updated[big]) big = j;
        double big_sq = updated[big] * updated[big];
        if (nonzero_pivots == cols && big_sq < threshold_helper * (double) (rows - k)) nonzero_pivots = k;
        trans[k] = big;
        if (k != big) {
            for (int i = 0; i < rows; i++) std::swap(A[i][k], A[i][big]);
            std::swap(updated[k], updated[big]);
            std::swap(direct[k], direct[big]);
        }
        /* makeHouseholderInPlace on A[k..4][k] */
        double tail_sq = 0;
        for (int i = k + 1; i < rows; i++) tail_sq += A[i][k] * A[i][k];
        double c0 = A[k][k], beta, tau;
        if (tail_sq <= DBL_MIN) {
            tau  = 0;
            beta = c0;
            for (int i = k + 1; i < rows; i++) A[i][k] = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail_sq);
            if (c0 >= 0) beta = -beta;
            for (int i = k + 1; i < rows; i++) A[i][k] = A[i][k] / (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hc[k]   = tau;
        A[k][k] = beta;
        /* applyHouseholderOnTheLeft to the trailing columns */
        if (tau != 0) {
            for (int j = k + 1; j < cols; j++) {
                double tmp = 0;
                for (int i = k + 1; i < rows; i++) tmp += A[i][k] * A[i][j];
                tmp += A[k][j];
                A[k][j] -= tau * tmp;
                for (int i = k + 1; i < rows; i++) A[i][j] -= tau * A[i][k] * tmp;
            }
        }
        for (int j = k + 1; j < cols; j++) {
            if (updated[j] != 0) {
                double temp  = std::fabs(A[k][j]) / updated[j];
                temp         = (1.0 + temp) * (1.0 - temp);
                temp         = temp < 0 ? 0 : temp;
                double ratio = updated[j] / direct[j];
                double temp2 = temp * (ratio * ratio);
                if (temp2 <= downdate_thr) {
                    double s = 0;
                    for (int i = k + 1; i < rows; i++) s += A[i][j] * A[i][j];
                    direct[j]  = std::sqrt(s);
                    updated[j] = direct[j];
                } else {
                    updated[j] *= std::sqrt(temp);
                }
            }
        }
    }
    /* solve */
    double c[5];
    for (int i = 0; i < rows; i++) c[i] = bvec[i];
    double x[3] = {0, 0, 0};
    if (nonzero_pivots > 0) {
        for (int k = 0; k < nonzero_pivots; k++) {
            double tau = hc[k];
            if (tau != 0) {
                double tmp = 0;
                for (int i = k + 1; i < rows; i++) tmp += A[i][k] * c[i];
                tmp += c[k];
                c[k] -= tau * tmp;
                for (int i = k + 1; i < rows; i++) c[i] -= tau * A[i][k] * tmp;
            }
        }
        for (int i = nonzero_pivots - 1; i >= 0; i--) {
            double s = 0;
            for (int j = i + 1; j < nonzero_pivots; j++) s += A[i][j] * c[j];
            c[i] = (c[i] - s) / A[i][i];
        }
        for (int i = 0; i < nonzero_pivots; i++) x[i] = c[i];
    }
    /* undo the column transpositions: dst = P * x */
    for (int k = cols - 1; k >= 0; k--) std::swap(x[k], x[trans[k]]);
    xout[0] = x[0];
    xout[1] = x[1];
    xout[2] = x[2];
}

/* PointCloudCommon::planeEstimation (lidar/pointcloud.cc:61-93). */
bool plane_estimation(const float *pts, double threshold, double *unit, double *norm_inverse, double *distance) {
    double A[5][3], b[5];
    for (int k = 0; k < 5; k++) {
        A[k][0] = (double) pts[4 * k];
        A[k][1] = (double) pts[4 * k + 1];
        A[k][2] = (double) pts[4 * k + 2];
        b[k]    = -1.0;
    }
    double n[3];
    colpiv_qr_solve_5x3(A, b, n);
    double norm   = std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    *norm_inverse = 1.0 / norm;
    unit[0]       = n[0] * *norm_inverse;
    unit[1]       = n[1] * *norm_inverse;
    unit[2]       = n[2] * *norm_inverse;
    bool ok       = true;
    for (int k = 0; k < 5; k++) distance[k] = 0;
    for (int k = 0; k < 5; k++) {
        double dis =
            ((unit[0] * (double) pts[4 * k] + unit[1] * (double) pts[4 * k + 1]) + unit[2] * (double) pts[4 * k + 2]) +
            *norm_inverse;
        distance[k] = dis;
        if (std::fabs(dis) > threshold) {
            ok = false;
            break; /* reference returns at the first violation (:88-90) */
        }
    }
    return ok;
}

/* Ceres HuberLoss(a)::Evaluate. */
inline void huber(double a, double s, double rho[3]) {
    double b = a * a;
    if (s > b) {
        double r = std::sqrt(s);
        rho[0]   = 2.0 * a * r - b;
        rho[1]   = std::max(DBL_MIN, a / r);
        rho[2]   = -rho[1] / (2.0 * s);
    } else {
        rho[0] = s;
        rho[1] = 1.0;
        rho[2] = 0.0;
    }
}

struct FactorConst {
    V3 p, n;
    double d, sqrt_info;
    V3 v0, w0, v1, w1;
    double td0, td1;
};

/* LidarFactor::Evaluate (factors/lidar_factor.cc:41-118), expression order as written there. */
void lidar_factor(const FactorConst &f, const double *const *par, double *res, double **jac) {
    V3 t_w_b0   = ld3(par[0]);
    Quat q_w_b0 = {par[0][3], par[0][4], par[0][5], par[0][6]};
    V3 t_w_b1   = ld3(par[1]);
    Quat q_w_b1 = {par[1][3], par[1][4], par[1][5], par[1][6]};
    V3 t_b_l    = ld3(par[2]);
    Quat q_b_l  = {par[2][3], par[2][4], par[2][5], par[2][6]};
    double td   = par[3][0];

    double dt0  = td - f.td0;
    M3 R_w_b0   = quat_to_R(q_w_b0);
    M3 R_b0_bb0 = madd(ident(), skew(scale(f.w0, dt0)));
    M3 R_w_bb0  = mm(R_w_b0, R_b0_bb0);
    V3 t_w_bb0  = add(t_w_b0, scale(f.v0, dt0));
    double dt1  = td - f.td1;
    M3 R_w_b1   = quat_to_R(q_w_b1);
    M3 R_b1_bb1 = madd(ident(), skew(scale(f.w1, dt1)));
    M3 R_w_bb1  = mm(R_w_b1, R_b1_bb1);
    V3 t_w_bb1  = add(t_w_b1, scale(f.v1, dt1));

    M3 R_b_l = quat_to_R(q_b_l);
    V3 pb1   = add(mv(R_b_l, f.p), t_b_l);
    V3 pw    = add(mv(R_w_bb1, pb1), t_w_bb1);
    V3 pb0   = mv(tr(R_w_bb0), sub(pw, t_w_bb0));
    V3 pl0   = mv(tr(R_b_l), sub(pb0, t_b_l));

    V3 sn  = scale(f.n, f.sqrt_info); /* sqrt_info * n^T, shared by the Jacobians */
    res[0] = f.sqrt_info * (dot(f.n, pl0) + f.d);

    if (!jac) return;
    V3 sqrt_common = vm(sn, tr(R_b_l)); /* s n^T R_bl^T */
    if (jac[0]) {
        double *J = jac[0];
        for (int i = 0; i < 7; i++) J[i] = 0;
        V3 a  = vm(sqrt_common, tr(R_w_bb0));
        J[0]  = -a.x;
        J[1]  = -a.y;
        J[2]  = -a.z;
        V3 u  = mv(tr(R_w_b0), sub(pw, t_w_bb0));
        V3 bb = vm(vm(sqrt_common, tr(R_b0_bb0)), skew(u));
        J[3]  = bb.x;
        J[4]  = bb.y;
        J[5]  = bb.z;
    }
    if (jac[1]) {
        double *J = jac[1];
        for (int i = 0; i < 7; i++) J[i] = 0;
        V3 a  = vm(sqrt_common, tr(R_w_bb0));
        J[0]  = a.x;
        J[1]  = a.y;
        J[2]  = a.z;
        V3 bb = vm(vm(vm(scale(sqrt_common, -1.0), tr(R_w_bb0)), R_w_b1), skew(mv(R_b1_bb1, pb1)));
        J[3]  = bb.x;
        J[4]  = bb.y;
        J[5]  = bb.z;
    }
    if (jac[2]) {
        double *J = jac[2];
        for (int i = 0; i < 7; i++) J[i] = 0;
        M3 common = mm(mm(tr(R_b_l), tr(R_w_bb0)), R_w_bb1);
        V3 a      = vm(sn, msub(common, tr(R_b_l)));
        J[0]      = a.x;
        J[1]      = a.y;
        J[2]      = a.z;
        M3 inner  = msub(skew(mv(tr(R_b_l), sub(pb0, t_b_l))), mm(mm(common, R_b_l), skew(f.p)));
        V3 bb     = vm(sn, inner);
        J[3]      = bb.x;
        J[4]      = bb.y;
        J[5]      = bb.z;
    }
    if (jac[3]) {
        V3 term0 = mv(tr(mm(R_w_b0, skew(f.w0))), sub(pw, t_w_bb0));
        /* (R_w_b1 * skew(w1) * pb1 + v1 - v0): left-to-right => ((A + v1) - v0) */
        V3 inner = sub(add(mv(mm(R_w_b1, skew(f.w1)), pb1), f.v1), f.v0);
        V3 term1 = mv(tr(R_w_bb0), inner);
        jac[3][0] = dot(sqrt_common, add(term0, term1));
    }
}

inline FactorConst make_const(const float *p, const double *n, double d, double std, const double *motion) {
    FactorConst f;
    f.p         = {(double) p[0], (double) p[1], (double) p[2]};
    f.n         = ld3(n);
    f.d         = d;
    f.sqrt_info = 1.0 / std;
    f.v0        = ld3(motion);
    f.w0        = ld3(motion + 3);
    f.td0       = motion[6];
    f.v1        = ld3(motion + 7);
    f.w1        = ld3(motion + 10);
    f.td1       = motion[13];
    return f;
}

/* ResidualBlockInfo::Evaluate loss part (factors/residual_block_info.cc:49-77), 1-dim residual. */
double huber_correct(double delta, double *r, double *J22) {
    double rho[3];
    double sq = r[0] * r[0];
    huber(delta, sq, rho);
    double sqrt_rho1 = std::sqrt(rho[1]);
    double residual_scaling, alpha_sq_norm;
    if (sq == 0.0 || rho[2] <= 0.0) {
        residual_scaling = sqrt_rho1;
        alpha_sq_norm    = 0.0;
    } else {
        double D         = 1.0 + 2.0 * sq * rho[2] / rho[1];
        double alpha     = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm    = alpha / sq;
    }
    if (J22) {
        for (int i = 0; i < 22; i++) {
            double rtJ = r[0] * J22[i];
            J22[i]     = sqrt_rho1 * (J22[i] - alpha_sq_norm * r[0] * rtJ);
        }
    }
    r[0] *= residual_scaling;
    return rho[0];
}

const int LOCAL_OF_J[19] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 21};

} // namespace

extern "C" {

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_ins_window_index(const double *ins_t, int w, double time) { return ins_window_index(ins_t, w, time); }

int orc_pose_from_ins_window(const double *ins_t, const double *ins_p, const double *ins_q, int w, const double *R_bl,
                             const double *t_bl, double time, double *R, double *t) {
    Pose ext = {ldm(R_bl), ld3(t_bl)}, pose;
    bool ok  = pose_from_window(ins_t, ins_p, ins_q, w, ext, time, pose);
    std::memcpy(R, pose.R.m, sizeof(pose.R.m));
    t[0] = pose.t.x;
    t[1] = pose.t.y;
    t[2] = pose.t.z;
    return ok ? 1 : 0;
}

void orc_state_to_pose(const double *p, const double *q, const double *R_bl, const double *t_bl, double *R,
                       double *t) {
    Pose ext  = {ldm(R_bl), ld3(t_bl)};
    Pose pose = state_to_pose(ld3(p), ldq(q), ext);
    std::memcpy(R, pose.R.m, sizeof(pose.R.m));
    t[0] = pose.t.x;
    t[1] = pose.t.y;
    t[2] = pose.t.z;
}

/* LidarMap::lidarFrameProcessing undistortion loop (lidar/lidar_map.cc:223-243). */
int orc_undistort(const float *xyzi, const double *time, int n, const double *ins_t, const double *ins_p,
                  const double *ins_q, int w, const double *R_bl, const double *t_bl, double stamp, double td,
                  float *out_xyzi, double *pose_R, double *pose_t, int threads) {
    Pose ext = {ldm(R_bl), ld3(t_bl)}, lidar_pose;
    pose_from_window(ins_t, ins_p, ins_q, w, ext, stamp, lidar_pose);
    std::memcpy(pose_R, lidar_pose.R.m, sizeof(lidar_pose.R.m));
    pose_t[0]     = lidar_pose.t.x;
    pose_t[1]     = lidar_pose.t.y;
    pose_t[2]     = lidar_pose.t.z;
    int fallbacks = 0;
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) reduction(+ : fallbacks) schedule(static)
    for (int k = 0; k < n; k++) {
        Pose curr;
        if (!pose_from_window(ins_t, ins_p, ins_q, w, ext, time[k] + td, curr)) fallbacks++;
        Pose rel = relative_pose(lidar_pose, curr);
        project_point(rel, xyzi + 4 * k, out_xyzi + 4 * k);
    }
    return fallbacks;
}

/* lidar_map.cc:249-256 */
int orc_decimate(const float *xyzi, int n, int filter_num, float *out) {
    int m = 0, num_points = 0;
    for (int k = 0; k < n; k++) {
        if (num_points++ % filter_num == 0) {
            std::memcpy(out + 4 * m, xyzi + 4 * k, 16);
            m++;
        }
    }
    return m;
}

/* pcl::VoxelGrid<PointXYZI>::applyFilter as used at lidar_map.cc:209-210,260-261. */
int orc_voxel_filter(const float *xyzi, int n, double leaf_d, float *out, int32_t *keys, int32_t *order) {
    if (n <= 0) return 0;
    float leaf = (float) leaf_d;
    float inv  = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            mn[a] = std::min(mn[a], xyzi[4 * i + a]);
            mx[a] = std::max(mx[a], xyzi[4 * i + a]);
        }
    int64_t dx = (int64_t) ((mx[0] - mn[0]) * inv) + 1;
    int64_t dy = (int64_t) ((mx[1] - mn[1]) * inv) + 1;
    int64_t dz = (int64_t) ((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t) INT32_MAX) return -1;
    int min_b[3], max_b[3], div_b[3];
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int) std::floor(mn[a] * inv);
        max_b[a] = (int) std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
    std::vector<std::pair<uint32_t, int>> kv(n);
    for (int i = 0; i < n; i++) {
        int ijk0 = (int) (std::floor(xyzi[4 * i] * inv) - (float) min_b[0]);
        int ijk1 = (int) (std::floor(xyzi[4 * i + 1] * inv) - (float) min_b[1]);
        int ijk2 = (int) (std::floor(xyzi[4 * i + 2] * inv) - (float) min_b[2]);
        int idx  = ijk0 + ijk1 * mul1 + ijk2 * mul2;
        kv[i]    = {(uint32_t) idx, i};
        if (keys) keys[i] = idx;
    }
    std::stable_sort(kv.begin(), kv.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    int m = 0;
    for (int i = 0; i < n;) {
        int j = i;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        while (j < n && kv[j].first == kv[i].first) {
            const float *p = xyzi + 4 * kv[j].second;
            sx += p[0];
            sy += p[1];
            sz += p[2];
            si += p[3];
            j++;
        }
        float cnt      = (float) (j - i);
        out[4 * m]     = sx / cnt;
        out[4 * m + 1] = sy / cnt;
        out[4 * m + 2] = sz / cnt;
        out[4 * m + 3] = si / cnt;
        m++;
        i = j;
    }
    if (order)
        for (int i = 0; i < n; i++) order[i] = kv[i].second;
    return m;
}

/* PointCloudCommon::pointCloudProjection (lidar/pointcloud.cc:38-51). */
void orc_project(const double *R, const double *t, const float *src, int n, float *dst, int threads) {
    Pose T = {ldm(R), ld3(t)};
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int k = 0; k < n; k++) project_point(T, src + 4 * k, dst + 4 * k);
}

void orc_relative_pose(const double *R0, const double *t0, const double *R1, const double *t1, double *R01,
                       double *t01) {
    Pose a = {ldm(R0), ld3(t0)}, b = {ldm(R1), ld3(t1)};
    Pose r = relative_pose(a, b);
    std::memcpy(R01, r.R.m, sizeof(r.R.m));
    t01[0] = r.t.x;
    t01[1] = r.t.y;
    t01[2] = r.t.z;
}

int orc_knn5(const float *map_xyzi, int m, const float *query_xyz, double max_dist, int32_t *idx, float *d2) {
    return knn5_brute(map_xyzi, m, query_xyz, max_dist, idx, d2);
}

int orc_plane_estimation(const float *pts, double threshold, double *unit_normal, double *norm_inverse,
                         double *distance) {
    return plane_estimation(pts, threshold, unit_normal, norm_inverse, distance) ? 1 : 0;
}

/* One query of LidarMap::findFeaturesInMap (lidar/lidar_map.cc:349-398). tree == nullptr: brute force. */
static inline void associate_query(const Pose &T, const KdTree *tree, const float *map_xyzi, int m, const float *world_k,
                                   const float *lidar_k, double max_sq, double plane_thr, double sigma, double scalar_thr,
                                   double sigma_gate, int8_t *type, uint8_t *covered, int32_t *nn_idx, float *nn_d2,
                                   double *normal, double *dd) {
    float ref[4];
    project_point(T, world_k, ref);
    int32_t idx[5] = {-1, -1, -1, -1, -1};
    float d2[5]    = {0, 0, 0, 0, 0};
    int found;
    if (tree) {
        KdTree::Top5 top;
        if (m > 0) tree->search_rec(0, ref, (float) (max_sq * max_sq), top);
        found = top.n;
        for (int j = 0; j < found; j++) {
            idx[j] = top.idx[j];
            d2[j]  = top.d2[j];
        }
    } else {
        found = knn5_brute(map_xyzi, m, ref, max_sq, idx, d2);
    }
    double unit[3] = {0, 0, 0}, ninv = 0;
    int8_t ftype = -1;
    uint8_t cov  = 0;
    if (found == 5) {
        cov = 1;
        if ((double) d2[4] < max_sq) {
            float pts[20];
            for (int j = 0; j < 5; j++) std::memcpy(pts + 4 * j, map_xyzi + 4 * idx[j], 16);
            double dist[5];
            if (plane_estimation(pts, plane_thr, unit, &ninv, dist)) {
                double dis = std::fabs(
                    ((unit[0] * (double) ref[0] + unit[1] * (double) ref[1]) + unit[2] * (double) ref[2]) + ninv);
                const float *l = lidar_k;
                float nrm      = std::sqrt((l[0] * l[0] + l[1] * l[1]) + l[2] * l[2]);
                double scalar  = dis / (double) std::sqrt(nrm); /* fp32 sqrt of fp32 norm (:384) */
                if (scalar < scalar_thr && dis < sigma_gate * sigma) ftype = 0;
            }
        }
    }
    *type    = ftype;
    *covered = cov;
    for (int j = 0; j < 5; j++) {
        nn_idx[j] = j < found ? idx[j] : -1;
        nn_d2[j]  = j < found ? d2[j] : 0.f;
    }
    normal[0] = unit[0];
    normal[1] = unit[1];
    normal[2] = unit[2];
    *dd       = ninv;
}

static inline Pose world_to_ref(const double *ref_R, const double *ref_t) {
    Pose T;
    T.R = tr(ldm(ref_R));
    M3 negR = T.R;
    for (int i = 0; i < 9; i++) negR.m[i] = -negR.m[i];
    T.t = mv(negR, ld3(ref_t));
    return T;
}

/* LidarMap::findFeaturesInMap (lidar/lidar_map.cc:326-408). ref_R/ref_t is the REF frame's
 * pose (lidar->world); the world->ref transform is formed as at :343-345. */
void orc_associate(const double *ref_R, const double *ref_t, const float *cur_world, const float *cur_lidar, int q,
                   const float *map_xyzi, int m, double max_sq, double plane_thr, double sigma, double scalar_thr,
                   double sigma_gate, int knn_mode, int threads, int8_t *type, uint8_t *covered, int32_t *nn_idx,
                   float *nn_d2, double *normal, double *dd, int32_t *num_features, int32_t *num_covered) {
    const Pose T = world_to_ref(ref_R, ref_t);
    KdTree tree;
    if (knn_mode == 1) tree.build(map_xyzi, m);
    int nf = 0, nc = 0;
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) reduction(+ : nf, nc) schedule(dynamic, 64)
    for (int k = 0; k < q; k++) {
        associate_query(T, knn_mode == 1 ? &tree : nullptr, map_xyzi, m, cur_world + 4 * k, cur_lidar + 4 * k, max_sq, plane_thr,
                        sigma, scalar_thr, sigma_gate, type + k, covered + k, nn_idx + 5 * k, nn_d2 + 5 * k, normal + 3 * k, dd + k);
        nf += (type[k] == 0);
        nc += covered[k];
    }
    *num_features = nf;
    *num_covered  = nc;
}

/* ---- the reference's keyframe life cycle for the CPU BASELINE: the tree of a keyframe map is built ONCE when the
 * keyframe is added (LidarMap::addNewKeyFrame, lidar_map.cc:100-118) and searched by every later association. The map
 * array must outlive the handle. */
void *orc_tree_build(const float *map_xyzi, int m) {
    KdTree *t = new KdTree();
    t->build(map_xyzi, m);
    return t;
}
void orc_tree_free(void *tree) { delete static_cast<KdTree *>(tree); }

/* LidarMap::updateLidarFeaturesInMap (lidar_map.cc:428-459): every reference keyframe against the current one, the
 * reference's nested tbb::parallel_for (refs :436-444, queries :349-398) restated as one OpenMP loop over all
 * (ref, query) pairs. Per-ref inputs are arrays of pointers; outputs are [n_refs][q] row-major. */
void orc_associate_window(int n_refs, void *const *trees, const float *const *maps, const int *m, const double *ref_R,
                          const double *ref_t, const float *cur_world, const float *cur_lidar, int q, double max_sq,
                          double plane_thr, double sigma, double scalar_thr, double sigma_gate, int threads, int8_t *type,
                          uint8_t *covered, int32_t *nn_idx, float *nn_d2, double *normal, double *dd, int32_t *num_features,
                          int32_t *num_covered) {
    std::vector<Pose> T(n_refs);
    for (int r = 0; r < n_refs; r++) T[r] = world_to_ref(ref_R + 9 * r, ref_t + 3 * r);
    (void) threads;
    const long total = (long) n_refs * q;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 64)
    for (long e = 0; e < total; e++) {
        const int r = (int) (e / q), k = (int) (e % q);
        const size_t o = (size_t) r * q + k;
        associate_query(T[r], static_cast<const KdTree *>(trees[r]), maps[r], m[r], cur_world + 4 * k, cur_lidar + 4 * k, max_sq,
                        plane_thr, sigma, scalar_thr, sigma_gate, type + o, covered + o, nn_idx + 5 * o, nn_d2 + 5 * o,
                        normal + 3 * o, dd + o);
    }
    for (int r = 0; r < n_refs; r++) {
        int nf = 0, nc = 0;
        for (int k = 0; k < q; k++) {
            nf += type[(size_t) r * q + k] == 0;
            nc += covered[(size_t) r * q + k];
        }
        num_features[r] = nf;
        num_covered[r]  = nc;
    }
}

int orc_lidar_factor_evaluate(const float *point_xyz, const double *n, double d, double std, const double *motion,
                              const double *const *parameters, double *residuals, double **jacobians) {
    FactorConst f = make_const(point_xyz, n, d, std, motion);
    lidar_factor(f, parameters, residuals, jacobians);
    return 1;
}

double orc_huber_correct(double delta, double *r, double *J22) { return huber_correct(delta, r, J22); }

/* Batch of factors: evaluate (F1), optionally correct (F4), optionally reduce (F5,
 * factors/marginalization_info.cc:181-210: H += Ji^T Jj for i<=j then mirror; b -= Ji^T r),
 * sequentially in residual order. */
void orc_factor_batch(int n, const float *point_xyz, const double *normal, const double *d, const double *std,
                      const uint8_t *valid, const double *motion, const double *pose_ref, const double *pose_cur,
                      const double *ext, double td, uint32_t flags, double huber_delta, double *r, double *J,
                      double *H, double *b, double *cost, int threads) {
    const double *par[4] = {pose_ref, pose_cur, ext, &td};
    std::vector<double> rr(n, 0.0), JJ((size_t) 22 * n, 0.0), rho0(n, 0.0);
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int k = 0; k < n; k++) {
        if (valid && !valid[k]) continue;
        FactorConst f = make_const(point_xyz + 3 * k, normal + 3 * k, d[k], std[k], motion);
        double *Jk    = JJ.data() + (size_t) 22 * k;
        double *jac[4] = {Jk, Jk + 7, Jk + 14, Jk + 21};
        lidar_factor(f, par, &rr[k], jac);
        if (flags & 2u)
            for (int i = 0; i < 7; i++) Jk[i] = 0;
        if (flags & 4u)
            for (int i = 0; i < 7; i++) Jk[7 + i] = 0;
        if (flags & 8u)
            for (int i = 0; i < 7; i++) Jk[14 + i] = 0;
        if (flags & 16u) Jk[21] = 0;
        if (flags & 1u)
            rho0[k] = huber_correct(huber_delta, &rr[k], Jk);
        else
            rho0[k] = rr[k] * rr[k];
    }
    if (r) std::memcpy(r, rr.data(), sizeof(double) * n);
    if (J) std::memcpy(J, JJ.data(), sizeof(double) * 22 * n);
    if (H || b || cost) {
        double Hl[19 * 19], bl[19], c0 = 0, c1 = 0;
        std::memset(Hl, 0, sizeof(Hl));
        std::memset(bl, 0, sizeof(bl));
        for (int k = 0; k < n; k++) {
            if (valid && !valid[k]) continue;
            const double *Jk = JJ.data() + (size_t) 22 * k;
            double jl[19];
            for (int a = 0; a < 19; a++) jl[a] = Jk[LOCAL_OF_J[a]];
            for (int a = 0; a < 19; a++) {
                for (int c = a; c < 19; c++) Hl[19 * a + c] += jl[a] * jl[c];
                bl[a] -= jl[a] * rr[k];
            }
            c0 += rho0[k];
            c1 += rr[k] * rr[k];
        }
        for (int a = 0; a < 19; a++)
            for (int c = a + 1; c < 19; c++) Hl[19 * c + a] = Hl[19 * a + c];
        if (H) std::memcpy(H, Hl, sizeof(Hl));
        if (b) std::memcpy(b, bl, sizeof(bl));
        if (cost) {
            cost[0] = 0.5 * c0;
            cost[1] = c1;
        }
    }
}

/* Optimizer::lidarOutlierCullingByChi2 (ff_lins/optimizer.cc:515-548). */
int orc_chi2(int n, const float *point_xyz, const double *normal, const double *d, const double *std,
             const uint8_t *valid, const double *motion, const double *pose_ref, const double *pose_cur,
             const double *ext, double td, double threshold, uint8_t *outlier, int threads) {
    const double *par[4] = {pose_ref, pose_cur, ext, &td};
    int count            = 0;
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) reduction(+ : count) schedule(static)
    for (int k = 0; k < n; k++) {
        outlier[k] = 0;
        if (valid && !valid[k]) continue;
        FactorConst f = make_const(point_xyz + 3 * k, normal + 3 * k, d[k], std[k], motion);
        double res;
        lidar_factor(f, par, &res, nullptr);
        if (res * res > threshold) {
            outlier[k] = 1;
            count++;
        }
    }
    return count;
}


/* One solver iteration's worth of LiDAR factor work for the CPU BASELINE, in one call: every valid residual of every
 * (ref, cur) pair evaluated in parallel (Ceres evaluates residual blocks on its thread pool, optimizer.cc:103) and
 * reduced the way MarginalizationInfo::constructEquation does it (marginalization_info.cc:133-179): the factors are
 * split into `threads` chunks, each chunk accumulates its own H0 / b0 (constructSingleEquation :181-210), the chunks
 * are summed serially (:174-177). Inputs per pair are [n_pairs][q] row-major (normal [n_pairs][q][3]); motions
 * [n_pairs][14]; pose_refs [n_pairs][7]. Outputs H [n_pairs][361], b [n_pairs][19], cost [n_pairs][2]. */
void orc_window_eval(int n_pairs, int q, const float *point_xyz, const double *normal, const double *d, double std,
                     const uint8_t *valid, const double *motions, const double *pose_refs, const double *pose_cur,
                     const double *ext, double td, uint32_t flags, double huber_delta, double *H, double *b, double *cost,
                     int threads) {
    const int T = threads > 0 ? threads : 1;
    const long total = (long) n_pairs * q;
    struct Acc {
        double H[19 * 19], b[19], c0, c1;
    };
    std::vector<Acc> acc((size_t) T * n_pairs);
    std::memset(acc.data(), 0, acc.size() * sizeof(Acc));
#pragma omp parallel num_threads(T)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        /* contiguous chunk of the factor list per thread, as constructEquation hands them out */
        const long lo = total * tid / T, hi = total * (tid + 1) / T;
        for (long e = lo; e < hi; e++) {
            const int p = (int) (e / q), k = (int) (e % q);
            if (valid && !valid[e]) continue;
            const double *par[4] = {pose_refs + 7 * p, pose_cur, ext, &td};
            FactorConst f        = make_const(point_xyz + 3 * k, normal + 3 * e, d[e], std, motions + 14 * p);
            double Jk[22], r;
            double *jac[4] = {Jk, Jk + 7, Jk + 14, Jk + 21};
            lidar_factor(f, par, &r, jac);
            if (flags & 2u)
                for (int i = 0; i < 7; i++) Jk[i] = 0;
            if (flags & 4u)
                for (int i = 0; i < 7; i++) Jk[7 + i] = 0;
            if (flags & 8u)
                for (int i = 0; i < 7; i++) Jk[14 + i] = 0;
            if (flags & 16u) Jk[21] = 0;
            double rho0 = (flags & 1u) ? huber_correct(huber_delta, &r, Jk) : r * r;
            Acc &a = acc[(size_t) tid * n_pairs + p];
            double jl[19];
            for (int i = 0; i < 19; i++) jl[i] = Jk[LOCAL_OF_J[i]];
            for (int i = 0; i < 19; i++) {
                for (int c = i; c < 19; c++) a.H[19 * i + c] += jl[i] * jl[c];
                a.b[i] -= jl[i] * r;
            }
            a.c0 += rho0;
            a.c1 += r * r;
        }
    }
    for (int p = 0; p < n_pairs; p++) {
        Acc s;
        std::memset(&s, 0, sizeof(s));
        for (int t = 0; t < T; t++) {
            const Acc &a = acc[(size_t) t * n_pairs + p];
            for (int i = 0; i < 361; i++) s.H[i] += a.H[i];
            for (int i = 0; i < 19; i++) s.b[i] += a.b[i];
            s.c0 += a.c0;
            s.c1 += a.c1;
        }
        for (int i = 0; i < 19; i++)
            for (int c = i + 1; c < 19; c++) s.H[19 * c + i] = s.H[19 * i + c];
        if (H) std::memcpy(H + 361 * (size_t) p, s.H, sizeof(s.H));
        if (b) std::memcpy(b + 19 * (size_t) p, s.b, sizeof(s.b));
        if (cost) {
            cost[2 * p]     = 0.5 * s.c0;
            cost[2 * p + 1] = s.c1;
        }
    }
}

/* Optimizer::lidarOutlierCullingByChi2 over the whole window in one call (optimizer.cc:515-548: one tbb loop over all
 * residual blocks). outlier [n_pairs][q] is OR-ed in place; counts [n_pairs]. */
void orc_window_chi2(int n_pairs, int q, const float *point_xyz, const double *normal, const double *d, double std,
                     const uint8_t *valid, const double *motions, const double *pose_refs, const double *pose_cur,
                     const double *ext, double td, double threshold, uint8_t *outlier, int32_t *counts, int threads) {
    const long total = (long) n_pairs * q;
    (void) threads;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (long e = 0; e < total; e++) {
        const int p = (int) (e / q), k = (int) (e % q);
        if ((valid && !valid[e]) || outlier[e]) continue;
        const double *par[4] = {pose_refs + 7 * p, pose_cur, ext, &td};
        FactorConst f        = make_const(point_xyz + 3 * k, normal + 3 * e, d[e], std, motions + 14 * p);
        double res;
        lidar_factor(f, par, &res, nullptr);
        if (res * res > threshold) outlier[e] = 2; /* newly culled */
    }
    for (int p = 0; p < n_pairs; p++) {
        int c = 0;
        for (int k = 0; k < q; k++) {
            uint8_t &o = outlier[(size_t) p * q + k];
            if (o == 2) {
                o = 1;
                c++;
            }
        }
        counts[p] = c;
    }
}

} /* extern "C" */
