"""ctypes binding of oracle/_ref/libff_ref.so: the REFERENCE'S OWN lidar_factor.cc, misc.cc, rotation.cc and pointcloud.cc
compiled where they lie under /root/reference (oracle/Makefile `ref`, oracle/ffref_wrap.cc) against the stand-in headers of
oracle/refshim/. Same argument layouts as oracle/pyoracle.py. Test infrastructure only: used by tests/ and by
tests/golden/make_golden_ffref.py to pin the oracle; nothing under ff-lins_b200/ imports it."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libff_ref.so")
PATH_OMP = os.path.join(HERE, "_ref", "libff_ref_omp.so")   # -O3, TBB sites as OpenMP tasks: the CPU-baseline build
_L = None
_LOMP = None
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)


def available():
    return os.path.exists(PATH)


def omp_available():
    return os.path.exists(PATH_OMP)


def lib(omp=False):
    global _L, _LOMP
    if omp:
        if _LOMP is None:
            _LOMP = C.CDLL(PATH_OMP)
        return _LOMP
    if _L is None:
        L = C.CDLL(PATH)
        L.ffref_ins_window_index.argtypes = [dp, C.c_int, C.c_double]
        L.ffref_pose_from_ins_window.argtypes = [dp, dp, dp, C.c_int, dp, dp, C.c_double, dp, dp]
        L.ffref_state_to_pose.argtypes = [dp, dp, dp, dp, dp, dp]
        L.ffref_state_to_pose.restype = None
        L.ffref_undistort.argtypes = [fp, dp, C.c_int, dp, dp, dp, C.c_int, dp, dp, C.c_double, C.c_double, fp, dp, dp]
        L.ffref_project.argtypes = [dp, dp, fp, C.c_int, fp]
        L.ffref_project.restype = None
        L.ffref_plane_estimation.argtypes = [fp, C.c_double, dp, dp, dp]
        L.ffref_lidar_factor_evaluate.argtypes = [fp, dp, C.c_double, C.c_double, dp, C.POINTER(dp), dp, C.POINTER(dp)]
        L.ffref_ins_mechanization.argtypes = [dp, dp, dp, dp, dp]
        L.ffref_ins_mechanization.restype = None
        _L = L
    return _L


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _P(a):
    return a.ctypes.data_as(dp)


def _F(a):
    return a.ctypes.data_as(fp)


def ins_window_index(ins_t, time):
    ins_t = _d(ins_t)
    return lib().ffref_ins_window_index(_P(ins_t), len(ins_t), float(time))


def pose_from_ins_window(ins_t, ins_p, ins_q, R_bl, t_bl, time):
    ins_t, ins_p, ins_q, R_bl, t_bl = _d(ins_t), _d(ins_p), _d(ins_q), _d(R_bl).reshape(-1), _d(t_bl)
    R, t = np.zeros(9), np.zeros(3)
    ok = lib().ffref_pose_from_ins_window(_P(ins_t), _P(ins_p), _P(ins_q), len(ins_t), _P(R_bl), _P(t_bl), float(time),
                                          _P(R), _P(t))
    return bool(ok), R.reshape(3, 3), t


def state_to_pose(p, q, R_bl, t_bl):
    p, q, R_bl, t_bl = _d(p), _d(q), _d(R_bl).reshape(-1), _d(t_bl)
    R, t = np.zeros(9), np.zeros(3)
    lib().ffref_state_to_pose(_P(p), _P(q), _P(R_bl), _P(t_bl), _P(R), _P(t))
    return R.reshape(3, 3), t


def undistort(xyzi, time, ins_t, ins_p, ins_q, R_bl, t_bl, stamp, td):
    xyzi, time = _f(xyzi), _d(time)
    ins_t, ins_p, ins_q, R_bl, t_bl = _d(ins_t), _d(ins_p), _d(ins_q), _d(R_bl).reshape(-1), _d(t_bl)
    out = np.zeros((len(xyzi), 4), np.float32)
    R, t = np.zeros(9), np.zeros(3)
    fb = lib().ffref_undistort(_F(xyzi), _P(time), len(xyzi), _P(ins_t), _P(ins_p), _P(ins_q), len(ins_t), _P(R_bl),
                               _P(t_bl), float(stamp), float(td), _F(out), _P(R), _P(t))
    return out, R.reshape(3, 3), t, fb


def project(R, t, src):
    R, t, src = _d(R).reshape(-1), _d(t), _f(src)
    dst = np.zeros_like(src)
    lib().ffref_project(_P(R), _P(t), _F(src), len(src), _F(dst))
    return dst


def plane_estimation(pts5, threshold):
    pts5 = _f(pts5)
    n, ninv, dist = np.zeros(3), np.zeros(1), np.zeros(5)
    ok = lib().ffref_plane_estimation(_F(pts5), float(threshold), _P(n), _P(ninv), _P(dist))
    return bool(ok), n, float(ninv[0]), dist


def lidar_factor_evaluate(point_xyz, normal, d, std, motion, pose_ref, pose_cur, ext, td, want_jac=(1, 1, 1, 1)):
    params = [_d(pose_ref), _d(pose_cur), _d(ext), np.array([td], np.float64)]
    par = (dp * 4)(*[_P(p) for p in params])
    res = np.zeros(1)
    jacs = [np.full(s, np.nan) for s in (7, 7, 7, 1)]
    jp = None if want_jac is None else (dp * 4)(*[(_P(j) if w else dp()) for j, w in zip(jacs, want_jac)])
    pt, nn, mo = _f(point_xyz), _d(normal), _d(motion)
    lib().ffref_lidar_factor_evaluate(_F(pt), _P(nn), float(d), float(std), _P(mo), par, _P(res), jp)
    return float(res[0]), jacs


def ins_mechanization(gravity, imu_pre, imu_cur, state16, state_time):
    """imu = [time, dt, dtheta(3), dvel(3)]; state16 = [p, q xyzw, v, bg, ba]; returns (state16, time)."""
    g, a, b = _d(gravity), _d(imu_pre), _d(imu_cur)
    s = _d(state16).copy()
    tm = np.array([state_time], np.float64)
    lib().ffref_ins_mechanization(_P(g), _P(a), _P(b), _P(s), _P(tm))
    return s, float(tm[0])


class RefLidarMap:
    """The reference's own lidar::LidarMap (lidar_map.cc compiled from /root/reference, oracle/ffref_map_wrap.cc), driven in
    the order ff_lins.cc drives it. Keyframes are addressed by their index in the window (0 = oldest, -1 = newest)."""
    CLOUD_LIDAR, CLOUD_WORLD, CLOUD_MAP, CLOUD_MAP_FULL, CLOUD_LIDAR_FULL = range(5)

    def __init__(self, point_filter_num, point_std=0.1, window_size=10, downsample=0.5, frame_rate=10.0, plane_threshold=0.1,
                 min_keyframe_translation=0.3, omp=False, threads=None):
        L = lib(omp)
        L.ffref_map_set_motion.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.ffref_map_add_factors.argtypes = [C.c_void_p]
        L.ffref_map_evaluate_factors.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp]
        L.ffref_map_evaluate_factors.restype = None
        L.ffref_map_chi2.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_double, C.POINTER(C.c_int32)]
        L.ffref_map_reproject.argtypes = [C.c_void_p]
        L.ffref_map_create.restype = C.c_void_p
        L.ffref_map_create.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.ffref_map_frame.argtypes = [C.c_void_p, fp, dp, C.c_int, dp, dp, dp, C.c_int, dp, dp, C.c_double, C.c_double]
        for name in ("ffref_map_destroy", "ffref_map_keyframe", "ffref_map_update_features", "ffref_map_remove_oldest",
                     "ffref_map_remove_newest", "ffref_map_window"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ffref_map_set_pose.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.ffref_map_get_pose.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.ffref_map_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, C.c_int]
        L.ffref_map_features.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), dp, dp, fp, C.POINTER(C.c_int32), dp,
                                         C.POINTER(C.c_uint8)]
        L.ffref_map_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        self.L = L
        self.threads = L.ffref_set_threads(int(threads or os.cpu_count() or 1)) if omp else 1
        self.h = L.ffref_map_create(point_std, window_size, downsample, frame_rate, plane_threshold, min_keyframe_translation,
                                    point_filter_num)

    def close(self):
        if self.h:
            self.L.ffref_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _key(self, k):
        n = self.window()
        return k + n if k < 0 else k

    def frame(self, xyzi, time, ins_t, ins_p, ins_q, R_bl, t_bl, stamp, td):
        """lidarFrameProcessing on a new frame; returns the size of its down-sampled cloud."""
        xyzi, time = _f(xyzi), _d(time)
        ins_t, ins_p, ins_q, R_bl, t_bl = _d(ins_t), _d(ins_p), _d(ins_q), _d(R_bl).reshape(-1), _d(t_bl)
        return self.L.ffref_map_frame(self.h, _F(xyzi), _P(time), len(xyzi), _P(ins_t), _P(ins_p), _P(ins_q), len(ins_t),
                                      _P(R_bl), _P(t_bl), float(stamp), float(td))

    def keyframe(self):
        """The newest pending frame becomes a keyframe (merge, addNewKeyFrame, updateLidarFeaturesInMap)."""
        return self.L.ffref_map_keyframe(self.h)

    def update_features(self):
        self.L.ffref_map_update_features(self.h)

    def remove_oldest(self):
        self.L.ffref_map_remove_oldest(self.h)

    def remove_newest(self):
        self.L.ffref_map_remove_newest(self.h)

    def window(self):
        return self.L.ffref_map_window(self.h)

    def set_pose(self, k, R, t):
        R, t = _d(R).reshape(-1), _d(t)
        self.L.ffref_map_set_pose(self.h, self._key(k), _P(R), _P(t))

    def pose(self, k):
        R, t = np.zeros(9), np.zeros(3)
        self.L.ffref_map_get_pose(self.h, self._key(k), _P(R), _P(t))
        return R.reshape(3, 3), t

    def keyframe_selection(self, R, t, stamp):
        """LidarMap::keyframeSelection of a candidate frame against the newest keyframe -> (is_keyframe, stats[3])."""
        self.L.ffref_map_keyframe_selection.argtypes = [C.c_void_p, dp, dp, C.c_double, dp]
        R, t = _d(R).reshape(-1), _d(t)
        st = np.zeros(3)
        ok = self.L.ffref_map_keyframe_selection(self.h, _P(R), _P(t), float(stamp), _P(st))
        return bool(ok), st

    def cloud(self, k, which):
        n = self.L.ffref_map_cloud(self.h, self._key(k), which, None, 0)
        out = np.zeros((n, 4), np.float32)
        if n:
            self.L.ffref_map_cloud(self.h, self._key(k), which, _F(out), n)
        return out

    def features(self, k):
        k = self._key(k)
        i32 = C.POINTER(C.c_int32)
        n = self.L.ffref_map_features(self.h, k, None, None, None, None, None, None, None)
        out = dict(type=np.zeros(n, np.int32), normal=np.zeros((n, 3)), d=np.zeros(n), nearest=np.zeros((n, 5, 4), np.float32),
                   n_nearest=np.zeros(n, np.int32), std=np.zeros(n), outlier=np.zeros(n, np.uint8))
        if n:
            self.L.ffref_map_features(self.h, k, out["type"].ctypes.data_as(i32), _P(out["normal"]), _P(out["d"]),
                                      _F(out["nearest"]), out["n_nearest"].ctypes.data_as(i32), _P(out["std"]),
                                      out["outlier"].ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    # ---- the solver-facing part (optimizer.cc:437-548)
    def set_motion(self, k, v, omega):
        v, omega = _d(v), _d(omega)
        self.L.ffref_map_set_motion(self.h, self._key(k), _P(v), _P(omega))

    def add_factors(self):
        """addLidarFactors: one LidarFactor per PLANE, non-outlier feature of every (reference, newest) pair."""
        return self.L.ffref_map_add_factors(self.h)

    def evaluate_factors(self, pose_refs, pose_cur, ext, td, huber_delta=1.0, out=None):
        n = self.window() - 1
        pose_refs, pose_cur, ext = _d(pose_refs), _d(pose_cur), _d(ext)
        H, b, cost = out if out is not None else (np.zeros((n, 19, 19)), np.zeros((n, 19)), np.zeros(n))
        self.L.ffref_map_evaluate_factors(self.h, _P(pose_refs), _P(pose_cur), _P(ext), float(td), float(huber_delta), _P(H),
                                          _P(b), _P(cost))
        return H, b, cost

    def chi2(self, pose_refs, pose_cur, ext, td, threshold=3.841):
        n = self.window() - 1
        pose_refs, pose_cur, ext = _d(pose_refs), _d(pose_cur), _d(ext)
        per = np.zeros(n, np.int32)
        total = self.L.ffref_map_chi2(self.h, _P(pose_refs), _P(pose_cur), _P(ext), float(td), float(threshold),
                                      per.ctypes.data_as(C.POINTER(C.c_int32)))
        return total, per

    def reproject(self):
        self.L.ffref_map_reproject(self.h)

    def counts(self):
        a, b = C.c_int32(), C.c_int32()
        self.L.ffref_map_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def residual_block_evaluate(point_xyz, normal, d, std, motion, pose_ref, pose_cur, ext, td, huber_delta):
    """ResidualBlockInfo::Evaluate (residual_block_info.cc:33-80) on a LidarFactor with ceres::HuberLoss(huber_delta)
    (<= 0: no loss): corrected residual and the 22 corrected Jacobian entries [pose_ref 7 | pose_cur 7 | ext 7 | td]."""
    L = lib()
    L.ffref_residual_block_evaluate.argtypes = [fp, dp, C.c_double, C.c_double, dp, C.POINTER(dp), C.c_double, dp, dp]
    L.ffref_residual_block_evaluate.restype = None
    params = [_d(pose_ref), _d(pose_cur), _d(ext), np.array([td], np.float64)]
    par = (dp * 4)(*[_P(p) for p in params])
    res, J = np.zeros(1), np.zeros(22)
    pt, nn, mo = _f(point_xyz), _d(normal), _d(motion)
    L.ffref_residual_block_evaluate(_F(pt), _P(nn), float(d), float(std), _P(mo), par, float(huber_delta), _P(res), _P(J))
    return float(res[0]), J


def pose_plus(x, delta):
    """PoseManifold::Plus (pose_manifold.cc:35-51): x = [t, q xyzw], delta = [dt, rotation vector]."""
    x, delta = _d(x), _d(delta)
    out = np.zeros(7)
    lib().ffref_pose_plus.restype = None
    lib().ffref_pose_plus(_P(x), _P(delta), _P(out))
    return out


def pose_minus(y, x):
    y, x = _d(y), _d(x)
    out = np.zeros(6)
    lib().ffref_pose_minus.restype = None
    lib().ffref_pose_minus(_P(y), _P(x), _P(out))
    return out


def pose_jacobians(x):
    x = _d(x)
    jp, jm = np.zeros((7, 6)), np.zeros((6, 7))
    lib().ffref_pose_jacobians.restype = None
    lib().ffref_pose_jacobians(_P(x), _P(jp), _P(jm))
    return jp, jm


def marginalize_lidar_pair(point_xyz, normal, d, std, motion, pose_ref, pose_cur, ext, td, huber_delta=1.0, evaluate_at=None):
    """MarginalizationInfo (marginalization_info.cc) over the LidarFactor blocks of one (oldest reference, newest) pair with
    the reference pose marginalized: the prior (J0 13 x 13, e0 13) on [pose_cur 6 | ext 6 | td 1]. evaluate_at = (pose_cur7,
    ext7, td): additionally MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91) of that prior there -> also
    returns (residuals 13, tangent Jacobian 13 x 13)."""
    L = lib()
    L.ffref_marginalize_lidar_pair.argtypes = [C.c_int, fp, dp, dp, C.c_double, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp,
                                               dp, dp, C.c_double, dp, dp]
    pt, nn, dd, mo = _f(point_xyz).reshape(-1, 3), _d(normal), _d(d), _d(motion)
    pr, pc, ex = _d(pose_ref), _d(pose_cur), _d(ext)
    J0, e0 = np.zeros((13, 13)), np.zeros(13)
    if evaluate_at is None:
        ok = L.ffref_marginalize_lidar_pair(len(pt), _F(pt), _P(nn), _P(dd), float(std), _P(mo), _P(pr), _P(pc), _P(ex), float(td),
                                            float(huber_delta), _P(J0), _P(e0), None, None, 0.0, None, None)
        if not ok:
            raise RuntimeError("the reference reports the marginalization as invalid")
        return J0, e0
    ec, ee = _d(evaluate_at[0]), _d(evaluate_at[1])
    res, jac = np.zeros(13), np.zeros((13, 13))
    ok = L.ffref_marginalize_lidar_pair(len(pt), _F(pt), _P(nn), _P(dd), float(std), _P(mo), _P(pr), _P(pc), _P(ex), float(td),
                                        float(huber_delta), _P(J0), _P(e0), _P(ec), _P(ee), float(evaluate_at[2]), _P(res), _P(jac))
    if not ok:
        raise RuntimeError("the reference reports the marginalization as invalid")
    return J0, e0, res, jac


class RefPreintegration:
    """PreintegrationNormal of the reference (preintegration_base.cc, preintegration_normal.cc) through PreintegrationFactor.
    imu = [time, dt, dtheta(3), dvel(3)]; state16 = [p, q xyzw, v, bg, ba]; noise = [acc_vrw, gyr_arw, gyr_bias_std,
    acc_bias_std, corr_time, gravity]."""

    def __init__(self, noise, imu0, state16, time):
        L = self.L = lib()
        L.ffref_preint_create.restype = C.c_void_p
        L.ffref_preint_create.argtypes = [dp, dp, dp, C.c_double]
        L.ffref_preint_destroy.argtypes = [C.c_void_p]
        L.ffref_preint_add.argtypes = [C.c_void_p, dp]
        L.ffref_preint_states.argtypes = [C.c_void_p, dp, dp, dp]
        L.ffref_preint_matrices.argtypes = [C.c_void_p, dp, dp]
        L.ffref_preint_evaluate.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, dp, dp, dp]
        for name in ("ffref_preint_destroy", "ffref_preint_add", "ffref_preint_states", "ffref_preint_matrices", "ffref_preint_evaluate"):
            getattr(L, name).restype = None
        n, i0, s = _d(noise), _d(imu0), _d(state16)
        self.h = L.ffref_preint_create(_P(n), _P(i0), _P(s), float(time))

    def __del__(self):
        try:
            if self.h:
                self.L.ffref_preint_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add(self, imu):
        m = _d(imu)
        self.L.ffref_preint_add(self.h, _P(m))

    def states(self):
        cur, dlt, dt = np.zeros(16), np.zeros(16), np.zeros(1)
        self.L.ffref_preint_states(self.h, _P(cur), _P(dlt), _P(dt))
        return cur, dlt, float(dt[0])

    def matrices(self):
        cov, jac = np.zeros((15, 15)), np.zeros((15, 15))
        self.L.ffref_preint_matrices(self.h, _P(cov), _P(jac))
        return cov, jac

    def evaluate(self, pose0, mix0, pose1, mix1):
        a = [_d(x) for x in (pose0, mix0, pose1, mix1)]
        r = np.zeros(15)
        J = [np.zeros((15, 7)), np.zeros((15, 9)), np.zeros((15, 7)), np.zeros((15, 9))]
        self.L.ffref_preint_evaluate(self.h, _P(a[0]), _P(a[1]), _P(a[2]), _P(a[3]), _P(r), *[_P(j) for j in J])
        return r, J


def ins_window_scenario(gravity, imu, state16, series_start, series_end, updated16, updated_time, reserved):
    """An INS window built as LINS builds it (insMechanization per epoch), then MISC::getImuSeriesFromTo(start, end) and
    MISC::redoInsMechanization(updated, reserved): returns (series_ok, series rows [time, dt, dtheta, dvel], window times,
    window states16 after the redo)."""
    L = lib()
    L.ffref_ins_window_scenario.argtypes = [dp, dp, C.c_int, dp, C.c_double, C.c_double, C.c_double, dp, C.c_int,
                                            C.POINTER(C.c_int), dp, C.c_double, C.c_int, dp, dp, C.c_int]
    g, imu, s0, up = _d(gravity), _d(imu), _d(state16), _d(updated16)
    n = len(imu)
    series, ok = np.zeros((n + 4, 8)), C.c_int()
    wt, ws = np.zeros(n + 4), np.zeros((n + 4, 16))
    code = L.ffref_ins_window_scenario(_P(g), _P(imu), n, _P(s0), float(imu[0, 0]), float(series_start), float(series_end),
                                       _P(series), n + 4, C.byref(ok), _P(up), float(updated_time), int(reserved), _P(wt), _P(ws),
                                       n + 4)
    ns, nw = code // 100000, code % 100000
    return bool(ok.value), series[:ns], wt[:nw], ws[:nw]


REC32 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad", "<f4"), ("intensity", "<f4"), ("pad2", "<f4"), ("time", "<f8")])
LIVOX19 = np.dtype([("offset_time", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("reflectivity", "u1"), ("tag", "u1"),
                    ("line", "u1")])


def convert_livox(points19, sec, nsec, scan_line, min_range, max_range, to_gps_time=False):
    """LidarConverter::livoxPointCloudConvertion (ROS/lidar/lidar_converter.cc:36-68) on a CustomMsg with stamp (sec, nsec)
    and the packed points (LIVOX19 records). Returns (32-byte records, start, end)."""
    L = lib()
    L.ffref_convert_livox.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_int,
                                      C.c_void_p, C.c_int, dp, dp]
    pts = np.ascontiguousarray(points19, LIVOX19)
    n = len(pts)
    out = np.zeros(max(n, 1), REC32)
    s, e = np.zeros(1), np.zeros(1)
    m = L.ffref_convert_livox(pts.ctypes.data, n, int(sec), int(nsec), int(scan_line), float(min_range), float(max_range),
                              int(to_gps_time), out.ctypes.data, n, _P(s), _P(e))
    return out[:m].copy(), float(s[0]), float(e[0])


def convert_generic(kind, xyzi, time, sec, nsec, min_range, max_range, to_gps_time=False):
    """velodyne (kind 2) / ouster (3) / hesai (4) PointCloudConvertion (lidar_converter.cc:70-213); `time` is the sensor's
    per-point time field in its own unit (s relative, ns relative, absolute s)."""
    L = lib()
    L.ffref_convert_generic.argtypes = [C.c_int, fp, dp, C.c_int, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_int,
                                        C.c_void_p, C.c_int, dp, dp]
    xyzi, time = _f(xyzi).reshape(-1, 4), _d(time)
    n = len(xyzi)
    out = np.zeros(max(n, 1), REC32)
    s, e = np.zeros(1), np.zeros(1)
    m = L.ffref_convert_generic(int(kind), _F(xyzi), _P(time), n, int(sec), int(nsec), float(min_range), float(max_range),
                                int(to_gps_time), out.ctypes.data, n, _P(s), _P(e))
    return out[:m].copy(), float(s[0]), float(e[0])
