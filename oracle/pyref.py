"""ctypes binding of oracle/_ref/libff_ref.so: the REFERENCE'S OWN lidar_factor.cc, misc.cc, rotation.cc and pointcloud.cc
compiled where they lie under /root/reference (oracle/Makefile `ref`, oracle/ffref_wrap.cc) against the stand-in headers of
oracle/refshim/. Same argument layouts as oracle/pyoracle.py. Test infrastructure only: used by tests/ and by
tests/golden/make_golden_ffref.py to pin the oracle; nothing under ff-lins_b200/ imports it."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libff_ref.so")
_L = None
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)


def available():
    return os.path.exists(PATH)


def lib():
    global _L
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
