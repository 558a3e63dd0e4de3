"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the reference's own
ikd-Tree (oracle/_ref/libikd_ref.so). TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package (ff-lins_b200/)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_IKD = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _LIB
    if _LIB is None:
        # FFLINS_ORACLE_LIB: an alternative build of the same source (bench.py's labelled -march=native figure)
        path = os.environ.get("FFLINS_ORACLE_LIB") or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_ins_window_index.argtypes = [f64p, C.c_int, C.c_double]
        L.orc_pose_from_ins_window.argtypes = [f64p, f64p, f64p, C.c_int, f64p, f64p, C.c_double, f64p, f64p]
        L.orc_state_to_pose.argtypes = [f64p, f64p, f64p, f64p, f64p, f64p]
        L.orc_undistort.argtypes = [f32p, f64p, C.c_int, f64p, f64p, f64p, C.c_int, f64p, f64p, C.c_double,
                                    C.c_double, f32p, f64p, f64p, C.c_int]
        L.orc_decimate.argtypes = [f32p, C.c_int, C.c_int, f32p]
        L.orc_voxel_filter.argtypes = [f32p, C.c_int, C.c_double, f32p, C.c_void_p, C.c_void_p]
        L.orc_project.argtypes = [f64p, f64p, f32p, C.c_int, f32p, C.c_int]
        L.orc_relative_pose.argtypes = [f64p, f64p, f64p, f64p, f64p, f64p]
        L.orc_knn5.argtypes = [f32p, C.c_int, f32p, C.c_double, i32p, f32p]
        L.orc_plane_estimation.argtypes = [f32p, C.c_double, f64p, f64p, f64p]
        L.orc_associate.argtypes = [f64p, f64p, f32p, f32p, C.c_int, f32p, C.c_int, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, i8p, u8p, i32p, f32p,
                                    f64p, f64p, i32p, i32p]
        L.orc_lidar_factor_evaluate.argtypes = [f32p, f64p, C.c_double, C.c_double, f64p,
                                                C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_double),
                                                C.POINTER(C.POINTER(C.c_double))]
        L.orc_huber_correct.argtypes = [C.c_double, f64p, C.c_void_p]
        L.orc_huber_correct.restype = C.c_double
        L.orc_factor_batch.argtypes = [C.c_int, f32p, f64p, f64p, f64p, C.c_void_p, f64p, f64p, f64p, f64p,
                                       C.c_double, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]
        L.orc_factor_batch.restype = None
        L.orc_chi2.argtypes = [C.c_int, f32p, f64p, f64p, f64p, C.c_void_p, f64p, f64p, f64p, f64p, C.c_double,
                               C.c_double, u8p, C.c_int]
        pp = C.POINTER(C.c_void_p)
        L.orc_tree_build.argtypes = [f32p, C.c_int]
        L.orc_tree_build.restype = C.c_void_p
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_tree_free.restype = None
        L.orc_associate_window.argtypes = [C.c_int, pp, pp, i32p, f64p, f64p, f32p, f32p, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_double, C.c_int, i8p, u8p, i32p, f32p, f64p, f64p,
                                           i32p, i32p]
        L.orc_associate_window.restype = None
        L.orc_window_eval.argtypes = [C.c_int, C.c_int, f32p, f64p, f64p, C.c_double, u8p, f64p, f64p, f64p, f64p,
                                      C.c_double, C.c_uint32, C.c_double, f64p, f64p, f64p, C.c_int]
        L.orc_window_eval.restype = None
        L.orc_window_chi2.argtypes = [C.c_int, C.c_int, f32p, f64p, f64p, C.c_double, u8p, f64p, f64p, f64p, f64p,
                                      C.c_double, C.c_double, u8p, i32p, C.c_int]
        L.orc_window_chi2.restype = None
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def max_threads():
    return lib().orc_max_threads()


def ins_window_index(ins_t, time):
    ins_t = _c(ins_t, np.float64)
    return lib().orc_ins_window_index(ins_t, len(ins_t), float(time))


def pose_from_ins_window(ins_t, ins_p, ins_q, R_bl, t_bl, time):
    R = np.zeros(9)
    t = np.zeros(3)
    ok = lib().orc_pose_from_ins_window(_c(ins_t, np.float64), _c(ins_p, np.float64), _c(ins_q, np.float64),
                                        len(ins_t), _c(R_bl, np.float64).reshape(-1), _c(t_bl, np.float64),
                                        float(time), R, t)
    return bool(ok), R.reshape(3, 3), t


def state_to_pose(p, q, R_bl, t_bl):
    R = np.zeros(9)
    t = np.zeros(3)
    lib().orc_state_to_pose(_c(p, np.float64), _c(q, np.float64), _c(R_bl, np.float64).reshape(-1),
                            _c(t_bl, np.float64), R, t)
    return R.reshape(3, 3), t


def undistort(xyzi, time, ins_t, ins_p, ins_q, R_bl, t_bl, stamp, td, threads=1):
    xyzi = _c(xyzi, np.float32)
    n = len(xyzi)
    out = np.zeros((n, 4), np.float32)
    R = np.zeros(9)
    t = np.zeros(3)
    fb = lib().orc_undistort(xyzi, _c(time, np.float64), n, _c(ins_t, np.float64), _c(ins_p, np.float64),
                             _c(ins_q, np.float64), len(ins_t), _c(R_bl, np.float64).reshape(-1),
                             _c(t_bl, np.float64), float(stamp), float(td), out, R, t, threads)
    return out, R.reshape(3, 3), t, fb


def decimate(xyzi, filter_num):
    xyzi = _c(xyzi, np.float32)
    out = np.zeros_like(xyzi)
    m = lib().orc_decimate(xyzi, len(xyzi), filter_num, out)
    return out[:m].copy()


def voxel_filter(xyzi, leaf, want_keys=False):
    xyzi = _c(xyzi, np.float32)
    n = len(xyzi)
    out = np.zeros((max(n, 1), 4), np.float32)
    keys = np.zeros(max(n, 1), np.int32)
    order = np.zeros(max(n, 1), np.int32)
    m = lib().orc_voxel_filter(xyzi.reshape(-1) if n else np.zeros(4, np.float32), n, float(leaf), out,
                               keys.ctypes.data, order.ctypes.data)
    if m < 0:
        raise OverflowError("voxel grid index overflow")
    if want_keys:
        return out[:m].copy(), keys[:n].copy(), order[:n].copy()
    return out[:m].copy()


def project(R, t, src, threads=1):
    src = _c(src, np.float32)
    dst = np.zeros_like(src)
    lib().orc_project(_c(R, np.float64).reshape(-1), _c(t, np.float64), src, len(src), dst, threads)
    return dst


def relative_pose(R0, t0, R1, t1):
    R = np.zeros(9)
    t = np.zeros(3)
    lib().orc_relative_pose(_c(R0, np.float64).reshape(-1), _c(t0, np.float64), _c(R1, np.float64).reshape(-1),
                            _c(t1, np.float64), R, t)
    return R.reshape(3, 3), t


def knn5(map_xyzi, query_xyz, max_dist=5.0):
    map_xyzi = _c(map_xyzi, np.float32)
    idx = np.full(5, -1, np.int32)
    d2 = np.zeros(5, np.float32)
    n = lib().orc_knn5(map_xyzi if len(map_xyzi) else np.zeros((1, 4), np.float32), len(map_xyzi),
                       _c(query_xyz, np.float32), float(max_dist), idx, d2)
    return n, idx, d2


def plane_estimation(pts5, threshold):
    pts5 = _c(pts5, np.float32).reshape(5, 4)
    unit = np.zeros(3)
    ninv = np.zeros(1)
    dist = np.zeros(5)
    ok = lib().orc_plane_estimation(pts5, float(threshold), unit, ninv, dist)
    return bool(ok), unit, float(ninv[0]), dist


def associate(ref_R, ref_t, cur_world, cur_lidar, map_xyzi, max_sq=5.0, plane_thr=0.1, sigma=0.1,
              scalar_thr=0.1, sigma_gate=4.5, knn_mode=0, threads=1):
    cur_world = _c(cur_world, np.float32)
    cur_lidar = _c(cur_lidar, np.float32)
    map_xyzi = _c(map_xyzi, np.float32)
    q = len(cur_world)
    m = len(map_xyzi)
    if m == 0:
        map_xyzi = np.zeros((1, 4), np.float32)
    out = dict(type=np.zeros(q, np.int8), covered=np.zeros(q, np.uint8), nn_idx=np.zeros((q, 5), np.int32),
               nn_d2=np.zeros((q, 5), np.float32), normal=np.zeros((q, 3)), d=np.zeros(q))
    nf = np.zeros(1, np.int32)
    nc = np.zeros(1, np.int32)
    lib().orc_associate(_c(ref_R, np.float64).reshape(-1), _c(ref_t, np.float64), cur_world, cur_lidar, q,
                        map_xyzi, m, max_sq, plane_thr, sigma, scalar_thr, sigma_gate, knn_mode, threads,
                        out["type"], out["covered"], out["nn_idx"].reshape(-1), out["nn_d2"].reshape(-1),
                        out["normal"].reshape(-1), out["d"], nf, nc)
    out["num_features"] = int(nf[0])
    out["num_covered"] = int(nc[0])
    return out


def lidar_factor_evaluate(point_xyz, normal, d, std, motion, pose_ref, pose_cur, ext, td, want_jac=(1, 1, 1, 1)):
    """LidarFactor::Evaluate for one residual with the Ceres calling convention."""
    dp = C.POINTER(C.c_double)
    params = [_c(pose_ref, np.float64), _c(pose_cur, np.float64), _c(ext, np.float64),
              np.array([td], np.float64)]
    par = (dp * 4)(*[p.ctypes.data_as(dp) for p in params])
    res = np.zeros(1)
    jacs = [np.full(s, np.nan) for s in (7, 7, 7, 1)]
    if want_jac is None:
        jp = None
    else:
        jp = (dp * 4)(*[(j.ctypes.data_as(dp) if w else dp()) for j, w in zip(jacs, want_jac)])
    lib().orc_lidar_factor_evaluate(_c(point_xyz, np.float32), _c(normal, np.float64), float(d), float(std),
                                    _c(motion, np.float64), par, res.ctypes.data_as(dp), jp)
    return float(res[0]), jacs


def huber_correct(delta, r, J22=None):
    r = np.array([r], np.float64)
    J = None if J22 is None else np.array(J22, np.float64)
    rho0 = lib().orc_huber_correct(float(delta), r, None if J is None else J.ctypes.data)
    return float(r[0]), J, rho0


def factor_batch(point_xyz, normal, d, std, valid, motion, pose_ref, pose_cur, ext, td, flags=0, huber_delta=1.0,
                 reduce=True, threads=1):
    point_xyz = _c(point_xyz, np.float32).reshape(-1, 3)
    n = len(point_xyz)
    r = np.zeros(n)
    J = np.zeros((n, 22))
    H = np.zeros((19, 19))
    b = np.zeros(19)
    cost = np.zeros(2)
    v = None if valid is None else _c(valid, np.uint8)
    lib().orc_factor_batch(n, point_xyz, _c(normal, np.float64), _c(d, np.float64), _c(std, np.float64),
                           None if v is None else v.ctypes.data, _c(motion, np.float64), _c(pose_ref, np.float64),
                           _c(pose_cur, np.float64), _c(ext, np.float64), float(td), int(flags), float(huber_delta),
                           r.ctypes.data, J.ctypes.data, H.ctypes.data if reduce else None,
                           b.ctypes.data if reduce else None, cost.ctypes.data if reduce else None, threads)
    return r, J, H, b, cost


def chi2(point_xyz, normal, d, std, valid, motion, pose_ref, pose_cur, ext, td, threshold=3.841, threads=1):
    point_xyz = _c(point_xyz, np.float32).reshape(-1, 3)
    n = len(point_xyz)
    out = np.zeros(n, np.uint8)
    v = None if valid is None else _c(valid, np.uint8)
    c = lib().orc_chi2(n, point_xyz, _c(normal, np.float64), _c(d, np.float64), _c(std, np.float64),
                       None if v is None else v.ctypes.data, _c(motion, np.float64), _c(pose_ref, np.float64),
                       _c(pose_cur, np.float64), _c(ext, np.float64), float(td), float(threshold), out, threads)
    return out, c


# ---- CPU-baseline forms: the reference's keyframe life cycle (tree built once per keyframe) and window-level calls ----
class MapTree:
    """Exact static kd-tree over one keyframe map, built once (LidarMap::addNewKeyFrame, lidar_map.cc:100-118)."""

    def __init__(self, map_xyzi):
        self.map = _c(map_xyzi, np.float32).reshape(-1, 4)
        if len(self.map) == 0:
            self.map = np.zeros((1, 4), np.float32)
            self.m = 0
        else:
            self.m = len(self.map)
        self.h = lib().orc_tree_build(self.map, self.m)

    def close(self):
        if self.h:
            lib().orc_tree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def associate_window(trees, ref_R, ref_t, cur_world, cur_lidar, max_sq=5.0, plane_thr=0.1, sigma=0.1, scalar_thr=0.1,
                     sigma_gate=4.5, threads=1):
    """updateLidarFeaturesInMap: every ref (MapTree) against the current keyframe in one parallel loop."""
    n = len(trees)
    cur_world = _c(cur_world, np.float32)
    cur_lidar = _c(cur_lidar, np.float32)
    q = len(cur_world)
    out = dict(type=np.zeros((n, q), np.int8), covered=np.zeros((n, q), np.uint8), nn_idx=np.zeros((n, q, 5), np.int32),
               nn_d2=np.zeros((n, q, 5), np.float32), normal=np.zeros((n, q, 3)), d=np.zeros((n, q)),
               num_features=np.zeros(n, np.int32), num_covered=np.zeros(n, np.int32))
    hs = (C.c_void_p * n)(*[t.h for t in trees])
    ms = (C.c_void_p * n)(*[t.map.ctypes.data for t in trees])
    m = np.array([t.m for t in trees], np.int32)
    R = _c(np.stack([np.asarray(r).reshape(9) for r in ref_R]), np.float64).reshape(-1)
    T = _c(np.stack([np.asarray(t).reshape(3) for t in ref_t]), np.float64).reshape(-1)
    lib().orc_associate_window(n, hs, ms, m, R, T, cur_world.reshape(-1), cur_lidar.reshape(-1), q, max_sq, plane_thr, sigma,
                               scalar_thr, sigma_gate, threads, out["type"].reshape(-1), out["covered"].reshape(-1),
                               out["nn_idx"].reshape(-1), out["nn_d2"].reshape(-1), out["normal"].reshape(-1),
                               out["d"].reshape(-1), out["num_features"], out["num_covered"])
    return out


def window_eval(point_xyz, normal, d, std, valid, motions, pose_refs, pose_cur, ext, td, flags=0, huber_delta=1.0, threads=1,
                out=None):
    """All pairs of the window in one call; inputs [n, q(, 3)]; returns H [n,19,19], b [n,19], cost [n,2]."""
    n, q = d.shape
    if out is None:
        out = (np.zeros((n, 19, 19)), np.zeros((n, 19)), np.zeros((n, 2)))
    H, b, cost = out
    lib().orc_window_eval(n, q, point_xyz, normal.reshape(-1), d.reshape(-1), float(std), valid.reshape(-1),
                          motions.reshape(-1), pose_refs.reshape(-1), pose_cur, ext, float(td), int(flags),
                          float(huber_delta), H.reshape(-1), b.reshape(-1), cost.reshape(-1), threads)
    return H, b, cost


def window_chi2(point_xyz, normal, d, std, valid, motions, pose_refs, pose_cur, ext, td, outlier, threshold=3.841, threads=1):
    """outlier [n, q] uint8 is updated in place; returns the per-pair counts of newly culled residuals."""
    n, q = d.shape
    counts = np.zeros(n, np.int32)
    lib().orc_window_chi2(n, q, point_xyz, normal.reshape(-1), d.reshape(-1), float(std), valid.reshape(-1),
                          motions.reshape(-1), pose_refs.reshape(-1), pose_cur, ext, float(td), float(threshold),
                          outlier.reshape(-1), counts, threads)
    return counts


# ---- the reference's own ikd-Tree (compiled from /root/reference, oracle/Makefile `ref`) ----
def ikd_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libikd_ref.so"))


def _ikd():
    global _IKD
    if _IKD is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libikd_ref.so"))
        L.ikdref_build.argtypes = [f32p, C.c_int, C.c_double]
        L.ikdref_build.restype = C.c_void_p
        L.ikdref_search.argtypes = [C.c_void_p, f32p, C.c_int, C.c_double, f32p, f32p]
        L.ikdref_free.argtypes = [C.c_void_p]
        _IKD = L
    return _IKD


class IkdTreeRef:
    """KD_TREE<PointXYZI>::build + nearestSearch exactly as FF-LINS calls them."""

    def __init__(self, map_xyzi, downsample=0.5):
        self.map = _c(map_xyzi, np.float32)
        self.h = _ikd().ikdref_build(self.map, len(self.map), downsample)

    def search(self, query_xyz, k=5, max_dist=5.0):
        pts = np.zeros((k, 4), np.float32)
        d2 = np.zeros(k, np.float32)
        n = _ikd().ikdref_search(self.h, _c(query_xyz, np.float32), k, max_dist, pts, d2)
        return n, pts[:n], d2[:n]

    def close(self):
        if self.h:
            _ikd().ikdref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
