"""ctypes binding of libfflins_b200.so (include/fflins.h). Thin by design: every method is one
C-ABI call with numpy buffers; there is no CPU path — loading fails loudly if the CUDA library
is missing, and fflins_create fails without a device."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "libfflins_b200.so")
_LIB = None

CLOUD_LIDAR_FULL, CLOUD_LIDAR, CLOUD_WORLD, CLOUD_MAP_FULL, CLOUD_MAP = range(5)
HUBER, CONST_REF, CONST_CUR, CONST_EXT, CONST_TD = 1, 2, 4, 8, 16

EXPORTS = [
    "fflins_config_default", "fflins_create", "fflins_destroy", "fflins_last_error", "fflins_version",
    "fflins_synchronize", "fflins_host_register", "fflins_host_unregister", "fflins_frame_process",
    "fflins_frame_process_aos", "fflins_frame_process_dev", "fflins_frame_process_static", "fflins_frame_info_get",
    "fflins_frame_download", "fflins_frame_upload", "fflins_frame_set_pose", "fflins_frame_get_pose",
    "fflins_frame_set_motion", "fflins_frame_release", "fflins_reproject", "fflins_keyframe_selection",
    "fflins_voxel_downsample", "fflins_cloud_projection", "fflins_plane_estimation", "fflins_knn5",
    "fflins_keyframe_merge", "fflins_keyframe_add", "fflins_keyframe_remove_oldest", "fflins_keyframe_remove_newest",
    "fflins_keyframe_count", "fflins_keyframe_ids", "fflins_associate", "fflins_associate_pair",
    "fflins_features_download", "fflins_features_set_outlier", "fflins_factor_eval_batch", "fflins_factor_eval",
    "fflins_window_eval", "fflins_window_chi2", "fflins_profile_enable", "fflins_profile_reset", "fflins_profile_get",
    "fflins_launch_count", "fflins_timer_start", "fflins_timer_stop", "fflins_reproject_window",
    "fflins_eval_stats_get", "fflins_measure_fp64", "fflins_host_buffer_done", "fflins_host_prefetch", "fflins_solve_options_default",
    "fflins_window_solve",
]


class Config(C.Structure):
    _fields_ = [("point_to_plane_std", C.c_double), ("window_size", C.c_int32), ("point_filter_num", C.c_int32),
                ("downsample_size", C.c_double), ("frame_rate", C.c_double),
                ("plane_estimation_threshold", C.c_double), ("minimum_keyframe_translation", C.c_double),
                ("max_search_square_distance", C.c_double), ("max_keyframe_interval", C.c_double),
                ("min_keyframe_rotation", C.c_double), ("plane_scalar_threshold", C.c_double),
                ("plane_sigma_gate", C.c_double), ("huber_delta", C.c_double), ("max_points_per_frame", C.c_int32),
                ("max_merge_frames", C.c_int32)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3)]

    @staticmethod
    def make(R, t):
        p = Pose()
        p.R[:] = list(np.asarray(R, np.float64).reshape(-1))
        p.t[:] = list(np.asarray(t, np.float64).reshape(-1))
        return p

    def numpy(self):
        return np.array(self.R[:]).reshape(3, 3), np.array(self.t[:])


class SolveOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_int32 * 2), ("chi2_threshold", C.c_double), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double), ("initial_radius", C.c_double),
                ("flags", C.c_uint32), ("fixed_refs", C.c_uint32)]


class SolvePrior(C.Structure):
    _fields_ = [("H0", C.c_void_p), ("b0", C.c_void_p), ("x0", C.c_void_p), ("c0", C.c_double)]


class SolveReport(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32 * 2), ("successful", C.c_int32 * 2),
                ("evaluations", C.c_int32 * 2), ("cost_initial", C.c_double * 2), ("cost_final", C.c_double * 2),
                ("n_outliers", C.c_int32 * 16), ("n_res", C.c_int32 * 16), ("phase_cycles", C.c_double * 8)]


class FrameInfo(C.Structure):
    _fields_ = [("id", C.c_uint64), ("n_raw", C.c_int32), ("n_dec", C.c_int32), ("n_down", C.c_int32),
                ("n_map", C.c_int32), ("n_fallback", C.c_int32), ("reserved", C.c_int32), ("pose", Pose)]


class AssocStats(C.Structure):
    _fields_ = [("n_refs", C.c_int32), ("n_queries", C.c_int32), ("num_features", C.c_int32 * 16),
                ("num_covered", C.c_int32 * 16), ("ref_ids", C.c_uint64 * 16)]


class EvalStats(C.Structure):
    _fields_ = [("resident_launches", C.c_int64), ("resident_evals", C.c_int64), ("launched_evals", C.c_int64),
                ("resident_fallbacks", C.c_int64)]


class FflinsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fflins error {code}: {msg}")
        self.code = code


def load():
    """Load the CUDA library. No fallback: a missing .so is an error."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `make -C ff-lins_b200` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.fflins_last_error.restype = C.c_char_p
        L.fflins_version.restype = C.c_char_p
        L.fflins_launch_count.restype = C.c_int64
        L.fflins_destroy.restype = None
        L.fflins_config_default.restype = None
        L.fflins_solve_options_default.restype = None
        _LIB = L
    return _LIB


def default_config(**kw):
    cfg = Config()
    load().fflins_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _pose7(a):
    a = _c(a, np.float64).reshape(-1)
    return a


class Context:
    def __init__(self, cfg=None, device=0, **kw):
        self.L = load()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        h = C.c_void_p()
        rc = self.L.fflins_create(C.byref(self.cfg), device, C.byref(h))
        if rc != 0:
            raise FflinsError(rc, "fflins_create failed (no CUDA device?)" if rc == -6 else "fflins_create failed")
        self.h = h
        self._window_eval = self.L.fflins_window_eval   # bound once: the solver loop calls it 15 times per keyframe

    def close(self):
        if getattr(self, "h", None):
            self.L.fflins_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FflinsError(rc, self.L.fflins_last_error(self.h).decode())

    # ---- frames -----------------------------------------------------------------------------
    def frame_process(self, fid, xyzi, time, ins_t, ins_p, ins_q, R_bl, t_bl, stamp, td, sync=True):
        """sync=False: asynchronous mode (out = NULL), returns None; counts via frame_info() later."""
        xyzi = _c(xyzi, np.float32)
        time = _c(time, np.float64)
        ins_t, ins_p, ins_q = _c(ins_t, np.float64), _c(ins_p, np.float64), _c(ins_q, np.float64)
        info = FrameInfo() if sync else None
        pose = R_bl if isinstance(R_bl, Pose) else Pose.make(R_bl, t_bl)
        self._ck(self.L.fflins_frame_process(self.h, C.c_uint64(fid), _p(xyzi), _p(time), len(time), _p(ins_t),
                                             _p(ins_p), _p(ins_q), len(ins_t), C.byref(pose), C.c_double(stamp),
                                             C.c_double(td), C.byref(info) if sync else None))
        return info

    def frame_process_aos(self, fid, rec, ins_t, ins_p, ins_q, R_bl, t_bl, stamp, td, sync=True):
        assert rec.dtype.itemsize == 32
        ins_t, ins_p, ins_q = _c(ins_t, np.float64), _c(ins_p, np.float64), _c(ins_q, np.float64)
        info = FrameInfo() if sync else None
        pose = Pose.make(R_bl, t_bl)
        self._ck(self.L.fflins_frame_process_aos(self.h, C.c_uint64(fid), _p(rec), len(rec), _p(ins_t), _p(ins_p),
                                                 _p(ins_q), len(ins_t), C.byref(pose), C.c_double(stamp),
                                                 C.c_double(td), C.byref(info) if sync else None))
        return info

    def host_prefetch(self, rec):
        """Announce a raw frame (32-byte records, page-locked) ahead of frame_process_aos: its copy starts now."""
        assert rec.dtype.itemsize == 32 and rec.flags["C_CONTIGUOUS"]
        self._ck(self.L.fflins_host_prefetch(self.h, _p(rec), len(rec)))

    def frame_process_dev(self, fid, d_xyzi, d_time, n, ins_t, ins_p, ins_q, pose_bl, stamp, td, sync=True):
        """d_xyzi / d_time are raw device addresses (ints); the INS window arrays are host numpy arrays.
        sync=False: asynchronous mode (out = NULL), returns None."""
        info = FrameInfo() if sync else None
        self._ck(self.L.fflins_frame_process_dev(self.h, C.c_uint64(fid), C.c_void_p(d_xyzi), C.c_void_p(d_time), n,
                                                 _p(ins_t), _p(ins_p), _p(ins_q), len(ins_t), C.byref(pose_bl),
                                                 C.c_double(stamp), C.c_double(td), C.byref(info) if sync else None))
        return info

    # Argument tuples converted once per frame ("prepared" calls): what a C++ caller pays per call. The tuple keeps
    # the arrays alive; asynchronous mode only (out = NULL).
    @staticmethod
    def prepare_frame_dev(d_xyzi, d_time, n, ins_t, ins_p, ins_q, pose_bl, stamp, td):
        keep = (ins_t, ins_p, ins_q, pose_bl)
        return (C.c_void_p(d_xyzi), C.c_void_p(d_time), n, _p(ins_t), _p(ins_p), _p(ins_q), len(ins_t), C.byref(pose_bl),
                C.c_double(stamp), C.c_double(td), None), keep

    def frame_process_dev_prepared(self, fid, prep):
        self._ck(self.L.fflins_frame_process_dev(self.h, C.c_uint64(fid), *prep[0]))

    @staticmethod
    def prepare_frame(xyzi, time, ins_t, ins_p, ins_q, pose_bl, stamp, td):
        xyzi, time = _c(xyzi, np.float32), _c(time, np.float64)
        ins_t, ins_p, ins_q = _c(ins_t, np.float64), _c(ins_p, np.float64), _c(ins_q, np.float64)
        keep = (xyzi, time, ins_t, ins_p, ins_q, pose_bl)
        return (_p(xyzi), _p(time), len(time), _p(ins_t), _p(ins_p), _p(ins_q), len(ins_t), C.byref(pose_bl),
                C.c_double(stamp), C.c_double(td), None), keep

    def frame_process_prepared(self, fid, prep):
        self._ck(self.L.fflins_frame_process(self.h, C.c_uint64(fid), *prep[0]))

    @staticmethod
    def prepare_frame_aos(rec, ins_t, ins_p, ins_q, pose_bl, stamp, td):
        """rec: the reference's raw 32-byte PointXYZIT records (what LidarMap::lidarFrameProcessing receives)."""
        assert rec.dtype.itemsize == 32 and rec.flags["C_CONTIGUOUS"]
        ins_t, ins_p, ins_q = _c(ins_t, np.float64), _c(ins_p, np.float64), _c(ins_q, np.float64)
        keep = (rec, ins_t, ins_p, ins_q, pose_bl)
        return (_p(rec), len(rec), _p(ins_t), _p(ins_p), _p(ins_q), len(ins_t), C.byref(pose_bl), C.c_double(stamp),
                C.c_double(td), None), keep

    def frame_process_aos_prepared(self, fid, prep):
        self._ck(self.L.fflins_frame_process_aos(self.h, C.c_uint64(fid), *prep[0]))

    def frame_process_static(self, fid, xyzi, p, q, R_bl, t_bl):
        xyzi = _c(xyzi, np.float32)
        info = FrameInfo()
        pose = Pose.make(R_bl, t_bl)
        p, q = _c(p, np.float64), _c(q, np.float64)
        self._ck(self.L.fflins_frame_process_static(self.h, C.c_uint64(fid), _p(xyzi), len(xyzi), _p(p), _p(q),
                                                    C.byref(pose), C.byref(info)))
        return info

    def frame_info(self, fid):
        info = FrameInfo()
        self._ck(self.L.fflins_frame_info_get(self.h, C.c_uint64(fid), C.byref(info)))
        return info

    def download(self, fid, which):
        n = C.c_int32()
        self._ck(self.L.fflins_frame_download(self.h, C.c_uint64(fid), which, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.float32)
        if n.value:
            self._ck(self.L.fflins_frame_download(self.h, C.c_uint64(fid), which, _p(out), n.value, C.byref(n)))
        return out

    def upload(self, fid, which, xyzi):
        xyzi = _c(xyzi, np.float32).reshape(-1, 4)
        self._ck(self.L.fflins_frame_upload(self.h, C.c_uint64(fid), which, _p(xyzi), len(xyzi)))

    def set_pose(self, fid, R, t):
        pose = Pose.make(R, t)
        self._ck(self.L.fflins_frame_set_pose(self.h, C.c_uint64(fid), C.byref(pose)))

    def get_pose(self, fid):
        pose = Pose()
        self._ck(self.L.fflins_frame_get_pose(self.h, C.c_uint64(fid), C.byref(pose)))
        return pose.numpy()

    def set_motion(self, fid, v, omega, td):
        d3 = C.c_double * 3
        self._ck(self.L.fflins_frame_set_motion(self.h, C.c_uint64(fid), d3(v[0], v[1], v[2]),
                                                d3(omega[0], omega[1], omega[2]), C.c_double(td)))

    def release(self, fid):
        self._ck(self.L.fflins_frame_release(self.h, C.c_uint64(fid)))

    def reproject(self, fid, R, t):
        pose = Pose.make(R, t)
        self._ck(self.L.fflins_reproject(self.h, C.c_uint64(fid), C.byref(pose)))

    def reproject_window(self, ids, poses):
        """poses: ctypes array (Pose * n) or list of (R, t)."""
        n = len(ids)
        ids = (C.c_uint64 * n)(*ids)
        if not isinstance(poses, C.Array):
            arr = (Pose * n)()
            for i, (R, t) in enumerate(poses):
                arr[i] = Pose.make(R, t)
            poses = arr
        self._ck(self.L.fflins_reproject_window(self.h, n, ids, poses))

    def get_pose_struct(self, fid):
        pose = Pose()
        self._ck(self.L.fflins_frame_get_pose(self.h, C.c_uint64(fid), C.byref(pose)))
        return pose

    def keyframe_selection(self, fid, stamp, last_stamp):
        k = C.c_int()
        st = np.zeros(3)
        self._ck(self.L.fflins_keyframe_selection(self.h, C.c_uint64(fid), C.c_double(stamp), C.c_double(last_stamp),
                                                  C.byref(k), _p(st)))
        return bool(k.value), st

    # ---- stage-level ---------------------------------------------------------------------------
    def voxel_downsample(self, xyzi, leaf, want_keys=False):
        xyzi = _c(xyzi, np.float32).reshape(-1, 4)
        n = len(xyzi)
        out = np.zeros((max(n, 1), 4), np.float32)
        keys = np.zeros(max(n, 1), np.int32) if want_keys else None
        m = C.c_int32()
        self._ck(self.L.fflins_voxel_downsample(self.h, _p(xyzi), n, C.c_double(leaf), _p(out), n, C.byref(m),
                                                _p(keys)))
        if want_keys:
            return out[:m.value].copy(), keys[:n].copy()
        return out[:m.value].copy()

    def cloud_projection(self, R, t, src):
        src = _c(src, np.float32).reshape(-1, 4)
        dst = np.zeros_like(src)
        pose = Pose.make(R, t)
        self._ck(self.L.fflins_cloud_projection(self.h, C.byref(pose), _p(src), len(src), _p(dst)))
        return dst

    def plane_estimation(self, pts, threshold):
        pts = _c(pts, np.float32).reshape(-1, 5, 4)
        n = len(pts)
        unit, ninv, dist, ok = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 5)), np.zeros(n, np.uint8)
        self._ck(self.L.fflins_plane_estimation(self.h, _p(pts), n, C.c_double(threshold), _p(unit), _p(ninv),
                                                _p(dist), _p(ok)))
        return ok.astype(bool), unit, ninv, dist

    def knn5(self, map_xyzi, query_xyzi, max_dist=5.0):
        map_xyzi = _c(map_xyzi, np.float32).reshape(-1, 4)
        query_xyzi = _c(query_xyzi, np.float32).reshape(-1, 4)
        q = len(query_xyzi)
        idx, d2, cnt = np.zeros((q, 5), np.int32), np.zeros((q, 5), np.float32), np.zeros(q, np.int32)
        self._ck(self.L.fflins_knn5(self.h, _p(map_xyzi) if len(map_xyzi) else None, len(map_xyzi), _p(query_xyzi), q,
                                    C.c_double(max_dist), _p(idx), _p(d2), _p(cnt)))
        return cnt, idx, d2

    # ---- keyframes -------------------------------------------------------------------------------
    def keyframe_merge(self, key_id, nonkey_ids, sync=True):
        """sync=False: out = NULL, the map is built on the map stream and adopted at first use."""
        n = len(nonkey_ids)
        ids = (C.c_uint64 * n)(*nonkey_ids) if n else None
        info = FrameInfo() if sync else None
        self._ck(self.L.fflins_keyframe_merge(self.h, C.c_uint64(key_id), ids, n, C.byref(info) if sync else None))
        return info

    def keyframe_add(self, key_id):
        self._ck(self.L.fflins_keyframe_add(self.h, C.c_uint64(key_id)))

    def keyframe_remove_oldest(self):
        rid = C.c_uint64()
        self._ck(self.L.fflins_keyframe_remove_oldest(self.h, C.byref(rid)))
        return rid.value

    def keyframe_remove_newest(self):
        rid = C.c_uint64()
        self._ck(self.L.fflins_keyframe_remove_newest(self.h, C.byref(rid)))
        return rid.value

    def keyframe_ids(self):
        buf = self.__dict__.get("_ids_buf")
        if buf is None:
            buf = self.__dict__["_ids_buf"] = ((C.c_uint64 * 16)(), C.c_int32())
        ids, n = buf
        self._ck(self.L.fflins_keyframe_ids(self.h, ids, 16, C.byref(n)))
        return ids[:n.value]

    # ---- association --------------------------------------------------------------------------------
    def associate(self):
        st = AssocStats()
        self._ck(self.L.fflins_associate(self.h, C.byref(st)))
        return st

    def associate_pair(self, ref_id, cur_id):
        nf, nc = C.c_int32(), C.c_int32()
        self._ck(self.L.fflins_associate_pair(self.h, C.c_uint64(ref_id), C.c_uint64(cur_id), C.byref(nf),
                                              C.byref(nc)))
        return nf.value, nc.value

    def features(self, ref_id, nn_points=False):
        q = C.c_int32()
        self._ck(self.L.fflins_features_download(self.h, C.c_uint64(ref_id), 0, C.byref(q), None, None, None, None,
                                                 None, None, None, None))
        n = q.value
        out = dict(type=np.zeros(n, np.int8), covered=np.zeros(n, np.uint8), outlier=np.zeros(n, np.uint8),
                   nn_idx=np.zeros((n, 5), np.int32), nn_d2=np.zeros((n, 5), np.float32), normal=np.zeros((n, 3)),
                   d=np.zeros(n))
        pts = np.zeros((n, 5, 4), np.float32) if nn_points else None
        if n:
            self._ck(self.L.fflins_features_download(self.h, C.c_uint64(ref_id), n, C.byref(q), _p(out["type"]),
                                                     _p(out["covered"]), _p(out["outlier"]), _p(out["nn_idx"]),
                                                     _p(out["nn_d2"]), _p(pts), _p(out["normal"]), _p(out["d"])))
        if nn_points:
            out["nn_points"] = pts
        return out

    def set_outlier(self, ref_id, mask):
        mask = _c(mask, np.uint8)
        self._ck(self.L.fflins_features_set_outlier(self.h, C.c_uint64(ref_id), _p(mask), len(mask)))

    # ---- factors ---------------------------------------------------------------------------------------
    def factor_eval_batch(self, point, normal, d, std, motion, pose_ref, pose_cur, ext, td, flags=0, want_J=True):
        point = _c(point, np.float32).reshape(-1, 3)
        n = len(point)
        normal, d, std = _c(normal, np.float64), _c(d, np.float64), _c(std, np.float64)
        motion = _c(motion, np.float64)
        r = np.zeros(n)
        J = np.zeros((n, 22)) if want_J else None
        self._ck(self.L.fflins_factor_eval_batch(self.h, n, _p(point), _p(normal), _p(d), _p(std), _p(motion),
                                                 _p(_pose7(pose_ref)), _p(_pose7(pose_cur)), _p(_pose7(ext)),
                                                 C.c_double(td), C.c_uint32(flags), _p(r), _p(J)))
        return r, J

    def factor_eval(self, ref_id, cur_id, pose_ref, pose_cur, ext, td, flags=0, per_residual=True):
        q = self.frame_info(cur_id).n_down
        n_res = C.c_int32()
        valid = np.zeros(q, np.uint8) if per_residual else None
        r = np.zeros(q) if per_residual else None
        J = np.zeros((q, 22)) if per_residual else None
        H, b, cost = np.zeros((19, 19)), np.zeros(19), np.zeros(2)
        self._ck(self.L.fflins_factor_eval(self.h, C.c_uint64(ref_id), C.c_uint64(cur_id), _p(_pose7(pose_ref)),
                                           _p(_pose7(pose_cur)), _p(_pose7(ext)), C.c_double(td), C.c_uint32(flags),
                                           q, C.byref(n_res), _p(valid), _p(r), _p(J), _p(H), _p(b), _p(cost)))
        return dict(n_res=n_res.value, valid=valid, r=r, J=J, H=H, b=b, cost=cost)

    def window_eval(self, ref_ids, pose_refs, pose_cur, ext, td, flags=0, out=None):
        ids = np.array(list(ref_ids), np.uint64)
        n = len(ids)
        pose_refs = _c(pose_refs, np.float64).reshape(n, 7)
        if out is None:
            out = dict(H=np.zeros((n, 19, 19)), b=np.zeros((n, 19)), cost=np.zeros((n, 2)), n_res=np.zeros(n, np.int32))
        self._ck(self.L.fflins_window_eval(self.h, n, _p(ids), _p(pose_refs), _p(_pose7(pose_cur)), _p(_pose7(ext)),
                                           C.c_double(td), C.c_uint32(flags), _p(out["H"]), _p(out["b"]),
                                           _p(out["cost"]), _p(out["n_res"])))
        return out

    def window_eval_prepared(self, prep, td, flags, out):
        """Same call with the argument arrays converted once (`prepare_window`): what a C++ caller pays."""
        key = (td, flags)
        args = prep.get("_eval_args")
        if args is None or prep.get("_eval_key") != key:
            args = (self.h, prep["n"], prep["ids_p"], prep["refs_p"], prep["cur_p"], prep["ext_p"], C.c_double(td),
                    C.c_uint32(flags), out["H_p"], out["b_p"], out["cost_p"], out["n_res_p"])
            prep["_eval_args"], prep["_eval_key"] = args, key
        rc = self._window_eval(*args)
        if rc != 0:
            self._ck(rc)
        return out

    @staticmethod
    def prepare_window(ref_ids, pose_refs, pose_cur, ext, out=None):
        """out: a result dict of an earlier call with the same number of pairs may be reused (overwritten in place)."""
        ids = np.array(list(ref_ids), np.uint64)
        n = len(ids)
        refs = _c(pose_refs, np.float64).reshape(n, 7)
        cur, ex = _pose7(pose_cur), _pose7(ext)
        if out is None or len(out["n_res"]) != n:
            out = dict(H=np.zeros((n, 19, 19)), b=np.zeros((n, 19)), cost=np.zeros((n, 2)), n_res=np.zeros(n, np.int32))
            for k in ("H", "b", "cost", "n_res"):
                out[k + "_p"] = _p(out[k])
        prep = dict(n=n, ids=ids, refs=refs, cur=cur, ext=ex, ids_p=_p(ids), refs_p=_p(refs), cur_p=_p(cur), ext_p=_p(ex))
        return prep, out

    def window_chi2_prepared(self, prep, td, threshold=3.841):
        """prep from prepare_window; returns the per-pair outlier counts."""
        cnt = np.zeros(max(prep["n"], 1), np.int32)
        self._ck(self.L.fflins_window_chi2(self.h, prep["n"], prep["ids_p"], prep["refs_p"], prep["cur_p"], prep["ext_p"],
                                           C.c_double(td), C.c_double(threshold), _p(cnt)))
        return cnt[:prep["n"]]

    def solve_options(self, **kw):
        o = SolveOptions()
        self.L.fflins_solve_options_default(C.byref(o))
        for k, v in kw.items():
            if k == "max_iterations":
                o.max_iterations[0], o.max_iterations[1] = int(v[0]), int(v[1])
            else:
                setattr(o, k, v)
        return o

    def window_solve(self, ref_ids, pose_refs, pose_cur, ext, td, options=None, prior=None):
        """fflins_window_solve: the 5 + chi2 + 10 Levenberg-Marquardt schedule on the device, one launch.
        prior = (H0 [D, D], b0 [D], x0 [7 n + 15], c0) or None. Returns (pose_refs, pose_cur, ext, td, report)."""
        ids = np.array(list(ref_ids), np.uint64)
        n = len(ids)
        refs = np.array(_c(pose_refs, np.float64).reshape(n, 7), copy=True)
        cur, ex = np.array(_pose7(pose_cur), copy=True), np.array(_pose7(ext), copy=True)
        tdv = C.c_double(td)
        opt = options if options is not None else self.solve_options()
        rep = SolveReport()
        pr, keep = None, None
        if prior is not None:
            H0, b0, x0, c0 = prior
            keep = (_c(H0, np.float64), _c(b0, np.float64), _c(x0, np.float64))
            D = 6 * n + 13
            assert keep[0].shape == (D, D) and keep[1].shape == (D,) and keep[2].shape == (7 * n + 15,)
            pr = SolvePrior(keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, float(c0))
        self._ck(self.L.fflins_window_solve(self.h, n, _p(ids), _p(refs), _p(cur), _p(ex), C.byref(tdv), C.byref(opt),
                                            C.byref(pr) if pr is not None else None, C.byref(rep)))
        return refs, cur, ex, tdv.value, rep

    def window_chi2(self, ref_ids, pose_refs, pose_cur, ext, td, threshold=3.841):
        ids = np.array(list(ref_ids), np.uint64)
        n = len(ids)
        pose_refs = _c(pose_refs, np.float64).reshape(n, 7)
        cnt = np.zeros(max(n, 1), np.int32)
        self._ck(self.L.fflins_window_chi2(self.h, n, _p(ids), _p(pose_refs), _p(_pose7(pose_cur)), _p(_pose7(ext)),
                                           C.c_double(td), C.c_double(threshold), _p(cnt)))
        return cnt[:n]

    # ---- measurement --------------------------------------------------------------------------------------
    def synchronize(self):
        self._ck(self.L.fflins_synchronize(self.h))

    def profile_enable(self, on=True):
        self._ck(self.L.fflins_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._ck(self.L.fflins_profile_reset(self.h))

    def profile_get(self):
        n = C.c_int32()
        self._ck(self.L.fflins_profile_get(self.h, None, 0, None, None, 0, C.byref(n)))
        k = n.value
        if k == 0:
            return {}
        names = C.create_string_buffer(64 * k + 16)
        ms = np.zeros(k)
        cnt = np.zeros(k, np.int64)
        self._ck(self.L.fflins_profile_get(self.h, names, len(names), _p(ms), _p(cnt), k, C.byref(n)))
        nm = names.value.decode().strip().split("\n")
        return {nm[i]: (float(ms[i]), int(cnt[i])) for i in range(k)}

    def timer_start(self):
        self._ck(self.L.fflins_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.L.fflins_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.fflins_launch_count(self.h))

    def eval_stats(self):
        st = EvalStats()
        self._ck(self.L.fflins_eval_stats_get(self.h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in EvalStats._fields_}

    def measure_fp64(self):
        v = C.c_double()
        self._ck(self.L.fflins_measure_fp64(self.h, C.byref(v)))
        return v.value


def host_register(arr):
    rc = load().fflins_host_register(arr.ctypes.data_as(C.c_void_p), C.c_uint64(arr.nbytes))
    if rc != 0:
        raise FflinsError(rc, "cudaHostRegister failed")


def host_unregister(arr):
    load().fflins_host_unregister(arr.ctypes.data_as(C.c_void_p))
