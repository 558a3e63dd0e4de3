"""ctypes binding of libfflins_host.so (include/fflins_host.h): INS mechanization / INS window, IMU preintegration,
the sliding-window solver and the converters / writers — the CPU side either side of the LiDAR hot path."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "libfflins_host.so")
_LIB = None

EXPORTS = [
    "fflins_ins_mechanization", "fflins_imu_need_interpolation", "fflins_imu_interpolation", "fflins_insw_create",
    "fflins_insw_destroy", "fflins_insw_reset", "fflins_insw_add_imu", "fflins_insw_size", "fflins_insw_back",
    "fflins_insw_export", "fflins_insw_redo", "fflins_insw_imu_series", "fflins_preint_create", "fflins_preint_destroy",
    "fflins_preint_add_imu", "fflins_preint_reintegrate", "fflins_preint_current_state", "fflins_preint_delta_state",
    "fflins_preint_delta_time", "fflins_preint_evaluate", "fflins_preint_matrices", "fflins_preint_frame_omega",
    "fflins_solver_options_default", "fflins_solver_create", "fflins_solver_destroy", "fflins_solver_set_extrinsic",
    "fflins_solver_get_extrinsic", "fflins_solver_set_first_node", "fflins_solver_add_node", "fflins_solver_num_nodes",
    "fflins_solver_get_node", "fflins_solver_set_node_lidar", "fflins_solver_set_node_state", "fflins_solver_optimize", "fflins_solver_marginalize", "fflins_solver_prior",
    "fflins_convert_livox", "fflins_convert_generic", "fflins_format_trajectory", "fflins_format_imu_error",
    "fflins_format_statistics",
]


class Imu(C.Structure):
    _fields_ = [("time", C.c_double), ("dt", C.c_double), ("dtheta", C.c_double * 3), ("dvel", C.c_double * 3),
                ("odovel", C.c_double)]

    @staticmethod
    def make(time, dt, dtheta, dvel):
        m = Imu()
        m.time, m.dt = float(time), float(dt)
        m.dtheta[:] = [float(x) for x in dtheta]
        m.dvel[:] = [float(x) for x in dvel]
        return m


class NavState(C.Structure):
    _fields_ = [("time", C.c_double), ("p", C.c_double * 3), ("q", C.c_double * 4), ("v", C.c_double * 3),
                ("bg", C.c_double * 3), ("ba", C.c_double * 3)]

    @staticmethod
    def make(time, p, q, v, bg=(0, 0, 0), ba=(0, 0, 0)):
        s = NavState()
        s.time = float(time)
        s.p[:] = [float(x) for x in p]
        s.q[:] = [float(x) for x in q]
        s.v[:] = [float(x) for x in v]
        s.bg[:] = [float(x) for x in bg]
        s.ba[:] = [float(x) for x in ba]
        return s

    def pose7(self):
        return np.array(list(self.p) + list(self.q))

    def mix9(self):
        return np.array(list(self.v) + list(self.bg) + list(self.ba))


class InsConfig(C.Structure):
    _fields_ = [("gravity", C.c_double * 3)]


class ImuNoise(C.Structure):
    _fields_ = [("acc_vrw", C.c_double), ("gyr_arw", C.c_double), ("gyr_bias_std", C.c_double),
                ("acc_bias_std", C.c_double), ("corr_time", C.c_double), ("gravity", C.c_double)]


class SolverOptions(C.Structure):
    _fields_ = [("first_iterations", C.c_int32), ("second_iterations", C.c_int32), ("chi2_threshold", C.c_double),
                ("estimate_extrinsic", C.c_int32), ("estimate_td", C.c_int32), ("window_size", C.c_int32),
                ("min_speed_estimate_td", C.c_double)]


class SolverSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32 * 2), ("successful", C.c_int32 * 2), ("outliers", C.c_int32),
                ("cost_initial", C.c_double), ("cost_final", C.c_double)]


class LivoxPoint(C.Structure):
    _fields_ = [("offset_time", C.c_uint32), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float),
                ("reflectivity", C.c_uint8), ("tag", C.c_uint8), ("line", C.c_uint8)]


EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.POINTER(C.c_int32))
CHI2_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_double, C.POINTER(C.c_int32))


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not built: run `make -C ff-lins_b200`")
        L = C.CDLL(LIB_PATH)
        for name in ("fflins_insw_create", "fflins_preint_create", "fflins_solver_create"):
            getattr(L, name).restype = C.c_void_p
        L.fflins_preint_delta_time.restype = C.c_double
        for name in ("fflins_insw_destroy", "fflins_preint_destroy", "fflins_solver_destroy"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


GRAVITY = -9.80   # z-up navigation frame


def ins_config(g=GRAVITY):
    c = InsConfig()
    c.gravity[:] = [0.0, 0.0, g]
    return c


def imu_noise(acc_vrw=0.02, gyr_arw=0.002, gyr_bias_std=1e-4, acc_bias_std=1e-3, corr_time=3600.0, g=GRAVITY):
    return ImuNoise(acc_vrw, gyr_arw, gyr_bias_std, acc_bias_std, corr_time, g)


def mechanization(cfg, pre, cur, state):
    load().fflins_ins_mechanization(C.byref(cfg), C.byref(pre), C.byref(cur), C.byref(state))
    return state


class InsWindow:
    def __init__(self, cfg=None):
        self.L = load()
        self.cfg = cfg or ins_config()
        self.h = C.c_void_p(self.L.fflins_insw_create(C.byref(self.cfg)))

    def close(self):
        if self.h:
            self.L.fflins_insw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, imu0, state0):
        self.L.fflins_insw_reset(self.h, C.byref(imu0), C.byref(state0))

    def add(self, imu):
        self.L.fflins_insw_add_imu(self.h, C.byref(imu))

    def __len__(self):
        return self.L.fflins_insw_size(self.h)

    def back(self):
        imu, st = Imu(), NavState()
        self.L.fflins_insw_back(self.h, C.byref(imu), C.byref(st))
        return imu, st

    def export(self, cap=None):
        n = cap or len(self)
        t, p, q = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 4))
        k = self.L.fflins_insw_export(self.h, _dp(t), _dp(p), _dp(q), n)
        return t[:k], p[:k], q[:k]

    def redo(self, state, reserved):
        return self.L.fflins_insw_redo(self.h, C.byref(state), int(reserved))

    def series(self, start, end, cap=4096):
        buf = (Imu * cap)()
        n = self.L.fflins_insw_imu_series(self.h, C.c_double(start), C.c_double(end), buf, cap)
        if n < 0:
            raise RuntimeError("getImuSeriesFromTo failed")
        return [buf[i] for i in range(n)]


class Preintegration:
    def __init__(self, noise, imu0, state, own=True):
        self.L = load()
        self.noise = noise
        self.h = C.c_void_p(self.L.fflins_preint_create(C.byref(noise), C.byref(imu0), C.byref(state)))
        self.own = own

    def release(self):
        """Hand the native object over (fflins_solver_add_node takes ownership)."""
        self.own = False
        return self.h

    def __del__(self):
        try:
            if self.own and self.h:
                self.L.fflins_preint_destroy(self.h)
        except Exception:
            pass

    def add(self, imu):
        self.L.fflins_preint_add_imu(self.h, C.byref(imu))

    def reintegrate(self, state):
        self.L.fflins_preint_reintegrate(self.h, C.byref(state))

    def current_state(self):
        s = NavState()
        self.L.fflins_preint_current_state(self.h, C.byref(s))
        return s

    def delta_state(self):
        s = NavState()
        self.L.fflins_preint_delta_state(self.h, C.byref(s))
        return s

    def delta_time(self):
        return self.L.fflins_preint_delta_time(self.h)

    def evaluate(self, pose0, mix0, pose1, mix1, jac=True):
        a = [np.ascontiguousarray(x, np.float64) for x in (pose0, mix0, pose1, mix1)]
        r = np.zeros(15)
        J = [np.zeros((15, 7)), np.zeros((15, 9)), np.zeros((15, 7)), np.zeros((15, 9))] if jac else [None] * 4
        self.L.fflins_preint_evaluate(self.h, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(r),
                                      *[(_dp(j) if j is not None else None) for j in J])
        return r, J

    def matrices(self):
        cov, jac = np.zeros((15, 15)), np.zeros((15, 15))
        self.L.fflins_preint_matrices(self.h, _dp(cov), _dp(jac))
        return cov, jac

    def frame_omega(self, lidar_frame_dt, imu_dt, bg):
        bg = np.ascontiguousarray(bg, np.float64)
        out = np.zeros(3)
        self.L.fflins_preint_frame_omega(self.h, C.c_double(lidar_frame_dt), C.c_double(imu_dt), _dp(bg), _dp(out))
        return out


class Solver:
    """Optimizer::optimization / marginalization over IMU factors + the caller's LiDAR blocks."""

    def __init__(self, **kw):
        self.L = load()
        self.opt = SolverOptions()
        self.L.fflins_solver_options_default(C.byref(self.opt))
        for k, v in kw.items():
            setattr(self.opt, k, v)
        self.h = C.c_void_p(self.L.fflins_solver_create(C.byref(self.opt)))
        self._cb = None

    def close(self):
        if self.h:
            self.L.fflins_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_extrinsic(self, ext7, td):
        e = np.ascontiguousarray(ext7, np.float64)
        self.L.fflins_solver_set_extrinsic(self.h, _dp(e), C.c_double(td))

    def get_extrinsic(self):
        e, td = np.zeros(7), C.c_double()
        self.L.fflins_solver_get_extrinsic(self.h, _dp(e), C.byref(td))
        return e, td.value

    def set_first_node(self, time, state, pose_std=None, mix_std=None):
        ps = np.ascontiguousarray(pose_std, np.float64) if pose_std is not None else None
        ms = np.ascontiguousarray(mix_std, np.float64) if mix_std is not None else None
        self.L.fflins_solver_set_first_node(self.h, C.c_double(time), C.byref(state), _dp(ps) if ps is not None else None,
                                            _dp(ms) if ms is not None else None)

    def add_node(self, time, preint, lidar_id=None):
        lid = C.c_uint64(0xFFFFFFFFFFFFFFFF if lidar_id is None else lidar_id)
        return self.L.fflins_solver_add_node(self.h, C.c_double(time), preint.release(), lid)

    def set_node_lidar(self, index, lidar_id):
        return self.L.fflins_solver_set_node_lidar(self.h, index, C.c_uint64(lidar_id))

    def set_node_state(self, index, state):
        return self.L.fflins_solver_set_node_state(self.h, index, C.byref(state))

    def num_nodes(self):
        return self.L.fflins_solver_num_nodes(self.h)

    def node(self, i):
        s, lid = NavState(), C.c_uint64()
        if self.L.fflins_solver_get_node(self.h, i, C.byref(s), C.byref(lid)) != 0:
            raise IndexError(i)
        return s, (None if lid.value == 0xFFFFFFFFFFFFFFFF else lid.value)

    @staticmethod
    def wrap_callbacks(evaluate, cull):
        """evaluate(ref_ids, pose_refs[n,7], pose_cur, ext, td, flags) -> (H[n,19,19], b[n,19], cost[n,2]);
        cull(ref_ids, pose_refs, pose_cur, ext, td, threshold) -> counts[n]."""
        def _eval(user, n, ids, prefs, pcur, ext, td, flags, H, b, cost, nres):
            try:
                ref_ids = [ids[i] for i in range(n)]
                pr = np.ctypeslib.as_array(prefs, (n, 7)).copy()
                Hh, bb, cc = evaluate(ref_ids, pr, np.ctypeslib.as_array(pcur, (7,)).copy(),
                                      np.ctypeslib.as_array(ext, (7,)).copy(), td, flags)
                np.ctypeslib.as_array(H, (n, 19, 19))[:] = Hh
                np.ctypeslib.as_array(b, (n, 19))[:] = bb
                np.ctypeslib.as_array(cost, (n, 2))[:] = cc
                return 0
            except Exception as ex:   # never let an exception cross the C boundary
                print("lidar evaluate callback failed:", ex)
                return -1

        def _cull(user, n, ids, prefs, pcur, ext, td, thr, counts):
            try:
                ref_ids = [ids[i] for i in range(n)]
                pr = np.ctypeslib.as_array(prefs, (n, 7)).copy()
                cnt = cull(ref_ids, pr, np.ctypeslib.as_array(pcur, (7,)).copy(), np.ctypeslib.as_array(ext, (7,)).copy(),
                           td, thr)
                np.ctypeslib.as_array(counts, (n,))[:] = cnt
                return 0
            except Exception as ex:
                print("lidar chi2 callback failed:", ex)
                return -1
        return EVAL_FN(_eval), (CHI2_FN(_cull) if cull is not None else C.cast(None, CHI2_FN))

    def optimize(self, evaluate=None, cull=None, native=None):
        """native = (eval_fn_ptr, chi2_fn_ptr, user): e.g. the addresses of fflins_window_eval / fflins_window_chi2 and the
        context handle — the solver then calls the GPU library directly, no Python in the loop."""
        sm = SolverSummary()
        if native is not None:
            ev, ch, user = native
            rc = self.L.fflins_solver_optimize(self.h, ev, ch, user, C.byref(sm))
        elif evaluate is not None:
            ev, ch = self.wrap_callbacks(evaluate, cull)
            self._cb = (ev, ch)
            rc = self.L.fflins_solver_optimize(self.h, ev, ch, None, C.byref(sm))
        else:
            rc = self.L.fflins_solver_optimize(self.h, C.cast(None, EVAL_FN), C.cast(None, CHI2_FN), None, C.byref(sm))
        if rc != 0:
            raise RuntimeError(f"fflins_solver_optimize failed ({rc})")
        return sm

    def marginalize(self, evaluate=None, native=None):
        if native is not None:
            rc = self.L.fflins_solver_marginalize(self.h, native[0], native[2])
        elif evaluate is not None:
            ev, _ = self.wrap_callbacks(evaluate, None)
            self._cb = (ev,)
            rc = self.L.fflins_solver_marginalize(self.h, ev, None)
        else:
            rc = self.L.fflins_solver_marginalize(self.h, C.cast(None, EVAL_FN), None)
        if rc != 0:
            raise RuntimeError(f"fflins_solver_marginalize failed ({rc})")

    def prior(self):
        r, c = C.c_int(), C.c_int()
        self.L.fflins_solver_prior(self.h, C.byref(r), C.byref(c), None, None, 0)
        J, e = np.zeros((r.value, c.value)), np.zeros(r.value)
        if r.value:
            self.L.fflins_solver_prior(self.h, C.byref(r), C.byref(c), _dp(J), _dp(e), J.size)
        return J, e


REC32 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("pad", "<f4"), ("intensity", "<f4"), ("pad2", "<f4"),
                  ("time", "<f8")])


def convert_livox(points, timebase_ns, scan_line, min_range, max_range, to_gps_time=False):
    """points: ctypes array of LivoxPoint. Returns (records, start, end)."""
    n = len(points)
    out = np.zeros(max(n, 1), REC32)
    s, e = C.c_double(), C.c_double()
    m = load().fflins_convert_livox(points, n, C.c_uint64(timebase_ns), scan_line, C.c_double(min_range), C.c_double(max_range),
                                    int(to_gps_time), out.ctypes.data_as(C.c_void_p), n, C.byref(s), C.byref(e))
    if m < 0:
        raise RuntimeError("conversion failed")
    return out[:m].copy(), s.value, e.value


def convert_generic(xyzi, time, stamp, time_mode, min_range, max_range, to_gps_time=False):
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    time = np.ascontiguousarray(time, np.float64)
    n = len(xyzi)
    out = np.zeros(max(n, 1), REC32)
    s, e = C.c_double(), C.c_double()
    m = load().fflins_convert_generic(xyzi.ctypes.data_as(C.POINTER(C.c_float)), _dp(time), n, C.c_double(stamp), int(time_mode),
                                      C.c_double(min_range), C.c_double(max_range), int(to_gps_time),
                                      out.ctypes.data_as(C.c_void_p), n, C.byref(s), C.byref(e))
    if m < 0:
        raise RuntimeError("conversion failed")
    return out[:m].copy(), s.value, e.value


def format_trajectory(state):
    buf = C.create_string_buffer(512)
    n = load().fflins_format_trajectory(C.byref(state), buf, 512)
    return buf.raw[:n].decode()


def format_imu_error(state):
    buf = C.create_string_buffer(512)
    n = load().fflins_format_imu_error(C.byref(state), buf, 512)
    return buf.raw[:n].decode()


def format_statistics(values):
    v = np.ascontiguousarray(values, np.float64)
    buf = C.create_string_buffer(64 * len(v) + 16)
    n = load().fflins_format_statistics(_dp(v), len(v), buf, len(buf))
    return buf.raw[:n].decode()
