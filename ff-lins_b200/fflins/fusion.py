"""The LiDAR-inertial loop of LINS::runFusion (ff_lins/ff_lins/ff_lins.cc:166-300) in the reduced form the scope allows:
raw IMU increments -> INS mechanization -> INS window -> per-frame undistortion (hot path) -> keyframe time nodes with
IMU preintegration -> window optimisation (IMU factors + LiDAR blocks + marginalization prior) -> poses written back,
INS window re-mechanized from the optimised state (MISC::redoInsMechanization) -> marginalization when the window is
full. Everything LiDAR goes through a `backend` object, so the same loop runs on the GPU library (GpuLidar below) and on
the CPU oracle (oracle/cpu_lidar.py, tests only); the host side (fflins.host) is identical in both."""
import ctypes as C

import numpy as np

from . import api, host
from .window_solver import lidar_pose


class GpuLidar:
    """LiDAR backend over the C-ABI (include/fflins.h)."""

    def __init__(self, seq, window=10):
        self.seq = seq
        self.ctx = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=window))
        L = self.ctx.L
        # the solver calls the library directly: addresses of fflins_window_eval / fflins_window_chi2 + the context handle
        self.native = (C.cast(L.fflins_window_eval, host.EVAL_FN), C.cast(L.fflins_window_chi2, host.CHI2_FN), self.ctx.h)

    def process_frame(self, fid, fr, ins_t, ins_p, ins_q):
        s = self.seq
        return self.ctx.frame_process(fid, fr["xyzi"], fr["time"], ins_t, ins_p, ins_q, s.R_bl, s.t_bl, fr["stamp"], fr["td"])

    def make_keyframe(self, fid, nonkeys, v, omega, td):
        c = self.ctx
        c.keyframe_merge(fid, nonkeys)
        for nk in nonkeys:
            c.release(nk)
        c.keyframe_add(fid)
        c.set_motion(fid, v, omega, td)

    def associate(self):
        st = self.ctx.associate()
        return int(sum(st.num_features[:st.n_refs]))

    def callbacks(self):
        return dict(native=self.native)

    def set_poses(self, ids, poses):
        self.ctx.reproject_window(ids, poses)

    def remove_oldest(self):
        return self.ctx.keyframe_remove_oldest()

    def close(self):
        self.ctx.close()


class FusionLoop:
    def __init__(self, seq, backend, frames_per_key=3, window=5, imu_rng=None, arw=0.0, vrw=0.0, init_error=(0.0, 0.0)):
        self.seq, self.lidar = seq, backend
        self.K, self.window = frames_per_key, window
        self.noise = host.imu_noise()
        self.insw = host.InsWindow(host.ins_config())
        self.solver = host.Solver(window_size=window, estimate_extrinsic=0, estimate_td=0)
        self.solver.set_extrinsic(seq.ext7(), 0.0)
        self.imu_dt = 1.0 / seq.imu_hz
        self.k_imu = None           # last IMU epoch fed
        self.imu_rng, self.arw, self.vrw = imu_rng, arw, vrw
        self.init_error = init_error
        self.nonkeys, self.key_ids, self.traj, self.stats = [], [], [], []
        self.last_node_time = None

    def _feed_imu(self, t_rel_end):
        from .synth import GPS_T0
        k1 = int(np.ceil(t_rel_end / self.imu_dt))
        if self.k_imu is None:
            k0 = k1 - 400
            t, dt, dth, dv = self.seq.imu_samples(k0, k1, gravity=host.GRAVITY, rng=self.imu_rng, arw=self.arw, vrw=self.vrw)
            p, q, v = self.seq.body_state(t[0] - GPS_T0)
            st = host.NavState.make(t[0], p + np.array([self.init_error[0], 0, 0]), q, v + np.array([self.init_error[1], 0, 0]))
            self.insw.reset(host.Imu.make(t[0], dt[0], dth[0], dv[0]), st)
            rng = range(1, len(t))
        else:
            if k1 <= self.k_imu:
                return
            t, dt, dth, dv = self.seq.imu_samples(self.k_imu + 1, k1, gravity=host.GRAVITY, rng=self.imu_rng, arw=self.arw,
                                                  vrw=self.vrw)
            rng = range(len(t))
        for i in rng:
            self.insw.add(host.Imu.make(t[i], dt[i], dth[i], dv[i]))
        self.k_imu = k1

    def step(self, fid, fr):
        """One LiDAR frame (linsLidarProcessing, ff_lins.cc:349-367) and, on keyframes, the window update."""
        from .synth import GPS_T0
        self._feed_imu(fr["stamp"] - GPS_T0 + 5 * self.imu_dt)      # waitImuDoInsMechanization: end time + 5 IMU epochs
        ins_t, ins_p, ins_q = self.insw.export(1000)
        self.lidar.process_frame(fid, fr, ins_t, ins_p, ins_q)
        if (len(self.nonkeys) + 1) < self.K:
            self.nonkeys.append(fid)
            return False
        stamp = fr["stamp"]
        # ---- time node (addNewLidarFrameTimeNode / addNewTimeNode, ff_lins.cc:458-526)
        if self.solver.num_nodes() == 0:
            # the first keyframe's node: state of the INS window at the stamp (initialisation stands in for the alignment)
            st = self._ins_state_at(stamp)
            self.solver.set_first_node(stamp, st, (0.02,) * 3 + (0.003,) * 3, (0.05,) * 3 + (2e-4,) * 3 + (2e-3,) * 3)
            self.solver.set_node_lidar(0, fid)
            omega = self.seq.body_rate(np.array([stamp - GPS_T0]))[0]
            v = np.array(st.v[:])
        else:
            series = self.insw.series(self.last_node_time, stamp)
            prev, _ = self.solver.node(self.solver.num_nodes() - 1)
            pre = host.Preintegration(self.noise, series[0], prev)
            for m in series[1:]:
                pre.add(m)
            cur = pre.current_state()
            omega = pre.frame_omega(0.1, self.imu_dt, np.array(cur.bg[:]))
            v = np.array(cur.v[:])
            self.solver.add_node(stamp, pre, fid)
        self.last_node_time = stamp
        self.lidar.make_keyframe(fid, self.nonkeys, v, omega, fr["td"])
        self.nonkeys = []
        self.key_ids.append(fid)
        n_feat = outl = 0
        if len(self.key_ids) >= 2:
            n_feat = self.lidar.associate()
            sm = self.solver.optimize(**self.lidar.callbacks())
            outl = sm.outliers
            self._write_back()
        st, _ = self.solver.node(self.solver.num_nodes() - 1)
        self.traj.append(np.array(st.p[:]))
        self.stats.append((n_feat, outl))
        # ---- marginalization when the window is full (ff_lins.cc:262-283)
        if len(self.key_ids) > self.window:
            cb = self.lidar.callbacks()
            self.solver.marginalize(**({"native": cb["native"]} if "native" in cb else {"evaluate": cb["evaluate"]}))
            self.lidar.remove_oldest()
            self.key_ids.pop(0)
        return True

    def _ins_state_at(self, stamp):
        """State of the INS window at a time node (linear position / slerp-free small-step interpolation is enough at 200 Hz)."""
        t, p, q = self.insw.export(1000)
        i = int(np.searchsorted(t, stamp))
        i = min(max(i, 1), len(t) - 1)
        s = (stamp - t[i - 1]) / (t[i] - t[i - 1])
        pos = p[i - 1] + s * (p[i] - p[i - 1])
        qq = q[i - 1] + s * (q[i] - q[i - 1])
        qq /= np.linalg.norm(qq)
        vel = (p[i] - p[i - 1]) / (t[i] - t[i - 1])
        return host.NavState.make(stamp, pos, qq, vel)

    def _write_back(self):
        """updateParametersFromOptimizer (optimizer.cc:187-219) + redoInsMechanization from the newest node."""
        ext7, _ = self.solver.get_extrinsic()
        ids, poses = [], []
        for i in range(self.solver.num_nodes()):
            st, lid = self.solver.node(i)
            if lid is None:
                continue
            R, t = lidar_pose(st.pose7(), ext7)
            ids.append(lid)
            poses.append((R, t))
        self.lidar.set_poses(ids, poses)
        newest, _ = self.solver.node(self.solver.num_nodes() - 1)
        self.insw.redo(newest, 400)
