"""The LiDAR call sequence of LINS::runFusion (ff_lins/ff_lins/ff_lins.cc:236-262) and of
Optimizer::optimization / marginalization (ff_lins/ff_lins/optimizer.cc:85-185, 221-404) as far as
they touch the hot path, driven through the C-ABI.

Per frame:      fflins_frame_process                        (lidarFrameProcessing)
Per keyframe:   fflins_keyframe_merge -> fflins_keyframe_add -> fflins_frame_set_motion
                -> fflins_associate -> n_eval x fflins_window_eval (+ fflins_window_chi2 between the
                two solves) -> fflins_reproject of every keyframe -> fflins_keyframe_remove_oldest
                when the window is full.
Keyframes are chosen every `frames_per_key` frames (the synthetic trajectory makes the reference's
own keyframeSelection fire at that cadence; the selection call is exercised separately)."""
import time

import numpy as np

from . import api


class _Timed:
    """Optional wall-clock trace of the host-side calls (FFLINS trace, off by default)."""

    def __init__(self, ctx, sink):
        self._ctx, self._sink = ctx, sink

    def __getattr__(self, name):
        fn = getattr(self._ctx, name)
        if not callable(fn):
            return fn

        def wrap(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            e = self._sink.setdefault(name, [0.0, 0])
            e[0] += time.perf_counter() - t0
            e[1] += 1
            return r
        return wrap


class Session:
    def __init__(self, ctx, seq, frames_per_key=5, n_eval=(5, 10), flags=api.HUBER, trace=None, sync_frames=True):
        self.ctx = ctx if trace is None else _Timed(ctx, trace)
        self.seq = seq
        self.K = frames_per_key
        self.sync_frames = sync_frames   # False: frames are enqueued asynchronously (out = NULL)
        self.n_eval = n_eval
        self.flags = flags
        self.nonkeys = []
        self.key_stamps = {}
        self.key_pose7 = {}
        self.key_pose = {}
        self.next_id = 0
        self.window = ctx.cfg.window_size
        self.pose_bl = api.Pose.make(seq.R_bl, seq.t_bl)
        self.ext7 = seq.ext7()
        self.last = {}

    # -- inputs are host numpy arrays (the e2e path) --------------------------------------------------
    def process_frame(self, fr):
        fid = self.next_id
        self.next_id += 1
        if not self.sync_frames and "aos" in fr:   # the reference-facing input: raw 32-byte PointXYZIT records
            prep = fr.get("_prep_aos")
            if prep is None:
                prep = fr["_prep_aos"] = api.Context.prepare_frame_aos(fr["aos"], fr["ins_t"], fr["ins_p"], fr["ins_q"],
                                                                       self.pose_bl, fr["stamp"], fr["td"])
            self.ctx.frame_process_aos_prepared(fid, prep)
            return fid, None
        if not self.sync_frames:   # asynchronous: the ctypes argument tuple is built once per frame dict
            prep = fr.get("_prep_host")
            if prep is None:
                prep = fr["_prep_host"] = api.Context.prepare_frame(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"],
                                                                    fr["ins_q"], self.pose_bl, fr["stamp"], fr["td"])
            self.ctx.frame_process_prepared(fid, prep)
            return fid, None
        info = self.ctx.frame_process(fid, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"],
                                      self.pose_bl, None, fr["stamp"], fr["td"], sync=True)
        return fid, info

    def process_frame_dev(self, fr_dev):
        """fr_dev: raw device pointers of the point data (resident in HBM) + host INS window + scalars."""
        fid = self.next_id
        self.next_id += 1
        if not self.sync_frames:
            prep = fr_dev.get("_prep_dev")
            if prep is None:
                prep = fr_dev["_prep_dev"] = api.Context.prepare_frame_dev(
                    fr_dev["xyzi"], fr_dev["time"], fr_dev["n"], fr_dev["ins_t"], fr_dev["ins_p"], fr_dev["ins_q"],
                    self.pose_bl, fr_dev["stamp"], fr_dev["td"])
            self.ctx.frame_process_dev_prepared(fid, prep)
            return fid, None
        info = self.ctx.frame_process_dev(fid, fr_dev["xyzi"], fr_dev["time"], fr_dev["n"], fr_dev["ins_t"],
                                          fr_dev["ins_p"], fr_dev["ins_q"], self.pose_bl, fr_dev["stamp"],
                                          fr_dev["td"], sync=True)
        return fid, info

    def after_frame(self, fid, fr, force_key=None):
        """Keyframe bookkeeping after a frame was processed. Returns True if it became a keyframe."""
        is_key = force_key if force_key is not None else (len(self.nonkeys) + 1 >= self.K)
        if not is_key:
            self.nonkeys.append(fid)
            return False
        ctx = self.ctx
        self.last["merge"] = ctx.keyframe_merge(fid, self.nonkeys, sync=self.sync_frames)
        for nk in self.nonkeys:
            ctx.release(nk)
        self.nonkeys = []
        ctx.keyframe_add(fid)
        ctx.set_motion(fid, fr["v"], fr["omega"], fr["td"])
        self.key_stamps[fid] = fr["stamp"]
        self.key_pose7[fid] = fr["pose7"] if "pose7" in fr else self.seq.imu_pose7(fr["stamp"])
        self.key_pose[fid] = ctx.get_pose_struct(fid)   # host-evaluated, available without a sync
        ids = ctx.keyframe_ids()
        if len(ids) >= 2:
            self.last["assoc"] = ctx.associate()
            refs = ids[:-1]
            # solver state: IMU pose blocks of the keyframes (precomputed per frame when available)
            pose_refs = np.array([self.key_pose7[r] for r in refs])
            pose_cur = self.key_pose7[fid]
            prep, out = api.Context.prepare_window(refs, pose_refs, pose_cur, self.ext7, out=self.last.get("eval"))
            # first solve: n_eval[0] LM iterations' worth of evaluations
            for _ in range(self.n_eval[0]):
                ctx.window_eval_prepared(prep, fr["td"], self.flags, out)
            self.last["chi2"] = ctx.window_chi2_prepared(prep, fr["td"])
            for _ in range(self.n_eval[1]):
                ctx.window_eval_prepared(prep, fr["td"], self.flags, out)
            self.last["eval"] = out
            self.last["eval_prep"] = (prep, out, fr["td"])
            # updateParametersFromOptimizer: re-project every keyframe (optimizer.cc:203-218)
            poses = (api.Pose * len(ids))(*[self.key_pose[k] for k in ids])
            ctx.reproject_window(ids, poses)
        if len(ids) > self.window:
            rid = ctx.keyframe_remove_oldest()
            self.key_stamps.pop(rid, None)
            self.key_pose7.pop(rid, None)
            self.key_pose.pop(rid, None)
        return True
