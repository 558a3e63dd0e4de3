"""The reference's per-frame / per-keyframe LiDAR call sequence executed by the CPU oracle
(OpenMP at the reference's tbb::parallel_for sites). TEST INFRASTRUCTURE / CPU BASELINE ONLY:
used by tests/ for end-to-end comparisons and by bench.py's cpu_baseline and --impl reference
legs. Mirrors ff-lins_b200/fflins/pipeline.py call for call."""
import time

import numpy as np

from . import pyoracle as O


class CpuSession:
    def __init__(self, seq, threads=None, frames_per_key=5, window=10, n_eval=(5, 10), knn_mode=1, leaf=0.5,
                 sigma=0.1, flags=1):
        self.seq = seq
        self.threads = threads or O.max_threads()
        self.K = frames_per_key
        self.window = window
        self.n_eval = n_eval
        self.knn_mode = knn_mode
        self.leaf = leaf
        self.sigma = sigma
        self.flags = flags
        self.nonkeys = []
        self.keys = []
        self.ext7 = seq.ext7()
        self.t = {}
        self.last = {}

    def _tick(self, name, t0):
        self.t[name] = self.t.get(name, 0.0) + (time.perf_counter() - t0)

    def process_frame(self, fr):
        s = self.seq
        t0 = time.perf_counter()
        und, R, t, fb = O.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], s.R_bl, s.t_bl,
                                    fr["stamp"], fr["td"], threads=self.threads)
        self._tick("undistort", t0)
        t0 = time.perf_counter()
        dec = O.decimate(und, s.filter_num)
        down = O.voxel_filter(dec, self.leaf)
        self._tick("voxel_frame", t0)
        t0 = time.perf_counter()
        world = O.project(R, t, down, threads=self.threads)
        self._tick("project_world", t0)
        return dict(und=und, dec=dec, down=down, world=world, R=R, t=t, fr=fr, fallback=fb)

    def after_frame(self, f, force_key=None):
        is_key = force_key if force_key is not None else (len(self.nonkeys) + 1 >= self.K)
        if not is_key:
            self.nonkeys.append(f)
            return False
        t0 = time.perf_counter()
        clouds = [f["und"]]
        for g in self.nonkeys:
            Rr, tr = O.relative_pose(f["R"], f["t"], g["R"], g["t"])
            clouds.append(O.project(Rr, tr, g["und"], threads=self.threads))
        f["map_full"] = np.concatenate(clouds)
        self._tick("merge_project", t0)
        t0 = time.perf_counter()
        f["map"] = O.voxel_filter(f["map_full"], self.leaf)
        self._tick("voxel_map", t0)
        self.nonkeys = []
        self.keys.append(f)
        if len(self.keys) >= 2:
            cur = self.keys[-1]
            refs = self.keys[:-1]
            t0 = time.perf_counter()
            feats = []
            for ref in refs:   # kd-tree build + search per ref (lidar_map.cc:100-118, 436-444)
                feats.append(O.associate(ref["R"], ref["t"], cur["world"], cur["down"], ref["map"], sigma=self.sigma,
                                         knn_mode=self.knn_mode, threads=self.threads))
            self._tick("associate", t0)
            for ref, ft in zip(refs, feats):
                ref["feat"] = ft
                ref["outlier"] = np.zeros(len(cur["down"]), np.uint8)
            s = self.seq
            pose_refs = [s.imu_pose7(r["fr"]["stamp"]) for r in refs]
            pose_cur = s.imu_pose7(cur["fr"]["stamp"])
            td = cur["fr"]["td"]
            mc = np.concatenate([cur["fr"]["v"], cur["fr"]["omega"], [cur["fr"]["td"]]])
            q = len(cur["down"])
            std = np.full(q, self.sigma)
            pts = np.ascontiguousarray(cur["down"][:, :3])

            def evaluate():
                out = []
                for ref, pr in zip(refs, pose_refs):
                    ft = ref["feat"]
                    valid = ((ft["type"] == 0) & (ref["outlier"] == 0)).astype(np.uint8)
                    motion = np.concatenate([ref["fr"]["v"], ref["fr"]["omega"], [ref["fr"]["td"]], mc])
                    out.append(O.factor_batch(pts, ft["normal"], ft["d"], std, valid, motion, pr, pose_cur, self.ext7,
                                              td, flags=self.flags, threads=self.threads))
                return out
            t0 = time.perf_counter()
            ev = None
            for _ in range(self.n_eval[0]):
                ev = evaluate()
            self._tick("factor_eval", t0)
            t0 = time.perf_counter()
            n_out = []
            for ref, pr in zip(refs, pose_refs):
                ft = ref["feat"]
                valid = ((ft["type"] == 0) & (ref["outlier"] == 0)).astype(np.uint8)
                motion = np.concatenate([ref["fr"]["v"], ref["fr"]["omega"], [ref["fr"]["td"]], mc])
                mask, c = O.chi2(pts, ft["normal"], ft["d"], std, valid, motion, pr, pose_cur, self.ext7, td,
                                 threads=self.threads)
                ref["outlier"] |= mask
                n_out.append(c)
            self._tick("chi2", t0)
            t0 = time.perf_counter()
            for _ in range(self.n_eval[1]):
                ev = evaluate()
            self._tick("factor_eval", t0)
            t0 = time.perf_counter()
            for k in self.keys:
                k["world"] = O.project(k["R"], k["t"], k["down"], threads=self.threads)
            self._tick("reproject", t0)
            self.last = dict(eval=ev, chi2=n_out, feats=feats)
        if len(self.keys) > self.window:
            self.keys.pop(0)
        return True
