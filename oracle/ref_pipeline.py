"""The per-frame / per-keyframe LiDAR call sequence executed by the REFERENCE'S OWN sources: lidar::LidarMap
(lidar_map.cc + ikd-Tree), PointCloudCommon, MISC and LidarFactor compiled from /root/reference into
oracle/_ref/libff_ref_omp.so (-O3, the reference's tbb::parallel_for sites as OpenMP tasks; Eigen / PCL / Ceres are the
stand-in headers of oracle/refshim/). CPU BASELINE ONLY: used by bench.py's cpu_baseline and --impl reference legs. Same
workload and life cycle as oracle/cpu_pipeline.py (which runs the oracle port): per frame lidarFrameProcessing; per keyframe
mergeNonKeyframes + addNewKeyFrame (kd-tree built once) + updateLidarFeaturesInMap, then the solver's LiDAR work -
addLidarFactors, 5 evaluations, lidarOutlierCullingByChi2, addLidarFactors, 10 evaluations - and the window's world clouds;
removeOldestLidarFrame when the window is full."""
import time

import numpy as np

from . import pyref as R


class RefSession:
    def __init__(self, seq, threads=None, frames_per_key=5, window=10, n_eval=(5, 10), sigma=0.1, huber=1.0):
        self.seq = seq
        self.K, self.window, self.n_eval, self.huber = frames_per_key, window, n_eval, huber
        self.map = R.RefLidarMap(seq.filter_num, point_std=sigma, window_size=window, omp=True)
        self.pending = 0
        self.frs = []
        self.ext7 = np.ascontiguousarray(seq.ext7(), dtype=np.float64)
        self.t = {}
        self.bufs = None

    def close(self):
        self.map.close()

    def _tick(self, name, t0):
        self.t[name] = self.t.get(name, 0.0) + (time.perf_counter() - t0)

    def step_frame(self, fr):
        s, m = self.seq, self.map
        t0 = time.perf_counter()
        m.frame(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], s.R_bl, s.t_bl, fr["stamp"], fr["td"])
        self._tick("frame_processing", t0)
        self.pending += 1
        if self.pending < self.K:
            return False
        self.pending = 0
        t0 = time.perf_counter()
        m.keyframe()          # merge + map voxel filter + kd-tree build + updateLidarFeaturesInMap
        self._tick("keyframe_merge_tree_associate", t0)
        m.set_motion(-1, fr["v"], fr["omega"])
        self.frs.append(fr)
        n = m.window() - 1
        if n >= 1:
            pose_refs = np.ascontiguousarray(np.stack([s.imu_pose7(f["stamp"]) for f in self.frs[:-1]]))
            pose_cur = np.ascontiguousarray(s.imu_pose7(fr["stamp"]))
            td = fr["td"]
            if self.bufs is None or len(self.bufs[0]) != n:
                self.bufs = (np.zeros((n, 19, 19)), np.zeros((n, 19)), np.zeros(n))
            t0 = time.perf_counter()
            m.add_factors()
            for _ in range(self.n_eval[0]):
                m.evaluate_factors(pose_refs, pose_cur, self.ext7, td, self.huber, out=self.bufs)
            self._tick("factor_eval", t0)
            t0 = time.perf_counter()
            m.chi2(pose_refs, pose_cur, self.ext7, td)
            self._tick("chi2", t0)
            t0 = time.perf_counter()
            m.add_factors()
            for _ in range(self.n_eval[1]):
                m.evaluate_factors(pose_refs, pose_cur, self.ext7, td, self.huber, out=self.bufs)
            self._tick("factor_eval", t0)
            t0 = time.perf_counter()
            m.reproject()
            self._tick("reproject", t0)
        if m.window() > self.window:
            t0 = time.perf_counter()
            m.remove_oldest()
            self.frs.pop(0)
            self._tick("remove_oldest", t0)
        return True
