"""CPU-oracle LiDAR backend for fflins.fusion.FusionLoop (TEST INFRASTRUCTURE ONLY): the same interface as
fflins.fusion.GpuLidar, every LiDAR operation by the oracle. Used by tests/test_gpu_closed_loop_imu.py to run the
identical host loop once on the GPU library and once on the oracle."""
import numpy as np

from . import pyoracle as O


class CpuLidar:
    def __init__(self, seq, window=10, threads=8):
        self.seq, self.threads = seq, threads
        self.frames, self.keys = {}, []
        self.feats, self.outl = {}, {}

    def process_frame(self, fid, fr, ins_t, ins_p, ins_q):
        s = self.seq
        und, R, t, _ = O.undistort(fr["xyzi"], fr["time"], ins_t, ins_p, ins_q, s.R_bl, s.t_bl, fr["stamp"], fr["td"],
                                   threads=self.threads)
        down = O.voxel_filter(O.decimate(und, s.filter_num), 0.5)
        self.frames[fid] = dict(und=und, down=down, R=R, t=t, fr=fr)

    def make_keyframe(self, fid, nonkeys, v, omega, td):
        f = self.frames[fid]
        clouds = [f["und"]]
        for nk in nonkeys:
            g = self.frames.pop(nk)
            Rr, tr = O.relative_pose(f["R"], f["t"], g["R"], g["t"])
            clouds.append(O.project(Rr, tr, g["und"]))
        f["map"] = O.voxel_filter(np.concatenate(clouds), 0.5)
        f["motion"] = np.concatenate([v, omega, [td]])
        f["id"] = fid
        self.keys.append(f)

    def associate(self):
        cur, refs = self.keys[-1], self.keys[:-1]
        world = O.project(cur["R"], cur["t"], cur["down"])
        self.feats = {r["id"]: O.associate(r["R"], r["t"], world, cur["down"], r["map"], threads=self.threads) for r in refs}
        self.outl = {r["id"]: np.zeros(len(cur["down"]), np.uint8) for r in refs}
        return int(sum(ft["num_features"] for ft in self.feats.values()))

    def callbacks(self):
        cur = self.keys[-1]
        pts = np.ascontiguousarray(cur["down"][:, :3])
        std = np.full(len(pts), 0.1)
        motions = {r["id"]: np.concatenate([r["motion"], cur["motion"]]) for r in self.keys[:-1]}

        def evaluate(ref_ids, pose_refs, pose_cur, ext, td, flags):
            n = len(ref_ids)
            H, b, c = np.zeros((n, 19, 19)), np.zeros((n, 19)), np.zeros((n, 2))
            for i, rid in enumerate(ref_ids):
                ft = self.feats[rid]
                valid = ((ft["type"] == 0) & (self.outl[rid] == 0)).astype(np.uint8)
                _, _, H[i], b[i], c[i] = O.factor_batch(pts, ft["normal"], ft["d"], std, valid, motions[rid], pose_refs[i], pose_cur,
                                                        ext, td, flags=flags, threads=self.threads)
            return H, b, c

        def cull(ref_ids, pose_refs, pose_cur, ext, td, thr):
            cnt = []
            for i, rid in enumerate(ref_ids):
                ft = self.feats[rid]
                valid = ((ft["type"] == 0) & (self.outl[rid] == 0)).astype(np.uint8)
                mask, c = O.chi2(pts, ft["normal"], ft["d"], std, valid, motions[rid], pose_refs[i], pose_cur, ext, td, thr,
                                 threads=self.threads)
                self.outl[rid] |= mask
                cnt.append(c)
            return np.array(cnt)
        return dict(evaluate=evaluate, cull=cull)

    def set_poses(self, ids, poses):
        by_id = {k["id"]: k for k in self.keys}
        for fid, (R, t) in zip(ids, poses):
            if fid in by_id:
                by_id[fid]["R"], by_id[fid]["t"] = np.array(R), np.array(t)

    def remove_oldest(self):
        k = self.keys.pop(0)
        self.frames.pop(k["id"], None)
        return k["id"]

    def close(self):
        pass
