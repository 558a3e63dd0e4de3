"""Generates the committed golden fixtures under tests/golden/ (run here, in the build container):

  knn_ikdtree_ref.npz   outputs of the REFERENCE'S OWN ikd-Tree (compiled from /root/reference into
                        oracle/_ref/libikd_ref.so): build(map) + nearestSearch(q, 5, ., 5.0) exactly
                        as lidar_map.cc:113-114,368 call them. Pins kNN semantics against the reference.
  frame_small.npz       a 2,000-point Livox-shaped frame, 20-entry-per-100ms INS window, and the oracle's
                        undistorted / decimated / voxel-filtered / world clouds.
  factors_small.npz     64 residuals: inputs, oracle r / J (raw and Huber-corrected), H, b, cost.

The reference itself cannot be run here (no Eigen/PCL/Ceres), so only the kNN fixture comes from
reference code; the others freeze the oracle (regression pins) and give the GPU tests fixed vectors.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
sys.path.insert(0, ROOT)

from fflins import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def main():
    rng = np.random.default_rng(20231019)
    # ---- kNN from the reference's ikd-Tree
    assert O.ikd_available(), "build oracle/_ref first (make -C oracle)"
    m = 5000
    mp = np.zeros((m, 4), np.float32)
    mp[:2500, 0:2] = rng.uniform(-12, 12, (2500, 2))
    mp[:2500, 2] = rng.normal(0, 0.02, 2500)
    mp[2500:, :3] = rng.uniform(-12, 12, (2500, 3)) * [1, 1, 0.3]
    mp = O.voxel_filter(mp, 0.5)          # the reference's maps are voxel centroids
    qs = np.zeros((64, 3), np.float32)
    qs[:48] = rng.uniform(-13, 13, (48, 3)) * [1, 1, 0.2]
    qs[48:] = rng.uniform(-40, 40, (16, 3))   # far away: fewer than 5 neighbours within 5 m
    tree = O.IkdTreeRef(mp, 0.5)
    cnt = np.zeros(64, np.int32)
    pts = np.zeros((64, 5, 4), np.float32)
    d2 = np.zeros((64, 5), np.float32)
    for i, q in enumerate(qs):
        n, p, d = tree.search(q, 5, 5.0)
        cnt[i] = n
        pts[i, :n] = p
        d2[i, :n] = d
    tree.close()
    np.savez_compressed(os.path.join(HERE, "knn_ikdtree_ref.npz"), map=mp, query=qs, count=cnt, points=pts, d2=d2)

    # ---- small frame
    seq = synth.Sequence("robot", n_points=2000, ins_entries=200)
    fr = seq.frame(4)
    und, R, t, fb = O.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl,
                                fr["stamp"], fr["td"])
    dec = O.decimate(und, seq.filter_num)
    down, keys, order = O.voxel_filter(dec, 0.5, want_keys=True)
    world = O.project(R, t, down)
    np.savez_compressed(os.path.join(HERE, "frame_small.npz"), xyzi=fr["xyzi"], time=fr["time"], ins_t=fr["ins_t"],
                        ins_p=fr["ins_p"], ins_q=fr["ins_q"], R_bl=seq.R_bl, t_bl=seq.t_bl, stamp=fr["stamp"],
                        td=fr["td"], filter_num=seq.filter_num, undistorted=und, pose_R=R, pose_t=t, decimated=dec,
                        down=down, keys=keys, world=world)

    # ---- factors
    n = 64
    p = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rng.uniform(-5, 5, n)
    std = np.full(n, 0.1)

    def pose7(scale):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        return np.concatenate([rng.uniform(-10, 10, 3) * scale, q])
    motion = np.concatenate([rng.normal(0, 3, 3), rng.normal(0, 0.3, 3), [0.004], rng.normal(0, 3, 3),
                             rng.normal(0, 0.3, 3), [-0.007]])
    pr, pc, ext, td = pose7(1), pose7(1), pose7(0.02), 0.011
    r0, _, _, _, _ = O.factor_batch(p, nrm, d, std, None, motion, pr, pc, ext, td, reduce=False)
    d = d - 0.1 * r0 + 0.15 * rng.normal(size=n)
    r, J, H, b, cost = O.factor_batch(p, nrm, d, std, None, motion, pr, pc, ext, td, flags=0)
    rh, Jh, Hh, bh, costh = O.factor_batch(p, nrm, d, std, None, motion, pr, pc, ext, td, flags=1)
    np.savez_compressed(os.path.join(HERE, "factors_small.npz"), point=p, normal=nrm, d=d, std=std, motion=motion,
                        pose_ref=pr, pose_cur=pc, ext=ext, td=td, r=r, J=J, H=H, b=b, cost=cost, r_huber=rh,
                        J_huber=Jh, H_huber=Hh, b_huber=bh, cost_huber=costh)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
