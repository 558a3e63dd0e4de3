"""Runs the frame chain on a 26x-tiled AT128 frame (4 M points per launch): used under ncu to look at
the undistortion kernel where it is bandwidth-relevant (tools/, not part of the product)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import fflins
from fflins import api, synth
seq = synth.Sequence("at128")
batch = 26
n = seq.n * batch
ctx = fflins.Context(fflins.default_config(point_filter_num=seq.filter_num, max_points_per_frame=n, max_merge_frames=2))
pose_bl = api.Pose.make(seq.R_bl, seq.t_bl)
bufs = []
for k in range(3):
    fr = seq.frame(k)
    bufs.append((fr, torch.from_numpy(np.tile(fr["xyzi"], (batch, 1))).cuda(), torch.from_numpy(np.tile(fr["time"], batch)).cuda()))
torch.cuda.synchronize()
for rep in range(6):
    if rep == 2:
        ctx.profile_enable(True); ctx.profile_reset()
    for k, (fr, x, t) in enumerate(bufs):
        ctx.frame_process_dev(10 * rep + k, x.data_ptr(), t.data_ptr(), n, fr["ins_t"], fr["ins_p"], fr["ins_q"], pose_bl, fr["stamp"], fr["td"])
        ctx.release(10 * rep + k)
prof = ctx.profile_get()
for name, (ms, cnt) in prof.items():
    if cnt: print(f"{name:24s} {ms / cnt * 1e3:9.2f} us/launch" + (f"  {40 * n / (ms / cnt * 1e-3) / 1e9:8.1f} GB/s" if name == "undistort" else ""))
print("ok")
