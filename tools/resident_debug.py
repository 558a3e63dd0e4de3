"""Window fill with the resident evaluation kernel: prints how each keyframe's evaluations were served."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
import numpy as np
import fflins
from fflins import api, synth
from fflins.pipeline import Session
seq = synth.Sequence("robot")
ctx = fflins.Context(fflins.default_config(point_filter_num=seq.filter_num))
sess = Session(ctx, seq, 2, (5, 10), sync_frames=True)
for i in range(30):
    fr = seq.frame(i)
    fid, _ = sess.process_frame(fr)
    if sess.after_frame(fid, fr):
        print("keyframes", len(ctx.keyframe_ids()), ctx.eval_stats(), flush=True)
ctx.close()
