"""Runs only bench.batched_roofline (26 AT128 frames per launch). FFLINS_VOXEL_STAMPS=1 prints the cooperative filter's phases."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
import bench
from fflins import synth
seq = synth.Sequence("at128")
frames = [seq.frame(k) for k in range(bench.FRAMES_PER_KEY)]
print(json.dumps(bench.batched_roofline(seq, frames, 0)))
