"""Micro-benchmark + bit-exactness check of the large-cloud voxel filter (fflins_voxel_downsample) on a keyframe-map
sized input (5 merged AT128 frames = 768 k points). Usage: python tools/vox_bench.py [config] [reps]
FFLINS_VOXEL_CHAIN=1 selects the multi-kernel chain, FFLINS_VOXEL_STAMPS=1 prints the cooperative kernel's phases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200")); sys.path.insert(0, ROOT)
import fflins
from fflins import synth
from oracle import pyoracle as O

cfg = sys.argv[1] if len(sys.argv) > 1 else "at128"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
seq = synth.Sequence(cfg)
pts = []
for f in range(5):
    fr = seq.frame(f)
    und, R, t, _ = O.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"], fr["td"])
    w = und.copy()
    w[:, :3] = (und[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)
    pts.append(w)
cloud = np.ascontiguousarray(np.concatenate(pts), np.float32)
ctx = fflins.Context(fflins.default_config(point_filter_num=seq.filter_num))
ref = O.voxel_filter(cloud, 0.5)
out = ctx.voxel_downsample(cloud, 0.5)
ok = out.tobytes() == ref.tobytes()
print(f"{cfg}: n={len(cloud)} voxels={len(out)} bit_exact_vs_oracle={ok}")
for n in (5000, 8191, 8193, 70000, 300001):
    sub = cloud[:n]
    a, k = ctx.voxel_downsample(sub, 0.5, want_keys=True)
    b, kb, _ = O.voxel_filter(sub, 0.5, want_keys=True)
    e = a.tobytes() == b.tobytes() and np.array_equal(k, kb)
    ok = ok and e
    print(f"  n={n}: voxels={len(a)} exact={a.tobytes() == b.tobytes()} keys={np.array_equal(k, kb)}")
big = np.ascontiguousarray(cloud[:20000] * np.float32(100.0))
tiny = ctx.voxel_downsample(big, 0.01)   # PCL overflow rule: input returned unchanged
print("  overflow rule:", len(tiny) == 20000 and tiny.tobytes() == big.tobytes())
ok = ok and len(tiny) == 20000
ctx.profile_enable(True); ctx.profile_reset()
for _ in range(reps):
    ctx.voxel_downsample(cloud, 0.5)
prof = ctx.profile_get()
ctx.profile_enable(False)
tot = 0.0
for name, (ms, cnt) in sorted(prof.items()):
    if name.startswith("voxel"):
        print(f"  {name}: {ms / reps * 1e3:.1f} us per filter ({cnt} scopes)")
        tot += ms / reps
print(f"  total {tot * 1e3:.1f} us per filter of {len(cloud)} points; algorithmic bytes {16 * len(cloud) + 16 * len(out)} -> {(16 * len(cloud) + 16 * len(out)) / tot / 1e6:.1f} GB/s")
if not ok:
    sys.exit(1)
