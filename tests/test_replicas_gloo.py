"""Multi-GPU = independent replicas: the only cross-rank steps are the barrier and
the max-over-ranks of the timing. Covered here with world_size-2 gloo on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["FFLINS_ROOT"])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ms = bench.max_over_ranks([10.0 + 5 * rank, 1.0 - rank], dist, device="cpu")
frames = bench.aggregate_frames(steps=7, world=world)
dist.barrier()
if rank == 0:
    print(json.dumps({"ms": ms, "frames": frames, "world": world}))
dist.destroy_process_group()
'''


def _torchrun(args, env_extra=None, timeout=600):
    env = dict(os.environ, FFLINS_ROOT=ROOT, **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731"] + args
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_max_over_ranks_and_aggregate(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(line) == 1
    out = json.loads(line[0])
    assert out["ms"] == [15.0, 1.0]                 # MAX over ranks, element-wise
    assert out["frames"] == 7 * 5 * 2 and out["world"] == 2    # whole-job aggregate: every replica's frames


def test_reference_arm_only_rank0_prints():
    """bench.py --impl reference under torchrun: rank 0 runs and prints ONE line, the others exit 0."""
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--config", "robot",
                   "--frames", "10", "--prefill", "2"], timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["value"] > 0
    assert out["cpu_baseline"]["kind"] == "port" and out["e2e"]["h2d_bytes_per_step"] == 0
