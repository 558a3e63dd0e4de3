"""The product's __host__ __device__ math headers (ff-lins_b200/csrc/math_*.cuh) executed on the CPU
against the oracle (tests/host_math_check.cu): catches algebra slips without GPU time."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_math_against_oracle(tmp_path, oracle):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "host_math_check"
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    subprocess.check_call([nvcc, *ccbin, "-O2", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets", "-I",
                           os.path.join(ROOT, "ff-lins_b200", "csrc"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "host_math_check.cu"), "-L", os.path.join(ROOT, "oracle"), "-loracle",
                           "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")])
    res = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    res = {k: float(v) for k, v in res.items()}
    # fp64 factors: <= 1e-9 relative (BASELINE.json); observed 1e-14
    for k in ("factor_r", "factor_J_ref", "factor_J_cur", "factor_J_ext", "factor_J_td", "huber"):
        assert res[k] <= 1e-9, (k, res[k])
    assert res["plane_rel"] <= 1e-9 and res["plane_ok_mismatch"] == 0
    assert res["plane_bit_mismatch"] == 0          # -fmad=false: identical to the oracle bit for bit
    assert res["frame_pose_bit_mismatch"] == 0     # host frame pose == oracle
    assert res["undistort_pose_rel"] <= 1e-13      # closed-form interval rows vs the per-point chain
    assert res["window_index_mismatch"] == 0
