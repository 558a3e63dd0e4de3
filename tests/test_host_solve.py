"""The numerics of the device-side window solve (ff-lins_b200/csrc/math_solve.cuh) executed on the CPU
(tests/host_solve_check.cu, one thread standing in for the solver CTA) against an independent numpy
Levenberg-Marquardt (tests/lm_ref.py) over the ORACLE's per-pair blocks: catches slips in the assembly,
the elimination and the accept / reject control before any GPU time is spent. The hand-offs between the
CTAs of the kernel are what tests/test_gpu_solve.py adds on the device."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import lm_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = tmp_path_factory.mktemp("solve") / "libhost_solve_check.so"
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    subprocess.check_call([nvcc, *ccbin, "-O2", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler",
                           "-fPIC", "-I", os.path.join(ROOT, "ff-lins_b200", "csrc"), "-o", str(so),
                           os.path.join(ROOT, "tests", "host_solve_check.cu")])
    return C.CDLL(str(so))


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def rand_unit(rng, shape):
    v = rng.normal(size=shape + (3,))
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def make_problem(oracle, rng, n, q, noise=0.02, outlier_frac=0.02):
    """n references, q points seen by the current frame; planes chosen so that every residual vanishes at the
    true state, then measurement noise and a few gross outliers."""
    def pose(scale):
        t = rng.normal(size=3) * scale
        qv = rng.normal(size=4)
        qv[3] += 4.0
        return np.concatenate([t, qv / np.linalg.norm(qv)])
    refs = np.stack([pose(2.0) for _ in range(n)])
    cur = pose(2.0)
    ext = np.concatenate([rng.normal(size=3) * 0.1, [0.01, -0.02, 0.015, 1.0]])
    ext[3:] /= np.linalg.norm(ext[3:])
    td = 0.004
    x_true = np.concatenate([refs.reshape(-1), cur, ext, [td]])
    pts = (rng.normal(size=(q, 3)) * np.array([15.0, 15.0, 3.0])).astype(np.float32)
    normal = rand_unit(rng, (n, q))
    motions = np.zeros((n, 14))
    motions[:, 0:3] = rng.normal(size=(n, 3)) * 0.5
    motions[:, 3:6] = rng.normal(size=(n, 3)) * 0.05
    motions[:, 6] = 0.001
    motions[:, 7:10] = rng.normal(size=3) * 0.5
    motions[:, 10:13] = rng.normal(size=3) * 0.05
    motions[:, 13] = 0.002
    d = np.zeros((n, q))
    valid = np.ones((n, q), np.uint8)
    std = 0.1
    for p in range(n):
        for k in range(q):
            r, _ = oracle.lidar_factor_evaluate(pts[k], normal[p, k], 0.0, std, motions[p], refs[p], cur, ext, td,
                                                want_jac=(0, 0, 0, 0))
            d[p, k] = -r * std
    d += rng.normal(size=d.shape) * noise
    gross = rng.random(size=d.shape) < outlier_frac
    d[gross] += rng.choice([-1.0, 1.0], size=gross.sum()) * 1.5
    type_ = np.zeros((n, q), np.int8)
    type_[rng.random(size=(n, q)) < 0.2] = -1   # not every query is a plane feature of every reference
    return dict(n=n, q=q, pts=pts, normal=normal, d=d, std=std, motions=motions, type=type_, x_true=x_true)


def perturb(rng, P, trans=0.05, rot=0.01):
    n = P["n"]
    x = P["x_true"].copy()
    for k in range(n + 1):
        dl = np.concatenate([rng.normal(size=3) * trans, rng.normal(size=3) * rot])
        x[7 * k:7 * k + 7] = lm_ref.pose_plus(x[7 * k:7 * k + 7], dl)
    return x


def oracle_fns(oracle, P, flags, outlier):
    n, q = P["n"], P["q"]

    def valid():
        return ((P["type"] == 0) & (outlier == 0)).astype(np.uint8)

    def ev(x):
        return oracle.window_eval(P["pts"], P["normal"], P["d"], P["std"], valid(), P["motions"],
                                  np.ascontiguousarray(x[:7 * n].reshape(n, 7)), np.ascontiguousarray(x[7 * n:7 * n + 7]),
                                  np.ascontiguousarray(x[7 * n + 7:7 * n + 14]), float(x[7 * n + 14]), flags)

    def chi2(x, thr):
        return oracle.window_chi2(P["pts"], P["normal"], P["d"], P["std"], (P["type"] == 0).astype(np.uint8), P["motions"],
                                  np.ascontiguousarray(x[:7 * n].reshape(n, 7)), np.ascontiguousarray(x[7 * n:7 * n + 7]),
                                  np.ascontiguousarray(x[7 * n + 7:7 * n + 14]), float(x[7 * n + 14]), outlier, thr)
    return ev, chi2


def run_harness(lib, P, x, flags, prior=None, max_iterations=(5, 10), chi2=3.841, tol=(1e-6, 1e-10, 1e-8), fixed_refs=0):
    n, q = P["n"], P["q"]
    pts4 = np.zeros((q, 4), np.float32)
    pts4[:, :3] = P["pts"]
    outlier = np.zeros((n, q), np.uint8)
    motions = np.zeros((n + 1, 7))
    motions[:n] = P["motions"][:, :7]
    motions[n] = P["motions"][0, 7:]
    xs = np.array(x, dtype=np.float64)
    opt = np.array([max_iterations[0], max_iterations[1], chi2, tol[0], tol[1], tol[2], 1e4, flags, fixed_refs], np.float64)
    rep = np.zeros(10 + n)
    H0 = b0 = x0 = None
    c0 = 0.0
    if prior is not None:
        H0, b0, x0, c0 = [np.ascontiguousarray(a, dtype=np.float64) if not np.isscalar(a) else a for a in prior]
    rc = lib.solve_host_run(n, q, _p(pts4), _p(P["type"]), _p(outlier), _p(np.ascontiguousarray(P["normal"])),
                            _p(np.ascontiguousarray(P["d"])), _p(motions), C.c_double(P["std"]), C.c_double(1.0), _p(xs), _p(opt),
                            _p(H0), _p(b0), _p(x0), C.c_double(c0), _p(rep))
    assert rc == 0
    return xs, rep, outlier


def make_prior(rng, P, x0, weight=25.0):
    """A dense positive definite prior around x0, as a marginalization would leave it (J0^T J0)."""
    D = 6 * P["n"] + 13
    J0 = rng.normal(size=(D + 5, D)) * np.sqrt(weight)
    e0 = rng.normal(size=D + 5) * 0.1
    return J0.T @ J0, -J0.T @ e0, x0.copy(), 0.5 * e0 @ e0


def test_spd_solve_matches_numpy(harness):
    rng = np.random.default_rng(3)
    for n in (1, 4, 10, 15):
        D = 6 * n + 13
        M = rng.normal(size=(D + 3, D))
        A = M.T @ M + np.eye(D) * 1e-3
        b = rng.normal(size=D)
        x = np.zeros(D)
        assert harness.solve_host_spd(n, _p(np.ascontiguousarray(A)), _p(b), _p(x)) == 0
        ref = np.linalg.solve(A, b)
        assert np.max(np.abs(x - ref)) <= 1e-9 * np.max(np.abs(ref))
    A = -np.eye(19)
    assert harness.solve_host_spd(1, _p(A), _p(np.ones(19)), _p(np.zeros(19))) == 1   # not positive definite


@pytest.mark.parametrize("case", ["cur_only", "all_free_prior", "gauge_fixed_ref", "ext_td_const_prior"])
def test_solve_numerics_match_numpy_lm(harness, oracle, case):
    rng = np.random.default_rng({"cur_only": 1, "all_free_prior": 2, "gauge_fixed_ref": 3, "ext_td_const_prior": 4}[case])
    n, q = (3, 260) if case != "all_free_prior" else (5, 300)
    P = make_problem(oracle, rng, n, q)
    x_init = perturb(rng, P)
    flags, prior, fixed_refs = 1, None, 0
    if case == "cur_only":
        flags |= 2 | 8 | 16            # references, extrinsic and td constant: plain scan-to-map registration
    elif case == "all_free_prior":
        prior = make_prior(rng, P, x_init)
    elif case == "gauge_fixed_ref":
        flags |= 8 | 16
        fixed_refs = 0b001
        prior = make_prior(rng, P, x_init, weight=1.0)
    else:
        flags |= 8 | 16
        prior = make_prior(rng, P, x_init)
    xs, rep, outlier_h = run_harness(harness, P, x_init, flags, prior, fixed_refs=fixed_refs)
    outlier_o = np.zeros((n, q), np.uint8)
    ev, chi2 = oracle_fns(oracle, P, flags, outlier_o)
    xr, (s0, s1), counts = lm_ref.window_solve(ev, chi2, n, x_init, flags, prior=prior, fixed_refs=fixed_refs)
    assert np.array_equal(outlier_h, outlier_o) and np.array_equal(rep[10:10 + n].astype(int), counts)
    assert counts.sum() > 0
    assert [int(v) for v in rep[0:3]] == [s0["iterations"], s0["successful"], s0["evaluations"]]
    assert [int(v) for v in rep[5:8]] == [s1["iterations"], s1["successful"], s1["evaluations"]]
    assert s0["successful"] >= 1
    assert abs(rep[3] - s0["cost_initial"]) <= 1e-9 * s0["cost_initial"]
    assert abs(rep[9] - s1["cost_final"]) <= 1e-9 * s1["cost_final"]
    assert np.max(np.abs(xs - xr)) <= 1e-9, np.max(np.abs(xs - xr))
    # and the solve does what a solve should: the cost went down and the current pose moved towards the truth
    assert rep[9] < rep[3]
    if case == "cur_only":
        e0 = np.linalg.norm(x_init[7 * n:7 * n + 3] - P["x_true"][7 * n:7 * n + 3])
        e1 = np.linalg.norm(xs[7 * n:7 * n + 3] - P["x_true"][7 * n:7 * n + 3])
        assert e1 < 0.5 * e0   # (the references keep their own perturbation: the optimum is not the truth)
        assert np.array_equal(xs[:7 * n], x_init[:7 * n]) and np.array_equal(xs[7 * n + 7:], x_init[7 * n + 7:])


def test_solve_without_iterations_only_culls(harness, oracle):
    rng = np.random.default_rng(9)
    P = make_problem(oracle, rng, 2, 200)
    x_init = perturb(rng, P, 0.01, 0.002)
    xs, rep, outlier = run_harness(harness, P, x_init, 1, max_iterations=(0, 0))
    assert np.array_equal(xs, x_init) and rep[0] == 0 and rep[2] == 1 and rep[7] == 1 and outlier.sum() > 0
