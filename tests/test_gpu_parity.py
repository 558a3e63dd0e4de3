"""GPU-vs-oracle parity, stage by stage, through the C-ABI (SURVEY.md §7 kernel/parity map).

Bit-exact stages are fed the ORACLE's output of the previous stage, so a 1-ulp difference in an
upstream transcendental cannot leak into a bit-exact claim. Tolerances (BASELINE.json north_star):
voxel keys / output sets / kNN index sets bit-exact; fp64 residuals, Jacobians, H, b <= 1e-9
relative per block; fp32 undistortion output within 1 ulp."""
import os
import sys

import numpy as np
import pytest

from util import bits_equal, block_rel, rand_pose, random_scene_cloud, ulp_diff
from util import build_window as _build_window, window_params as _window_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def robot():
    from fflins import synth
    return synth.Sequence("robot")


# ---------------------------------------------------------------- K1 undistortion
@pytest.mark.parametrize("name,frame", [("robot", 3), ("lili", 2), ("ouster64", 1), ("at128", 4)])
def test_undistort_parity(ctx, oracle, name, frame):
    from fflins import api, synth
    seq = synth.Sequence(name)
    fr = seq.frame(frame)
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    info = c.frame_process(7, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl,
                           fr["stamp"], fr["td"])
    und, R, t, fb = oracle.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl,
                                     seq.t_bl, fr["stamp"], fr["td"], threads=8)
    Rg, tg = info.pose.numpy()
    assert np.max(np.abs(Rg - R)) < 1e-12 and np.max(np.abs(tg - t)) < 1e-9
    assert info.n_fallback == fb == 0
    full = c.download(7, api.CLOUD_MAP_FULL)
    assert full.shape == und.shape
    ulps = ulp_diff(full[:, :3], und[:, :3])
    frac = float(np.mean(ulps > 0))
    print(f"{name}: undistort ulp histogram: 0:{np.sum(ulps == 0)} 1:{np.sum(ulps == 1)} >1:{np.sum(ulps > 1)}")
    assert ulps.max() <= 1, "fp32 undistorted coordinates must agree within 1 ulp"
    assert frac < 1e-3
    assert bits_equal(full[:, 3], und[:, 3])
    # index decimation (lidar_map.cc:249-256): identical set by construction
    dec = c.download(7, api.CLOUD_LIDAR_FULL)
    assert bits_equal(dec, full[::seq.filter_num])
    assert info.n_dec == len(dec)
    # the frame chain downstream of the GPU's own undistorted cloud is bit-exact with the oracle
    down = c.download(7, api.CLOUD_LIDAR)
    assert bits_equal(down, oracle.voxel_filter(dec, 0.5))
    world = c.download(7, api.CLOUD_WORLD)
    assert bits_equal(world, oracle.project(Rg, tg, down))
    c.close()


def test_undistort_aos_matches_soa(ctx, robot):
    from fflins import api, synth
    fr = robot.frame(1)
    a = ctx.frame_process(100, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl,
                          robot.t_bl, fr["stamp"], fr["td"])
    rec = synth.pack_aos(fr["xyzi"], fr["time"])
    b = ctx.frame_process_aos(101, rec, fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl,
                              fr["stamp"], fr["td"])
    assert a.n_down == b.n_down
    for which in range(3):
        assert bits_equal(ctx.download(100, which), ctx.download(101, which))
    assert bits_equal(ctx.download(100, api.CLOUD_MAP_FULL), ctx.download(101, api.CLOUD_MAP_FULL))
    ctx.release(100)
    ctx.release(101)


def test_undistort_outside_window_falls_back(ctx, oracle, robot):
    """Points outside the INS window use window.back() (misc.cc:76-79)."""
    from fflins import api
    fr = robot.frame(2)
    time = fr["time"].copy()
    time[::7] += 100.0      # far beyond the window
    time[3::11] -= 100.0    # before the window
    info = ctx.frame_process(102, fr["xyzi"], time, fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl,
                             fr["stamp"], fr["td"])
    und, R, t, fb = oracle.undistort(fr["xyzi"], time, fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl,
                                     robot.t_bl, fr["stamp"], fr["td"])
    assert info.n_fallback == fb > 0
    full = ctx.download(102, api.CLOUD_MAP_FULL)
    assert ulp_diff(full[:, :3], und[:, :3]).max() <= 1
    ctx.release(102)


def test_identity_motion_is_identity(ctx):
    """KAT: a static INS window and identity extrinsic leave every point untouched bit for bit."""
    from fflins import api
    rng = np.random.default_rng(5)
    n, w = 5000, 50
    xyzi = rng.uniform(-50, 50, (n, 4)).astype(np.float32)
    ins_t = 1000.0 + 0.005 * np.arange(w)
    ins_p = np.zeros((w, 3))
    ins_q = np.tile(np.array([0, 0, 0, 1.0]), (w, 1))
    time = rng.uniform(ins_t[1], ins_t[-2], n)
    info = ctx.frame_process(103, xyzi, time, ins_t, ins_p, ins_q, np.eye(3), np.zeros(3), ins_t[-2], 0.0)
    assert bits_equal(ctx.download(103, api.CLOUD_MAP_FULL), xyzi)
    assert info.n_fallback == 0
    ctx.release(103)


def test_static_overload(ctx, oracle, robot):
    from fflins import api
    fr = robot.frame(0)
    p, q = fr["ins_p"][-1], fr["ins_q"][-1]
    info = ctx.frame_process_static(104, fr["xyzi"], p, q, robot.R_bl, robot.t_bl)
    R, t = oracle.state_to_pose(p, q, robot.R_bl, robot.t_bl)
    Rg, tg = info.pose.numpy()
    assert bits_equal(Rg, R) and bits_equal(tg, t)
    assert bits_equal(ctx.download(104, api.CLOUD_MAP_FULL), fr["xyzi"])
    dec = oracle.decimate(fr["xyzi"], ctx.cfg.point_filter_num)
    down = oracle.voxel_filter(dec, 0.5)
    assert bits_equal(ctx.download(104, api.CLOUD_LIDAR), down)
    assert bits_equal(ctx.download(104, api.CLOUD_WORLD), oracle.project(R, t, down))
    ctx.release(104)


def test_empty_frame(ctx, robot):
    fr = robot.frame(0)
    info = ctx.frame_process(105, np.zeros((0, 4), np.float32), np.zeros(0), fr["ins_t"], fr["ins_p"], fr["ins_q"],
                             robot.R_bl, robot.t_bl, fr["stamp"], fr["td"])
    assert (info.n_raw, info.n_dec, info.n_down) == (0, 0, 0)
    ctx.release(105)


def test_bad_window_is_rejected(ctx, robot):
    from fflins import api
    fr = robot.frame(0)
    bad = fr["ins_t"].copy()
    bad[10] = bad[9]
    with pytest.raises(api.FflinsError):
        ctx.frame_process(106, fr["xyzi"], fr["time"], bad, fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl,
                          fr["stamp"], fr["td"])


# ---------------------------------------------------------------- K2 voxel downsample
@pytest.mark.parametrize("n", [1, 2, 5, 257, 4096, 4097, 20000, 150001, 768000])
def test_voxel_bit_exact(ctx, oracle, n):
    rng = np.random.default_rng(1000 + n)
    pts = random_scene_cloud(rng, n, extent=30.0 if n > 1000 else 3.0)
    out, keys = ctx.voxel_downsample(pts, 0.5, want_keys=True)
    ref, rkeys, _ = oracle.voxel_filter(pts, 0.5, want_keys=True)
    assert bits_equal(keys, rkeys), "voxel keys must be bit-exact"
    assert len(out) == len(ref)
    assert bits_equal(out, ref), "centroids must be bit-exact (ascending key, index-ordered fp32 sums)"


def test_voxel_large_input_multi_cta_scan(ctx, oracle):
    """Above ~1 M points the radix passes scan their histogram table with the three-launch multi-CTA scan."""
    rng = np.random.default_rng(4242)
    pts = random_scene_cloud(rng, 2_500_000, extent=120.0)
    out, keys = ctx.voxel_downsample(pts, 0.5, want_keys=True)
    ref, rkeys, _ = oracle.voxel_filter(pts, 0.5, want_keys=True)
    assert bits_equal(keys, rkeys) and len(out) == len(ref) > 100000
    assert bits_equal(out, ref)


def test_voxel_cooperative_filter_ten_million_points(oracle):
    """The cooperative filter at a bandwidth-bound size: two CTAs per SM, four 8-bit passes over keys wider than 27 bits,
    several sub-tiles per chunk. Overflow of the cooperative path (n > 4,096): the input comes back unchanged."""
    from conftest import new_ctx
    rng = np.random.default_rng(991)
    n = 10_200_000
    pts = np.empty((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-700, 700, n)
    pts[:, 1] = rng.uniform(-700, 700, n)
    pts[:, 2] = rng.uniform(-20, 20, n) * (rng.uniform(0, 1, n) < 0.5)   # half of the points on a ground sheet: dense voxels
    pts[:, 3] = rng.uniform(0, 255, n)
    c = new_ctx()
    try:
        out, keys = c.voxel_downsample(pts, 0.5, want_keys=True)
        ref, rkeys, _ = oracle.voxel_filter(pts, 0.5, want_keys=True)
        assert int(rkeys.max()).bit_length() > 27
        assert bits_equal(keys, rkeys) and len(out) == len(ref) > 1_000_000
        assert bits_equal(out, ref)
        big = (pts[:50_000] * np.float32(100.0)).astype(np.float32)
        assert bits_equal(c.voxel_downsample(big, 0.01), big)
    finally:
        c.close()


def test_voxel_cooperative_filter_wide_count_table():
    """Above 65,535 points per chunk the digit counts of the cooperative filter no longer fit 16 bits: the uint32
    instantiation (otherwise only reached by the 20 M-point batched roofline launch), forced here at every size."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "vox_bench.py"), "robot", "1"], env=dict(os.environ, FFLINS_VOXEL_WIDE="1"),
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "bit_exact_vs_oracle=True" in out.stdout and "voxel.coop" in out.stdout, out.stdout + out.stderr


def test_voxel_other_leaf_and_dense_voxels(ctx, oracle):
    rng = np.random.default_rng(77)
    pts = random_scene_cloud(rng, 60000, extent=2.0)   # thousands of points per voxel
    for leaf in (0.3, 0.5, 1.0):
        assert bits_equal(ctx.voxel_downsample(pts, leaf), oracle.voxel_filter(pts, leaf))


def test_voxel_cell_centres_identity_set(ctx):
    """KAT: points at distinct cell centres come back unchanged, sorted by key (z, y, x major)."""
    g = np.arange(-3, 4, dtype=np.float32) * 0.5 + 0.25
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.arange(xx.size, dtype=np.float32)], 1)
    rng = np.random.default_rng(3)
    out = ctx.voxel_downsample(pts[rng.permutation(len(pts))], 0.5)
    order = np.lexsort((pts[:, 0], pts[:, 1], pts[:, 2]))
    assert bits_equal(out, pts[order])


def test_voxel_overflow_returns_input(ctx, oracle):
    """PCL: when the index would overflow int32 the input cloud is returned unchanged."""
    rng = np.random.default_rng(4)
    pts = rng.uniform(-1e4, 1e4, (1000, 4)).astype(np.float32)
    with pytest.raises(OverflowError):
        oracle.voxel_filter(pts, 0.01)
    out = ctx.voxel_downsample(pts, 0.01)
    assert bits_equal(out, pts)


# ---------------------------------------------------------------- K3 projection
def test_projection_bit_exact(ctx, oracle):
    rng = np.random.default_rng(11)
    pts = rng.uniform(-100, 100, (100003, 4)).astype(np.float32)
    R, t = rand_pose(rng)
    assert bits_equal(ctx.cloud_projection(R, t, pts), oracle.project(R, t, pts))


# ---------------------------------------------------------------- A2 plane fit
def test_plane_estimation_parity(ctx, oracle):
    rng = np.random.default_rng(12)
    n = 4000
    pts = np.zeros((n, 5, 4), np.float32)
    for i in range(n):
        mode = i % 4
        nrm = rng.normal(size=3)
        a, b = rng.uniform(-5, 5, 5), rng.uniform(-5, 5, 5)
        if mode == 2:
            p = rng.uniform(-5, 5, (5, 3))
        elif mode == 3:
            p = np.stack([a, 2 * a + 1e-4 * rng.normal(size=5), 1 - a + 1e-4 * rng.normal(size=5)], 1)
        else:
            z = (-4 - nrm[0] * a - nrm[1] * b) / (abs(nrm[2]) + 0.3) + (0.03 * rng.normal(size=5) if mode == 1 else 0)
            p = np.stack([a, b, z], 1)
        pts[i, :, :3] = p
    ok, unit, ninv, dist = ctx.plane_estimation(pts, 0.1)
    n_ok = 0
    for i in range(n):
        o, u, d, ds = oracle.plane_estimation(pts[i], 0.1)
        assert o == ok[i]
        if i % 4 != 3:   # well conditioned: 1e-9 relative (observed: bit-identical)
            assert block_rel(unit[i], u) <= 1e-9 and abs(ninv[i] - d) <= 1e-9 * abs(d)
            assert block_rel(dist[i], ds) <= 1e-9 or np.max(np.abs(dist[i] - ds)) < 1e-12
        n_ok += o
    assert 0 < n_ok < n


def test_plane_kat_z_equals_one(ctx):
    """KAT: five points on z = 1 give unit normal (0,0,-1), d = 1, all distances 0."""
    pts = np.zeros((1, 5, 4), np.float32)
    pts[0, :, :3] = [[0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1], [0.3, 0.6, 1]]
    ok, unit, ninv, dist = ctx.plane_estimation(pts, 0.1)
    assert ok[0]
    assert np.allclose(unit[0], [0, 0, -1], atol=1e-15) and abs(ninv[0] - 1.0) < 1e-15
    assert np.max(np.abs(dist[0])) < 1e-15


# ---------------------------------------------------------------- A1 kNN
@pytest.mark.parametrize("m,q,extent", [(0, 10, 5.0), (3, 50, 2.0), (5, 50, 1.0), (2000, 500, 8.0), (60000, 3000, 30.0)])
def test_knn5_exact(ctx, oracle, m, q, extent):
    rng = np.random.default_rng(2000 + m)
    mp = random_scene_cloud(rng, m, extent) if m else np.zeros((0, 4), np.float32)
    qs = random_scene_cloud(rng, q, extent * 1.2)
    cnt, idx, d2 = ctx.knn5(mp, qs)
    for k in range(q):
        n, ri, rd = oracle.knn5(mp, qs[k, :3])
        assert cnt[k] == n
        assert np.array_equal(idx[k, :n], ri[:n]), f"query {k}: kNN index set must be exact"
        assert bits_equal(d2[k, :n], rd[:n])
        assert np.all(idx[k, n:] == -1)


def test_knn5_ties_broken_by_index(ctx, oracle):
    """Lattice map + queries at lattice centres: many exactly equidistant neighbours."""
    g = np.arange(-4, 5, dtype=np.float32)
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    mp = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.zeros(xx.size, np.float32)], 1)
    rng = np.random.default_rng(9)
    mp = mp[rng.permutation(len(mp))]
    qs = np.zeros((200, 4), np.float32)
    qs[:, :3] = rng.integers(-3, 3, (200, 3)) + 0.5
    cnt, idx, d2 = ctx.knn5(mp, qs)
    for k in range(len(qs)):
        n, ri, rd = oracle.knn5(mp, qs[k, :3])
        assert cnt[k] == n == 5
        assert np.array_equal(idx[k], ri) and bits_equal(d2[k], rd)
        assert len(set(d2[k].tolist())) == 1   # all five are equidistant: pure index tie-break


# ---------------------------------------------------------------- M1 merge + A3/A4 association
@pytest.fixture(scope="module", params=["robot", "ouster64", "lili"])
def window(request):
    """A small sliding window per sensor config: robot (identity extrinsic), ouster64 (65 k points, 400 Hz
    IMU, td = -0.12 s, lever arm), lili (Horizon pattern, lever arm)."""
    from fflins import api, synth
    from oracle import pyoracle
    seq = synth.Sequence(request.param)
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    keys = _build_window(c, pyoracle, seq, n_keys=4 if request.param == "robot" else 3, frames_per_key=3)
    yield c, seq, keys
    c.close()


def test_merge_bit_exact(window):
    from fflins import api
    c, seq, keys = window
    for key in keys:
        assert bits_equal(c.download(key["id"], api.CLOUD_MAP_FULL), key["map_full"])
        assert bits_equal(c.download(key["id"], api.CLOUD_MAP), key["map"])


def test_association_parity(window, oracle):
    c, seq, keys = window
    st = c.associate()
    cur = keys[-1]
    assert st.n_refs == len(keys) - 1 and st.n_queries == len(cur["down"])
    total_plane = 0
    for i, ref in enumerate(keys[:-1]):
        o = oracle.associate(ref["R"], ref["t"], cur["world"], cur["down"], ref["map"], threads=8)
        g = c.features(ref["id"], nn_points=True)
        assert np.array_equal(g["nn_idx"], o["nn_idx"]), "kNN index 5-tuples must be bit-exact"
        assert bits_equal(g["nn_d2"], o["nn_d2"])
        assert np.array_equal(g["covered"], o["covered"])
        assert np.array_equal(g["type"], o["type"]), "feature type flags must be identical"
        assert block_rel(g["normal"], o["normal"]) <= 1e-9 and block_rel(g["d"], o["d"]) <= 1e-9
        assert bits_equal(g["normal"], o["normal"]) and bits_equal(g["d"], o["d"])   # observed: identical
        assert not g["outlier"].any()
        assert st.num_features[i] == o["num_features"] and st.num_covered[i] == o["num_covered"]
        found = g["nn_idx"] >= 0
        assert bits_equal(g["nn_points"][found], ref["map"][g["nn_idx"][found]])
        total_plane += o["num_features"]
    assert total_plane > 100, "the synthetic scene must produce plane features"


def test_association_empty_map(ctx, oracle):
    """a keyframe whose map is empty yields FEATURE_NONE everywhere."""
    from fflins import api
    rng = np.random.default_rng(8)
    down = random_scene_cloud(rng, 300, 5.0)
    ctx.upload(2000, api.CLOUD_MAP, np.zeros((0, 4), np.float32))
    ctx.set_pose(2000, np.eye(3), np.zeros(3))
    ctx.upload(2001, api.CLOUD_LIDAR, down)
    ctx.upload(2001, api.CLOUD_WORLD, down)
    nf, nc = ctx.associate_pair(2000, 2001)
    assert (nf, nc) == (0, 0)
    g = ctx.features(2000)
    assert np.all(g["type"] == -1) and np.all(g["nn_idx"] == -1)
    ctx.release(2000)
    ctx.release(2001)


# ---------------------------------------------------------------- F1/F4 factors
def _factor_inputs(rng, n):
    p = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rng.uniform(-5, 5, n)
    std = np.full(n, 0.1)

    def pose7(scale_t):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q) * (1 + 1e-3 * rng.uniform(-1, 1))
        return np.concatenate([rng.uniform(-10, 10, 3) * scale_t, q])
    motion = np.concatenate([rng.normal(0, 3, 3), rng.normal(0, 0.3, 3), [rng.uniform(-0.02, 0.02)],
                             rng.normal(0, 3, 3), rng.normal(0, 0.3, 3), [rng.uniform(-0.02, 0.02)]])
    return p, nrm, d, std, motion, pose7(1), pose7(1), pose7(0.02), rng.uniform(-0.02, 0.02)


@pytest.mark.parametrize("flags", [0, 1, 1 | 8 | 16, 2, 4])
def test_factor_batch_parity(ctx, oracle, flags):
    rng = np.random.default_rng(31 + flags)
    n = 3001
    p, nrm, d, std, motion, pr, pc, ext, td = _factor_inputs(rng, n)
    # make residuals O(1) so that Huber sees both branches
    r0, _, _, _, _ = oracle.factor_batch(p, nrm, d, std, None, motion, pr, pc, ext, td, reduce=False)
    d = d - 0.1 * r0 + 0.2 * rng.normal(size=n) * (rng.random(n) < 0.5)
    r, J = ctx.factor_eval_batch(p, nrm, d, std, motion, pr, pc, ext, td, flags)
    ro, Jo, _, _, _ = oracle.factor_batch(p, nrm, d, std, None, motion, pr, pc, ext, td, flags=flags, reduce=False)
    assert (np.abs(ro) > 1).any() and (np.abs(ro) < 1).any()
    assert block_rel(r, ro) <= 1e-9
    for lo, hi in ((0, 7), (7, 14), (14, 21), (21, 22)):
        assert block_rel(J[:, lo:hi], Jo[:, lo:hi]) <= 1e-9
    assert np.all(J[:, [6, 13, 20]] == 0)


def test_factor_kat_coincident_poses(ctx):
    """KAT: coincident ref/cur poses and a point on the plane give r = 0 and equal-and-opposite
    position Jacobians (lidar_factor.cc:84,93)."""
    p = np.array([[1.0, 2.0, 3.0]], np.float32)
    nrm = np.array([[0.0, 0.0, 1.0]])
    d = np.array([-3.0])
    pose = np.array([1.0, -2.0, 0.5, 0, 0, 0, 1.0])
    ext = np.array([0, 0, 0, 0, 0, 0, 1.0])
    r, J = ctx.factor_eval_batch(p, nrm, d, np.array([0.1]), np.zeros(14), pose, pose, ext, 0.0)
    assert abs(r[0]) < 1e-12
    assert np.allclose(J[0, 0:3], -J[0, 7:10], atol=0, rtol=0)
    assert np.allclose(J[0, 7:10], [0, 0, 10.0], atol=1e-12)


def test_marginalization_prior_from_device_blocks(window, oracle):
    """Next row 2: the LiDAR share of the marginalization prior. The (oldest, current) 19 x 19
    block of fflins_window_eval goes through the host Schur complement / linearisation
    (marginalization_info.cc:93-131); the same with the oracle's block must give the same prior."""
    from fflins import marginalization as M
    c, seq, keys = window
    c.associate()
    rng = np.random.default_rng(77)
    pose_refs, pose_cur, ext, td = _window_params(seq, keys, rng)
    cur = keys[-1]
    refs = [k["id"] for k in keys[:-1]]
    out = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    f = c.features(refs[0])
    valid = ((f["type"] == 0) & (f["outlier"] == 0)).astype(np.uint8)
    motion_cur = np.concatenate([cur["fr"]["v"], cur["fr"]["omega"], [cur["fr"]["td"]]])
    motion = np.concatenate([keys[0]["fr"]["v"], keys[0]["fr"]["omega"], [keys[0]["fr"]["td"]], motion_cur])
    q = len(valid)
    _, _, Ho, bo, _ = oracle.factor_batch(cur["down"][:, :3], f["normal"], f["d"], np.full(q, 0.1), valid, motion, pose_refs[0],
                                          pose_cur, ext, td, flags=1)
    Hg, bg = out["H"][0], out["b"][0]
    # A lone relative factor leaves NO information on the remaining states once its reference pose is eliminated
    # (J_ref_t = -J_cur_t): in the estimator the oldest keyframe also carries the previous prior / its IMU factor.
    # Stand in for that share with a Gaussian prior on the reference pose (sigma 1 cm / 0.2 deg).
    W = np.zeros((19, 19))
    W[:6, :6] = np.diag([1e4] * 3 + [8.2e4] * 3)

    def marginal(H, b):
        H0, b0 = M.assemble([H + W], [b], [0], 1)
        return M.schur(H0, b0, 6)
    Hp_g, bp_g = marginal(Hg, bg)
    Hp_o, bp_o = marginal(Ho, bo)
    scale = np.abs(Hg).max()
    assert np.abs(Hp_g).max() > 1e-3 * scale, "the marginal must carry information"
    assert np.abs(Hp_g - Hp_o).max() <= 1e-9 * scale and np.abs(bp_g - bp_o).max() <= 1e-9 * max(np.abs(bg).max(), 1.0)
    J0, e0 = M.linearize(Hp_g, bp_g)
    # directions the eigenvalue cut (EPS = 1e-8) keeps reproduce the marginal normal equations
    assert np.abs(J0.T @ J0 - Hp_g).max() <= 1e-9 * scale + 1e-7
    dx = rng.normal(size=13) * 1e-2
    e = M.prior_residual(J0, e0, dx)
    quad = dx @ Hp_g @ dx - 2 * bp_g @ dx
    assert abs((e @ e - e0 @ e0) - quad) <= 1e-6 * max(1.0, abs(quad))


@pytest.mark.parametrize("flags", [0, 1, 1 | 8 | 16])
def test_factor_reduce_parity(window, oracle, flags):
    """K6: per-pair H = sum J^T J, b = -sum J^T r, cost vs. the oracle's sequential sums."""
    c, seq, keys = window
    c.associate()
    rng = np.random.default_rng(50)
    pose_refs, pose_cur, ext, td = _window_params(seq, keys, rng)
    cur = keys[-1]
    refs = [k["id"] for k in keys[:-1]]
    out = c.window_eval(refs, pose_refs, pose_cur, ext, td, flags)
    motion_cur = np.concatenate([cur["fr"]["v"], cur["fr"]["omega"], [cur["fr"]["td"]]])
    for i, ref in enumerate(keys[:-1]):
        f = c.features(ref["id"])
        valid = ((f["type"] == 0) & (f["outlier"] == 0)).astype(np.uint8)
        motion = np.concatenate([ref["fr"]["v"], ref["fr"]["omega"], [ref["fr"]["td"]], motion_cur])
        q = len(valid)
        ro, Jo, Ho, bo, co = oracle.factor_batch(cur["down"][:, :3], f["normal"], f["d"], np.full(q, 0.1), valid, motion,
                                                 pose_refs[i], pose_cur, ext, td, flags=flags)
        assert out["n_res"][i] == valid.sum() > 50
        assert block_rel(out["H"][i], Ho) <= 1e-9
        assert block_rel(out["b"][i], bo) <= 1e-9
        assert block_rel(out["cost"][i], co) <= 1e-9
        assert np.array_equal(out["H"][i], out["H"][i].T)
        # the single-pair entry point returns the same block plus per-residual r, J
        one = c.factor_eval(ref["id"], cur["id"], pose_refs[i], pose_cur, ext, td, flags)
        assert bits_equal(one["H"], out["H"][i]) and bits_equal(one["b"], out["b"][i])
        assert np.array_equal(one["valid"], valid)
        assert block_rel(one["r"], ro) <= 1e-9 and block_rel(one["J"], Jo) <= 1e-9
        if flags & 8:
            assert np.all(out["H"][i][12:19, :] == 0) and np.all(out["b"][i][12:19] == 0)


def test_factor_reduce_is_deterministic(window):
    c, seq, keys = window
    c.associate()
    rng = np.random.default_rng(51)
    pose_refs, pose_cur, ext, td = _window_params(seq, keys, rng)
    refs = [k["id"] for k in keys[:-1]]
    a = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    b = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    assert bits_equal(a["H"], b["H"]) and bits_equal(a["b"], b["b"])


def test_chi2_cull_parity(window, oracle):
    c, seq, keys = window
    c.associate()
    rng = np.random.default_rng(52)
    pose_refs, pose_cur, ext, td = _window_params(seq, keys, rng)
    cur = keys[-1]
    refs = [k["id"] for k in keys[:-1]]
    before = [c.features(r) for r in refs]
    cnt = c.window_chi2(refs, pose_refs, pose_cur, ext, td, 3.841)
    motion_cur = np.concatenate([cur["fr"]["v"], cur["fr"]["omega"], [cur["fr"]["td"]]])
    for i, ref in enumerate(keys[:-1]):
        f = before[i]
        valid = (f["type"] == 0).astype(np.uint8)
        motion = np.concatenate([ref["fr"]["v"], ref["fr"]["omega"], [ref["fr"]["td"]], motion_cur])
        mask, n = oracle.chi2(cur["down"][:, :3], f["normal"], f["d"], np.full(len(valid), 0.1), valid, motion,
                              pose_refs[i], pose_cur, ext, td, 3.841)
        after = c.features(ref["id"])
        assert np.array_equal(after["outlier"], mask), "outlier mask must be identical"
        assert cnt[i] == n
    # culled residuals drop out of the next evaluation (optimizer.cc:149-153,468)
    out = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    for i, r in enumerate(refs):
        f = c.features(r)
        assert out["n_res"][i] == ((f["type"] == 0) & (f["outlier"] == 0)).sum()


def test_reproject_and_keyframe_window(window, oracle):
    from fflins import api
    c, seq, keys = window
    rng = np.random.default_rng(53)
    R, t = rand_pose(rng)
    key = keys[1]
    c.reproject(key["id"], R, t)
    assert bits_equal(c.download(key["id"], api.CLOUD_WORLD), oracle.project(R, t, key["down"]))
    Rg, tg = c.get_pose(key["id"])
    assert bits_equal(Rg, R) and bits_equal(tg, t)
    c.reproject(key["id"], key["R"], key["t"])
    ids = c.keyframe_ids()
    assert ids == [k["id"] for k in keys]
    is_key, st = c.keyframe_selection(keys[-1]["id"], 10.0, 9.9)
    assert not is_key and st[1] == 0 and st[2] < 1e-7
    is_key, st = c.keyframe_selection(keys[0]["id"], 10.0, 9.9)
    assert is_key and (st[1] > 0.4 or st[2] > np.radians(10))


# ---------------------------------------------------------------- committed golden fixtures
import os  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_knn_reference_ikdtree(ctx):
    """GPU kNN vs. outputs of the REFERENCE'S OWN ikd-Tree (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "knn_ikdtree_ref.npz"))
    q = np.zeros((len(g["query"]), 4), np.float32)
    q[:, :3] = g["query"]
    cnt, idx, d2 = ctx.knn5(g["map"], q, 5.0)
    assert np.array_equal(cnt, g["count"])
    for i in range(len(q)):
        n = cnt[i]
        assert bits_equal(d2[i, :n], g["d2"][i, :n]) and bits_equal(g["map"][idx[i, :n]], g["points"][i, :n])


def test_golden_frame(ctx):
    from fflins import api
    g = np.load(os.path.join(GOLD, "frame_small.npz"))
    c = api.Context(api.default_config(point_filter_num=int(g["filter_num"])))
    info = c.frame_process(1, g["xyzi"], g["time"], g["ins_t"], g["ins_p"], g["ins_q"], g["R_bl"], g["t_bl"],
                           float(g["stamp"]), float(g["td"]))
    R, t = info.pose.numpy()
    assert bits_equal(R, g["pose_R"]) and bits_equal(t, g["pose_t"]), "frame pose is evaluated with the reference's fp64 chain"
    full = c.download(1, api.CLOUD_MAP_FULL)
    assert ulp_diff(full[:, :3], g["undistorted"][:, :3]).max() <= 1
    if bits_equal(full, g["undistorted"]):       # no 1-ulp flip in this fixture: the whole chain is bit-exact
        assert bits_equal(c.download(1, api.CLOUD_LIDAR), g["down"]) and bits_equal(c.download(1, api.CLOUD_WORLD), g["world"])
    out, keys = c.voxel_downsample(g["decimated"], 0.5, want_keys=True)
    assert bits_equal(out, g["down"]) and bits_equal(keys, g["keys"])
    c.close()


def test_golden_factors(ctx):
    g = np.load(os.path.join(GOLD, "factors_small.npz"))
    for flags, suf in ((0, ""), (1, "_huber")):
        r, J = ctx.factor_eval_batch(g["point"], g["normal"], g["d"], g["std"], g["motion"], g["pose_ref"], g["pose_cur"],
                                     g["ext"], float(g["td"]), flags)
        assert block_rel(r, g["r" + suf]) <= 1e-9
        for lo, hi in ((0, 7), (7, 14), (14, 21), (21, 22)):
            assert block_rel(J[:, lo:hi], g["J" + suf][:, lo:hi]) <= 1e-9


# ---------------------------------------------------------------- full-size runs: size-independent properties
def test_full_size_at128_properties(oracle):
    """BASELINE.json's AT128 configuration (153,600 points/frame, 5 frames per keyframe) through the
    whole call sequence; checked through properties that do not need an O(Q*M) oracle pass."""
    from fflins import api, synth
    from fflins.pipeline import Session
    seq = synth.Sequence("at128")
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    sess = Session(c, seq, frames_per_key=5, n_eval=(1, 1))
    frames = {}
    for i in range(15):
        fr = seq.frame(i)
        fid, info = sess.process_frame(fr)
        assert info.n_raw == 153600 and info.n_dec == 1920 and 0 < info.n_down <= 1920 and info.n_fallback == 0
        frames[fid] = fr
        sess.after_frame(fid, fr)
    ids = c.keyframe_ids()
    assert len(ids) == 3
    for k in ids:
        m = c.download(k, api.CLOUD_MAP)
        full = c.download(k, api.CLOUD_MAP_FULL)
        assert len(full) == 5 * 153600
        # (1) voxel keys of the centroids are strictly ascending (sortedness) and one centroid per occupied voxel
        _, keys, _ = oracle.voxel_filter(full, 0.5, want_keys=True)
        assert len(m) == len(np.unique(keys))
        # (2) idempotence of the set: filtering the full cloud again gives the same cloud bit for bit
        assert bits_equal(c.voxel_downsample(full, 0.5), m)
        # (3) each centroid lies inside the bounding box of the input
        assert np.all(m[:, :3] >= full[:, :3].min(0) - 1e-4) and np.all(m[:, :3] <= full[:, :3].max(0) + 1e-4)
    # (4) association invariants
    cur = ids[-1]
    for r in ids[:-1]:
        f = c.features(r)
        found = f["nn_idx"] >= 0
        d2 = np.where(found, f["nn_d2"], np.float32(3e38))
        assert np.all(np.diff(d2, axis=1)[found[:, 1:]] >= 0), "neighbours ascending"
        assert np.all(f["nn_d2"][found] <= 25.0)
        assert np.all(f["covered"] == found.all(1)) and np.all(f["covered"][f["type"] == 0] == 1)
        assert np.allclose(np.linalg.norm(f["normal"][f["type"] == 0], axis=1), 1.0, atol=1e-12)
        # spot check 64 queries against the brute-force oracle
        Rr, tr = c.get_pose(r)
        mp = c.download(r, api.CLOUD_MAP)
        world = c.download(cur, api.CLOUD_WORLD)
        ref_pts = oracle.project(Rr.T, -(Rr.T) @ tr, world)   # approximate transform: only used to pick candidates
        o = oracle.associate(Rr, tr, world[:64], c.download(cur, api.CLOUD_LIDAR)[:64], mp)
        assert np.array_equal(o["nn_idx"], f["nn_idx"][:64]) and np.array_equal(o["type"], f["type"][:64])
        del ref_pts
    # (5) normal-equation blocks: symmetric, positive semi-definite, consistent counts; chi2 idempotent
    out = sess.last["eval"]
    for i in range(len(ids) - 1):
        H = out["H"][i]
        assert np.array_equal(H, H.T) and np.linalg.eigvalsh(H).min() >= -1e-9 * np.abs(H).max()
        assert out["cost"][i][0] >= 0 and out["n_res"][i] > 100
    pose_refs = np.stack([seq.imu_pose7(frames[r]["stamp"]) for r in ids[:-1]])
    again = c.window_chi2(ids[:-1], pose_refs, seq.imu_pose7(frames[cur]["stamp"]), seq.ext7(), frames[cur]["td"])
    assert np.all(again == 0), "a second cull with the same parameters finds nothing new"
    c.close()


def test_end_to_end_vs_cpu_pipeline(oracle):
    """Whole call sequence, GPU vs. the oracle pipeline fed the same raw frames (no staging): reports the
    discrete flips a 1-ulp undistortion difference can cause and bounds the drift of H."""
    from fflins import api, synth
    from fflins.pipeline import Session
    from oracle.cpu_pipeline import CpuSession
    seq = synth.Sequence("robot")
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    gs = Session(c, seq, frames_per_key=3, n_eval=(1, 1))
    cs = CpuSession(seq, threads=8, frames_per_key=3, n_eval=(1, 1), knn_mode=1)
    flips_q = 0
    for i in range(12):
        fr = seq.frame(i)
        fid, info = gs.process_frame(fr)
        f = cs.process_frame(fr)
        flips_q += abs(info.n_down - len(f["down"]))
        gs.after_frame(fid, fr)
        cs.after_frame(f)
    ids = c.keyframe_ids()
    assert len(ids) == len(cs.keys) == 4
    flips_type = 0
    for i, (r, ck) in enumerate(zip(ids[:-1], cs.keys[:-1])):
        g = c.features(r)
        if len(g["type"]) == len(ck["feat"]["type"]):
            flips_type += int(np.sum(g["type"] != ck["feat"]["type"]))
        Hg, Ho = gs.last["eval"]["H"][i], cs.last["eval"][i][2]
        assert block_rel(Hg, Ho) < 1e-3, "end-to-end H drift"
    print(f"end-to-end flips: voxel-count {flips_q}, feature-type {flips_type}")
    assert flips_q <= 2 and flips_type <= 5
    c.close()


def test_async_frames_match_sync(robot):
    """out = NULL (asynchronous) frame processing: counts are read back lazily and every cloud is
    bit-identical to the synchronous call; the frame pose is available immediately."""
    from fflins import api
    c = api.Context(api.default_config(point_filter_num=robot.filter_num))
    frames = [robot.frame(i) for i in range(6)]
    sync_infos = []
    for i, fr in enumerate(frames):
        sync_infos.append(c.frame_process(100 + i, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl,
                                          robot.t_bl, fr["stamp"], fr["td"]))
    for i, fr in enumerate(frames):
        r = c.frame_process(200 + i, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl,
                            fr["stamp"], fr["td"], sync=False)
        assert r is None
        R, t = c.get_pose(200 + i)                      # no synchronisation needed for the pose
        Rs, ts = sync_infos[i].pose.numpy()
        assert bits_equal(R, Rs) and bits_equal(t, ts)
    for i in range(6):
        info = c.frame_info(200 + i)
        assert (info.n_down, info.n_dec, info.n_fallback) == (sync_infos[i].n_down, sync_infos[i].n_dec, 0)
        for which in (api.CLOUD_LIDAR_FULL, api.CLOUD_LIDAR, api.CLOUD_WORLD, api.CLOUD_MAP_FULL):
            assert bits_equal(c.download(200 + i, which), c.download(100 + i, which))
    # the same out of page-locked memory: copies deferred to the launch, batch split over two streams
    pinned = []
    for i, fr in enumerate(frames[:4]):
        x, t = np.ascontiguousarray(fr["xyzi"], np.float32).copy(), np.ascontiguousarray(fr["time"], np.float64).copy()
        api.host_register(x)
        api.host_register(t)
        pinned.append((x, t))
        c.frame_process(400 + i, x, t, fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl, fr["stamp"], fr["td"],
                        sync=False)
    for i in (3, 0, 1, 2):
        info = c.frame_info(400 + i)
        assert (info.n_down, info.n_dec, info.n_fallback) == (sync_infos[i].n_down, sync_infos[i].n_dec, 0)
        for which in (api.CLOUD_LIDAR_FULL, api.CLOUD_LIDAR, api.CLOUD_WORLD, api.CLOUD_MAP_FULL):
            assert bits_equal(c.download(400 + i, which), c.download(100 + i, which))
    c.synchronize()
    for x, t in pinned:
        api.host_unregister(x)
        api.host_unregister(t)
    # merge straight after asynchronous frames (the keyframe path of the benchmark)
    for i, fr in enumerate(frames[:3]):
        c.frame_process(300 + i, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl,
                        fr["stamp"], fr["td"], sync=False)
    a = c.keyframe_merge(302, [300, 301])
    b = c.keyframe_merge(102, [100, 101])
    assert a.n_map == b.n_map and a.n_down == sync_infos[2].n_down
    assert bits_equal(c.download(302, api.CLOUD_MAP), c.download(102, api.CLOUD_MAP))
    c.close()


_ALT_SCRIPT = r"""
import sys, os, hashlib
sys.path.insert(0, os.path.join(sys.argv[1], "ff-lins_b200"))
import numpy as np
from fflins import api, synth
from fflins.pipeline import Session
seq = synth.Sequence("robot")
c = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=3))
sess = Session(c, seq, frames_per_key=3, n_eval=(1, 1), sync_frames=False)
for i in range(12):
    fr = seq.frame(i)
    fid, _ = sess.process_frame(fr)
    sess.after_frame(fid, fr)
h = hashlib.sha256()
for k in c.keyframe_ids():
    h.update(c.download(k, api.CLOUD_MAP).tobytes())
h.update(sess.last["eval"]["H"].tobytes())
print("DIGEST", h.hexdigest())
"""


def test_alternative_host_paths_same_bits(tmp_path):
    """FFLINS_NO_POLL (stream synchronisation instead of polled completion flags) and FFLINS_NO_HELPER_THREAD (map job
    issued by the calling thread) are debugging switches: same maps, same H blocks."""
    import subprocess
    script = tmp_path / "alt.py"
    script.write_text(_ALT_SCRIPT)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for extra in ({}, {"FFLINS_NO_POLL": "1"}, {"FFLINS_NO_HELPER_THREAD": "1"}):
        env = dict(os.environ, **extra)
        out = subprocess.run([sys.executable, str(script), root], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1] == digests[2]


def test_frame_queue_edge_cases(robot):
    """The asynchronous frame queue: more frames than one batch holds, input kinds alternating inside the queue
    (SoA / 32-byte AoS records), an empty frame, and a decimated cloud too large for the one-CTA voxel filter in the
    same batch as small ones — every cloud and count bit-identical to the synchronous calls."""
    from fflins import api, synth
    frames = [robot.frame(i) for i in range(11)]
    clouds = (api.CLOUD_LIDAR_FULL, api.CLOUD_LIDAR, api.CLOUD_WORLD, api.CLOUD_MAP_FULL)

    def run(c, fid, fr, kind, sync):
        if kind == "aos":
            return c.frame_process_aos(fid, synth.pack_aos(fr["xyzi"], fr["time"]), fr["ins_t"], fr["ins_p"], fr["ins_q"],
                                       robot.R_bl, robot.t_bl, fr["stamp"], fr["td"], sync=sync)
        x, t = fr["xyzi"], fr["time"]
        if kind == "empty":
            x, t = x[:0], t[:0]
        return c.frame_process(fid, x, t, fr["ins_t"], fr["ins_p"], fr["ins_q"], robot.R_bl, robot.t_bl, fr["stamp"], fr["td"],
                               sync=sync)

    # 1. eleven frames in flight (batch = 8), kinds alternating, one empty
    kinds = ["soa", "soa", "aos", "soa", "empty", "aos", "aos", "soa", "soa", "soa", "aos"]
    c = api.Context(api.default_config(point_filter_num=robot.filter_num))
    ref = [run(c, 100 + i, fr, k, True) for i, (fr, k) in enumerate(zip(frames, kinds))]
    for i, (fr, k) in enumerate(zip(frames, kinds)):
        assert run(c, 200 + i, fr, k, False) is None
    for i in reversed(range(len(frames))):
        info = c.frame_info(200 + i)
        assert (info.n_raw, info.n_dec, info.n_down, info.n_fallback) == (ref[i].n_raw, ref[i].n_dec, ref[i].n_down, ref[i].n_fallback)
        for which in clouds:
            assert bits_equal(c.download(200 + i, which), c.download(100 + i, which))
    assert ref[4].n_raw == 0 and ref[4].n_down == 0
    c.close()
    # 2. filter_num = 1: the decimated cloud (> 4096 points) takes the multi-kernel voxel filter inside a batch
    c = api.Context(api.default_config(point_filter_num=1))
    big = [frames[0], frames[1], frames[2]]
    small = dict(frames[3])
    small["xyzi"], small["time"] = small["xyzi"][:3000], small["time"][:3000]
    seqs = [big[0], small, big[1], small, big[2]]
    assert len(big[0]["xyzi"]) > 4096
    ref = [run(c, 100 + i, fr, "soa", True) for i, fr in enumerate(seqs)]
    for i, fr in enumerate(seqs):
        run(c, 200 + i, fr, "soa", False)
    for i in range(len(seqs)):
        info = c.frame_info(200 + i)
        assert (info.n_dec, info.n_down) == (ref[i].n_dec, ref[i].n_down) and info.n_down > 0
        for which in clouds:
            assert bits_equal(c.download(200 + i, which), c.download(100 + i, which))
    c.close()


def test_async_keyframe_pipeline_matches_sync(robot):
    """The benchmark's asynchronous call sequence (frames and keyframe merge with out = NULL, map built on
    the map stream and adopted at first use) produces bit-identical maps, features and H blocks."""
    from fflins import api
    from fflins.pipeline import Session
    frames = [robot.frame(i) for i in range(15)]
    res = []
    # "pinned": asynchronous frames out of page-locked host memory — the batch is split, the newest frame goes first
    # on the compute stream and the older ones follow on the copy stream behind their own copies
    for mode in ("sync", "async", "pinned"):
        c = api.Context(api.default_config(point_filter_num=robot.filter_num, window_size=3))
        sess = Session(c, robot, frames_per_key=3, n_eval=(1, 1), sync_frames=(mode == "sync"))
        use = frames
        if mode == "pinned":
            use = [dict(fr) for fr in frames]
            for fr in use:
                fr["xyzi"] = np.ascontiguousarray(fr["xyzi"], np.float32).copy()
                fr["time"] = np.ascontiguousarray(fr["time"], np.float64).copy()
                api.host_register(fr["xyzi"])
                api.host_register(fr["time"])
        for fr in use:
            fid, _ = sess.process_frame(fr)
            sess.after_frame(fid, fr)
        ids = c.keyframe_ids()
        maps = [c.download(k, api.CLOUD_MAP) for k in ids]
        feats = [c.features(k) for k in ids[:-1]]
        downs = [c.download(k, api.CLOUD_WORLD) for k in ids]
        res.append((list(ids), maps, feats, sess.last["eval"]["H"].copy(), [c.frame_info(k).n_map for k in ids], downs))
        c.close()
        if mode == "pinned":
            for fr in use:
                api.host_unregister(fr["xyzi"])
                api.host_unregister(fr["time"])
    ids_a, maps_a, feats_a, H_a, nm_a, downs_a = res[0]
    assert all(n > 0 for n in nm_a)
    for ids_b, maps_b, feats_b, H_b, nm_b, downs_b in res[1:]:
        assert ids_a == ids_b and nm_a == nm_b
        for ma, mb in zip(maps_a, maps_b):
            assert bits_equal(ma, mb)
        for da, db in zip(downs_a, downs_b):
            assert bits_equal(da, db)
        for fa, fb in zip(feats_a, feats_b):
            for k in ("type", "nn_idx", "nn_d2", "normal", "d", "outlier"):
                assert bits_equal(fa[k], fb[k])
        assert bits_equal(H_a, H_b)
