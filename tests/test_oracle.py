"""CPU tests of the oracle: pins against the reference's own ikd-Tree, independent numerical
libraries (LAPACK via scipy), finite differences, hand-computable KATs and the committed golden
fixtures. The reference has no tests or golden vectors for this path, so these
pins are what anchors the oracle ("parity unpinned" for the Eigen / PCL / Ceres arithmetic)."""
import os

import numpy as np
import pytest

from util import bits_equal, block_rel, rand_pose, random_scene_cloud

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- kNN: pinned on the reference's ikd-Tree
def test_knn_matches_reference_ikdtree_golden(oracle):
    """Golden outputs of KD_TREE::nearestSearch (reference code, tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "knn_ikdtree_ref.npz"))
    some_short = False
    for i, q in enumerate(g["query"]):
        n, idx, d2 = oracle.knn5(g["map"], q, 5.0)
        assert n == g["count"][i]
        assert bits_equal(d2[:n], g["d2"][i, :n]), "fp32 squared distances must be bit-exact"
        # the reference returns points, not indices: compare coordinates (no exact ties in this fixture)
        assert bits_equal(g["map"][idx[:n]], g["points"][i, :n])
        some_short |= n < 5
    assert some_short, "fixture must contain queries with fewer than 5 neighbours inside 5 m"


def test_knn_matches_reference_ikdtree_live(oracle):
    """Same pin against the live library when oracle/_ref was built (it travels to the GPU box)."""
    if not oracle.ikd_available():
        pytest.skip("oracle/_ref/libikd_ref.so not built")
    rng = np.random.default_rng(3)
    mp = oracle.voxel_filter((rng.uniform(-15, 15, (20000, 4)) * [1, 1, 0.2, 1]).astype(np.float32), 0.5)
    tree = oracle.IkdTreeRef(mp, 0.5)
    for q in (rng.uniform(-16, 16, (300, 3)) * [1, 1, 0.3]).astype(np.float32):
        n, pts, d2 = tree.search(q, 5, 5.0)
        m, idx, od2 = oracle.knn5(mp, q, 5.0)
        assert n == m and bits_equal(d2, od2[:n]) and bits_equal(pts, mp[idx[:n]])
    tree.close()


def test_knn_cutoff_is_five_metres_not_five_square_metres(oracle):
    """max_dist = 5.0 is squared inside Search (ikd_Tree.cc:902): neighbours up to 5 m are returned."""
    mp = np.array([[4.9, 0, 0, 0], [0, 4.99, 0, 0], [0, 0, 5.0, 0], [5.01, 0, 0, 0], [3, 4, 0, 0], [0, 3, 4.01, 0]], np.float32)
    n, idx, d2 = oracle.knn5(mp, np.zeros(3, np.float32), 5.0)
    assert n == 4 and set(idx[:4].tolist()) == {0, 1, 2, 4}
    assert np.all(np.diff(d2[:4]) >= 0)


def test_knn_kdtree_equals_brute_force(oracle):
    rng = np.random.default_rng(5)
    mp = (rng.uniform(-10, 10, (3000, 4)) * [1, 1, 0.1, 1]).astype(np.float32)
    qs = (rng.uniform(-11, 11, (400, 4)) * [1, 1, 0.2, 1]).astype(np.float32)
    a = oracle.associate(np.eye(3), np.zeros(3), qs, qs, mp, knn_mode=0)
    b = oracle.associate(np.eye(3), np.zeros(3), qs, qs, mp, knn_mode=1)
    for k in ("nn_idx", "nn_d2", "type", "covered", "normal", "d"):
        assert bits_equal(a[k], b[k])


# ---------------------------------------------------------------- plane fit vs LAPACK
def test_plane_fit_matches_lapack(oracle):
    scipy_linalg = pytest.importorskip("scipy.linalg")
    rng = np.random.default_rng(7)
    for trial in range(300):
        nrm = rng.normal(size=3)
        a, b = rng.uniform(-5, 5, 5), rng.uniform(-5, 5, 5)
        z = (-4 - nrm[0] * a - nrm[1] * b) / (abs(nrm[2]) + 0.3) + 0.02 * rng.normal(size=5)
        pts = np.zeros((5, 4), np.float32)
        pts[:, :3] = np.stack([a, b, z], 1)
        ok, unit, ninv, dist = oracle.plane_estimation(pts, 0.1)
        A = pts[:, :3].astype(np.float64)
        n_ls, *_ = scipy_linalg.lstsq(A, -np.ones(5), lapack_driver="gelsy")   # pivoted-QR driver
        ninv_ref = 1 / np.linalg.norm(n_ls)
        assert block_rel(unit, n_ls * ninv_ref) < 1e-10 and abs(ninv - ninv_ref) < 1e-10 * ninv_ref
        d_ref = A @ (n_ls * ninv_ref) + ninv_ref
        assert ok == bool(np.all(np.abs(d_ref) <= 0.1))


def test_plane_kat(oracle):
    pts = np.zeros((5, 4), np.float32)
    pts[:, :3] = [[0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1], [0.3, 0.6, 1]]
    ok, unit, ninv, dist = oracle.plane_estimation(pts, 0.1)
    assert ok and np.allclose(unit, [0, 0, -1], atol=1e-15) and abs(ninv - 1) < 1e-15 and np.max(np.abs(dist)) < 1e-15
    pts[4, 2] = 1.5                       # 0.5 m off the plane: rejected
    assert not oracle.plane_estimation(pts, 0.1)[0]


# ---------------------------------------------------------------- pose interpolation vs scipy Slerp
def test_pose_interpolation_matches_slerp(oracle):
    st = pytest.importorskip("scipy.spatial.transform")
    rng = np.random.default_rng(9)
    w = 40
    ins_t = 100.0 + 0.005 * np.arange(w)
    rot = st.Rotation.from_rotvec(np.cumsum(rng.normal(0, 0.004, (w, 3)), 0))
    ins_q = rot.as_quat()
    ins_p = np.cumsum(rng.normal(0, 0.01, (w, 3)), 0)
    Rbl = st.Rotation.from_rotvec([0.01, -0.02, 0.03]).as_matrix()
    tbl = np.array([0.05, -0.02, 0.1])
    slerp = st.Slerp(ins_t, rot)
    for time in rng.uniform(ins_t[0], ins_t[-1] - 1e-9, 200):
        ok, R, t = oracle.pose_from_ins_window(ins_t, ins_p, ins_q, Rbl, tbl, time)
        assert ok
        Rq = slerp([time]).as_matrix()[0]
        i = np.searchsorted(ins_t, time, side="right")
        s = (time - ins_t[i - 1]) / (ins_t[i] - ins_t[i - 1])
        p = ins_p[i - 1] + s * (ins_p[i] - ins_p[i - 1])
        assert np.max(np.abs(R - Rq @ Rbl)) < 1e-12 and np.max(np.abs(t - (p + Rq @ tbl))) < 1e-12
    # outside the window: newest state, returns false (misc.cc:76-79)
    ok, R, t = oracle.pose_from_ins_window(ins_t, ins_p, ins_q, Rbl, tbl, ins_t[-1] + 1)
    assert not ok and np.max(np.abs(R - rot[-1].as_matrix() @ Rbl)) < 1e-12


def test_window_index_edges(oracle):
    t = np.array([1.0, 2.0, 3.0, 4.0])
    assert oracle.ins_window_index(t, 0.5) == 0          # before front
    assert oracle.ins_window_index(t, 4.0) == 0          # back <= time
    assert oracle.ins_window_index(t, 1.0) == 1          # first <= time < second
    assert oracle.ins_window_index(t, 2.0) == 2
    assert oracle.ins_window_index(t, 3.999) == 3
    big = np.arange(65536, dtype=np.float64)
    for x in (0.5, 12345.25, 65534.5):
        assert oracle.ins_window_index(big, x) == int(x) + 1   # converges inside the 17-iteration cap


def test_identity_motion_undistortion_is_identity(oracle):
    rng = np.random.default_rng(5)
    n, w = 2000, 50
    xyzi = rng.uniform(-50, 50, (n, 4)).astype(np.float32)
    ins_t = 1000.0 + 0.005 * np.arange(w)
    out, R, t, fb = oracle.undistort(xyzi, rng.uniform(ins_t[1], ins_t[-2], n), ins_t, np.zeros((w, 3)),
                                     np.tile([0, 0, 0, 1.0], (w, 1)), np.eye(3), np.zeros(3), ins_t[-2], 0.0)
    assert bits_equal(out, xyzi) and fb == 0


# ---------------------------------------------------------------- voxel filter
def test_voxel_kats(oracle):
    g = np.arange(-3, 4, dtype=np.float32) * 0.5 + 0.25
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.arange(xx.size, dtype=np.float32)], 1)
    rng = np.random.default_rng(3)
    out, keys, order = oracle.voxel_filter(pts[rng.permutation(len(pts))], 0.5, want_keys=True)
    assert bits_equal(out, pts[np.lexsort((pts[:, 0], pts[:, 1], pts[:, 2]))])   # key = i + j*dx + k*dx*dy
    assert sorted(keys.tolist()) == list(range(343))
    # centroid of one voxel: sequential fp32 sum in index order, divided by float(count)
    one = np.array([[0.1, 0.2, 0.3, 1], [0.4, 0.1, 0.2, 2], [0.2, 0.45, 0.1, 4]], np.float32)
    acc = np.zeros(4, np.float32)
    for p in one:
        acc = (acc + p).astype(np.float32)
    assert bits_equal(oracle.voxel_filter(one, 0.5)[0], (acc / np.float32(3)).astype(np.float32))
    assert len(oracle.voxel_filter(np.zeros((0, 4), np.float32), 0.5)) == 0
    with pytest.raises(OverflowError):
        oracle.voxel_filter(rng.uniform(-1e4, 1e4, (100, 4)).astype(np.float32), 0.01)


def test_voxel_numpy_crosscheck(oracle):
    """Keys and voxel membership against an independent numpy formulation."""
    rng = np.random.default_rng(8)
    pts = (rng.uniform(-20, 20, (30000, 4)) * [1, 1, 0.2, 1]).astype(np.float32)
    out, keys, order = oracle.voxel_filter(pts, 0.5, want_keys=True)
    inv = np.float32(1.0) / np.float32(0.5)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    ijk -= np.floor(pts[:, :3].min(0) * inv).astype(np.int64)
    div = np.floor(pts[:, :3].max(0) * inv).astype(np.int64) - np.floor(pts[:, :3].min(0) * inv).astype(np.int64) + 1
    ref_keys = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    assert np.array_equal(keys, ref_keys)
    uniq = np.unique(ref_keys)
    assert len(out) == len(uniq)
    means = np.stack([pts[ref_keys == k].astype(np.float64).mean(0) for k in uniq[:200]])
    assert np.max(np.abs(out[:200] - means)) < 1e-4


# ---------------------------------------------------------------- factor: finite differences, KATs, Huber, H/b
def _rand_factor(rng):
    p = rng.uniform(-20, 20, 3).astype(np.float32)
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)

    def pose7(scale):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        return np.concatenate([rng.uniform(-10, 10, 3) * scale, q])
    motion = np.concatenate([rng.normal(0, 3, 3), rng.normal(0, 0.3, 3), [0.004], rng.normal(0, 3, 3),
                             rng.normal(0, 0.3, 3), [-0.007]])
    return p, n, rng.uniform(-5, 5), 0.1, motion, pose7(1), pose7(1), pose7(0.02), 0.011


def _plus(pose, delta):
    """Right perturbation of pose_manifold.cc:41-47: t + dt, q (x) exp(dtheta)."""
    out = pose.copy()
    out[:3] += delta[:3]
    th = delta[3:6]
    a = np.linalg.norm(th)
    dq = np.concatenate([np.sin(a / 2) * th / a, [np.cos(a / 2)]]) if a > 0 else np.array([0, 0, 0, 1.0])
    x1, y1, z1, w1 = pose[3:]
    x2, y2, z2, w2 = dq
    out[3:] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
               w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    return out


def test_factor_jacobians_by_central_differences(oracle):
    rng = np.random.default_rng(11)
    h = 1e-6
    for trial in range(20):
        p, n, d, std, motion, pr, pc, ext, td = _rand_factor(rng)
        r, J = oracle.lidar_factor_evaluate(p, n, d, std, motion, pr, pc, ext, td)
        blocks = [pr, pc, ext]
        for b in range(3):
            num = np.zeros(6)
            for k in range(6):
                e = np.zeros(6)
                e[k] = h
                args = [x.copy() for x in blocks]
                args[b] = _plus(blocks[b], e)
                rp, _ = oracle.lidar_factor_evaluate(p, n, d, std, motion, *args, td, want_jac=None)
                args[b] = _plus(blocks[b], -e)
                rm, _ = oracle.lidar_factor_evaluate(p, n, d, std, motion, *args, td, want_jac=None)
                num[k] = (rp - rm) / (2 * h)
            assert np.max(np.abs(J[b][:6] - num)) < 1e-6 * max(1.0, np.max(np.abs(num)))
            assert J[b][6] == 0
        rp, _ = oracle.lidar_factor_evaluate(p, n, d, std, motion, pr, pc, ext, td + h, want_jac=None)
        rm, _ = oracle.lidar_factor_evaluate(p, n, d, std, motion, pr, pc, ext, td - h, want_jac=None)
        assert abs(J[3][0] - (rp - rm) / (2 * h)) < 1e-6 * max(1.0, abs(J[3][0]))


def test_factor_null_jacobian_blocks(oracle):
    rng = np.random.default_rng(12)
    p, n, d, std, motion, pr, pc, ext, td = _rand_factor(rng)
    r0, J0 = oracle.lidar_factor_evaluate(p, n, d, std, motion, pr, pc, ext, td)
    r1, J1 = oracle.lidar_factor_evaluate(p, n, d, std, motion, pr, pc, ext, td, want_jac=(1, 0, 1, 0))
    assert r0 == r1 and bits_equal(J0[0], J1[0]) and bits_equal(J0[2], J1[2])
    assert np.all(np.isnan(J1[1])) and np.all(np.isnan(J1[3]))      # untouched, as with Ceres null blocks


def test_factor_kat_coincident(oracle):
    pose = np.array([1.0, -2.0, 0.5, 0, 0, 0, 1.0])
    r, J = oracle.lidar_factor_evaluate(np.array([1, 2, 3], np.float32), [0, 0, 1.0], -3.0, 0.1, np.zeros(14), pose, pose,
                                        [0, 0, 0, 0, 0, 0, 1.0], 0.0)
    assert abs(r) < 1e-12 and np.array_equal(J[0][:3], -J[1][:3]) and np.allclose(J[1][:3], [0, 0, 10.0])


def test_huber_corrector(oracle):
    for r in (0.0, 0.3, -0.99, 1.0, 1.5, -4.0, 30.0):
        J = np.arange(1.0, 23.0)
        rc, Jc, rho0 = oracle.huber_correct(1.0, r, J)
        s = r * r
        if s <= 1:
            assert rc == r and bits_equal(Jc, J) and rho0 == s
        else:
            # Huber's linear branch has rho'' < 0, so the corrector takes the `rho[2] <= 0` path
            # (residual_block_info.cc:60-62): J and r are both scaled by sqrt(rho') = |r|^-1/2
            k = np.sqrt(1 / abs(r))
            assert abs(rho0 - (2 * abs(r) - 1)) < 1e-15
            assert np.max(np.abs(Jc - k * J)) < 1e-14 * np.max(np.abs(J)) and abs(rc - k * r) < 1e-14 * abs(r)


def test_H_b_accumulation(oracle):
    g = np.load(os.path.join(GOLD, "factors_small.npz"))
    for flags, suf in ((0, ""), (1, "_huber")):
        r, J, H, b, cost = oracle.factor_batch(g["point"], g["normal"], g["d"], g["std"], None, g["motion"], g["pose_ref"],
                                               g["pose_cur"], g["ext"], float(g["td"]), flags=flags)
        # regression pin on the committed fixture
        assert bits_equal(r, g["r" + suf]) and bits_equal(J, g["J" + suf]) and bits_equal(H, g["H" + suf])
        loc = [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 21]
        Jl = J[:, loc]
        assert block_rel(H, Jl.T @ Jl) < 1e-12 and block_rel(b, -Jl.T @ r) < 1e-12
        assert abs(cost[1] - r @ r) < 1e-9 * (r @ r)
    # residual-validity mask and constant-block flags
    valid = (np.arange(64) % 3 != 0).astype(np.uint8)
    r, J, H, b, cost = oracle.factor_batch(g["point"], g["normal"], g["d"], g["std"], valid, g["motion"], g["pose_ref"],
                                           g["pose_cur"], g["ext"], float(g["td"]), flags=8 | 16)
    assert np.all(r[valid == 0] == 0) and np.all(H[12:, :] == 0) and np.all(H[:, 12:] == 0) and np.all(b[12:] == 0)
    mask, cnt = oracle.chi2(g["point"], g["normal"], g["d"], g["std"], valid, g["motion"], g["pose_ref"], g["pose_cur"],
                            g["ext"], float(g["td"]), 3.841)
    assert cnt == int(np.sum((g["r"] ** 2 > 3.841) & (valid == 1))) and not mask[valid == 0].any()


def test_frame_golden(oracle):
    g = np.load(os.path.join(GOLD, "frame_small.npz"))
    und, R, t, fb = oracle.undistort(g["xyzi"], g["time"], g["ins_t"], g["ins_p"], g["ins_q"], g["R_bl"], g["t_bl"],
                                     float(g["stamp"]), float(g["td"]))
    assert bits_equal(und, g["undistorted"]) and bits_equal(R, g["pose_R"]) and bits_equal(t, g["pose_t"]) and fb == 0
    dec = oracle.decimate(und, int(g["filter_num"]))
    assert bits_equal(dec, g["decimated"]) and len(dec) == (len(und) + int(g["filter_num"]) - 1) // int(g["filter_num"])
    down, keys, _ = oracle.voxel_filter(dec, 0.5, want_keys=True)
    assert bits_equal(down, g["down"]) and bits_equal(keys, g["keys"])
    assert bits_equal(oracle.project(R, t, down), g["world"])
    assert np.abs(und[:, :3] - g["xyzi"][:, :3]).max() > 1e-3, "the fixture must carry real motion distortion"


# ---------------------------------------------------------------- CPU-baseline forms (bench.py's reference arm)
def test_window_forms_match_per_ref_calls(oracle):
    """orc_associate_window / orc_window_eval / orc_window_chi2 (kd-tree built once per keyframe, all pairs per call,
    chunked reduction as marginalization_info.cc:133-179) give what the per-ref oracle calls give."""
    rng = np.random.default_rng(3)
    q, n = 700, 3
    maps = [random_scene_cloud(rng, 6000 + 500 * i, 12.0) for i in range(n)]
    poses = [rand_pose(rng, 0.5) for _ in range(n)]
    lidar = random_scene_cloud(rng, q, 10.0)
    world = lidar.copy()
    trees = [oracle.MapTree(m) for m in maps]
    aw = oracle.associate_window(trees, [p[0] for p in poses], [p[1] for p in poses], world, lidar, threads=4)
    for i in range(n):
        o = oracle.associate(poses[i][0], poses[i][1], world, lidar, maps[i], knn_mode=0)
        assert np.array_equal(aw["nn_idx"][i], o["nn_idx"]) and np.array_equal(aw["type"][i], o["type"])
        assert aw["normal"][i].tobytes() == o["normal"].tobytes() and aw["d"][i].tobytes() == o["d"].tobytes()
        assert aw["num_features"][i] == o["num_features"] and aw["num_covered"][i] == o["num_covered"]
    assert aw["num_features"].sum() > 50

    def p7():
        qq = rng.normal(size=4)
        return np.concatenate([rng.uniform(-1, 1, 3), qq / np.linalg.norm(qq)])
    pose_refs = np.stack([p7() for _ in range(n)])
    pose_cur, ext, td = p7(), np.array([0.01, 0, 0.02, 0, 0, 0, 1.0]), 0.004
    motions = rng.normal(0, 0.3, (n, 14))
    motions[:, 6] = motions[:, 13] = 0.001
    pts = np.ascontiguousarray(lidar[:, :3])
    valid = np.ascontiguousarray((aw["type"] == 0).astype(np.uint8))
    H, b, cost = oracle.window_eval(pts, aw["normal"], aw["d"], 0.1, valid, motions, pose_refs, pose_cur, ext, td, flags=1,
                                    threads=4)
    outlier = np.zeros((n, q), np.uint8)
    cnt = oracle.window_chi2(pts, aw["normal"], aw["d"], 0.1, valid, motions, pose_refs, pose_cur, ext, td, outlier,
                             threshold=3.841, threads=4)
    for i in range(n):
        _, _, Ho, bo, co = oracle.factor_batch(pts, aw["normal"][i], aw["d"][i], np.full(q, 0.1), valid[i], motions[i],
                                               pose_refs[i], pose_cur, ext, td, flags=1)
        assert np.abs(H[i] - Ho).max() <= 1e-12 * np.abs(Ho).max() and np.abs(b[i] - bo).max() <= 1e-12 * np.abs(bo).max()
        assert np.abs(cost[i] - co).max() <= 1e-12 * np.abs(co).max()
        mask, c = oracle.chi2(pts, aw["normal"][i], aw["d"][i], np.full(q, 0.1), valid[i], motions[i], pose_refs[i],
                              pose_cur, ext, td, 3.841)
        assert np.array_equal(mask, outlier[i]) and c == cnt[i]
    for t in trees:
        t.close()
