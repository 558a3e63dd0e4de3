"""libfflins_host.so — the CPU side either side of the hot path: INS mechanization, the INS
window, IMU preintegration and its factor, the window solver with marginalization, converters and writers. CPU tests:
each piece against an independent numpy restatement of the reference (misc.cc:138-181, preintegration_base.cc:39-72),
finite differences, or first principles."""
import ctypes as C
import math

import numpy as np
import pytest

G = -9.80


def qmul(a, b):
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rv2q(v):
    a = np.linalg.norm(v)
    if a == 0:
        return np.array([0, 0, 0, 1.0])
    return np.concatenate([np.sin(a / 2) * v / a, [np.cos(a / 2)]])


def np_mechanization(pre, cur, st, g):
    """numpy restatement of MISC::insMechanization without earth rotation (misc.cc:138-181)."""
    (tp, dtp, thp, dvp), (tc, dtc, thc, dvc) = pre, cur
    p, q, v, bg, ba = st
    thc2, dvc2 = thc - dtc * bg, dvc - dtc * ba
    thp2, dvp2 = thp - dtp * bg, dvp - dtp * ba
    dvfb = dvc2 + 0.5 * np.cross(thc2, dvc2) + (np.cross(thp2, dvc2) + np.cross(dvp2, thc2)) / 12.0
    dth = thc2 + np.cross(thp2, thc2) / 12.0
    dvel = q2R(q) @ dvfb + np.array([0, 0, g]) * dtc
    q = qmul(q, rv2q(dth))
    q /= np.linalg.norm(q)
    p = p + dtc * v + 0.5 * dtc * dvel
    v = v + dvel
    return p, q, v, bg, ba


@pytest.fixture(scope="module")
def seq():
    from fflins import synth
    return synth.Sequence("robot")


def _imus(seq, k0, k1, **kw):
    from fflins import host
    t, dt, dth, dv = seq.imu_samples(k0, k1, gravity=G, **kw)
    return [host.Imu.make(t[i], dt[i], dth[i], dv[i]) for i in range(len(t))], (t, dt, dth, dv)


def _truth_state(seq, t_abs, bg=(0, 0, 0), ba=(0, 0, 0)):
    from fflins import host, synth
    p, q, v = seq.body_state(t_abs - synth.GPS_T0)
    return host.NavState.make(t_abs, p, q, v, bg, ba)


def test_exports():
    from fflins import host
    L = host.load()
    missing = [e for e in host.EXPORTS if not hasattr(L, e)]
    assert not missing, missing


def test_mechanization_matches_numpy_and_truth(seq):
    from fflins import host, synth
    bg, ba = np.array([1e-4, -2e-4, 5e-5]), np.array([2e-3, 1e-3, -3e-3])
    imus, (t, dt, dth, dv) = _imus(seq, 100, 700, bg=bg, ba=ba)     # 3 s at 200 Hz
    st = _truth_state(seq, t[0], bg, ba)
    cfg = host.ins_config(G)
    ref = (np.array(st.p[:]), np.array(st.q[:]), np.array(st.v[:]), bg, ba)
    for i in range(1, len(imus)):
        host.mechanization(cfg, imus[i - 1], imus[i], st)
        ref = np_mechanization((t[i - 1], dt[i - 1], dth[i - 1], dv[i - 1]), (t[i], dt[i], dth[i], dv[i]), ref, G)
        if i % 50 == 0:
            assert np.max(np.abs(np.array(st.p[:]) - ref[0])) < 1e-12 and np.max(np.abs(np.array(st.q[:]) - ref[1])) < 1e-14
            assert np.max(np.abs(np.array(st.v[:]) - ref[2])) < 1e-12
    assert st.time == t[-1]
    tp, tq, tv = seq.body_state(t[-1] - synth.GPS_T0)
    # 3 s of strapdown integration of exact increments: millimetre / micro-radian level against the analytic trajectory
    assert np.linalg.norm(np.array(st.p[:]) - tp) < 2e-3
    assert np.linalg.norm(np.array(st.v[:]) - tv) < 2e-3
    assert 2 * np.linalg.norm(qmul(np.array([-tq[0], -tq[1], -tq[2], tq[3]]), np.array(st.q[:]))[:3]) < 1e-5


def test_ins_window_export_series_and_redo(seq):
    from fflins import host, synth
    imus, (t, *_rest) = _imus(seq, 0, 400)
    w = host.InsWindow(host.ins_config(G))
    w.reset(imus[0], _truth_state(seq, t[0]))
    for m in imus[1:]:
        w.add(m)
    assert len(w) == 401
    wt, wp, wq = w.export()
    assert np.array_equal(wt, t)
    truth = np.stack([seq.body_state(x - synth.GPS_T0)[0] for x in t[::40]])
    assert np.max(np.linalg.norm(wp[::40] - truth, axis=1)) < 2e-3
    # getImuSeriesFromTo: interpolated ends, the pieces add up to the raw increments, last time = end
    start, end = t[100] + 0.0021, t[180] + 0.0037
    s = w.series(start, end)
    assert s[0].time == pytest.approx(start) and s[-1].time == end
    assert sum(m.dt for m in s[1:]) == pytest.approx(end - start, abs=1e-9)   # (absolute GPS times: 6e-11 ulp)
    tot = np.sum([np.array(m.dtheta[:]) for m in s[1:]], axis=0)
    raw = np.sum([np.array(imus[k].dtheta[:]) for k in range(102, 181)], axis=0)
    part0 = np.array(imus[101].dtheta[:]) * (t[101] - start) / imus[101].dt
    part1 = np.array(imus[181].dtheta[:]) * (end - t[180]) / imus[181].dt
    assert np.allclose(tot, raw + part0 + part1, atol=1e-15)
    # nodes within MINIMUM_TIME_INTERVAL (1e-4 s) of an epoch take the epoch itself, no interpolation (misc.cc:240-264)
    s2 = w.series(t[100] + 5e-5, t[180] + 5e-5)
    assert [m.time for m in s2[:-1]] == list(t[100:180]) and s2[-1].time == t[180] + 5e-5
    assert np.array_equal(np.array(s2[-1].dtheta[:]), np.array(imus[180].dtheta[:]))
    # redoInsMechanization from a corrected state at an IMU epoch: the window behind it follows, stale epochs go
    st_k = _truth_state(seq, t[300] + 0.002)        # between two epochs: the bracketing increment is split (need = 2)
    st_k.p[0] += 0.05
    before = w.export()[1][-1].copy()
    assert w.redo(st_k, 200) == 0
    wt2, wp2, _ = w.export()
    assert len(wt2) == 401 - (301 - 200) and wt2[-1] == t[-1]
    assert wp2[-1][0] - before[0] == pytest.approx(0.05, abs=2e-3)
    w.close()


def _preint(seq, k0, k1, bg=(0, 0, 0), ba=(0, 0, 0), noise=None):
    from fflins import host
    imus, (t, *_r) = _imus(seq, k0, k1)
    st0 = _truth_state(seq, t[0], bg, ba)
    pre = host.Preintegration(noise or host.imu_noise(g=G), imus[0], st0)
    for m in imus[1:]:
        pre.add(m)
    return pre, st0, _truth_state(seq, t[-1], bg, ba), t


def test_preintegration_delta_and_residual(seq):
    pre, st0, st1, t = _preint(seq, 200, 240)       # 0.2 s
    assert pre.delta_time() == pytest.approx(t[-1] - t[0], abs=1e-9)
    # the integrated current state tracks the truth, the residual of the true states is ~0 before whitening
    cur = pre.current_state()
    assert np.linalg.norm(np.array(cur.p[:]) - np.array(st1.p[:])) < 1e-4
    r, J = pre.evaluate(st0.pose7(), st0.mix9(), st1.pose7(), st1.mix9())
    cov, jac = pre.matrices()
    assert np.all(np.linalg.eigvalsh(cov) > 0) and np.allclose(cov, cov.T)
    W = np.linalg.cholesky(np.linalg.inv(cov)).T          # LLT(cov^-1).matrixL().transpose()
    raw = np.linalg.solve(W, r)
    assert np.linalg.norm(raw[:3]) < 1e-4 and np.linalg.norm(raw[3:6]) < 1e-3 and np.linalg.norm(raw[6:9]) < 1e-5
    assert np.all(raw[9:] == 0)
    assert np.all(J[0][:, 6] == 0) and np.all(J[2][:, 6] == 0)
    # bias Jacobian of the preintegrated quantities (first-order): re-integrating at a shifted bias moves delta p by dp_dbg dbg
    from fflins import host
    dbg = np.array([2e-4, -1e-4, 3e-4])
    pre2, *_ = _preint(seq, 200, 240, bg=dbg)
    d0, d2 = pre.delta_state(), pre2.delta_state()
    pred = jac[0:3, 9:12] @ dbg
    assert np.linalg.norm((np.array(d2.p[:]) - np.array(d0.p[:])) - pred) < 0.15 * np.linalg.norm(pred) + 1e-9   # (first-order phi = I + F dt propagation, as in the reference)


def test_preintegration_factor_jacobians_by_finite_differences(seq):
    pre, st0, st1, _ = _preint(seq, 300, 330)
    rng = np.random.default_rng(3)
    pose0, mix0, pose1, mix1 = st0.pose7(), st0.mix9(), st1.pose7(), st1.mix9()
    pose1[:3] += rng.normal(0, 0.01, 3)          # a non-trivial linearisation point
    mix0[3:6] += rng.normal(0, 1e-4, 3)
    mix1[:3] += rng.normal(0, 0.01, 3)
    r0, J = pre.evaluate(pose0, mix0, pose1, mix1)

    def plus_pose(x, d):
        out = x.copy()
        out[:3] += d[:3]
        q = qmul(x[3:], rv2q(d[3:6]))
        out[3:] = q / np.linalg.norm(q)
        return out
    h = 1e-6
    args = [pose0, mix0, pose1, mix1]
    for blk, (x, Jb) in enumerate(zip(args, J)):
        nloc = 6 if blk in (0, 2) else 9
        num = np.zeros((15, nloc))
        for j in range(nloc):
            d = np.zeros(nloc)
            d[j] = h
            a = list(args)
            a[blk] = plus_pose(x, d) if blk in (0, 2) else x + d
            rp, _ = pre.evaluate(*a, jac=False)
            d[j] = -h
            a[blk] = plus_pose(x, d) if blk in (0, 2) else x + d
            rm, _ = pre.evaluate(*a, jac=False)
            num[:, j] = (rp - rm) / (2 * h)
        scale = max(np.abs(num).max(), 1.0)
        assert np.abs(Jb[:, :nloc] - num).max() < 2e-5 * scale, f"block {blk}"


def test_frame_omega(seq):
    pre, st0, st1, t = _preint(seq, 400, 440)
    from fflins import synth
    om = pre.frame_omega(0.1, 1.0 / seq.imu_hz, np.zeros(3))
    truth = seq.body_rate(np.array([t[-1] - synth.GPS_T0 - 0.025]))[0]
    assert np.linalg.norm(om - truth) < 0.02 * max(np.linalg.norm(truth), 1e-3) + 1e-4


# ---- solver -------------------------------------------------------------------------------------------------------
def _build_chain(seq, n_nodes, step=40, k0=200, perturb=0.0, rng=None, pose_std=(0.01,) * 3 + (0.002,) * 3,
                 mix_std=(0.05,) * 3 + (1e-4,) * 3 + (1e-3,) * 3):
    """A window of IMU-only nodes on the analytic trajectory; node states start from the mechanized chain."""
    from fflins import host
    noise = host.imu_noise(g=G)
    imus, (t, *_r) = _imus(seq, k0, k0 + step * (n_nodes - 1))
    s = host.Solver(window_size=50)
    first = _truth_state(seq, t[0])
    if perturb and rng is not None:
        first.p[0] += perturb
    s.set_first_node(t[0], first, pose_std, mix_std)
    state = first
    for n in range(1, n_nodes):
        seg = imus[step * (n - 1): step * n + 1]
        pre = host.Preintegration(noise, seg[0], state)
        for m in seg[1:]:
            pre.add(m)
        state = pre.current_state()
        s.add_node(t[step * n], pre)
    return s, t


def test_solver_imu_only_chain_is_stationary_and_pulls_back(seq):
    """IMU factors + the initial priors only: the mechanized chain is already the minimiser (cost ~ 0); a displaced first
    node is pulled back to its prior and the chain follows."""
    s, t = _build_chain(seq, 5)
    sm = s.optimize()
    assert sm.cost_final < 1e-6
    rng = np.random.default_rng(1)
    s2, _ = _build_chain(seq, 5, perturb=0.0)
    # displace every node by the same 5 cm: the IMU factors stay consistent, only the prior on node 0 disagrees
    # (set_first_node remembers the undisplaced prior)
    from fflins import host
    L = host.load()
    st0, _ = s2.node(0)
    sm2 = s2.optimize()
    p_ref = np.array([s2.node(i)[0].p[:] for i in range(5)])
    truth = np.stack([seq.body_state(t[40 * i] - 345600.0)[0] for i in range(5)])
    assert np.max(np.linalg.norm(p_ref - truth, axis=1)) < 5e-3


def _lidar_callbacks(oracle, feats, pts, motions):
    """Synthetic LiDAR share for CPU tests: per-pair blocks from the oracle's factor_batch (what fflins_window_eval returns).
    feats / motions are keyed by the reference keyframe id."""
    outl = {k: np.zeros(len(pts), np.uint8) for k in feats}

    def evaluate(ref_ids, pose_refs, pose_cur, ext, td, flags):
        n = len(ref_ids)
        H, b, c = np.zeros((n, 19, 19)), np.zeros((n, 19)), np.zeros((n, 2))
        for i, rid in enumerate(ref_ids):
            ft = feats[rid]
            valid = ((ft["type"] == 0) & (outl[rid] == 0)).astype(np.uint8)
            _, _, H[i], b[i], c[i] = oracle.factor_batch(pts, ft["normal"], ft["d"], np.full(len(pts), 0.1), valid, motions[rid],
                                                         pose_refs[i], pose_cur, ext, td, flags=flags)
        return H, b, c

    def cull(ref_ids, pose_refs, pose_cur, ext, td, thr):
        cnt = []
        for i, rid in enumerate(ref_ids):
            ft = feats[rid]
            valid = ((ft["type"] == 0) & (outl[rid] == 0)).astype(np.uint8)
            mask, c = oracle.chi2(pts, ft["normal"], ft["d"], np.full(len(pts), 0.1), valid, motions[rid], pose_refs[i],
                                  pose_cur, ext, td, thr)
            outl[rid] |= mask
            cnt.append(c)
        return np.array(cnt)
    return evaluate, cull


def _plane_world(rng, n):
    """Points on the walls / floor / ceiling of a 20 x 20 x 6 m box with their plane parameters n . p + d = 0 (world)."""
    pts, nrm, d = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    for k in range(n):
        ax = int(rng.integers(0, 3))
        p = rng.uniform(-9, 9, 3)
        p[2] = rng.uniform(0.5, 5.5)
        val = float(rng.choice([-10.0, 10.0])) if ax < 2 else float(rng.choice([0.0, 6.0]))
        p[ax] = val
        nv = np.zeros(3)
        nv[ax] = 1.0
        pts[k], nrm[k], d[k] = p, nv, -val
    return pts, nrm, d


def _lidar_window(seq, oracle, n_nodes=4, refs=(0, 1), step=40, k0=200, n_pts=400, seed=7):
    """IMU chain of n_nodes time nodes; keyframes at `refs` and at the last node, whose points lie on the planes of a box.
    Features = exact planes expressed in each reference keyframe's LiDAR frame (what the association would produce)."""
    from fflins import host
    from fflins.window_solver import lidar_pose
    rng = np.random.default_rng(seed)
    noise = host.imu_noise(g=G)
    imus, (tt, *_r) = _imus(seq, k0, k0 + step * (n_nodes - 1))
    ext7 = np.array([0.02, -0.01, 0.03, 0, 0, 0, 1.0])
    s = host.Solver(window_size=50)
    first = _truth_state(seq, tt[0])
    s.set_first_node(tt[0], first, (0.01,) * 3 + (0.002,) * 3, (0.05,) * 3 + (1e-4,) * 3 + (1e-3,) * 3)
    s.set_extrinsic(ext7, 0.0)
    state = first
    for n in range(1, n_nodes):
        seg = imus[step * (n - 1): step * n + 1]
        pre = host.Preintegration(noise, seg[0], state)
        for m in seg[1:]:
            pre.add(m)
        state = pre.current_state()
        s.add_node(tt[step * n], pre)
    cur = n_nodes - 1
    for n in list(refs) + [cur]:
        s.set_node_lidar(n, 10 + n)
    truth = {i: _truth_state(seq, tt[step * i]) for i in range(n_nodes)}
    world, nrm_w, d_w = _plane_world(rng, n_pts)
    Rc, tc = lidar_pose(truth[cur].pose7(), ext7)
    pts_cur = np.ascontiguousarray(((world - tc) @ Rc).astype(np.float32))      # R^T (p - t)
    feats, motions = {}, {}
    for n in refs:
        R, tl = lidar_pose(truth[n].pose7(), ext7)
        feats[10 + n] = dict(type=np.zeros(n_pts, np.int8), normal=np.ascontiguousarray(nrm_w @ R),
                             d=np.ascontiguousarray(d_w + nrm_w @ tl))
        motions[10 + n] = np.zeros(14)
    evaluate, cull = _lidar_callbacks(oracle, feats, pts_cur, motions)
    return s, truth, evaluate, cull


def test_solver_with_lidar_blocks_recovers_pose(seq, oracle):
    """IMU chain + LiDAR factors of the newest node against two reference nodes (oracle blocks standing in for the GPU):
    a 5 cm error of the newest node's position is removed."""
    from fflins import host
    s, truth, evaluate, cull = _lidar_window(seq, oracle)
    st3, lid = s.node(3)
    assert lid == 13
    st3.p[0] += 0.05
    st3.p[1] -= 0.03
    s.set_node_state(3, st3)
    sm = s.optimize(evaluate, cull)
    assert sm.cost_final < 0.05 * sm.cost_initial and sm.successful[0] >= 1
    p3 = np.array(s.node(3)[0].p[:])
    assert np.linalg.norm(p3 - np.array(truth[3].p[:])) < 3e-3
    # extrinsic and td stay put while the window is not full (optimizer.cc:418-435)
    e, td = s.get_extrinsic()
    assert np.array_equal(e, np.array([0.02, -0.01, 0.03, 0, 0, 0, 1.0])) and td == 0.0


def test_marginalization_prior_keeps_the_optimum(seq, oracle):
    """After marginalizing the oldest node the remaining window's optimum does not move: the prior carries what the
    eliminated IMU factor, initial priors and LiDAR factors said about the kept states."""
    s, truth, evaluate, cull = _lidar_window(seq, oracle, seed=11)
    s.optimize(evaluate, cull)
    before = [np.array(s.node(i)[0].p[:]) for i in range(1, 4)]
    v_before = [np.array(s.node(i)[0].v[:]) for i in range(1, 4)]
    s.marginalize(evaluate)
    assert s.num_nodes() == 3 and s.node(0)[1] == 11
    J0, e0 = s.prior()
    assert J0.shape[0] > 0 and J0.shape[1] >= 15
    # disturb the kept states, solve again with the prior in place of node 0
    st, _ = s.node(0)
    st.p[0] += 0.02
    s.set_node_state(0, st)
    sm = s.optimize(evaluate, cull)
    after = [np.array(s.node(i)[0].p[:]) for i in range(3)]
    v_after = [np.array(s.node(i)[0].v[:]) for i in range(3)]
    assert max(np.linalg.norm(a - b) for a, b in zip(after, before)) < 2e-3
    assert max(np.linalg.norm(a - b) for a, b in zip(v_after, v_before)) < 5e-3


# ---- converters and writers --------------------------------------------------------------------------------------
def test_converters_and_writers():
    from fflins import host
    n = 6
    pts = (host.LivoxPoint * n)()
    vals = [(0, 1.0, 0, 0, 10, 0x00, 0), (1000, 0.1, 0, 0, 20, 0x00, 0),       # too near (min 0.5)
            (2000, 3.0, 4.0, 0, 30, 0x10, 1), (3000, 5.0, 0, 0, 40, 0x20, 0),   # echo 2: dropped before the time range
            (4000, 300.0, 0, 0, 50, 0x00, 0),                                    # too far
            (5000, 2.0, 0, 0, 60, 0x00, 7)]                                      # line >= scan_line
    for p, v in zip(pts, vals):
        p.offset_time, p.x, p.y, p.z, p.reflectivity, p.tag, p.line = v
    rec, s, e = host.convert_livox(pts, 1_000_000_000, 6, 0.5, 250.0)
    assert len(rec) == 2 and list(rec["intensity"]) == [10.0, 30.0]
    assert rec["time"][1] == pytest.approx(1.000002) and s == pytest.approx(1.0) and e == pytest.approx(1.000004)
    assert rec.dtype.itemsize == 32
    xyzi = np.array([[1, 0, 0, 5], [0, 0.2, 0, 6], [0, 0, 7, 8]], np.float32)
    r1, s1, e1 = host.convert_generic(xyzi, np.array([0.0, 0.01, 0.02]), 100.0, 0, 0.5, 250.0)
    assert len(r1) == 2 and r1["time"][1] == pytest.approx(100.02) and (s1, e1) == (100.0, pytest.approx(100.02))
    r2, *_ = host.convert_generic(xyzi, np.array([0.0, 1e7, 2e7]), 100.0, 1, 0.5, 250.0)
    assert r2["time"][1] == pytest.approx(100.02)
    r3, *_ = host.convert_generic(xyzi, np.array([1.6e9, 1.6e9 + 1, 1.6e9 + 2]), 0.0, 2, 0.5, 250.0, to_gps_time=True)
    sec = 1.6e9 + 18 - 315964800
    assert r3["time"][0] == pytest.approx(sec - math.floor(sec / 604800) * 604800)
    st = host.NavState.make(345600.123456789, (1, 2, 3), (0, 0, 0, 1), (0, 0, 0), (1e-5, 0, 0), (0, 1e-4, 0))
    line = host.format_trajectory(st)
    assert line == "%-15.9f %-15.9f %-15.9f %-15.9f %-15.9f %-15.9f %-15.9f %-15.9f \n" % (345600.123456789, 1, 2, 3, 0, 0, 0, 1)
    err = host.format_imu_error(st).split()
    assert float(err[1]) == pytest.approx(1e-5 * 180 / math.pi * 3600, rel=1e-6) and float(err[5]) == pytest.approx(10.0)
    assert host.format_statistics([1.0, 2.5]) == "%-15.9f %-15.9f \n" % (1.0, 2.5)
