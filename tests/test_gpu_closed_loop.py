"""Closed-loop trajectory check (BASELINE.json: "end-to-end trajectory on the same synthetic sequence
must differ by less than 1 mm ATE"), in the reduced form the scope allows without Ceres / IMU factors
(SURVEY.md §8f row 1): every new keyframe's pose is solved by the same dense LM (fflins.window_solver)
against the window of older keyframes; the loop is closed (solved pose -> setPose + re-projection ->
next association). The loop runs twice on the same raw frames: once with every LiDAR operation on the
GPU through the C-ABI, once with the CPU oracle. ATE = RMS position difference over the keyframes.
A third run replaces the host LM by fflins_window_solve (the whole 5 + chi2 + 10 schedule on the device, one
launch per keyframe) on a full window of 10 references."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K = 3          # frames per keyframe
N_FRAMES = 24  # 8 keyframes
WINDOW = 5


def _init_guess(seq, stamp, k):
    rng = np.random.default_rng(9000 + k)             # identical perturbation for both backends
    return seq.imu_pose7(stamp, rng, 0.02, 0.003)     # 2 cm, ~0.17 deg off the INS pose


def run_gpu(seq, frames, K=K, WINDOW=WINDOW, device_solve=False):
    from fflins import api
    from fflins.window_solver import WindowSolver, lidar_pose
    c = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=WINDOW))
    ext = seq.ext7()
    nonkeys, poses7, traj, stats = [], {}, [], []
    for i, fr in enumerate(frames):
        c.frame_process(i, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"], fr["td"])
        if (i + 1) % K:
            nonkeys.append(i)
            continue
        c.keyframe_merge(i, nonkeys)
        for nk in nonkeys:
            c.release(nk)
        nonkeys = []
        c.keyframe_add(i)
        c.set_motion(i, fr["v"], fr["omega"], fr["td"])
        ids = c.keyframe_ids()
        if len(ids) == 1:
            poses7[i] = seq.imu_pose7(fr["stamp"])
        else:
            st = c.associate()
            refs = ids[:-1]
            pose_refs = np.stack([poses7[r] for r in refs])
            flags = api.HUBER | api.CONST_REF | api.CONST_EXT | api.CONST_TD

            def evaluate(pose, e, td):
                o = c.window_eval(refs, pose_refs, pose, e, td, flags)
                return o["H"], o["b"], o["cost"]

            def cull(pose, e, td):
                return c.window_chi2(refs, pose_refs, pose, e, td)
            if device_solve:
                opt = c.solve_options(flags=flags, function_tolerance=1e-14, parameter_tolerance=1e-14)
                _, pc, _, _, rep = c.window_solve(refs, pose_refs, _init_guess(seq, fr["stamp"], len(traj)), ext, fr["td"], options=opt)
                assert rep.status == 0 and rep.successful[0] >= 1
                sol = dict(pose=pc, outliers=int(sum(rep.n_outliers[:len(refs)])))
            else:
                sol = WindowSolver(evaluate, cull).solve(_init_guess(seq, fr["stamp"], len(traj)), ext, fr["td"])
            poses7[i] = sol["pose"]
            stats.append((int(sum(st.num_features[:st.n_refs])), sol["outliers"]))
        R, t = lidar_pose(poses7[i], ext)
        c.reproject(i, R, t)                          # optimizer.cc:203-218
        traj.append(poses7[i][:3].copy())
        if len(ids) > WINDOW:
            c.keyframe_remove_oldest()
    c.close()
    return np.array(traj), stats


def run_cpu(seq, frames, O, K=K, WINDOW=WINDOW):
    from fflins.window_solver import WindowSolver, lidar_pose
    ext = seq.ext7()
    nonkeys, keys, traj, stats = [], [], [], []
    for i, fr in enumerate(frames):
        und, R, t, _ = O.undistort(fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl,
                                   fr["stamp"], fr["td"], threads=8)
        dec = O.decimate(und, seq.filter_num)
        down = O.voxel_filter(dec, 0.5)
        f = dict(und=und, down=down, R=R, t=t, fr=fr)
        if (i + 1) % K:
            nonkeys.append(f)
            continue
        clouds = [und]
        for g in nonkeys:
            Rr, tr = O.relative_pose(R, t, g["R"], g["t"])
            clouds.append(O.project(Rr, tr, g["und"]))
        f["map"] = O.voxel_filter(np.concatenate(clouds), 0.5)
        nonkeys = []
        keys.append(f)
        if len(keys) == 1:
            f["pose7"] = seq.imu_pose7(fr["stamp"])
        else:
            world = O.project(R, t, down)
            refs = keys[:-1]
            feats = [O.associate(r["R"], r["t"], world, down, r["map"], threads=8) for r in refs]
            outl = [np.zeros(len(down), np.uint8) for _ in refs]
            pts = np.ascontiguousarray(down[:, :3])
            std = np.full(len(down), 0.1)
            mc = np.concatenate([fr["v"], fr["omega"], [fr["td"]]])
            motions = [np.concatenate([r["fr"]["v"], r["fr"]["omega"], [r["fr"]["td"]], mc]) for r in refs]
            flags = 1 | 2 | 8 | 16

            def evaluate(pose, e, td):
                Hs, bs, cs = [], [], []
                for r, ft, ol, mo in zip(refs, feats, outl, motions):
                    valid = ((ft["type"] == 0) & (ol == 0)).astype(np.uint8)
                    _, _, H, b, cst = O.factor_batch(pts, ft["normal"], ft["d"], std, valid, mo, r["pose7"], pose, e, td,
                                                     flags=flags, threads=8)
                    Hs.append(H)
                    bs.append(b)
                    cs.append(cst)
                return np.stack(Hs), np.stack(bs), np.stack(cs)

            def cull(pose, e, td):
                n = []
                for r, ft, ol, mo in zip(refs, feats, outl, motions):
                    valid = ((ft["type"] == 0) & (ol == 0)).astype(np.uint8)
                    mask, cnt = O.chi2(pts, ft["normal"], ft["d"], std, valid, mo, r["pose7"], pose, e, td, threads=8)
                    ol |= mask
                    n.append(cnt)
                return np.array(n)
            sol = WindowSolver(evaluate, cull).solve(_init_guess(seq, fr["stamp"], len(traj)), ext, fr["td"])
            f["pose7"] = sol["pose"]
            stats.append((int(sum(ft["num_features"] for ft in feats)), sol["outliers"]))
        f["R"], f["t"] = lidar_pose(f["pose7"], ext)
        traj.append(f["pose7"][:3].copy())
        if len(keys) > WINDOW:
            keys.pop(0)
    return np.array(traj), stats


def test_closed_loop_ate_below_1mm(oracle):
    from fflins import synth
    seq = synth.Sequence("robot")
    frames = [seq.frame(i) for i in range(N_FRAMES)]
    g_traj, g_stats = run_gpu(seq, frames)
    c_traj, c_stats = run_cpu(seq, frames, oracle)
    assert g_traj.shape == c_traj.shape == (N_FRAMES // K, 3)
    ate = float(np.sqrt(np.mean(np.sum((g_traj - c_traj) ** 2, axis=1))))
    truth = np.stack([seq.imu_pose7(frames[i]["stamp"])[:3] for i in range(K - 1, N_FRAMES, K)])
    err_truth = float(np.sqrt(np.mean(np.sum((g_traj - truth) ** 2, axis=1))))
    print(f"closed loop: ATE(GPU vs CPU oracle) = {ate * 1e3:.6f} mm over {len(g_traj)} keyframes; "
          f"RMS vs ground truth = {err_truth * 1e3:.2f} mm; features/outliers GPU {g_stats} CPU {c_stats}")
    assert ate < 1e-3, "trajectory must agree within 1 mm"
    assert err_truth < 0.05, "the solve must pull the 2 cm initial error down to the noise level"
    assert all(s[0] > 200 for s in g_stats)


def test_closed_loop_with_the_device_solve_window10(oracle):
    """The same loop with each keyframe's solve as ONE fflins_window_solve call, on a full window of 10 references
    (13 keyframes of 3 frames), against the CPU-oracle loop with the host LM: two different LM controls (Ceres-style
    trust region on the device, plain lambda damping in window_solver.py) must land on the same optimum."""
    from fflins import synth
    seq = synth.Sequence("robot")
    k, window, n_frames = 3, 10, 39
    frames = [seq.frame(i) for i in range(n_frames)]
    g_traj, g_stats = run_gpu(seq, frames, K=k, WINDOW=window, device_solve=True)
    c_traj, c_stats = run_cpu(seq, frames, oracle, K=k, WINDOW=window)
    assert g_traj.shape == c_traj.shape == (n_frames // k, 3)
    ate = float(np.sqrt(np.mean(np.sum((g_traj - c_traj) ** 2, axis=1))))
    truth = np.stack([seq.imu_pose7(frames[i]["stamp"])[:3] for i in range(k - 1, n_frames, k)])
    err_truth = float(np.sqrt(np.mean(np.sum((g_traj - truth) ** 2, axis=1))))
    print(f"closed loop, device solve, window 10: ATE(device-solved GPU loop vs host-LM CPU oracle loop) = {ate * 1e3:.6f} mm over "
          f"{len(g_traj)} keyframes; RMS vs ground truth = {err_truth * 1e3:.2f} mm; features/outliers GPU {g_stats[-3:]} CPU {c_stats[-3:]}")
    assert ate < 1e-3, "trajectory must agree within 1 mm"
    assert err_truth < 0.05
