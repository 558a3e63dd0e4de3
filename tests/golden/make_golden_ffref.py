"""Generates tests/golden/ffref_reference_sources.npz: outputs of the REFERENCE'S OWN source files for the hot path -
factors/lidar_factor.cc, common/misc.cc, common/rotation.cc, lidar/pointcloud.cc - compiled where they lie under
/root/reference into oracle/_ref/libff_ref.so (oracle/Makefile `ref`; Eigen / PCL / Ceres / TBB / glog are stand-in headers
under oracle/refshim/, see refshim/mini_eigen.h for what that pins). Run in the build container, where /root/reference exists:

  make -C oracle && python tests/golden/make_golden_ffref.py

Contents (inputs + reference outputs), all seeded:
  und_*      a 1,500-point frame through the per-point loop of lidarFrameProcessing (MISC::getPoseFromInsWindow +
             PointCloudCommon::relativeProjection), 40 of the point times outside the INS window (fallback branch)
  pw_*       getPoseFromInsWindow / getInsWindowIndex at 64 times incl. the window's edges and entries
  fac_*      LidarFactor::Evaluate: 96 residuals, residual + the four Jacobian blocks
  rb_*       ResidualBlockInfo::Evaluate (the reference's corrector, residual_block_info.cc:49-77) over the same factors
             with ceres::HuberLoss(1.0) (the loss itself is the stand-in of oracle/refshim/ceres/ceres.h)
  pl_*       planeEstimation on 400 five-point sets (planes, noisy planes, scatter, near-collinear)
  prj_*      pointCloudProjection of 2,000 points
  mech_*     MISC::insMechanization over 200 IMU epochs (iswithearth = false)
  mg_*       MarginalizationInfo::marginalization (marginalization_info.cc:43-260) over the 240 LidarFactor residual blocks
             of one (oldest reference, newest) pair with HuberLoss(1.0), the reference pose marginalized: the prior J0, e0
             on [pose_cur | ext | td] (eigenvectors come from the stand-in's Jacobi solver: compare J0^T J0 and J0^T e0)
  pre_*      PreintegrationNormal (preintegration_base.cc, preintegration_normal.cc) over 60 IMU epochs: integrated and
             preintegrated states, covariance, bias Jacobian, and PreintegrationFactor::Evaluate (whitened residual and the
             four Jacobian blocks) at a displaced pair of states
  iw_*       an INS window of 401 epochs built by insMechanization, MISC::getImuSeriesFromTo between two off-epoch times and
             MISC::redoInsMechanization from an updated state with 50 reserved epochs (misc.cc:186-336)
  mgf_*      MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91) of that prior at displaced parameter blocks
  cv_*       LidarConverter (ROS/lidar/lidar_converter.cc:36-232): a Livox CustomMsg (tag / line / range filters) and
             Velodyne / Ouster / Hesai clouds with their three time conventions, with and without the GPS-time conversion
  pm_*       PoseManifold::Plus / Minus and their Jacobians (pose_manifold.cc:35-83) on 64 poses and increments

and tests/golden/ffref_lidar_map.npz: the reference's own lidar::LidarMap (lidar_map.cc + the vendored ikd-Tree) driven
over 3 keyframes x 2 frames of the `robot` sequence (2,000 points per frame): per frame the undistorted cloud and pose,
per keyframe the down-sampled / world / map clouds, and per (reference, newest) pair of every association the
LidarFeatures (type, unit normal, norm inverse, the five nearest map points in the kd-tree's order) and the counters.
The VoxelGrid under it is the stand-in of oracle/refshim/pcl/filters/voxel_grid.h (PCL is not installed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
sys.path.insert(0, ROOT)

from fflins import synth  # noqa: E402
from oracle import pyref as R  # noqa: E402


def rand_pose7(rng, span=10.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([rng.uniform(-span, span, 3), q])


def main():
    assert R.available(), "build oracle/_ref/libff_ref.so first (make -C oracle, needs /root/reference)"
    rng = np.random.default_rng(20261020)
    out = {}

    rng = np.random.default_rng(20261020 + 1)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- undistortion loop
    seq = synth.Sequence("lili", n_points=1500, ins_entries=200)
    fr = seq.frame(7)
    time = np.array(fr["time"], np.float64)
    lo, hi = fr["ins_t"][0] - fr["td"], fr["ins_t"][-1] - fr["td"]
    time[:20] = lo - rng.uniform(1e-4, 0.05, 20)       # before the window: index 0 -> window.back() pose
    time[20:40] = hi + rng.uniform(0.0, 0.05, 20)      # at / behind the last entry: ditto
    und, pR, pt, fb = R.undistort(fr["xyzi"], time, fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"],
                                  fr["td"])
    assert fb == 40
    out.update(und_xyzi=np.asarray(fr["xyzi"], np.float32), und_time=time, und_ins_t=fr["ins_t"], und_ins_p=fr["ins_p"],
               und_ins_q=fr["ins_q"], und_R_bl=np.asarray(seq.R_bl, np.float64), und_t_bl=np.asarray(seq.t_bl, np.float64),
               und_stamp=float(fr["stamp"]), und_td=float(fr["td"]), und_out=und, und_pose_R=pR, und_pose_t=pt,
               und_fallbacks=fb, und_filter_num=seq.filter_num)

    rng = np.random.default_rng(20261020 + 2)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- pose / index queries
    it = np.asarray(fr["ins_t"], np.float64)
    times = np.concatenate([[it[0] - 1e-3, it[0], it[-1], it[-1] + 1e-3, np.nextafter(it[-1], 0), it[1], it[len(it) // 2]],
                            rng.uniform(it[0], it[-1], 57)])
    pw_ok, pw_R, pw_t, pw_idx = [], [], [], []
    for tm in times:
        ok, Rm, tv = R.pose_from_ins_window(it, fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, tm)
        pw_ok.append(ok), pw_R.append(Rm), pw_t.append(tv)
        pw_idx.append(R.ins_window_index(it, tm))
    out.update(pw_times=times, pw_ok=np.array(pw_ok), pw_R=np.array(pw_R), pw_t=np.array(pw_t), pw_index=np.array(pw_idx))

    rng = np.random.default_rng(20261020 + 3)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- LidarFactor::Evaluate
    n = 96
    fac = dict(point=rng.uniform(-30, 30, (n, 3)).astype(np.float32), normal=np.zeros((n, 3)), d=rng.uniform(-5, 5, n),
               std=np.full(n, 0.1), motion=np.zeros((n, 14)), pose_ref=np.zeros((n, 7)), pose_cur=np.zeros((n, 7)),
               ext=np.zeros((n, 7)), td=rng.normal(size=n) * 0.01, r=np.zeros(n), J=np.zeros((n, 22)))
    for i in range(n):
        nv = rng.normal(size=3)
        fac["normal"][i] = nv / np.linalg.norm(nv)
        fac["motion"][i] = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 0.3, [rng.normal() * 0.01],
                                           rng.normal(size=3), rng.normal(size=3) * 0.3, [rng.normal() * 0.01]])
        fac["pose_ref"][i], fac["pose_cur"][i], fac["ext"][i] = rand_pose7(rng), rand_pose7(rng), rand_pose7(rng, 0.3)
        if i % 3 == 0:      # nearby poses, as in a window
            fac["pose_cur"][i, :3] = fac["pose_ref"][i, :3] + rng.normal(size=3) * 0.3
            q = fac["pose_ref"][i, 3:] + rng.normal(size=4) * 0.02
            fac["pose_cur"][i, 3:] = q / np.linalg.norm(q)
        r, J = R.lidar_factor_evaluate(fac["point"][i], fac["normal"][i], fac["d"][i], fac["std"][i], fac["motion"][i],
                                       fac["pose_ref"][i], fac["pose_cur"][i], fac["ext"][i], fac["td"][i])
        fac["r"][i] = r
        fac["J"][i] = np.concatenate(J)
    out.update({"fac_" + k: v for k, v in fac.items()})

    rng = np.random.default_rng(20261020 + 4)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- ResidualBlockInfo::Evaluate with ceres::HuberLoss(1.0): the plane offsets are moved so that the residuals spread
    # over both branches of the loss (|r| = 0, 0.4, 3, 12 sigma)
    rb_d, rb_r, rb_J = fac["d"].copy(), np.zeros(n), np.zeros((n, 22))
    for i in range(n):
        rb_d[i] = fac["d"][i] - fac["std"][i] * fac["r"][i] + fac["std"][i] * [0.0, 0.4, -3.0, 12.0][i % 4]
        rb_r[i], rb_J[i] = R.residual_block_evaluate(fac["point"][i], fac["normal"][i], rb_d[i], fac["std"][i],
                                                     fac["motion"][i], fac["pose_ref"][i], fac["pose_cur"][i], fac["ext"][i],
                                                     fac["td"][i], 1.0)
    assert (np.abs(rb_r) > 1).sum() >= n // 3 and (np.abs(rb_r) < 1).sum() >= n // 3
    out.update(rb_d=rb_d, rb_r=rb_r, rb_J=rb_J)

    rng = np.random.default_rng(20261020 + 5)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- planeEstimation
    m = 400
    pts = np.zeros((m, 5, 4), np.float32)
    for i in range(m):
        c = rng.uniform(-20, 20, 3)
        nv = rng.normal(size=3)
        nv /= np.linalg.norm(nv)
        u = rng.normal(size=(5, 3))
        u -= np.outer(u @ nv, nv)
        mode = i % 5
        if mode == 3:
            p = rng.uniform(-5, 5, (5, 3)) + c
        elif mode == 4:
            a = rng.uniform(-3, 3, 5)
            p = c + np.outer(a, rng.normal(size=3)) + 1e-3 * rng.normal(size=(5, 3))
        else:
            p = c + u * rng.uniform(0.2, 1.5) + np.outer(rng.normal(size=5) * [0.001, 0.03, 0.09][mode], nv)
        pts[i, :, :3] = p
        pts[i, :, 3] = rng.uniform(0, 255, 5)
    pl_ok, pl_n, pl_ninv, pl_dist = np.zeros(m, bool), np.zeros((m, 3)), np.zeros(m), np.zeros((m, 5))
    for i in range(m):
        pl_ok[i], pl_n[i], pl_ninv[i], pl_dist[i] = R.plane_estimation(pts[i], 0.1)
    assert 0 < pl_ok.sum() < m
    out.update(pl_points=pts, pl_ok=pl_ok, pl_normal=pl_n, pl_norm_inverse=pl_ninv, pl_distance=pl_dist)

    rng = np.random.default_rng(20261020 + 6)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- pointCloudProjection
    src = rng.uniform(-80, 80, (2000, 4)).astype(np.float32)
    pose = rand_pose7(rng, 200.0)
    Rw, _ = R.state_to_pose(np.zeros(3), pose[3:], np.eye(3), np.zeros(3))
    out.update(prj_src=src, prj_R=Rw, prj_t=pose[:3], prj_dst=R.project(Rw, pose[:3], src), prj_q=pose[3:])

    rng = np.random.default_rng(20261020 + 7)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- insMechanization
    k = 200
    g = np.array([0.0, 0.0, -9.8])
    imu = np.zeros((k + 1, 8))
    imu[:, 0] = 100.0 + 0.005 * np.arange(k + 1)
    imu[:, 1] = 0.005
    imu[:, 2:5] = rng.normal(size=(k + 1, 3)) * 0.002 + [0.001, -0.0005, 0.002]
    imu[:, 5:8] = rng.normal(size=(k + 1, 3)) * 0.01 + [0.0, 0.0, 9.8 * 0.005]
    st = np.concatenate([[1.0, -2.0, 0.5], rand_pose7(rng)[3:], [0.4, -0.2, 0.05], [1e-4, -2e-4, 5e-5], [2e-3, 1e-3, -3e-3]])
    states = np.zeros((k + 1, 16))
    states[0] = st
    tm = imu[0, 0]
    for i in range(1, k + 1):
        st, tm = R.ins_mechanization(g, imu[i - 1], imu[i], st, tm)
        states[i] = st
    out.update(mech_gravity=g, mech_imu=imu, mech_states=states)

    rng = np.random.default_rng(20261020 + 8)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- PreintegrationNormal through PreintegrationFactor::Evaluate (preintegration_base.cc, preintegration_normal.cc)
    from fflins import host
    G = host.GRAVITY
    pseq = synth.Sequence("robot", n_points=100)
    bg, ba = np.array([1e-4, -2e-4, 5e-5]), np.array([2e-3, 1e-3, -3e-3])
    t, dt, dth, dv = pseq.imu_samples(200, 260, gravity=G, bg=bg, ba=ba)        # 0.3 s at 200 Hz
    p0, q0, v0 = pseq.body_state(t[0] - synth.GPS_T0)
    p1, q1, v1 = pseq.body_state(t[-1] - synth.GPS_T0)
    nz = host.imu_noise(g=G)
    noise = np.array([nz.acc_vrw, nz.gyr_arw, nz.gyr_bias_std, nz.acc_bias_std, nz.corr_time, nz.gravity])
    imu = np.column_stack([t, dt, dth, dv])
    state0 = np.concatenate([p0, q0, v0, bg, ba])
    pre = R.RefPreintegration(noise, imu[0], state0, t[0])
    for row in imu[1:]:
        pre.add(row)
    cur, dlt, dtime = pre.states()
    cov, jac = pre.matrices()
    pose0, mix0 = np.concatenate([p0, q0]), np.concatenate([v0, bg + rng.normal(0, 1e-4, 3), ba])
    pose1, mix1 = np.concatenate([p1 + rng.normal(0, 0.01, 3), q1]), np.concatenate([v1 + rng.normal(0, 0.01, 3), bg, ba + rng.normal(0, 1e-3, 3)])
    r, J = pre.evaluate(pose0, mix0, pose1, mix1)
    out.update(pre_noise=noise, pre_imu=imu, pre_state0=state0, pre_current=cur, pre_delta=dlt, pre_delta_time=dtime, pre_cov=cov,
               pre_jac=jac, pre_pose0=pose0, pre_mix0=mix0, pre_pose1=pose1, pre_mix1=mix1, pre_r=r, pre_J0=J[0], pre_J1=J[1],
               pre_J2=J[2], pre_J3=J[3])
    del pre

    rng = np.random.default_rng(20261020 + 9)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- the INS window: getImuSeriesFromTo (both ends between epochs) and redoInsMechanization (misc.cc:186-336)
    t, dt, dth, dv = pseq.imu_samples(0, 400, gravity=G)
    wimu = np.column_stack([t, dt, dth, dv])
    p0, q0, v0 = pseq.body_state(t[0] - synth.GPS_T0)
    ws0 = np.concatenate([p0, q0, v0, np.zeros(6)])
    s_a, s_b = t[120] + 0.0013, t[180] + 0.0027
    tu = t[300] + 0.002
    pu, qu, vu = pseq.body_state(tu - synth.GPS_T0)
    upd = np.concatenate([pu + 0.01, qu, vu - 0.02, [1e-4, 0, -1e-4], [1e-3, 2e-3, 0]])
    ok, series, wt, wstates = R.ins_window_scenario([0, 0, G], wimu, ws0, s_a, s_b, upd, tu, 50)
    assert ok and len(series) > 50 and len(wt) < len(wimu)
    out.update(iw_gravity=G, iw_imu=wimu, iw_state0=ws0, iw_series_from=s_a, iw_series_to=s_b, iw_series=series, iw_updated=upd,
               iw_updated_time=tu, iw_reserved=50, iw_times=wt, iw_states=wstates)

    rng = np.random.default_rng(20261020 + 10)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- LidarConverter (ROS/lidar/lidar_converter.cc): Livox CustomMsg and Velodyne / Ouster / Hesai clouds
    n = 1500
    lv = np.zeros(n, R.LIVOX19)
    lv["offset_time"] = np.sort(rng.integers(0, 100_000_000, n))
    xyz = rng.normal(size=(n, 3)) * [20, 20, 3]
    xyz[::50] = 0.0                                     # invalid returns: inside the minimum range
    lv["x"], lv["y"], lv["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    lv["reflectivity"] = rng.integers(0, 255, n)
    lv["tag"] = rng.choice([0x00, 0x10, 0x20, 0x30, 0x05, 0x15], n)
    lv["line"] = rng.integers(0, 8, n)
    sec, nsec = 1_690_000_123, 456_789_012
    out.update(cv_livox=lv, cv_sec=sec, cv_nsec=nsec)
    for gps in (0, 1):
        rec, s0, e0 = R.convert_livox(lv, sec, nsec, 6, 1.0, 100.0, bool(gps))
        out.update({f"cv_livox_out{gps}": rec, f"cv_livox_span{gps}": np.array([s0, e0])})
    gx = np.concatenate([rng.normal(size=(n, 3)) * [30, 30, 4], rng.uniform(0, 255, (n, 1))], 1).astype(np.float32)
    gx[::40, :3] = 0.1
    stamp = sec + 1e-9 * nsec
    times = {2: np.sort(rng.uniform(0, 0.1, n)).astype(np.float32).astype(np.float64),
             3: np.sort(rng.integers(0, 100_000_000, n)).astype(np.float64), 4: stamp + np.sort(rng.uniform(0, 0.1, n))}
    out.update(cv_xyzi=gx)
    for kind in (2, 3, 4):
        out[f"cv_time{kind}"] = times[kind]
        for gps in (0, 1):
            rec, s0, e0 = R.convert_generic(kind, gx, times[kind], sec, nsec, 1.0, 100.0, bool(gps))
            out.update({f"cv_kind{kind}_out{gps}": rec, f"cv_kind{kind}_span{gps}": np.array([s0, e0])})

    rng = np.random.default_rng(20261020 + 11)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- PoseManifold::Plus / Minus (pose_manifold.cc:35-83)
    pm_x = np.array([rand_pose7(rng) for _ in range(64)])
    pm_d = rng.normal(size=(64, 6)) * np.array([1, 1, 1, 0.3, 0.3, 0.3])
    pm_d[:8, 3:] *= 1e-6        # small rotations
    pm_d[8, 3:] = 0.0           # none: AngleAxis of a zero vector
    pm_plus = np.array([R.pose_plus(a, b) for a, b in zip(pm_x, pm_d)])
    pm_minus = np.array([R.pose_minus(a, b) for a, b in zip(pm_plus, pm_x)])
    jp, jm = R.pose_jacobians(pm_x[0])
    out.update(pm_x=pm_x, pm_delta=pm_d, pm_plus=pm_plus, pm_minus=pm_minus, pm_plus_jacobian=jp, pm_minus_jacobian=jm)

    rng = np.random.default_rng(20261020 + 12)   # (one generator per section: adding a section leaves the others unchanged)
    # ---- MarginalizationInfo over the LidarFactor blocks of one (oldest reference, newest) pair (marginalization_info.cc)
    m = 240
    mg = dict(point=rng.uniform(-20, 20, (m, 3)).astype(np.float32), normal=rng.normal(size=(m, 3)), d=rng.uniform(-1, 1, m),
              motion=rng.normal(size=14) * np.array([1, 1, 1, .3, .3, .3, .01] * 2), pose_ref=rand_pose7(rng), ext=rand_pose7(rng, 0.2),
              td=0.004)
    mg["normal"] /= np.linalg.norm(mg["normal"], axis=1, keepdims=True)
    mg["pose_cur"] = mg["pose_ref"].copy()
    mg["pose_cur"][:3] += rng.normal(size=3) * 0.3
    q = mg["pose_ref"][3:] + rng.normal(size=4) * 0.02
    mg["pose_cur"][3:] = q / np.linalg.norm(q)
    for i in range(m):      # residuals of 0 .. 3 sigma around the planes: both branches of the loss
        r, _ = R.lidar_factor_evaluate(mg["point"][i], mg["normal"][i], mg["d"][i], 0.1, mg["motion"], mg["pose_ref"],
                                       mg["pose_cur"], mg["ext"], mg["td"], want_jac=None)
        mg["d"][i] += 0.1 * (rng.normal() * 1.5 - r)
    J0, e0 = R.marginalize_lidar_pair(mg["point"], mg["normal"], mg["d"], 0.1, mg["motion"], mg["pose_ref"], mg["pose_cur"],
                                      mg["ext"], mg["td"], 1.0)
    out.update({"mg_" + k: v for k, v in mg.items()})
    out.update(mg_J0=J0, mg_e0=e0)
    # MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91) of that prior at displaced blocks; the extrinsic's
    # quaternion is handed over with its sign flipped (dq.w() < 0 branch of the pose difference)
    ecur = R.pose_plus(mg["pose_cur"], rng.normal(size=6) * np.array([.05, .05, .05, .02, .02, .02]))
    eext = R.pose_plus(mg["ext"], rng.normal(size=6) * np.array([.01, .01, .01, .005, .005, .005]))
    eext[3:] *= -1
    etd = mg["td"] + 0.002
    J0b, e0b, eres, ejac = R.marginalize_lidar_pair(mg["point"], mg["normal"], mg["d"], 0.1, mg["motion"], mg["pose_ref"],
                                                    mg["pose_cur"], mg["ext"], mg["td"], 1.0, evaluate_at=(ecur, eext, etd))
    assert np.array_equal(J0b, J0) and np.array_equal(e0b, e0)
    out.update(mgf_pose_cur=ecur, mgf_ext=eext, mgf_td=etd, mgf_residuals=eres, mgf_jacobian=ejac)

    path = os.path.join(HERE, "ffref_reference_sources.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    lidar_map_golden()


MAP_SEQ = dict(name="robot", n_points=2000, n_keys=3, frames_per_key=2)


def lidar_map_golden():
    seq = synth.Sequence(MAP_SEQ["name"], n_points=MAP_SEQ["n_points"])
    ref = R.RefLidarMap(seq.filter_num, min_keyframe_translation=0.4)
    out = dict(filter_num=seq.filter_num, n_keys=MAP_SEQ["n_keys"], frames_per_key=MAP_SEQ["frames_per_key"],
               R_bl=np.asarray(seq.R_bl, np.float64), t_bl=np.asarray(seq.t_bl, np.float64))
    fi = 0
    for k in range(MAP_SEQ["n_keys"]):
        for j in range(MAP_SEQ["frames_per_key"]):
            fr = seq.frame(fi)
            args = (fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"], fr["td"])
            ref.frame(*args)
            und, pR, pt, fb = R.undistort(*args)        # the same loop over the same reference functions
            assert fb == 0
            out[f"f{fi}_und"], out[f"f{fi}_R"], out[f"f{fi}_t"] = und, pR, pt
            fi += 1
        ref.keyframe()
        cur = fi - 1
        Rk, tk = ref.pose(-1)
        assert Rk.tobytes() == out[f"f{cur}_R"].tobytes() and tk.tobytes() == out[f"f{cur}_t"].tobytes()
        assert ref.cloud(-1, ref.CLOUD_LIDAR_FULL).tobytes() == out[f"f{cur}_und"][::seq.filter_num].tobytes()
        out[f"k{k}_down"] = ref.cloud(-1, ref.CLOUD_LIDAR)
        out[f"k{k}_world"] = ref.cloud(-1, ref.CLOUD_WORLD)
        out[f"k{k}_map"] = ref.cloud(-1, ref.CLOUD_MAP)
        out[f"k{k}_map_full_n"] = len(ref.cloud(-1, ref.CLOUD_MAP_FULL))
        if k >= 1:
            out[f"k{k}_counts"] = np.array(ref.counts())
            for i in range(k):
                ft = ref.features(i)
                for name in ("type", "normal", "d", "nearest", "n_nearest", "std"):
                    out[f"k{k}_r{i}_{name}"] = ft[name]
    # ---- keyframeSelection (lidar_map.cc:75-98) of 48 candidate poses / stamps against the newest keyframe, with the
    # AT128 configuration's thresholds (0.4 m; 10 deg and 0.475 s are constants of lidar_map.h)
    rng = np.random.default_rng(7)
    Rk, tk = ref.pose(-1)
    stamp_k = seq.frame(fi - 1)["stamp"]
    cand_R, cand_t, cand_stamp, cand_key, cand_stats = [], [], [], [], []
    for i in range(48):
        ang = [0.02, 0.1, 0.17, 0.18, 0.3, 3.0][i % 6] * (1 if i % 2 else -1)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        q = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])
        dR, _ = R.state_to_pose(np.zeros(3), q, np.eye(3), np.zeros(3))
        Rc = Rk @ dR
        tc = tk + rng.normal(size=3) * [0.05, 0.2, 0.24, 0.5][i % 4]
        st = stamp_k + [0.1, 0.3, 0.47, 0.48, 0.6][i % 5]
        ok, stats = ref.keyframe_selection(Rc, tc, st)
        cand_R.append(Rc), cand_t.append(tc), cand_stamp.append(st), cand_key.append(ok), cand_stats.append(stats)
    assert 5 < sum(cand_key) < 48
    out.update(ks_key_R=Rk, ks_key_t=tk, ks_key_stamp=stamp_k, ks_R=np.array(cand_R), ks_t=np.array(cand_t),
               ks_stamp=np.array(cand_stamp), ks_is_key=np.array(cand_key), ks_stats=np.array(cand_stats))
    ref.close()
    path = os.path.join(HERE, "ffref_lidar_map.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
