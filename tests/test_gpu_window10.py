"""Parity on the configuration bench.py times: the AT128-shaped sequence with a FULL sliding window
(11 keyframes x 5 frames -> 10 reference keyframes), plus a 15-keyframe window (14 references, the
largest the context accepts), compared with the oracle for EVERY (ref, query) pair.

  association   idx / d2 / type / normal bit-exact              lidar_map.cc:326-459
  factors       H, b, cost <= 1e-9 per block                    lidar_factor.cc:41-118, marginalization_info.cc:181-210
  chi2 cull     outlier mask identical                          optimizer.cc:515-548
and each factor / chi2 result once through the RESIDENT evaluation kernel (mailbox) and once through
the launched kernel (the profiler forces it), which must agree bit for bit.
Also: removeNewestLidarFrame, LidarFeature::setOutlier round trip, the single-keyframe association
after an asynchronous frame, the kNN radius bound, and the unstaged
end-to-end flip counts on at128 / ouster64."""
import time

import numpy as np
import pytest

from util import bits_equal, block_rel, build_window, window_params

pytestmark = pytest.mark.gpu


def _make_window(name, n_keys, frames_per_key, window_size=10):
    from fflins import api, synth
    from oracle import pyoracle
    seq = synth.Sequence(name)
    c = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=window_size))
    keys = build_window(c, pyoracle, seq, n_keys=n_keys, frames_per_key=frames_per_key)
    return c, seq, keys


@pytest.fixture(scope="module")
def window10():
    """The benched configuration: at128, 11 keyframes of 5 frames (10 references)."""
    c, seq, keys = _make_window("at128", 11, 5)
    yield c, seq, keys
    c.close()


@pytest.fixture(scope="module")
def window15():
    """14 references (robot frames: the pair count is what is being exercised)."""
    c, seq, keys = _make_window("robot", 15, 2, window_size=15)
    yield c, seq, keys
    c.close()


def _motion(key):
    return np.concatenate([key["fr"]["v"], key["fr"]["omega"], [key["fr"]["td"]]])


def _check_association(c, oracle, keys):
    st = c.associate()
    cur = keys[-1]
    assert st.n_refs == len(keys) - 1 and st.n_queries == len(cur["down"])
    planes = 0
    for i, ref in enumerate(keys[:-1]):
        o = oracle.associate(ref["R"], ref["t"], cur["world"], cur["down"], ref["map"], threads=8)
        g = c.features(ref["id"])
        assert np.array_equal(g["nn_idx"], o["nn_idx"]), f"ref {i}: kNN index 5-tuples must be bit-exact"
        assert bits_equal(g["nn_d2"], o["nn_d2"])
        assert np.array_equal(g["covered"], o["covered"]) and np.array_equal(g["type"], o["type"])
        assert bits_equal(g["normal"], o["normal"]) and bits_equal(g["d"], o["d"])
        assert st.num_features[i] == o["num_features"] and st.num_covered[i] == o["num_covered"]
        planes += o["num_features"]
    return st, planes


def _eval_both_paths(c, refs, pose_refs, pose_cur, ext, td, flags):
    """window_eval through the resident kernel and through a launch; returns (resident, launched)."""
    s0 = c.eval_stats()
    a = c.window_eval(refs, pose_refs, pose_cur, ext, td, flags)
    s1 = c.eval_stats()
    c.profile_enable(True)      # the per-stage profiler needs stream-ordered kernels: forces the launched path
    b = c.window_eval(refs, pose_refs, pose_cur, ext, td, flags)
    c.profile_enable(False)
    s2 = c.eval_stats()
    assert s2["launched_evals"] == s1["launched_evals"] + 1
    resident = s1["resident_evals"] == s0["resident_evals"] + 1
    return a, b, resident


def _check_factors(c, oracle, seq, keys, flags, seed):
    rng = np.random.default_rng(seed)
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    cur = keys[-1]
    refs = [k["id"] for k in keys[:-1]]
    a, b, resident = _eval_both_paths(c, refs, pose_refs, pose_cur, ext, td, flags)
    for k in ("H", "b", "cost", "n_res"):
        assert bits_equal(a[k], b[k]), f"resident and launched evaluation differ in {k}"
    q = len(cur["down"])
    total = 0
    for i, ref in enumerate(keys[:-1]):
        f = c.features(ref["id"])
        valid = ((f["type"] == 0) & (f["outlier"] == 0)).astype(np.uint8)
        motion = np.concatenate([_motion(ref), _motion(cur)])
        _, _, Ho, bo, co = oracle.factor_batch(cur["down"][:, :3], f["normal"], f["d"], np.full(q, 0.1), valid, motion,
                                               pose_refs[i], pose_cur, ext, td, flags=flags)
        assert a["n_res"][i] == valid.sum()
        total += int(valid.sum())
        if valid.sum() == 0:
            assert not a["H"][i].any()
            continue
        assert block_rel(a["H"][i], Ho) <= 1e-9 and block_rel(a["b"][i], bo) <= 1e-9
        assert block_rel(a["cost"][i], co) <= 1e-9
        assert np.array_equal(a["H"][i], a["H"][i].T)
    return total, resident


def _check_chi2(c, oracle, seq, keys, seed):
    rng = np.random.default_rng(seed)
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    cur = keys[-1]
    refs = [k["id"] for k in keys[:-1]]
    q = len(cur["down"])
    before = [c.features(r) for r in refs]
    want = []
    for i, ref in enumerate(keys[:-1]):
        f = before[i]
        valid = ((f["type"] == 0) & (f["outlier"] == 0)).astype(np.uint8)
        motion = np.concatenate([_motion(ref), _motion(cur)])
        want.append(oracle.chi2(cur["down"][:, :3], f["normal"], f["d"], np.full(q, 0.1), valid, motion, pose_refs[i],
                                pose_cur, ext, td, 3.841))
    s0 = c.eval_stats()
    cnt = c.window_chi2(refs, pose_refs, pose_cur, ext, td, 3.841)          # resident (after a settled association)
    s1 = c.eval_stats()
    for i, r in enumerate(refs):
        mask, n = want[i]
        assert np.array_equal(c.features(r)["outlier"], mask | before[i]["outlier"]) and cnt[i] == n
    # LidarFeature::setOutlier round trip: restore the flags, cull again through the launched kernel
    for i, r in enumerate(refs):
        c.set_outlier(r, before[i]["outlier"])
        assert np.array_equal(c.features(r)["outlier"], before[i]["outlier"])
    c.profile_enable(True)
    cnt2 = c.window_chi2(refs, pose_refs, pose_cur, ext, td, 3.841)
    c.profile_enable(False)
    assert np.array_equal(cnt, cnt2)
    for i, r in enumerate(refs):
        assert np.array_equal(c.features(r)["outlier"], want[i][0] | before[i]["outlier"])
    assert c.eval_stats()["launched_evals"] == s1["launched_evals"] + 1
    return int(np.sum(cnt)), s1["resident_evals"] == s0["resident_evals"] + 1


# ------------------------------------------------------------------------------------------------ at128, 10 refs
def test_window10_association_every_pair(window10, oracle):
    c, seq, keys = window10
    st, planes = _check_association(c, oracle, keys)
    assert st.n_refs == 10
    print(f"at128 window10: Q={st.n_queries} M={[len(k['map']) for k in keys]} plane features={planes}")
    assert planes > 1000


@pytest.mark.parametrize("flags", [1, 0, 1 | 8 | 16])
def test_window10_factor_blocks(window10, oracle, flags):
    c, seq, keys = window10
    c.associate()
    total, resident = _check_factors(c, oracle, seq, keys, flags, 60 + flags)
    assert total > 1000
    assert resident, "after a settled association the evaluation must be served by the resident kernel"


def test_window10_chi2(window10, oracle):
    c, seq, keys = window10
    c.associate()
    n_out, resident = _check_chi2(c, oracle, seq, keys, 71)
    assert resident
    print(f"at128 window10: chi2 culled {n_out}")


def test_window10_resident_retires_and_restarts(window10):
    """The resident kernel retires after its idle period (1 ms) and is restarted by the next evaluation; results
    are unchanged and no evaluation is lost across the boundary."""
    c, seq, keys = window10
    c.associate()
    rng = np.random.default_rng(5)
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    refs = [k["id"] for k in keys[:-1]]
    first = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    first = {k: v.copy() for k, v in first.items()}
    s0 = c.eval_stats()
    for _ in range(3):
        time.sleep(0.02)
        again = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
        assert bits_equal(again["H"], first["H"]) and bits_equal(again["b"], first["b"])
    for _ in range(50):      # back to back: served without a restart
        again = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    assert bits_equal(again["H"], first["H"])
    s1 = c.eval_stats()
    assert s1["resident_launches"] >= s0["resident_launches"] + 2
    assert s1["resident_evals"] + s1["launched_evals"] == s0["resident_evals"] + s0["launched_evals"] + 53
    assert s1["resident_evals"] >= s0["resident_evals"] + 50


# ------------------------------------------------------------------------------------------------ 14 refs
def test_window15_all_pairs(window15, oracle):
    c, seq, keys = window15
    st, planes = _check_association(c, oracle, keys)
    assert st.n_refs == 14 and planes > 200
    for flags in (1, 1 | 2, 1 | 4 | 16):
        total, resident = _check_factors(c, oracle, seq, keys, flags, 90 + flags)
        assert total > 200 and resident
    _check_chi2(c, oracle, seq, keys, 93)


# ------------------------------------------------------------------------------------------------ ADVICE r1
def test_associate_single_keyframe_after_async_frame(oracle):
    """A lone keyframe (no reference yet) right after an asynchronous frame: the association launches nothing, so the
    frame's counts must be adopted behind a real synchronisation, never from a stale pinned slot."""
    from fflins import api, synth
    seq = synth.Sequence("robot")
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    # occupy every pinned count slot once with a different (larger) frame, so stale values exist
    big = synth.Sequence("lili")
    for i in range(70):
        fr = big.frame(i % 2)
        c.frame_process(5000 + i, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], big.R_bl, big.t_bl,
                        fr["stamp"], fr["td"])
        c.release(5000 + i)
    fr = seq.frame(0)
    want = c.frame_process(1, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl,
                           fr["stamp"], fr["td"])
    c.release(1)
    c.frame_process(2, fr["xyzi"], fr["time"], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"],
                    fr["td"], sync=False)
    c.keyframe_merge(2, [], sync=False)
    c.keyframe_add(2)
    st = c.associate()
    assert st.n_refs == 0
    info = c.frame_info(2)
    assert info.n_down == want.n_down and info.n_fallback == want.n_fallback
    assert st.n_queries == want.n_down
    c.close()


def test_knn_search_radius_bound(ctx):
    """The packed queue fields of the kNN hold offsets up to 495 cells: a leaf too small for the search radius is
    rejected instead of silently wrapping."""
    from fflins import api
    with pytest.raises(api.FflinsError) as e:
        api.Context(api.default_config(downsample_size=0.005))
    assert e.value.code == -1
    ok = api.Context(api.default_config(downsample_size=0.02))   # 5 m / 2 cm = 251 cells: accepted
    ok.close()
    mp = np.zeros((10, 4), np.float32)
    with pytest.raises(api.FflinsError):
        ctx.knn5(mp, mp, max_dist=400.0)


# ------------------------------------------------------------------------------------------------ unstaged flips
@pytest.mark.parametrize("name,frames_per_key", [("at128", 5), ("ouster64", 5)])
def test_end_to_end_flip_counts(oracle, name, frames_per_key):
    """Whole call sequence with NO staging (raw frames into both pipelines) over 25 frames: undistortion is allowed
    1 ulp of fp32 difference, which may flip a voxel assignment or a validity test now and then. Reports and bounds
    the discrete flips and the drift of the normal equations."""
    from fflins import api, synth
    from fflins.pipeline import Session
    from oracle.cpu_pipeline import CpuSession
    seq = synth.Sequence(name)
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    gs = Session(c, seq, frames_per_key=frames_per_key, n_eval=(1, 1))
    cs = CpuSession(seq, threads=8, frames_per_key=frames_per_key, n_eval=(1, 1), knn_mode=1)
    flips_q = 0
    for i in range(25):
        fr = seq.frame(i)
        fid, info = gs.process_frame(fr)
        f = cs.process_frame(fr)
        flips_q += abs(info.n_down - len(f["down"]))
        gs.after_frame(fid, fr)
        cs.after_frame(f)
    ids = c.keyframe_ids()
    assert len(ids) == len(cs.keys) == 25 // frames_per_key
    flips_type = flips_idx = feats = 0
    for i, (r, ck) in enumerate(zip(ids[:-1], cs.keys[:-1])):
        g = c.features(r)
        assert c.frame_info(r).n_map == len(ck["map"]) or abs(c.frame_info(r).n_map - len(ck["map"])) <= 2
        if len(g["type"]) == len(ck["feat"]["type"]):
            flips_type += int(np.sum(g["type"] != ck["feat"]["type"]))
            flips_idx += int(np.sum(np.any(g["nn_idx"] != ck["feat"]["nn_idx"], axis=1)))
            feats += len(g["type"])
        Hg, Ho = gs.last["eval"]["H"][i], cs.last["eval"][i][2]
        assert block_rel(Hg, Ho) < 1e-3, "end-to-end H drift"
    print(f"{name}: unstaged flips over 25 frames: voxel-count {flips_q}, feature-type {flips_type}, "
          f"kNN 5-tuple {flips_idx} of {feats} (ref, query) pairs")
    assert flips_q <= 3 and flips_type <= max(5, feats // 500) and flips_idx <= max(10, feats // 200)
    c.close()


# ------------------------------------------------------------------------------------------------ native caller of bench.py
def test_native_session_driver_issues_the_same_calls():
    """tools/session_driver.cc (what bench.py times) against fflins/pipeline.py (the readable statement of the call sequence):
    the same frames through two contexts must leave bit-identical evaluation blocks and association counts."""
    import ctypes
    import bench
    import fflins
    from fflins import api, synth
    from fflins.pipeline import Session
    seq = synth.Sequence("robot", n_points=8000)
    frames = [seq.frame(i) for i in range(12)]
    for fr in frames:
        fr["pose7"] = seq.imu_pose7(fr["stamp"])
    K, steps = 3, 4
    # Python session, synchronous frames on host buffers
    c1 = fflins.Context(fflins.default_config(point_filter_num=seq.filter_num, window_size=3))
    s1 = Session(c1, seq, K, (2, 1), sync_frames=False)
    for i in range(K * steps):
        fid, _ = s1.process_frame(frames[i])
        s1.after_frame(fid, frames[i])
    want = {k: np.array(s1.last["eval"][k], copy=True) for k in ("H", "b", "cost", "n_res")}
    a1 = s1.last["assoc"]
    # native driver, device-resident inputs
    import torch
    c2 = fflins.Context(fflins.default_config(point_filter_num=seq.filter_num, window_size=3))
    ns = bench.NativeSession(c2, seq, K, (2, 1), api.HUBER)
    keep, items = [], []
    for fr in frames:
        x = torch.from_numpy(np.ascontiguousarray(fr["xyzi"])).cuda()
        t = torch.from_numpy(np.ascontiguousarray(fr["time"])).cuda()
        ins = [np.ascontiguousarray(fr[k], np.float64) for k in ("ins_t", "ins_p", "ins_q")]
        keep.append((x, t, ins))
        items.append({"pts": x.data_ptr(), "time": t.data_ptr(), "ins_t": ins[0], "ins_p": ins[1], "ins_q": ins[2], "n": len(fr["xyzi"]),
                      "w": len(ins[0]), "stamp": fr["stamp"], "td": fr["td"], "v": fr["v"], "omega": fr["omega"], "pose7": fr["pose7"]})
    torch.cuda.synchronize()
    ns.table("dev", items)
    ns.run("dev", steps)
    assert ns.last_assoc() == (a1.n_refs, a1.n_queries)
    n = a1.n_refs
    H, b, cost = np.zeros((16, 361)), np.zeros((16, 19)), np.zeros((16, 2))
    ns.L.drv_last_blocks.argtypes = [ctypes.c_void_p] * 4
    assert ns.L.drv_last_blocks(ns.h, H.ctypes.data, b.ctypes.data, cost.ctypes.data) == n
    assert bits_equal(H[:n].reshape(want["H"][:n].shape), want["H"][:n]) and bits_equal(b[:n].reshape(want["b"][:n].shape), want["b"][:n])
    assert ns.last_nres() == int(np.sum(want["n_res"][:n])) and ns.last_nres() > 100
    ns.close()
    c1.close()
    c2.close()


def test_voxel_chain_fallback_still_bit_exact():
    """The multi-kernel voxel chain (FFLINS_VOXEL_CHAIN=1: the fallback behind the cooperative filter) against the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FFLINS_VOXEL_CHAIN="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "vox_bench.py"), "robot", "1"], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bit_exact_vs_oracle=True" in out.stdout and "voxel.radix_sort" in out.stdout


_QUIET_SCRIPT = r'''
import os, sys
import numpy as np
root = sys.argv[1]
sys.path.insert(0, os.path.join(root, "ff-lins_b200")); sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from fflins import api, synth
from oracle import pyoracle
from util import bits_equal, build_window, window_params
seq = synth.Sequence("robot")
c = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=10))
keys = build_window(c, pyoracle, seq, n_keys=6, frames_per_key=2)
c.associate()
refs = [k["id"] for k in keys[:-1]]
rng = np.random.default_rng(3)
n_res = 0
for it in range(40):
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    s0 = c.eval_stats()
    a = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    a = {k: np.array(a[k], copy=True) for k in ("H", "b", "cost", "n_res")}
    s1 = c.eval_stats()
    assert s1["resident_evals"] == s0["resident_evals"] + 1 or it == 0, (it, s0, s1)
    c.profile_enable(True)                      # profiling forces the launched kernel
    b = c.window_eval(refs, pose_refs, pose_cur, ext, td, 1)
    c.profile_enable(False)
    for k in ("H", "b", "cost", "n_res"):
        assert bits_equal(a[k], np.array(b[k])), (it, k)
    n_res += int(a["n_res"].sum())
assert n_res > 1000
print("QUIET_OK", c.eval_stats())
c.close()
'''


def test_resident_quiet_mode_matches_launched(tmp_path):
    """Quiet waiting of the resident kernel (only row 0 polls the host, the other rows wait in L2 for its go: what the
    library switches to while host-to-device copies are in flight), forced for every message with FFLINS_QUIET=1: 40
    evaluations of a 5-reference window, each bit-identical to the launched kernel."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "quiet.py"
    script.write_text(_QUIET_SCRIPT)
    out = subprocess.run([sys.executable, str(script), root], env=dict(os.environ, FFLINS_QUIET="1", FFLINS_COMBINED="0"),
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "QUIET_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_resident_combined_quiet_message_matches_launched(tmp_path):
    """The same with the combined message (FFLINS_COMBINED=1): row 0's warp polls the whole window message and hands it to
    the other rows through L2 - one PCIe read per message instead of two in series."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "quiet.py"
    script.write_text(_QUIET_SCRIPT)
    out = subprocess.run([sys.executable, str(script), root], env=dict(os.environ, FFLINS_QUIET="1", FFLINS_COMBINED="1"),
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "QUIET_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ---- fflins_window_solve: the 5 + chi2 + 10 schedule on the device (one launch per keyframe) ------------------------------
def _solve_setup(c, seq, keys, seed, trans=0.03, rot=0.004):
    import lm_ref
    rng = np.random.default_rng(seed)
    c.associate()
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    refs = [k["id"] for k in keys[:-1]]
    n = len(refs)
    x = np.concatenate([pose_refs.reshape(-1), pose_cur, ext, [td]])
    for k in range(n + 1):   # a solver has something to do: every pose starts off its optimum
        x[7 * k:7 * k + 7] = lm_ref.pose_plus(x[7 * k:7 * k + 7], np.concatenate([rng.normal(size=3) * trans, rng.normal(size=3) * rot]))
    return rng, refs, n, x


def _clear_outliers(c, refs):
    for r in refs:
        c.set_outlier(r, np.zeros(len(c.features(r)["outlier"]), np.uint8))


def _dense_prior(rng, n, x0, weight):
    D = 6 * n + 13
    J0 = rng.normal(size=(D + 5, D)) * np.sqrt(weight)
    e0 = rng.normal(size=D + 5) * 0.1
    return J0.T @ J0, -J0.T @ e0, x0.copy(), 0.5 * e0 @ e0


def _host_lm(c, refs, n, x, flags, prior, fixed_refs=0, **kw):
    """The same solve driven from the host: numpy LM (tests/lm_ref.py) over fflins_window_eval / fflins_window_chi2."""
    import lm_ref

    def ev(xx):
        o = c.window_eval(refs, xx[:7 * n].reshape(n, 7), xx[7 * n:7 * n + 7], xx[7 * n + 7:7 * n + 14], float(xx[7 * n + 14]), flags)
        return o["H"], o["b"], o["cost"]

    def chi2(xx, thr):
        return c.window_chi2(refs, xx[:7 * n].reshape(n, 7), xx[7 * n:7 * n + 7], xx[7 * n + 7:7 * n + 14], float(xx[7 * n + 14]), thr)
    return lm_ref.window_solve(ev, chi2, n, x, flags, prior=prior, fixed_refs=fixed_refs, **kw)


def _device_solve(c, refs, n, x, flags, prior, fixed_refs=0, **kw):
    opt = c.solve_options(flags=flags, fixed_refs=fixed_refs, **kw)
    pr, pc, ex, td, rep = c.window_solve(refs, x[:7 * n].reshape(n, 7), x[7 * n:7 * n + 7], x[7 * n + 7:7 * n + 14], float(x[7 * n + 14]),
                                         options=opt, prior=prior)
    return np.concatenate([pr.reshape(-1), pc, ex, [td]]), rep


def _check_solve(c, seq, keys, seed, flags, with_prior, fixed_refs=0, weight=400.0):
    rng, refs, n, x = _solve_setup(c, seq, keys, seed)
    prior = _dense_prior(rng, n, x, weight) if with_prior else None
    _clear_outliers(c, refs)
    xd, rep = _device_solve(c, refs, n, x, flags, prior, fixed_refs)
    out_d = [c.features(r)["outlier"].copy() for r in refs]
    _clear_outliers(c, refs)
    xh, (s0, s1), counts = _host_lm(c, refs, n, x, flags, prior, fixed_refs)
    out_h = [c.features(r)["outlier"].copy() for r in refs]
    _clear_outliers(c, refs)
    assert rep.status == 0
    assert [rep.iterations[0], rep.successful[0], rep.evaluations[0]] == [s0["iterations"], s0["successful"], s0["evaluations"]]
    assert [rep.iterations[1], rep.successful[1], rep.evaluations[1]] == [s1["iterations"], s1["successful"], s1["evaluations"]]
    assert s0["successful"] >= 1 and rep.cost_final[1] < rep.cost_initial[0]
    assert list(rep.n_outliers[:n]) == list(counts) and all(np.array_equal(a, b) for a, b in zip(out_d, out_h))
    assert abs(rep.cost_initial[0] - s0["cost_initial"]) <= 1e-9 * s0["cost_initial"]
    assert abs(rep.cost_final[1] - s1["cost_final"]) <= 1e-9 * s1["cost_final"]
    assert np.max(np.abs(xd - xh)) <= 1e-9, np.max(np.abs(xd - xh))   # fp64 state: <= 1e-9 absolute (m, quaternion units, s)
    return xd, rep, x


def test_window_solve_full_window_with_prior(window10):
    """The benched shape: 10 references, every pose, the extrinsic and td free, a dense prior (D = 73)."""
    c, seq, keys = window10
    xd, rep, x0 = _check_solve(c, seq, keys, 41, 1, True)
    assert np.max(np.abs(xd - x0)) > 1e-4


def test_window_solve_registration_only(window10):
    """References, extrinsic and td constant, no prior: scan-to-window registration of the newest keyframe."""
    c, seq, keys = window10
    xd, rep, x0 = _check_solve(c, seq, keys, 42, 1 | 2 | 8 | 16, False)
    n = len(keys) - 1
    assert np.array_equal(xd[:7 * n], x0[:7 * n]) and np.array_equal(xd[7 * n + 7:], x0[7 * n + 7:])
    truth = seq.imu_pose7(keys[-1]["fr"]["stamp"])
    assert np.linalg.norm(xd[7 * n:7 * n + 3] - truth[:3]) < np.linalg.norm(x0[7 * n:7 * n + 3] - truth[:3])


def test_window_solve_gauge_fixed_and_ext_const(window10):
    c, seq, keys = window10
    _check_solve(c, seq, keys, 43, 1 | 8 | 16, True, fixed_refs=0b1, weight=25.0)


def test_window_solve_14_references(window15):
    """D = 97: the largest window the context accepts."""
    c, seq, keys = window15
    _check_solve(c, seq, keys, 44, 1, True)


def test_window_solve_is_deterministic_and_coexists_with_the_resident_kernel(window10):
    c, seq, keys = window10
    rng, refs, n, x = _solve_setup(c, seq, keys, 45)
    prior = _dense_prior(rng, n, x, 400.0)
    args = (refs, x[:7 * n].reshape(n, 7), x[7 * n:7 * n + 7], x[7 * n + 7:7 * n + 14], float(x[7 * n + 14]))
    _clear_outliers(c, refs)
    before = c.window_eval(*args, 1)           # (starts the resident evaluation kernel)
    xa, ra = _device_solve(c, refs, n, x, 1, prior)
    _clear_outliers(c, refs)
    xb, rb = _device_solve(c, refs, n, x, 1, prior)
    _clear_outliers(c, refs)
    after = c.window_eval(*args, 1)
    assert bits_equal(xa, xb) and ra.cost_final[1] == rb.cost_final[1] and list(ra.n_outliers) == list(rb.n_outliers)
    for k in ("H", "b", "cost", "n_res"):
        assert bits_equal(before[k], after[k])
    # tolerances off: the schedule runs to its iteration limits (what bench.py times)
    xc, rc = _device_solve(c, refs, n, x, 1, prior, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    _clear_outliers(c, refs)
    assert rc.iterations[0] == 5 and rc.iterations[1] == 10 and rc.evaluations[0] + rc.evaluations[1] <= 17


# ---- mutates the shared window10 fixture: keep it the LAST test of this file that uses it --------------------------------
def test_window10_remove_newest(window10, oracle):
    """removeNewestLidarFrame (lidar_map.cc:175-188): the newest keyframe leaves the window, the previous one becomes
    the current frame of the next association."""
    from fflins import api
    c, seq, keys = window10
    ids = c.keyframe_ids()
    assert ids == [k["id"] for k in keys]
    rid = c.keyframe_remove_newest()
    assert rid == keys[-1]["id"] and c.keyframe_ids() == ids[:-1]
    with pytest.raises(api.FflinsError):
        c.frame_info(rid)
    st, planes = _check_association(c, oracle, keys[:-1])
    assert st.n_refs == 9 and planes > 500
    total, _ = _check_factors(c, oracle, seq, keys[:-1], 1, 81)
    assert total > 500


# ---- fflins_host_prefetch: a raw frame announced ahead of its fflins_frame_process_aos call -------------------------------------
def test_host_prefetch_is_transparent():
    """An announced frame (its copy already on the copy stream) gives bit-identical clouds to a frame first seen by
    fflins_frame_process_aos; announcements that are replaced, never consumed or overtaken by the staging ring are harmless."""
    from fflins import api, synth
    seq = synth.Sequence("robot")
    c = api.Context(api.default_config(point_filter_num=seq.filter_num))
    frames = [seq.frame(i) for i in range(6)]
    recs = []
    for fr in frames:
        rec = np.ascontiguousarray(synth.pack_aos(fr["xyzi"], fr["time"])).copy()
        api.host_register(rec)
        recs.append(rec)

    def process(fid, i, sync):
        fr = frames[i]
        return c.frame_process_aos(fid, recs[i], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"], fr["td"],
                                   sync=sync)
    want = [process(100 + i, i, True) for i in range(6)]
    which_all = (api.CLOUD_LIDAR_FULL, api.CLOUD_LIDAR, api.CLOUD_WORLD, api.CLOUD_MAP_FULL)
    # (1) announce, then process synchronously
    c.host_prefetch(recs[0])
    info = process(200, 0, True)
    assert (info.n_down, info.n_dec) == (want[0].n_down, want[0].n_dec)
    for which in which_all:
        assert bits_equal(c.download(200, which), c.download(100, which))
    # (2) the bench's pattern: the newest frame of an interval announced ahead, the interval queued asynchronously, merged
    c.host_prefetch(recs[4])
    for i in range(5):
        process(300 + i, i, False)
    a = c.keyframe_merge(304, [300, 301, 302, 303])
    b = c.keyframe_merge(104, [100, 101, 102, 103])
    assert a.n_map == b.n_map and a.n_down == want[4].n_down
    for i in range(5):
        for which in which_all[:3]:
            assert bits_equal(c.download(300 + i, which), c.download(100 + i, which)), (i, which)
    assert bits_equal(c.download(304, api.CLOUD_MAP), c.download(104, api.CLOUD_MAP))
    # (2b) the whole next interval announced, newest first; processed oldest first
    for i in (4, 3, 2, 1, 0):
        c.host_prefetch(recs[i])
    for i in range(5):
        process(600 + i, i, False)
    a = c.keyframe_merge(604, [600, 601, 602, 603])
    assert a.n_map == b.n_map
    for i in range(5):
        for which in which_all[:3]:
            assert bits_equal(c.download(600 + i, which), c.download(100 + i, which)), (i, which)
    assert bits_equal(c.download(604, api.CLOUD_MAP), c.download(104, api.CLOUD_MAP))
    # (3) a buffer announced twice, one announced but not the one processed next, and one that is never consumed while the
    # staging ring wraps over it
    c.host_prefetch(recs[1])
    c.host_prefetch(recs[1])
    c.host_prefetch(recs[2])
    info = process(400, 1, True)          # announced (twice): taken from its staging slot
    for which in which_all:
        assert bits_equal(c.download(400, which), c.download(101, which))
    info = process(403, 3, True)          # never announced: copied as usual
    for which in which_all:
        assert bits_equal(c.download(403, which), c.download(103, which))
    for k in range(40):                   # recs[2] stays announced; 40 more frames walk the ring past its slot
        info = process(500 + k, k % 6, True)
        assert info.n_down == want[k % 6].n_down
        c.release(500 + k)
    info = process(401, 2, True)          # the dropped announcement is simply copied again
    for which in which_all:
        assert bits_equal(c.download(401, which), c.download(102, which))
    # (4) a changed point count under the same pointer is not taken for the announced frame
    c.host_prefetch(recs[3])
    fr = frames[3]
    half = len(recs[3]) // 2
    info = c.frame_process_aos(402, recs[3][:half], fr["ins_t"], fr["ins_p"], fr["ins_q"], seq.R_bl, seq.t_bl, fr["stamp"], fr["td"])
    assert info.n_raw == half
    c.synchronize()
    for rec in recs:
        api.host_unregister(rec)
    c.close()
