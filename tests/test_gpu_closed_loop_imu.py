"""Closed loop FROM RAW IMU SAMPLES: synthetic IMU increments -> INS
mechanization -> INS window -> undistortion -> keyframe time nodes with IMU preintegration -> window solve (IMU factors +
LiDAR blocks + marginalization prior) -> poses and INS window corrected -> next frame. The identical host loop
(fflins.fusion.FusionLoop over libfflins_host.so) runs once with every LiDAR operation on the GPU library — the solver
calling fflins_window_eval / fflins_window_chi2 through their C addresses — and once on the CPU oracle.
BASELINE.json: the two trajectories must differ by < 1 mm ATE."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K, N_FRAMES, WINDOW = 3, 36, 5      # 12 keyframes: the window fills and 6 marginalizations happen


def _run(seq, frames, backend):
    from fflins.fusion import FusionLoop
    loop = FusionLoop(seq, backend, frames_per_key=K, window=WINDOW, imu_rng=np.random.default_rng(42), arw=2e-4, vrw=2e-3,
                      init_error=(0.0, 0.08))
    for i, fr in enumerate(frames):
        loop.step(i, fr)
    backend.close()
    return np.array(loop.traj), loop.stats


def test_closed_loop_from_raw_imu(oracle):
    from fflins import synth
    from fflins.fusion import GpuLidar
    from oracle.cpu_lidar import CpuLidar
    seq = synth.Sequence("robot")
    frames = [seq.frame(i) for i in range(N_FRAMES)]
    g_traj, g_stats = _run(seq, frames, GpuLidar(seq, window=WINDOW))
    c_traj, c_stats = _run(seq, frames, CpuLidar(seq, window=WINDOW))
    assert g_traj.shape == c_traj.shape == (N_FRAMES // K, 3)
    ate = float(np.sqrt(np.mean(np.sum((g_traj - c_traj) ** 2, axis=1))))
    truth = np.stack([seq.body_state(frames[i]["stamp"] - synth.GPS_T0)[0] for i in range(K - 1, N_FRAMES, K)])
    # frame-to-frame LiDAR factors observe RELATIVE motion: the offset accumulated before the first keyframe stays, the
    # drift stops. The INS starts with a 8 cm/s velocity error: dead reckoning would add 0.08 m/s x 3.3 s = 26 cm.
    rel = np.linalg.norm((g_traj[-1] - g_traj[0]) - (truth[-1] - truth[0]))
    span = frames[N_FRAMES - 1]["stamp"] - frames[K - 1]["stamp"]
    print(f"closed loop from raw IMU: ATE(GPU vs CPU oracle) = {ate * 1e3:.6f} mm over {len(g_traj)} keyframes; relative-motion "
          f"error over {span:.1f} s = {rel * 1e3:.1f} mm (dead reckoning with the 8 cm/s initial velocity error: "
          f"{0.08 * span * 1e3:.0f} mm); features/outliers GPU {g_stats[1:4]} CPU {c_stats[1:4]}")
    assert ate < 1e-3, "trajectory must agree within 1 mm"
    assert rel < 0.02, "IMU + LiDAR factors must stop the drift of the velocity error"
    assert all(s[0] > 200 for s in g_stats[1:])
