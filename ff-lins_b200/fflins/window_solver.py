"""Dense Levenberg-Marquardt over the LiDAR factors of the sliding window ("next" row 1 of SURVEY.md
§8f, reduced): the newest keyframe's IMU pose (and optionally the LiDAR-IMU extrinsic and the time
delay) is solved against the older keyframes, which are held fixed — the schedule of
Optimizer::optimization (ff_lins/ff_lins/optimizer.cc:85-185): 5 iterations, chi-square cull
(3.841), 10 iterations, Huber(1.0) on every residual. IMU preintegration factors and the
marginalization prior are out of scope here, so this is frame-to-keyframe-window registration, not
the full estimator; it exists so that the GPU path can be run closed-loop (solve -> setPose ->
re-project -> next association) and its trajectory compared with the same loop driven by the CPU
oracle (tests/test_gpu_closed_loop.py).

The solver is evaluation-agnostic: `evaluate(pose_cur, ext, td)` must return per-pair blocks
(H [n,19,19], b [n,19], cost [n,2]) in the layout of fflins_window_eval, `cull(pose_cur, ext, td)`
applies the chi-square test. Everything here is host-side 6..13-dimensional linear algebra."""
import numpy as np


def quat_mul(a, b):
    """Hamilton product, quaternions as x,y,z,w."""
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def pose_plus(pose7, delta6):
    """PoseManifold::Plus (ff_lins/factors/pose_manifold.cc:35-50): t + dt, q (x) exp(dtheta), normalised."""
    out = np.array(pose7, np.float64)
    out[:3] += delta6[:3]
    th = np.asarray(delta6[3:6], np.float64)
    a = np.linalg.norm(th)
    dq = np.concatenate([np.sin(a / 2) * th / a, [np.cos(a / 2)]]) if a > 0 else np.array([0.0, 0.0, 0.0, 1.0])
    q = quat_mul(out[3:], dq)
    out[3:] = q / np.linalg.norm(q)
    return out


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def lidar_pose(pose7, ext7):
    """MISC::stateToPoseWithExtrinsic (misc.cc:99-105): lidar pose from the IMU pose block and the extrinsic."""
    R = quat_to_R(pose7[3:])
    return R @ quat_to_R(ext7[3:]), pose7[:3] + R @ ext7[:3]


class WindowSolver:
    def __init__(self, evaluate, cull, n_iters=(5, 10), solve_ext=False, solve_td=False, prior_sigma=None,
                 lm_lambda=1e-4):
        self.evaluate = evaluate
        self.cull = cull
        self.n_iters = n_iters
        self.solve_ext = solve_ext
        self.solve_td = solve_td
        self.prior_sigma = prior_sigma        # optional Gaussian prior on the free block (dim-sized vector)
        self.lm_lambda = lm_lambda
        idx = list(range(6, 12))              # cur pose block of the 19-vector
        if solve_ext:
            idx += list(range(12, 18))
        if solve_td:
            idx += [18]
        self.idx = np.array(idx)

    def _normal_equations(self, pose, ext, td, x0):
        H, b, cost = self.evaluate(pose, ext, td)
        Hs = H.sum(0)[np.ix_(self.idx, self.idx)]
        bs = b.sum(0)[self.idx]
        c = float(cost[:, 0].sum())
        if self.prior_sigma is not None:      # prior pulling the free block to its initial value
            w = 1.0 / np.asarray(self.prior_sigma) ** 2
            dx = self._delta_from(x0, (pose, ext, td))
            Hs = Hs + np.diag(w)
            bs = bs - w * dx
            c += 0.5 * float(np.sum(w * dx * dx))
        return Hs, bs, c

    def _delta_from(self, x0, x):
        """Tangent-space difference x (-) x0 of the free block (small-angle for rotations)."""
        (p0, e0, t0), (p, e, t) = x0, x

        def dpose(a, b):
            dq = quat_mul(np.array([-a[3], -a[4], -a[5], a[6]]), b[3:])
            return np.concatenate([b[:3] - a[:3], 2 * dq[:3] * np.sign(dq[3])])
        d = [dpose(p0, p)]
        if self.solve_ext:
            d.append(dpose(e0, e))
        if self.solve_td:
            d.append([t - t0])
        return np.concatenate(d)

    def _apply(self, pose, ext, td, delta):
        pose = pose_plus(pose, delta[:6])
        k = 6
        if self.solve_ext:
            ext = pose_plus(ext, delta[k:k + 6])
            k += 6
        if self.solve_td:
            td = td + float(delta[k])
        return pose, ext, td

    def _run(self, pose, ext, td, x0, iters):
        lam = self.lm_lambda
        H, b, c = self._normal_equations(pose, ext, td, x0)
        for _ in range(iters):
            D = np.diag(np.maximum(np.diag(H), 1e-12))
            try:
                delta = np.linalg.solve(H + lam * D, b)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            cand = self._apply(pose, ext, td, delta)
            Hn, bn, cn = self._normal_equations(*cand, x0)
            if cn < c:                        # accept, relax the damping
                pose, ext, td = cand
                H, b, c = Hn, bn, cn
                lam = max(lam / 3, 1e-9)
                if np.linalg.norm(delta) < 1e-10:
                    break
            else:
                lam *= 4
        return pose, ext, td, c

    def solve(self, pose_cur, ext, td):
        pose, ext, td = np.array(pose_cur, np.float64), np.array(ext, np.float64), float(td)
        x0 = (pose.copy(), ext.copy(), td)
        pose, ext, td, c0 = self._run(pose, ext, td, x0, self.n_iters[0])
        n_out = self.cull(pose, ext, td)
        pose, ext, td, c1 = self._run(pose, ext, td, x0, self.n_iters[1])
        return dict(pose=pose, ext=ext, td=td, cost=c1, outliers=int(np.sum(n_out)))
