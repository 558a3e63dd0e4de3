"""Independent numpy statement of the window solve that fflins_window_solve runs on the device: the
Levenberg-Marquardt loop of ff-lins_b200/host/host.cc (lm_solve: Ceres' documented trust-region
algorithm) over per-pair 19 x 19 LiDAR blocks and a quadratic prior, with the 5 + chi2 + 10 schedule
of Optimizer::optimization (optimizer.cc:85-185). Test infrastructure only.

State x: n reference poses, the current pose, the extrinsic (7 doubles: t, q = x y z w), td.
Tangent: [ref 0 .. ref n-1 | cur | ext | td], D = 6 n + 13."""
import numpy as np


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def pose_plus(x, d):
    out = np.array(x, dtype=np.float64)
    out[:3] = x[:3] + d[:3]
    a = np.linalg.norm(d[3:6])
    dq = np.array([0, 0, 0, 1.0]) if a == 0 else np.concatenate([d[3:6] * (np.sin(a / 2) / a), [np.cos(a / 2)]])
    q = qmul(x[3:7], dq)
    out[3:7] = q / np.linalg.norm(q)
    return out


def pose_minus(x, x0):
    q0 = x0[3:7]
    qi = np.array([-q0[0], -q0[1], -q0[2], q0[3]]) / np.dot(q0, q0)
    dq = qmul(qi, x[3:7])
    s = -2.0 if dq[3] < 0 else 2.0
    return np.concatenate([x[:3] - x0[:3], s * dq[:3]])


def cols(n, p):
    return np.concatenate([np.arange(6 * p, 6 * p + 6), np.arange(6 * n, 6 * n + 13)])


def fixed_mask(n, flags, fixed_refs=0):
    f = np.zeros(6 * n + 13, bool)
    for p in range(n):
        f[6 * p:6 * p + 6] = bool(flags & 2) or bool((fixed_refs >> p) & 1)
    f[6 * n:6 * n + 6] = bool(flags & 4)
    f[6 * n + 6:6 * n + 12] = bool(flags & 8)
    f[6 * n + 12] = bool(flags & 16)
    return f


def delta(n, x, x0):
    d = np.zeros(6 * n + 13)
    for k in range(n + 2):
        d[6 * k:6 * k + 6] = pose_minus(x[7 * k:7 * k + 7], x0[7 * k:7 * k + 7])
    d[6 * n + 12] = x[7 * n + 14] - x0[7 * n + 14]
    return d


def linearize(eval_fn, n, x, prior):
    D = 6 * n + 13
    H = np.zeros((D, D))
    g = np.zeros(D)
    cost = 0.0
    if prior is not None:
        H0, b0, x0, c0 = prior
        d = delta(n, x, x0)
        H += H0
        g += b0 - H0 @ d
        cost += c0 + 0.5 * d @ H0 @ d - b0 @ d
    Hb, bb, cb = eval_fn(x)
    for p in range(n):
        c = cols(n, p)
        H[np.ix_(c, c)] += Hb[p]
        g[c] += bb[p]
        cost += cb[p, 0]
    return H, g, cost


def apply_step(n, x, dx, fixed):
    out = np.array(x, dtype=np.float64)
    for k in range(n + 2):
        if not fixed[6 * k]:
            out[7 * k:7 * k + 7] = pose_plus(x[7 * k:7 * k + 7], dx[6 * k:6 * k + 6])
    out[7 * n + 14] = x[7 * n + 14] + dx[6 * n + 12]
    return out


def lm_solve(eval_fn, n, x, fixed, max_iters, prior=None, ftol=1e-6, gtol=1e-10, ptol=1e-8, radius=1e4):
    H, g, cost = linearize(eval_fn, n, x, prior)
    st = dict(iterations=0, successful=0, evaluations=1, cost_initial=cost)
    decrease = 2.0
    free = ~fixed
    for _ in range(max_iters):
        st["iterations"] += 1
        if np.max(np.abs(g[free]), initial=0.0) < gtol:
            break
        A = H.copy()
        rhs = g.copy()
        A[fixed, :] = 0
        A[:, fixed] = 0
        A[fixed, fixed] = 1
        rhs[fixed] = 0
        idx = np.where(free)[0]
        A[idx, idx] += np.clip(H[idx, idx], 1e-6, 1e32) / radius
        ok = True
        try:
            L = np.linalg.cholesky(A)
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
        except np.linalg.LinAlgError:
            ok = False
        model = 0.0
        if ok:
            Hf = H[np.ix_(idx, idx)]
            model = dx[idx] @ (g[idx] - 0.5 * Hf @ dx[idx])
        if not ok or not model > 0:
            radius /= decrease
            decrease *= 2
            continue
        xn = apply_step(n, x, dx, fixed)
        Hn, gn, cost_n = linearize(eval_fn, n, xn, prior)
        st["evaluations"] += 1
        rho = (cost - cost_n) / model
        if rho > 1e-3:
            st["successful"] += 1
            dcost = cost - cost_n
            x, H, g, cost = xn, Hn, gn, cost_n
            radius = radius / max(1.0 / 3.0, 1.0 - (2 * rho - 1) ** 3)
            decrease = 2.0
            if abs(dcost) <= ftol * cost:
                break
            if np.linalg.norm(dx) <= ptol * (np.linalg.norm(x[:7 * (n + 1)]) + ptol):
                break
        else:
            radius /= decrease
            decrease *= 2
    st["cost_final"] = cost
    return x, st


def window_solve(eval_fn, chi2_fn, n, x, flags, max_iterations=(5, 10), chi2_threshold=3.841, prior=None, fixed_refs=0,
                 ftol=1e-6, gtol=1e-10, ptol=1e-8, radius=1e4):
    """eval_fn(x) -> (H [n,19,19], b [n,19], cost [n,2]); chi2_fn(x, threshold) -> per-pair outlier counts (and
    marks the outliers so that later evaluations skip them)."""
    fixed = fixed_mask(n, flags, fixed_refs)
    x, s0 = lm_solve(eval_fn, n, np.array(x, dtype=np.float64), fixed, max_iterations[0], prior, ftol, gtol, ptol, radius)
    outliers = chi2_fn(x, chi2_threshold) if chi2_threshold > 0 else np.zeros(n, np.int32)
    x, s1 = lm_solve(eval_fn, n, x, fixed, max_iterations[1], prior, ftol, gtol, ptol, radius)
    return x, (s0, s1), outliers
