"""Marginalization of the oldest keyframe from the LiDAR normal equations ("next" row 2 of SURVEY.md §8f).

The reference builds H0/b0 by evaluating every factor connected to the state being dropped
(MarginalizationInfo::constructEquation, ff_lins/factors/marginalization_info.cc:133-210), eliminates it with a
Schur complement (schurElimination, :109-131) and turns the result into a linearised prior
(linearization, :93-107) that MarginalizationFactor::Evaluate (marginalization_factor.cc:37-91) applies as
e = e0 + J0 dx. The LiDAR share of H0/b0 is exactly what the device path already produces: the robustified 19 x 19
block of the (oldest, current) pair from fflins_window_eval — H = sum J^T J over [ref pose 6 | cur pose 6 | ext 6 |
td 1] in the right-perturbation tangent space (pose_manifold.cc:35-59), b = -sum J^T r (the sign of :208). What is
left is dense 19-dimensional host algebra, restated here with numpy; IMU preintegration factors and an earlier
prior would simply be added into H0/b0 before `schur` (they are out of scope).

Evaluation-agnostic like window_solver.py: the blocks may come from the CUDA path or from the CPU oracle, which is
how tests/test_gpu_marginalization.py compares the two."""
import numpy as np

EPS = 1e-8   # marginalization_info.h:141


def assemble(blocks_H, blocks_b, ref_slots, n_ref_states):
    """constructEquation for LiDAR pair blocks: state order [ref_0 .. ref_{n-1} | cur | ext | td], 6 per pose.
    blocks_H[i] (19 x 19) / blocks_b[i] (19) belong to the pair whose reference state is ref_slots[i]."""
    dim = 6 * n_ref_states + 13
    H0, b0 = np.zeros((dim, dim)), np.zeros(dim)
    tail = np.arange(6 * n_ref_states, dim)
    for H, b, slot in zip(blocks_H, blocks_b, ref_slots):
        idx = np.concatenate([np.arange(6 * slot, 6 * slot + 6), tail])
        H0[np.ix_(idx, idx)] += H
        b0[idx] += b
    return H0, b0


def _pinv_sym(A):
    """Hmm^-1 as in schurElimination: eigenvalues <= EPS are dropped (:120-126)."""
    w, V = np.linalg.eigh(A)
    winv = np.where(w > EPS, 1.0 / np.where(w > EPS, w, 1.0), 0.0)
    return (V * winv) @ V.T


def schur(H0, b0, m):
    """Eliminate the first m states: Hp = Hrr - Hrm Hmm^-1 Hmr, bp = br - Hrm Hmm^-1 bm (:109-131)."""
    Hmm = 0.5 * (H0[:m, :m] + H0[:m, :m].T)
    Hmr, Hrm, Hrr = H0[:m, m:], H0[m:, :m], H0[m:, m:]
    Hmm_inv = _pinv_sym(Hmm)
    return Hrr - Hrm @ Hmm_inv @ Hmr, b0[m:] - Hrm @ Hmm_inv @ b0[:m]


def linearize(Hp, bp):
    """J0 = S^1/2 V^T, e0 = -S^-1/2 V^T bp with eigenvalues <= EPS zeroed (:93-107)."""
    w, V = np.linalg.eigh(0.5 * (Hp + Hp.T))
    keep = w > EPS
    S = np.where(keep, w, 0.0)
    S_inv = np.where(keep, 1.0 / np.where(keep, w, 1.0), 0.0)
    J0 = np.sqrt(S)[:, None] * V.T
    e0 = np.sqrt(S_inv) * (V.T @ -bp)
    return J0, e0


def marginalize_oldest(H_pair, b_pair):
    """Drop the reference pose of ONE (oldest, current) pair block: prior (J0, e0) on [cur 6 | ext 6 | td 1]."""
    H0, b0 = assemble([H_pair], [b_pair], [0], 1)
    return linearize(*schur(H0, b0, 6))


def prior_residual(J0, e0, dx):
    """MarginalizationFactor::Evaluate: e = e0 + J0 dx (marginalization_factor.cc:72-74)."""
    return e0 + J0 @ dx
