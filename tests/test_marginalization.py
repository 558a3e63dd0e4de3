"""Host algebra of the marginalization step (CPU): Schur complement and linearised prior against first principles."""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
from fflins import marginalization as M   # noqa: E402


def _random_problem(rng, rows=200, dim=19):
    J = rng.normal(size=(rows, dim))
    r = rng.normal(size=rows)
    return J, r, J.T @ J, -J.T @ r


def test_schur_equals_minimising_over_the_dropped_block():
    """The prior cost equals the full quadratic minimised over the marginalised states, for every dx of the rest."""
    rng = np.random.default_rng(1)
    J, r, H, b = _random_problem(rng)
    J0, e0 = M.marginalize_oldest(H, b)
    assert J0.shape == (13, 13) and e0.shape == (13,)
    for _ in range(5):
        dx = rng.normal(size=13) * 0.1
        # minimise |r + J [dm; dx]|^2 over dm
        Jm, Jr = J[:, :6], J[:, 6:]
        dm = np.linalg.lstsq(Jm, -(r + Jr @ dx), rcond=None)[0]
        full = np.sum((r + Jm @ dm + Jr @ dx) ** 2)
        const = np.sum((r + Jm @ np.linalg.lstsq(Jm, -(r + Jr @ np.zeros(13)), rcond=None)[0]) ** 2) - np.sum(M.prior_residual(J0, e0, np.zeros(13)) ** 2)
        prior = np.sum(M.prior_residual(J0, e0, dx) ** 2) + const
        assert abs(full - prior) <= 1e-9 * max(1.0, full)


def test_linearize_reproduces_normal_equations():
    rng = np.random.default_rng(2)
    _, _, H, b = _random_problem(rng)
    Hp, bp = M.schur(H, b, 6)
    J0, e0 = M.linearize(Hp, bp)
    assert np.allclose(J0.T @ J0, Hp, rtol=1e-10, atol=1e-10)
    assert np.allclose(J0.T @ e0, -bp, rtol=1e-9, atol=1e-9)      # gradient of the prior = -bp (sign of :106)


def test_rank_deficient_block_is_dropped_not_inverted():
    """Eigenvalues <= EPS are zeroed (marginalization_info.cc:96-99,120-126): unobservable directions carry no prior."""
    rng = np.random.default_rng(3)
    J = rng.normal(size=(100, 19))
    J[:, 2] = 0.0       # ref z unobservable
    J[:, 18] = 0.0      # td unobservable
    r = rng.normal(size=100)
    J0, e0 = M.marginalize_oldest(J.T @ J, -J.T @ r)
    assert np.all(np.isfinite(J0)) and np.all(np.isfinite(e0))
    assert np.allclose(J0[:, 12], 0.0, atol=1e-9)   # no information on td
    assert np.linalg.matrix_rank(J0, tol=1e-6) == 12


def test_assemble_places_blocks():
    rng = np.random.default_rng(4)
    Hs = [np.eye(19) * (i + 1) for i in range(3)]
    bs = [np.full(19, float(i + 1)) for i in range(3)]
    H0, b0 = M.assemble(Hs, bs, [0, 1, 2], 3)
    assert H0.shape == (31, 31)
    assert np.allclose(np.diag(H0)[:18], np.repeat([1.0, 2.0, 3.0], 6))
    assert np.allclose(np.diag(H0)[18:], 6.0) and np.allclose(b0[18:], 6.0)
    assert np.allclose(b0[:18], np.repeat([1.0, 2.0, 3.0], 6))
