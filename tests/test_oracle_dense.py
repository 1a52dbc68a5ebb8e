"""Pins the dense half of the oracle with the reference's own test relations (CPU only)."""

import numpy as np
import pytest
import scipy.stats

from oracle import tinygp_np as o


@pytest.fixture
def data():
    rng = np.random.default_rng(84930)  # tests/test_solvers/test_quasisep/test_solver.py:16
    X = np.sort(rng.uniform(-3, 3, 50))
    return X, np.sin(X)


KERNELS = [
    lambda: o.ExpSquared(1.5),
    lambda: o.Exp(0.7),
    lambda: o.Matern32(1.5),
    lambda: o.Matern52(2.5),
    lambda: o.Cosine(2.3),
    lambda: o.ExpSineSquared(2.3, gamma=1.3),
    lambda: o.RationalQuadratic(1.2, alpha=1.7),
    lambda: 1.8 * o.ExpSquared(0.9) + o.Matern32(3.0) * o.Constant(0.4),
]


@pytest.mark.parametrize("mk", KERNELS)
def test_logp_is_mvn_logpdf(mk, data):
    X, y = data
    gp = o.GaussianProcess(mk(), X, diag=0.1)
    ref = scipy.stats.multivariate_normal(np.zeros(len(X)), gp.covariance).logpdf(y)
    assert abs(gp.log_probability(y) - ref) <= 1e-10 * abs(ref)


def test_scalar_definitions():
    # kernels/stationary.py formulas on one pair, written out by hand
    x1, x2 = np.array([[0.3, -1.2]]), np.array([[1.1, 0.4]])
    l1 = abs(0.3 - 1.1) + abs(-1.2 - 0.4)
    l2sq = (0.3 - 1.1) ** 2 + (-1.2 - 0.4) ** 2
    assert np.isclose(o.ExpSquared(1.3)(x1, x2)[0, 0], np.exp(-0.5 * l2sq / 1.3**2), rtol=1e-15)
    a = np.sqrt(3) * l1 / 0.8
    assert np.isclose(o.Matern32(0.8)(x1, x2)[0, 0], (1 + a) * np.exp(-a), rtol=1e-15)
    a = np.sqrt(5) * l1 / 0.8
    assert np.isclose(o.Matern52(0.8)(x1, x2)[0, 0], (1 + a + a * a / 3) * np.exp(-a), rtol=1e-15)
    # RationalQuadratic's default metric is L1 (stationary.py:56,232-235)
    r2 = l1**2 / 1.1**2
    assert np.isclose(o.RationalQuadratic(1.1, alpha=2.0)(x1, x2)[0, 0], (1 + 0.5 * r2 / 2.0) ** -2.0, rtol=1e-15)
    r2 = l2sq / 1.1**2
    assert np.isclose(o.RationalQuadratic(1.1, o.L2Distance(), alpha=2.0)(x1, x2)[0, 0],
                      (1 + 0.5 * r2 / 2.0) ** -2.0, rtol=1e-15)
    # L2 distance is exactly zero on the diagonal
    assert o.Matern32(0.8, o.L2Distance())(x1, x1)[0, 0] == 1.0


def test_solve_triangular_vs_solve(data):
    # tests/test_kernels/test_kernels.py:72-83
    X, y = data
    k = o.ExpSquared(1.5)
    s = o.DirectSolver(k, X, o.Diagonal(np.full(len(X), 0.1)))
    K = s.covariance()
    z = s.solve_triangular(s.solve_triangular(y), transpose=True)
    np.testing.assert_allclose(z, np.linalg.solve(K, y), rtol=1e-9)
    np.testing.assert_allclose(s.dot_triangular(s.solve_triangular(y)), y, rtol=1e-10, atol=1e-12)


def test_condition_relations(data):
    X, y = data
    gp = o.GaussianProcess(o.Matern52(1.5), X, diag=0.1)
    Xt = np.linspace(-2.5, 2.5, 17)
    lp, cond = gp.condition(y, Xt)
    K = gp.covariance
    Ks = o.Matern52(1.5)(X, Xt)
    mu = Ks.T @ np.linalg.solve(K, y)
    cov = o.Matern52(1.5)(Xt, Xt) + np.sqrt(np.finfo(float).eps) * np.eye(17) - Ks.T @ np.linalg.solve(K, Ks)
    np.testing.assert_allclose(cond.loc, mu, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(cond.covariance, cov, rtol=1e-7, atol=1e-9)
    assert lp == gp.log_probability(y)
    # predicting at the inputs uses the noise shortcut (gp.py:342-346)
    mu0 = gp.predict(y)
    np.testing.assert_allclose(mu0, (K - 0.1 * np.eye(len(X))) @ np.linalg.solve(K, y), rtol=1e-8, atol=1e-10)


def test_non_pd_gives_minus_inf():
    X = np.zeros(5)
    gp = o.GaussianProcess(o.ExpSquared(1.0), X, diag=-2.0)
    assert gp.log_probability(np.ones(5)) == -np.inf
