"""Pins the quasiseparable half of the oracle with the reference's own test relations (CPU only)."""

import numpy as np
import pytest
import scipy.linalg

from oracle import tinygp_np as o

QS_KERNELS = [
    lambda: o.qs.Matern32(1.5),                                  # test_solver.py:30
    lambda: 1.5 * o.qs.Matern52(1.5) + 0.3 * o.qs.Exp(1.5),      # test_solver.py:31
    lambda: o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9),   # BASELINE config 4
    lambda: o.qs.SHO(1.5, 0.3),
    lambda: o.qs.SHO(1.5, 0.5),
    lambda: o.qs.Celerite(1.1, 0.8, 0.9, 0.1),
    lambda: o.qs.Cosine(2.5) + o.qs.Exp(0.7, 1.3),
]


@pytest.fixture
def data():
    rng = np.random.default_rng(84930)
    X = np.sort(rng.uniform(-3, 3, 50))
    return X, np.sin(X)


@pytest.mark.parametrize("mk", QS_KERNELS)
def test_transition_is_expm(mk):
    # tests/test_kernels/test_quasisep.py:66-72
    k = mk()
    F = k.design_matrix()
    for dt in (0.0, 0.13, 1.7):
        np.testing.assert_allclose(k.transition_matrix(0.4, 0.4 + dt), scipy.linalg.expm(F.T * dt),
                                   rtol=5e-7, atol=5e-7)


@pytest.mark.parametrize("mk", QS_KERNELS)
def test_generators_match_dense_kernel(mk, data):
    # tests/test_kernels/test_quasisep.py:53-64 : to_symm_qsm(x).to_dense() == kernel(x, x)
    X, _ = data
    k = mk()
    s = o.QuasisepSolver(k, X, o.Diagonal(np.zeros(len(X)) + 0.1))
    np.testing.assert_allclose(s.covariance(), k(X, X) + 0.1 * np.eye(len(X)), rtol=5e-7, atol=5e-7)


def test_celerite_closed_form(data):
    # tests/test_kernels/test_quasisep.py:83-97
    X, _ = data
    a, b, c, d = 1.1, 0.8, 0.9, 0.1
    tau = np.abs(X[:, None] - X[None, :])
    expect = np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
    np.testing.assert_allclose(o.qs.Celerite(a, b, c, d)(X, X), expect, rtol=5e-7, atol=5e-7)


@pytest.mark.parametrize("mk", QS_KERNELS)
def test_quasisep_equals_direct_equals_kalman(mk, data):
    # tests/test_solvers/test_quasisep/test_solver.py:27-103 and tests/test_solvers/test_kalman.py:47-69
    X, y = data
    k = mk()
    noise = o.Diagonal(np.full(len(X), 0.1))
    qsolver = o.QuasisepSolver(k, X, noise)
    K = k(X, X) + 0.1 * np.eye(len(X))
    L = np.linalg.cholesky(K)
    np.testing.assert_allclose(qsolver.normalization(),
                               np.sum(np.log(np.diag(L))) + 0.5 * len(X) * np.log(2 * np.pi), rtol=1e-10)
    np.testing.assert_allclose(qsolver.solve_triangular(y), scipy.linalg.solve_triangular(L, y, lower=True),
                               rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(qsolver.solve_triangular(y, transpose=True),
                               scipy.linalg.solve_triangular(L, y, lower=True, trans=1), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(qsolver.dot_triangular(y), L @ y, rtol=1e-7, atol=1e-9)
    gp = o.GaussianProcess(k, X, diag=0.1)
    lp = gp.log_probability(y)
    lp_dense = -0.5 * y @ np.linalg.solve(K, y) - np.sum(np.log(np.diag(L))) - 0.5 * len(X) * np.log(2 * np.pi)
    assert abs(lp - lp_dense) <= 1e-10 * abs(lp_dense)
    assert abs(o.KalmanLogp(k, X, y, np.full(len(X), 0.1)) - lp) <= 1e-9 * abs(lp)


def test_cholesky_factor_is_dense_cholesky(data):
    # tests/test_solvers/test_quasisep/test_core.py:308-323
    X, _ = data
    k = o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9)
    s = o.QuasisepSolver(k, X, o.Diagonal(np.full(len(X), 0.1)))
    N = len(X)
    Lq = np.diag(s.c) + o.qs_lower_matmul(s.p, s.w, s.a, np.eye(N))
    np.testing.assert_allclose(Lq, np.linalg.cholesky(k(X, X) + 0.1 * np.eye(N)), rtol=5e-7, atol=5e-7)


def test_unsorted_raises(data):
    # tests/test_solvers/test_quasisep/test_solver.py:127-143
    X, _ = data
    with pytest.raises(ValueError, match="sorted"):
        o.QuasisepSolver(o.qs.Matern32(1.0), X[::-1], o.Diagonal(np.full(len(X), 0.1)))
    # ties are allowed (diff == 0 is not < 0)
    Xt = np.array([0.0, 0.0, 1.0])
    o.QuasisepSolver(o.qs.Matern32(1.0), Xt, o.Diagonal(np.full(3, 0.1)))


def test_fast_generators_equal_loop(data):
    X, _ = data
    k = o.qs.SHO(1.5, 3.0, 1.8) + 0.7 * o.qs.Matern32(1.5, 0.9)
    for a, b in zip(k.to_symm_qsm(X), o.qs_generators_fast(k, X)):
        np.testing.assert_array_equal(a, b)


def test_c_restatement_equals_numpy(data):
    from oracle import cref
    X, y = data
    k = o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9)
    d, p, q, a = k.to_symm_qsm(X)
    d = d + 0.1
    c, w, bad = cref.qs_cholesky(d, p, q, a)
    c2, w2 = o.qs_cholesky(d, p, q, a)
    assert bad == 0
    np.testing.assert_allclose(c, c2, rtol=1e-14)
    np.testing.assert_allclose(w, w2, rtol=1e-12, atol=1e-15)
    lp = cref.qs_log_probability(d, p, q, a, y)
    lp2 = o.GaussianProcess(k, X, diag=0.1).log_probability(y)
    assert abs(lp - lp2) <= 1e-12 * abs(lp2)
    assert cref.qs_log_probability(d - 10.0, p, q, a, y) == -np.inf
