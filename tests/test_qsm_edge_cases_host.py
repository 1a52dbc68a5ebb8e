"""Edge cases of the device QSM algebra on the CPU (host build of qsm.cu through the real Python classes): one and two
points, order 1, an order larger than a warp can hold in one pass (24), chunk lengths longer than the series, associativity
with dense vectors, and the consistency-checked Riccati scan's sequential redo."""

import numpy as np
import pytest

from oracle import qsm_np as oq
from qsmhost import HostBackend
from qsmutil import qsmcases
from tinygp_b200.solvers.quasisep import core


@pytest.fixture(scope="module")
def backend():
    return HostBackend()


@pytest.fixture(autouse=True)
def _use_host_build(backend, monkeypatch):
    monkeypatch.setattr(core, "_backend", lambda: backend)
    backend.set_option("qsm_chunk", 0)
    yield
    backend.set_option("qsm_chunk", 0)


def _square(n, ml, mu, seed):
    p1, q1, a1 = qsmcases.generators(seed, ml, n=n)
    p2, q2, a2 = qsmcases.generators(seed + 1, mu, n=n)
    d = qsmcases.diag(seed + 2, n=n) + 3.0
    return (core.SquareQSM(diag=core.DiagQSM(d), lower=core.StrictLowerTriQSM(p1, q1, a1), upper=core.StrictUpperTriQSM(p2, q2, a2)),
            oq.QSM(d, (p1, q1, a1), (p2, q2, a2)))


@pytest.mark.parametrize("n", [1, 2, 3, 9])
@pytest.mark.parametrize("ml,mu", [(1, 1), (1, 4), (3, 2)])
def test_tiny_series_and_order_one(n, ml, mu):
    A, Ao = _square(n, ml, mu, seed=10 * n + ml)
    B, Bo = _square(n, mu, ml, seed=77 + n)
    x = np.random.default_rng(n).normal(size=(n, 2))
    np.testing.assert_allclose(A @ x, Ao.matmul(x), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose((A @ B).to_dense(), Ao.to_dense() @ Bo.to_dense(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose((A + B).to_dense(), Ao.to_dense() + Bo.to_dense(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(A.gram().to_dense(), Ao.to_dense().T @ Ao.to_dense(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(A.inv().to_dense(), np.linalg.inv(Ao.to_dense()), rtol=1e-8, atol=1e-8)
    L = core.LowerTriQSM(diag=A.diag, lower=A.lower)
    np.testing.assert_allclose(L.solve(x), np.linalg.solve(np.tril(Ao.to_dense()), x), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("chunk", [0, 4, 1000])
def test_order_24_generators(backend, chunk):
    """orders beyond one element per lane: every small product loops over the lanes"""
    backend.set_option("qsm_chunk", chunk)
    n, m = 60, 24
    p, q, a = qsmcases.generators(5, m, n=n, contract=0.5)
    d = qsmcases.diag(6, n=n) + 40.0
    L, Lo = core.LowerTriQSM(diag=core.DiagQSM(d), lower=core.StrictLowerTriQSM(p, q, a)), oq.QSM(d, (p, q, a))
    x = np.random.default_rng(1).normal(size=(n, 3))
    np.testing.assert_allclose(L @ x, Lo.matmul(x), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(L.solve(x), Lo.solve(x), rtol=1e-10, atol=1e-10)
    Li = L.inv()
    Sq = Li.T @ Li
    G = core.SymmQSM(diag=Sq.diag, lower=Sq.lower)
    Go = Lo.inv().gram()
    np.testing.assert_allclose(G.to_dense(), Go.to_dense(), rtol=1e-9, atol=1e-12)
    ch = G.cholesky()
    assert ch.info == 0
    np.testing.assert_allclose(ch.to_dense(), np.linalg.cholesky(Go.to_dense()), rtol=1e-7, atol=1e-10)


def test_vector_products_associate():
    A, Ao = _square(40, 3, 2, seed=3)
    B, Bo = _square(40, 2, 4, seed=9)
    x = np.random.default_rng(4).normal(size=40)
    np.testing.assert_allclose((A @ B) @ x, A @ (B @ x), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(x @ A, x @ Ao.to_dense(), rtol=1e-11, atol=1e-11)


def test_consistency_check_redoes_an_ill_conditioned_scan_sequentially(backend):
    """the order-4J conditioned covariance of the CARMA(2,1) golden case carries two large, almost cancelling state
    covariances: chunk composites lose digits there, the scan notices (replay end state vs next entering state) and redoes
    itself sequentially -- the factor then equals the reference's to its own accuracy"""
    from oracle import tinygp_np as o
    rng = np.random.default_rng(33)
    n = 120
    X = np.sort(rng.uniform(0, 30.0, n))
    ko = o.qs.CARMA(np.array([0.1, 1.1]), np.array([1.0, 3.0]))
    d, p, q, a = o.qs_generators_fast(ko, X)
    c, w = oq.cholesky(d + 0.05, p, q, a)
    cov = oq.condition_qsm(oq.QSM(c, (p, w, a)), oq.QSM(d, (p, q, a), symm=True), np.full(n, 0.05))
    lo = cov.lower
    M = core.SymmQSM(diag=core.DiagQSM(cov.d), lower=core.StrictLowerTriQSM(*lo))
    before = _redos(backend)
    backend.set_option("qsm_chunk", 8)
    ch = M.cholesky()
    assert _redos(backend) == before + 1
    co, wo = oq.cholesky(cov.d, *lo)
    # two implementations of this ill-conditioned sequential recursion agree to ~1e-10 (the chunk composites: 3e-9 / 3e-6)
    np.testing.assert_allclose(ch.diag.d, co, rtol=3e-10, atol=1e-12)
    np.testing.assert_allclose(ch.lower.q, wo, rtol=0, atol=2e-8)


def _redos(backend):
    from ctypes import byref, c_int64
    v = c_int64()
    backend.lib.b200gp_get_option.restype = int
    backend.check(backend.lib.b200gp_get_option(backend.handle, b"qsm_sequential_redos", byref(v)))
    return v.value
