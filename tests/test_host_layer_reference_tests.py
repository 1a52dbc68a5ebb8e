"""The reference's own host-level tests (tests/test_gp.py, tests/test_noise.py, tests/test_kernels/test_kernels.py,
tests/test_kernels/test_distance.py, tests/test_transforms.py of dfm/tinygp), restated against `tinygp_b200` -- same
data, same assertions, same tolerances (5e-7, src/tinygp/test_utils.py:9-26) -- and run on the CPU with the C-ABI replaced
by tests/hostmock.py.  What they check is the product's Python layer (algebra, shapes, error behaviour, lowering);
cases the B200 backend refuses (Custom, DotProduct, Polynomial, noise.Dense/Banded, user Distance) assert the refusal."""

import os
import sys
from ctypes import c_void_p

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import hostmock  # noqa: E402
import tinygp_b200 as tinygp  # noqa: E402
from tinygp_b200 import GaussianProcess, _cabi, kernels, noise, transforms  # noqa: E402
from tinygp_b200.solvers import DirectSolver  # noqa: E402


def assert_allclose(a, b, atol=5e-7, rtol=5e-7):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), atol=atol, rtol=rtol)


@pytest.fixture(autouse=True)
def mocklib():
    lib = hostmock.MockLib()
    ctx = _cabi.Context.__new__(_cabi.Context)
    ctx.lib, ctx.handle, ctx.device = lib, c_void_p(1), -1
    previous = _cabi._ctx
    _cabi.set_context(ctx)
    try:
        yield lib
    finally:
        _cabi.set_context(previous)


@pytest.fixture
def random():
    return np.random.default_rng(1058390)


# ---- tests/test_gp.py ---------------------------------------------------------------------------------------
@pytest.fixture
def gp_data(random):
    X = random.uniform(-3, 3, (50, 5))
    y = random.normal(len(X))
    return X, y


def test_sample(gp_data):                                            # test_gp.py:24-38
    X, _ = gp_data
    gp = GaussianProcess(kernels.Matern32(1.5), X, diag=0.01, mean=np.sum)
    y = gp.sample(543)
    assert y.shape == (len(X),)
    y = gp.sample(543, shape=(7, 3))
    assert y.shape == (7, 3, len(X))
    y = gp.sample(543, shape=(100_000,))
    assert y.shape == (100_000, len(X))
    assert_allclose(np.mean(y, axis=0), np.sum(X, axis=1), atol=0.015)
    assert_allclose(np.cov(y, rowvar=False), gp.covariance, atol=0.015)


def test_means(gp_data):                                             # test_gp.py:41-52
    X, _ = gp_data
    y = np.sin(X[:, 0])
    gp1 = GaussianProcess(kernels.Matern32(1.5), X, diag=0.01, mean=lambda x: 0.0)
    gp2 = GaussianProcess(kernels.Matern32(1.5), X, diag=0.01, mean=0.0)
    gp3 = GaussianProcess(kernels.Matern32(1.5), X, diag=0.01)
    assert_allclose(gp1.mean, gp2.mean)
    assert_allclose(gp1.mean, gp3.mean)
    assert_allclose(gp1.log_probability(y), gp2.log_probability(y))
    assert_allclose(gp1.log_probability(y), gp3.log_probability(y))


def test_condition_shape_error(gp_data):                             # test_gp.py:55-76 (array branch; pytrees refused)
    X, _ = gp_data
    y = np.sin(X[:, 0])
    gp = GaussianProcess(kernels.ExpSquared(distance=kernels.L2Distance()), X, diag=0.1)
    gp.condition(y, X[0][None])
    with pytest.raises(ValueError):
        gp.condition(y, X[0])
    with pytest.raises((ValueError, TypeError, NotImplementedError)):
        GaussianProcess(kernels.ExpSquared(), {"x": X}, diag=0.1)


# ---- tests/test_noise.py ------------------------------------------------------------------------------------
def check_noise_model(nz, dense):                                    # test_noise.py:10-37
    rng = np.random.default_rng(6675)
    assert_allclose(nz.diagonal(), np.diag(dense))
    assert_allclose(nz + np.zeros_like(dense), dense)
    y1 = rng.normal(size=dense.shape)
    assert_allclose(nz + y1, dense + y1)
    assert_allclose(y1 + nz, y1 + dense)
    assert_allclose(nz @ y1, dense @ y1)
    y2 = rng.normal(size=(dense.shape[1], 3))
    assert_allclose(nz @ y2, dense @ y2)
    y3 = rng.normal(size=dense.shape[1])
    assert_allclose(nz @ y3, dense @ y3)
    try:
        qsm = nz.to_qsm()
    except NotImplementedError:
        pass
    else:
        assert_allclose(qsm @ y1, dense @ y1)
        assert_allclose(qsm @ y2, dense @ y2)
        assert_allclose(qsm @ y3, dense @ y3)


def test_diagonal():                                                 # test_noise.py:40-45
    N = 50
    diag = np.random.default_rng(9432).normal(size=N)
    check_noise_model(tinygp.noise.Diagonal(diag=diag), np.diag(diag))
    with pytest.raises(ValueError):                                  # noise.py:67-72
        tinygp.noise.Diagonal(diag=np.float64(0.1))


def test_banded():                                                   # test_noise.py:48-65
    N, J = 50, 5
    R = np.random.default_rng(9432).normal(size=(N, N))
    R[np.triu_indices(N, J + 1)] = 0
    R[np.tril_indices(N)] = R.T[np.tril_indices(N)]
    off_diags = np.zeros((N, J))
    for j in range(J):
        off_diags[: N - j - 1, j] = R[(np.arange(0, N - j - 1), np.arange(j + 1, N))]
    check_noise_model(tinygp.noise.Banded(diag=np.diag(R), off_diags=off_diags), R)


def test_dense():                                                    # test_noise.py:68-73
    M = np.random.default_rng(9432).normal(size=(50, 50))
    check_noise_model(tinygp.noise.Dense(value=M), M)


def test_gp_with_banded_and_dense_noise_on_both_solvers():
    """noise.Banded on the QuasisepSolver (solver.py:73-74: a SymmQSM sum) equals the same model on the DirectSolver
    (direct.py:47-48), which equals noise.Dense of the dense band -- log_probability, conditioning at new points, at the inputs"""
    rng = np.random.default_rng(5)
    x = np.sort(rng.uniform(0, 10, 60))
    y, xs = np.sin(x), np.linspace(-1, 11, 9)
    banded = tinygp.noise.Banded(diag=rng.uniform(0.1, 0.2, 60), off_diags=0.02 * rng.normal(size=(60, 2)))
    dense = tinygp.noise.Dense(value=banded + np.zeros((60, 60)))
    kq = quasisep.Matern32(1.5) + quasisep.SHO(omega=1.2, quality=2.0, sigma=0.7)
    gps = [GaussianProcess(kq, x, noise=banded), GaussianProcess(kq, x, noise=banded, solver=DirectSolver),
           GaussianProcess(kq, x, noise=dense, solver=DirectSolver)]
    assert isinstance(gps[0].solver, QuasisepSolver) and gps[0].solver._generic
    ref = gps[2]
    for gp in gps[:2]:
        assert_allclose(gp.log_probability(y), ref.log_probability(y))
        assert_allclose(gp.covariance, ref.covariance)
        for kw in (dict(X_test=xs), dict()):
            a, b = gp.condition(y, **kw).gp, ref.condition(y, **kw).gp
            assert_allclose(a.loc, b.loc)
            assert_allclose(a.covariance, b.covariance)
    with pytest.raises(NotImplementedError):                         # noise.py:121-123
        GaussianProcess(kq, x, noise=dense)                          # -> QuasisepSolver -> dense.to_qsm()
    with pytest.raises(ValueError, match="Input coordinates must be sorted"):
        GaussianProcess(kq, x[::-1], noise=banded)


# ---- tests/test_kernels/test_kernels.py -----------------------------------------------------------------------
@pytest.fixture
def kdata(random):
    x1 = random.uniform(-3, 3, (50, 5))
    x2 = random.uniform(-5, 5, (50, 5))
    return x1, x2


def test_constant(kdata):                                            # test_kernels.py:24-43
    x1, x2 = kdata
    v = np.ones(3)
    with pytest.raises(ValueError):
        kernels.Constant(np.ones(3)).evaluate(v, v)
    with pytest.raises(ValueError):
        (np.ones(3) * kernels.Matern32(1.5)).evaluate(v, v)
    factor = 2.5
    k1 = kernels.Matern32(2.5)
    assert_allclose(factor * k1(x1, x2), (factor * k1)(x1, x2))


def test_custom_and_nonstationary_kernels_are_refused(kdata):        # test_kernels.py:46-63, 93-97
    x1, x2 = kdata
    for k in (kernels.Custom(lambda a, b: 1.0), kernels.DotProduct(), kernels.Polynomial(order=1.5, scale=0.5, sigma=1.3)):
        with pytest.raises(NotImplementedError, match="unsupported by the B200"):
            k(x1, x2)


def test_ops(kdata):                                                 # test_kernels.py:66-73
    x1, x2 = kdata
    k1 = 1.5 * kernels.Matern32(2.5)
    k2 = 0.9 * kernels.ExpSineSquared(scale=1.5, gamma=0.3)
    assert_allclose(k1(x1, x2) + k2(x1, x2), (k1 + k2)(x1, x2))
    assert_allclose(k1(x1, x2) * k2(x1, x2), (k1 * k2)(x1, x2))
    assert_allclose(sum([k1, k2])(x1, x2), (k1 + k2)(x1, x2))        # __radd__ with 0 (base.py:110-116)


def test_conditioned(kdata):                                         # test_kernels.py:76-90
    x1, x2 = kdata
    k1 = 1.5 * kernels.Matern32(2.5, distance=kernels.L2Distance())
    k2 = 0.9 * kernels.ExpSquared(scale=1.5)
    K = k1(x1, x1) + 0.1 * np.eye(x1.shape[0])
    solver = DirectSolver.init(k1, x1, noise.Diagonal(np.full(x1.shape[0], 0.1)))
    cond = kernels.Conditioned(x1, solver, k2)
    assert_allclose(cond(x1, x2), k2(x1, x2) - k2(x1, x1) @ np.linalg.solve(K, k2(x1, x2)))


@pytest.mark.parametrize("kernel", [kernels.ExpSineSquared, kernels.RationalQuadratic])
def test_required_parameters(kernel):                                # test_kernels.py:123-133
    with pytest.raises(ValueError):
        kernel(0.5)(np.zeros((2, 1)), np.zeros((2, 1)))


# ---- tests/test_kernels/test_distance.py ------------------------------------------------------------------------
def test_distances_through_the_lowering(kdata):
    """L1 / L2 distance definitions (distance.py:41-59) as seen through the leaves that use them"""
    x1, x2 = kdata
    d1 = np.sum(np.abs(x1[:, None, :] - x2[None, :, :]), axis=-1)
    d2 = np.sqrt(np.sum((x1[:, None, :] - x2[None, :, :]) ** 2, axis=-1))
    assert_allclose(kernels.Exp(1.0, distance=kernels.L1Distance())(x1, x2), np.exp(-d1))
    assert_allclose(kernels.Exp(1.0, distance=kernels.L2Distance())(x1, x2), np.exp(-d2))
    assert_allclose(kernels.ExpSquared(1.0)(x1, x1)[np.arange(50), np.arange(50)], np.ones(50))   # zero distance, no NaN
    with pytest.raises(NotImplementedError, match="unsupported by the B200"):
        class Custom(kernels.Distance):
            def distance(self, X1, X2):
                return 0.0
        kernels.Exp(1.0, distance=Custom())(x1, x2)


# ---- tests/test_transforms.py -----------------------------------------------------------------------------------
def test_linear():                                                   # test_transforms.py:9-16
    kernel0 = kernels.Matern32(4.5)
    kernel1 = transforms.Linear(1 / 4.5, kernels.Matern32())
    assert_allclose(kernel0.evaluate(0.5, 0.1), kernel1.evaluate(0.5, 0.1))


def test_multivariate_linear(random):                                # test_transforms.py:19-25
    x1, x2 = random.normal(size=(2, 3))
    L = np.linalg.cholesky(np.cov(random.normal(size=(3, 100))))
    kernel1 = transforms.Linear(np.linalg.inv(L), kernels.Matern32(distance=kernels.L2Distance()))
    r = np.sqrt(np.sum(np.linalg.solve(L, x1 - x2) ** 2))
    assert_allclose(kernel1.evaluate(x1, x2), (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r))


def test_cholesky_and_subspace(random):                              # test_transforms.py:28-44
    kernel0 = kernels.Matern32(4.5)
    kernel1 = transforms.Cholesky(4.5, kernels.Matern32())
    assert_allclose(kernel0.evaluate(0.5, 0.1), kernel1.evaluate(0.5, 0.1))
    x1, x2 = random.normal(size=(2, 3))
    L = np.linalg.cholesky(np.cov(random.normal(size=(3, 100))))
    kernel2 = transforms.Cholesky(L, kernels.Matern32(distance=kernels.L2Distance()))
    kernel3 = transforms.Linear(np.linalg.inv(L), kernels.Matern32(distance=kernels.L2Distance()))
    assert_allclose(kernel2.evaluate(x1, x2), kernel3.evaluate(x1, x2))
    kernel4 = transforms.Subspace(1, kernels.Matern32())
    assert_allclose(kernel4.evaluate(np.array([0.5, 0.1]), np.array([-0.4, 0.7])),
                    kernels.Matern32().evaluate(np.array([0.1]), np.array([0.7])))


# ---- tests/test_solvers/test_quasisep/test_solver.py -------------------------------------------------------------
from tinygp_b200.kernels import quasisep  # noqa: E402
from tinygp_b200.solvers import QuasisepSolver  # noqa: E402

KERNEL_PAIRS = [
    (lambda: quasisep.Matern32(sigma=1.8, scale=1.5), lambda: 1.8 ** 2 * kernels.Matern32(1.5)),
    (lambda: 1.8 ** 2 * quasisep.Matern32(1.5), lambda: 1.8 ** 2 * kernels.Matern32(1.5)),
    (lambda: quasisep.Matern52(sigma=1.8, scale=1.5), lambda: 1.8 ** 2 * kernels.Matern52(1.5)),
    (lambda: quasisep.Exp(sigma=1.8, scale=1.5), lambda: 1.8 ** 2 * kernels.Exp(1.5)),
    (lambda: quasisep.Cosine(sigma=1.8, scale=1.5), lambda: 1.8 ** 2 * kernels.Cosine(1.5)),
    (lambda: quasisep.Matern32(sigma=1.8, scale=1.5) + quasisep.Matern52(sigma=0.9, scale=0.7),
     lambda: 1.8 ** 2 * kernels.Matern32(1.5) + 0.9 ** 2 * kernels.Matern52(0.7)),
]


@pytest.mark.parametrize("parallel", [False, True], ids=["sequential", "parallel"])
@pytest.mark.parametrize("pair", range(len(KERNEL_PAIRS)))
def test_consistent_with_direct(pair, parallel):                     # test_solver.py:62-103
    rng = np.random.default_rng(84930)
    x = np.sort(rng.uniform(-3, 3, 50))
    y = np.sin(x)
    t = np.sort(rng.uniform(-3, 3, 10))
    kernel0 = quasisep.Matern32(sigma=3.8, scale=4.5)
    kernel1, kernel2 = KERNEL_PAIRS[pair][0](), KERNEL_PAIRS[pair][1]()
    gp1 = GaussianProcess(kernel1, x, diag=0.1, solver=QuasisepSolver, parallel=parallel)
    gp2 = GaussianProcess(kernel2, x, diag=0.1, solver=DirectSolver)
    assert_allclose(gp1.covariance, gp2.covariance)
    assert_allclose(gp1.solver.normalization(), gp2.solver.normalization())
    assert_allclose(gp1.log_probability(y), gp2.log_probability(y))
    assert_allclose(gp1.sample(0), gp2.sample(0))
    assert_allclose(gp1.sample(0, shape=(5, 7)), gp2.sample(0, shape=(5, 7)))

    gp1p, gp2p = gp1.condition(y), gp2.condition(y)
    assert_allclose(gp1p.log_probability, gp2p.log_probability)
    assert_allclose(gp1p.gp.loc, gp2p.gp.loc)
    assert_allclose(gp1p.gp.variance, gp2p.gp.variance)
    assert_allclose(gp1p.gp.covariance, gp2p.gp.covariance)

    gp1p, gp2p = gp1.condition(y, kernel=kernel0), gp2.condition(y, kernel=kernel0)
    assert_allclose(gp1p.log_probability, gp2p.log_probability)
    assert_allclose(gp1p.gp.loc, gp2p.gp.loc)
    assert_allclose(gp1p.gp.variance, gp2p.gp.variance)
    assert_allclose(gp1p.gp.covariance, gp2p.gp.covariance)

    gp1p, gp2p = gp1.condition(y, X_test=t, kernel=kernel0), gp2.condition(y, X_test=t, kernel=kernel0)
    assert not isinstance(gp1p.gp.solver, QuasisepSolver)
    assert_allclose(gp1p.log_probability, gp2p.log_probability)
    assert_allclose(gp1p.gp.loc, gp2p.gp.loc)
    assert_allclose(gp1p.gp.variance, gp2p.gp.variance)
    assert_allclose(gp1p.gp.covariance, gp2p.gp.covariance)


def test_unsorted():                                                 # test_solver.py:127-143
    rng = np.random.default_rng(0)
    x_ = rng.uniform(-3, 3, 50)                                      # not sorted
    kernel = quasisep.Matern32(sigma=1.8, scale=1.5)
    with pytest.raises(ValueError, match="Input coordinates must be sorted"):
        GaussianProcess(kernel, x_, diag=0.1)


# ---- tests/test_kernels/test_quasisep.py ----------------------------------------------------------------------------
QS_KERNELS = [
    lambda: quasisep.Matern32(sigma=1.8, scale=1.5),
    lambda: quasisep.Matern32(1.5),
    lambda: quasisep.Matern52(sigma=1.8, scale=1.5),
    lambda: quasisep.Matern52(1.5),
    lambda: quasisep.Celerite(1.1, 0.8, 0.9, 0.1),
    lambda: quasisep.SHO(omega=1.5, quality=0.5, sigma=1.3),
    lambda: quasisep.SHO(omega=1.5, quality=3.5, sigma=1.3),
    lambda: quasisep.SHO(omega=1.5, quality=0.1, sigma=1.3),
    lambda: quasisep.Exp(sigma=1.8, scale=1.5),
    lambda: quasisep.Exp(1.5),
    lambda: 1.5 * quasisep.Matern52(1.5) + 0.3 * quasisep.Exp(1.5),
    lambda: quasisep.Cosine(sigma=1.8, scale=1.5),
    lambda: 1.8 * quasisep.Cosine(1.5),
    lambda: quasisep.Matern52(1.5) * quasisep.SHO(omega=1.5, quality=0.1),      # test_quasisep.py's Product entry
    lambda: quasisep.CARMA(alpha=np.array([1.4, 2.3, 1.5]), beta=np.array([0.1, 0.5])),      # test_quasisep.py:43-46
    lambda: quasisep.CARMA(alpha=np.array([1, 1.2]), beta=np.array([1.0, 3.0])),
    lambda: quasisep.CARMA(alpha=np.array([0.1, 1.1]), beta=np.array([1.0, 3.0])),
    lambda: quasisep.CARMA(alpha=np.array([1.0 / 100]), beta=np.array([0.3])),
    lambda: (quasisep.Matern52(1.5) + 0.4 * quasisep.Exp(0.7)) * quasisep.SHO(omega=1.5, quality=0.1),      # multiplied out
    lambda: (quasisep.Exp(1.5) + quasisep.Exp(0.7)) * (quasisep.Cosine(2.5) + 0.5 * quasisep.Exp(1.1)),
    lambda: quasisep.CARMA(alpha=np.array([1, 1.2]), beta=np.array([1.0, 3.0])) * quasisep.Exp(2.0),       # CARMA2 row first
]


@pytest.mark.parametrize("which", range(len(QS_KERNELS)))
def test_quasisep_kernels(which):                                    # test_quasisep.py:53-72 (matrix identities)
    rng = np.random.default_rng(84930)
    x = np.sort(rng.uniform(-3, 3, 50))
    y = np.sin(x)
    t = np.sort(rng.uniform(-3, 3, 12))
    kernel = QS_KERNELS[which]()
    K = kernel(x, x)
    diag = np.full(50, 0.3)
    dense = QuasisepSolver(kernel, x, noise.Diagonal(diag)).covariance()      # to_symm_qsm(x).to_dense() + noise
    assert_allclose(dense - np.diag(diag), K)
    assert_allclose(kernel.to_symm_qsm(x).to_dense(), K)                       # test_quasisep.py:58 as written
    assert_allclose(kernel.matmul(x, y), K @ y)
    assert_allclose(kernel.matmul(t, x, y), kernel(t, x) @ y)
    gq = kernel.to_general_qsm(t, x)                                           # kernels/quasisep.py:118-145, general.py
    assert gq.shape == (12, 50)
    assert_allclose(gq @ y, kernel(t, x) @ y)
    assert np.array_equal(gq.idx, np.searchsorted(x, t, side="right") - 1)
    # the state-space model itself (test_quasisep.py:66-72): the kernel value from (h, Pinf, A), and "F is defined
    # consistently with the transition matrix"
    from scipy.linalg import expm
    from tinygp_b200.solvers.quasisep.block import ensure_dense
    F, Pinf = kernel.design_matrix(), kernel.stationary_covariance()
    for x1, x2 in ((x[3], x[9]), (x[0], x[0]), (x[7], x[41])):
        A = kernel.transition_matrix(x1, x2)                  # a Block for a Sum (quasisep.py:262-270)
        assert_allclose(expm(ensure_dense(F).T * (x2 - x1)), ensure_dense(A))
        assert_allclose(kernel.observation_model(x2) @ A.T @ Pinf @ kernel.observation_model(x1), kernel.evaluate(x1, x2))
        assert_allclose(kernel.observation_model(x2) @ ensure_dense(A).T @ ensure_dense(Pinf) @ kernel.observation_model(x1),
                        kernel.evaluate(x1, x2))


# ---- tests/test_solvers/test_quasisep/test_block.py ------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(10,), (10, 3), (4, 10, 3), (2, 4, 10, 3)])
def test_block(shape):                                               # test_block.py:13-33
    from scipy.linalg import block_diag
    from tinygp_b200.solvers.quasisep.block import Block
    rng = np.random.default_rng(1234)
    block = Block(*(rng.uniform(size=(s, s)) for s in [1, 3, 2, 4]))
    x = rng.uniform(size=shape)
    xt = x if x.ndim == 1 else np.swapaxes(x, -1, -2)
    block_ = block_diag(*block.blocks)
    assert len(block) == len(block_)
    assert block.shape == block_.shape
    assert_allclose(block @ x, block_ @ x)
    assert_allclose(xt @ block, xt @ block_)
    assert_allclose(block.T @ x, block_.T @ x)
    assert_allclose(xt @ block.T, xt @ block_.T)
    assert_allclose(block.to_dense(), block_)
    assert_allclose((block @ block.T).to_dense(), block_ @ block_.T)
    assert_allclose((2.0 * block + block).to_dense(), 3.0 * block_)
    assert_allclose(block - block_, np.zeros_like(block_))


def test_sum_state_space_is_blocked_like_the_reference():            # kernels/quasisep.py:241-295
    from tinygp_b200.solvers.quasisep.block import Block
    k = quasisep.Matern32(1.5) + quasisep.SHO(omega=1.5, quality=0.7) + quasisep.Exp(0.4)
    T = k.transition_matrix(0.1, 0.9)
    assert isinstance(T, Block) and [b.shape for b in T.blocks] == [(2, 2), (2, 2), (1, 1)]      # not nested (issue 265)
    assert isinstance(k.design_matrix(), Block) and isinstance(k.stationary_covariance(), Block)
    dense = quasisep.Sum(quasisep.Matern32(1.5) + quasisep.SHO(omega=1.5, quality=0.7), quasisep.Exp(0.4), use_block=False)
    assert isinstance(dense.transition_matrix(0.1, 0.9), np.ndarray)
    assert_allclose(dense.transition_matrix(0.1, 0.9), T.to_dense())
    assert_allclose(dense.transition_matrix(0.1, 0.9), quasisep.Quasisep.transition_matrix(k, 0.1, 0.9))   # from the device rows


def test_wrapper_kernels():                                          # kernels/quasisep.py:218-238
    class TimeColumn(quasisep.Wrapper):                              # structured inputs (t, band): the process lives on t
        def coord_to_sortable(self, X):
            return np.asarray(X)[..., 0]

    rng = np.random.default_rng(11)
    t = np.sort(rng.uniform(0, 10, 40))
    X = np.stack((t, rng.integers(0, 2, 40).astype(float)), axis=1)
    tt = rng.uniform(-1, 11, 6)
    Xt = np.stack((tt, np.zeros(6)), axis=1)
    y = np.sin(t)
    base = quasisep.Matern32(1.5) + 0.5 * quasisep.SHO(omega=1.2, quality=2.0)
    wrapped = TimeColumn(base)
    assert_allclose(wrapped(X, Xt), base(t, tt))
    assert_allclose(wrapped.to_symm_qsm(X).to_dense(), base(t, t))
    assert_allclose(wrapped.matmul(Xt, X, y), base(tt, t) @ y)
    assert_allclose(wrapped.to_general_qsm(Xt, X) @ y, base(tt, t) @ y)
    assert_allclose(wrapped.transition_matrix(X[3], X[8]).to_dense(), base.transition_matrix(t[3], t[8]).to_dense())
    g1, g2 = GaussianProcess(wrapped, X, diag=0.1), GaussianProcess(base, t, diag=0.1)
    assert isinstance(g1.solver, QuasisepSolver)
    assert_allclose(g1.log_probability(y), g2.log_probability(y))
    for kw1, kw2 in ((dict(X_test=Xt), dict(X_test=tt)), (dict(), dict())):
        c1, c2 = g1.condition(y, **kw1).gp, g2.condition(y, **kw2).gp
        assert_allclose(c1.loc, c2.loc)
        assert_allclose(c1.covariance, c2.covariance)
    with pytest.raises(ValueError, match="Input coordinates must be sorted"):
        GaussianProcess(wrapped, X[::-1], diag=0.1)

    class Multiband(TimeColumn):     # a coordinate-dependent observation model (the reference's multiband tutorial): no device
        amp = np.array([1.0, 0.6])   # state-space rows for that -> generators from these Python methods, device QSM algebra
        def observation_model(self, X):
            return self.amp[int(X[1])] * self.kernel.observation_model(X[0])

    mb = Multiband(base)
    assert not mb._on_device() and not mb._has_closed_form()
    band = X[:, 1].astype(int)
    want = np.outer(mb.amp[band], mb.amp[band]) * base(t, t)
    assert_allclose(mb(X, X), want)
    assert_allclose(mb(X), np.diag(want))
    assert_allclose(mb.to_symm_qsm(X).to_dense(), want)
    cross = mb.amp[0] * mb.amp[band][None, :] * base(tt, t)
    assert_allclose(mb(Xt, X), cross)
    assert_allclose(mb.matmul(Xt, X, y), cross @ y)
    assert_allclose(mb.evaluate(X[3], X[9]), want[3, 9])
    assert_allclose(mb.evaluate(X[9], X[3]), want[3, 9])
    g3 = GaussianProcess(mb, X, diag=0.1)
    g4 = GaussianProcess(mb, X, diag=0.1, solver=DirectSolver)
    assert isinstance(g3.solver, QuasisepSolver) and g3.solver._generic
    Sigma = want + 0.1 * np.eye(40)
    sign, logdet = np.linalg.slogdet(Sigma)
    lp = -0.5 * y @ np.linalg.solve(Sigma, y) - 0.5 * logdet - 20 * np.log(2 * np.pi)
    assert_allclose(g3.log_probability(y), lp)
    assert_allclose(g4.log_probability(y), lp)
    for kw in (dict(X_test=Xt), dict()):
        c3, c4 = g3.condition(y, **kw).gp, g4.condition(y, **kw).gp
        assert_allclose(c3.loc, c4.loc)
        assert_allclose(c3.covariance, c4.covariance)
    assert_allclose(g3.condition(y, Xt).gp.loc, cross @ np.linalg.solve(Sigma, y))


# ---- tests/test_kernels/test_quasisep_nonreversible.py (structured (N, 2) arrays instead of pytree coordinates) ----------
class CausalFilter(quasisep.Quasisep):
    """A minimal non-reversible two-state process: driver -> response."""

    def design_matrix(self):
        return np.array([[-1.0, 0.0], [0.8, -2.0]])

    def stationary_covariance(self):
        return np.array([[0.5, 2.0 / 15.0], [2.0 / 15.0, 91.0 / 300.0]])

    def observation_model(self, X):
        return np.eye(2)[int(X[1])]

    def coord_to_sortable(self, X):
        return X[0]

    def transition_matrix(self, X1, X2):
        from scipy.linalg import expm
        return expm(self.design_matrix().T * (X2[0] - X1[0]))


def _direct_covariance(kernel, X1, X2):                               # test_quasisep_nonreversible.py:36-50
    Pinf = kernel.stationary_covariance()
    out = np.empty((len(X1), len(X2)))
    for i, x1 in enumerate(X1):
        for j, x2 in enumerate(X2):
            h1, h2 = kernel.observation_model(x1), kernel.observation_model(x2)
            out[i, j] = (h2 @ kernel.transition_matrix(x1, x2).T @ Pinf @ h1 if x1[0] < x2[0]
                         else h1 @ kernel.transition_matrix(x2, x1).T @ Pinf @ h2)
    return out


_NR_X = np.stack(([0.0, 0.3, 0.8, 1.4, 2.2, 3.1], [0, 1, 0, 1, 1, 0]), axis=1)
_NR_XT = np.stack(([0.1, 0.9, 1.8, 2.8], [1, 0, 1, 0]), axis=1)


def test_nonreversible_covariance_and_cross_matmul():                # test_quasisep_nonreversible.py:53-78
    kernel = CausalFilter()
    expected = _direct_covariance(kernel, _NR_X, _NR_X)
    assert_allclose(kernel(_NR_X, _NR_X), expected)
    assert_allclose(kernel.to_symm_qsm(_NR_X).to_dense(), expected)
    cross = _direct_covariance(kernel, _NR_XT, _NR_X)
    y = np.linspace(-0.5, 0.7, 6)
    assert_allclose(kernel.matmul(_NR_XT, _NR_X, y), cross @ y)
    a = _direct_covariance(kernel, np.array([[1.0, 1]]), np.array([[0.0, 0]]))[0, 0]
    b = _direct_covariance(kernel, np.array([[1.0, 0]]), np.array([[0.0, 1]]))[0, 0]
    assert not np.isclose(a, b)
    assert_allclose(kernel.evaluate(np.array([1.0, 1]), np.array([0.0, 0])), a)


def test_nonreversible_solvers_and_conditioning_agree():             # test_quasisep_nonreversible.py:81-106
    kernel = CausalFilter()
    y = np.array([0.2, -0.1, 0.3, 0.15, -0.2, 0.05])
    diag = np.full(6, 0.05)
    gp_direct = GaussianProcess(kernel, _NR_X, diag=diag, solver=DirectSolver)
    gp_quasisep = GaussianProcess(kernel, _NR_X, diag=diag, solver=QuasisepSolver)
    assert_allclose(gp_quasisep.covariance, gp_direct.covariance)
    assert_allclose(gp_quasisep.log_probability(y), gp_direct.log_probability(y))
    Sigma = _direct_covariance(kernel, _NR_X, _NR_X) + np.diag(diag)
    sign, logdet = np.linalg.slogdet(Sigma)
    assert_allclose(gp_direct.log_probability(y), -0.5 * y @ np.linalg.solve(Sigma, y) - 0.5 * logdet - 3 * np.log(2 * np.pi))
    for kw in (dict(), dict(X_test=_NR_XT)):
        cd, cq = gp_direct.condition(y, **kw), gp_quasisep.condition(y, **kw)
        assert_allclose(cq.gp.loc, cd.gp.loc)
        assert_allclose(cq.gp.covariance, cd.gp.covariance)


def test_user_defined_kernel_reproduces_the_reference():
    """the same CausalFilter run by the UNMODIFIED reference (tests/golden/make_golden_custom.py -> custom_kernel_vectors.json,
    pytree coordinates there, (N, 2) arrays here): kernel values, QSM, cross product, both solvers, conditioning"""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "custom_kernel_vectors.json")) as f:
        g = json.load(f)
    X = np.stack((g["time"], g["channel"]), axis=1).astype(float)
    Xt = np.stack((g["time_test"], g["channel_test"]), axis=1).astype(float)
    y, diag = np.array(g["y"]), np.full(len(g["time"]), g["diag"])
    kernel = CausalFilter()
    tol = dict(atol=1e-12, rtol=1e-10)
    assert_allclose(kernel(X, X), g["K"], **tol)
    assert_allclose(kernel(Xt, X), g["K_cross"], **tol)
    assert_allclose(kernel.to_symm_qsm(X).to_dense(), g["symm_qsm_dense"], **tol)
    assert_allclose(kernel.matmul(Xt, X, y), g["cross_matmul"], **tol)
    for name, solver in (("quasisep", QuasisepSolver), ("direct", DirectSolver)):
        gp, want = GaussianProcess(kernel, X, diag=diag, solver=solver), g[name]
        assert_allclose(gp.log_probability(y), want["log_probability"], **tol)
        assert_allclose(gp.covariance, want["covariance"], **tol)
        c_in, c_test = gp.condition(y), gp.condition(y, X_test=Xt)
        assert_allclose(c_in.gp.loc, want["cond_in_loc"], atol=1e-10, rtol=1e-9)
        assert_allclose(c_in.gp.covariance, want["cond_in_cov"], atol=1e-10, rtol=1e-9)
        assert_allclose(c_test.gp.loc, want["cond_test_loc"], atol=1e-10, rtol=1e-9)
        assert_allclose(c_test.gp.covariance, want["cond_test_cov"], atol=1e-10, rtol=1e-9)


def test_multiband_wrapper_reproduces_the_reference():
    """a Wrapper whose observation model depends on the band column, run by the unmodified reference (make_golden_custom.py)"""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "custom_kernel_vectors.json")) as f:
        g = json.load(f)["multiband"]

    class Multiband(quasisep.Wrapper):
        def __init__(self, kernel, amplitudes):
            self.kernel, self.amplitudes = kernel, np.asarray(amplitudes)

        def coord_to_sortable(self, X):
            return X[0]

        def observation_model(self, X):
            return self.amplitudes[int(X[1])] * self.kernel.observation_model(X[0])

    X = np.stack((g["t"], g["band"]), axis=1).astype(float)
    Xt = np.stack((g["t_test"], g["band_test"]), axis=1).astype(float)
    y = np.array(g["y"])
    mb = Multiband(quasisep.Matern32(1.5) + 0.5 * quasisep.SHO(omega=1.2, quality=2.0), g["amplitudes"])
    tol = dict(atol=1e-11, rtol=1e-9)
    assert_allclose(mb(X, X), g["K"], **tol)
    assert_allclose(mb(Xt, X), g["K_cross"], **tol)
    gp = GaussianProcess(mb, X, diag=0.1)
    assert_allclose(gp.log_probability(y), g["log_probability"], **tol)
    c_in, c_test = gp.condition(y), gp.condition(y, X_test=Xt)
    assert_allclose(c_in.gp.loc, g["cond_in_loc"], **tol)
    assert_allclose(c_in.gp.variance, g["cond_in_var"], **tol)
    assert_allclose(c_test.gp.loc, g["cond_test_loc"], **tol)
    assert_allclose(c_test.gp.covariance, g["cond_test_cov"], **tol)


def test_models_with_more_than_eight_states_use_generator_arrays():
    """Matern52 + Matern52 + SHO + SHO = 10 states: above the 8 the model-specialised kernels compile, so the generators are
    evaluated on the host and the device works on generator arrays of order 10 (QuasisepSolver's generic mode)"""
    rng = np.random.default_rng(2)
    t = np.sort(rng.uniform(0, 10, 50))
    y, tt = np.sin(t), rng.uniform(-1, 11, 5)
    k = (quasisep.Matern52(1.5) + 0.5 * quasisep.Matern52(0.4) + quasisep.SHO(omega=1.2, quality=2.0, sigma=0.7)
         + quasisep.SHO(omega=3.0, quality=0.3, sigma=0.4))
    assert not k._on_device() and k._has_closed_form()
    gq, gd = GaussianProcess(k, t, diag=0.1), GaussianProcess(k, t, diag=0.1, solver=DirectSolver)
    assert isinstance(gq.solver, QuasisepSolver) and gq.solver._generic
    assert_allclose(gq.log_probability(y), gd.log_probability(y))
    assert_allclose(k.to_symm_qsm(t).to_dense(), k(t, t))
    assert_allclose(k.matmul(tt, t, y), k(tt, t) @ y)
    cq, cd = gq.condition(y, tt).gp, gd.condition(y, tt).gp
    assert_allclose(cq.loc, cd.loc)
    assert_allclose(cq.covariance, cd.covariance + 0.0)


def test_vectorised_generators_equal_the_per_point_formulas():
    """models with more than 8 states: `_host_symm_qsm` evaluates the component rows' closed forms for all time steps at once;
    the per-point path (quasisep.py:102-116 through the kernel's methods, what a user-defined subclass gets) gives the same arrays"""
    t = np.sort(np.random.default_rng(3).uniform(0, 60, 400))
    t[7] = t[6]
    k = (quasisep.Celerite(1.1, 0.1, 0.3, 1.5) + quasisep.Cosine(3.0, 0.7) * quasisep.Exp(2.0) + quasisep.SHO(1.2, 0.5, 1.1)
         + quasisep.SHO(1.2, 0.2, 1.1) + 0.3 * quasisep.Exp(0.5)
         + quasisep.CARMA(alpha=np.array([1.4, 2.3, 1.5]), beta=np.array([0.1, 0.5])))
    assert k.state_dim() == 12 and not k._on_device()
    fast = k._host_symm_qsm(t)

    def no_rows():
        raise NotImplementedError
    k.components = no_rows                       # the Sum's own state-space methods still work (its leaves keep their rows)
    try:
        slow = k._host_symm_qsm(t)
    finally:
        del k.components
    assert_allclose(fast.diag.d, slow.diag.d, atol=1e-13, rtol=1e-12)
    for part in ("p", "q", "a"):
        assert_allclose(getattr(fast.lower, part), getattr(slow.lower, part), atol=1e-13, rtol=1e-12)


def test_oversized_products_use_generator_arrays():                  # more states than the model-specialised kernels compile
    t = np.linspace(0, 4, 30)
    y = np.cos(t)
    for k in ((quasisep.Matern52(1.5) + quasisep.Matern32(0.7)) * quasisep.SHO(omega=1.5, quality=0.1),      # 6 + 4 states
              quasisep.Matern52(1.5) * quasisep.Matern52(0.5)):                                              # 3 x 3 = 9 states
        assert not k._on_device()
        gq = GaussianProcess(k, t, diag=0.1)
        assert gq.solver._generic
        assert_allclose(gq.log_probability(y), GaussianProcess(k, t, diag=0.1, solver=DirectSolver).log_probability(y))
        assert_allclose(k.to_symm_qsm(t).to_dense(), k(t, t))


def test_carma():                                                    # test_quasisep.py:100-122
    rng = np.random.default_rng(84930)
    x = np.sort(rng.uniform(-3, 3, 50))
    y = np.sin(x)
    carma2_kernels = [
        quasisep.CARMA(alpha=np.array([0.01]), beta=np.array([0.1])),
        quasisep.CARMA(alpha=np.array([1.0, 1.2]), beta=np.array([1.0, 3.0])),
        quasisep.CARMA(alpha=np.array([0.1, 1.1]), beta=np.array([1.0, 3.0])),
    ]
    validate_kernels = [                                             # equivalent Celerite + Exp kernels
        quasisep.Exp(scale=100.0, sigma=np.sqrt(0.5)),
        quasisep.Celerite(25.0 / 6, 2.5, 0.6, -0.8),
        quasisep.Exp(1.0, np.sqrt(4.04040404)) + quasisep.Exp(10.0, np.sqrt(4.5959596)),
    ]
    for k1, k2 in zip(carma2_kernels, validate_kernels):
        gp1 = GaussianProcess(k1, x, diag=0.1)
        gp2 = GaussianProcess(k2, x, diag=0.1)
        assert_allclose(gp1.log_probability(y), gp2.log_probability(y))
        assert_allclose(gp1.solver.normalization(), gp2.solver.normalization())


def test_carma_quads():                                              # test_quasisep.py:141-160
    alpha = np.array([1.4, 2.3, 1.5])
    beta = np.array([0.1, 0.5])
    alpha_quads = quasisep.carma_poly2quads(np.append(alpha, 1.0))
    beta_quads = quasisep.carma_poly2quads(beta)
    alpha_quads, beta_mult, beta_quads = alpha_quads[:-1], beta_quads[-1], beta_quads[:-1]
    carma31 = quasisep.CARMA.init(alpha=alpha, beta=beta)
    carma31_quads = quasisep.CARMA.from_quads(alpha_quads=alpha_quads, beta_quads=beta_quads, beta_mult=beta_mult)
    assert_allclose(carma31.arroots, carma31_quads.arroots)
    assert_allclose(carma31.acf, carma31_quads.acf)
    assert_allclose(carma31.obsmodel, carma31_quads.obsmodel)


def test_celerite():                                                 # test_quasisep.py:83-97
    a, b, c, d = 1.1, 0.8, 0.9, 0.1
    kernel = quasisep.Celerite(a, b, c, d)
    rng = np.random.default_rng(84930)
    x = np.sort(rng.uniform(-3, 3, 50))
    t = np.sort(rng.uniform(-3, 3, 12))
    tau = np.abs(x[:, None] - x[None, :])
    assert_allclose(kernel(x, x), np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau)))
    tau = np.abs(x[:, None] - t[None, :])
    assert_allclose(kernel(x, t), np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau)))
