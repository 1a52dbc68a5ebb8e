"""QSM algebra on the GPU (libb200gp.so: qsm.cu, one warp per chunk): the same checks as the CPU twin
(test_qsm_device_code_on_host.py) against the reference's golden outputs and the oracle, plus the end-to-end
QSM-valued conditioning of a quasiseparable GaussianProcess (solver.py:124-129) and a size with thousands of chunks."""

import numpy as np
import pytest

import qsmchecks
from oracle import tinygp_np as o
from qsmutil import GOLD, qsmcases
from tinygp_b200 import GaussianProcess
from tinygp_b200.kernels import quasisep
from tinygp_b200.solvers import QuasisepSolver
from tinygp_b200.solvers.quasisep import core

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 3, 7])
def ops(ctx, request):
    ctx.set_option("qsm_chunk", request.param)
    return qsmchecks.operands()


def test_dense_forms_parts_and_scaling(ops):
    qsmchecks.check_dense_and_parts(ops)


def test_qsm_mul_every_type_pair(ops):
    qsmchecks.check_products(ops)


def test_elementwise_sum_and_product(ops):
    qsmchecks.check_sums(ops)


def test_inverses_gram_cholesky_solves(ops):
    qsmchecks.check_inverses_and_factor(ops)


@pytest.mark.parametrize("chunk", [0, 2, 5])
@pytest.mark.parametrize("case", qsmcases.CONDITION, ids=lambda c: c["name"])
def test_conditioned_covariance_generators(ctx, case, chunk):
    ctx.set_option("qsm_chunk", chunk)
    qsmchecks.check_condition_algebra(case)


@pytest.mark.parametrize("n,m1,m2,chunk", [(300, 3, 2, 0), (300, 2, 5, 16), (1000, 4, 4, 0), (3000, 6, 3, 0)])
def test_many_chunks_against_the_oracle(ctx, n, m1, m2, chunk):
    ctx.set_option("qsm_chunk", chunk)
    qsmchecks.check_against_oracle_large(n, m1, m2, seed=n + m1)


@pytest.mark.parametrize("case", qsmcases.CONDITION, ids=lambda c: c["name"])
def test_gp_condition_at_the_inputs_matches_the_reference(case):
    """GaussianProcess.condition(y) with a quasiseparable kernel: SymmQSM-valued covariance, QuasisepSolver(covariance=)
    for the conditioned process; generators, factor and log-probability equal the reference's (qsm_vectors.json)"""
    t, y = qsmcases.condition_inputs(case)
    env = {"quasisep": quasisep, "np": np}
    k = eval(case["kernel"], env)
    kp = None if case["pred"] is None else eval(case["pred"], env)
    gp = GaussianProcess(k, t, diag=case["diag"])
    lp, cgp = gp.condition(y, diag=case["pdiag"], kernel=kp)
    g = GOLD["condition"][case["name"]]
    assert isinstance(cgp.solver, QuasisepSolver) and isinstance(cgp.solver.matrix, core.SymmQSM)
    m = cgp.solver.matrix
    if kp is None:      # own kernel: the order-J form N - N Sigma^-1 N of the same matrix
        assert m.lower.p.shape[1] == k.state_dim()
    else:               # another predictive kernel: the reference's algebra, generator for generator
        np.testing.assert_allclose(m.diag.d, g["d"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(m.lower.p, g["p"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(m.lower.q, g["q"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(m.lower.a, g["a"], rtol=1e-9, atol=1e-9)
    assert abs(lp - g["cond_log_probability"]) < 1e-9 * abs(g["cond_log_probability"])
    np.testing.assert_allclose(cgp.loc, g["loc"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(cgp.variance, g["variance"], rtol=1e-8, atol=1e-10)
    c, _ = cgp.solver.factor_arrays()
    np.testing.assert_allclose(c, g["factor_c"], rtol=1e-9, atol=1e-11)      # the Cholesky diagonal does not depend on the realisation
    assert abs(cgp.log_probability(y + 0.01) - g["cgp_log_probability"]) < 1e-7 * abs(g["cgp_log_probability"])
    np.testing.assert_allclose(cgp.covariance, g["dense"], rtol=1e-8, atol=1e-9)


def test_conditioning_a_long_series_never_densifies():
    """N = 200 000 (2048 chunks): variance from the QSM branch equals the O(N) inverse-diagonal scan of the model-based
    solver (an independent device path) and, on a 400-point prefix problem, the oracle's dense conditioning"""
    n = 200_000
    rng = np.random.default_rng(12)
    t = np.sort(rng.uniform(0, n / 10.0, n))
    y = np.sin(t) + 0.1 * rng.normal(size=n)
    k = quasisep.SHO(omega=1.5, quality=3.0, sigma=1.8) + quasisep.Matern32(scale=1.5, sigma=0.9)
    gp = GaussianProcess(k, t, diag=0.1)
    lp, cgp = gp.condition(y, diag=0.05)
    var_qsm = cgp.variance
    var_scan = gp.solver.conditioned_variance(cgp.noise)
    np.testing.assert_allclose(var_qsm, var_scan, rtol=1e-7, atol=1e-10)
    assert cgp.solver.info == 0 and np.isfinite(cgp.log_probability(y))
    # a different predictive kernel (one component of the sum): the reference's order-(J' + 2J' + J) algebra at N = 2e5
    k1 = quasisep.Matern32(scale=1.5, sigma=0.9)
    _, c1 = gp.condition(y, kernel=k1, diag=0.05)
    assert c1.solver.matrix.lower.p.shape[1] == 2 + (2 + 4 + 2) and c1.solver.info == 0
    assert np.all(c1.variance > 0) and np.all(c1.variance <= 0.81 + 0.05 + 1e-9)
    z = rng.normal(size=n)
    back = cgp.solver.solve_triangular(cgp.solver.dot_triangular(z))
    np.testing.assert_allclose(back, z, rtol=1e-6, atol=1e-8)
    m = 400                 # the oracle evaluates k(t_i, t_j) in Python loops: keep it small
    ko = o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9)
    _, cs = GaussianProcess(k, t[:m], diag=0.1).condition(y[:m], diag=0.05)
    _, co = o.GaussianProcess(ko, t[:m], diag=0.1).condition(y[:m], diag=0.05)
    np.testing.assert_allclose(cs.variance, co.variance, rtol=1e-7, atol=1e-10)
    assert abs(cs.log_probability(y[:m]) - co.log_probability(y[:m])) < 1e-7 * abs(co.log_probability(y[:m]))


# ---- the reference's own test_core.py (tests/qsm_reference_tests.py) on the GPU ---------------------------------------
import qsm_reference_tests as R  # noqa: E402


def test_reference_quasisep_def():
    R.check_quasisep_def()


@pytest.mark.parametrize("parallel", [False, True])
@pytest.mark.parametrize("name", ["random", "celerite"])
def test_reference_matmuls(name, parallel):
    R.check_strict_tri_matmul(name, parallel)
    R.check_tri_matmul(name, parallel)
    R.check_square_matmul(True, name, parallel)
    R.check_square_matmul(False, name, parallel)


@pytest.mark.parametrize("chunk", [0, 5])
def test_reference_inverses_solves_gram_cholesky(ctx, chunk):
    ctx.set_option("qsm_chunk", chunk)
    R.check_tri_inv()
    for parallel in (False, True):
        R.check_tri_solve(parallel)
        R.check_cholesky(parallel)
    for name in ("random", "celerite"):
        R.check_gram(name)
        for symm in (True, False):
            R.check_square_inv(symm, name)


@pytest.mark.parametrize("chunk", [0, 5])
def test_reference_products_and_ops(ctx, chunk):
    ctx.set_option("qsm_chunk", chunk)
    R.check_tri_qsmul()
    R.check_square_qsmul()
    R.check_ops()
