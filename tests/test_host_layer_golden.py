"""The Python host layer of tinygp_b200 on a CPU-only machine: the product's GaussianProcess / kernels / solvers /
transforms run unchanged, with the C-ABI library replaced by tests/hostmock.py (the oracle behind the same entry
points), and must reproduce the reference-generated goldens.  This checks what the GPU parity tests cannot isolate:
lowering to kernel programs, argument marshalling, which entry point is called with what, noise placement and the
error behaviour of the host code -- not the CUDA kernels (those are the `-m gpu` tests' job)."""

import os
import sys
from ctypes import c_void_p

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import hostmock  # noqa: E402
import refcases  # noqa: E402
from test_reference_golden import GOLD, compare, product_namespace  # noqa: E402

from tinygp_b200 import _cabi  # noqa: E402


@pytest.fixture()
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


@pytest.mark.parametrize("case", refcases.CASES, ids=[c["name"] for c in refcases.CASES])
def test_host_layer_reproduces_reference_goldens(mocklib, case):
    with np.errstate(all="ignore"):
        got = refcases.run_case(product_namespace(), case)
    compare(got, GOLD["cases"][case["name"]], case["name"], tol=1e-9)
    if not np.isfinite(got["log_probability"]):
        return                                  # non-PD: the reference says -inf and nothing else is defined
    if "noise" in case:
        # noise.Banded / noise.Dense: a factor of a precomputed covariance; Kss - A^T A is the device GEMM entry point
        assert "gram_downdate" in mocklib.calls
        assert ("qsm_cholesky" if case["kind"] == "quasisep" else "dense_create_from_cov") in mocklib.calls
    elif case["kind"] == "quasisep":
        # conditioning goes through the device entry point, never through host linear algebra
        assert "qs_condition" in mocklib.calls and "qs_kernel_matmul" in mocklib.calls
    else:
        assert "dense_condition" in mocklib.calls


def test_unsorted_raises_reference_text(mocklib):
    import tinygp_b200 as tg
    from tinygp_b200.kernels import quasisep
    with pytest.raises(ValueError) as e:
        tg.GaussianProcess(quasisep.Matern32(1.5), np.array([0.0, 2.0, 1.0]), diag=0.1)
    assert str(e.value) == GOLD["unsorted_raises"]
    tg.GaussianProcess(quasisep.Matern32(1.5), np.array([0.0, 2.0, 1.0]), diag=0.1, assume_sorted=True)


def test_condition_shape_error_like_the_reference(mocklib):
    """tests/test_gp.py:52-76 of the reference (array branch)"""
    import tinygp_b200 as tg
    rng = np.random.default_rng(0)
    X, y = rng.uniform(-3, 3, (20, 2)), rng.normal(size=20)
    gp = tg.GaussianProcess(tg.kernels.ExpSquared(distance=tg.kernels.L2Distance()), X, diag=0.1)
    gp.condition(y, X[0][None])
    with pytest.raises(ValueError):
        gp.condition(y, X[0])


def test_matrix_rhs_kernel_matmul_runs_on_the_device_entry_point(mocklib):
    import tinygp_b200 as tg
    rng = np.random.default_rng(1)
    X1, X2, Y = rng.uniform(0, 3, (5, 2)), rng.uniform(0, 3, (7, 2)), rng.normal(size=(7, 3))
    k = 1.3 * tg.kernels.Matern52(0.8)
    got = k.matmul(X1, X2, Y)
    assert got.shape == (5, 3) and mocklib.calls.count("kernel_matvec") == 3 and "kernel_matrix" not in mocklib.calls
    from oracle import tinygp_np as o
    np.testing.assert_allclose(got, (1.3 * o.Matern52(0.8))(X1, X2) @ Y, rtol=1e-12, atol=1e-14)


def test_mean_only_predict_skips_the_conditioned_covariance(mocklib):
    """gp.py:225-229: predict is jitted with return_var / return_cov static, so the reference never materialises the
    conditioned covariance for a mean-only prediction; neither do we (an N x N matrix for a long time series)."""
    import tinygp_b200 as tg
    from tinygp_b200.kernels import quasisep
    from oracle import tinygp_np as o
    rng = np.random.default_rng(2)
    t = np.sort(rng.uniform(0, 30, 150)); y = np.sin(t); tt = rng.uniform(-1, 31, 11)
    k, ko = quasisep.SHO(1.5, 3.0, 1.8) + quasisep.Matern32(1.5, 0.9), o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9)
    gp, gpo = tg.GaussianProcess(k, t, diag=0.1), o.GaussianProcess(ko, t, diag=0.1)
    np.testing.assert_allclose(gp.predict(y), gpo.predict(y), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gp.predict(y, tt), gpo.predict(y, tt), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gp.predict(y, tt, include_mean=False), gpo.predict(y, tt, include_mean=False),
                               rtol=1e-10, atol=1e-12)
    assert "qs_condition" not in mocklib.calls and "kernel_matrix" not in mocklib.calls
    mu, var = gp.predict(y, tt, return_var=True)
    muo, varo = gpo.predict(y, tt, return_var=True)
    np.testing.assert_allclose(var, varo, rtol=1e-9, atol=1e-12)
    assert "qs_condition" in mocklib.calls
    Xd = rng.uniform(0, 4, (40, 2)); yd = np.sin(Xd[:, 0])
    gpd = tg.GaussianProcess(tg.kernels.Matern32(1.2), Xd, diag=0.1, mean=0.4)
    gpdo = o.GaussianProcess(o.Matern32(1.2), Xd, diag=0.1, mean=0.4)
    n0 = len(mocklib.calls)
    np.testing.assert_allclose(gpd.predict(yd, Xd[:5] + 0.1), gpdo.predict(yd, Xd[:5] + 0.1), rtol=1e-10, atol=1e-12)
    assert "dense_condition" not in mocklib.calls[n0:]
    with pytest.raises(ValueError):
        gpd.predict(yd, Xd[0])


def test_conditioning_a_quasisep_process_at_its_inputs_is_qsm_valued(mocklib):
    """solver.py:124-129: the conditioned covariance at the inputs is a SymmQSM, factored by QuasisepSolver(covariance=);
    neither the N x N conditioning entry point nor a dense factorisation is called for anything but `.covariance`"""
    import tinygp_b200 as tg
    from tinygp_b200.kernels import quasisep
    from tinygp_b200.solvers import QuasisepSolver
    from tinygp_b200.solvers.quasisep.core import SymmQSM
    from oracle import tinygp_np as o
    rng = np.random.default_rng(6)
    t = np.sort(rng.uniform(0, 25, 120)); y = np.sin(t) + 0.1 * rng.normal(size=120)
    k, ko = quasisep.Matern52(2.5, 1.3) + quasisep.Exp(1.1, 0.4), o.qs.Matern52(2.5, 1.3) + o.qs.Exp(1.1, 0.4)
    gp = tg.GaussianProcess(k, t, diag=0.08)
    lp, cond = gp.condition(y, diag=0.02)
    lpo, condo = o.GaussianProcess(ko, t, diag=0.08).condition(y, diag=0.02)
    assert isinstance(cond.solver, QuasisepSolver) and isinstance(cond.solver.matrix, SymmQSM)
    assert cond.solver.matrix.lower.p.shape == (120, 4)        # own kernel: N - N Sigma^-1 N, order J (the reference: 4J)
    np.testing.assert_allclose(lp, lpo, rtol=1e-10)
    np.testing.assert_allclose(cond.loc, condo.loc, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(cond.variance, condo.variance, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(cond.log_probability(y + 0.01), condo.log_probability(y + 0.01), rtol=1e-8)
    z = rng.normal(size=(120, 3))
    np.testing.assert_allclose(cond.solver.dot_triangular(z), np.linalg.cholesky(condo.covariance) @ z, rtol=1e-7, atol=1e-9)
    assert "qs_condition" not in mocklib.calls and "dense_create_from_cov" not in mocklib.calls
    assert {"qs_kernel_qsm", "qsm_inv", "qsm_scale", "qsm_add", "qsm_cholesky"} <= set(mocklib.calls)
    np.testing.assert_allclose(cond.covariance, condo.covariance, rtol=1e-8, atol=1e-11)
    # predicting one component of a sum (kernel=): the QSM branch with a different predictive kernel
    k1, k1o = quasisep.Exp(1.1, 0.4), o.qs.Exp(1.1, 0.4)
    n0 = len(mocklib.calls)
    _, c1 = gp.condition(y, kernel=k1, diag=0.02)
    assert {"qs_factor_qsm", "qsm_mul", "qsm_gram"} <= set(mocklib.calls[n0:]) and c1.solver.matrix.lower.p.shape == (120, 7)
    _, c1o = o.GaussianProcess(ko, t, diag=0.08).condition(y, kernel=k1o, diag=0.02)
    np.testing.assert_allclose(c1.loc, c1o.loc, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(c1.variance, c1o.variance, rtol=1e-8, atol=1e-12)
