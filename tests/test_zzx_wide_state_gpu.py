"""Quasiseparable models with 7 and 8 states (the widest the backend compiles, B200GP_QS_MAX_J) and Products of Sums, which
the host lowering multiplies out into several Kronecker-structured terms (tinygp_b200/kernels/quasisep.py Product.components;
reference kernels/quasisep.py:298-331).  The oracle keeps the reference's interleaved state order, so agreement here also
shows that the block-wise permutation of the state changes no solver result."""

import numpy as np
import pytest

import oracle.tinygp_np as o
from tinygp_b200 import GaussianProcess
from tinygp_b200.kernels import quasisep as Q
from util import LOGP_RTOL, rel, to_oracle

pytestmark = pytest.mark.gpu

MODELS = {
    "m52*cosine+exp (7)": Q.Matern52(2.5, 1.3) * Q.Cosine(3.0, 0.7) + Q.Exp(2.0, 0.5),
    "m52+m52+sho (8)": Q.Matern52(2.5, 1.3) + Q.Matern52(0.6, 0.4) + Q.SHO(1.5, 3.0, 0.8),
    "(m52+0.4exp)*sho (8)": (Q.Matern52(1.5) + 0.4 * Q.Exp(0.7)) * Q.SHO(omega=1.5, quality=0.1),
    "(exp+exp)*(cosine+0.5exp) (6)": (Q.Exp(1.5) + Q.Exp(0.7)) * (Q.Cosine(2.5) + 0.5 * Q.Exp(1.1)),
    "carma21*exp (2)": Q.CARMA(alpha=np.array([1, 1.2]), beta=np.array([1.0, 3.0])) * Q.Exp(2.0),
}


def _data(n, seed=11):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, n / 6.0, n))
    return t, np.sin(t) + 0.1 * rng.normal(size=n), rng.uniform(0.05, 0.2, n)


@pytest.mark.parametrize("n", [1, 40, 3000])
@pytest.mark.parametrize("name", list(MODELS))
def test_log_probability(name, n):
    t, y, noise = _data(n)
    got = GaussianProcess(MODELS[name], t, diag=noise).log_probability(y)
    want = o.GaussianProcess(to_oracle(MODELS[name]), t, diag=noise).log_probability(y)
    assert rel(got, want) < LOGP_RTOL


@pytest.mark.parametrize("name", list(MODELS))
def test_condition_predict_and_inverse_diagonal(name):
    t, y, noise = _data(140)      # the oracle evaluates k(t, t) point by point in Python: keep it small
    xs = np.sort(np.random.default_rng(5).uniform(-2.0, t[-1] + 2.0, 23))
    gp, gpo = GaussianProcess(MODELS[name], t, diag=noise), o.GaussianProcess(to_oracle(MODELS[name]), t, diag=noise)
    lp, cond = gp.condition(y, xs)
    lpo, condo = gpo.condition(y, xs)
    assert rel(lp, lpo) < LOGP_RTOL
    np.testing.assert_allclose(cond.loc, condo.loc, rtol=5e-7, atol=5e-7)
    np.testing.assert_allclose(cond.variance, condo.variance, rtol=5e-7, atol=5e-7)
    mu, var = gp.predict(y, return_var=True)
    muo, varo = gpo.predict(y, return_var=True)
    np.testing.assert_allclose(mu, muo, rtol=5e-7, atol=5e-7)
    np.testing.assert_allclose(var, varo, rtol=1e-7, atol=1e-10)
    want = np.diag(np.linalg.inv(gpo.solver.covariance()))
    np.testing.assert_allclose(gp.solver.inverse_diagonal(), want, rtol=1e-8, atol=0)
