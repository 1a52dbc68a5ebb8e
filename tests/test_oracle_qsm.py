"""The oracle's QSM algebra (oracle/qsm_np.py) pinned against the unmodified reference's outputs
(tests/golden/qsm_vectors.json <- tests/golden/make_golden_qsm.py: core.py / ops.py / solver.py:124-129 over jaxshim)."""

import numpy as np
import pytest

from oracle import qsm_np as oq
from oracle import tinygp_np as o
from qsmutil import GOLD, TYPE_TO_KIND, oracle_operands, qsmcases

OPS = oracle_operands()
TOL = dict(rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", list(GOLD["dense"]))
def test_dense_form(name):
    np.testing.assert_allclose(OPS[name].to_dense(), GOLD["dense"][name], **TOL)
    np.testing.assert_allclose(OPS[name].transpose().to_dense(), np.asarray(GOLD["dense"][name]).T, **TOL)


@pytest.mark.parametrize("key", list(GOLD["mul"]))
def test_qsm_mul(key):
    a, b = key.split("@")
    r = OPS[a] @ OPS[b]
    assert r.kind == TYPE_TO_KIND[GOLD["mul"][key]["type"]]
    np.testing.assert_allclose(r.to_dense(), GOLD["mul"][key]["dense"], **TOL)


@pytest.mark.parametrize("key", list(GOLD["add"]))
def test_elementwise_add(key):
    a, b = key.split("+")
    r = OPS[a] + OPS[b]
    assert r.kind == TYPE_TO_KIND[GOLD["add"][key]["type"]]
    np.testing.assert_allclose(r.to_dense(), GOLD["add"][key]["dense"], **TOL)
    np.testing.assert_allclose((OPS[a] - OPS[b]).to_dense(), OPS[a].to_dense() - OPS[b].to_dense(), **TOL)


@pytest.mark.parametrize("key", list(GOLD["emul"]))
def test_elementwise_mul(key):
    a, b = key.split("*")
    r = OPS[a] * OPS[b]
    assert r.kind == TYPE_TO_KIND[GOLD["emul"][key]["type"]]
    np.testing.assert_allclose(r.to_dense(), GOLD["emul"][key]["dense"], **TOL)


@pytest.mark.parametrize("key", ["L", "U", "SY", "SQ"])
def test_inverse(key):
    r = OPS[key].inv()
    assert r.kind == TYPE_TO_KIND[GOLD["inv"][key]["type"]]
    np.testing.assert_allclose(r.to_dense(), GOLD["inv"][key]["dense"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(r.to_dense() @ OPS[key].to_dense(), np.eye(qsmcases.N), atol=1e-8)


def test_gram_cholesky_solve():
    for k in GOLD["gram"]:
        np.testing.assert_allclose(OPS[k].gram().to_dense(), GOLD["gram"][k], **TOL)
    ch = OPS["SY"].cholesky()
    np.testing.assert_allclose(ch.d, GOLD["cholesky"]["SY"]["c"], **TOL)
    np.testing.assert_allclose(ch.lower[1], GOLD["cholesky"]["SY"]["w"], **TOL)
    x = np.asarray(GOLD["x"])
    np.testing.assert_allclose(OPS["L"].solve(x), GOLD["solve"]["L"], **TOL)
    np.testing.assert_allclose(OPS["U"].solve(x), GOLD["solve"]["U"], **TOL)
    np.testing.assert_allclose(ch.solve(x), GOLD["solve"]["chol"], **TOL)
    np.testing.assert_allclose(ch.transpose().solve(x), GOLD["solve"]["cholT"], **TOL)


@pytest.mark.parametrize("case", qsmcases.CONDITION, ids=lambda c: c["name"])
def test_condition_qsm(case):
    """solver.py:124-129 through the oracle's algebra: generators AND dense values equal the reference's"""
    t, y = qsmcases.condition_inputs(case)
    env = {"quasisep": o.qs, "np": np}
    k = eval(case["kernel"], env)
    kp = k if case["pred"] is None else eval(case["pred"], env)
    d, p, q, a = o.qs_generators_fast(k, t)
    c, w = oq.cholesky(d + case["diag"], p, q, a)
    factor = oq.QSM(c, (p, w, a))
    dm, pm, qm, am = o.qs_generators_fast(kp, t)
    cov = oq.condition_qsm(factor, oq.QSM(dm, (pm, qm, am), symm=True), np.full(case["n"], case["pdiag"]))
    g = GOLD["condition"][case["name"]]
    assert cov.kind == "symm"
    np.testing.assert_allclose(cov.d, g["d"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cov.lower[0], g["p"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cov.lower[1], g["q"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cov.lower[2], g["a"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cov.to_dense(), g["dense"], rtol=1e-9, atol=1e-9)
    ch = cov.cholesky()
    np.testing.assert_allclose(ch.d, g["factor_c"], rtol=1e-7, atol=1e-9)
