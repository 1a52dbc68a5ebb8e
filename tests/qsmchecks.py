"""The QSM-algebra checks, written once against tinygp_b200.solvers.quasisep.core and run (a) on the CPU over the host
build of the device source (tests/test_qsm_device_code_on_host.py) and (b) on the GPU over libb200gp.so
(tests/test_qsm_gpu.py).  Expected values: the unmodified reference's outputs (tests/golden/qsm_vectors.json) and the
oracle (oracle/qsm_np.py) for sizes beyond the goldens."""

import numpy as np

from oracle import qsm_np as oq
from qsmutil import GOLD, TYPE_TO_KIND, qsmcases
from tinygp_b200.solvers.quasisep import core

TOL = dict(rtol=1e-9, atol=1e-9)
KIND_NAME = {core.DIAG: "diag", core.STRICT_LOWER: "strict_lower", core.STRICT_UPPER: "strict_upper", core.LOWER: "lower",
             core.UPPER: "upper", core.SQUARE: "square", core.SYMM: "symm"}


def build(spec):
    k = spec["kind"]
    d = core.DiagQSM(spec["d"]) if "d" in spec else None
    lo = core.StrictLowerTriQSM(*spec["lower"]) if "lower" in spec else None
    up = core.StrictUpperTriQSM(*spec["upper"]) if "upper" in spec else None
    return {"diag": lambda: d, "strict_lower": lambda: lo, "strict_upper": lambda: up,
            "lower": lambda: core.LowerTriQSM(diag=d, lower=lo), "upper": lambda: core.UpperTriQSM(diag=d, upper=up),
            "square": lambda: core.SquareQSM(diag=d, lower=lo, upper=up), "symm": lambda: core.SymmQSM(diag=d, lower=lo)}[k]()


def operands():
    return {k: build(v) for k, v in qsmcases.operands().items()}


def kind_of(m):
    return KIND_NAME[{v: k for k, v in core._CLASS_OF_KIND.items()}[type(m)]]


def check_dense_and_parts(ops):
    specs = qsmcases.operands()
    for name, m in ops.items():
        assert kind_of(m) == specs[name]["kind"]
        np.testing.assert_allclose(m.to_dense(), GOLD["dense"][name], **TOL)
        np.testing.assert_allclose(m.T.to_dense(), np.asarray(GOLD["dense"][name]).T, **TOL)
        np.testing.assert_allclose((-m).to_dense(), -np.asarray(GOLD["dense"][name]), **TOL)
        np.testing.assert_allclose((m * 2.5).to_dense(), 2.5 * np.asarray(GOLD["dense"][name]), **TOL)
        x = np.asarray(GOLD["x"])
        np.testing.assert_allclose(m @ x, np.asarray(GOLD["dense"][name]) @ x, **TOL)
        np.testing.assert_allclose(m @ x[:, 0], np.asarray(GOLD["dense"][name]) @ x[:, 0], **TOL)
    # generator arrays round-trip through the device, parts share them
    sq, spec = ops["SQ"], specs["SQ"]
    np.testing.assert_array_equal(sq.diag.d, spec["d"])
    for got, want in zip(sq.lower, spec["lower"]):
        np.testing.assert_array_equal(got, want)
    for got, want in zip(sq.upper, spec["upper"]):
        np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(ops["SY"].lower.a, specs["SY"]["lower"][2])
    v = np.linspace(0.5, 1.5, qsmcases.N)
    np.testing.assert_allclose(sq.scale(v).to_dense(), v[:, None] * np.asarray(GOLD["dense"]["SQ"]), **TOL)


def check_products(ops):
    for key, g in GOLD["mul"].items():
        a, b = key.split("@")
        r = ops[a] @ ops[b]
        assert kind_of(r) == TYPE_TO_KIND[g["type"]], key
        np.testing.assert_allclose(r.to_dense(), g["dense"], err_msg=key, **TOL)
    for key in GOLD["mul_unsupported"]:          # the reference fails on these; the backend says so instead of guessing
        a, b = key.split("@")
        try:
            ops[a] @ ops[b]
        except Exception as e:
            assert "unequal widths" in str(e), key
        else:
            raise AssertionError(f"{key} should be refused")


def check_sums(ops):
    for key, g in GOLD["add"].items():
        a, b = key.split("+")
        r = ops[a] + ops[b]
        assert kind_of(r) == TYPE_TO_KIND[g["type"]], key
        np.testing.assert_allclose(r.to_dense(), g["dense"], err_msg=key, **TOL)
        np.testing.assert_allclose((ops[a] - ops[b]).to_dense(),
                                   np.asarray(GOLD["dense"][a]) - np.asarray(GOLD["dense"][b]), err_msg=key, **TOL)
    for key, g in GOLD["emul"].items():
        a, b = key.split("*")
        r = ops[a] * ops[b]
        assert kind_of(r) == TYPE_TO_KIND[g["type"]], key
        np.testing.assert_allclose(r.to_dense(), g["dense"], err_msg=key, **TOL)


def check_inverses_and_factor(ops):
    for key in ("L", "U", "SY", "SQ"):
        r = ops[key].inv()
        assert kind_of(r) == TYPE_TO_KIND[GOLD["inv"][key]["type"]]
        np.testing.assert_allclose(r.to_dense(), GOLD["inv"][key]["dense"], rtol=1e-8, atol=1e-8)
    for k in GOLD["gram"]:
        if hasattr(ops[k], "gram"):
            g = ops[k].gram()
            assert kind_of(g) == "symm"
            np.testing.assert_allclose(g.to_dense(), GOLD["gram"][k], **TOL)
    ch = ops["SY"].cholesky()
    assert ch.info == 0 and kind_of(ch) == "lower"
    np.testing.assert_allclose(ch.diag.d, GOLD["cholesky"]["SY"]["c"], **TOL)
    np.testing.assert_allclose(ch.lower.q, GOLD["cholesky"]["SY"]["w"], **TOL)
    x = np.asarray(GOLD["x"])
    np.testing.assert_allclose(ops["L"].solve(x), GOLD["solve"]["L"], **TOL)
    np.testing.assert_allclose(ops["U"].solve(x), GOLD["solve"]["U"], **TOL)
    np.testing.assert_allclose(ch.solve(x), GOLD["solve"]["chol"], **TOL)
    np.testing.assert_allclose(ch.T.solve(x), GOLD["solve"]["cholT"], **TOL)
    np.testing.assert_allclose(ch.solve(x[:, 1]), np.asarray(GOLD["solve"]["chol"])[:, 1], **TOL)
    # a non-positive pivot: NaNs from there on and the index of the first one (the reference only has the NaNs)
    spec = qsmcases.operands()["SY"]
    bad = core.SymmQSM(diag=core.DiagQSM(np.where(np.arange(qsmcases.N) == 5, -1.0, spec["d"])),
                       lower=core.StrictLowerTriQSM(*spec["lower"])).cholesky()
    assert bad.info == 6 and np.isnan(bad.diag.d[5])


def check_condition_algebra(case, from_kernel=None):
    """solver.py:124-129 spelled with the public classes: generators equal the reference's"""
    from oracle import tinygp_np as o
    t, y = qsmcases.condition_inputs(case)
    env = {"quasisep": o.qs, "np": np}
    k = eval(case["kernel"], env)
    kp = k if case["pred"] is None else eval(case["pred"], env)
    d, p, q, a = o.qs_generators_fast(k, t)
    K = core.SymmQSM(diag=core.DiagQSM(d + case["diag"]), lower=core.StrictLowerTriQSM(p, q, a))
    factor = K.cholesky()
    dm, pm, qm, am = o.qs_generators_fast(kp, t)
    M = core.SymmQSM(diag=core.DiagQSM(dm), lower=core.StrictLowerTriQSM(pm, qm, am))
    delta = (factor.inv() @ M).gram()
    M = M + core.DiagQSM(np.full(case["n"], case["pdiag"]))
    cov = M - delta
    g = GOLD["condition"][case["name"]]
    assert kind_of(cov) == "symm"
    np.testing.assert_allclose(cov.diag.d, g["d"], **TOL)
    lo = cov.lower
    np.testing.assert_allclose(lo.p, g["p"], **TOL)
    np.testing.assert_allclose(lo.q, g["q"], **TOL)
    np.testing.assert_allclose(lo.a, g["a"], **TOL)
    np.testing.assert_allclose(cov.to_dense(), g["dense"], **TOL)
    ch = cov.cholesky()
    np.testing.assert_allclose(ch.diag.d, g["factor_c"], rtol=1e-7, atol=1e-9)
    assert abs(np.linalg.norm(ch.lower.q) - g["factor_w_norm"]) < 1e-7 * g["factor_w_norm"]


def check_against_oracle_large(n, m1, m2, seed):
    """sizes with many chunks: every scan against the oracle's point-by-point loops"""
    p1, q1, a1 = qsmcases.generators(seed, m1, n=n)
    p2, q2, a2 = qsmcases.generators(seed + 1, m2, n=n)
    d1, d2 = qsmcases.diag(seed + 2, n=n) + 3.0, qsmcases.diag(seed + 3, n=n) + 3.0
    A = core.SquareQSM(diag=core.DiagQSM(d1), lower=core.StrictLowerTriQSM(p1, q1, a1), upper=core.StrictUpperTriQSM(p2, q2, a2))
    B = core.SquareQSM(diag=core.DiagQSM(d2), lower=core.StrictLowerTriQSM(p2, q2, a2), upper=core.StrictUpperTriQSM(p1, q1, a1))
    Ao, Bo = oq.QSM(d1, (p1, q1, a1), (p2, q2, a2)), oq.QSM(d2, (p2, q2, a2), (p1, q1, a1))
    x = np.random.default_rng(seed).normal(size=(n, 19))
    np.testing.assert_allclose(A @ x, Ao.matmul(x), **TOL)
    C, Co = A @ B, oq.qsm_mul(Ao, Bo)
    for got, want in zip(C.lower, Co.lower):
        np.testing.assert_allclose(got, want, **TOL)
    for got, want in zip(C.upper, Co.upper):
        np.testing.assert_allclose(got, want, **TOL)
    np.testing.assert_allclose(C.diag.d, Co.d, **TOL)
    L, Lo = core.LowerTriQSM(diag=core.DiagQSM(d1), lower=core.StrictLowerTriQSM(p1, q1, a1)), oq.QSM(d1, (p1, q1, a1))
    np.testing.assert_allclose(L.solve(x), Lo.solve(x), **TOL)
    np.testing.assert_allclose(L.T.solve(x), Lo.transpose().solve(x), **TOL)
    Li = L.inv()
    Sq = Li.T @ Li                                              # SPD by construction: (L L^T)^-1
    G, Go = core.SymmQSM(diag=Sq.diag, lower=Sq.lower), Lo.inv().gram()
    np.testing.assert_allclose(G.diag.d, Go.d, **TOL)
    ch, (co, wo) = G.cholesky(), oq.cholesky(Go.d, *Go.lower)
    assert ch.info == 0
    np.testing.assert_allclose(ch.diag.d, co, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(ch.lower.q, wo, rtol=1e-7, atol=1e-8)
    Gi, (lam, t, s, ell) = G.inv(), oq.symm_inv(Go.d, *Go.lower)
    np.testing.assert_allclose(Gi.diag.d, lam, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(Gi.lower.p, t, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(Gi.lower.q, s, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(Gi.lower.a, ell, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(core.sum_log_diag(ch), np.sum(np.log(co)), rtol=1e-12)
