"""Inputs of the QSM-algebra golden vectors (tests/golden/qsm_vectors.json): seeded NumPy generators and the list of
operations.  Imported by the generator (make_golden_qsm.py, which runs the UNMODIFIED reference over jaxshim) and by
the tests (which rebuild the same inputs for the oracle and for the CUDA path).  Test infrastructure."""

import numpy as np

N = 20


def generators(seed, m, n=N, contract=0.6):
    """random order-m generators with contracting transition matrices (products of many a's stay bounded)"""
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, m))
    q = rng.normal(size=(n, m))
    a = contract * rng.normal(size=(n, m, m)) / np.sqrt(m)
    return p, q, a


def diag(seed, n=N):
    return np.exp(0.3 * np.random.default_rng(seed).normal(size=n)) + 2.0


def celerite_symm(seed, n=N):
    """the positive-definite celerite SymmQSM of the reference's own test fixture (tests/.../test_core.py:61-110)"""
    rng = np.random.default_rng(seed)
    d0 = np.exp(rng.normal(size=n))
    t = np.sort(rng.uniform(0, 10, n))
    a = np.array([1.0, 2.5]); b = np.array([0.5, 1.5]); c = np.array([1.2, 0.5]); d = np.array([0.5, 0.1])
    cos, sin = np.cos(d[None] * t[:, None]), np.sin(d[None] * t[:, None])
    p = np.concatenate((a[None] * cos + b[None] * sin, a[None] * sin - b[None] * cos), axis=1)
    q = np.concatenate((cos, sin), axis=1)
    cc = np.append(c, c)
    dt = np.append(0, np.diff(t))
    am = np.stack([np.diag(v) for v in np.exp(-cc[None] * dt[:, None])], axis=0)
    p = np.einsum("ni,nij->nj", p, am)
    dd = d0 + np.sum(a)
    return dd, p, q, am


# operand table: name -> (kind, builder args); kinds follow core.py's seven classes
def operands():
    p1, q1, a1 = generators(1, 2)
    p2, q2, a2 = generators(2, 3)
    p3, q3, a3 = generators(3, 2)
    dd, pc, qc, ac = celerite_symm(4)
    return {
        "D": dict(kind="diag", d=diag(10)),
        "SL": dict(kind="strict_lower", lower=(p1, q1, a1)),
        "SU": dict(kind="strict_upper", upper=(p2, q2, a2)),
        "L": dict(kind="lower", d=diag(11), lower=(p2, q2, a2)),
        "U": dict(kind="upper", d=diag(12), upper=(p3, q3, a3)),
        "SQ": dict(kind="square", d=diag(13) + 3.0, lower=(p1, q1, a1), upper=(p2, q2, a2)),
        "SY": dict(kind="symm", d=dd, lower=(pc, qc, ac)),
    }


UNARY = ["transpose", "neg", "to_dense", "scale"]
INVERTIBLE = ["L", "U", "SY", "SQ"]

# conditioning cases: training kernel, predictive kernel (None = the same), n, seed
CONDITION = [
    dict(name="cond_c4", kernel="quasisep.SHO(omega=1.5, quality=3.0, sigma=1.8) + quasisep.Matern32(scale=1.5, sigma=0.9)",
         pred=None, n=24, span=12.0, diag=0.1, pdiag=0.05, seed=501),
    dict(name="cond_m52", kernel="quasisep.Matern52(scale=2.5, sigma=1.3)", pred=None, n=20, span=12.0, diag=0.05,
         pdiag=0.02, seed=502),
    dict(name="cond_other_kernel", kernel="quasisep.Celerite(1.1, 0.1, 0.3, 1.5) + quasisep.Exp(scale=2.0, sigma=0.5)",
         pred="quasisep.Exp(scale=2.0, sigma=0.5)", n=24, span=12.0, diag=0.05, pdiag=0.02, seed=503),
]


def condition_inputs(case):
    rng = np.random.default_rng(case["seed"])
    t = np.sort(rng.uniform(0, case["span"], case["n"]))
    y = np.sin(t) + 0.1 * rng.normal(size=case["n"])
    return t, y
