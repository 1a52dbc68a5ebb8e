"""Generate tests/golden/qsm_vectors.json: the reference's quasiseparable-matrix algebra (solvers/quasisep/core.py,
ops.py) and its QSM-valued conditioning (solvers/quasisep/solver.py:124-129), executed from the UNMODIFIED sources under
/root/reference/src over the NumPy stand-ins of tests/golden/jaxshim.  Run from the repo root:
    python tests/golden/make_golden_qsm.py
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import qsmcases  # noqa: E402
import refimport  # noqa: E402


def main():
    tinygp = refimport.install()
    from tinygp import GaussianProcess, noise
    from tinygp.kernels import quasisep
    from tinygp.solvers.quasisep import core

    def build(spec):
        k = spec["kind"]
        lo = core.StrictLowerTriQSM(*spec["lower"]) if "lower" in spec else None
        up = core.StrictUpperTriQSM(*spec["upper"]) if "upper" in spec else None
        d = core.DiagQSM(spec["d"]) if "d" in spec else None
        return {"diag": lambda: d, "strict_lower": lambda: lo, "strict_upper": lambda: up,
                "lower": lambda: core.LowerTriQSM(diag=d, lower=lo), "upper": lambda: core.UpperTriQSM(diag=d, upper=up),
                "square": lambda: core.SquareQSM(diag=d, lower=lo, upper=up),
                "symm": lambda: core.SymmQSM(diag=d, lower=lo)}[k]()

    def dense(m):
        return np.asarray(m.to_dense()).tolist()

    ops = qsmcases.operands()
    objs = {k: build(v) for k, v in ops.items()}
    out = {"generator": "tests/golden/make_golden_qsm.py",
           "reference": "dfm/tinygp sources at /root/reference/src executed over tests/golden/jaxshim", "dense": {},
           "mul": {}, "mul_unsupported": [], "add": {}, "emul": {}, "inv": {}, "gram": {}, "cholesky": {}, "solve": {}, "condition": {}}
    rng = np.random.default_rng(99)
    x = rng.normal(size=(qsmcases.N, 3))
    out["x"] = x.tolist()
    for k, m in objs.items():
        out["dense"][k] = dense(m)
    names = list(objs)
    for a in names:
        for b in names:
            if not (a == "D" and b == "D"):     # the reference's diag @ diag wraps a DiagQSM in a DiagQSM (ops.py:57-60)
                try:
                    r = objs[a] @ objs[b]                             # ops.py:52-214
                    rd = dense(r)
                except Exception as e:     # e.g. strict x strict: the reference builds generators of unequal widths
                    out["mul_unsupported"].append(f"{a}@{b}")
                    continue
                want = np.asarray(out["dense"][a]) @ np.asarray(out["dense"][b])
                # symm x symm is declared symmetric by ops.py:212-214 (its upper part is dropped): recorded as it is
                out["mul"][f"{a}@{b}"] = {"type": type(r).__name__, "dense": rd,
                                          "is_the_dense_product": bool(np.allclose(rd, want, atol=1e-9))}
            try:
                r = objs[a] + objs[b]                                 # ops.py:24-35
                if r is not None:
                    out["add"][f"{a}+{b}"] = {"type": type(r).__name__, "dense": dense(r)}
            except Exception:
                pass
            try:
                r = objs[a] * objs[b]                                 # ops.py:38-49
                if r is not None:
                    out["emul"][f"{a}*{b}"] = {"type": type(r).__name__, "dense": dense(r)}
            except Exception:
                pass
    for a in qsmcases.INVERTIBLE:
        r = objs[a].inv()
        out["inv"][a] = {"type": type(r).__name__, "dense": dense(r)}
    out["inv"]["SY_parallel"] = {"type": "SymmQSM", "dense": dense(objs["SY"].inv(parallel=True))}
    for a in ("SQ", "L", "U"):
        if hasattr(objs[a], "gram"):
            out["gram"][a] = dense(objs[a].gram())
    chol = objs["SY"].cholesky()
    out["cholesky"]["SY"] = {"c": np.asarray(chol.diag.d).tolist(), "w": np.asarray(chol.lower.q).tolist()}
    out["solve"]["L"] = np.asarray(objs["L"].solve(x)).tolist()
    out["solve"]["U"] = np.asarray(objs["U"].solve(x)).tolist()
    out["solve"]["chol"] = np.asarray(chol.solve(x)).tolist()
    out["solve"]["cholT"] = np.asarray(chol.transpose().solve(x)).tolist()

    env = {"quasisep": quasisep, "np": np}
    for case in qsmcases.CONDITION:
        t, y = qsmcases.condition_inputs(case)
        k = eval(case["kernel"], env)
        kp = k if case["pred"] is None else eval(case["pred"], env)
        gp = GaussianProcess(k, t, diag=case["diag"])
        cov = gp.solver.condition(kp, None, noise.Diagonal(np.full(case["n"], case["pdiag"])))   # solver.py:124-129
        assert type(cov).__name__ == "SymmQSM"
        rec = {"d": np.asarray(cov.diag.d).tolist(), "p": np.asarray(cov.lower.p).tolist(),
               "q": np.asarray(cov.lower.q).tolist(), "a": np.asarray(cov.lower.a).tolist(), "dense": dense(cov)}
        cond = gp.condition(y, diag=case["pdiag"], kernel=None if case["pred"] is None else kp)
        cgp = cond.gp
        rec["cond_log_probability"] = float(cond.log_probability)
        rec["loc"] = np.asarray(cgp.loc).tolist()
        rec["variance"] = np.asarray(cgp.variance).tolist()
        rec["cgp_log_probability"] = float(cgp.log_probability(y + 0.01))
        f = cgp.solver.factor
        rec["factor_c"] = np.asarray(f.diag.d).tolist()
        rec["factor_w_norm"] = float(np.linalg.norm(np.asarray(f.lower.q)))
        out["condition"][case["name"]] = rec
        print(case["name"], "order", np.asarray(cov.lower.p).shape[1], "logp", rec["cgp_log_probability"])
    with open(os.path.join(HERE, "qsm_vectors.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote qsm_vectors.json", os.path.getsize(os.path.join(HERE, "qsm_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
