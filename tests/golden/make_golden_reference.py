"""Generate tests/golden/reference_vectors.json by running the UNMODIFIED reference (dfm/tinygp sources under
/root/reference/src) on the cases of refcases.py.  jax/equinox are not installable here, so the reference runs over
the NumPy stand-ins in tests/golden/jaxshim (see its README: only the array library is substituted; tinygp's own code
is executed as is).  Run from the repo root:  python tests/golden/make_golden_reference.py
"""

import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import refcases  # noqa: E402
import refimport  # noqa: E402


def reference_namespace():
    tinygp = refimport.install()
    from tinygp import GaussianProcess, kernels, noise, transforms
    from tinygp.kernels import quasisep

    def qs_factor(gp):
        f = gp.solver.factor                          # LowerTriQSM(diag=DiagQSM(c), lower=StrictLowerTriQSM(p, q=w, a))
        return np.asarray(f.diag.d), np.asarray(f.lower.q)

    assert tinygp.__file__.startswith(refimport.REFERENCE_SRC)
    return refcases.Namespace("reference", GaussianProcess, kernels, quasisep, transforms, qs_factor, noise=noise)


def main():
    ns = reference_namespace()
    out = {"generator": "tests/golden/make_golden_reference.py",
           "reference": "dfm/tinygp sources at /root/reference/src executed over tests/golden/jaxshim (NumPy %s)" % np.__version__,
           "cases": {}}
    try:
        out["reference_tree_sha256"] = subprocess.run(
            "cd %s && find tinygp -name '*.py' | sort | xargs sha256sum | sha256sum" % refimport.REFERENCE_SRC,
            shell=True, capture_output=True, text=True, check=True).stdout.split()[0]
    except Exception:
        pass
    for case in refcases.CASES:
        with np.errstate(all="ignore"):
            res = refcases.run_case(ns, case)
        out["cases"][case["name"]] = res
        print(f"{case['name']:32s} logp = {res['log_probability']!r}")
    # error behaviour: unsorted quasisep input (solver.py:142-146)
    import tinygp
    try:
        tinygp.GaussianProcess(ns.quasisep.Matern32(1.5), np.array([0.0, 2.0, 1.0]), diag=0.1)
        out["unsorted_raises"] = None
    except ValueError as e:
        out["unsorted_raises"] = str(e)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:      # one line per case
        f.write("{\n")
        for k in [k for k in out if k != "cases"]:
            f.write(f" {json.dumps(k)}: {json.dumps(out[k])},\n")
        f.write(' "cases": {\n')
        names = list(out["cases"])
        for i, n in enumerate(names):
            body = json.dumps(out["cases"][n], separators=(",", ":"))
            f.write(f"  {json.dumps(n)}: {body}{',' if i + 1 < len(names) else ''}\n")
        f.write(" }\n}\n")
    print("wrote reference_vectors.json")


if __name__ == "__main__":
    main()
