"""Generate tests/golden/known_answers.json.

The reference (tinygp on JAX) cannot be imported in this container (no jax / equinox, no network), so these are
NOT reference-produced vectors.  They are known answers from an INDEPENDENT formulation that shares no code with
oracle/ or tinygp_b200/: covariance entries from scalar Python formulas written out from the kernel definitions
(src/tinygp/kernels/stationary.py:76-235, kernels/quasisep.py docstrings: the closed-form k(tau)), and the
log-density from scipy.stats.multivariate_normal.  Run:  python tests/golden/make_golden.py
"""
import json
import math
import os

import numpy as np
import scipy.stats


def k_expsq(x1, x2, scale):
    return math.exp(-0.5 * sum((a - b) ** 2 for a, b in zip(x1, x2)) / scale ** 2)


def k_m32_l2(x1, x2, scale):
    r = math.sqrt(sum((a - b) ** 2 for a, b in zip(x1, x2))) / scale
    return (1 + math.sqrt(3) * r) * math.exp(-math.sqrt(3) * r)


def k_m52_l1(x1, x2, scale):
    r = sum(abs(a - b) for a, b in zip(x1, x2)) / scale
    return (1 + math.sqrt(5) * r + 5 * r * r / 3) * math.exp(-math.sqrt(5) * r)


def k_rq_l1(x1, x2, scale, alpha):
    r2 = (sum(abs(a - b) for a, b in zip(x1, x2)) / scale) ** 2
    return (1 + 0.5 * r2 / alpha) ** (-alpha)


def k_sho(tau, w, q, sigma):  # kernels/quasisep.py:404-430 docstring, Q > 1/2
    g = math.sqrt(4 * q * q - 1)
    arg = g * w * tau / (2 * q)
    return sigma ** 2 * math.exp(-w * tau / (2 * q)) * (math.cos(arg) + math.sin(arg) / g)


def k_qm32(tau, scale, sigma):  # kernels/quasisep.py:528-545 docstring
    f = math.sqrt(3) / scale
    return sigma ** 2 * (1 + f * tau) * math.exp(-f * tau)


def main():
    out = {"note": "independent SciPy/Python known answers; see make_golden.py", "cases": []}
    rng = np.random.default_rng(20260923)
    # dense cases
    for name, nd, n in [("expsq3d", 3, 40), ("m32_l2_2d", 2, 35), ("m52_l1_1d", 1, 30), ("combo_1d", 1, 25)]:
        X = rng.uniform(-2, 2, (n, nd))
        y = np.sin(X[:, 0]) + 0.05 * rng.normal(size=n)
        diag = 0.1
        K = np.empty((n, n))
        for i in range(n):
            for j in range(n):
                if name == "expsq3d":
                    K[i, j] = 1.7 * k_expsq(X[i], X[j], 0.9)
                elif name == "m32_l2_2d":
                    K[i, j] = k_m32_l2(X[i], X[j], 1.3)
                elif name == "m52_l1_1d":
                    K[i, j] = 0.8 * k_m52_l1(X[i], X[j], 1.1)
                else:
                    K[i, j] = 1.2 * k_expsq(X[i], X[j], 0.7) + k_rq_l1(X[i], X[j], 1.5, 1.5) * 0.5 + 0.05
        K += diag * np.eye(n)
        lp = float(scipy.stats.multivariate_normal(np.zeros(n), K).logpdf(y))
        out["cases"].append({"kind": "dense", "name": name, "X": X.tolist(), "y": y.tolist(), "diag": diag, "logp": lp,
                             "K00": float(K[0, 0]), "K01": float(K[0, 1])})
    # quasisep case: SHO + Matern32 (BASELINE config 4 kernel) via its closed form
    n = 45
    t = np.sort(rng.uniform(0, 20, n))
    y = np.sin(t) + 0.05 * rng.normal(size=n)
    K = np.empty((n, n))
    for i in range(n):
        for j in range(n):
            tau = abs(t[i] - t[j])
            K[i, j] = k_sho(tau, 1.5, 3.0, 1.8) + k_qm32(tau, 1.5, 0.9)
    K += 0.1 * np.eye(n)
    lp = float(scipy.stats.multivariate_normal(np.zeros(n), K).logpdf(y))
    out["cases"].append({"kind": "quasisep", "name": "sho+m32", "t": t.tolist(), "y": y.tolist(), "diag": 0.1, "logp": lp})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
