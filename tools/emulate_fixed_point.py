"""CPU emulation of the int8 fixed-point Cholesky update (tinygp_b200/csrc/ozaki.cu) in NumPy: same row scales, same
digit cutting (cut_digits_kernel), same dropped-pair diagonal correction, same left-looking block columns; the integer
dot products are done as fp64 BLAS products of integer-valued matrices (exact: every sum is < 2^53).

Purpose: how does the log-probability error depend on the number of digit planes S?  (The GPU answer at full size is
the first measurement of the next round; this gives the trend at oracle-friendly sizes with the bench's point density.)

    python tools/emulate_fixed_point.py [N] [nb] [S,S,...]   # developer tool, not part of the product or the tests
"""

import sys
import time

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bench import make_dense_problem  # noqa: E402


def cut(L, rs, S):
    x = L / rs[:, None] * 64.0
    planes = []
    for _ in range(S):
        q = np.clip(np.rint(x), -127, 127)
        x = (x - q) * 128.0
        planes.append(q)
    return planes


def logp_fixed_point(K, y, S, nb):
    n = K.shape[0]
    rs = 2.0 ** np.ceil(np.log2(np.sqrt(np.diag(K))))
    L = np.zeros_like(K)
    planes = [np.zeros_like(K) for _ in range(S)]
    corr = np.zeros(n)
    for c0 in range(0, n, nb):
        c1 = min(n, c0 + nb)
        C = K[c0:, c0:c1].copy()
        if c0 > 0:
            for g in range(S):                      # one fp64 rounding per digit group, like the kernel's epilogue
                acc = np.zeros_like(C)
                for s in range(g + 1):
                    acc += planes[s][c0:, :c0] @ planes[g - s][c0:c1, :c0].T
                C -= (2.0 ** -(12 + 7 * g)) * (rs[c0:, None] * rs[None, c0:c1]) * acc
            idx = np.arange(c1 - c0)
            C[idx, idx] -= corr[c0:c1]
        Ljj = sla.cholesky(C[: c1 - c0], lower=True, check_finite=False)
        L[c0:c1, c0:c1] = Ljj
        if c1 < n:
            L[c1:, c0:c1] = sla.solve_triangular(Ljj, C[c1 - c0:].T, lower=True, check_finite=False).T
        P = cut(L[c0:, c0:c1], rs[c0:], S)
        for s in range(S):
            planes[s][c0:, c0:c1] = P[s]
        for g in range(S, 2 * (S - 1) + 1):         # dropped pairs of this panel's columns, exact integer sums
            acc = np.zeros(n - c0)
            for s in range(g - (S - 1), S):
                acc += np.sum(P[s] * P[g - s], axis=1)
            corr[c0:] += rs[c0:] ** 2 * acc * 2.0 ** -(12 + 7 * g)
    alpha = sla.solve_triangular(L, y, lower=True, check_finite=False)
    return -0.5 * alpha @ alpha - np.sum(np.log(np.diag(L))) - 0.5 * n * np.log(2 * np.pi)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    X, y, diag, scale = make_dense_problem(n)
    d2 = np.sum((X[:, None, :] - X[None, :, :]) ** 2, axis=-1)
    K = np.exp(-0.5 * d2 / scale ** 2) + np.diag(diag)
    del d2
    Lx = sla.cholesky(K, lower=True, check_finite=False)
    ax = sla.solve_triangular(Lx, y, lower=True, check_finite=False)
    exact = -0.5 * ax @ ax - np.sum(np.log(np.diag(Lx))) - 0.5 * n * np.log(2 * np.pi)
    ev = np.linalg.eigvalsh(K)
    print(f"N={n} nb={nb} exact logp={exact!r} cond(K)={ev[-1] / ev[0]:.3g}")
    planes_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 7, 6, 5, 4, 3]
    for S in planes_list:
        t0 = time.time()
        lp = logp_fixed_point(K, y, S, nb)
        print(f"  S={S}: logp={lp!r} rel_err={abs(lp - exact) / abs(exact):.3e} ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
