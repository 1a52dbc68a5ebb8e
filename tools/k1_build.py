"""K1 stand-alone: the kernel-matrix build tile (build_rect_kernel) through the public Kernel.__call__ (kernels/base.py:84-103).
Used under ncu to capture achieved HBM GB/s of the build (north_star evidence)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinygp_b200 import kernels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(49382)
X = rng.uniform(0, 20.0 * (n / 65536.0) ** (1 / 3), (n, 3))
k = 1.0 * kernels.ExpSquared(scale=1.0)
for _ in range(2):
    t0 = time.perf_counter()
    K = k(X, X)
    print("build", n, time.perf_counter() - t0, K[0, 0], K[5, 7])
L2 = kernels.L2Distance()
k3 = 1.5 * kernels.Matern52(2.0, L2) + 0.7 * kernels.RationalQuadratic(1.5, L2, alpha=1.5)
K = k3(X, X)
print("c3 kernel", K[0, 0], K[5, 7])
