"""One-off: BASELINE config 3 at full size (N = 131072, Matern52 + RationalQuadratic with the Euclidean metric) through the
NATIVE fp64 DMMA path (ozaki_slices = 0: no int8 emulation anywhere; 137 GB matrix on one B200), as an independent known
answer for the int8 / sharded runs.  No CPU in this container can hold the 69 GB lower triangle (62 GB of RAM), so unlike
tests/golden/full_size.json["c2"/"c3s"] this anchor is not LAPACK; the DMMA path itself is pinned to LAPACK at N = 65536 /
65537 (tests/test_golden.py).  Writes gpurun_out/c3_golden.json; merged by hand into tests/golden/full_size.json["c3"]."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tinygp_b200 import GaussianProcess, _cabi, kernels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
rng = np.random.default_rng(49383)
side = 25.0 * (n / 131072.0) ** (1.0 / 3.0)
X = np.ascontiguousarray(rng.uniform(0.0, side, (n, 3)))
y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=n)
L2 = kernels.L2Distance()
k = 1.5 * kernels.Matern52(2.0, L2) + 0.7 * kernels.RationalQuadratic(1.5, L2, alpha=1.5)
ctx = _cabi.get_context()
ctx.set_option("ozaki_slices", 0)
t0 = time.perf_counter()
lp = GaussianProcess(k, X, diag=0.1).log_probability(y)
dt = time.perf_counter() - t0
out = {"n": n, "seed": 49383, "diag": 0.1, "log_probability": lp, "x_checksum": float(X.sum()), "y_checksum": float(y.sum()),
       "lapack": "B200 native fp64 DMMA path (ozaki_slices=0), not LAPACK: see tools/make_c3_golden_gpu.py", "seconds": dt}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c3_golden.json", "w"), indent=1)
print(json.dumps(out))
