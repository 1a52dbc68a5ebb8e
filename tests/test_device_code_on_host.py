"""The __host__ __device__ core of the quasiseparable CUDA path (tinygp_b200/csrc/qs_core.cuh: model lowering,
per-point generators `qs_gen`, the `GramBack` scan monoid and its chunk / tree / replay decomposition) compiled for
the CPU by tests/csrc/qs_hostcheck.cu and compared with the oracle.  Same source as the kernels; what is NOT covered
here is the launch geometry and memory system, which the `-m gpu` tests cover."""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import tinygp_np as o
from tinygp_b200.kernels import quasisep as Q

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "qs_hostcheck.cu")
OUT = os.path.join(HERE, "csrc", "_build", "libqs_hostcheck.so")
DEPS = [SRC, os.path.join(HERE, "..", "tinygp_b200", "csrc", "qs_core.cuh"),
        os.path.join(HERE, "..", "tinygp_b200", "csrc", "common.cuh")]

KERNELS = {
    "sho+m32": (Q.SHO(1.5, 3.0, 1.8) + Q.Matern32(1.5, 0.9), o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Matern32(1.5, 0.9)),
    "sho_critical": (Q.SHO(1.2, 0.5, 1.1), o.qs.SHO(1.2, 0.5, 1.1)),
    "sho_overdamped": (Q.SHO(1.2, 0.2, 1.1), o.qs.SHO(1.2, 0.2, 1.1)),
    "exp": (Q.Exp(1.7, 0.8), o.qs.Exp(1.7, 0.8)),
    "m52": (Q.Matern52(2.5, 1.3), o.qs.Matern52(2.5, 1.3)),
    "celerite": (Q.Celerite(1.1, 0.1, 0.3, 1.5), o.qs.Celerite(1.1, 0.1, 0.3, 1.5)),
    "cosine+exp": (Q.Cosine(3.0, 0.7) + Q.Exp(2.0, 0.5), o.qs.Cosine(3.0, 0.7) + o.qs.Exp(2.0, 0.5)),
    "scaled_sum3": (2.0 * Q.Matern32(1.2) + Q.SHO(0.8, 4.0, 0.6) + 0.5 * Q.Exp(5.0),
                    o.qs.Scale(o.qs.Matern32(1.2, 1.0), 2.0) + o.qs.SHO(0.8, 4.0, 0.6) + o.qs.Scale(o.qs.Exp(5.0, 1.0), 0.5)),
}


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                        "-o", OUT, SRC], check=True)
    return ctypes.CDLL(OUT)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _data(n, seed=0, ties=True):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, n / 8.0, n))
    if ties and n > 10:
        t[7] = t[6]
    return t, rng.uniform(0.05, 0.2, n)


@pytest.mark.parametrize("name", list(KERNELS))
def test_model_and_generators_match_the_oracle(lib, name):
    """build_model + qs_gen (kernels/quasisep.py:102-116 and the state-space models :343-673)"""
    k, ko = KERNELS[name]
    comps = k.component_array()
    t, _ = _data(200)
    d, p, q, a = ko.to_symm_qsm(t)
    J = ctypes.c_int()
    qm, hm, d0 = np.zeros(8), np.zeros(8), ctypes.c_double()
    assert lib.hostcheck_model(_p(comps), comps.shape[0], ctypes.byref(J), _p(qm), _p(hm), ctypes.byref(d0)) == 0
    assert J.value == p.shape[1] == k.state_dim()
    np.testing.assert_allclose(qm[:J.value], q[0], rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(d0.value, d[0], rtol=1e-14)
    ad, pd = np.zeros(a.shape), np.zeros(p.shape)      # (zeros_like would copy the oracle's swapped strides)
    assert lib.hostcheck_generators(_p(comps), comps.shape[0], _p(t), ctypes.c_int64(t.size), _p(ad), _p(pd)) == 0
    np.testing.assert_allclose(ad, a, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(pd, p, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("chunk", [1, 7, 64])
@pytest.mark.parametrize("n", [1, 5, 300, 5000])
@pytest.mark.parametrize("name", ["sho+m32", "m52", "exp", "scaled_sum3"])
def test_inverse_diagonal_scan_matches_the_oracle(lib, name, n, chunk):
    """diag((K + N)^-1) by the GramBack scan: chunk composites, the fan-in-16 tree (two levels at n = 5000, chunk 1)
    and the replay give the diagonal of the dense inverse"""
    k, ko = KERNELS[name]
    comps = k.component_array()
    t, noise = _data(n, seed=n)
    s = o.QuasisepSolver(ko, t, o.Diagonal(noise))
    out = np.zeros(n)
    c, w = np.ascontiguousarray(s.c), np.ascontiguousarray(s.w)
    assert lib.hostcheck_inverse_diagonal(_p(comps), comps.shape[0], _p(t), ctypes.c_int64(n), _p(c), _p(w), chunk,
                                          _p(out)) == 0
    if n <= 300:
        want = np.diag(np.linalg.inv(s.covariance()))
    else:   # columns of L^-1 by the oracle's sequential solve: (Sigma^-1)_ii = |L^-1 e_i|^2 on a sample of columns
        idx = np.unique(np.r_[0, 1, n // 2, n - 2, n - 1, np.random.default_rng(1).integers(0, n, 12)])
        E = np.zeros((n, idx.size)); E[idx, np.arange(idx.size)] = 1.0
        Z = s.solve_triangular(s.solve_triangular(E), transpose=True)      # Sigma^-1 e_i
        want, out = Z[idx, np.arange(idx.size)], out[idx]
    np.testing.assert_allclose(out, want, rtol=1e-10, atol=0)
