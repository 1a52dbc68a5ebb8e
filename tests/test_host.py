"""Host logic and the C-ABI surface without a GPU."""

import ctypes
import os
import re

import numpy as np
import pytest

from tinygp_b200 import _cabi, kernels, noise
from tinygp_b200.kernels import quasisep as Q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "b200gp.h")).read()
    declared = set(re.findall(r"\b(b200gp_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200gp_ctx", "b200gp_dense", "b200gp_qs", "b200gp_profile"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200gp.h but not exported"
        assert name in _cabi.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.b200gp_version() >= 100


def test_no_cpu_fallback_when_no_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.B200Error, match="no CPU fallback"):
        _cabi.Context()


def test_kernel_program_lowering():
    k = 1.5 * kernels.ExpSquared(scale=1.2) + kernels.Matern32(2.0) * kernels.RationalQuadratic(alpha=1.5)
    prog = k.program()
    assert prog.shape == (7, 4)
    assert prog[:, 0].tolist() == [0, 2, 17, 3, 7, 17, 16]
    assert prog[1, 1] == 1 and prog[3, 1] == 0 and prog[4, 1] == 0   # L2 default only for ExpSquared
    assert prog[4, 3] == 1.5
    assert sum([kernels.Exp(1.0), kernels.Exp(2.0)]).program().shape == (3, 4)   # __radd__ with 0
    assert (kernels.Exp(1.0) + 2.0).program()[1].tolist() == [0, 0, 2.0, 0]


def test_unsupported_kernels_raise():
    for k in (kernels.DotProduct(), kernels.Polynomial(order=2), kernels.Custom(lambda a, b: a * b)):
        with pytest.raises(NotImplementedError, match="B200"):
            k.program()

    class MyDist(kernels.Distance):
        pass

    with pytest.raises(NotImplementedError, match="B200"):
        kernels.Matern32(1.0, distance=MyDist()).program()
    with pytest.raises(ValueError, match="scalar"):
        kernels.Exp(scale=np.ones(2)).program()
    with pytest.raises(ValueError, match="alpha"):
        kernels.RationalQuadratic()
    with pytest.raises(ValueError, match="gamma"):
        kernels.ExpSineSquared()
    with pytest.raises(ValueError):
        kernels.Constant(np.ones(3))
    with pytest.raises(NotImplementedError):      # noise.py:121-123: no quasiseparable form
        noise.Dense(np.eye(3)).to_qsm()


def test_quasisep_components():
    k = Q.SHO(omega=1.5, quality=3.0, sigma=1.8) + 0.7 * Q.Matern32(scale=1.5, sigma=0.9)
    c = k.component_array()
    assert c.shape == (2, 8)
    assert c[0, :5].tolist() == [3, 1.0, 1.5, 3.0, 1.8]
    assert c[1, :4].tolist() == [1, 0.7, 1.5, 0.9]
    assert k.state_dim() == 4
    with pytest.raises(ValueError):
        Q.Matern32(1.0) + kernels.Exp(1.0)
    # Product (quasisep.py:298-331): one term of chained rows, scaled once; state dimension = product of the factors'
    kp = 0.5 * (Q.Matern32(1.0) * Q.Exp(2.0)) + Q.Exp(3.0)
    cp = kp.component_array()
    assert cp[:, 6].tolist() == [1.0, 0.0, 0.0] and cp[:, 1].tolist() == [0.5, 1.0, 1.0] and kp.state_dim() == 3
    # a Sum inside a Product is multiplied out: (a + b) * c -> a * c + b * c, two chained terms
    cs = ((Q.Matern32(1.0) + 0.3 * Q.Exp(1.0)) * Q.Exp(2.0)).component_array()
    assert cs[:, 0].tolist() == [Q.QS_MATERN32, Q.QS_EXP, Q.QS_EXP, Q.QS_EXP]
    assert cs[:, 6].tolist() == [1.0, 0.0, 1.0, 0.0] and cs[:, 1].tolist() == [1.0, 1.0, 0.3, 1.0]
    with pytest.raises(NotImplementedError):      # ... more than 8 states (3 * 2 + 2 * 2 = 10): no device rows (generator arrays instead)
        ((Q.Matern52(1.0) + Q.Matern32(1.0)) * Q.SHO(1.0, 2.0)).component_array()


QS_PAIRS = [
    (lambda: Q.SHO(1.5, 3.0, 1.8) + 0.7 * Q.Matern32(1.5, 0.9),
     lambda o: o.qs.SHO(1.5, 3.0, 1.8) + o.qs.Scale(o.qs.Matern32(1.5, 0.9), 0.7)),
    (lambda: Q.SHO(1.5, 0.3), lambda o: o.qs.SHO(1.5, 0.3)),
    (lambda: Q.SHO(1.5, 0.5), lambda o: o.qs.SHO(1.5, 0.5)),
    (lambda: Q.Matern52(1.3, 0.8), lambda o: o.qs.Matern52(1.3, 0.8)),
    (lambda: Q.Exp(1.3, 0.8), lambda o: o.qs.Exp(1.3, 0.8)),
    (lambda: Q.Cosine(1.3, 0.8), lambda o: o.qs.Cosine(1.3, 0.8)),
    (lambda: Q.Celerite(1.1, 0.8, 0.9, 0.1), lambda o: o.qs.Celerite(1.1, 0.8, 0.9, 0.1)),
    (lambda: Q.SHO(1.5, 3.0, 1.8) * Q.Matern32(1.5, 0.9), lambda o: o.qs.SHO(1.5, 3.0, 1.8) * o.qs.Matern32(1.5, 0.9)),
]


@pytest.mark.parametrize("case", range(len(QS_PAIRS)))
def test_quasisep_dense_lowering(case):
    """the closed-form k(tau) programs that evaluate a quasiseparable kernel densely on the device agree with the
    oracle's state-space evaluation (kernels/quasisep.py:118-145); checked here through the Python restatement of
    the device interpreter, on the GPU in tests/test_quasisep_gpu.py"""
    from oracle import tinygp_np as o
    from test_transforms import _interp
    kk, oo = QS_PAIRS[case][0](), QS_PAIRS[case][1](o)
    X = np.linspace(0, 3, 7)
    prog, x = kk.lower_for(X)
    got = np.array([[_interp(prog, a, b) for b in x] for a in x])
    np.testing.assert_allclose(got, oo(X, X), rtol=1e-11, atol=1e-13)


def test_noise_diagonal():
    d = noise.Diagonal(np.array([1.0, 2.0]))
    assert (d + np.zeros((2, 2))).tolist() == [[1, 0], [0, 2]]
    assert (np.ones((2, 2)) + d).tolist() == [[2, 1], [1, 3]]
    assert (d @ np.array([3.0, 4.0])).tolist() == [3, 8]
    with pytest.raises(ValueError):
        noise.Diagonal(1.0)
