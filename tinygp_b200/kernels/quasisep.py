"""Quasiseparable (celerite-style) kernels (reference: src/tinygp/kernels/quasisep.py).

Each kernel is a parameter holder that lowers to a list of state-space *components*
(include/b200gp.h, B200GP_QS_*): a ``Sum`` is block diagonal (quasisep.py:241-295) so it is the
concatenation of its terms' components, a ``Scale`` multiplies the stationary covariance
(quasisep.py:334-340).  The per-point generators d, p, q, a (``to_symm_qsm``, quasisep.py:102-116)
are computed on the device in registers from t[k] - t[k-1]; they are never materialised unless
``QuasisepSolver.generators()`` asks for them.
"""

from __future__ import annotations

__all__ = ["Quasisep", "Sum", "Product", "Scale", "Celerite", "SHO", "Exp", "Matern32", "Matern52",
           "Cosine", "CARMA"]

import numpy as np

from tinygp_b200.kernels.base import Kernel

QS_EXP, QS_MATERN32, QS_MATERN52, QS_SHO, QS_CELERITE, QS_COSINE = range(6)
STATE_DIM = {QS_EXP: 1, QS_MATERN32: 2, QS_MATERN52: 3, QS_SHO: 2, QS_CELERITE: 2, QS_COSINE: 2}


class Quasisep(Kernel):
    """Base class (quasisep.py:47-215)."""

    def components(self) -> list[tuple]:
        """[(kind, pinf_scale, p0, p1, p2, p3)] for the device."""
        raise NotImplementedError(
            f"{type(self).__name__} is unsupported by the B200 quasiseparable solver backend")

    def component_array(self) -> np.ndarray:
        comps = self.components()
        out = np.zeros((len(comps), 8))
        for i, c in enumerate(comps):
            out[i, : len(c)] = c
        return out

    def state_dim(self) -> int:
        return sum(STATE_DIM[int(c[0])] for c in self.components())

    def coord_to_sortable(self, X):
        return X

    # dense evaluation of a quasiseparable kernel goes through its closed form k(tau)
    def _ktau(self, tau):
        raise NotImplementedError

    def __call__(self, X1, X2=None):
        X1 = np.asarray(X1, dtype=np.float64)
        if X2 is None:
            return self._ktau(np.zeros_like(X1))
        X2 = np.asarray(X2, dtype=np.float64)
        if X1.ndim != 1 or X2.ndim != 1:
            raise ValueError("quasiseparable kernels take 1-D sortable coordinates")
        return self._ktau(np.abs(X1[:, None] - X2[None, :]))

    def matmul(self, X1, X2=None, y=None):
        if y is None:
            y, X2 = X2, None
        if X2 is None:
            X2 = X1
        return self(X1, X2) @ np.asarray(y, dtype=np.float64)

    def lower(self, lc=None):
        raise NotImplementedError(
            "quasiseparable kernels are handled by QuasisepSolver; they do not lower to a dense kernel program")

    # algebra (quasisep.py:165-199)
    def __add__(self, other):
        if not isinstance(other, Quasisep):
            raise ValueError("Quasisep kernels can only be added to other Quasisep kernels")
        return Sum(self, other)

    def __radd__(self, other):
        if not isinstance(other, Kernel) and np.ndim(other) == 0 and other == 0:
            return self
        if not isinstance(other, Quasisep):
            raise ValueError("Quasisep kernels can only be added to other Quasisep kernels")
        return Sum(other, self)

    def __mul__(self, other):
        if isinstance(other, Quasisep):
            return Product(self, other)
        if isinstance(other, Kernel) or np.ndim(other) != 0:
            raise ValueError("Quasisep kernels can only be multiplied by scalars and other Quasisep kernels")
        return Scale(kernel=self, scale=other)

    def __rmul__(self, other):
        if isinstance(other, Quasisep):
            return Product(other, self)
        if isinstance(other, Kernel) or np.ndim(other) != 0:
            raise ValueError("Quasisep kernels can only be multiplied by scalars and other Quasisep kernels")
        return Scale(kernel=self, scale=other)


class Sum(Quasisep):
    """quasisep.py:241-295"""

    def __init__(self, kernel1: Quasisep, kernel2: Quasisep, use_block: bool = True):
        self.kernel1, self.kernel2, self.use_block = kernel1, kernel2, use_block

    def components(self):
        return self.kernel1.components() + self.kernel2.components()

    def _ktau(self, tau):
        return self.kernel1._ktau(tau) + self.kernel2._ktau(tau)


class Scale(Quasisep):
    """quasisep.py:334-340"""

    def __init__(self, kernel: Quasisep, scale):
        self.kernel, self.scale = kernel, scale

    def components(self):
        s = float(self.scale)
        return [(c[0], c[1] * s) + tuple(c[2:]) for c in self.kernel.components()]

    def _ktau(self, tau):
        return float(self.scale) * self.kernel._ktau(tau)


class Product(Quasisep):
    """quasisep.py:298-331 -- Kronecker-product states: out of scope of the first B200 pass."""

    def __init__(self, kernel1, kernel2):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def _ktau(self, tau):
        return self.kernel1._ktau(tau) * self.kernel2._ktau(tau)


class Celerite(Quasisep):
    """exp(-c tau) [a cos(d tau) + b sin(d tau)], quasisep.py:343-401"""

    def __init__(self, a, b, c, d):
        self.a, self.b, self.c, self.d = a, b, c, d

    def components(self):
        return [(QS_CELERITE, 1.0, float(self.a), float(self.b), float(self.c), float(self.d))]

    def _ktau(self, tau):
        return np.exp(-self.c * tau) * (self.a * np.cos(self.d * tau) + self.b * np.sin(self.d * tau))


class SHO(Quasisep):
    """quasisep.py:404-488"""

    def __init__(self, omega, quality, sigma=1.0):
        self.omega, self.quality, self.sigma = omega, quality, sigma

    def components(self):
        return [(QS_SHO, 1.0, float(self.omega), float(self.quality), float(self.sigma), 0.0)]

    def _ktau(self, tau):
        w, q, s2 = float(self.omega), float(self.quality), float(self.sigma) ** 2
        e = np.exp(-0.5 * w * tau / q)
        if np.allclose(q, 0.5):
            return s2 * np.exp(-w * tau) * (1 + w * tau)
        if q > 0.5:
            g = np.sqrt(4 * q * q - 1)
            arg = 0.5 * g * w * tau / q
            return s2 * e * (np.cos(arg) + np.sin(arg) / g)
        f = np.sqrt(1 - 4 * q * q)
        arg = 0.5 * f * w * tau / q
        return s2 * e * (np.cosh(arg) + np.sinh(arg) / f)


class Exp(Quasisep):
    """sigma^2 exp(-tau / scale), quasisep.py:491-525"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_EXP, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def _ktau(self, tau):
        return float(self.sigma) ** 2 * np.exp(-tau / self.scale)


class Matern32(Quasisep):
    """quasisep.py:528-569"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_MATERN32, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def _ktau(self, tau):
        f = np.sqrt(3) / self.scale
        return float(self.sigma) ** 2 * (1 + f * tau) * np.exp(-f * tau)


class Matern52(Quasisep):
    """quasisep.py:572-633"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_MATERN52, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def _ktau(self, tau):
        f = np.sqrt(5) / self.scale
        return float(self.sigma) ** 2 * (1 + f * tau + np.square(f * tau) / 3) * np.exp(-f * tau)


class Cosine(Quasisep):
    """quasisep.py:636-673"""

    def __init__(self, scale, sigma=1.0):
        self.scale, self.sigma = scale, sigma

    def components(self):
        return [(QS_COSINE, 1.0, float(self.scale), float(self.sigma), 0.0, 0.0)]

    def _ktau(self, tau):
        return float(self.sigma) ** 2 * np.cos(2 * np.pi * tau / self.scale)


class CARMA(Quasisep):
    """quasisep.py:695-1030 -- complex-root state space: out of scope of the first B200 pass."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("CARMA is unsupported by the B200 quasiseparable solver backend")
