"""Observation noise models (reference: src/tinygp/noise.py).  ``Diagonal`` is the one the fused device paths take (it enters
the build kernels / the state-space generators); ``Dense`` and ``Banded`` go through the precomputed-covariance entry points
(``b200gp_dense_create_from_cov``; a device ``SymmQSM`` sum factored by ``b200gp_qsm_cholesky``)."""

from __future__ import annotations

__all__ = ["Noise", "Diagonal", "Dense", "Banded"]

import numpy as np


class Noise:
    __array_priority__ = 2001  # noise.py:30

    def diagonal(self):
        raise NotImplementedError


class Diagonal(Noise):
    """noise.py:55-95"""

    def __init__(self, diag):
        diag = np.asarray(diag, dtype=np.float64)
        if diag.ndim != 1:  # noise.py:67-72
            raise ValueError(
                "The diagonal for the noise model be the same shape as the data; "
                "if passing a constant, it should be broadcasted first"
            )
        self.diag = np.ascontiguousarray(diag)

    def diagonal(self):
        return self.diag

    def _add(self, other):  # noise.py:77-78
        out = np.array(other, dtype=np.float64, copy=True)
        idx = np.arange(out.shape[0])
        out[idx, idx] += self.diag
        return out

    __add__ = _add
    __radd__ = _add

    def __matmul__(self, other):  # noise.py:86-90
        other = np.asarray(other, dtype=np.float64)
        return self.diag * other if other.ndim == 1 else self.diag[:, None] * other

    def to_qsm(self):  # noise.py:92-95
        from tinygp_b200.solvers.quasisep.core import DiagQSM
        return DiagQSM(d=self.diag)


class Dense(Noise):
    """A full-rank N x N observation model (noise.py:98-123).  DirectSolver only: the kernel matrix is built on the device,
    this matrix is added to it on its way to the device factorisation (solvers/direct.py)."""

    def __init__(self, value):
        value = np.asarray(value, dtype=np.float64)
        if value.ndim != 2 or value.shape[0] != value.shape[1]:
            raise ValueError("noise.Dense takes a square (N, N) matrix")
        self.value = np.ascontiguousarray(value)

    def diagonal(self):
        return np.diag(self.value).copy()

    def __add__(self, other):
        return self.value + np.asarray(other, dtype=np.float64)

    def __radd__(self, other):
        return np.asarray(other, dtype=np.float64) + self.value

    def __matmul__(self, other):
        return self.value @ np.asarray(other, dtype=np.float64)

    def to_qsm(self):  # noise.py:121-123
        """This cannot be compactly represented as a quasiseparable matrix"""
        raise NotImplementedError


class Banded(Noise):
    """``diag`` (N,) plus ``J`` symmetric off-diagonals ``off_diags`` (N, J), row n holding N[n, n+1 .. n+J] (noise.py:126-240;
    Delisle et al. 2020).  For the QuasisepSolver it is a SymmQSM of order J (``to_qsm``) added to the kernel's on the device."""

    def __init__(self, diag, off_diags):
        self.diag = np.ascontiguousarray(np.asarray(diag, dtype=np.float64))
        self.off_diags = np.ascontiguousarray(np.asarray(off_diags, dtype=np.float64))
        if self.diag.ndim != 1 or self.off_diags.ndim != 2 or self.off_diags.shape[0] != self.diag.shape[0]:
            raise ValueError("noise.Banded takes diag (N,) and off_diags (N, J)")

    def diagonal(self):
        return self.diag

    def _add(self, other):  # noise.py:196-215
        out = np.array(other, dtype=np.float64, copy=True)
        n = out.shape[0]
        idx = np.arange(n)
        out[idx, idx] += self.diag
        for j in range(self.off_diags.shape[1]):      # the j-th off-diagonal on both sides
            r = idx[: max(n - j - 1, 0)]
            out[r, r + j + 1] += self.off_diags[: r.size, j]
            out[r + j + 1, r] += self.off_diags[: r.size, j]
        return out

    __add__ = _add
    __radd__ = _add

    def __matmul__(self, other):  # noise.py:223-224
        return self.to_qsm() @ other

    def to_qsm(self):  # noise.py:226-240: p = e_1, q = the row of off-diagonals, a = the shift
        from tinygp_b200.solvers.quasisep import core
        n, j = self.off_diags.shape
        p = np.zeros((n, j))
        p[:, 0] = 1.0
        a = np.broadcast_to(np.eye(j, k=1), (n, j, j)).copy()
        return core.SymmQSM(diag=core.DiagQSM(d=self.diag), lower=core.StrictLowerTriQSM(p=p, q=self.off_diags, a=a))
