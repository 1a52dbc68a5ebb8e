"""Observation noise models (reference: src/tinygp/noise.py).  Only ``Diagonal`` is supported by the
B200 backend: ``Dense`` defeats the fused build and ``Banded`` raises the quasiseparable order."""

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
    def __init__(self, *a, **k):
        raise NotImplementedError("noise.Dense is unsupported by the B200 solver backend")


class Banded(Noise):
    def __init__(self, *a, **k):
        raise NotImplementedError("noise.Banded is unsupported by the B200 solver backend")
