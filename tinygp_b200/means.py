"""Mean functions (reference: src/tinygp/means.py).  Evaluated on the host: O(N), off the hot path."""

from __future__ import annotations

__all__ = ["MeanBase", "Mean", "Conditioned"]

import numpy as np


class MeanBase:
    def __call__(self, X):
        raise NotImplementedError

    def vmap(self, X):
        X = np.asarray(X, dtype=np.float64)
        return np.array([self(x) for x in X], dtype=np.float64).reshape(X.shape[0])


class Mean(MeanBase):
    """A constant or a callable of a single coordinate (means.py:31-55)."""

    def __init__(self, value):
        if callable(value):
            self.func, self.value = value, 0.0
        else:
            self.func, self.value = None, value

    def __call__(self, X):
        return self.func(X) if self.func is not None else self.value

    def vmap(self, X):
        if self.func is None:
            v = float(self.value)
            if v == 0.0:                # the default mean: calloc-backed zeros, no 8 N bytes written on the host (N = 10^7: 80 MB)
                return np.zeros(np.shape(X)[0])
            return np.full(np.shape(X)[0], v)
        return super().vmap(X)


class Conditioned(MeanBase):
    """means.py:58-86: mu(x) = k(x, X) @ alpha (+ mean_function(x))"""

    def __init__(self, X, alpha, kernel, include_mean, mean_function=None):
        self.X, self.alpha, self.kernel = X, alpha, kernel
        self.include_mean, self.mean_function = include_mean, mean_function

    def vmap(self, X):
        mu = self.kernel.matmul(X, self.X, self.alpha)
        if self.include_mean and self.mean_function is not None:
            mu = mu + self.mean_function.vmap(X)
        return mu

    def __call__(self, X):
        return self.vmap(np.asarray(X)[None])[0]
