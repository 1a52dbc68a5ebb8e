"""The solver plugin protocol, re-declared verbatim from src/tinygp/solvers/solver.py:15-82."""

from __future__ import annotations

__all__ = ["Solver"]

from typing import Any


class Solver:
    def __init__(self, kernel, X, noise, *, covariance: Any | None = None):
        del kernel, X, noise, covariance
        raise NotImplementedError

    @classmethod
    def init(cls, kernel, X, noise, *, covariance: Any | None = None):  # solver.py:29-38
        return cls(kernel, X, noise, covariance=covariance)

    def variance(self):
        """The diagonal of the covariance matrix"""
        raise NotImplementedError

    def covariance(self):
        """The evaluated covariance matrix"""
        raise NotImplementedError

    def normalization(self):
        """(log_det + n*log(2*pi))/2"""
        raise NotImplementedError

    def solve_triangular(self, y, *, transpose: bool = False):
        """Solve L @ x = y (or L.T @ x = y) for K = L @ L.T"""
        raise NotImplementedError

    def dot_triangular(self, y):
        """L @ y"""
        raise NotImplementedError

    def condition(self, kernel, X_test, noise) -> Any:
        raise NotImplementedError
