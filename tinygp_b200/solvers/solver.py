"""The solver plugin protocol, re-declared verbatim from src/tinygp/solvers/solver.py:15-82."""

from __future__ import annotations

__all__ = ["Solver", "ConditionedCovariance"]

from typing import Any

import numpy as np


class ConditionedCovariance(np.ndarray):
    """The dense matrix a ``Solver.condition`` returns, tagged with whether the predictive noise is already on its
    diagonal.  The reference's DirectSolver.condition adds it (direct.py:88-92) and so does QuasisepSolver's QSM
    branch (solvers/quasisep/solver.py:124-129), its dense branch does not (:131-139) -- while the conditioned GP's
    ``variance`` always carries it (direct.py:49 / solver.py:84-85).  ``DirectSolver(covariance=...)`` reads the tag."""

    includes_noise = True

    def __array_finalize__(self, obj):
        self.includes_noise = getattr(obj, "includes_noise", True)

    @classmethod
    def tag(cls, matrix, includes_noise: bool) -> "ConditionedCovariance":
        out = np.asarray(matrix).view(cls)
        out.includes_noise = bool(includes_noise)
        return out


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
