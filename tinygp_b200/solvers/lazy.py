"""The conditioned process at the inputs of a QuasisepSolver without an N x N matrix unless one is asked for.

The reference keeps ``gp.condition(y)`` linear in N for quasiseparable kernels by returning the conditioned covariance
as a SymmQSM (solvers/quasisep/solver.py:124-129) and factoring that with another QuasisepSolver.  Here the same
*values* are served in two tiers: ``variance()`` -- what ``cond.gp.variance`` / ``predict(return_var=True)`` read,
solver.py:84-85 -- comes from the O(N) backward scan ``QuasisepSolver.conditioned_variance``; every other method
needs the matrix and materialises the dense ``DirectSolver`` (``b200gp_qs_condition`` + Cholesky) on first use.
"""

from __future__ import annotations

__all__ = ["LazyConditionedSolver"]

from typing import Any

from tinygp_b200.solvers.direct import DirectSolver
from tinygp_b200.solvers.solver import Solver


class LazyConditionedSolver(Solver):
    def __init__(self, parent, conditioned_kernel, predictive_kernel, X, noise):
        self.parent, self.kernel, self.predictive_kernel = parent, conditioned_kernel, predictive_kernel
        self.X, self.noise = X, noise
        self._inner: DirectSolver | None = None

    def _dense(self) -> DirectSolver:
        if self._inner is None:
            cov = self.parent.condition(self.predictive_kernel, None, self.noise)     # solver.py:124-129, densified
            self._inner = DirectSolver(self.kernel, self.X, self.noise, covariance=cov)
        return self._inner

    def variance(self):
        if self._inner is not None:
            return self._inner.variance()
        return self.parent.conditioned_variance(self.noise)

    def covariance(self):
        return self._dense().covariance()

    def normalization(self):
        return self._dense().normalization()

    def solve_triangular(self, y, *, transpose: bool = False):
        return self._dense().solve_triangular(y, transpose=transpose)

    def dot_triangular(self, y):
        return self._dense().dot_triangular(y)

    def condition(self, kernel, X_test, noise) -> Any:
        return self._dense().condition(kernel, X_test, noise)
