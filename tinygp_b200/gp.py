"""``GaussianProcess`` re-declared without JAX (reference: src/tinygp/gp.py:30-393).

Same constructor and method signatures, so ``solver=`` drops in; all linear algebra goes through
the six ``Solver`` methods exactly where the reference calls them (gp.py:120,124,201,310,315,320,334).
"""

from __future__ import annotations

__all__ = ["GaussianProcess", "ConditionResult"]

from typing import Any, NamedTuple

import numpy as np

from tinygp_b200 import kernels, means
from tinygp_b200.kernels.quasisep import Quasisep
from tinygp_b200.noise import Diagonal, Noise
from tinygp_b200.solvers import DirectSolver, QuasisepSolver
from tinygp_b200.solvers.quasisep.core import SymmQSM


class GaussianProcess:
    def __init__(self, kernel, X, *, diag=None, noise: Noise | None = None, mean=None, solver: Any | None = None,
                 mean_value=None, covariance_value: Any | None = None, **solver_kwargs: Any):
        self.kernel = kernel
        if isinstance(X, (dict, tuple)) or np.asarray(X).dtype == object or np.ndim(X) not in (1, 2):
            raise ValueError("coordinates must be an array of shape (N,) or (N, D); pytree coordinates (gp.py:40-44) "
                             "are unsupported by the B200 solver backend")
        self.X = X
        self._zero_mean = mean is None and mean_value is None      # default Mean(0.0): y - loc == y (skips an N-vector pass)
        if isinstance(mean, means.MeanBase):  # gp.py:81-86
            self.mean_function = mean
        elif mean is None:
            self.mean_function = means.Mean(0.0)
        else:
            self.mean_function = means.Mean(mean)
        if mean_value is None:
            mean_value = self.mean_function.vmap(self.X)  # gp.py:87-88
        mean_value = np.asarray(mean_value, dtype=np.float64)
        if mean_value.ndim != 1:
            raise ValueError(f"Invalid mean shape: expected ndim = 1, got ndim={mean_value.ndim}")
        self.num_data = mean_value.shape[0]
        self.dtype = mean_value.dtype
        self.mean = mean_value
        if noise is None:  # gp.py:96-99
            diag = _default_diag(self.mean) if diag is None else diag
            noise = Diagonal(diag=_as_full_diag(diag, self.mean.shape))
        self.noise = noise
        if solver is None:  # gp.py:101-105
            solver = QuasisepSolver if (isinstance(covariance_value, SymmQSM) or isinstance(kernel, Quasisep)) else DirectSolver
        self.solver = solver(kernel, self.X, self.noise, covariance=covariance_value, **solver_kwargs)

    @property
    def loc(self):
        return self.mean

    @property
    def variance(self):
        return self.solver.variance()

    @property
    def covariance(self):
        return self.solver.covariance()

    def log_probability(self, y):  # gp.py:126-138
        hook = getattr(self.solver, "whitened_sumsq", None)
        if hook is not None:     # same value as below with |alpha|^2 reduced on the device (no N-vector round trip)
            with np.errstate(all="ignore"):
                yv = np.asarray(y, dtype=np.float64)
                loglike = -0.5 * hook(yv if self._zero_mean else yv - self.loc) - self.solver.normalization()
            return loglike if np.isfinite(loglike) else -np.inf
        return self._compute_log_prob(self._get_alpha(y))

    def condition(self, y, X_test=None, *, diag=None, noise: Noise | None = None, include_mean: bool = True,
                  kernel=None) -> "ConditionResult":
        """gp.py:140-223"""
        self._check_X_test(X_test)
        alpha, log_prob, mean_value = self._condition(y, X_test, include_mean, kernel)
        if kernel is None:
            kernel = self.kernel
        if noise is None:
            diag = _default_diag(mean_value) if diag is None else diag
            noise = Diagonal(diag=np.broadcast_to(np.asarray(diag, dtype=np.float64), mean_value.shape).copy())
        cond_kernel = kernels.Conditioned(self.X, self.solver, kernel)
        cond_mean = means.Conditioned(self.X, alpha, kernel, include_mean=include_mean, mean_function=self.mean_function)
        covariance_value = self.solver.condition(kernel, X_test, noise)  # gp.py:201
        if X_test is None:
            X_test = self.X
        gp = GaussianProcess(  # gp.py:208-221: the conditional GP factors the M x M covariance again
            cond_kernel,
            X_test,
            noise=noise,
            mean=cond_mean,
            mean_value=mean_value,
            covariance_value=covariance_value,
            # a SymmQSM (solver.py:124-129) is factored in O(N) by QuasisepSolver(covariance=...), gp.py:101-103
            solver=None if isinstance(covariance_value, SymmQSM) else DirectSolver,
        )
        return ConditionResult(log_prob, gp)

    def predict(self, y, X_test=None, *, kernel=None, include_mean: bool = True, return_var: bool = False,
                return_cov: bool = False):
        """gp.py:225-271.  The reference's ``predict`` is jitted with ``return_var`` / ``return_cov`` static
        (gp.py:225-229), so XLA never computes the conditioned covariance when only the mean is asked for; the same
        holds here explicitly -- the mean needs ``alpha`` and one kernel product, not ``solver.condition`` (which for a
        quasiseparable N = 10^7 series would be an N x N matrix)."""
        if not (return_var or return_cov):
            self._check_X_test(X_test)
            return self._condition(y, X_test, include_mean, kernel)[2]
        if return_var and X_test is None and kernel is None and getattr(self.solver, "conditioned_variance", None) is not None:
            # QuasisepSolver at the inputs: the reference's QSM branch (solver.py:124-129) keeps this O(N); so does the
            # backward scan behind conditioned_variance -- same values as cond.variance below
            mean_value = self._condition(y, None, include_mean, None)[2]
            noise = Diagonal(diag=np.full(mean_value.shape, _default_diag(mean_value)))
            return mean_value, self.solver.conditioned_variance(noise)
        _, cond = self.condition(y, X_test, kernel=kernel, include_mean=include_mean)
        if return_var:
            return cond.loc, cond.variance
        if return_cov:
            return cond.loc, cond.covariance
        return cond.loc

    def sample(self, key, shape=None):
        """gp.py:273-311.  ``key`` seeds numpy's Generator (JAX's threefry stream is not reproducible
        without JAX -- sample-value parity is statistical only, as in tests/test_gp.py:24-38)."""
        rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
        shape = (self.num_data,) if shape is None else (self.num_data,) + tuple(shape)
        normal_samples = rng.standard_normal(shape)
        return self.mean + np.moveaxis(self.solver.dot_triangular(normal_samples), 0, -1)

    def _check_X_test(self, X_test):  # gp.py:180-191
        if X_test is not None:
            a, b = np.asarray(self.X), np.asarray(X_test)
            if a.ndim != b.ndim or a.shape[1:] != b.shape[1:]:
                raise ValueError(
                    "`X_test` must have the same tree structure as the input `X`, "
                    "and all but the leading dimension must have matching sizes"
                )

    # -- internals (gp.py:313-361) -------------------------------------------------------------
    def _compute_log_prob(self, alpha):
        with np.errstate(all="ignore"):
            loglike = -0.5 * np.sum(np.square(alpha)) - self.solver.normalization()
        return loglike if np.isfinite(loglike) else -np.inf

    def _get_alpha(self, y):
        return self.solver.solve_triangular(np.asarray(y, dtype=np.float64) - self.loc)

    def _condition(self, y, X_test, include_mean, kernel=None):
        y = np.asarray(y, dtype=np.float64)
        alpha = self._get_alpha(y)
        log_prob = self._compute_log_prob(alpha)
        alpha = self.solver.solve_triangular(alpha, transpose=True)  # gp.py:334
        if X_test is None:
            if kernel is None:
                delta = self.noise @ alpha  # gp.py:342-346
                mean_value = y - delta
                if not include_mean:
                    mean_value = mean_value - self.loc
            else:
                mean_value = kernel.matmul(self.X, y=alpha)
                if include_mean:
                    mean_value = mean_value + self.loc
        else:
            if kernel is None:
                kernel = self.kernel
            mean_value = kernel.matmul(X_test, self.X, alpha)  # gp.py:357
            if include_mean:
                mean_value = mean_value + self.mean_function.vmap(X_test)
        return alpha, log_prob, mean_value


class ConditionResult(NamedTuple):
    """gp.py:364-385"""
    log_probability: Any
    gp: GaussianProcess


def _as_full_diag(diag, shape):
    """jnp.broadcast_to(diag, shape) of gp.py:98; an array that already has the full shape is passed through as it is (no
    copy: a caller's page-locked buffer stays page-locked all the way to the device copy)"""
    d = np.asarray(diag, dtype=np.float64)
    if d.shape == tuple(shape) and d.flags.c_contiguous:
        return d
    return np.broadcast_to(d, shape).copy()


def _default_diag(reference):  # gp.py:388-393
    return np.sqrt(np.finfo(np.asarray(reference).dtype).eps)
