"""B200 QuasisepSolver: the contract of src/tinygp/solvers/quasisep/solver.py:19-139 over the chunked
scan kernels of libb200gp.so.  The factor (c, w) lives in HBM for the object's lifetime."""

from __future__ import annotations

__all__ = ["QuasisepSolver"]

from ctypes import byref, c_double, c_int, c_void_p
from typing import Any

import numpy as np

from tinygp_b200 import _cabi
from tinygp_b200.kernels.quasisep import Quasisep
from tinygp_b200.solvers.solver import ConditionedCovariance, Solver

_UNSORTED_MSG = "Input coordinates must be sorted in order to use the QuasisepSolver"  # solver.py:142-146


class QuasisepSolver(Solver):
    def __init__(self, kernel, X, noise, *, covariance: Any | None = None, assume_sorted: bool = False,
                 parallel: bool = False):
        """``parallel`` is accepted for API compatibility (solver.py:33,60-64); the device scans are
        always the chunked parallel form, and give the sequential recursion's values."""
        if covariance is not None:
            raise NotImplementedError("QuasisepSolver(covariance=SymmQSM) is unsupported by the B200 backend")
        if not isinstance(kernel, Quasisep):
            raise ValueError("QuasisepSolver requires a tinygp_b200.kernels.quasisep.Quasisep kernel")
        self._ctx = _cabi.get_context()
        self._h = c_void_p()
        self.kernel, self.noise, self.parallel = kernel, noise, parallel
        t = _cabi.f64(kernel.coord_to_sortable(X))
        if t.ndim != 1:
            raise ValueError("QuasisepSolver takes 1-D sortable coordinates")
        self.X = t
        self._n = t.shape[0]
        diag = _cabi.f64(noise.diagonal())
        if diag.shape != t.shape:
            raise ValueError("noise diagonal must have shape (N,)")
        comps = kernel.component_array()
        unsorted, info = c_int(0), c_int(0)
        lib = self._ctx.lib
        self._ctx.check(lib.b200gp_qs_create(self._ctx.handle, _cabi.ptr(comps), comps.shape[0], _cabi.ptr(t),
                                             self._n, _cabi.ptr(diag), int(bool(assume_sorted)), byref(self._h),
                                             byref(unsorted), byref(info)))
        if unsorted.value:
            raise ValueError(_UNSORTED_MSG)
        self.info = info.value
        J = c_int(0)
        self._ctx.check(lib.b200gp_qs_state_dim(self._h, byref(J)))
        self._J = J.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._ctx.lib.b200gp_qs_free(h)
            except Exception:
                pass
            self._h = c_void_p()

    # -- Solver contract ------------------------------------------------------------------
    def variance(self):  # solver.py:84-85
        out = np.empty(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_variance(self._h, _cabi.ptr(out)))
        return out

    def covariance(self):  # solver.py:87-88: to_dense() = matmul with the identity (core.py:84-90)
        eye = np.eye(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_matmul(self._h, _cabi.ptr(eye), self._n))
        return eye

    def normalization(self):  # solver.py:90-93
        ld = c_double()
        self._ctx.check(self._ctx.lib.b200gp_qs_logdet_half(self._h, byref(ld)))
        if getattr(self, "info", 0) != 0:     # failed factorisation: NaN like the reference's log of a NaN pivot
            return np.nan
        return ld.value + 0.5 * self._n * np.log(2 * np.pi)

    def _apply(self, fn, y, *extra):
        y = np.asarray(y, dtype=np.float64)
        if y.shape[0] != self._n:
            raise ValueError("dimension mismatch")
        buf = np.array(y.reshape(self._n, -1), dtype=np.float64, order="C", copy=True)  # core.py:35-44
        self._ctx.check(fn(self._h, _cabi.ptr(buf), buf.shape[1], *extra))
        return buf.reshape(y.shape)

    def solve_triangular(self, y, *, transpose: bool = False):  # solver.py:95-99
        return self._apply(self._ctx.lib.b200gp_qs_solve_triangular, y, int(bool(transpose)))

    def whitened_sumsq(self, y):
        """``sum(solve_triangular(y) ** 2)`` (the data term of gp.py:313-316) reduced on the device: at N = 10^7 the
        N-vector ``alpha`` is neither copied back nor squared on the host.  Optional hook read by GaussianProcess."""
        y = _cabi.f64(y)
        if y.shape != (self._n,):
            raise ValueError("dimension mismatch")
        out = c_double()
        self._ctx.check(self._ctx.lib.b200gp_qs_solve_sumsq(self._h, _cabi.ptr(y), byref(out)))
        return out.value

    def dot_triangular(self, y):  # solver.py:101-102
        return self._apply(self._ctx.lib.b200gp_qs_dot_triangular, y)

    def matmul(self, y):
        """covariance @ y without densifying (core.py:499-505)."""
        return self._apply(self._ctx.lib.b200gp_qs_matmul, y)

    def factor(self):
        """(c, w) of the LowerTriQSM factor (core.py:524-539)."""
        c, w = np.empty(self._n), np.empty((self._n, self._J))
        self._ctx.check(self._ctx.lib.b200gp_qs_get_factor(self._h, _cabi.ptr(c), _cabi.ptr(w)))
        return c, w

    def generators(self):
        """(d, p, q, a) of the SymmQSM incl. the noise diagonal (kernels/quasisep.py:102-116)."""
        n, J = self._n, self._J
        d, p, q, a = np.empty(n), np.empty((n, J)), np.empty((n, J)), np.empty((n, J, J))
        self._ctx.check(self._ctx.lib.b200gp_qs_get_generators(self._h, _cabi.ptr(d), _cabi.ptr(p), _cabi.ptr(q),
                                                               _cabi.ptr(a)))
        return d, p, q, a

    def inverse_diagonal(self):
        """diag((K + N)^-1): the diagonal of ``factor.inv().gram()`` (core.py:310-317, 424-434) by one backward scan."""
        out = np.empty(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_inverse_diagonal(self._h, _cabi.ptr(out)))
        return out

    def conditioned_variance(self, noise):
        """Variance of the conditioned process at the inputs for this solver's kernel -- the diagonal of
        solver.py:124-129 as solver.py:84-85 reads it -- in O(N) on the device (no N x N matrix)."""
        diag = _cabi.f64(noise.diagonal())
        if diag.shape != (self._n,):
            raise ValueError("noise diagonal must match the number of predicted points")
        out = np.empty(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_conditioned_variance(self._h, _cabi.ptr(diag), _cabi.ptr(out)))
        return out

    def condition(self, kernel, X_test, noise) -> Any:
        """solver.py:104-139: ``Kss [+ noise] - A^T A`` with ``A = factor.solve(Ks)`` -- computed entirely on the
        device by ``b200gp_qs_condition`` (build kernel for ``Ks^T`` from the predictive kernel's program, one
        forward-substitution scan per test point, NT GEMM on the tensor pipe with ``k(X*, X*)`` generated in its
        epilogue).  The reference adds the predictive noise in its QSM branch (:124-129: ``X_test is None`` with a
        quasiseparable kernel) and not in its dense branch (:131-139); the result is tagged accordingly
        (``ConditionedCovariance.includes_noise``).  The QSM branch's *values* are returned as a dense matrix: a
        QSM-valued result needs qsm_mul / inv / gram on the device and is a 'next' row."""
        diag_ptr, with_noise = None, False
        if X_test is None:
            prog, x = kernel.lower_for(self.X)
            xt_ptr, m = None, self._n
            if isinstance(kernel, Quasisep):                                   # solver.py:124-129
                diag = _cabi.f64(noise.diagonal())
                if diag.shape != (m,):
                    raise ValueError("noise diagonal must match the number of predicted points")
                diag_ptr, with_noise = _cabi.ptr(diag), True
        else:
            xt = np.asarray(kernel.coord_to_sortable(X_test) if hasattr(kernel, "coord_to_sortable") else X_test,
                            dtype=np.float64)
            if xt.ndim != 1:
                raise ValueError("QuasisepSolver.condition takes 1-D test coordinates")
            prog, x = kernel.lower_for(xt)
            xt_ptr, m = _cabi.ptr(x), x.shape[0]
        if x.shape[1] != 1:
            raise NotImplementedError("a predictive kernel with host-side Transform columns is unsupported by the "
                                      "B200 QuasisepSolver.condition")
        out = np.empty((m, m))
        self._ctx.check(self._ctx.lib.b200gp_qs_condition(self._h, _cabi.ptr(prog), prog.shape[0], xt_ptr, m,
                                                          diag_ptr, _cabi.ptr(out)))
        return ConditionedCovariance.tag(out, with_noise)
