"""B200 QuasisepSolver: the contract of src/tinygp/solvers/quasisep/solver.py:19-139 over the chunked
scan kernels of libb200gp.so.  The factor (c, w) lives in HBM for the object's lifetime."""

from __future__ import annotations

__all__ = ["QuasisepSolver"]

from ctypes import byref, c_double, c_int, c_void_p
from typing import Any

import numpy as np

from tinygp_b200 import _cabi
from tinygp_b200.kernels.quasisep import Quasisep
from tinygp_b200.noise import Diagonal
from tinygp_b200.solvers.quasisep import core as qcore
from tinygp_b200.solvers.solver import ConditionedCovariance, Solver

_UNSORTED_MSG = "Input coordinates must be sorted in order to use the QuasisepSolver"  # solver.py:142-146


class QuasisepSolver(Solver):
    def __init__(self, kernel, X, noise, *, covariance: Any | None = None, assume_sorted: bool = False,
                 parallel: bool = False):
        """``parallel`` is accepted for API compatibility (solver.py:33,60-64); the device scans are
        always the chunked parallel form, and give the sequential recursion's values.

        ``covariance`` (solver.py:75-78): a precomputed ``core.SymmQSM`` -- what ``condition`` returns for a
        quasiseparable predictive kernel at the inputs -- is factored as it is (generator arrays of any order, qsm.cu);
        otherwise the kernel's state-space model is factored by the model-specialised scans (quasisep.cu / qs_fast.cu)."""
        self._ctx = _cabi.get_context()
        self._h = c_void_p()
        self.kernel, self.noise, self.parallel = kernel, noise, parallel
        self._matrix = self._factor = None
        if covariance is None and isinstance(kernel, Quasisep) and not (isinstance(noise, Diagonal) and kernel._on_device()):
            # solver.py:73-74 as written -- the kernel's SymmQSM plus the noise's, factored from generator arrays -- for
            #  * noise.Banded (noise.py:226-240): a device SymmQSM of order J + J_band;
            #  * kernels without a device model (a user-defined state-space subclass, or more than 8 states): their
            #    generators are evaluated by the kernel's own Python methods on the host and uploaded (any order J)
            t = _cabi.f64(kernel._sortable(X))
            if t.ndim != 1:
                raise ValueError("QuasisepSolver takes 1-D sortable coordinates")
            if not assume_sorted:
                unsorted = c_int(0)
                self._ctx.check(self._ctx.lib.b200gp_qs_check_sorted(self._ctx.handle, _cabi.ptr(t), t.shape[0], byref(unsorted)))
                if unsorted.value:
                    raise ValueError(_UNSORTED_MSG)
            if np.shape(noise.diagonal()) != t.shape:
                raise ValueError("noise diagonal must have shape (N,)")
            covariance = kernel.to_symm_qsm(X) + noise.to_qsm()
        if covariance is not None:
            if not isinstance(covariance, qcore.SymmQSM):
                raise ValueError("QuasisepSolver(covariance=...) takes a tinygp_b200.solvers.quasisep.core.SymmQSM")
            self.X = X
            self._n = covariance.shape[0]
            self._matrix = covariance
            self._factor = covariance.cholesky()                              # solver.py:82
            self.info = int(self._factor.info)
            self._J = covariance._ml
            # model-specific shortcuts of the kernel-built solver do not exist for generator arrays
            self.whitened_sumsq = self.conditioned_variance = self.inverse_diagonal = None
            return
        if not isinstance(kernel, Quasisep):
            raise ValueError("QuasisepSolver requires a tinygp_b200.kernels.quasisep.Quasisep kernel")
        t = _cabi.f64(kernel._sortable(X))
        if t.ndim != 1:
            raise ValueError("QuasisepSolver takes 1-D sortable coordinates")
        self.X = X                     # as given (solver.py:66); the kernel's coord_to_sortable is applied where it is used
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

    @property
    def _generic(self) -> bool:
        """built from a precomputed SymmQSM (generator arrays) rather than from a kernel's state-space model"""
        return not self._h

    @property
    def matrix(self):
        """solver.py:81: the covariance incl. noise as a SymmQSM (device generators)"""
        if self._matrix is None:
            self._matrix = self.kernel.to_symm_qsm(self.X) + self.noise.to_qsm()      # solver.py:73-74
        return self._matrix

    @property
    def factor(self):
        """solver.py:82: LowerTriQSM(diag = c, lower = (p, w, a)) on the device"""
        if self._factor is None:
            out = c_void_p()
            self._ctx.check(self._ctx.lib.b200gp_qs_factor_qsm(self._h, byref(out)))
            self._factor = qcore.QSM._wrap(self._ctx, out)
        return self._factor

    # -- Solver contract ------------------------------------------------------------------
    def variance(self):  # solver.py:84-85
        if self._generic:
            return self._matrix.diag.d
        out = np.empty(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_variance(self._h, _cabi.ptr(out)))
        return out

    def covariance(self):  # solver.py:87-88: to_dense() = matmul with the identity (core.py:84-90)
        if self._generic:
            return self._matrix.to_dense()
        eye = np.eye(self._n)
        self._ctx.check(self._ctx.lib.b200gp_qs_matmul(self._h, _cabi.ptr(eye), self._n))
        return eye

    def normalization(self):  # solver.py:90-93
        if self._generic:
            if self.info != 0:
                return np.nan
            return qcore.sum_log_diag(self._factor) + 0.5 * self._n * np.log(2 * np.pi)
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
        if self._generic:
            return (self._factor.transpose() if transpose else self._factor).solve(y)
        return self._apply(self._ctx.lib.b200gp_qs_solve_triangular, y, int(bool(transpose)))

    def whitened_sumsq(self, y):
        """``sum(solve_triangular(y) ** 2)`` (the data term of gp.py:313-316) reduced on the device: at N = 10^7 the
        N-vector ``alpha`` is neither copied back nor squared on the host.  Optional hook read by GaussianProcess."""
        y = _cabi.f64(y)
        if y.shape != (self._n,):
            raise ValueError("dimension mismatch")
        if self._generic:
            return float(np.sum(np.square(self._factor.solve(y))))
        out = c_double()
        self._ctx.check(self._ctx.lib.b200gp_qs_solve_sumsq(self._h, _cabi.ptr(y), byref(out)))
        return out.value

    def dot_triangular(self, y):  # solver.py:101-102
        if self._generic:
            return self._factor.matmul(y)
        return self._apply(self._ctx.lib.b200gp_qs_dot_triangular, y)

    def matmul(self, y):
        """covariance @ y without densifying (core.py:499-505)."""
        if self._generic:
            return self._matrix.matmul(y)
        return self._apply(self._ctx.lib.b200gp_qs_matmul, y)

    def factor_arrays(self):
        """(c, w) of the LowerTriQSM factor (core.py:524-539) as host arrays."""
        if self._generic:
            return self._factor.diag.d, self._factor.lower.q
        c, w = np.empty(self._n), np.empty((self._n, self._J))
        self._ctx.check(self._ctx.lib.b200gp_qs_get_factor(self._h, _cabi.ptr(c), _cabi.ptr(w)))
        return c, w

    def generators(self):
        """(d, p, q, a) of the SymmQSM incl. the noise diagonal (kernels/quasisep.py:102-116)."""
        n, J = self._n, self._J
        if self._generic:
            lo = self._matrix.lower
            return self._matrix.diag.d, lo.p, lo.q, lo.a
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

    def _is_own_kernel(self, kernel) -> bool:
        if kernel is self.kernel:
            return True
        try:
            a, b = kernel.component_array(), self.kernel.component_array()
        except NotImplementedError:
            return False
        return a.shape == b.shape and bool(np.all(a == b))

    def condition(self, kernel, X_test, noise) -> Any:
        """solver.py:104-139.

        *QSM branch* (:124-129; ``X_test is None`` and a quasiseparable predictive kernel): ``M - (factor.inv() @ M).gram()``
        with ``M = kernel.to_symm_qsm(X) + noise`` as a ``core.SymmQSM`` of order up to 4J whose generators stay on the
        device -- O(N), nothing densified; a GaussianProcess built on it factors it with QuasisepSolver(covariance=...).
        When the predictive kernel is the solver's own kernel the same matrix is returned in its order-J form (below).

        *Dense branch* (:131-139): ``Kss - A^T A`` with ``A = factor.solve(Ks)`` computed entirely on the device by
        ``b200gp_qs_condition`` (build kernel for ``Ks^T`` from the predictive kernel's program, one forward-substitution
        scan per test point, NT GEMM on the tensor pipe with ``k(X*, X*)`` generated in its epilogue).  The reference
        adds the predictive noise in the QSM branch and not in the dense one; the dense result is tagged accordingly
        (``ConditionedCovariance.includes_noise``)."""
        if X_test is None and isinstance(kernel, Quasisep):                       # solver.py:124-129
            if not self._generic and self._is_own_kernel(kernel):
                # Predictive kernel = training kernel: with Sigma = K + N and M = K,  K - K Sigma^-1 K = N - N Sigma^-1 N,
                # so the same matrix is  diag(noise* + N) - diag(N) Sigma^-1 diag(N)  with Sigma^-1 = symm_inv (ops.py:403-460)
                # of order J -- instead of the order-4J difference of two large, almost cancelling terms that the four
                # reference lines below build (same dense values to rounding; 64x fewer flops per point at J = 4, and a
                # minimal realisation, whose Cholesky carry is small and contracting).
                n = _cabi.f64(self.noise.diagonal())
                Sinv = self.matrix.inv()
                lower = Sinv.lower.scale(-n).transpose().scale(n).transpose()      # p <- -N p (rows), q <- q N (columns)
                diag = Sinv.diag.scale(-(n * n)) + (noise.to_qsm() + self.noise.to_qsm())
                return qcore.SymmQSM(diag=diag, lower=lower)
            M = kernel.to_symm_qsm(self.X)
            if M.shape[0] != self._n:
                raise ValueError("dimension mismatch")
            delta = (self.factor.inv() @ M).gram()
            M = M + noise.to_qsm()
            return M - delta
        if self._generic:
            # generator arrays (noise.Banded, or a conditioned process): solver.py:131-139 call by call -- Ks and Kss from the
            # build kernel, A = factor.solve(Ks) by the QSM scans, Kss - A^T A from the fp64 GEMM (b200gp_gram_downdate)
            Xs = self.X if X_test is None else X_test
            A = self._factor.solve(kernel(self.X, Xs))
            out = np.ascontiguousarray(kernel(Xs, Xs))
            At = np.ascontiguousarray(A.T)
            self._ctx.check(self._ctx.lib.b200gp_gram_downdate(self._ctx.handle, _cabi.ptr(At), At.shape[0], At.shape[1],
                                                               _cabi.ptr(out)))
            return ConditionedCovariance.tag(out, False)
        if X_test is None:
            prog, x = kernel.lower_for(kernel._sortable(self.X) if isinstance(kernel, Quasisep) else self.X)
            xt_ptr, m = None, self._n
        else:
            xt = np.asarray(kernel._sortable(X_test) if isinstance(kernel, Quasisep) else X_test,
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
                                                          None, _cabi.ptr(out)))
        return ConditionedCovariance.tag(out, False)
