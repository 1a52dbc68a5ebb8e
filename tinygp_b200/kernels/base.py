"""Kernel expression surface (reference: src/tinygp/kernels/base.py).

A kernel here is a plain parameter holder.  ``Kernel.__call__`` (base.py:84-103)
does not vmap a Python ``evaluate``; it lowers the expression to a postfix *kernel
program* (include/b200gp.h) and runs the pairwise-build CUDA kernel.
"""

from __future__ import annotations

__all__ = ["Kernel", "Conditioned", "Custom", "Sum", "Product", "Constant", "DotProduct", "Polynomial"]

import numpy as np

from tinygp_b200 import _cabi

OP_CONST, OP_ADD, OP_MUL = 0, 16, 17


def _as_coords(X):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        return np.ascontiguousarray(X[:, None]), True
    if X.ndim == 2:
        return np.ascontiguousarray(X), False
    raise ValueError("coordinates must have shape (N,) or (N, D); pytrees are not supported by the B200 backend")


class Kernel:
    """Base class (base.py:29-126)."""

    def lower(self) -> list[tuple[int, int, float, float]]:
        raise NotImplementedError(
            f"{type(self).__name__} cannot be lowered to a device kernel program: "
            "unsupported by the B200 solver backend (only stationary kernels and their sums/products)"
        )

    def program(self) -> np.ndarray:
        instr = self.lower()
        return np.ascontiguousarray(np.array(instr, dtype=np.float64).reshape(-1, _cabi.PROG_STRIDE))

    # -- evaluation ---------------------------------------------------------------------
    def evaluate(self, X1, X2):
        """Scalar evaluation at a single pair of coordinates (base.py:37-56)."""
        x1 = np.atleast_1d(np.asarray(X1, dtype=np.float64))[None, :]
        x2 = np.atleast_1d(np.asarray(X2, dtype=np.float64))[None, :]
        return self(x1, x2)[0, 0]

    def evaluate_diag(self, X):
        return self.evaluate(X, X)

    def __call__(self, X1, X2=None):
        ctx = _cabi.get_context()
        prog = self.program()
        x1, _ = _as_coords(X1)
        if X2 is None:  # base.py:85-93
            out = np.empty(x1.shape[0])
            ctx.check(ctx.lib.b200gp_kernel_diag(ctx.handle, _cabi.ptr(prog), prog.shape[0], _cabi.ptr(x1),
                                                 x1.shape[0], x1.shape[1], _cabi.ptr(out)))
            return out
        x2, _ = _as_coords(X2)
        if x1.shape[1] != x2.shape[1]:
            raise ValueError("X1 and X2 must have the same number of input dimensions")
        out = np.empty((x1.shape[0], x2.shape[0]))
        ctx.check(ctx.lib.b200gp_kernel_matrix(ctx.handle, _cabi.ptr(prog), prog.shape[0], _cabi.ptr(x1), x1.shape[0],
                                               _cabi.ptr(x2), x2.shape[0], x1.shape[1], _cabi.ptr(out)))
        return out

    def matmul(self, X1, X2=None, y=None):
        """base.py:68-82 -- k(X1, X2) @ y without materialising K for vector y."""
        if y is None:
            assert X2 is not None
            y = X2
            X2 = None
        if X2 is None:
            X2 = X1
        y = np.asarray(y, dtype=np.float64)
        if y.ndim != 1:
            return np.dot(self(X1, X2), y)
        ctx = _cabi.get_context()
        prog = self.program()
        x1, _ = _as_coords(X1)
        x2, _ = _as_coords(X2)
        yy = _cabi.f64(y)
        out = np.empty(x1.shape[0])
        ctx.check(ctx.lib.b200gp_kernel_matvec(ctx.handle, _cabi.ptr(prog), prog.shape[0], _cabi.ptr(x1), x1.shape[0],
                                               _cabi.ptr(x2), x2.shape[0], x1.shape[1], _cabi.ptr(yy),
                                               _cabi.ptr(out)))
        return out

    # -- algebra (base.py:105-126) --------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, Kernel):
            return Sum(self, other)
        return Sum(self, Constant(other))

    def __radd__(self, other):
        if not isinstance(other, Kernel) and np.ndim(other) == 0 and other == 0:
            return self
        if isinstance(other, Kernel):
            return Sum(other, self)
        return Sum(Constant(other), self)

    def __mul__(self, other):
        if isinstance(other, Kernel):
            return Product(self, other)
        return Product(self, Constant(other))

    def __rmul__(self, other):
        if isinstance(other, Kernel):
            return Product(other, self)
        return Product(Constant(other), self)


class Sum(Kernel):
    """base.py:170-177"""

    def __init__(self, kernel1: Kernel, kernel2: Kernel):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def lower(self):
        return self.kernel1.lower() + self.kernel2.lower() + [(OP_ADD, 0, 0.0, 0.0)]


class Product(Kernel):
    """base.py:180-187"""

    def __init__(self, kernel1: Kernel, kernel2: Kernel):
        self.kernel1, self.kernel2 = kernel1, kernel2

    def lower(self):
        return self.kernel1.lower() + self.kernel2.lower() + [(OP_MUL, 0, 0.0, 0.0)]


class Constant(Kernel):
    """base.py:190-209"""

    def __init__(self, value):
        if np.ndim(value) != 0:
            raise ValueError("The value of a constant kernel must be a scalar")
        self.value = float(value)

    def lower(self):
        return [(OP_CONST, 0, self.value, 0.0)]


class Conditioned(Kernel):
    """The covariance of a process conditioned on data (base.py:129-153), evaluated matrix-wise:
    k*(X1, X2) - [L^-1 k(X, X1)]^T [L^-1 k(X, X2)]."""

    def __init__(self, X, solver, kernel: Kernel):
        self.X, self.solver, self.kernel = X, solver, kernel

    def __call__(self, X1, X2=None):
        if X2 is None:
            K = self.solver.solve_triangular(self.kernel(self.X, X1))
            return self.kernel(X1) - np.sum(K * K, axis=0)
        K1 = self.solver.solve_triangular(self.kernel(self.X, X1))
        K2 = self.solver.solve_triangular(self.kernel(self.X, X2))
        return self.kernel(X1, X2) - K1.T @ K2

    def matmul(self, X1, X2=None, y=None):
        if y is None:
            y, X2 = X2, None
        if X2 is None:
            X2 = X1
        return self(X1, X2) @ np.asarray(y, dtype=np.float64)


class Custom(Kernel):
    """base.py:156-167 -- arbitrary Python callables cannot run inside the CUDA build kernel."""

    def __init__(self, function):
        self.function = function


class DotProduct(Kernel):
    """base.py:212-225 -- non-stationary: out of scope of the B200 backend (lower() raises)."""


class Polynomial(Kernel):
    """base.py:228-249 -- non-stationary: out of scope of the B200 backend (lower() raises)."""

    def __init__(self, order, scale=1.0, sigma=0.0):
        self.order, self.scale, self.sigma = order, scale, sigma
