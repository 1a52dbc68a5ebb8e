"""Rectangular quasiseparable matrices (src/tinygp/solvers/quasisep/general.py:42-106).

``Quasisep.to_general_qsm(X1, X2)`` (kernels/quasisep.py:118-145) -- the (n1, n2) cross-covariance between test points X1 and
the sorted training coordinates X2 -- is only ever multiplied into dense vectors by its callers (the predictive mean,
gp.py:357).  On the B200 that product is ONE entry point, ``b200gp_qs_kernel_matmul``: a forward and a backward state scan
over X2, then per row of X1 the bit-exact ``searchsorted(X2, x, side="right") - 1`` and two transition matrices, all with the
kernel's state-space model in registers -- the reference's generator arrays ``pl, ql, pu, qu, a, idx`` are never written to
memory.  This class is the handle to that product."""

from __future__ import annotations

__all__ = ["GeneralQSM"]

import numpy as np


class GeneralQSM:
    def __init__(self, kernel, X1, X2):
        self.kernel = kernel
        self._raw = (X1, X2)                      # kernel.matmul applies coord_to_sortable itself
        self.X1 = np.asarray(kernel._sortable(X1), dtype=np.float64)
        self.X2 = np.asarray(kernel._sortable(X2), dtype=np.float64)
        if self.X1.ndim != 1 or self.X2.ndim != 1:
            raise ValueError("quasiseparable kernels take 1-D sortable coordinates")

    @property
    def shape(self):
        return (self.X1.shape[0], self.X2.shape[0])

    @property
    def idx(self):
        """jnp.searchsorted(X2, X1, side="right") - 1 (kernels/quasisep.py:121), computed on the device, bit-exact"""
        from tinygp_b200 import _cabi
        ctx = _cabi.get_context()
        out = np.empty(self.X1.shape[0], dtype=np.int64)
        x2, x1 = _cabi.f64(self.X2), _cabi.f64(self.X1)
        ctx.check(ctx.lib.b200gp_searchsorted_right_m1(ctx.handle, _cabi.ptr(x2), x2.shape[0], _cabi.ptr(x1), x1.shape[0],
                                                       _cabi.ptr(out)))
        return out

    def matmul(self, x):
        """general.py:66-103: (n2, ...) -> (n1, ...)"""
        return self.kernel.matmul(self._raw[0], self._raw[1], x)

    def __matmul__(self, other):
        return self.matmul(other)
