"""Quasiseparable matrices on the device: the seven classes of src/tinygp/solvers/quasisep/core.py over `b200gp_qsm_*`.

Every object holds a handle to device-resident generator arrays (``d`` (n,), ``p``, ``q`` (n, m), ``a`` (n, m, m)); the
operations of core.py / ops.py -- ``@`` between QSMs (`qsm_mul`, ops.py:52-214), ``+``, ``-``, element-wise ``*``,
``transpose``, ``scale``, ``inv`` (core.py:310-317, ops.py:403-460), ``gram`` (core.py:424-434), ``cholesky``
(core.py:522-537), ``solve`` and the dense products -- run as O(n m^3) chunked scans in libb200gp.so and return new
handles; parts (``.diag``, ``.lower``, ``.upper``) and transposes share the device arrays.  ``.d/.p/.q/.a`` download.
There is no CPU fallback.
"""

from __future__ import annotations

__all__ = ["QSM", "DiagQSM", "StrictLowerTriQSM", "StrictUpperTriQSM", "LowerTriQSM", "UpperTriQSM", "SquareQSM", "SymmQSM"]

from ctypes import byref, c_int, c_int64, c_void_p, c_double

import numpy as np

from tinygp_b200 import _cabi

DIAG, STRICT_LOWER, STRICT_UPPER, LOWER, UPPER, SQUARE, SYMM = range(7)


def _backend():
    """the C-ABI context (tests substitute the host build of the same device source)"""
    return _cabi.get_context()


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(x):
    return None if x is None else c_void_p(x.ctypes.data)


class QSM:
    """Base class (core.py:47-131): a handle + the blanket operators."""

    __array_priority__ = 2000
    _kind = None

    def __init__(self, *args, **kwargs):
        pass          # construction happens in __new__ of the concrete classes (the handle decides the class)

    # ---- handle management --------------------------------------------------------------------------------------
    @classmethod
    def _wrap(cls, ctx, handle):
        n, kind, ml, mu = c_int64(), c_int(), c_int(), c_int()
        ctx.check(ctx.lib.b200gp_qsm_info(handle, byref(n), byref(kind), byref(ml), byref(mu)))
        obj = object.__new__(_CLASS_OF_KIND[kind.value])
        obj._ctx, obj._h, obj._n, obj._ml, obj._mu = ctx, handle, n.value, ml.value, mu.value
        return obj

    @classmethod
    def _create(cls, kind, d=None, lower=None, upper=None):
        ctx = _backend()
        arrs = []
        n = None
        if d is not None:
            d = _f64(d)
            if d.ndim != 1:
                raise ValueError("the diagonal must be one-dimensional")
            n = d.shape[0]
        ms = [0, 0]
        for idx, part in enumerate((lower, upper)):
            if part is None:
                arrs += [None, None, None]
                continue
            p, q, a = (_f64(v) for v in part)
            if p.ndim != 2 or q.shape != p.shape or a.shape != (p.shape[0], p.shape[1], p.shape[1]):
                raise ValueError("generators must have shapes p, q: (n, m), a: (n, m, m)")
            if n is not None and p.shape[0] != n:
                raise ValueError("dimension mismatch")
            n = p.shape[0]
            ms[idx] = p.shape[1]
            arrs += [p, q, a]
        h = c_void_p()
        ctx.check(ctx.lib.b200gp_qsm_create(ctx.handle, n, kind, ms[0], ms[1], _p(d), *[_p(x) for x in arrs], byref(h)))
        return QSM._wrap(ctx, h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._ctx.lib.b200gp_qsm_free(h)
            except Exception:
                pass
            self._h = None

    def _unary(self, fn, *extra):
        out = c_void_p()
        self._ctx.check(fn(self._h, *extra, byref(out)))
        return QSM._wrap(self._ctx, out)

    def _binary(self, fn, other):
        out = c_void_p()
        self._ctx.check(fn(self._h, other._h, byref(out)))
        return QSM._wrap(self._ctx, out)

    def _part_impl(self, which):
        out = c_void_p()
        self._ctx.check(self._ctx.lib.b200gp_qsm_part(self._h, which, byref(out)))
        return QSM._wrap(self._ctx, out)

    def _get(self, **want):
        """download the named generator arrays: d, lp, lq, la, up, uq, ua"""
        n, ml, mu = self._n, self._ml, self._mu
        shapes = {"d": (n,), "lp": (n, ml), "lq": (n, ml), "la": (n, ml, ml), "up": (n, mu), "uq": (n, mu), "ua": (n, mu, mu)}
        bufs = {k: (np.empty(shapes[k]) if want.get(k) else None) for k in shapes}
        self._ctx.check(self._ctx.lib.b200gp_qsm_get(self._h, *[_p(bufs[k]) for k in ("d", "lp", "lq", "la", "up", "uq", "ua")]))
        return bufs

    # ---- core.py:58-131 ---------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self._n, self._n)

    def transpose(self):
        return self._unary(self._ctx.lib.b200gp_qsm_transpose)

    @property
    def T(self):
        return self.transpose()

    def matmul(self, x, *, parallel: bool = False):
        """dense product (core.py:62-73; ``parallel`` accepted: the device scans are always the chunked form)"""
        x = np.asarray(x, dtype=np.float64)
        if x.shape[0] != self._n:
            raise ValueError("dimension mismatch")
        buf = np.array(x.reshape(self._n, -1), dtype=np.float64, order="C", copy=True)   # core.py:35-44
        self._ctx.check(self._ctx.lib.b200gp_qsm_matmul(self._h, _p(buf), buf.shape[1]))
        return buf.reshape(x.shape)

    def scale(self, other):
        c = np.asarray(other, dtype=np.float64)
        if c.ndim > 1:
            raise ValueError("scale takes a scalar or a vector")
        if c.ndim == 1 and c.shape[0] != self._n:
            raise ValueError("dimension mismatch")
        cc = _f64(c.reshape(-1))
        return self._unary(self._ctx.lib.b200gp_qsm_scale, _p(cc), int(c.ndim == 1))

    def to_dense(self):
        """core.py:84-90 (testing only: O(n^2 m))"""
        return self.matmul(np.eye(self._n))

    def __neg__(self):
        return self._unary(self._ctx.lib.b200gp_qsm_neg)

    def __add__(self, other):
        if not isinstance(other, QSM):
            return NotImplemented
        from tinygp_b200.solvers.quasisep.ops import elementwise_add
        return elementwise_add(self, other)

    def __sub__(self, other):
        return self.__add__(-other)                                   # core.py:104-105

    def __mul__(self, other):
        if isinstance(other, QSM):
            from tinygp_b200.solvers.quasisep.ops import elementwise_mul
            return elementwise_mul(self, other)
        assert np.ndim(other) <= 1
        return self.scale(other)

    def __rmul__(self, other):
        assert not isinstance(other, QSM)
        assert np.ndim(other) <= 1
        return self.scale(other)

    def __matmul__(self, other):
        if isinstance(other, QSM):
            from tinygp_b200.solvers.quasisep.ops import qsm_mul
            return qsm_mul(self, other)
        return self.matmul(other)

    def __rmatmul__(self, other):
        assert not isinstance(other, QSM)
        return (self.transpose() @ np.asarray(other).transpose()).transpose()


class DiagQSM(QSM):
    """core.py:134-165"""

    def __new__(cls, d):
        return QSM._create(DIAG, d=d)

    @property
    def d(self):
        return self._get(d=True)["d"]

    @property
    def shape(self):
        return (self._n, self._n)

    def self_add(self, other):
        return self + other

    def self_mul(self, other):
        return self * other


class _StrictTri(QSM):
    def self_add(self, other):
        return self + other

    def self_mul(self, other):
        return self * other


class StrictLowerTriQSM(_StrictTri):
    """core.py:168-236: M[i, j] = p[i] . a[i-1] ... a[j+1] . q[j] for i > j"""

    def __new__(cls, p, q, a):
        return QSM._create(STRICT_LOWER, lower=(p, q, a))

    p = property(lambda self: self._get(lp=True)["lp"])
    q = property(lambda self: self._get(lq=True)["lq"])
    a = property(lambda self: self._get(la=True)["la"])

    def __iter__(self):
        g = self._get(lp=True, lq=True, la=True)
        return iter((g["lp"], g["lq"], g["la"]))


class StrictUpperTriQSM(_StrictTri):
    """core.py:239-292: the transpose of the StrictLowerTriQSM with the same (p, q, a)"""

    def __new__(cls, p, q, a):
        return QSM._create(STRICT_UPPER, upper=(p, q, a))

    p = property(lambda self: self._get(up=True)["up"])
    q = property(lambda self: self._get(uq=True)["uq"])
    a = property(lambda self: self._get(ua=True)["ua"])

    def __iter__(self):
        g = self._get(up=True, uq=True, ua=True)
        return iter((g["up"], g["uq"], g["ua"]))


def _compose(diag, lower, upper, symm):
    ref = diag if diag is not None else (lower if lower is not None else upper)
    ctx = ref._ctx
    out = c_void_p()
    ctx.check(ctx.lib.b200gp_qsm_compose(diag._h if diag is not None else None, lower._h if lower is not None else None,
                                         upper._h if upper is not None else None, int(symm), byref(out)))
    return QSM._wrap(ctx, out)


class _WithDiag(QSM):
    @property
    def diag(self):
        return self._part_impl(0)


class LowerTriQSM(_WithDiag):
    """core.py:295-345"""

    def __new__(cls, diag, lower):
        return _compose(diag, lower, None, False)

    @property
    def lower(self):
        return self._part_impl(1)

    def inv(self):
        return self._unary(self._ctx.lib.b200gp_qsm_inv)              # core.py:310-317

    def solve(self, y, *, parallel: bool = False):                   # core.py:319-336
        return _solve(self, y)

    def __iter__(self):
        return iter((self.diag, self.lower))


class UpperTriQSM(_WithDiag):
    """core.py:348-393"""

    def __new__(cls, diag, upper):
        return _compose(diag, None, upper, False)

    @property
    def upper(self):
        return self._part_impl(2)

    def inv(self):
        return self._unary(self._ctx.lib.b200gp_qsm_inv)              # core.py:362-363

    def solve(self, y, *, parallel: bool = False):                   # core.py:366-383
        return _solve(self, y)

    def __iter__(self):
        return iter((self.diag, self.upper))


class SquareQSM(_WithDiag):
    """core.py:396-481"""

    def __new__(cls, diag, lower, upper):
        return _compose(diag, lower, upper, False)

    @property
    def lower(self):
        return self._part_impl(1)

    @property
    def upper(self):
        return self._part_impl(2)

    def gram(self):
        return self._unary(self._ctx.lib.b200gp_qsm_gram)             # core.py:424-434

    def inv(self):
        return self._unary(self._ctx.lib.b200gp_qsm_inv)              # core.py:436-478 (sequential on the device)

    def __iter__(self):
        return iter((self.diag, self.lower, self.upper))


class SymmQSM(_WithDiag):
    """core.py:484-540"""

    def __new__(cls, diag, lower):
        return _compose(diag, lower, None, True)

    @property
    def lower(self):
        return self._part_impl(1)

    def inv(self, *, parallel: bool = False):
        return self._unary(self._ctx.lib.b200gp_qsm_inv)              # core.py:507-520 / ops.py:403-460

    def cholesky(self, *, parallel: bool = False):
        """core.py:522-537.  A non-positive pivot gives NaNs from there on, as in the reference; the 1-based index
        of the first one is kept on the result as ``.info`` (0 = none)."""
        out, info = c_void_p(), c_int64()
        self._ctx.check(self._ctx.lib.b200gp_qsm_cholesky(self._h, byref(out), byref(info)))
        r = QSM._wrap(self._ctx, out)
        r.info = info.value
        return r

    def gram(self):
        return self._unary(self._ctx.lib.b200gp_qsm_gram)

    def __iter__(self):
        return iter((self.diag, self.lower))


def _solve(mat, y):
    y = np.asarray(y, dtype=np.float64)
    if y.shape[0] != mat._n:
        raise ValueError("dimension mismatch")
    buf = np.array(y.reshape(mat._n, -1), dtype=np.float64, order="C", copy=True)
    mat._ctx.check(mat._ctx.lib.b200gp_qsm_solve(mat._h, _p(buf), buf.shape[1]))
    return buf.reshape(y.shape)


def sum_log_diag(mat) -> float:
    out = c_double()
    mat._ctx.check(mat._ctx.lib.b200gp_qsm_sum_log_diag(mat._h, byref(out)))
    return out.value


_CLASS_OF_KIND = {DIAG: DiagQSM, STRICT_LOWER: StrictLowerTriQSM, STRICT_UPPER: StrictUpperTriQSM, LOWER: LowerTriQSM,
                  UPPER: UpperTriQSM, SQUARE: SquareQSM, SYMM: SymmQSM}
