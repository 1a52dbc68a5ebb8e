"""`jax.numpy` on NumPy (fp64, eager).  Arrays are an ndarray subclass that adds `.at[...]` and
`.block_until_ready()`; everything else falls through to NumPy."""

import sys
import types

import numpy as _np


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _apply(self, op, value):
        out = _np.array(self.arr, copy=True)
        op(out, self.idx, value)
        return out.view(Array)

    def set(self, value):
        def op(o, i, v):
            o[i] = v
        return self._apply(op, value)

    def add(self, value):
        return self._apply(lambda o, i, v: _np.add.at(o, i, v), value)

    def multiply(self, value):
        return self._apply(lambda o, i, v: _np.multiply.at(o, i, v), value)

    def get(self):
        return _cast(self.arr[self.idx])


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


def _defers(other):
    # a jax array returns NotImplemented for a non-array operand, so `array * kernel` reaches Kernel.__rmul__
    return getattr(type(other), "__is_shim_module__", False)


class Array(_np.ndarray):
    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def _install_deferring_binops():
    for name in ("add", "sub", "mul", "truediv", "matmul", "pow"):
        for dunder in (f"__{name}__", f"__r{name}__"):
            base = getattr(_np.ndarray, dunder)

            def op(self, other, _base=base):
                if _defers(other):
                    return NotImplemented
                return _base(self, other)

            op.__name__ = dunder
            setattr(Array, dunder, op)


_install_deferring_binops()


def _cast(x):
    if isinstance(x, Array):
        return x
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, tuple):
        return tuple(_cast(v) for v in x)
    if isinstance(x, list):
        return [_cast(v) for v in x]
    return x


def _wrap(f):
    def g(*args, **kwargs):
        return _cast(f(*args, **kwargs))

    g.__name__ = getattr(f, "__name__", "f")
    g.__doc__ = getattr(f, "__doc__", None)
    return g


def _chol(a, upper=False):
    a = _np.asarray(a, dtype=float)
    try:
        L = _np.linalg.cholesky(a)
    except _np.linalg.LinAlgError:
        return _np.full_like(a, _np.nan).view(Array)  # JAX yields NaNs, it does not raise
    return _cast(_np.swapaxes(L, -1, -2) if upper else L)


linalg = types.ModuleType("jax.numpy.linalg")
for _name in dir(_np.linalg):
    _obj = getattr(_np.linalg, _name)
    if callable(_obj) and not _name.startswith("_") and not isinstance(_obj, type):
        setattr(linalg, _name, _wrap(_obj))
linalg.cholesky = _chol
linalg.LinAlgError = _np.linalg.LinAlgError
sys.modules["jax.numpy.linalg"] = linalg

ndarray = _np.ndarray
_passthrough_types = (type, types.ModuleType)


def __getattr__(name):
    obj = getattr(_np, name)
    if callable(obj) and not isinstance(obj, _passthrough_types) and not isinstance(obj, _np.ufunc):
        return _wrap(obj)
    if isinstance(obj, _np.ufunc):
        return _wrap(obj)
    return obj


def array(x, dtype=None, copy=True, **kw):
    return _cast(_np.array(x, dtype=dtype, copy=copy, **kw))


def asarray(x, dtype=None, **kw):
    return _cast(_np.asarray(x, dtype=dtype, **kw))


def roots(p, strip_zeros=True):
    p = _np.asarray(p)
    if not strip_zeros:
        # jax: no stripping of leading/trailing zeros; tinygp only calls it with a non-zero leading coefficient
        return _cast(_np.roots(p).astype(complex))
    return _cast(_np.roots(p))


def finfo(x):
    if isinstance(x, (_np.ndarray, _np.generic)):
        return _np.finfo(x.dtype)
    if isinstance(x, (float, int)):
        return _np.finfo(_np.float64)
    return _np.finfo(x)
