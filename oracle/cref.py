"""ctypes wrapper of oracle/csrc/qs_oracle.c (TEST ORACLE / CPU baseline; see oracle/__init__.py)."""

import ctypes
import os

import numpy as np

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liboracle_qs.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} not built (python -c 'import __graft_entry__ as g; g.build()')")
_lib = ctypes.CDLL(_path)
_P = ctypes.c_void_p
_lib.qs_log_probability.restype = ctypes.c_double
_lib.qs_log_probability.argtypes = [ctypes.c_int64, ctypes.c_int, _P, _P, _P, _P, _P, _P, _P, _P]
_lib.qs_cholesky.restype = ctypes.c_int64
_lib.qs_cholesky.argtypes = [ctypes.c_int64, ctypes.c_int, _P, _P, _P, _P, _P, _P]


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def qs_cholesky(d, p, q, a):
    d, p, q, a = map(_c, (d, p, q, a))
    n, J = p.shape
    c, w = np.empty(n), np.empty((n, J))
    bad = _lib.qs_cholesky(n, J, d.ctypes.data, p.ctypes.data, q.ctypes.data, a.ctypes.data, c.ctypes.data,
                           w.ctypes.data)
    return c, w, int(bad)


def qs_log_probability(d, p, q, a, resid):
    d, p, q, a, resid = map(_c, (d, p, q, a, resid))
    n, J = p.shape
    c, w, al = np.empty(n), np.empty((n, J)), np.empty(n)
    return float(_lib.qs_log_probability(n, J, d.ctypes.data, p.ctypes.data, q.ctypes.data, a.ctypes.data,
                                         resid.ctypes.data, c.ctypes.data, w.ctypes.data, al.ctypes.data))
