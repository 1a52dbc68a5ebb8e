"""`jax.scipy.linalg` on SciPy/LAPACK."""

import numpy as _np
import scipy.linalg as _sl

from ..numpy import _cast


def cholesky(a, lower=False, **_kw):
    a = _np.asarray(a, dtype=float)
    try:
        return _cast(_sl.cholesky(a, lower=lower, check_finite=False))
    except _sl.LinAlgError:
        return _cast(_np.full_like(a, _np.nan))  # jax returns NaNs for a non-PD input (tinygp relies on it: gp.py:316)


def solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, **_kw):
    a = _np.asarray(a, dtype=float)
    b = _np.asarray(b, dtype=float)
    if not _np.all(_np.isfinite(a)) or not _np.all(_np.isfinite(b)):
        return _cast(_np.full_like(b, _np.nan))
    return _cast(_sl.solve_triangular(a, b, trans=trans, lower=lower, unit_diagonal=unit_diagonal, check_finite=False))


def block_diag(*arrs):
    return _cast(_sl.block_diag(*[_np.asarray(a) for a in arrs]))


def solve(a, b, **kw):
    return _cast(_sl.solve(_np.asarray(a), _np.asarray(b), **{k: v for k, v in kw.items() if k in ("assume_a", "lower")}))


def expm(a, **_kw):
    return _cast(_sl.expm(_np.asarray(a, dtype=float)))
