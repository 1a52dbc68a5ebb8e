import numpy as _np


def _tol(t, dtype, default):
    if isinstance(t, dict):
        return t.get(str(_np.dtype(dtype)), t.get("float64", default))
    return default if t is None else t


def check_close(xs, ys, atol=None, rtol=None, err_msg=""):
    xs, ys = _np.asarray(xs), _np.asarray(ys)
    dtype = _np.result_type(xs.dtype, ys.dtype)
    if not _np.issubdtype(dtype, _np.floating):
        dtype = _np.float64
    _np.testing.assert_allclose(xs, ys, atol=_tol(atol, dtype, 0), rtol=_tol(rtol, dtype, 1e-7), err_msg=err_msg)
