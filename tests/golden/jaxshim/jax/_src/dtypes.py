import numpy as _np


def to_complex_dtype(dtype):
    return _np.result_type(dtype, _np.complex64)
