"""src/tinygp/solvers/quasisep/ops.py:24-214 over the device algebra (`b200gp_qsm_add / _elementwise_mul / _mul`)."""

from __future__ import annotations

__all__ = ["elementwise_add", "elementwise_mul", "qsm_mul"]

from tinygp_b200.solvers.quasisep.core import QSM


def elementwise_add(a: QSM, b: QSM) -> QSM:       # ops.py:24-35
    return a._binary(a._ctx.lib.b200gp_qsm_add, b)


def elementwise_mul(a: QSM, b: QSM) -> QSM:       # ops.py:38-49
    return a._binary(a._ctx.lib.b200gp_qsm_elementwise_mul, b)


def qsm_mul(a: QSM, b: QSM) -> QSM:               # ops.py:52-214
    return a._binary(a._ctx.lib.b200gp_qsm_mul, b)
