"""Distance metrics (reference: src/tinygp/kernels/distance.py:41-59).

Only the two built-in metrics can be lowered to the device kernel program; custom
``Distance`` subclasses are arbitrary Python and are rejected by the B200 backend.
"""

from __future__ import annotations

__all__ = ["Distance", "L1Distance", "L2Distance"]


class Distance:
    code: int | None = None

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self))

    def __repr__(self):
        return f"{type(self).__name__}()"


class L1Distance(Distance):
    """sum_d |x1_d - x2_d|  (distance.py:41-45)"""
    code = 0


class L2Distance(Distance):
    """sqrt(sum_d (x1_d - x2_d)^2) with the r2 == 0 guard (distance.py:48-59)"""
    code = 1
