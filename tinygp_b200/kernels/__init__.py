"""Kernel building blocks, same names as ``tinygp.kernels`` (src/tinygp/kernels/__init__.py)."""

from tinygp_b200.kernels import quasisep as quasisep
from tinygp_b200.kernels.base import (
    Conditioned as Conditioned,
    Constant as Constant,
    Custom as Custom,
    DotProduct as DotProduct,
    Kernel as Kernel,
    Polynomial as Polynomial,
    Product as Product,
    Sum as Sum,
)
from tinygp_b200.kernels.distance import (
    Distance as Distance,
    L1Distance as L1Distance,
    L2Distance as L2Distance,
)
from tinygp_b200.kernels.stationary import (
    Cosine as Cosine,
    Exp as Exp,
    ExpSineSquared as ExpSineSquared,
    ExpSquared as ExpSquared,
    Matern32 as Matern32,
    Matern52 as Matern52,
    RationalQuadratic as RationalQuadratic,
    Stationary as Stationary,
)
