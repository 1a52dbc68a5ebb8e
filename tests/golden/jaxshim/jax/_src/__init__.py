from . import dtypes  # noqa: F401
