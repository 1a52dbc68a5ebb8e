"""`equinox.Module` stand-in: every subclass becomes a dataclass (a user-written `__init__` is kept and may assign
fields), `field(static=True)` marks non-leaf fields for the shim's pytrees, `__check_init__` hooks run after
construction.  No freezing, no tracing."""

import abc
import dataclasses

_MISSING = dataclasses.MISSING


def field(*, default=_MISSING, default_factory=_MISSING, static=False, converter=None, init=True, repr=True,  # noqa: A002
          metadata=None, **_kw):
    md = dict(metadata or {})
    md["static"] = static
    if converter is not None:
        md["converter"] = converter
    kwargs = {"metadata": md, "init": init, "repr": repr}
    if default is not _MISSING:
        kwargs["default"] = default
    if default_factory is not _MISSING:
        kwargs["default_factory"] = default_factory
    return dataclasses.field(**kwargs)


def static_field(**kw):
    return field(static=True, **kw)


class _ModuleMeta(abc.ABCMeta):
    def __new__(mcs, name, bases, ns, **kwargs):
        cls = super().__new__(mcs, name, bases, ns, **kwargs)
        own_init = "__init__" in ns
        inherited_custom = any(getattr(b, "__shim_custom_init__", False) for b in bases)
        declares_fields = bool(ns.get("__annotations__"))
        # equinox: a user __init__ wins; a class that adds fields without writing __init__ gets the dataclass one
        make_init = (not own_init) and (declares_fields or not inherited_custom)
        cls.__shim_custom_init__ = own_init or (inherited_custom and not make_init)
        cls = dataclasses.dataclass(init=make_init, eq=False, repr=False, match_args=False)(cls)
        cls.__is_shim_module__ = True
        return cls

    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        for f in dataclasses.fields(cls):
            conv = f.metadata.get("converter")
            if conv is not None and hasattr(obj, f.name):
                object.__setattr__(obj, f.name, conv(getattr(obj, f.name)))
        for klass in reversed(cls.__mro__):
            check = klass.__dict__.get("__check_init__")
            if check is not None:
                check(obj)
        return obj


class Module(metaclass=_ModuleMeta):
    def __repr__(self):
        parts = ", ".join(f"{f.name}={getattr(self, f.name, None)!r}" for f in dataclasses.fields(self))
        return f"{type(self).__name__}({parts})"


def tree_at(*_a, **_k):
    raise NotImplementedError("equinox.tree_at is not provided by the shim")


def filter_jit(fun=None, **_k):
    return fun if fun is not None else (lambda f: f)


def is_array(x):
    import numpy as np

    return isinstance(x, np.ndarray)


def partition(tree, _filter):
    return tree, None


def combine(a, _b):
    return a
