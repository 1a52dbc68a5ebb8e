"""`jax` on NumPy — eager, fp64, no tracing (see ../README.md).  Only what dfm/tinygp uses."""

import contextlib
import functools

import numpy as _np

from . import debug, errors, lax, random, scipy, test_util, tree_util  # noqa: F401
from . import numpy as _jnp
from . import numpy  # noqa: F401
from .numpy import Array, _cast
from .tree_util import tree_flatten, tree_map, tree_unflatten

__shim__ = True


class _Config:
    jax_enable_x64 = True

    def update(self, key, value):
        setattr(self, key, value)


config = _Config()


@contextlib.contextmanager
def enable_x64(flag=True):
    yield


def jit(fun=None, **_kwargs):
    if fun is None:
        return lambda f: f
    return fun


def _take(x, i, axis):
    x = _np.asarray(x)
    return _cast(_np.take(x, i, axis=axis)) if x.ndim else x


def vmap(fun, in_axes=0, out_axes=0):
    """A loop over slices followed by a stack.  `in_axes` may be an int/None or a tuple with one int/None per
    positional argument (each argument may itself be a pytree)."""

    @functools.wraps(fun, assigned=("__name__", "__doc__"), updated=())
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("vmap: in_axes does not match the arguments")
        size = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            for leaf in tree_util.tree_leaves(a):
                n = _np.shape(leaf)[ax]
                if size is None:
                    size = n
                elif size != n:
                    raise ValueError(f"vmap: inconsistent mapped axis sizes {size} vs {n}")
        if size is None:
            raise ValueError("vmap: nothing to map over")
        outs, treedef = [], None
        for i in range(size):
            sliced = [a if ax is None else tree_map(lambda x, ax=ax: _take(x, i, ax), a) for a, ax in zip(args, axes)]
            leaves, treedef = tree_flatten(fun(*sliced))
            outs.append(leaves)
        if size == 0:
            raise ValueError("vmap over an empty axis is not supported by the shim")
        stacked = [_cast(_np.stack([_np.asarray(o[k]) for o in outs], axis=out_axes)) for k in range(len(outs[0]))]
        return tree_unflatten(treedef, stacked)

    return mapped


def grad(*_a, **_k):
    import sys

    if "pytest" in sys.modules:
        sys.modules["pytest"].skip("jax.grad: the NumPy stand-in has no autodiff")
    raise NotImplementedError("jax.grad: the NumPy stand-in has no autodiff")


value_and_grad = grad


def block_until_ready(x):
    return x


def device_get(x):
    return x
