"""`jax.lax` control flow, eagerly."""

import numpy as _np

from .numpy import _cast
from .tree_util import tree_flatten, tree_leaves, tree_map, tree_unflatten


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def scan(f, init, xs, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
    else:
        leaves = tree_leaves(xs)
        n = _np.shape(leaves[0])[0] if leaves else length
    carry, ys, treedef = init, [None] * n, None
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        x = None if xs is None else tree_map(lambda v: _cast(_np.asarray(v)[i]), xs)
        carry, y = f(carry, x)
        leaves, treedef = tree_flatten(y)
        ys[i] = leaves
    if n == 0:
        return carry, None
    stacked = [_cast(_np.stack([_np.asarray(y[k]) for y in ys])) for k in range(len(ys[0]))]
    return carry, tree_unflatten(treedef, stacked)


def associative_scan(fn, elems, reverse=False, axis=0):
    """Inclusive scan as a left fold (right fold if `reverse`): `fn` is applied to batches of one element, which
    is how the reference's combine functions (written for batched operands) are meant to be called."""
    assert axis == 0
    leaves, treedef = tree_flatten(elems)
    n = _np.shape(leaves[0])[0]
    if n == 0:
        return elems

    def pick(i):
        return tree_unflatten(treedef, [_cast(_np.asarray(l)[i:i + 1]) for l in leaves])

    out = [None] * n
    if not reverse:
        acc = pick(0)
        out[0] = tree_leaves(acc)
        for i in range(1, n):
            acc = fn(acc, pick(i))
            out[i] = tree_leaves(acc)
    else:
        # jax: flip, scan with fn(a, b), flip back -> element i = fn(fn(..fn(x[n-1], x[n-2])..), x[i])
        acc = pick(n - 1)
        out[n - 1] = tree_leaves(acc)
        for i in range(n - 2, -1, -1):
            acc = fn(acc, pick(i))
            out[i] = tree_leaves(acc)
    stacked = [_cast(_np.concatenate([_np.asarray(o[k]) for o in out], axis=0)) for k in range(len(leaves))]
    return tree_unflatten(treedef, stacked)


def stop_gradient(x):
    return x
