"""Minimal pytrees: tuples, lists, dicts, None, namedtuples and equinox-shim Modules are nodes; the rest are leaves."""

import dataclasses

import numpy as np


class _Leaf:
    def __repr__(self):
        return "*"


LEAF = _Leaf()


def _is_module(x):
    return getattr(type(x), "__is_shim_module__", False)


def _module_fields(x):
    dyn, static = [], []
    for f in dataclasses.fields(x):
        if f.metadata.get("static", False):
            static.append((f.name, getattr(x, f.name, None)))
        elif hasattr(x, f.name):
            dyn.append(f.name)
    return dyn, tuple(static)


def _flatten(tree, leaves, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        leaves.append(tree)
        return LEAF
    if tree is None:
        return ("none",)
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return ("namedtuple", type(tree), [_flatten(t, leaves, is_leaf) for t in tree])
    if isinstance(tree, tuple):
        return ("tuple", [_flatten(t, leaves, is_leaf) for t in tree])
    if isinstance(tree, list):
        return ("list", [_flatten(t, leaves, is_leaf) for t in tree])
    if isinstance(tree, dict):
        keys = sorted(tree.keys())
        return ("dict", keys, [_flatten(tree[k], leaves, is_leaf) for k in keys])
    if _is_module(tree):
        dyn, static = _module_fields(tree)
        return ("module", type(tree), dyn, static, [_flatten(getattr(tree, n), leaves, is_leaf) for n in dyn])
    leaves.append(tree)
    return LEAF


def _unflatten(treedef, it):
    if treedef is LEAF:
        return next(it)
    kind = treedef[0]
    if kind == "none":
        return None
    if kind == "tuple":
        return tuple(_unflatten(t, it) for t in treedef[1])
    if kind == "namedtuple":
        return treedef[1](*[_unflatten(t, it) for t in treedef[2]])
    if kind == "list":
        return [_unflatten(t, it) for t in treedef[1]]
    if kind == "dict":
        return {k: _unflatten(t, it) for k, t in zip(treedef[1], treedef[2])}
    if kind == "module":
        _, cls, dyn, static, children = treedef
        obj = object.__new__(cls)
        for name, child in zip(dyn, children):
            object.__setattr__(obj, name, _unflatten(child, it))
        for name, value in static:
            object.__setattr__(obj, name, value)
        return obj
    raise TypeError(treedef)


def tree_flatten(tree, is_leaf=None):
    leaves = []
    treedef = _flatten(tree, leaves, is_leaf)
    return leaves, treedef


def tree_unflatten(treedef, leaves):
    return _unflatten(treedef, iter(leaves))


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] for r in rest]
    for o in others:
        if len(o) != len(leaves):
            raise ValueError("tree_map: mismatched tree structures")
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def tree_reduce(f, tree, *init):
    import functools

    return functools.reduce(f, tree_leaves(tree), *init)


def tree_all(tree):
    return all(tree_leaves(tree))


def _treedef_eq(a, b):
    return repr(a) == repr(b)


__all__ = ["tree_flatten", "tree_unflatten", "tree_leaves", "tree_structure", "tree_map", "tree_reduce"]
_ = np
