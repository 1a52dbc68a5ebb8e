"""Block-diagonal matrices on the host (reference: src/tinygp/solvers/quasisep/block.py:19-137).

The reference uses ``Block`` for the transition matrix, design matrix and stationary covariance of a ``Sum`` of quasiseparable
kernels, so that XLA skips the zero blocks.  On the B200 backend that sparsity is exploited inside the CUDA kernels (compile-time
block layouts, csrc/qs_fast.cuh); this class only keeps the host-side state-space queries of ``kernels.quasisep`` drop-in
compatible: small J x J matrices, optionally with leading batch axes, NumPy arithmetic."""

from __future__ import annotations

__all__ = ["Block", "ensure_dense"]

import numpy as np


def ensure_dense(x):
    """a ``Block`` as a dense array; anything else unchanged (block.py:12-16)"""
    return x.to_dense() if isinstance(x, Block) else x


class Block:
    __array_priority__ = 1999      # block.py:21: NumPy defers ``ndarray @ Block`` etc. to the reflected operators below

    def __init__(self, *blocks):
        self.blocks = tuple(np.asarray(b, dtype=np.float64) for b in blocks)

    def _map(self, fn):
        return Block(*(fn(b) for b in self.blocks))

    def _zip(self, other, fn):
        if len(self.blocks) != len(other.blocks) or any(a.shape != b.shape for a, b in zip(self.blocks, other.blocks)):
            raise ValueError("Block operands must have the same block structure")
        return Block(*(fn(a, b) for a, b in zip(self.blocks, other.blocks)))

    def __getitem__(self, idx):
        return self._map(lambda b: b[idx])

    def __len__(self):
        if any(b.ndim != 2 for b in self.blocks):
            raise TypeError("len() of a batched Block")
        return sum(b.shape[0] for b in self.blocks)

    @property
    def ndim(self):
        (ndim,) = {b.ndim for b in self.blocks}
        return ndim

    @property
    def shape(self):
        return (len(self), len(self))

    def transpose(self):
        return self._map(lambda b: b.transpose())

    @property
    def T(self):
        return self.transpose()

    @property
    def mT(self):
        return self._map(lambda b: np.swapaxes(b, -1, -2))

    def to_dense(self):
        size = sum(b.shape[-1] for b in self.blocks)
        out = np.zeros(self.blocks[0].shape[:-2] + (size, size))
        o = 0
        for b in self.blocks:
            out[..., o:o + b.shape[-2], o:o + b.shape[-1]] = b
            o += b.shape[-1]
        return out

    def __mul__(self, other):
        return self._map(lambda b: b * other)

    __rmul__ = __mul__

    def __add__(self, other):
        return self._zip(other, np.add) if isinstance(other, Block) else self.to_dense() + other

    def __radd__(self, other):
        return other + self.to_dense()

    def __sub__(self, other):
        return self._zip(other, np.subtract) if isinstance(other, Block) else self.to_dense() - other

    def __rsub__(self, other):
        return other - self.to_dense()

    def __matmul__(self, other):
        if isinstance(other, Block):
            return self._zip(other, np.matmul)
        other = np.asarray(other, dtype=np.float64)
        if other.ndim < 1:
            raise ValueError("Block @ scalar")
        o, ys = 0, []
        for b in self.blocks:                       # row slab o : o + size of the right operand per block
            size = b.shape[-1]
            ys.append(b @ (other[o:o + size] if other.ndim == 1 else other[..., o:o + size, :]))
            o += size
        return np.concatenate(ys, axis=0 if other.ndim == 1 else -2)

    def __rmatmul__(self, other):
        other = np.asarray(other, dtype=np.float64)
        o, ys = 0, []
        for b in self.blocks:                       # column slab of the left operand per block
            size = b.shape[-2]
            ys.append(other[..., o:o + size] @ b)
            o += size
        return np.concatenate(ys, axis=-1)
