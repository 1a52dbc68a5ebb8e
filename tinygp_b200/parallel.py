"""Host-side helpers for the one-process-per-GPU runs (bench.py): work sharding and timing reduction.

The dense and batched paths shard by *independent problems* (hyper-parameter points / grid entries): no
data-path collective is needed, only (i) a deterministic problem -> rank map, (ii) a MAX reduction of the
per-rank device time and (iii) a gather of the per-problem scalars.  These are backend-agnostic
(`nccl` on the GPU box, `gloo` in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_indices(nprob: int, rank: int, world: int) -> np.ndarray:
    """Round-robin problem indices owned by `rank` (matches `grid[rank::world]` in bench.py)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return np.arange(rank, nprob, world)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a per-rank scalar (device time); identity when not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local_idx: np.ndarray, local_vals: np.ndarray, nprob: int, device=None) -> np.ndarray:
    """Assemble the per-problem results of all ranks into one array on every rank (sum of disjoint scatters)."""
    import torch
    import torch.distributed as dist

    out = torch.zeros(nprob, dtype=torch.float64, device=device)
    out[torch.as_tensor(local_idx, dtype=torch.long, device=device)] = torch.as_tensor(
        local_vals, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out.cpu().numpy()
