"""torchrun --nproc-per-node G tools/mg_check.py [n] : sharded dense log_probability vs single-GPU result."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tinygp_b200 import kernels, multigpu  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
slices = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ctx = multigpu.make_context(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(49382)
X = np.ascontiguousarray(rng.uniform(0, 20.0 * (n / 65536.0) ** (1 / 3), (n, 3)))
y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=n)
diag = np.full(n, 0.1)
k = 1.0 * kernels.ExpSquared(1.0)
for it in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lp = multigpu.log_probability_sharded(k, X, diag, y, slices=slices, ctx=ctx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"world={world} n={n} slices={slices} it={it} logp={lp!r} sec={dt:.3f}", flush=True)
if rank == 0 and n <= 16384:
    from oracle import tinygp_np as o
    lpo = o.GaussianProcess(o.Constant(1.0) * o.ExpSquared(1.0), X, diag=0.1).log_probability(y)
    print(f"oracle={lpo!r} rel={abs(lp - lpo) / abs(lpo):.3e}", flush=True)
if world > 1:
    dist.destroy_process_group()
