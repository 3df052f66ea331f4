"""Multi-GPU batch sharding: one process per GPU, the batch split contiguously, one NCCL all-gather of the results.

The reference has no distributed path; its only parallelism is user-side process pools over independent samples
(examples/tools/sampling/multiprocessing/closed_loop/mp_sampling_closed_loop_02.py:69-70).  Problem instances are
independent NLPs, so there is no data-path collective during the solve; ranks exchange only the converged
controls / status flags (B x n_u doubles) at the end — latency-bound over NVLink 5 / NVSwitch.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def shard_bounds(total: int, world_size: int, rank: int):
    """Contiguous shard [lo, hi) of `total` instances for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local: dict, total: int, group=None):
    """All-gather per-rank result tensors (u0_scaled [b,n_u], status [b], iter_count [b]) into global arrays.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors); shards may be ragged (padded to the largest)."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    sizes = [shard_bounds(total, ws, r) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    out = {}
    for key, t in local.items():
        pad_shape = (mx,) + tuple(t.shape[1:])
        buf = torch.zeros(pad_shape, dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        full = torch.empty((ws,) + pad_shape, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full.view(-1), buf.view(-1), group=group) if full.is_cuda else \
            dist.all_gather(list(full.unbind(0)), buf, group=group)
        out[key] = torch.cat([full[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
    return out


class ShardedBatchSolver:
    """Solve a global batch across the ranks of an initialised torch.distributed group (backend nccl)."""

    def __init__(self, mpc, group=None):
        import torch.distributed as dist
        self.mpc, self.group = mpc, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def solve(self, P_local, x0s_local, u0s_local, z0s_local=None, total: Optional[int] = None, gather=True):
        """Inputs: this rank's shard as CUDA float64 tensors.  Returns (local results, gathered results or None)."""
        local = self.mpc.S.solve_batch_device(P_local, x0s_local, u0s_local, z0s_local)
        if not gather or self.world == 1:
            return local, None
        total = total if total is not None else P_local.shape[0] * self.world
        g = gather_results({k: local[k] for k in ("u0_scaled", "status", "iter_count")}, total, self.group)
        return local, g
