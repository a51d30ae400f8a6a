"""
Partitioning of the hot path over the GPUs of one box (one process per GPU, ``torch.distributed``).

The reference has no multi-device path at all (its only parallelism is ``jax.vmap`` over
chains on one device, ``goose/engine.py:442-448``). Two partitionings follow from the structure of
the computation:

* **chains** — evaluations of different chains are independent given the data: replicate the data,
  give every rank its own chains, no communication (:func:`chain_shard`).
* **rows** — log-likelihood, gradient and information are sums over rows: give every rank a
  contiguous block of rows and ALL chains, evaluate with ``LSLB_SKIP_PRIOR`` on ranks > 0 and
  all-reduce ``[C, P+1]`` (:class:`RowSharded`). A random intercept is sharded by whole groups so its
  gradient needs no reduction (:func:`row_shard` with ``group_idx``).
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist

SKIP_PRIOR = 1  # LSLB_SKIP_PRIOR in include/liesel_b200.h


def chain_shard(num_chains: int, rank: int, world_size: int) -> tuple[int, int]:
    """Half-open range of chains owned by ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(num_chains, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def row_shard(n: int, rank: int, world_size: int, group_idx: np.ndarray | None = None) -> tuple[int, int]:
    """
    Half-open range of rows owned by ``rank``. With ``group_idx`` (ascending) every boundary is
    moved forward to the next group start, so the rows of one group never straddle two ranks.
    """
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    bounds = [(r * n) // world_size for r in range(world_size + 1)]
    if group_idx is not None:
        g = np.asarray(group_idx)
        if g.shape != (n,):
            raise ValueError("group_idx must have one entry per row")
        for r in range(1, world_size):
            b = bounds[r]
            while 0 < b < n and g[b] == g[b - 1]:
                b += 1
            bounds[r] = b
        for r in range(1, world_size + 1):
            bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds[rank], bounds[rank + 1]


class RowSharded:
    """
    Sums per-rank partial evaluations. ``evaluate_local(params, aux, flags) -> (logp [C], grad [C,P]
    or None)`` is ``Plan.logp_grad`` of the plan built on this rank's rows; priors and cached
    constants are contributed by rank 0 only. One all-reduce of ``C*(P+1)`` values per evaluation
    (NCCL on GPUs: ~1 MB at C=1024, P=256 — latency-bound, on the evaluation's own stream).
    """

    def __init__(self, evaluate_local: Callable, group=None):
        self.evaluate_local = evaluate_local
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def logp_grad(self, params: torch.Tensor, aux: torch.Tensor | None = None, want_grad: bool = True):
        flags = 0 if self.rank == 0 else SKIP_PRIOR
        logp, grad = self.evaluate_local(params, aux, flags)
        if self.world == 1:
            return logp, grad
        if want_grad:
            buf = torch.cat([logp[:, None], grad], dim=1).contiguous()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            return buf[:, 0].contiguous(), buf[:, 1:].contiguous()
        buf = logp.contiguous().clone()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf, None


def gather_chains(local: torch.Tensor, num_chains: int, group=None) -> torch.Tensor:
    """All-gather per-rank chain blocks (leading axis) into ``[num_chains, ...]`` on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [chain_shard(num_chains, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)
