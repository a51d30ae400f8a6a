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

The sum over ranks can run inside the C call itself (``Plan.logp_grad(..., shard=...)``):
:class:`PeerGroup` — the finalize kernel writes into an IPC-mapped exchange buffer and one kernel
sums the ranks' partials through NVLink peer loads in rank order (bitwise identical on all ranks);
:class:`NcclComm` — ``ncclAllReduce`` on a communicator owned by the library's caller.
``torch.distributed`` is only the bootstrap that exchanges the 64-byte IPC handles / the NCCL id.
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


def _exchange_bytes(payload: bytes, group=None) -> list[bytes]:
    """All-gather of a small fixed-size host blob over ``torch.distributed`` (bootstrap only)."""
    world = dist.get_world_size(group)
    out: list = [None] * world
    dist.all_gather_object(out, payload, group=group)
    return out


class PeerGroup:
    """
    Peer-memory exchange group of the ranks of one box (``lslb_peer_*`` in the C ABI).
    ``max_elems`` = the largest result of a sharded call: ``C*(P+1)`` for logp+grad,
    ``C*p*(p+1)`` for IWLS. All ranks must make the same sharded calls in the same order.
    """

    transport = "peer"

    def __init__(self, max_elems: int, device: torch.device, group=None):
        import ctypes as C

        from . import _lib

        self._lib = _lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        self.handle = C.c_void_p()
        mine = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.lslb_peer_create(self.rank, self.world, int(max_elems),
                                                 C.byref(self.handle), mine), "lslb_peer_create")
            if self.world > 1:
                blobs = _exchange_bytes(mine.raw, group)
                _lib.check(_lib.lib.lslb_peer_connect(self.handle, C.c_char_p(b"".join(blobs))),
                           "lslb_peer_connect")
                dist.barrier(group)  # nobody publishes before every rank has mapped every buffer

    def timed_out(self) -> bool:
        import ctypes as C

        flag = C.c_int(0)
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.lslb_peer_timed_out(self.handle, C.byref(flag)), "lslb_peer_timed_out")
        return bool(flag.value)

    def abort(self):
        """Fails the whole group: every rank's current or next sharded call returns NaN."""
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib.lslb_peer_abort(self.handle, self._lib.stream_ptr()), "lslb_peer_abort")

    def reset(self, group=None):
        """Collective recovery after a failure: barrier, clear every rank's flags and epoch, barrier."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group)
            self._lib.check(self._lib.lib.lslb_peer_reset(self.handle), "lslb_peer_reset")
            if self.world > 1:
                dist.barrier(group)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            with torch.cuda.device(self.device):
                self._lib.lib.lslb_peer_destroy(self.handle)
            self.handle.value = None


class NcclComm:
    """An ``ncclComm_t`` created through the library's own thin NCCL wrappers (``lslb_nccl_*``)."""

    transport = "nccl"

    def __init__(self, device: torch.device, group=None):
        import ctypes as C

        from . import _lib

        self._lib = _lib
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        self.handle = C.c_void_p()
        uid = C.create_string_buffer(_lib.NCCL_ID_BYTES)
        if self.rank == 0:
            _lib.check(_lib.lib.lslb_nccl_unique_id(uid), "lslb_nccl_unique_id")
        blob = _exchange_bytes(uid.raw, group)[0] if self.world > 1 else uid.raw
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.lslb_nccl_comm_create(self.rank, self.world, C.c_char_p(blob),
                                                      C.byref(self.handle)), "lslb_nccl_comm_create")

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            with torch.cuda.device(self.device):
                self._lib.lib.lslb_nccl_comm_destroy(self.handle)
            self.handle.value = None


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
