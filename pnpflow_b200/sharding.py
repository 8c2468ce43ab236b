"""Batch sharding across the GPUs of one box.

Images in a batch are independent on the whole path (operators, GroupNorm, attention are per image; t is shared), so
the batch is split into contiguous slices, one process per GPU, and NCCL is used ONLY to scatter the measurements
and gather the restored images once per batch — there is no collective inside the step loop.

Parity caveat handled here: the reference's noise (`torch.randn_like` of the FULL batch, Philox grid depends on
numel) and its random/paintbrush masks (sequential host RNG over the FULL batch) are functions of the full-batch
tensor.  ``FullBatchNoise`` therefore draws the full-batch tensor on every rank with identical seeds and hands each
rank its slice; ``shard_operator`` slices a full-batch mask.
"""
from __future__ import annotations

import copy
from typing import Iterator, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank ``rank``; the first ``batch % world`` ranks get one extra image."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_batch(full: Optional[torch.Tensor], shape, dtype, device, src: int = 0, group=None) -> torch.Tensor:
    """Rank ``src`` holds ``full`` [B,...]; every rank returns its contiguous shard (dist.scatter over NCCL/gloo)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = shape[0]
    lo, hi = shard_bounds(B, world, rank)
    out = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=device)
    if B % world == 0:
        chunks = list(full.contiguous().chunk(world, dim=0)) if rank == src else None
        dist.scatter(out, chunks, src=src, group=group)
    else:                                  # ragged: broadcast and slice (rare; collectives need equal sizes)
        buf = full.contiguous() if rank == src else torch.empty(tuple(shape), dtype=dtype, device=device)
        dist.broadcast(buf, src=src, group=group)
        out.copy_(buf[lo:hi])
    return out


def gather_batch(shard: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All ranks receive the full batch [B,...] assembled from the shards (dist.all_gather)."""
    world = dist.get_world_size(group)
    if batch % world == 0:
        parts = [torch.empty_like(shard) for _ in range(world)]
        dist.all_gather(parts, shard.contiguous(), group=group)
        return torch.cat(parts, dim=0)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        buf = shard.contiguous() if r == dist.get_rank(group) else torch.empty((hi - lo,) + tuple(shard.shape[1:]),
                                                                                dtype=shard.dtype, device=shard.device)
        dist.broadcast(buf, src=r, group=group)
        parts.append(buf)
    return torch.cat(parts, dim=0)


class FullBatchNoise:
    """Iterator of per-(step, draw) noise slices that reproduces the reference's full-batch ``randn_like`` stream.

    Every rank seeds the same generator and draws the FULL [B,C,H,W] tensor per draw, then keeps rows [lo, hi).
    Cost: B*C*H*W Philox samples per draw per rank (<< 1 % of a step)."""
    def __init__(self, full_shape, lo: int, hi: int, device, generator: Optional[torch.Generator] = None):
        self.full_shape, self.lo, self.hi, self.device, self.gen = tuple(full_shape), lo, hi, device, generator

    def __iter__(self) -> Iterator[torch.Tensor]:
        return self

    def __next__(self) -> torch.Tensor:
        eps = torch.randn(self.full_shape, device=self.device, generator=self.gen)
        return eps[self.lo:self.hi]


def shard_operator(op, lo: int, hi: int, full_batch: int):
    """Restrict an engine operator to images [lo, hi) of a ``full_batch``-image batch (only masks are per image)."""
    from .degradations import _CachedMask
    if not isinstance(op, _CachedMask):
        return op
    view = copy.copy(op)
    view._cache = {}
    parent_host_mask = op._host_mask

    def _host_mask(B, H, W):
        assert B == hi - lo
        return parent_host_mask(full_batch, H, W)[lo:hi]

    view._host_mask = _host_mask
    return view
