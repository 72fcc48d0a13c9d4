"""Batch sharding across the GPUs of one box: one process per GPU, contiguous slices of the parameter
batch, ONE collective at the end to gather final states / observables.

The reference has nothing distributed; every pulse point is an independent ODE, so the time
loop needs no communication.  Works with backend "nccl" (CUDA tensors, NVLink/NVSwitch) and "gloo" (CPU
tensors; used by the CPU tests of this logic).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` items for `rank`; the first total % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_last_dim(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Slice a [..., B] parameter block for this rank."""
    lo, hi = shard_bounds(t.shape[-1], rank, world)
    return t[..., lo:hi].contiguous()


def gather_batch(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None,
                 dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Gather per-rank results [B_local, ...] into [total, ...] (rank order = batch order).

    dst=None -> every rank gets the result (all_gather); dst=r -> only rank r (others return None).
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    is_complex = local.is_complex()
    buf = torch.view_as_real(local) if is_complex else local
    max_b = -(-int(total) // world)
    pad = max_b - buf.shape[0]
    if pad > 0:
        buf = torch.cat([buf, buf.new_zeros((pad,) + tuple(buf.shape[1:]))], dim=0)
    buf = buf.contiguous()
    out = None
    if dst is None:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, parts, dst=dst, group=group)
    if parts is not None:
        pieces = []
        for r in range(world):
            lo, hi = shard_bounds(total, r, world)
            pieces.append(parts[r][: hi - lo])
        out = torch.cat(pieces, dim=0)
        if is_complex:
            out = torch.view_as_complex(out.contiguous())
    return out
