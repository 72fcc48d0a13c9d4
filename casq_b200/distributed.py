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


class PeerResultBuffer:
    """Result buffer of a batch-sharded solve that lives on ONE rank and is written by every rank's kernels directly.

    On a B200 box every GPU can address every peer's HBM through NVLink 5 / NVSwitch.  The buffer is a symmetric-memory
    allocation (``torch.distributed._symmetric_memory``); ``slot`` is this rank's [B_local, ...] slice of the DESTINATION
    rank's buffer, peer-mapped, so passing it as the ``out`` tensor of a solve makes the kernel's final stores land in the
    consumer's memory: the gather is fused into the producing kernel's epilogue (no collective launch, no staging copy).

    The hand-over is one-sided.  ``sync()`` on a producer rank raises a flag in the destination's signal pad (after its
    stores, in stream order) and returns; on the destination it waits for the flag of every producer.  Nobody else waits:
    a producer goes on with its next solve, into the next of ``depth`` ring slots, and blocks only when it is ``depth``
    solves ahead of the consumer's ``release()``.  So the consumer sees the pace of the slowest GPU, not the sum over
    steps of the slowest GPU of each step plus a barrier, which is what a collective or an all-rank barrier per solve
    costs (measured at N = 8 on the headline sweep, per 1e5-point step: NCCL gather 11.0 ms, all-rank device barrier
    10.6 ms, no exchange at all 10.2 ms).

    Order of calls per solve, every rank:  ``out = buf.slot`` -> solve -> ``buf.sync()`` -> (destination: use
    ``buf.result()``) -> ``buf.release()``.

    Without symmetric memory (CPU tensors / gloo, or a platform without peer access) the same interface falls back to a
    local slot plus ``gather_batch`` in ``sync()``.
    """

    def __init__(self, shape_per_rank, dtype, device, dst: int = 0, group: Optional[dist.ProcessGroup] = None, depth: int = 2):
        self.dst, self.group = int(dst), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = tuple(int(x) for x in shape_per_rank)
        self.depth = max(1, int(depth))
        self._hdl = None
        self._gathered = None
        self._cur = 0        # ring position of the solve being produced
        self._ready = None   # ring position handed over by the last sync()
        self._unacked = 0    # producer: solves published and not yet released by the destination
        device = torch.device(device)
        if self.world > 1 and device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                full = (self.depth, self.world) + self.shape
                self._buf = symm_mem.empty(full, dtype=dtype, device=device)
                self._hdl = symm_mem.rendezvous(self._buf, dist.group.WORLD if group is None else group)
                self._remote = self._hdl.get_buffer(self.dst, full, dtype)
            except Exception:  # noqa: BLE001 - no peer access / no symmetric memory: fall back to a collective
                self._hdl = None
        if self._hdl is None:
            self._buf = None
            self._local = torch.empty(self.shape, dtype=dtype, device=device)

    @property
    def fused(self) -> bool:
        """True when the slot is peer memory of the destination rank (stores of the producing kernel are the exchange)."""
        return self._hdl is not None

    @property
    def slot(self) -> torch.Tensor:
        """Where this rank's solve writes its [B_local, ...] results (changes with every sync(): re-read it per solve)."""
        if self._hdl is None:
            return self._local
        return self._remote[self._cur][self.rank]

    _DATA, _ACK = 0, 1  # signal-pad channels

    def sync(self) -> None:
        """Hand this rank's slot over to the destination rank / wait there for every rank's slot (stream-ordered)."""
        if self._hdl is not None:
            if self.rank == self.dst:
                for src in range(self.world):
                    if src != self.dst:
                        self._hdl.wait_signal(src, self._DATA)
            else:
                self._hdl.put_signal(self.dst, self._DATA)
                self._unacked += 1
            self._ready = self._cur
            self._cur = (self._cur + 1) % self.depth
            if self.rank != self.dst and self._unacked >= self.depth:
                # the ring slot the next solve writes still belongs to the consumer: wait for its release
                self._hdl.wait_signal(self.dst, self._ACK)
                self._unacked -= 1
        elif self.world > 1:
            self._gathered = gather_batch(self._local, self.shape[0] * self.world, self.group, dst=self.dst)

    def release(self) -> None:
        """Destination: the results handed over by the last sync() are consumed (stream-ordered); producers may reuse
        that ring slot.  No-op elsewhere and in the fallback."""
        if self._hdl is not None and self.rank == self.dst:
            for src in range(self.world):
                if src != self.dst:
                    self._hdl.put_signal(src, self._ACK)

    def drain(self) -> None:
        """Producer: consume outstanding releases (call once after the last solve so the signal pads end clean)."""
        if self._hdl is not None and self.rank != self.dst:
            while self._unacked > 0:
                self._hdl.wait_signal(self.dst, self._ACK)
                self._unacked -= 1

    def result(self) -> Optional[torch.Tensor]:
        """[world * B_local, ...] on the destination rank (rank order = batch order), None elsewhere.  Call after sync()."""
        if self.world == 1:
            return self.slot
        if self.rank != self.dst:
            return None
        if self._hdl is not None:
            pos = self._cur if self._ready is None else self._ready
            return self._buf[pos].reshape((self.world * self.shape[0],) + self.shape[1:])
        return self._gathered
