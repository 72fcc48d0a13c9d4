"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch and gather results in batch order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casq_b200.distributed import gather_batch, shard_bounds, shard_last_dim


def test_shard_bounds_cover_batch_exactly():
    for total in (0, 1, 7, 100000, 100001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = torch.arange(6 * total, dtype=torch.float64).reshape(1, 6, total)
        mine = shard_last_dim(params, rank, world)
        lo, hi = shard_bounds(total, rank, world)
        assert mine.shape[-1] == hi - lo
        # stand-in for the per-rank solve: a complex "state" per point derived from its parameters
        local = torch.complex(mine[0, 0], mine[0, 4]).reshape(-1, 1, 1).repeat(1, 2, 3)
        full = gather_batch(local, total)
        only0 = gather_batch(local.real.contiguous(), total, dst=0)
        q.put((rank, full, only0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_two_rank_gloo_shard_and_gather(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params = torch.arange(6 * total, dtype=torch.float64).reshape(1, 6, total)
    want = torch.complex(params[0, 0], params[0, 4]).reshape(-1, 1, 1).repeat(1, 2, 3)
    for rank, full, only0 in results:
        assert torch.equal(full, want)
        if rank == 0:
            assert torch.equal(only0, want.real)
        else:
            assert only0 is None


def _peer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from casq_b200.distributed import PeerResultBuffer

        buf = PeerResultBuffer((4, 2, 3), torch.complex128, "cpu", dst=0)
        assert not buf.fused and buf.slot.shape == (4, 2, 3)  # CPU / gloo: the collective fallback
        buf.slot.copy_(torch.full((4, 2, 3), complex(rank + 1, -rank), dtype=torch.complex128))
        buf.sync()
        res = buf.result()
        buf.release()  # no-ops in the fallback, part of the call order every rank follows
        buf.drain()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_peer_result_buffer_falls_back_to_a_gather_without_symmetric_memory():
    """PeerResultBuffer on CPU tensors (gloo): same interface, results in rank order on the destination rank only.  On a B200
    box the slot is peer memory and the producing kernel's stores are the exchange (covered by bench.py --gpus N)."""
    from casq_b200.distributed import PeerResultBuffer

    single = PeerResultBuffer((3, 2), torch.float64, "cpu")
    single.slot.fill_(2.0)
    single.sync()
    assert single.world == 1 and torch.equal(single.result(), torch.full((3, 2), 2.0, dtype=torch.float64))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[1] is None
    want = torch.cat([torch.full((4, 2, 3), complex(1, 0), dtype=torch.complex128), torch.full((4, 2, 3), complex(2, -1), dtype=torch.complex128)])
    assert torch.equal(results[0], want)
