"""Probe: torch symmetric memory on this box (peer-mapped buffers over NVLink).  torchrun --nproc-per-node 2 scripts/gpu_symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B = 100000
t = symm_mem.empty((world, B, 2, 3), dtype=torch.complex128, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "backend", symm_mem.get_backend(dev), "ptrs", len(hdl.buffer_ptrs), "multicast", hdl.has_multicast_support, flush=True)
remote0 = hdl.get_buffer(0, (world, B, 2, 3), torch.complex128)
mine = remote0[rank]
mine.fill_(complex(rank + 1, -rank))
hdl.barrier()
if rank == 0:
    print("rank0 sees", [complex(t[r, 0, 0, 0].item()) for r in range(world)], flush=True)
# time: fill into peer slot + barrier vs gather
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
local = torch.randn((B, 2, 3), dtype=torch.complex128, device=dev)
def peer():
    mine.copy_(local); hdl.barrier()
parts = [torch.empty((B, 2, 3, 2), dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
def gather():
    dist.gather(torch.view_as_real(local), parts, dst=0)
print(rank, "peer copy + barrier ms", timed(peer), "nccl gather ms", timed(gather), flush=True)
dist.destroy_process_group()
