"""BASELINE configs 3, 4, 5 sharded by batch across the GPUs of one box: one process per GPU, each
rank builds its own model handle and solves its contiguous slice, ONE NCCL gather of the results per solve
(observables only for cfg 4), device-timed per rank, max over ranks.  cfg 4 is the config BASELINE names
"batch-sharded 8 GPUs": B = 1024 density matrices = 128 per GPU at N = 8 (weak scaling: 128 per GPU at every N).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/bench_configs_multi.py --out gpurun_out/configs_multi.json
"""
import argparse, json, os, sys
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from casq_b200 import engine
from casq_b200.distributed import gather_batch, shard_bounds
from scripts import _problems as P

DT = P.DT


def timed_ranks(fn, reps, warm):
    """best-of-reps of max-over-ranks device time, every repetition bracketed by a barrier"""
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t))
    return best


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"n_gpus": world, "gpu": torch.cuda.get_device_name(local)}

    # ---- cfg 4: 5 transmons, T1/T2, DRAG on d0, RK4 dt/4 with RWA, 128 density matrices per GPU
    st5, ops5, ch5 = P.manila(None)
    nm5 = P.native(st5, ops5, ch5, frame="diag", rwa=0.75 * P.CARRIERS["d0"], device=local)
    nm5.set_dissipators(P.t1t2(5))
    B4 = 128 * world
    lo, hi = shard_bounds(B4, rank, world)
    amp = np.linspace(0.05, 0.5, B4)[lo:hi]
    p4 = engine.pack_params([dict(amp=amp, angle=0.0, sigma=64.0, beta=1.0)], device="cuda")
    rho0 = np.zeros(243, dtype=complex); rho0[0] = 1
    keep = {}
    def run4():
        pops = engine.solve_lindblad_rk4(nm5, [("drag", ch5.index("d0"), 0, 256)], p4, rho0, (0, 257 * DT), DT / 4, want_rho=False)[1]
        keep["pops"] = gather_batch(pops, B4, dst=0)  # observables only: [B, T, 243] f64
    ms = timed_ranks(run4, reps=2, warm=1)
    if rank == 0:
        pops = keep["pops"]
        res["cfg4_lindblad_d243"] = {"B_total": B4, "B_per_gpu": 128, "steps": 1028, "ms_best": ms, "traj_per_s": B4 / ms * 1e3,
                                     "algorithmic_GBps_total": 8 * 243 * 243 * 16 * 1028 * B4 / (ms * 1e-3) / 1e9,
                                     "gathered_shape": list(pops.shape), "trace_min": float(pops[:, -1].sum(-1).min()),
                                     "trace_max": float(pops[:, -1].sum(-1).max())}

    # ---- cfg 3: 4096 cross-resonance unitaries per GPU
    st2, ops2, ch2 = P.chain(2)
    nm2 = P.native(st2, ops2, ch2, rwa=0.75 * P.CARRIERS["d0"], device=local)
    B3 = 4096 * world
    ww, aa = np.meshgrid(np.linspace(0, 768, 64 * world), np.linspace(0.05, 0.8, 64), indexing="ij")
    lo, hi = shard_bounds(B3, rank, world)
    p3 = engine.pack_params([dict(amp=aa.ravel()[lo:hi], angle=0.0, sigma=64.0, width=ww.ravel()[lo:hi])], device="cuda")
    def run3():
        U = engine.propagators_expm(nm2, [("gaussian_square", ch2.index("u0"), 0, 1024)], p3, (0, 1024 * DT), DT)
        keep["U"] = gather_batch(U, B3, dst=0)
    ms = timed_ranks(run3, reps=3, warm=1)
    if rank == 0:
        res["cfg3_expm_d9_rwa_m1"] = {"B_total": B3, "ms_best": ms, "traj_per_s": B3 / ms * 1e3, "gathered_shape": list(keep["U"].shape)}

    # ---- cfg 5: 1e4 Dopri5 candidates per GPU
    st1, ops1, ch1 = P.manila([0])
    nm1 = P.native(st1, ops1, ch1, device=local)
    B5 = 10000 * world
    rng = np.random.default_rng(20260101)
    sig, bet = rng.uniform(16, 128, B5), rng.uniform(-4, 4, B5)
    lo, hi = shard_bounds(B5, rank, world)
    p5 = engine.pack_params([dict(amp=0.12, angle=0.0, sigma=sig[lo:hi], beta=bet[lo:hi])], device="cuda")
    def run5():
        y = engine.solve_sv_dopri5(nm1, [("drag", 0, 0, 256)], p5, np.array([1, 0, 0]), (0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT)
        keep["y"] = gather_batch(y, B5, dst=0)
    ms = timed_ranks(run5, reps=3, warm=1)
    if rank == 0:
        res["cfg5_dopri5_d3"] = {"B_total": B5, "ms_best": ms, "traj_per_s": B5 / ms * 1e3, "gathered_shape": list(keep["y"].shape)}
        print(json.dumps(res, indent=1))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
