#!/usr/bin/env python
"""Headline benchmark: BASELINE.json configs[1] -- 3-level transmon, DRAG beta x amplitude sweep, 1e5 pulse
points, fixed-step RK4 (max_dt = dt/40 over [0, 257 dt] => S = 10 280 steps), state-vector fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over the whole 1e5-point batch on each GPU.  Prints ONE JSON line.
  value     trajectories/s, parameters resident in HBM when the timed region starts (table + RK4 kernels)
  e2e       same metric through the public API (B200PulseBackend.solve_slots_batch) from pinned HOST
            parameter arrays to HOST final states + populations, H2D/D2H inside the timed region
  roofline  RK4 kernel vs the fp64 FMA peak measured live on this GPU (casq_fp64_fma_probe)
  cpu_baseline  the oracle port (numpy restatement of casq's qiskit-dynamics path) timed on host cores
--impl reference times that CPU path alone (all host cores, one single-point solve per task, as casq runs a
sweep: a Python loop of solve calls).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload (SURVEY.md §8(d) cfg 2; constants from src/casq/models/transmon_model.py:94-114 and
# docs/source/user_guide/pulse_backend.rst:86-99 of the reference)
DT = 2.2222222222222221e-10
N_BETA, N_AMP = 250, 400
DURATION, SIGMA = 256, 64.0
T_FINAL_SAMPLES = 257
SUBSTEPS = 40
STEPS_PER_TRAJ = T_FINAL_SAMPLES * SUBSTEPS  # 10 280
# Algorithmic flops of ONE RK4 step for the ladder-coupled 3-level transmon with one driven channel, as the
# kernel evaluates it: per stage a signal-free matvec with the two complex couplings
# (6 mul + 10 fma = 26 flop), the real signal (3 flop) and its 1-2 coefficient products, and the fused
# stage combinations (12 fma): 36 mul/add + 85 fma = 121 fp64 instructions = 206 flop per step.
FLOP_PER_STEP = 206
FLOP_PER_STEP_DENSE_MODEL = 726  # SURVEY §8(d) dense d=3,K=1 count, reported for reference only
METRIC = "pulse trajectories/sec (d=3, S=10280 RK4 steps, fp64)"


def sweep_params(rank: int = 0):
    """beta x amp grid of cfg 2; rank r > 0 sweeps the same grid at sigma = 64 + r (weak scaling: every
    GPU gets its own 1e5-point sweep)."""
    beta = np.linspace(-5.0, 5.0, N_BETA)
    amp = np.linspace(0.02, 0.5, N_AMP)
    bb, aa = np.meshgrid(beta, amp, indexing="ij")
    return {"amp": aa.ravel(), "angle": 0.0, "sigma": SIGMA + rank, "beta": bb.ravel()}


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "cfg2: 3-level transmon (ibmq_manila q0), DragPulseGate(duration=256) sigma=64, "
                    "beta in linspace(-5,5,250) x amp in linspace(0.02,0.5,400) = 1e5 points per GPU, "
                    "RK4 max_dt=dt/40 over [0,257dt] (S=10280), no RWA, frame=H0",
        "points_per_gpu": N_BETA * N_AMP,
        "global_points": N_BETA * N_AMP * n_gpus,
        "dim": 3,
        "steps_per_trajectory": STEPS_PER_TRAJ,
        "parallelism": f"batch-sharded x{n_gpus}",
        "l2": "flushed between timed iterations (256 MiB write)",
    }


# --------------------------------------------------------------------------------------------------
def cpu_oracle_points(indices, rank=0):
    """Single-point oracle solves (what casq does per sweep point).  Returns seconds taken."""
    import oracle as O

    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    model = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=st)
    p = sweep_params(rank)
    y0 = np.array([1, 0, 0], dtype=complex)
    t0 = time.perf_counter()
    for i in indices:
        samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", DURATION)],
                                    [dict(amp=p["amp"][i], angle=0.0, sigma=p["sigma"], beta=p["beta"][i])])
        O.rk4_solve(model, samples, y0, (0.0, T_FINAL_SAMPLES * DT), DT / SUBSTEPS)
    return time.perf_counter() - t0


def _cpu_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return cpu_oracle_points(*args)


def run_reference(args) -> None:
    """--impl reference: the oracle port on all host cores (rank 0 only under torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    per_step = cores  # one single-point solve per core per step
    rng = np.random.default_rng(20260101)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            idx = rng.integers(0, N_BETA * N_AMP, size=per_step)
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [([int(i)], 0) for i in idx])
            el = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(el)
    ms = 1e3 * sum(times) / len(times)
    value = per_step / (ms * 1e-3)
    sample = f"{per_step} random points of the cfg2 grid per step, one single-point RK4 solve (S=10280) per core"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from casq_b200 import engine
    from casq_b200.backends import BackendLibrary, build
    from casq_b200.models import ControlModel, TransmonModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    carriers = {"d0": 4962770879.920025, "u0": 4838412258.764764}
    hamiltonian = TransmonModel.manila(extracted_qubits=[0], evaluation_mode=TransmonModel.EvaluationMode.DENSE)
    control = ControlModel(dt=DT, channel_carrier_freqs=carriers)
    backend = build(BackendLibrary.B200, hamiltonian, control, device=local_rank)
    model = backend.native
    slots = [("drag", hamiltonian.channels.index("d0"), 0, DURATION)]
    p = sweep_params(rank)
    B = N_BETA * N_AMP
    host_params = engine.pack_params([p], pin=True)  # [1,5,B] pinned
    dev_params = host_params.to(dev)
    y0 = torch.tensor(backend.ground_state(), dtype=torch.complex128, device=dev)
    t_span = (0.0, T_FINAL_SAMPLES * DT)
    max_dt = DT / SUBSTEPS
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = torch.empty((B, 2, 3), dtype=torch.complex128, device=dev)
    gathered = [torch.empty((B, 3, 2), dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def hot_step():
        """table kernel + RK4 kernel (+ the one result gather when sharded)."""
        model.invalidate_cache()
        engine.solve_sv_rk4(model, slots, dev_params, y0, t_span, max_dt, out=out)
        if world > 1:
            dist.all_gather(gathered, torch.view_as_real(out[:, 1, :]).contiguous())

    def timed(fn, k):
        tot = 0.0
        for _ in range(k):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / k

    # ---- value: device-resident inputs
    for _ in range(args.warmup):
        hot_step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        ms_step = timed(hot_step, args.steps)
        barrier()
        # ---- roofline: the RK4 kernel alone (table cached), same timed region for the clock record
        engine.solve_sv_rk4(model, slots, dev_params, y0, t_span, max_dt, out=out)

        def kernel_only():
            engine.solve_sv_rk4(model, slots, dev_params, y0, t_span, max_dt, out=out)

        ms_kernel = timed(kernel_only, args.steps)
    ms_step = max_over_ranks(ms_step)
    ms_kernel = max_over_ranks(ms_kernel)

    # ---- e2e: public API, host -> host
    host_final = torch.empty((B, 3), dtype=torch.complex128).pin_memory()
    host_pops = torch.empty((B, 2), dtype=torch.float64).pin_memory()

    def e2e_step():
        model.invalidate_cache()
        dparams = host_params.to(dev, non_blocking=True)
        sol = backend.solve_slots_batch(slots, None, T_FINAL_SAMPLES, "RK4", run_options={"max_dt": max_dt},
                                        params_device=dparams)
        host_final.copy_(sol.states[:, -1, :], non_blocking=True)
        host_pops.copy_(sol.populations[:, -1, :], non_blocking=True)
        if world > 1:
            dist.all_gather(gathered, torch.view_as_real(sol.states[:, -1, :]).contiguous())
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    ms_e2e = max_over_ranks(timed(e2e_step, args.steps))
    barrier()
    h2d = host_params.numel() * 8
    d2h = host_final.numel() * 16 + host_pops.numel() * 8

    fp64_peak = engine.fp64_fma_probe(local_rank) if rank == 0 else 0.0
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = 12
        idx = np.random.default_rng(20260101).integers(0, B, size=n)
        sec = cpu_oracle_points([int(i) for i in idx])
        cpu = {"value": n / sec, "unit": "trajectories/s", "cores": 1, "kind": "port",
               "sample": f"{n} random points of the cfg2 grid, single-point RK4 solves (S=10280) in a Python loop, 1 core "
                         f"of {os.cpu_count()}"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        total = B * world
        flops = FLOP_PER_STEP * STEPS_PER_TRAJ * B
        achieved = flops / (ms_kernel * 1e-3) / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r01_sv_rk4_small_dram_bytes.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        print(json.dumps({
            "metric": METRIC, "value": total / (ms_step * 1e-3), "unit": "trajectories/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "trajectories/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                         "kernel": "sv_rk4_small2_kernel<3,TRI> (2 trajectories per thread)", "kernel_ms": ms_kernel,
                         "flop_per_step": FLOP_PER_STEP, "dense_model_flop_per_step": FLOP_PER_STEP_DENSE_MODEL,
                         "peak_source": "measured live: casq_fp64_fma_probe (MEASURED_PEAKS.json has no fp64 entry)"},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
