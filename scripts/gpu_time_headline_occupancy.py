"""Headline RK4 kernel time against the number of two-trajectory warps per scheduler (592 schedulers on a B200): how much of
the kernel is throughput (time grows with warps per scheduler) and how much latency (it does not).  Usage: no arguments."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
y0 = np.array([1, 0, 0])
for B in (592 * 64, 592 * 2 * 64, 592 * 5 * 32, 100000, 592 * 3 * 64):
    rng = np.random.default_rng(1)
    p = engine.pack_params([dict(amp=rng.uniform(0.02, 0.5, B), angle=0.0, sigma=64.0, beta=rng.uniform(-5, 5, B))], device="cuda")
    best, mean = timed(lambda: engine.solve_sv_rk4(nm, [("drag", 0, 0, 256)], p, y0, (0, 257 * DT), DT / 40), reps=5, warm=2)
    print(f"B={B:7d} warps/scheduler={B / 64 / 592:.2f}: {best:.3f} ms -> {B / best * 1e3:.3e} traj/s")
