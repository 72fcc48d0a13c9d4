"""cfg 5 (1e4 DRAG candidates, Dopri5): time one solve and print the step-count statistics.
Usage: python scripts/gpu_dopri5_probe.py [B]"""
import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
y0 = np.array([1, 0, 0])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(20260101)
p = engine.pack_params([dict(amp=0.12, angle=0.0, sigma=rng.uniform(16, 128, B), beta=rng.uniform(-4, 4, B))], device="cuda")
run = lambda: engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], p, y0, (0.0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
y, ns, stt = run()
best, mean = timed(run, reps=5, warm=3)
print(f"B={B}: {best:.3f} ms ({B / best * 1e3:.3e} traj/s), attempted {int(ns[:, 0].sum())}, failed {int((stt != 0).sum())}")
ref = (y, ns)
n = ref[1][:, 0].cpu().numpy()
w = n[: (B // 32) * 32].reshape(-1, 32).max(axis=1)
print(f"attempted steps per trajectory: min {n.min()} mean {n.mean():.0f} max {n.max()}; per-warp max: mean {w.mean():.0f} min {w.min()} max {w.max()}")
