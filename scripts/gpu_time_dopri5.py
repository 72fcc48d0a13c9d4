"""BASELINE cfg 5 alone: 1e4 DRAG candidates, adaptive Dopri5 (atol 1e-6, rtol 1e-8, hmax dt).  Usage: gpu_time_dopri5.py [B]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
B5 = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(20260101)
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
p5 = engine.pack_params([dict(amp=0.12, angle=0.0, sigma=rng.uniform(16, 128, B5), beta=rng.uniform(-4, 4, B5))], device="cuda")
out5 = {}
def run5():
    out5["r"] = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], p5, np.array([1, 0, 0]), (0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
best, mean = timed(run5, reps=3, warm=1)
y5, ns5, st5 = out5["r"]
print(f"dopri5 d=3 B={B5}: {best:.3f} ms -> {B5/best*1e3:.0f} traj/s; attempted {int(ns5[:,0].sum())} accepted {int(ns5[:,1].sum())} failed {int((st5!=0).sum())}; "
      f"checksum {float(y5.real.sum()):.12f} {float(y5.imag.sum()):.12f}")
