"""One cfg 5 solve (for ncu): python scripts/gpu_dopri5_once.py [B]  (CASQ_DOPRI5_LANES picks the variant)"""
import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rng = np.random.default_rng(20260101)
p = engine.pack_params([dict(amp=0.12, angle=0.0, sigma=rng.uniform(16, 128, B), beta=rng.uniform(-4, 4, B))], device="cuda")
for _ in range(2):
    y, ns, stt = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], p, np.array([1, 0, 0]), (0.0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
torch.cuda.synchronize()
print("ok", int(ns[:, 0].sum()))
