"""Headline sweep (1e5 points, S = 10 280) solved by the constant-table and by the shared-memory-table variant of
sv_rk4_small2_kernel: prints the largest difference between the two.  Usage: python scripts/gpu_ct_compare.py"""
import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
y0 = np.array([1, 0, 0])
B = 100000
rng = np.random.default_rng(1)
p = engine.pack_params([dict(amp=rng.uniform(0.02, 0.5, B), angle=rng.uniform(-1, 1, B), sigma=64.0, beta=rng.uniform(-5, 5, B))], device="cuda")
def run():
    nm.invalidate_cache()
    return engine.solve_sv_rk4(nm, [("drag", 0, 0, 256)], p, y0, (0, 257 * DT), DT / 40).clone()
a = run()
os.environ["CASQ_SMALL_NO_CONST_TABLE"] = "1"
b = run()
print("CT vs table: max abs diff", float((a - b).abs().max()), "norm defect", float((a[:, -1].abs().pow(2).sum(-1) - 1).abs().max()))
