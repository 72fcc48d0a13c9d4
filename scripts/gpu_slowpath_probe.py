import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from casq_b200._native import NativeModel
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
bb, aa = np.meshgrid(np.linspace(-5, 5, 250), np.linspace(0.02, 0.5, 400), indexing="ij")
y0 = np.array([1, 0, 0])
for kind in ("drag", "gaussian", "constant"):
    p = engine.pack_params([dict(amp=aa.ravel(), angle=0.0, sigma=64.0, beta=bb.ravel())], device="cuda")
    best, mean = timed(lambda: engine.solve_sv_rk4(nm, [(kind, 0, 0, 256)], p, y0, (0, 257 * DT), DT / 40), reps=5, warm=2)
    print(kind, f"{best:.3f} ms")
# same work with a coarser sample grid (duration 32 -> 8x fewer boundaries) but identical step count
nm8 = NativeModel(st, ops, [P.CARRIERS[c] for c in ch], DT * 8, rotating_frame=st)
p = engine.pack_params([dict(amp=aa.ravel(), angle=0.0, sigma=8.0, beta=bb.ravel())], device="cuda")
best, mean = timed(lambda: engine.solve_sv_rk4(nm8, [("drag", 0, 0, 32)], p, y0, (0, 257 * DT), DT / 40), reps=5, warm=2)
print("drag, 8x coarser samples (same 10280 steps)", f"{best:.3f} ms")
