import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
st, ops, ch = P.manila([0])
nm = P.native(st, ops, ch)
y0 = np.array([1, 0, 0])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(1)
p = engine.pack_params([dict(amp=rng.uniform(0.02, 0.5, B), angle=0.0, sigma=64.0, beta=rng.uniform(-5, 5, B))], device="cuda")
best, mean = timed(lambda: engine.solve_sv_rk4(nm, [("drag", 0, 0, 256)], p, y0, (0, 257 * DT), DT / 40), reps=5, warm=2)
print(f"B={B} SLICED={os.environ.get('CASQ_SMALL_SLICED')} CHUNKS={os.environ.get('CASQ_SMALL_SLICE_CHUNKS')}: {best:.3f} ms")
