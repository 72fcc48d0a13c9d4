import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
st, ops, ch = P.chain(3)
nm = P.native(st, ops, ch, rwa=0.75 * P.CARRIERS["d0"])
B = 1024
rng = np.random.default_rng(1)
p = engine.pack_params([dict(amp=rng.uniform(0.05, 0.8, B), angle=0.0, sigma=64.0, width=rng.uniform(0, 768, B))], device="cuda")
keep = {}
def run():
    keep["U"] = engine.propagators_expm(nm, [("gaussian_square", ch.index("u0"), 0, 1024)], p, (0, 1024 * P.DT), P.DT)
best, _ = timed(run, reps=2, warm=1)
U = keep["U"]
print(f"expm d=27 B={B} S=1024: {best:.1f} ms -> {B/best*1e3:.0f} traj/s; unitarity {float((U.conj().transpose(1,2)@U - torch.eye(27, device='cuda', dtype=U.dtype)).abs().max()):.2e}")
