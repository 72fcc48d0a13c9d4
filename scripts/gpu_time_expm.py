"""BASELINE cfg 3 alone: NSxNS cross-resonance GaussianSquare unitaries (d = 9), expm-midpoint, RWA.  Usage: gpu_time_expm.py [NS]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
st2, ops2, ch2 = P.chain(2)
ww, aa3 = np.meshgrid(np.linspace(0, 768, NS), np.linspace(0.05, 0.8, NS), indexing="ij")
p3 = engine.pack_params([dict(amp=aa3.ravel(), angle=0.0, sigma=64.0, width=ww.ravel())], device="cuda")
nm2 = P.native(st2, ops2, ch2, rwa=0.75 * P.CARRIERS["d0"])
keep = {}
def run3():
    keep["U"] = engine.propagators_expm(nm2, [("gaussian_square", ch2.index("u0"), 0, 1024)], p3, (0, 1024 * DT), DT)
best, mean = timed(run3, reps=2, warm=1)
U = keep["U"]
print(f"expm d=9 B={NS*NS} S=1024: {best:.2f} ms -> {NS*NS/best*1e3:.0f} traj/s; unitarity {float((U.conj().transpose(1,2)@U - torch.eye(9, device='cuda', dtype=U.dtype)).abs().max()):.2e}; checksum {float(U.real.sum()):.12f}")
