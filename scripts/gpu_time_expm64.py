"""Propagators at d = 64 (random Hermitian model, global-scratch work matrices): usage gpu_time_expm64.py [B]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from casq_b200._native import NativeModel
from scripts.bench_configs import timed
d, B = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(4)
def herm(scale):
    a = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    return scale * (a + a.conj().T) / 2
H0 = np.diag(rng.uniform(0, 3e10, d)) + herm(2e7)
Hk = herm(3e8)[None]
nm = NativeModel(H0, Hk, [4.9e9], 2.2222222222222221e-10, rotating_frame=np.diag(H0).real, rwa_cutoff_freq=None)
p = engine.pack_params([dict(amp=rng.uniform(0.05, 0.8, B), angle=0.0, sigma=16.0, beta=0.5)], device="cuda")
keep = {}
def run():
    keep["U"] = engine.propagators_expm(nm, [("drag", 0, 0, 64)], p, (0, 64 * 2.2222222222222221e-10), 2.2222222222222221e-10 / 4)
best, _ = timed(run, reps=2, warm=1)
U = keep["U"]
print(f"expm d=64 B={B} S=256: {best:.1f} ms -> {B/best*1e3:.0f} traj/s; unitarity {float((U.conj().transpose(1,2)@U - torch.eye(d, device='cuda', dtype=U.dtype)).abs().max()):.2e}")
