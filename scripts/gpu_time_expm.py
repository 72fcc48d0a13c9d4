import sys; sys.path.insert(0, ".")
import numpy as np, torch, oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel
from scripts.bench_configs import timed
DT = O.MANILA_DT
v = O.MANILA["vars"]
hd = O.transmon_hamiltonian_dict({0: (v["wq0"], v["delta0"], v["omegad0"]), 1: (v["wq1"], v["delta1"], v["omegad1"])}, {(0, 1): v["jq0q1"]})
st2, ops2, ch2, _ = O.parse_hamiltonian_dict(hd, None)
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ww, aa3 = np.meshgrid(np.linspace(0, 768, NS), np.linspace(0.05, 0.8, NS), indexing="ij")
p3 = engine.pack_params([dict(amp=aa3.ravel(), angle=0.0, sigma=64.0, width=ww.ravel())], device="cuda")
nm2 = NativeModel(st2, ops2, [O.MANILA_CARRIERS[c] for c in ch2], DT, rotating_frame=st2, rwa_cutoff_freq=0.75 * O.MANILA_CARRIERS["d0"])
keep = {}
def run3():
    keep["U"] = engine.propagators_expm(nm2, [("gaussian_square", ch2.index("u0"), 0, 1024)], p3, (0, 1024 * DT), DT)
best, mean = timed(run3, reps=2, warm=1)
U = keep["U"]
print(f"expm d=9 B={NS*NS} S=1024: {best:.2f} ms -> {NS*NS/best*1e3:.0f} traj/s; unitarity {float((U.conj().transpose(1,2)@U - torch.eye(9, device='cuda', dtype=U.dtype)).abs().max()):.2e}; checksum {float(U.real.sum()):.12f}")
