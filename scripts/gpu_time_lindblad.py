"""BASELINE cfg 4 alone: 5 transmons (d = 243), T1/T2, DRAG on d0, RK4 dt/4 with RWA.  Usage: gpu_time_lindblad.py [B]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
st5, ops5, ch5 = P.manila(None)
nm5 = P.native(st5, ops5, ch5, frame="diag", rwa=0.75 * P.CARRIERS["d0"])
nm5.set_dissipators(P.t1t2(5))
B4 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
p4 = engine.pack_params([dict(amp=np.linspace(0.05, 0.5, B4), angle=0.0, sigma=64.0, beta=1.0)], device="cuda")
rho0 = np.zeros(243, dtype=complex); rho0[0] = 1
keep = {}
def run4():
    keep["r"] = engine.solve_lindblad_rk4(nm5, [("drag", ch5.index("d0"), 0, 256)], p4, rho0, (0, 257 * DT), DT / 4, want_rho=False)
best, mean = timed(run4, reps=1, warm=1)
pops = keep["r"][1][:, -1]
print(f"lindblad d=243 B={B4} S=1028: {best:.1f} ms -> {B4/best*1e3:.1f} traj/s, algorithmic {8*243*243*16*1028*B4/(best*1e-3)/1e9:.0f} GB/s; trace {float(pops.sum(-1).min()):.15f} P0 {float(pops[0,0]):.12f} {float(pops[-1,0]):.12f}")
