"""State-vector steppers beyond d = 4 (lane-group RK4 and warp-per-trajectory Dopri5): two and three coupled transmons.
Usage: gpu_time_general.py [B]"""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from casq_b200 import engine
from scripts import _problems as P
from scripts.bench_configs import timed
DT = P.DT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(20260101)
for nq in (2, 3):
    st, ops, ch = P.chain(nq)
    d = st.shape[0]
    nm = P.native(st, ops, ch)
    p = engine.pack_params([dict(amp=rng.uniform(0.05, 0.5, B), angle=0.0, sigma=64.0, beta=rng.uniform(-2, 2, B))], device="cuda")
    y0 = np.zeros(d, dtype=complex); y0[0] = 1
    slots = [("drag", ch.index("d0"), 0, 256)]
    keep = {}
    def rk4():
        keep["y"] = engine.solve_sv_rk4(nm, slots, p, y0, (0, 257 * DT), DT / 40)
    best, _ = timed(rk4, reps=2, warm=1)
    nrm = float((torch.linalg.vector_norm(keep["y"][:, -1], dim=1) - 1).abs().max())
    print(f"rk4 general d={d} B={B} S=10280: {best:.2f} ms -> {B/best*1e3:.0f} traj/s; norm defect {nrm:.2e}")
    def dp5():
        keep["r"] = engine.solve_sv_dopri5(nm, slots, p, y0, (0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
    best, _ = timed(dp5, reps=2, warm=1)
    y, ns, stt = keep["r"]
    print(f"dopri5 general d={d} B={B}: {best:.2f} ms -> {B/best*1e3:.0f} traj/s; attempted {int(ns[:,0].sum())} failed {int((stt!=0).sum())}")
