import sys; sys.path.insert(0, ".")
import numpy as np, torch, oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel
DT=O.MANILA_DT
st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st)
for kind, dur, tf in (("constant", 40, 24), ("drag", 24, 24)):
    p = dict(amp=0.6, angle=0.3, sigma=6.0, beta=1.2) if kind=="drag" else dict(amp=0.6, angle=0.3)
    samples = O.channel_samples(ch, [O.PulseSpec(kind,"d0",dur)],[p])
    span=(0.0, tf*DT)
    for tol in (1e-6, 1e-8, 1e-10):
        ts,y,s1 = O.dopri5_jax_solve(om, samples, np.array([1,0,0]), span, rtol=tol, atol=tol)
        got, ns, status = engine.solve_sv_dopri5(nm, [(kind,0,0,dur)], engine.pack_params([p], device="cuda"), np.array([1,0,0]), span, rtol=tol, atol=tol, return_stats=True)
        print(kind, tol, "oracle steps", s1[0], "cuda steps", ns.cpu().numpy()[0], "diff", np.abs(got.cpu().numpy()-y).max())
