"""Scratch GPU check of the small RK4 kernel against the oracle + first timing (not a test, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle as O
from casq_b200._native import NativeModel
from casq_b200 import engine as E

st, ops, ch, dims = O.parse_hamiltonian_dict(O.MANILA, [0])
nu = [O.MANILA_CARRIERS[c] for c in ch]
dt = O.MANILA_DT
for rwa in (None, 4.9e9):
    for frame in ("H0", "none"):
        rf = st if frame == "H0" else None
        om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, dt, rotating_frame=rf, rwa_cutoff_freq=rwa)
        m = NativeModel(st, ops, nu, dt, rotating_frame=rf, rwa_cutoff_freq=rwa)
        rng = np.random.default_rng(1)
        B = 37
        p = dict(amp=rng.uniform(0.02, 0.5, B), angle=rng.uniform(-1, 1, B), sigma=rng.uniform(8, 20, B), beta=rng.uniform(-3, 3, B))
        samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 48)], [p])
        tspan = (0.0, 48 * dt)
        te = np.linspace(0, 48 * dt, 7)
        md = dt / (40 if frame == "H0" else 200)
        _, yo = O.rk4_solve(om, samples, np.array([1, 0, 0]), tspan, md, te)
        params = E.pack_params([p], device="cuda")
        y = E.solve_sv_rk4(m, [("drag", 0, 0, 48)], params, np.array([1, 0, 0]), tspan, md, te).cpu().numpy()
        print(f"rwa={rwa} frame={frame}: max |cuda - oracle| = {np.abs(y - yo).max():.3e}  norm dev {np.abs(np.linalg.norm(y[:,-1],axis=-1)-1).max():.2e}")

# headline timing
om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, dt, rotating_frame=st)
m = NativeModel(st, ops, nu, dt, rotating_frame=st)
beta = np.linspace(-5, 5, 250); amp = np.linspace(0.02, 0.5, 400)
bb, aa = np.meshgrid(beta, amp, indexing="ij")
p = dict(amp=aa.ravel(), angle=0.0, sigma=64.0, beta=bb.ravel())
params = E.pack_params([p], device="cuda")
slots = [("drag", 0, 0, 256)]
tspan = (0.0, 257 * dt)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    y = E.solve_sv_rk4(m, slots, params, np.array([1, 0, 0]), tspan, dt / 40)
    torch.cuda.synchronize(); el = time.time() - t0
    print(f"headline 1e5 x 10280 steps: {el*1e3:.2f} ms -> {1e5/el:.3e} traj/s")
print("fp64 probe TFLOP/s:", E.fp64_fma_probe())
sub = np.arange(0, 100000, 5003)
samples = O.channel_samples(ch, slots and [O.PulseSpec("drag", "d0", 256)], [dict(amp=aa.ravel()[sub], angle=0.0, sigma=64.0, beta=bb.ravel()[sub])])
t0 = time.time(); _, yo = O.rk4_solve(om, samples, np.array([1, 0, 0]), tspan, dt / 40); print("oracle time", time.time() - t0)
print("headline parity (20 pts): ", np.abs(y.cpu().numpy()[sub] - yo).max())
