"""Generate the committed golden vectors under tests/golden/ FROM THE ORACLE (not from the reference:
sinel/casq cannot be imported here -- qiskit / qiskit-dynamics / jax are absent -- and it holds no golden
vectors of its own; parity is unpinned, see oracle/__init__.py).  The fixtures pin the oracle against
regressions and give the -m gpu tests inputs/outputs that do not need the oracle's slower integrators.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def main():
    rng = np.random.default_rng(20260101)
    dt = O.MANILA_DT
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, dt, rotating_frame=st)

    # 1. envelopes, all five families
    env = {}
    B = 6
    p = dict(amp=rng.uniform(0.05, 1.0, B), angle=rng.uniform(-np.pi, np.pi, B), sigma=rng.uniform(3, 40, B),
             width=rng.uniform(0, 48, B), beta=rng.uniform(-4, 4, B))
    for kind, keys in [("constant", ()), ("gaussian", ("sigma",)), ("drag", ("sigma", "beta")),
                       ("gaussian_square", ("sigma", "width")), ("gaussian_square_drag", ("sigma", "width", "beta"))]:
        env[kind] = O.envelope_samples(kind, 48, amp=p["amp"], angle=p["angle"], **{k: p[k] for k in keys})
    np.savez(os.path.join(HERE, "envelopes.npz"), duration=48, **{f"p_{k}": v for k, v in p.items()}, **env)

    # 2. headline-shaped RK4 case (cfg 2 at reduced duration): DRAG sweep, t_eval with 5 points
    B = 24
    p = dict(amp=rng.uniform(0.02, 0.5, B), angle=0.0, sigma=16.0, beta=rng.uniform(-5, 5, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 64)], [p])
    te = np.linspace(0, 65 * dt, 5)
    ts, y = O.rk4_solve(om, samples, np.array([1, 0, 0]), (0.0, 65 * dt), dt / 40, te)
    states = O.postprocess_statevectors(om, ts, y)
    np.savez(os.path.join(HERE, "rk4_drag_d3.npz"), amp=p["amp"], beta=p["beta"], sigma=16.0, duration=64, t_final=65,
             t_eval=te, raw=y, states=states)

    # 3. cfg 1: Gaussian amplitude sweep (Rabi), truth integrator + casq post-processing
    amps = np.linspace(0, 1, 11)
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian", "d0", 64)], [dict(amp=amps, angle=0.0, sigma=16.0)])
    ts, y = O.dop853_solve(om, samples, np.array([1, 0, 0]), (0.0, 64 * dt))
    np.savez(os.path.join(HERE, "rabi_gaussian_dop853.npz"), amp=amps, sigma=16.0, duration=64, raw=y,
             states=O.postprocess_statevectors(om, ts, y))

    # 4. cfg 5: Dopri5 (jax controller) candidates with the optimizer's documented options
    B = 8
    sig, beta = rng.uniform(6, 30, B), rng.uniform(-4, 4, B)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 64)], [dict(amp=0.45, angle=0.0, sigma=sig, beta=beta)])
    ts, y, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), (0.0, 64 * dt), rtol=1e-8, atol=1e-6, hmax=dt)
    np.savez(os.path.join(HERE, "dopri5_drag_d3.npz"), sigma=sig, beta=beta, amp=0.45, duration=64, raw=y, steps=steps)

    # 5. cfg 3 (reduced): 2 transmons, GaussianSquare on u0, exponential-midpoint propagators, RWA and no RWA
    hd = O.transmon_hamiltonian_dict({0: (O.MANILA["vars"]["wq0"], O.MANILA["vars"]["delta0"], O.MANILA["vars"]["omegad0"]),
                                      1: (O.MANILA["vars"]["wq1"], O.MANILA["vars"]["delta1"], O.MANILA["vars"]["omegad1"])},
                                     {(0, 1): O.MANILA["vars"]["jq0q1"]})
    st2, ops2, ch2, _ = O.parse_hamiltonian_dict(hd, None)
    B = 4
    p = dict(amp=rng.uniform(0.05, 0.8, B), angle=0.0, sigma=6.0, width=rng.uniform(0, 20, B))
    for tag, rwa, sub in (("rwa", 3.0e9, 1), ("norwa", None, 20)):
        m2 = O.OracleModel(st2, ops2, ch2, O.MANILA_CARRIERS, dt, rotating_frame=st2, rwa_cutoff_freq=rwa)
        samples = O.channel_samples(ch2, [O.PulseSpec("gaussian_square", "u0", 40)], [p])
        U = O.expm_midpoint_propagators(m2, samples, (0.0, 40 * dt), dt / sub)
        np.savez(os.path.join(HERE, f"expm_cr_d9_{tag}.npz"), amp=p["amp"], width=p["width"], sigma=6.0, duration=40,
                 substeps=sub, rwa_cutoff=np.nan if rwa is None else rwa, U=U)

    # 6. cfg 4 (reduced to 2 transmons, d=9): Lindblad T1/T2, diagonal frame, RWA
    a = np.diag(np.sqrt(np.arange(1, 3)), k=1).astype(complex)
    n = a.conj().T @ a
    eye = np.eye(3)
    T1, T2 = 100e-6, 80e-6
    gphi = 1 / T2 - 1 / (2 * T1)
    Ls = []
    for q in range(2):
        emb = (lambda m_, q=q: np.kron(m_, eye) if q == 1 else np.kron(eye, m_))
        Ls += [np.sqrt(1 / T1) * emb(a), np.sqrt(2 * gphi) * emb(n)]
    m3 = O.OracleModel(st2, ops2, ch2, O.MANILA_CARRIERS, dt, rotating_frame=np.diag(st2).real, rwa_cutoff_freq=3.0e9,
                       dissipators=np.array(Ls))
    B = 3
    p = dict(amp=np.linspace(0.2, 0.9, B), angle=0.0, sigma=8.0, beta=1.0)
    samples = O.channel_samples(ch2, [O.PulseSpec("drag", "d0", 32)], [p])
    rho0 = np.zeros(9, dtype=complex)
    rho0[0] = 1
    ts, rho = O.lindblad_rk4_solve(m3, samples, rho0, (0.0, 32 * dt), dt / 4)
    np.savez(os.path.join(HERE, "lindblad_d9.npz"), amp=p["amp"], sigma=8.0, beta=1.0, duration=32, T1=T1, T2=T2,
             rho=rho[:, -1], pops=np.real(np.einsum("bii->bi", O.postprocess_density_matrices(m3, ts, rho)[:, -1])))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
