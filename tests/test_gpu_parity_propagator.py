"""GPU parity for the batched expm propagator path (BASELINE cfg 3) against the oracle and its golden vectors."""
import os

import numpy as np
import pytest

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = O.MANILA_DT
TOL_U = 1e-9  # both sides run the same exponential-midpoint scheme; expm implementations agree to ~1e-13


def _two_transmon_dict():
    v = O.MANILA["vars"]
    return O.transmon_hamiltonian_dict({0: (v["wq0"], v["delta0"], v["omegad0"]), 1: (v["wq1"], v["delta1"], v["omegad1"])},
                                       {(0, 1): v["jq0q1"]})


@pytest.mark.parametrize("tag", ["rwa", "norwa"])
def test_cross_resonance_propagators_golden(tag):
    g = np.load(os.path.join(G, f"expm_cr_d9_{tag}.npz"))
    st, ops, ch, _ = O.parse_hamiltonian_dict(_two_transmon_dict(), None)
    rwa = None if np.isnan(g["rwa_cutoff"]) else float(g["rwa_cutoff"])
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st, rwa_cutoff_freq=rwa)
    p = dict(amp=g["amp"], angle=0.0, sigma=float(g["sigma"]), width=g["width"])
    dur = int(g["duration"])
    U = engine.propagators_expm(nm, [("gaussian_square", ch.index("u0"), 0, dur)], engine.pack_params([p], device="cuda"),
                                (0.0, dur * DT), DT / int(g["substeps"])).cpu().numpy()
    assert U.shape == g["U"].shape and np.abs(U - g["U"]).max() < TOL_U
    eye = np.eye(9)
    assert max(np.abs(u.conj().T @ u - eye).max() for u in U) < 1e-11  # expm of an anti-Hermitian generator


@pytest.mark.parametrize("frame", ["full", "diag", "none"])
def test_propagators_match_oracle_d3_all_pade_orders(frame):
    """Single transmon; max_dt spans three decades so every Pade order (3..13 with squarings) is exercised."""
    rng = np.random.default_rng(4)
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    rf = {"full": st, "diag": np.diag(st).real + 3e8, "none": None}[frame]
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf)
    B, dur = 5, 12
    p = dict(amp=rng.uniform(0.2, 1.0, B), angle=rng.uniform(-1, 1, B), sigma=4.0, beta=rng.uniform(-2, 2, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p])
    for max_dt in (DT / 64, DT / 4, DT, 4 * DT):
        want = O.expm_midpoint_propagators(om, samples, (0.0, dur * DT), max_dt)
        got = engine.propagators_expm(nm, [("drag", 0, 0, dur)], engine.pack_params([p], device="cuda"), (0.0, dur * DT), max_dt).cpu().numpy()
        assert np.abs(got - want).max() < TOL_U, (frame, max_dt)


def test_propagator_agrees_with_ode_solution():
    """U(T)|psi0> from the propagator path converges to the tight-tolerance ODE truth as substeps grow."""
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st)
    p = dict(amp=0.5, angle=0.0, sigma=6.0, beta=1.0)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 24)], [p])
    _, truth = O.dop853_solve(om, samples, np.array([1, 0, 0]), (0.0, 24 * DT))
    errs = []
    for sub in (10, 20, 40):
        U = engine.propagators_expm(nm, [("drag", 0, 0, 24)], engine.pack_params([p], device="cuda"), (0.0, 24 * DT), DT / sub).cpu().numpy()
        errs.append(np.abs(U[0] @ np.array([1, 0, 0]) - truth[0, -1]).max())
    assert errs[0] / errs[1] > 3 and errs[1] / errs[2] > 3 and errs[2] < 1e-4  # second-order scheme


@pytest.mark.parametrize("rwa", [None, 3.0e9])
def test_propagators_match_oracle_d9_all_pade_orders(rwa):
    """Two coupled transmons (the compile-time d = 9 kernel with the register-resident Gauss-Jordan), two driven channels;
    max_dt spans the Pade orders 3..9 and the scaled-and-squared regime; B = 7 leaves the last block partly empty."""
    rng = np.random.default_rng(11)
    st, ops, ch, _ = O.parse_hamiltonian_dict(_two_transmon_dict(), None)
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st, rwa_cutoff_freq=rwa)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st, rwa_cutoff_freq=rwa)
    B, dur = 7, 10
    specs = [O.PulseSpec("gaussian_square", "u0", dur), O.PulseSpec("drag", "d1", dur)]
    ps = [dict(amp=rng.uniform(0.2, 1.0, B), angle=rng.uniform(-1, 1, B), sigma=3.0, width=rng.uniform(0, 6, B)),
          dict(amp=rng.uniform(0.1, 0.6, B), angle=rng.uniform(-1, 1, B), sigma=3.0, beta=rng.uniform(-2, 2, B))]
    samples = O.channel_samples(ch, specs, ps)
    slots = [("gaussian_square", ch.index("u0"), 0, dur), ("drag", ch.index("d1"), 0, dur)]
    for max_dt in (DT / 64, DT / 8, DT, 5 * DT):
        want = O.expm_midpoint_propagators(om, samples, (0.0, dur * DT), max_dt)
        got = engine.propagators_expm(nm, slots, engine.pack_params(ps, device="cuda"), (0.0, dur * DT), max_dt).cpu().numpy()
        assert np.abs(got - want).max() < TOL_U, (rwa, max_dt)


def test_propagators_match_oracle_d27_three_transmons():
    """dim 27 (three coupled transmons): the block-per-trajectory kernel with 3x3 register tiles; unitaries of a
    DRAG pulse on d1 against scipy's expm per step, with and without RWA."""
    rng = np.random.default_rng(21)
    v = O.MANILA["vars"]
    qm = {q: (v[f"wq{q}"], v[f"delta{q}"], v[f"omegad{q}"]) for q in range(3)}
    hd = O.transmon_hamiltonian_dict(qm, {(0, 1): v["jq0q1"], (1, 2): v["jq1q2"]})
    st, ops, ch, _ = O.parse_hamiltonian_dict(hd, None)
    assert st.shape == (27, 27)
    B, dur = 3, 6
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=2.0, beta=rng.uniform(-1, 1, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d1", dur)], [p])
    for rwa, max_dt in ((None, DT / 8), (3.0e9, DT), (None, 2 * DT)):
        om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st, rwa_cutoff_freq=rwa)
        nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=st, rwa_cutoff_freq=rwa)
        want = O.expm_midpoint_propagators(om, samples, (0.0, dur * DT), max_dt)
        got = engine.propagators_expm(nm, [("drag", ch.index("d1"), 0, dur)], engine.pack_params([p], device="cuda"), (0.0, dur * DT),
                                      max_dt).cpu().numpy()
        assert got.shape == (B, 27, 27) and np.abs(got - want).max() < TOL_U, (rwa, max_dt)
        assert max(np.abs(u.conj().T @ u - np.eye(27)).max() for u in got) < 1e-11


@pytest.mark.parametrize("d", [20, 40, 64])
def test_propagators_match_oracle_random_hermitian_models_up_to_dim_64(d):
    """16 < dim <= 64: the block-per-trajectory kernel (tiles of 2 at even dims; work matrices in shared memory up to
    dim 32, in L2-resident global scratch above).  Random Hermitian H0 / drive with GHz-scale spectrum, full frame."""
    rng = np.random.default_rng(100 + d)
    def herm(scale):
        m = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        return scale * (m + m.conj().T) / 2
    h0 = np.diag(np.sort(rng.uniform(0, 2 * np.pi * 5e9, d))).astype(complex) + herm(2 * np.pi * 5e6)
    x = herm(2 * np.pi * 40e6)
    carriers = {"d0": 4.9e9}
    om = O.OracleModel(h0, x[None], ["d0"], carriers, DT, rotating_frame=h0)
    nm = NativeModel(h0, x[None], [carriers["d0"]], DT, rotating_frame=h0)
    B, dur = 2, 4
    p = dict(amp=rng.uniform(0.3, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=2.0)
    samples = O.channel_samples(["d0"], [O.PulseSpec("gaussian", "d0", dur)], [p])
    for max_dt in (DT / 4, 2 * DT):
        want = O.expm_midpoint_propagators(om, samples, (0.0, dur * DT), max_dt)
        got = engine.propagators_expm(nm, [("gaussian", 0, 0, dur)], engine.pack_params([p], device="cuda"), (0.0, dur * DT),
                                      max_dt).cpu().numpy()
        assert got.shape == (B, d, d) and np.abs(got - want).max() < TOL_U, (d, max_dt)
        assert max(np.abs(u.conj().T @ u - np.eye(d)).max() for u in got) < 1e-10
