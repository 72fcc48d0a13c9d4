"""GPU parity for the general-dimension RK4 path (coupled transmons, d = 5..27) against the oracle."""
import numpy as np
import pytest

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel

pytestmark = pytest.mark.gpu
DT = O.MANILA_DT
TOL_STATE = 1e-9


def _pair(h_dict, subsystems, frame, rwa, carriers=None):
    carriers = O.MANILA_CARRIERS if carriers is None else carriers
    st, ops, ch, _ = O.parse_hamiltonian_dict(h_dict, subsystems)
    rf = {"full": st, "diag": np.diag(st).real, "none": None}[frame]
    om = O.OracleModel(st, ops, ch, carriers, DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    nm = NativeModel(st, ops, [carriers[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    return om, nm, ch


@pytest.mark.parametrize("frame,rwa", [("full", None), ("diag", None), ("diag", 3.0e9), ("full", 3.0e9)])
def test_two_transmons_d9_cross_resonance(frame, rwa):
    """BASELINE cfg 3's model (qubits 0,1 + J01): GaussianSquare on u0 plus a DRAG on d1, random initial states."""
    rng = np.random.default_rng(17)
    om, nm, ch = _pair(O.MANILA, [0, 1], frame, rwa)
    assert om.dim == 9
    B = 21
    specs = [O.PulseSpec("gaussian_square", "u0", 32), O.PulseSpec("drag", "d1", 20, start=6)]
    ps = [dict(amp=rng.uniform(0.05, 0.8, B), angle=rng.uniform(-1, 1, B), sigma=rng.uniform(3, 8, B), width=rng.uniform(0, 20, B)),
          dict(amp=rng.uniform(0.05, 0.5, B), angle=0.0, sigma=rng.uniform(3, 6, B), beta=rng.uniform(-2, 2, B))]
    samples = O.channel_samples(ch, specs, ps)
    y0 = rng.normal(size=(B, 9)) + 1j * rng.normal(size=(B, 9))
    y0 /= np.linalg.norm(y0, axis=1, keepdims=True)
    span = (0.0, 32 * DT)
    te = np.linspace(0, 32 * DT, 4)
    _, want = O.rk4_solve(om, samples, y0, span, DT / 40, te)
    slots = [("gaussian_square", ch.index("u0"), 0, 32), ("drag", ch.index("d1"), 6, 20)]
    got = engine.solve_sv_rk4(nm, slots, engine.pack_params(ps, device="cuda"), y0, span, DT / 40, te).cpu().numpy()
    assert got.shape == (B, 4, 9) and np.abs(got - want).max() < TOL_STATE


def test_unitary_columns_d9():
    """Full propagator by integrating the d identity columns (y0 batched): U^+ U = 1 to RK4 accuracy."""
    om, nm, ch = _pair(O.MANILA, [0, 1], "full", None)
    p = dict(amp=np.full(9, 0.4), angle=0.0, sigma=5.0, width=12.0)
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian_square", "u0", 24)], [p])
    span = (0.0, 24 * DT)
    _, want = O.rk4_solve(om, samples, np.eye(9), span, DT / 40)
    got = engine.solve_sv_rk4(nm, [("gaussian_square", ch.index("u0"), 0, 24)], engine.pack_params([p], device="cuda"),
                              np.eye(9), span, DT / 40).cpu().numpy()
    assert np.abs(got - want).max() < TOL_STATE
    U = got[:, -1, :].T  # column b = image of basis state b
    assert np.abs(U.conj().T @ U - np.eye(9)).max() < 1e-6


def test_three_transmons_d27_and_five_level():
    rng = np.random.default_rng(23)
    om, nm, ch = _pair(O.MANILA, [0, 1, 2], "diag", None)
    assert om.dim == 27
    B = 7
    ps = [dict(amp=rng.uniform(0.1, 0.6, B), angle=0.0, sigma=4.0, beta=rng.uniform(-1, 1, B))]
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d1", 16)], ps)
    y0 = np.zeros(27, dtype=complex)
    y0[0] = 1
    span = (0.0, 16 * DT)
    _, want = O.rk4_solve(om, samples, y0, span, DT / 40)
    got = engine.solve_sv_rk4(nm, [("drag", ch.index("d1"), 0, 16)], engine.pack_params(ps, device="cuda"), y0, span, DT / 40).cpu().numpy()
    assert np.abs(got - want).max() < TOL_STATE
    # one 5-level Duffing oscillator, full (dense-eigenbasis-free) frame
    hd = {"h_str": ["w*O0", "al/2*O0*O0", "-al/2*O0", "om*X0||D0"], "qub": {"0": 5}, "vars": {"w": 3.1e10, "al": -2.1e9, "om": 9e8}}
    om5, nm5, ch5 = _pair(hd, None, "full", None, {"d0": 4.93e9})
    assert om5.dim == 5
    ps = [dict(amp=rng.uniform(0.2, 0.9, B), angle=0.2, sigma=5.0)]
    samples = O.channel_samples(ch5, [O.PulseSpec("gaussian", "d0", 20)], ps)
    y0 = np.array([1, 0, 0, 0, 0], dtype=complex)
    _, want = O.rk4_solve(om5, samples, y0, (0.0, 20 * DT), DT / 40)
    got = engine.solve_sv_rk4(nm5, [("gaussian", 0, 0, 20)], engine.pack_params(ps, device="cuda"), y0, (0.0, 20 * DT), DT / 40).cpu().numpy()
    assert np.abs(got - want).max() < TOL_STATE
