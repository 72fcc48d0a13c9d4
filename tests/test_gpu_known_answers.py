"""Closed-form answers checked DIRECTLY against the CUDA path (through the C ABI, no oracle in between): the oracle is a
restatement of un-vendored upstream code ("parity unpinned"), so these library-independent anchors are what
ties the kernels themselves to physics.  Every expected value below is a formula, not a computed reference.
Mirrors tests/test_oracle_known_answers.py, which pins the oracle with the same formulas."""
import numpy as np
import pytest
import torch
from scipy.linalg import expm

from casq_b200 import engine
from casq_b200._native import NativeModel

pytestmark = pytest.mark.gpu
DT = 1e-10
WQ = 2 * np.pi * 5e9   # qubit angular frequency
OM = 2 * np.pi * 50e6  # drive strength


def _two_level(rwa=7e9, frame=True):
    h0 = np.diag([0.0, WQ]).astype(complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex) * OM
    return NativeModel(h0, x[None], [5e9], DT, rotating_frame=h0 if frame else None, rwa_cutoff_freq=rwa)


def test_resonant_constant_drive_is_an_x_rotation_rk4_dopri5_expm():
    """2-level system, resonant constant drive, RWA: y(T) = (cos(theta/2), -i sin(theta/2)), theta = Omega * amp * T, for every
    amplitude of a batch; RK4 (fine steps), adaptive Dopri5 and the expm propagators must all land on the formula."""
    nm = _two_level()
    N = 200
    amp = np.linspace(0.05, 0.9, 16)
    T = N * DT
    theta = OM * amp * T
    want = np.stack([np.cos(theta / 2), -1j * np.sin(theta / 2)], axis=1)
    p = engine.pack_params([dict(amp=amp, angle=0.0)], device="cuda")
    slots = [("constant", 0, 0, N + 1)]  # one spare sample: the signal is 0 at t >= len * dt
    y0 = np.array([1, 0])
    rk4 = engine.solve_sv_rk4(nm, slots, p, y0, (0.0, T), DT / 8)[:, -1].cpu().numpy()
    assert np.abs(rk4 - want).max() < 1e-9
    dp5 = engine.solve_sv_dopri5(nm, slots, p, y0, (0.0, T), rtol=1e-11, atol=1e-11)[:, -1].cpu().numpy()
    assert np.abs(dp5 - want).max() < 1e-8
    U = engine.propagators_expm(nm, slots, p, (0.0, T), DT).cpu().numpy()
    assert np.abs(U[:, :, 0] - want).max() < 1e-11  # piecewise-constant generator: the exponential is exact
    # the envelope phase rotates the axis: angle = pi/2 turns the X rotation into a Y rotation
    p90 = engine.pack_params([dict(amp=amp, angle=np.pi / 2)], device="cuda")
    Uy = engine.propagators_expm(nm, slots, p90, (0.0, T), DT).cpu().numpy()
    assert np.abs(np.abs(Uy[:, 0, 0]) - np.abs(np.cos(theta / 2))).max() < 1e-11
    assert np.abs(Uy[:, 1, 0].imag).max() < 1e-9  # rotation about +-Y: a REAL amplitude on |1>
    assert np.abs(np.abs(Uy[:, 1, 0].real) - np.abs(np.sin(theta / 2))).max() < 1e-11


def test_lab_frame_constant_generator_matches_matrix_exponential():
    """No frame, zero carrier: H is constant, y(T) = expm(-i H T) y0 exactly (3 levels, non-trivial couplings)."""
    h0 = np.diag([0.0, 3e8, 5e8]).astype(complex)
    x = np.array([[0, 1, 0], [1, 0, 1.4], [0, 1.4, 0]], dtype=complex) * 2e8
    nm = NativeModel(h0, x[None], [0.0], DT, rotating_frame=None)
    N = 50
    T = N * DT
    amp = np.array([0.7, 0.2, -0.5])
    p = engine.pack_params([dict(amp=amp, angle=0.0)], device="cuda")
    slots = [("constant", 0, 0, N + 1)]
    y0 = np.array([1, 0, 0])
    want = np.stack([expm(-1j * (h0 + a * x) * T) @ y0 for a in amp])
    rk4 = engine.solve_sv_rk4(nm, slots, p, y0, (0.0, T), DT / 50)[:, -1].cpu().numpy()
    assert np.abs(rk4 - want).max() < 1e-10
    U = engine.propagators_expm(nm, slots, p, (0.0, T), DT).cpu().numpy()
    assert np.abs(U @ y0 - want).max() < 1e-12
    dp5 = engine.solve_sv_dopri5(nm, slots, p, y0, (0.0, T), rtol=1e-12, atol=1e-12)[:, -1].cpu().numpy()
    assert np.abs(dp5 - want).max() < 1e-9


def test_lindblad_amplitude_damping_and_dephasing_closed_forms():
    """Free evolution with T1 and pure dephasing on a 2-level system in its diagonal frame:
    P1(t) = P1(0) e^{-g1 t}; coherence |rho01(t)| = |rho01(0)| e^{-(g1/2 + g_phi) t} (L_phi = sqrt(2 g_phi) n); trace = 1."""
    h0 = np.diag([0.0, WQ]).astype(complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex) * OM
    nm = NativeModel(h0, x[None], [5e9], DT, rotating_frame=np.diag(h0).real)
    g1, gphi = 5e7, 3e7
    a = np.array([[0, 1], [0, 0]], dtype=complex)
    n = np.diag([0.0, 1.0]).astype(complex)
    nm.set_dissipators(np.array([np.sqrt(g1) * a, np.sqrt(2 * gphi) * n]))
    N = 100
    T = N * DT
    psi = np.array([0.6, 0.8j])
    rho0 = np.outer(psi, psi.conj())
    p = engine.pack_params([dict(amp=np.zeros(3), angle=0.0)], device="cuda")
    te = np.linspace(0, T, 5)
    rho, pops = engine.solve_lindblad_rk4(nm, [("constant", 0, 0, N)], p, rho0, (0.0, T), DT / 4, t_eval=te)
    rho = rho.cpu().numpy()
    for k, t in enumerate(te):
        r = rho[:, k]
        assert np.abs(np.trace(r, axis1=1, axis2=2) - 1).max() < 1e-12
        assert np.abs(r[:, 1, 1].real - 0.64 * np.exp(-g1 * t)).max() < 1e-9
        assert np.abs(np.abs(r[:, 0, 1]) - 0.48 * np.exp(-(g1 / 2 + gphi) * t)).max() < 1e-9
    assert np.abs(pops.cpu().numpy()[:, -1, 1] - 0.64 * np.exp(-g1 * T)).max() < 1e-9


def test_detuned_rabi_formula_with_rwa():
    """Drive detuned by Delta from the qubit (carrier = qubit + Delta/2pi): P1(T) = Omega_a^2/W^2 sin^2(W T/2),
    W = sqrt(Omega_a^2 + Delta^2), Omega_a = Omega * amp, under the RWA."""
    delta = 2 * np.pi * 20e6
    h0 = np.diag([0.0, WQ]).astype(complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex) * OM
    nu = (WQ + delta) / (2 * np.pi)
    nm = NativeModel(h0, x[None], [nu], DT, rotating_frame=h0, rwa_cutoff_freq=7e9)
    N = 300
    T = N * DT
    amp = np.linspace(0.1, 1.0, 10)
    oa = OM * amp
    W = np.sqrt(oa ** 2 + delta ** 2)
    want_p1 = oa ** 2 / W ** 2 * np.sin(W * T / 2) ** 2
    p = engine.pack_params([dict(amp=amp, angle=0.0)], device="cuda")
    slots = [("constant", 0, 0, N + 1)]
    y = engine.solve_sv_dopri5(nm, slots, p, np.array([1, 0]), (0.0, T), rtol=1e-11, atol=1e-11)[:, -1]
    assert np.abs((y.abs() ** 2).cpu().numpy()[:, 1] - want_p1).max() < 1e-8
    U = engine.propagators_expm(nm, slots, p, (0.0, T), DT / 8)
    # the carrier phase makes the generator time dependent in the frame of H0: the midpoint rule converges as h^2
    assert np.abs((U[:, 1, 0].abs() ** 2).cpu().numpy() - want_p1).max() < 1e-5
