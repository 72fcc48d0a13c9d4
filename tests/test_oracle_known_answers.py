"""The oracle has no reference golden vectors to lean on (parity unpinned), so it is pinned by
library-independent known answers instead."""
import numpy as np
import pytest
from scipy.linalg import expm

import oracle as O
from oracle import solvers


def _two_level_model(omega=2 * np.pi * 5e9, rabi=2 * np.pi * 50e6, frame=True, rwa=None, nu=5e9):
    h0 = np.diag([0.0, omega]).astype(complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    return O.OracleModel(h0, rabi * x[None], ["d0"], {"d0": nu}, 1e-10, rotating_frame=h0 if frame else None,
                         rwa_cutoff_freq=rwa)


def test_parser_matches_hand_built_transmon():
    st, ops, ch, dims = O.parse_hamiltonian_dict(O.MANILA, [0])
    wq, dl, om = O.MANILA["vars"]["wq0"], O.MANILA["vars"]["delta0"], O.MANILA["vars"]["omegad0"]
    assert ch == ["d0", "u0"] and dims == {0: 3}
    np.testing.assert_allclose(st, np.diag([0, wq, 2 * wq + dl]), rtol=1e-15)
    x = np.array([[0, 1, 0], [1, 0, np.sqrt(2)], [0, np.sqrt(2), 0]])
    np.testing.assert_allclose(ops[0], om * x, rtol=1e-15)
    np.testing.assert_allclose(ops[1], O.MANILA["vars"]["omegad1"] * x, rtol=1e-15)


def test_parser_little_endian_and_coupling():
    st, ops, ch, dims = O.parse_hamiltonian_dict(O.MANILA, [0, 1])
    assert st.shape == (9, 9) and ch == ["d0", "d1", "u0", "u1", "u2"]
    # index = n1*3 + n0: |n1=0,n0=1> is index 1 and carries wq0; |n1=1,n0=0> is index 3 and carries wq1
    assert st[1, 1].real == pytest.approx(O.MANILA["vars"]["wq0"]) and st[3, 3].real == pytest.approx(O.MANILA["vars"]["wq1"])
    assert st[1, 3].real == pytest.approx(O.MANILA["vars"]["jq0q1"])  # flip-flop <01|H|10>
    assert np.allclose(st, st.conj().T)


def test_envelope_known_values():
    g = O.envelope_samples("gaussian", 16, amp=1.0, sigma=8.0)
    assert g.shape == (16,) and np.allclose(g, g[::-1]) and np.all(np.diff(g[:8].real) > 0)
    # lifted gaussian reaches 0 one sample outside the window: envelope(t=-1) = 0
    t, c = -1.0, 8.0
    off = np.exp(-(((17 - c) / 8.0) ** 2) / 2)
    assert abs((np.exp(-(((t - c) / 8.0) ** 2) / 2) - off) / (1 - off)) < 1e-15
    assert np.allclose(O.envelope_samples("constant", 5, amp=0.3, angle=0.5), 0.3 * np.exp(0.5j))
    d = O.envelope_samples("drag", 16, amp=1.0, sigma=8.0, beta=2.0)
    assert np.allclose(d.real, g.real) and np.allclose(d.imag, -d.imag[::-1])  # odd quadrature part
    gs = O.envelope_samples("gaussian_square", 32, amp=0.5, sigma=4.0, width=16.0)
    assert np.allclose(gs[8:24], 0.5) and gs[0].real < gs[7].real < 0.5
    gsd = O.envelope_samples("gaussian_square_drag", 32, amp=0.5, sigma=4.0, width=16.0, beta=0.0)
    assert np.allclose(gsd, gs)
    batch = O.envelope_samples("drag", 16, amp=np.array([0.1, 0.2]), sigma=np.array([4.0, 8.0]), beta=np.array([0.0, 1.0]))
    assert batch.shape == (2, 16) and np.allclose(batch[1], 0.2 * O.envelope_samples("drag", 16, 1.0, sigma=8.0, beta=1.0))
    with pytest.raises(ValueError):
        O.envelope_samples("gaussian", 16, amp=1.0, sigma=0.0)
    with pytest.raises(ValueError):
        O.envelope_samples("gaussian_square", 16, amp=1.0, sigma=1.0, width=17.0)


def test_floor_divide_index_is_python_floor_division():
    dt = O.MANILA_DT
    ts = np.concatenate([np.arange(0, 300) * dt, np.arange(0, 300) * dt * (1 - 1e-16), np.array([-dt, 1e-30, 400 * dt])])
    got = O.floor_divide_index(ts, dt, 256)
    want = np.clip([int(t // dt) for t in ts], -1, 256)
    assert np.array_equal(got, want)


def test_fixed_step_sizes_rule():
    t_list, h, n = O.fixed_step_sizes((0.0, 1.0), None, 0.3)
    assert list(n) == [4] and h[0] == pytest.approx(0.25)
    t_list, h, n = O.fixed_step_sizes((0.0, 1.0), [0.0, 0.5, 1.0], 0.25)
    assert list(n) == [2, 2]
    t_list, h, n = O.fixed_step_sizes((0.0, 1.0), None, 5.0)
    assert list(n) == [1]
    # 257*dt / (dt/40) must give exactly 10280 steps (the benchmark's S)
    _, _, n = O.fixed_step_sizes((0.0, 257 * O.MANILA_DT), None, O.MANILA_DT / 40)
    assert int(n[0]) == 10280


def test_constant_resonant_drive_matches_closed_form_rwa_rotation():
    """2-level system, resonant constant drive, RWA: rotation by theta = Omega * amp * T about X."""
    m = _two_level_model(rwa=7e9)
    N, amp = 200, 0.4
    samples = np.full((1, 1, N), amp, dtype=complex)
    T = N * m.dt
    _, y = O.dop853_solve(m, samples, np.array([1, 0]), (0.0, T))
    theta = 2 * np.pi * 50e6 * amp * T
    want = np.array([np.cos(theta / 2), -1j * np.sin(theta / 2)])
    np.testing.assert_allclose(y[0, -1], want, atol=1e-9)


def test_lab_frame_constant_generator_matches_expm():
    """No frame, zero carrier: H is constant, y(T) = expm(-i H T) y0 exactly."""
    h0 = np.diag([0.0, 3e8, 5e8]).astype(complex)
    x = np.array([[0, 1, 0], [1, 0, 1.4], [0, 1.4, 0]], dtype=complex) * 2e8
    m = O.OracleModel(h0, x[None], ["d0"], {"d0": 0.0}, 1e-10)
    N = 50
    samples = np.full((1, 1, N + 1), 0.7, dtype=complex)  # one spare sample: the signal is 0 at t >= len*dt
    T = N * m.dt
    want = expm(-1j * (h0 + 0.7 * x) * T) @ np.array([1, 0, 0])
    _, y = O.rk4_solve(m, samples, np.array([1, 0, 0]), (0.0, T), m.dt / 50)
    np.testing.assert_allclose(y[0, -1], want, atol=1e-10)
    U = O.expm_midpoint_propagators(m, samples, (0.0, T), m.dt)
    np.testing.assert_allclose(U[0] @ np.array([1, 0, 0]), want, atol=1e-12)


def test_frame_independence(manila_q0):
    """Lab-frame state must not depend on the rotating frame used for integration."""
    om, st, ops, ch = manila_q0
    lab = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=None)
    diag = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=np.diag(st).real + 1e8)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 24)], [dict(amp=0.6, angle=0.3, sigma=6.0, beta=1.5)])
    span = (0.0, 24 * O.MANILA_DT)
    outs = []
    for m in (om, lab, diag):
        ts, y = O.dop853_solve(m, samples, np.array([1, 0, 0]), span)
        outs.append(O.postprocess_statevectors(m, ts, y)[0, -1])
    np.testing.assert_allclose(outs[0], outs[1], atol=2e-9)
    np.testing.assert_allclose(outs[0], outs[2], atol=2e-9)


def test_rabi_pi_pulse_amplitude(manila_q0):
    """lifted Gaussian dur 256 sigma 64 on manila q0 is a pi-pulse at amp ~ 0.1109."""
    om, st, ops, ch = manila_q0
    amps = np.array([0.1109, 0.2218])
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian", "d0", 256)], [dict(amp=amps, angle=0.0, sigma=64.0)])
    _, y = O.rk4_solve(om, samples, np.array([1, 0, 0]), (0.0, 256 * O.MANILA_DT), O.MANILA_DT / 20)
    p = np.abs(y[:, -1]) ** 2
    assert p[0, 1] > 0.999 and p[1, 0] > 0.98  # pi and 2 pi


def test_norm_conservation_and_rk4_order(manila_q0):
    """RK4 converges as h^4 on a smooth right-hand side (constant envelope, one spare sample so that no
    stage ever sees the end of the window).  With a sampled envelope the k4/k1 stages that land on a sample
    boundary make the method first order in h -- which is why parity needs the SAME grid, not a finer one."""
    om, st, ops, ch = manila_q0
    samples = np.zeros((1, 2, 17), dtype=complex)
    samples[0, 0, :] = 0.8
    span = (0.0, 16 * O.MANILA_DT)
    _, truth = O.dop853_solve(om, samples, np.array([1, 0, 0]), span)
    errs = []
    for sub in (20, 40, 80):
        _, y = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, O.MANILA_DT / sub)
        errs.append(np.abs(y[0, -1] - truth[0, -1]).max())
        assert abs(np.linalg.norm(y[0, -1]) - 1) < 1e-5
    assert 12 < errs[0] / errs[1] < 20 and 12 < errs[1] / errs[2] < 20
    # sampled envelope: first-order boundary term dominates
    gs = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 16)], [dict(amp=0.8, angle=0.0, sigma=4.0, beta=1.0)])
    _, truth = O.dop853_solve(om, gs, np.array([1, 0, 0]), span)
    e = [np.abs(O.rk4_solve(om, gs, np.array([1, 0, 0]), span, O.MANILA_DT / sub)[1][0, -1] - truth[0, -1]).max() for sub in (40, 80)]
    assert e[0] / e[1] < 4


def test_cross_integrator_agreement(manila_q0):
    om, st, ops, ch = manila_q0
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian", "d0", 16)], [dict(amp=0.5, angle=0.0, sigma=8.0)])
    span = (0.0, 16 * O.MANILA_DT)
    _, a = O.dop853_solve(om, samples, np.array([1, 0, 0]), span)
    _, b, steps = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, rtol=1e-12, atol=1e-12, hmax=O.MANILA_DT / 4)
    assert np.abs(a[0, -1] - b[0, -1]).max() < 1e-8
    assert steps[0, 0] >= steps[0, 1] > 16


def test_dopri5_dense_output_matches_direct_solve(manila_q0):
    om, st, ops, ch = manila_q0
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian", "d0", 16)], [dict(amp=0.5, angle=0.0, sigma=8.0)])
    span = (0.0, 16 * O.MANILA_DT)
    te = np.linspace(span[0], span[1], 9)
    ts, y, _ = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, te, rtol=1e-11, atol=1e-11)
    _, truth = O.dop853_solve(om, samples, np.array([1, 0, 0]), span, te)
    assert ts.shape == (9,) and np.abs(y - truth).max() < 1e-7


def test_rwa_keeps_resonant_physics():
    m_full = _two_level_model(rwa=None)
    m_rwa = _two_level_model(rwa=7e9)
    assert np.count_nonzero(m_rwa.pos[0]) == 1 and np.count_nonzero(m_rwa.neg[0]) == 1
    samples = np.full((1, 1, 100), 0.3, dtype=complex)
    span = (0.0, 100 * m_full.dt)
    _, a = O.dop853_solve(m_full, samples, np.array([1, 0]), span)
    _, b = O.dop853_solve(m_rwa, samples, np.array([1, 0]), span)
    assert 1e-6 < np.abs(a[0, -1] - b[0, -1]).max() < 2e-2  # Bloch-Siegert-sized difference, not more


def test_lindblad_trace_hermiticity_and_decay():
    """Amplitude damping of |1>: P1(t) = exp(-gamma t) exactly; trace and hermiticity conserved."""
    h0 = np.diag([0.0, 2 * np.pi * 1e9]).astype(complex)
    x = np.array([[0, 1], [1, 0]], dtype=complex)
    gamma = 5e7
    L = np.sqrt(gamma) * np.array([[0, 1], [0, 0]], dtype=complex)
    m = O.OracleModel(h0, x[None] * 0.0, ["d0"], {"d0": 1e9}, 1e-10, rotating_frame=h0, dissipators=L[None])
    samples = np.zeros((1, 1, 100), dtype=complex)
    T = 100 * m.dt
    ts, rho = O.lindblad_rk4_solve(m, samples, np.array([0, 1]), (0.0, T), m.dt / 4)
    r = rho[0, -1]
    assert abs(np.trace(r) - 1) < 1e-12 and np.allclose(r, r.conj().T)
    assert r[1, 1].real == pytest.approx(np.exp(-gamma * T), abs=1e-9)


def test_lindblad_reduces_to_schrodinger_without_dissipators(manila_q0):
    om, st, ops, ch = manila_q0
    m = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=st, dissipators=np.zeros((1, 3, 3)))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 8)], [dict(amp=0.7, angle=0.0, sigma=3.0, beta=0.5)])
    span = (0.0, 8 * O.MANILA_DT)
    _, y = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, O.MANILA_DT / 200)
    _, rho = O.lindblad_rk4_solve(m, samples, np.array([1, 0, 0]), span, O.MANILA_DT / 200)
    np.testing.assert_allclose(rho[0, -1], np.outer(y[0, -1], y[0, -1].conj()), atol=1e-9)


def test_postprocess_and_populations(manila_q0):
    om, st, ops, ch = manila_q0
    y = np.array([[[0.6, 0.0, 0.8j], [0.0, 1.0, 0.0]]], dtype=complex)
    ts = np.array([0.0, 1e-9])
    out = O.postprocess_statevectors(om, ts, y)
    np.testing.assert_allclose(out[0, 0], y[0, 0])  # t = 0: frame and (diagonal H0) dressed basis are identities
    np.testing.assert_allclose(np.abs(out[0, 1]), [0, 1, 0], atol=1e-15)
    pops = O.populations_dict(np.abs(out[0, 0]) ** 2)
    assert pops == pytest.approx({"0": 0.36, "1": 0.64})  # leakage folds into "1" 
    two = O.populations_dict(np.full(9, 1 / 9), subsystem_dims=[3, 3], measured=(0, 1))
    assert two == pytest.approx({"00": 1 / 9, "01": 2 / 9, "10": 2 / 9, "11": 4 / 9})
    smp, counts = O.sample_counts(pops, 1000, seed=5)
    assert sum(counts.values()) == 1000 and 550 < counts["1"] < 730


def test_dressed_decomposition_two_qubits():
    st, _, _, _ = O.parse_hamiltonian_dict(O.MANILA, [0, 1])
    ev, V = O.dressed_decomposition(st)
    assert np.allclose(V.conj().T @ V, np.eye(9), atol=1e-12)
    assert np.allclose(V.conj().T @ st @ V, np.diag(ev), atol=1e-3)  # rad/s on a 6e10 scale
    assert all(np.argmax(np.abs(V[:, k])) == k for k in range(9)) and np.all(np.diag(V).real > 0.99)


def test_lindblad_rhs_restructured_form_equals_textbook_form():
    """lindblad_rhs_dense uses A rho + rho A^+ + sum L rho L^+ with one cached D = sum L^+L and an elementwise product for
    diagonal L: compare with the textbook -i[H,rho] + sum (L rho L^+ - 1/2 {L^+L, rho}) with explicit Bohr phases."""
    rng = np.random.default_rng(4)
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0, 1])
    d = 9
    a = np.diag(np.sqrt(np.arange(1, 3)), k=1).astype(complex)
    Ls = [np.kron(np.eye(3), 3e3 * a), np.kron(2e3 * a.conj().T @ a, np.eye(3)),
          1e3 * (rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))]
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=np.diag(st).real, dissipators=np.array(Ls))
    x = rng.normal(size=(2, d, d)) + 1j * rng.normal(size=(2, d, d))
    rho = x @ np.conj(np.swapaxes(x, -1, -2))
    t = 3.3e-9
    z = rng.normal(size=(2, len(ch))) + 1j * rng.normal(size=(2, len(ch)))
    got = O.lindblad_rhs_dense(om, t, z, rho)
    H = om.hamiltonian_fb(t, z)
    ph = np.exp(1j * om.bohr * t)
    want = -1j * (H @ rho - rho @ H)
    for L in om.diss_fb:
        Lt = ph * L
        LdL = Lt.conj().T @ Lt
        want = want + Lt @ rho @ Lt.conj().T - 0.5 * (LdL @ rho + rho @ LdL)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_dopri5_tableau_is_anchored_to_scipy_rk45():
    """The Dormand-Prince constants the oracle (and, through the GPU parity tests, the CUDA steppers) use, checked against an
    INDEPENDENT implementation that ships in this image: scipy.integrate.RK45.  Nodes, stage matrix and 5th-order weights are
    the same numbers; the dense-output mid-point weights equal scipy's interpolation polynomial at theta = 1/2; the error
    weights are jax.experimental.ode's variant, which is exactly -2/3 of scipy's E (same embedded pair, scaled estimate)."""
    from scipy.integrate import RK45

    from oracle import solvers as S

    assert np.array_equal(S._DP_ALPHA[:5], RK45.C[1:6])
    assert np.abs(S._DP_BETA[:5, :5] - RK45.A[1:6, :5]).max() == 0.0
    assert np.abs(S._DP_BETA[5, :6] - RK45.B).max() == 0.0  # FSAL: the last stage row is the 5th-order solution
    assert np.abs(S._DP_CSOL[:6] - RK45.B).max() == 0.0 and S._DP_CSOL[6] == 0.0
    half = RK45.P @ np.array([0.5, 0.25, 0.125, 0.0625])
    assert np.abs(S._DP_CMID - half).max() < 5e-16  # scipy evaluates a cubic in theta: rounding of the sum
    assert np.abs(S._DP_CERR - (-2.0 / 3.0) * RK45.E).max() < 1e-16
