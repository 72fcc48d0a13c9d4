"""GPU parity for the sweep axes north_star names beyond amplitude / beta: per-point DETUNING (a frequency shift of the
Play) and per-point DURATION (PulseGate.duration, src/casq/gates/pulse_gate.py:69-72, which fixes
t_span, src/casq/backends/qiskit/dynamics_backend_patch.py:180-184), plus the cross-resonance gate on a control channel."""
import numpy as np
import pytest
import torch

import oracle as O
from casq_b200 import engine
from casq_b200._native import NativeModel
from casq_b200.backends import BackendLibrary, PulseBackend, build
from casq_b200.gates import CrossResonancePulseGate, DragPulseGate, GaussianSquarePulseGate, PulseCircuit
from casq_b200.models import ControlModel, TransmonModel
from casq_b200.pulse import DriveChannel, Play, Schedule, ShiftFrequency, ShiftPhase

pytestmark = pytest.mark.gpu
DT = O.MANILA_DT
RK4 = PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_RK4


def _q0(frame="full", rwa=None):
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    rf = {"full": st, "none": None, "shifted": np.diag(st).real + 2e8}[frame]
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    nm = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=rwa)
    return om, nm, ch


def test_envelope_sampler_with_frequency_shift():
    rng = np.random.default_rng(2)
    B, dur, start = 65, 40, 7
    p = dict(amp=rng.uniform(0.1, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=rng.uniform(4, 12, B), beta=rng.uniform(-2, 2, B),
             freq_shift=rng.uniform(-5e7, 5e7, B))
    want = O.channel_samples(["d0"], [O.PulseSpec("drag", "d0", dur, start=start)], [p], dt=DT)[:, 0, start:]
    got = engine.envelope_sample("drag", dur, engine.pack_params([p], device="cuda")[0], start=start, dt=DT).cpu().numpy()
    assert np.abs(got - want).max() < 1e-13


@pytest.mark.parametrize("frame,rwa", [("full", None), ("none", None), ("shifted", 3.0e9)])
def test_detuning_sweep_rk4_matches_oracle(frame, rwa):
    rng = np.random.default_rng(17)
    om, nm, ch = _q0(frame, rwa)
    B, dur = 33, 40
    p = dict(amp=rng.uniform(0.05, 0.9, B), angle=rng.uniform(-np.pi, np.pi, B), sigma=rng.uniform(4, 20, B),
             beta=rng.uniform(-3, 3, B), freq_shift=np.linspace(-4e7, 4e7, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p], dt=DT)
    span = (0.0, dur * DT)
    te = np.linspace(0, dur * DT, 5)
    sub = 40 if frame != "none" else 160
    _, want = O.rk4_solve(om, samples, np.array([1, 0, 0]), span, DT / sub, te)
    got = engine.solve_sv_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p], device="cuda"), np.array([1, 0, 0]),
                              span, DT / sub, te).cpu().numpy()
    assert np.abs(got - want).max() < 1e-9
    # the shift matters: the same sweep without it differs at the 1e-2 level
    p0 = dict(p, freq_shift=0.0)
    base = engine.solve_sv_rk4(nm, [("drag", ch.index("d0"), 0, dur)], engine.pack_params([p0], device="cuda"), np.array([1, 0, 0]),
                               span, DT / sub, te).cpu().numpy()
    assert np.abs(got - base).max() > 1e-3


def test_detuning_headline_kernel_and_second_slot():
    """B >= 16384 takes the two-trajectories-per-thread kernel; a second Play starting at sample 30 checks the GLOBAL sample
    index in the phase (start + n)."""
    rng = np.random.default_rng(5)
    om, nm, ch = _q0()
    B = 16384 + 3
    ps = [dict(amp=rng.uniform(0.05, 0.9, B), angle=0.0, sigma=rng.uniform(4, 10, B), beta=rng.uniform(-3, 3, B),
               freq_shift=rng.uniform(-4e7, 4e7, B)),
          dict(amp=rng.uniform(0.05, 0.9, B), angle=0.3, sigma=5.0, freq_shift=rng.uniform(-4e7, 4e7, B))]
    specs = [O.PulseSpec("drag", "d0", 24), O.PulseSpec("gaussian", "d0", 20, start=30)]
    slots = [("drag", 0, 0, 24), ("gaussian", 0, 30, 20)]
    span = (0.0, 50 * DT)
    got = engine.solve_sv_rk4(nm, slots, engine.pack_params(ps, device="cuda"), np.array([1, 0, 0]), span, DT / 40).cpu().numpy()
    idx = np.array([0, 1, 777, B - 2, B - 1])
    sub = [{k: (v[idx] if np.ndim(v) else v) for k, v in p.items()} for p in ps]
    _, want = O.rk4_solve(om, O.channel_samples(ch, specs, sub, dt=DT), np.array([1, 0, 0]), span, DT / 40)
    assert np.abs(got[idx] - want).max() < 1e-9


def test_detuning_dopri5_propagators_and_lindblad_paths():
    """The adaptive solver and the propagators read the envelope TABLE, the Lindblad path evaluates it in its coefficient
    kernel: all three carry the frequency shift."""
    rng = np.random.default_rng(23)
    om, nm, ch = _q0("full", 3.0e9)
    B, dur = 6, 24
    p = dict(amp=rng.uniform(0.2, 0.9, B), angle=rng.uniform(-1, 1, B), sigma=rng.uniform(4, 9, B), beta=rng.uniform(-2, 2, B),
             freq_shift=rng.uniform(-4e7, 4e7, B))
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p], dt=DT)
    span = (0.0, dur * DT)
    params = engine.pack_params([p], device="cuda")
    slots = [("drag", 0, 0, dur)]
    _, want, _ = O.dopri5_jax_solve(om, samples, np.array([1, 0, 0]), span, rtol=1e-10, atol=1e-10)
    got = engine.solve_sv_dopri5(nm, slots, params, np.array([1, 0, 0]), span, rtol=1e-10, atol=1e-10).cpu().numpy()
    assert np.abs(got - want).max() < 2e-6  # adaptive controller on a discontinuous RHS
    U = engine.propagators_expm(nm, slots, params, span, DT).cpu().numpy()
    assert np.abs(U - O.expm_midpoint_propagators(om, samples, span, DT)).max() < 1e-9
    # Lindblad, diagonal frame
    st, ops, _, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    a = np.diag(np.sqrt(np.arange(1, 3)), k=1).astype(complex)
    Ls = np.array([np.sqrt(1e7) * a, np.sqrt(2e6) * a.conj().T @ a])
    rf = np.diag(st).real
    oml = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=rf, rwa_cutoff_freq=3.0e9, dissipators=Ls)
    nml = NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], DT, rotating_frame=rf, rwa_cutoff_freq=3.0e9)
    nml.set_dissipators(Ls)
    _, wantl = O.lindblad_rk4_solve(oml, samples, np.array([1, 0, 0]), span, DT / 8)
    rho, _ = engine.solve_lindblad_rk4(nml, slots, params, np.array([1, 0, 0]), span, DT / 8)
    assert np.abs(rho.cpu().numpy() - wantl).max() < 1e-9


def _backend(qubits=(0,)):
    h = TransmonModel.manila(extracted_qubits=list(qubits))
    c = ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS)
    return build(BackendLibrary.B200, h, c, seed=11)


def test_solve_batch_detuning_reduces_to_solve_with_shift_frequency():
    """B = 1 semantics: a per-point detuning is a ShiftFrequency in front of the Play of the single circuit (and a phase shift
    joins the pulse angle)."""
    backend = _backend()
    gate = DragPulseGate(duration=32, amplitude=0.4)
    params = {"sigma": 8.0, "beta": 1.5}
    opts = {"max_dt": DT / 40}
    batch = backend.solve_batch(gate, {"sigma": np.array([8.0, 8.0]), "beta": np.array([1.5, 1.5])}, RK4, run_options=dict(opts),
                                detuning=np.array([2.5e7, -1e7]), angle=np.array([0.4, 0.0]))
    circuit = PulseCircuit(1, 1)
    circuit.pulse_gate(gate, params, 0)
    sched = Schedule(name="detuned")  # shift instructions first, then the gate's Play (same start sample: insertion order)
    sched.insert(0, ShiftFrequency(2.5e7, DriveChannel(0)))
    sched.insert(0, ShiftPhase(0.4, DriveChannel(0)))
    sched.insert(0, Play(gate.pulse(params), DriveChannel(0)))
    circuit.add_calibration(gate.name, [0], sched, [])
    circuit.measure(0, 0)
    sol = backend.solve(circuit, RK4, run_options=dict(opts))
    assert np.abs(sol.states[-1].data - batch.states[0, -1].cpu().numpy()).max() < 1e-13
    assert sum(sol.counts[-1].values()) == 1024


def test_duration_sweep_matches_per_duration_solves_and_oracle():
    backend = _backend()
    om, _, ch = _q0()
    rng = np.random.default_rng(3)
    durations = np.array([16, 32, 24, 16, 48, 32, 24])
    B = len(durations)
    sigma, beta, amp = rng.uniform(3, 7, B), rng.uniform(-2, 2, B), rng.uniform(0.2, 0.8, B)
    det = rng.uniform(-2e7, 2e7, B)
    gate = DragPulseGate(duration=64, amplitude=1.0)
    sweep = backend.solve_batch(gate, {"sigma": sigma, "beta": beta}, RK4, amplitude=amp, run_options={"max_dt": DT / 40},
                                duration=durations, detuning=det)
    assert sweep.states.shape == (B, 2, 3) and sweep.times.shape == (B, 2)
    assert np.array_equal(sweep.durations, durations) and np.allclose(sweep.times[:, 1], durations * DT)
    for b in range(B):
        p = dict(amp=amp[b:b + 1], angle=0.0, sigma=sigma[b:b + 1], beta=beta[b:b + 1], freq_shift=det[b:b + 1])
        d = int(durations[b])
        samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", d)], [p], dt=DT)
        ts, y = O.rk4_solve(om, samples, om.dressed()[1][:, 0], (0.0, d * DT), DT / 40)
        want = O.postprocess_statevectors(om, ts, y)[0, -1]
        assert np.abs(sweep.states[b, -1].cpu().numpy() - want).max() < 1e-9, b
    # scalar duration = a gate of that duration
    one = backend.solve_batch(gate, {"sigma": sigma[:1], "beta": beta[:1]}, RK4, amplitude=amp[:1], run_options={"max_dt": DT / 40},
                              duration=16, detuning=det[:1])
    assert torch.equal(one.states[0], sweep.states[0])
    with pytest.raises(Exception):
        backend.solve_batch(gate, {"sigma": sigma, "beta": beta}, RK4, run_options={"max_dt": DT / 40}, duration=durations + 0.5)


def test_cross_resonance_gate_on_control_channel_two_transmons():
    """BASELINE cfg 3's drive through the gate API: GaussianSquare on u0 (control qubit 0 driven at qubit 1's frequency)."""
    backend = _backend((0, 1))
    assert backend.dim == 9 and "u0" in backend.hamiltonian.channels
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0, 1])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
    gate = CrossResonancePulseGate(control_channel=0, duration=48, amplitude=1.0)
    B = 5
    amp, width = np.linspace(0.1, 0.8, B), np.linspace(0, 30, B)
    sol = backend.solve_batch(gate, {"sigma": np.full(B, 4.0), "width": width}, RK4, amplitude=amp, run_options={"max_dt": DT / 40})
    p = dict(amp=amp, angle=0.0, sigma=4.0, width=width)
    samples = O.channel_samples(ch, [O.PulseSpec("gaussian_square", "u0", 48)], [p])
    ts, y = O.rk4_solve(om, samples, om.dressed()[1][:, 0], (0.0, 48 * DT), DT / 40)
    want = O.postprocess_statevectors(om, ts, y)
    assert np.abs(sol.states.cpu().numpy() - want).max() < 1e-9
    # drop-in call: a two-qubit circuit with the CR gate on (control, target) and a measurement
    circuit = PulseCircuit(2, 1)
    circuit.pulse_gate(gate, {"sigma": 4.0, "width": float(width[2])}, (0, 1))
    gate.amplitude = float(amp[2])
    circuit.calibrations.clear()
    circuit.data.clear()
    circuit.pulse_gate(gate, {"sigma": 4.0, "width": float(width[2])}, (0, 1))
    circuit.measure(0, 0)
    one = backend.solve(circuit, RK4, run_options={"max_dt": DT / 40})
    assert np.abs(one.states[-1].data - want[2, -1]).max() < 1e-9
    with pytest.raises(Exception):
        circuit.pulse_gate(gate, {"sigma": 4.0, "width": 1.0}, 0)  # a two-qubit gate needs (control, target)


def test_solve_with_density_matrix_initial_state_matches_oracle():
    """The DensityMatrix branch of casq's result function (src/casq/backends/qiskit/helpers.py:110-121): solve() routes a
    DensityMatrix initial state to the Lindblad kernels (with and without a NoiseModel) and returns DensityMatrix states."""
    from casq_b200.models import NoiseModel
    from casq_b200.quantum_info import DensityMatrix

    rng = np.random.default_rng(12)
    x = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    rho0 = x @ x.conj().T
    rho0 /= np.trace(rho0).real
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    gate = DragPulseGate(duration=24, amplitude=0.6)
    circuit = PulseCircuit.from_pulse_gate(gate, {"sigma": 5.0, "beta": 1.2})
    for noise in (None, NoiseModel({0: NoiseModel.QubitNoise(t1=40e-9, t2=30e-9)})):
        h = TransmonModel.manila(extracted_qubits=[0])
        c = ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS)
        backend = build(BackendLibrary.B200, h, c, seed=5)
        if noise is not None:
            backend = type(backend)(h, c, seed=5, noise=noise)
        Ls = np.zeros((1, 3, 3), dtype=complex) if noise is None else noise.lindblad_operators({0: 3})
        om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st, dissipators=Ls)
        sol = backend.solve(circuit, RK4, initial_state=DensityMatrix(rho0), steps=4, run_options={"max_dt": DT / 40})
        p = dict(amp=0.6, angle=0.0, sigma=5.0, beta=1.2)
        samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 24)], [p])
        te = np.linspace(0, 24 * DT, 4)
        _, rho = O.lindblad_rk4_solve(om, samples, rho0, (0.0, 24 * DT), DT / 40, te)
        want = O.postprocess_density_matrices(om, te, rho, normalize=True)[0]
        assert len(sol.states) == 4 and isinstance(sol.states[-1], DensityMatrix)
        got = np.array([s.data for s in sol.states])
        assert np.abs(got - want).max() < 1e-9
        pr = np.real(np.diag(want[-1]))
        assert sol.populations[-1]["0"] == pytest.approx(pr[0], abs=1e-9) and sol.populations[-1]["1"] == pytest.approx(pr[1:].sum(), abs=1e-9)
        assert sum(sol.counts[-1].values()) == 1024
    with pytest.raises(Exception, match="fixed-step"):
        backend.solve(circuit, PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT, initial_state=DensityMatrix(rho0))


def test_forwarded_subsystem_dims_give_per_qubit_populations():
    """casq does not forward subsystem_dims (qiskit_pulse_backend.py:184-188): a two-transmon solve folds everything into one
    pseudo-qubit.  forward_subsystem_dims=True reports the measured qubits; the default keeps casq's behaviour."""
    h = TransmonModel.manila(extracted_qubits=[0, 1])
    c = ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS)
    plain = build(BackendLibrary.B200, h, c, seed=9)
    fwd = type(plain)(h, c, seed=9, forward_subsystem_dims=True)
    gate = DragPulseGate(duration=32, amplitude=0.35)
    circuit = PulseCircuit(2, 2)
    circuit.pulse_gate(gate, {"sigma": 8.0, "beta": 0.5}, 1)
    circuit.measure(0, 0)
    circuit.measure(1, 1)
    a = plain.solve(circuit, RK4, run_options={"max_dt": DT / 40})
    b = fwd.solve(circuit, RK4, run_options={"max_dt": DT / 40})
    assert np.abs(a.states[-1].data - b.states[-1].data).max() == 0.0
    probs = np.abs(a.states[-1].data) ** 2
    assert a.qubits == [0] and set(a.populations[-1]) <= {"0", "1"}
    assert a.populations[-1]["0"] == pytest.approx(probs[0]) and a.populations[-1]["1"] == pytest.approx(probs[1:].sum())
    assert b.qubits == [0, 1] and b.states[-1].dims() == (3, 3)
    want = O.populations_dict(probs, subsystem_dims=[3, 3], measured=(0, 1))
    assert {k: v for k, v in b.populations[-1].items()} == pytest.approx({k: v for k, v in want.items() if v > 0})
    assert b.populations[-1]["10"] > 0.2 > 10 * b.populations[-1].get("01", 0.0)  # the drive is on qubit 1 (left bit)
    assert sum(b.counts[-1].values()) == 1024 and np.array(b.iq_data[-1]).shape == (1024, 2, 2)


def test_gate_fidelities_from_propagators_match_oracle():
    """X-gate fidelity of a DRAG amplitude sweep from the batched propagators: the device numbers equal the same formula
    applied to the oracle's exponential-midpoint propagators, and the sweep has a high-fidelity point near the pi pulse."""
    from casq_b200.fidelity import gate_fidelities

    backend = _backend()
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
    B, dur = 24, 64
    amp = np.linspace(0.3, 0.6, B)
    p = dict(amp=amp, angle=0.0, sigma=16.0, beta=0.6)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    slots = [("drag", ch.index("d0"), 0, dur)]
    fa, fp, leak = backend.gate_fidelities_batch(slots, [p], dur, DT / 40, X, virtual_z=True)
    Uo = O.expm_midpoint_propagators(om, O.channel_samples(ch, [O.PulseSpec("drag", "d0", dur)], [p]), (0.0, dur * DT), DT / 40)
    fa_o, fp_o, leak_o = gate_fidelities(torch.as_tensor(Uo), X, [3], virtual_z=True)
    assert float((fa.cpu() - fa_o).abs().max()) < 1e-9 and float((leak.cpu() - leak_o).abs().max()) < 1e-9
    assert float(fa.max()) > 0.99 and float(fa.min()) < 0.9 and float(leak.max()) < 0.05
