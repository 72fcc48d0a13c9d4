"""Host-side mirror of casq's public surface (gates, circuit, models, factory): written to read like the
reference's own unit tests (tests/casq/gates/*_test.py echo tests, backend factory tests) plus parser parity
with the oracle's independent parser."""
import numpy as np
import pytest

import oracle as O
from casq_b200.backends import BackendLibrary, PulseBackend, build, build_from_backend
from casq_b200.common.exceptions import CasqError
from casq_b200.gates import (ConstantPulseGate, DragPulseGate, GaussianPulseGate, GaussianSquareDragPulseGate,
                             GaussianSquarePulseGate, PulseCircuit, PulseGate)
from casq_b200.models import ControlModel, HamiltonianModel, TransmonModel
from casq_b200.models.hamiltonian_parser import parse_hamiltonian_dict


def test_gaussian_pulse_gate():
    # mirrors tests/casq/gates/gaussian_pulse_gate_test.py: parameters echo back through the pulse
    gate = GaussianPulseGate(duration=16, amplitude=1, name="test")
    pulse = gate.pulse({"sigma": 8})
    assert pulse.pulse_type == "Gaussian" and pulse.duration == 16 and pulse.amp == 1 and pulse.params["sigma"] == 8
    assert gate.to_parameters_dict(np.array([3.0])) == {"sigma": 3.0}
    sched = gate.schedule({"sigma": 8}, 2)
    (start, play), = sched.instructions
    assert start == 0 and play.channel.name == "d2" and sched.duration == 16


def test_drag_pulse_gate():
    # mirrors tests/casq/gates/drag_pulse_gate_test.py:45-53
    gate = DragPulseGate(duration=16, amplitude=1)
    assert gate.angle is None  # drag_pulse_gate.py:54 default
    pulse = gate.pulse({"sigma": 8, "beta": 2})
    assert pulse.pulse_type == "Drag" and pulse.params == {"sigma": 8, "beta": 2} and pulse.angle == 0.0
    assert gate.to_parameters_dict([1, 2]) == {"sigma": 1, "beta": 2}


def test_other_gates_and_validation():
    assert ConstantPulseGate(8, 0.5).pulse().pulse_type == "Constant"
    assert ConstantPulseGate(8, 0.5).to_parameters_dict(np.array([])) == {}
    gs = GaussianSquarePulseGate(32, 0.5).pulse({"sigma": 4, "width": 16})
    assert gs.pulse_type == "GaussianSquare" and gs.params["width"] == 16
    gsd = GaussianSquareDragPulseGate(32, 0.5)
    assert gsd.to_parameters_dict([1, 2, 3]) == {"sigma": 1, "width": 2, "beta": 3}
    with pytest.raises(CasqError, match="amplitude"):
        GaussianPulseGate(16, 1.5).pulse({"sigma": 8})
    assert GaussianPulseGate(16, 1.5, limit_amplitude=False).pulse({"sigma": 8}).amp == 1.5
    with pytest.raises(CasqError, match="sigma"):
        GaussianPulseGate(16, 1.0).pulse({"sigma": 0})
    with pytest.raises(CasqError, match="width"):
        GaussianSquarePulseGate(16, 1.0).pulse({"sigma": 2, "width": 17})
    with pytest.raises(CasqError, match="needs parameter"):
        DragPulseGate(16, 1.0).pulse({"sigma": 2})
    with pytest.raises(TypeError):
        PulseGate(1, 16, 1.0)  # abstract


def test_pulse_circuit_from_pulse_gate():
    gate = DragPulseGate(duration=256, amplitude=1)
    circuit = PulseCircuit.from_pulse_gate(gate, {"sigma": 128, "beta": 2})
    assert circuit.num_qubits == 1 and circuit.num_clbits == 1
    assert [i.operation for i in circuit.data] == [gate, "measure"]
    assert gate.name in circuit.calibrations and circuit.measured_qubits() == [0]
    with pytest.raises(CasqError):
        circuit.pulse_gate(gate, {"sigma": 1, "beta": 0}, qubit=3)


@pytest.mark.parametrize("sub", [[0], [0, 1], [1, 3], [2, 3, 4], None])
def test_parser_agrees_with_oracle_parser(sub):
    a = parse_hamiltonian_dict(TransmonModel.MANILA, sub)
    b = O.parse_hamiltonian_dict(O.MANILA, sub)
    assert a[2] == b[2] and a[3] == b[3]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_parser_errors():
    with pytest.raises(CasqError, match="h_str"):
        parse_hamiltonian_dict({"qub": {"0": 2}, "vars": {}})
    with pytest.raises(CasqError, match="unknown symbol"):
        parse_hamiltonian_dict({"h_str": ["foo*X0"], "qub": {"0": 2}, "vars": {}})
    with pytest.raises(CasqError, match="not in the Hamiltonian"):
        parse_hamiltonian_dict({"h_str": ["X0"], "qub": {"0": 2}, "vars": {}}, [5])
    st, ops, ch, dims = parse_hamiltonian_dict({"h_str": ["2*pi*(X0+Y0)/2||D0", "-w*Z0/2"], "qub": {"0": 2}, "vars": {"w": 3.0}})
    assert ch == ["d0"] and np.allclose(st, np.diag([-1.5, 1.5])) and np.allclose(ops[0], np.pi * np.array([[0, 1 - 1j], [1 + 1j, 0]]))


def test_transmon_model_and_frame_default_quirk():
    P = TransmonModel.TransmonProperties
    qm = {0: P(31179079102.853794, -2165345334.8252344, 926545606.6640488),
          1: P(30397743782.610542, -2169482392.6367006, 892870223.8110852)}
    full = TransmonModel(qm, {(0, 1): 11845444.218797993})
    assert full.static_operator.shape == (9, 9) and full.channels == ["d0", "d1", "u0", "u1"]
    assert full.extracted_qubits == [0]  # hamiltonian_model.py:78 stores [0] but parses all qubits
    assert full.rotating_frame.shape == (9, 9)  # DENSE default -> full static operator (quirk 1)
    one = TransmonModel(qm, {(0, 1): 11845444.218797993}, extracted_qubits=[0])
    assert one.static_operator.shape == (3, 3) and one.channels == ["d0", "u0"]
    diag = HamiltonianModel(TransmonModel.MANILA, extracted_qubits=[0, 1], evaluation_mode=None)
    assert diag.rotating_frame.shape == (9,)  # raw None -> 1-D diagonal frame, mode still DENSE
    assert diag.evaluation_mode is HamiltonianModel.EvaluationMode.DENSE


def test_factory_and_no_cpu_fallback(native_lib):
    import torch

    h = TransmonModel.manila(extracted_qubits=[0])
    c = ControlModel(dt=O.MANILA_DT, channel_carrier_freqs=O.MANILA_CARRIERS)
    with pytest.raises(ValueError, match="Unknown backend library: QISKIT"):
        build(BackendLibrary.QISKIT, h, c)
    with pytest.raises(ValueError, match="Unknown backend"):
        build_from_backend(object())
    backend = build(BackendLibrary.B200, h, c, seed=1234)
    assert isinstance(backend, PulseBackend) and backend._seed == 1234 and backend.dim == 3
    np.testing.assert_allclose(backend.ground_state(), [1, 0, 0])
    with pytest.raises(CasqError, match="channel_carrier_freqs"):
        build(BackendLibrary.B200, h, ControlModel(dt=O.MANILA_DT, channel_carrier_freqs={"d0": 5e9}))
    circuit = PulseCircuit.from_pulse_gate(GaussianPulseGate(16, 1), {"sigma": 8})
    sched, t_acq, measured = backend._circuit_to_schedule(circuit)
    assert t_acq == 16 and measured == [0]
    assert backend._slots_from_schedule(sched)[0] == [("gaussian", 0, 0, 16)]
    if not torch.cuda.is_available():
        with pytest.raises(CasqError, match="no CPU fallback"):
            backend.solve(circuit, PulseBackend.ODESolverMethod.SCIPY_DOP853)
    no_measure = PulseCircuit(1, 1)
    no_measure.pulse_gate(GaussianPulseGate(16, 1), {"sigma": 8})
    with pytest.raises(CasqError, match="measurement"):
        backend._circuit_to_schedule(no_measure)


def test_solver_method_enum_matches_casq():
    M = PulseBackend.ODESolverMethod
    assert {m.name: m.value for m in M} == {
        "QISKIT_DYNAMICS_RK4": "RK4", "QISKIT_DYNAMICS_JAX_RK4": "jax_RK4", "QISKIT_DYNAMICS_JAX_ODEINT": "jax_odeint",
        "SCIPY_BDF": "BDF", "SCIPY_DOP853": "DOP853", "SCIPY_LSODA": "LSODA", "SCIPY_RADAU": "Radau",
        "SCIPY_RK23": "RK23", "SCIPY_RK45": "RK45"}
    assert [m.name for m in BackendLibrary][:3] == ["C3", "QISKIT", "QUTIP"]


def test_pack_params_layout():
    from casq_b200 import engine

    t = engine.pack_params([dict(amp=np.array([0.1, 0.2, 0.3]), angle=None, sigma=4.0, beta=np.array([1.0, 2.0, 3.0]))])
    assert t.shape == (1, 6, 3) and t[0, 5].tolist() == [0, 0, 0]
    assert t[0, 0].tolist() == [0.1, 0.2, 0.3] and t[0, 1].tolist() == [0, 0, 0] and t[0, 2].tolist() == [4, 4, 4]
    assert t[0, 3].tolist() == [0, 0, 0] and t[0, 4].tolist() == [1, 2, 3]
    with pytest.raises(CasqError, match="inconsistent"):
        engine.pack_params([dict(amp=np.zeros(3), sigma=np.zeros(4))])
    with pytest.raises(CasqError, match="unknown"):
        engine.pack_params([dict(amp=1.0, gamma=2.0)])


def test_decorators_and_per_qubit_populations(caplog):
    import logging

    from casq_b200.common import timer, trace
    from casq_b200.quantum_info import populations_by_qubit

    @trace(level="DEBUG")
    @timer(level="DEBUG", unit="msec")
    def f(x, y=2):
        return x + y

    with caplog.at_level(logging.DEBUG, logger="casq_b200"):
        assert f(1, y=3) == 4
    text = caplog.text
    assert "Entering [f" in text and "Exiting [f] with result [4]" in text and "milliseconds" in text
    two = populations_by_qubit(np.full(9, 1 / 9), [3, 3])
    assert two == pytest.approx({"00": 1 / 9, "01": 2 / 9, "10": 2 / 9, "11": 4 / 9})
    assert two == pytest.approx(O.populations_dict(np.full(9, 1 / 9), subsystem_dims=[3, 3], measured=(0, 1)))
    one = populations_by_qubit(np.array([0.5, 0.25, 0.25, 0, 0, 0, 0, 0, 0]), [3, 3], measured=[0])
    assert one == pytest.approx({"0": 0.5, "1": 0.5})


def test_measurement_script_problems_equal_the_oracle_definitions():
    """scripts/ builds its benchmark problems with the PRODUCT's models (it must not import the oracle); they have to be the
    same problems the parity tests define through the oracle: bit-equal operators, channels, carriers and dt."""
    import oracle as O
    from scripts import _problems as P
    assert P.DT == O.MANILA_DT and P.CARRIERS == O.MANILA_CARRIERS
    for eq in ([0], [0, 1], None):
        a, b = P.manila(eq), O.parse_hamiltonian_dict(O.MANILA, eq)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == list(b[2])
    v = O.MANILA["vars"]
    for nq in (2, 3):
        qm = {q: (v[f"wq{q}"], v[f"delta{q}"], v[f"omegad{q}"]) for q in range(nq)}
        cm = {(q, q + 1): v[f"jq{q}q{q + 1}"] for q in range(nq - 1)}
        b = O.parse_hamiltonian_dict(O.transmon_hamiltonian_dict(qm, cm), None)
        a = P.chain(nq)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == list(b[2])
    import pathlib
    for f in pathlib.Path(P.__file__).parent.glob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f.name


def test_gate_fidelity_definitions():
    """casq_b200.fidelity on CPU tensors: closed-form cases (identity, global phase, leakage, virtual Z, two qubits)."""
    import torch
    from casq_b200.fidelity import computational_indices, gate_fidelities

    assert computational_indices([3, 3]).tolist() == [0, 1, 3, 4] and computational_indices([3]).tolist() == [0, 1]
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    U = np.zeros((4, 3, 3), dtype=complex)
    U[0, :2, :2] = X
    U[0, 2, 2] = 1                                  # perfect X
    U[1, :2, :2] = np.exp(0.7j) * X
    U[1, 2, 2] = 1                                  # global phase: same fidelity
    c = np.cos(0.3)
    U[2] = np.array([[0, 1, 0], [c, 0, -np.sin(0.3)], [np.sin(0.3), 0, c]])  # |0> leaks into |2>
    U[3, :2, :2] = X @ np.diag([1, np.exp(1.1j)])
    U[3, 2, 2] = 1                                  # X up to a Z rotation
    fa, fp, leak = gate_fidelities(torch.as_tensor(U), X, [3])
    assert fa[:2].tolist() == pytest.approx([1, 1]) and fp[:2].tolist() == pytest.approx([1, 1]) and leak[:2].tolist() == pytest.approx([0, 0])
    # M = diag(c, 1): F_avg = (c^2 + 1 + (1 + c)^2) / 6, leakage = (1 - c^2) / 2
    assert float(fa[2]) == pytest.approx((c * c + 1 + (1 + c) ** 2) / 6) and float(leak[2]) == pytest.approx((1 - c * c) / 2)
    assert float(fp[3]) == pytest.approx(abs(1 + np.exp(1.1j)) ** 2 / 4)
    fz, _, _ = gate_fidelities(torch.as_tensor(U), X, [3], virtual_z=True)
    assert float(fz[3]) == pytest.approx(1.0) and float(fz[0]) == pytest.approx(1.0)
    # two qutrits: CNOT embedded in 9 x 9 is found on the 4 computational states
    cnot = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)  # control = subsystem 0
    U9 = np.eye(9, dtype=complex)
    idx = computational_indices([3, 3])
    U9[np.ix_(idx, idx)] = cnot
    fa9, fp9, l9 = gate_fidelities(torch.as_tensor(U9[None]), cnot, [3, 3])
    assert float(fa9[0]) == pytest.approx(1.0) and float(l9[0]) == pytest.approx(0.0)
    with pytest.raises(CasqError):
        gate_fidelities(torch.as_tensor(U9[None]), X, [3, 3])


def _cpu_backend(qubits=(0,), **kw):
    from casq_b200.backends import BackendLibrary, build
    from casq_b200.models import ControlModel, TransmonModel

    h = TransmonModel.manila(extracted_qubits=list(qubits))
    c = ControlModel(dt=O.MANILA_DT, channel_carrier_freqs=O.MANILA_CARRIERS)
    b = build(BackendLibrary.B200, h, c, seed=3)  # model construction is host-only work: no GPU needed
    return type(b)(h, c, seed=3, **kw) if kw else b


def test_schedule_phase_and_frequency_instructions_become_slot_parameters():
    """ShiftPhase / SetPhase / ShiftFrequency / SetFrequency act on the LATER plays of their channel (qiskit-dynamics
    InstructionToSignals): phases join the pulse angle, frequency shifts become the freq_shift row."""
    from casq_b200.pulse import (ControlChannel, Drag, DriveChannel, Gaussian, Play, Schedule, SetFrequency, SetPhase,
                                 ShiftFrequency, ShiftPhase)

    backend = _cpu_backend()
    s = Schedule()
    s.insert(0, Play(Gaussian(duration=8, amp=0.2, sigma=2.0), DriveChannel(0)))       # before any shift: untouched
    s.insert(8, ShiftPhase(0.25, DriveChannel(0)))
    s.insert(8, ShiftFrequency(3e6, DriveChannel(0)))
    s.insert(8, ShiftFrequency(1e6, ControlChannel(0)))                                 # other channel: no effect on d0
    s.insert(8, Play(Drag(duration=8, amp=0.3, angle=0.5, sigma=2.0, beta=1.0), DriveChannel(0)))
    s.insert(16, ShiftPhase(0.25, DriveChannel(0)))
    s.insert(16, SetFrequency(O.MANILA_CARRIERS["d0"] - 2e6, DriveChannel(0)))
    s.insert(16, Play(Gaussian(duration=8, amp=0.1, sigma=2.0), DriveChannel(0)))
    s.insert(24, SetPhase(-1.0, DriveChannel(0)))
    s.insert(24, Play(Gaussian(duration=8, amp=0.1, sigma=2.0), DriveChannel(0)))
    slots, params = backend._slots_from_schedule(s)
    assert [(k, st, d) for k, _, st, d in slots] == [("gaussian", 0, 8), ("drag", 8, 8), ("gaussian", 16, 8), ("gaussian", 24, 8)]
    assert "freq_shift" not in params[0] and params[0]["angle"] == 0.0
    assert float(params[1]["angle"]) == pytest.approx(0.75) and params[1]["freq_shift"] == pytest.approx(3e6)
    assert float(params[2]["angle"]) == pytest.approx(0.5) and params[2]["freq_shift"] == pytest.approx(-2e6)
    assert float(params[3]["angle"]) == pytest.approx(-1.0) and params[3]["freq_shift"] == pytest.approx(-2e6)
    packed = engine_pack(params)
    assert packed.shape == (4, 6, 1) and packed[1, 5, 0] == 3e6 and packed[0, 5, 0] == 0.0
    with pytest.raises(CasqError, match="not in the Hamiltonian"):
        bad = Schedule()
        bad.insert(0, Play(Gaussian(duration=8, amp=0.2, sigma=2.0), DriveChannel(3)))
        backend._slots_from_schedule(bad)


def engine_pack(params):
    from casq_b200 import engine

    return engine.pack_params(params).numpy()


def test_cross_resonance_gate_and_multi_qubit_circuit_host_logic():
    from casq_b200.gates import CrossResonancePulseGate, DragPulseGate, PulseCircuit
    from casq_b200.pulse import ControlChannel, GaussianSquare

    gate = CrossResonancePulseGate(control_channel=1, duration=64, amplitude=0.4, angle=0.1, name="cr")
    assert gate.num_qubits == 2 and gate.channel_name(0) == "u1" and gate.channel(5) == ControlChannel(1)
    assert gate.to_parameters_dict(np.array([8.0, 20.0])) == {"sigma": 8.0, "width": 20.0}
    pulse = gate.pulse({"sigma": 8.0, "width": 20.0})
    assert isinstance(pulse, GaussianSquare) and pulse.duration == 64 and pulse.amp == 0.4 and pulse.angle == 0.1
    sched = gate.schedule({"sigma": 8.0, "width": 20.0}, 0)
    assert sched.channels == ["u1"] and sched.duration == 64
    circuit = PulseCircuit(2, 2)
    circuit.pulse_gate(gate, {"sigma": 8.0, "width": 20.0}, (0, 1))
    circuit.pulse_gate(DragPulseGate(duration=16, amplitude=0.2, name="x"), {"sigma": 4.0, "beta": 0.0}, 1)
    circuit.measure(0, 0)
    circuit.measure(1, 1)
    assert [i.qubits for i in circuit.data] == [(0, 1), (1,), (0,), (1,)]
    with pytest.raises(CasqError, match="acts on 2"):
        circuit.pulse_gate(gate, {"sigma": 8.0, "width": 20.0}, 0)
    with pytest.raises(CasqError, match="width"):
        gate.pulse({"sigma": 8.0, "width": 65.0})
    # ASAP scheduling through the backend: the DRAG on qubit 1 starts when the CR gate (qubits 0 and 1) ends
    backend = _cpu_backend((0, 1))
    sched, t_acq, measured = backend._circuit_to_schedule(circuit)
    assert t_acq == 80 and measured == [0, 1]
    slots, params = backend._slots_from_schedule(sched)
    ch = backend.hamiltonian.channels
    assert slots == [("gaussian_square", ch.index("u1"), 0, 64), ("drag", ch.index("d1"), 64, 16)]


def test_to_solution_forwarded_dims_density_matrix_and_iq_layout():
    """_to_solution is host-only: populations / counts / IQ for one and two measured qubits, Statevector and DensityMatrix."""
    from casq_b200.quantum_info import DensityMatrix, Statevector

    rng = np.random.default_rng(0)
    psi = rng.normal(size=9) + 1j * rng.normal(size=9)
    psi /= np.linalg.norm(psi)
    probs = np.abs(psi) ** 2
    plain = _cpu_backend((0, 1))
    fwd = _cpu_backend((0, 1), forward_subsystem_dims=True)
    a = plain._to_solution("c", np.array([0.0, 1e-9]), np.array([psi, psi]), np.array([probs, probs]), 512, [0, 1])
    assert a.qubits == [0] and a.populations[0] == pytest.approx({"0": probs[0], "1": probs[1:].sum()})
    assert isinstance(a.states[0], Statevector) and a.states[0].dims() == (9,)
    assert sum(a.counts[1].values()) == 512 and len(a.iq_data[0]) == 512 and len(a.avg_iq_data[0]) == 2
    assert a.samples[0] == a.samples[1]  # same experiment seed at every time point (helpers.py:137-141)
    b = fwd._to_solution("c", np.array([0.0]), np.array([psi]), np.array([probs]), 512, [0, 1])
    assert b.qubits == [0, 1] and b.states[0].dims() == (3, 3)
    assert b.populations[0] == pytest.approx(O.populations_dict(probs, subsystem_dims=[3, 3], measured=(0, 1)))
    assert np.array(b.iq_data[0]).shape == (512, 2, 2) and np.array(b.avg_iq_data[0]).shape == (2, 2)
    one = fwd._to_solution("c", np.array([0.0]), np.array([psi]), np.array([probs]), 256, [1])
    assert one.qubits == [1] and one.populations[0] == pytest.approx(O.populations_dict(probs, subsystem_dims=[3, 3], measured=(1,)))
    rho = np.outer(psi, psi.conj())
    d = fwd._to_solution("c", np.array([0.0]), np.array([rho]), np.array([probs]), 128, [0], density=True)
    assert isinstance(d.states[0], DensityMatrix) and d.states[0].dims() == (3, 3) and sum(d.counts[0].values()) == 128
    with pytest.raises(CasqError, match="not in the Hamiltonian"):
        fwd._to_solution("c", np.array([0.0]), np.array([psi]), np.array([probs]), 16, [4])


def test_duration_sweep_assembly_on_cpu_tensors():
    import torch
    from casq_b200.backends.b200.b200_pulse_backend import BatchSolution, DurationSweepSolution

    def part(n, dur):
        return BatchSolution(times=np.array([0.0, dur * 1.0]), states=torch.full((n, 2, 3), float(dur), dtype=torch.complex128),
                             probabilities=torch.full((n, 2, 3), float(dur), dtype=torch.float64),
                             populations=torch.full((n, 2, 2), float(dur), dtype=torch.float64))

    sol = DurationSweepSolution.assemble(5, {16: (np.array([0, 3]), part(2, 16)), 32: (np.array([1, 2, 4]), part(3, 32))})
    assert sol.batch_size == 5 and sol.durations.tolist() == [16, 32, 32, 16, 32]
    assert sol.times[:, 1].tolist() == [16.0, 32.0, 32.0, 16.0, 32.0]
    assert sol.states[:, 0, 0].real.tolist() == [16.0, 32.0, 32.0, 16.0, 32.0] and sol.populations[3, 1, 1] == 16.0
