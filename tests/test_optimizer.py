"""PulseOptimizer: host-side checks on CPU, the scipy loop and the batched population search on the GPU
(the reference's own test only asserts the return type: tests/casq/pulse_optimizer_test.py:86-106)."""
import numpy as np
import pytest

import oracle as O
from casq_b200.backends import BackendLibrary, PulseBackend, build
from casq_b200.common.exceptions import CasqError
from casq_b200.gates import DragPulseGate, GaussianPulseGate
from casq_b200.models import ControlModel, TransmonModel
from casq_b200.optimizers import PulseOptimizer, XGateOptimizer, hellinger_fidelity

DT = O.MANILA_DT


def _backend(seed=3):
    h = TransmonModel.manila(extracted_qubits=[0])
    return build(BackendLibrary.B200, h, ControlModel(dt=DT, channel_carrier_freqs=O.MANILA_CARRIERS), seed=seed)


def test_hellinger_fidelity_and_ctor_checks(native_lib):
    assert hellinger_fidelity({"0": 0, "1": 1024}, {"0": 256, "1": 768}) == pytest.approx(0.75)
    assert hellinger_fidelity({"0": 1, "1": 1}, {"0": 5, "1": 5}) == pytest.approx(1.0)
    assert hellinger_fidelity({"0": 10}, {"1": 10}) == 0.0
    backend = _backend()
    with pytest.raises(CasqError, match="PulseGate"):
        XGateOptimizer("gate", backend, PulseBackend.ODESolverMethod.SCIPY_DOP853)
    with pytest.raises(CasqError, match="PulseBackend"):
        XGateOptimizer(GaussianPulseGate(16, 1), "backend", PulseBackend.ODESolverMethod.SCIPY_DOP853)
    opt = XGateOptimizer(GaussianPulseGate(16, 1), backend, PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT)
    assert opt.target_measurement == {"0": 0, "1": 1024} and opt.target_qubit == 0 and opt.use_jit is False
    assert {m.value for m in PulseOptimizer.OptimizationMethod} >= {"Nelder-Mead", "BFGS", "COBYLA", "trust-constr"}


@pytest.mark.gpu
def test_x_gate_optimizer_scipy_loop():
    """Same call as the reference test: Nelder-Mead, maxiter=10, tol=1 -> a Solution; plus sanity of its content."""
    backend = _backend()
    opt = XGateOptimizer(GaussianPulseGate(duration=64, amplitude=0.45), backend, PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_RK4,
                         method_options={"max_dt": DT / 20})
    sol = opt.solve(np.array([10.0]), PulseOptimizer.OptimizationMethod.SCIPY_NELDER_MEAD, maxiter=10, tol=1, verbose=False)
    assert isinstance(sol, PulseOptimizer.Solution)
    assert sol.num_iterations == len(sol.iterations) >= 2 and set(sol.parameters) == {"sigma"}
    assert 0.0 <= sol.fidelity <= 1.0 and sum(sol.measurement.values()) == 1024
    assert sol.iterations[0].parameters == {"sigma": 10.0} and sol.pulse.pulse_type == "Gaussian"


@pytest.mark.gpu
def test_population_search_finds_x_gate():
    """BASELINE cfg 5 in miniature: DRAG candidates (sigma, beta) at fixed amplitude, adaptive Dopri5 with the
    optimizer's documented options; the cross-entropy search must reach infidelity < 1e-3 and agree with the oracle."""
    backend = _backend()
    gate = DragPulseGate(duration=64, amplitude=0.45)
    opt = XGateOptimizer(gate, backend, PulseBackend.ODESolverMethod.QISKIT_DYNAMICS_JAX_ODEINT,
                         method_options={"atol": 1e-6, "rtol": 1e-8, "hmax": DT})
    res = opt.solve_population(lower=[6.0, -4.0], upper=[30.0, 4.0], population=512, generations=6, seed=11)
    assert res.evaluations == 512 * 6 and len(res.history) == 6
    assert all(a >= b for a, b in zip(res.history, res.history[1:]))
    assert res.infidelity < 1e-3 and res.leakage < 1e-3
    # oracle check of the winner (tight-tolerance truth)
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, DT, rotating_frame=st)
    ref = O.casq_solve(om, [O.PulseSpec("drag", "d0", 64)], [dict(amp=0.45, angle=0.0, **res.parameters)], method="DOP853",
                       run_options={"rtol": 1e-11, "atol": 1e-11})
    p1 = ref["populations"][-1].get("1", 0.0)
    assert abs((1 - p1) - res.infidelity) < 1e-4
    # batched objective == sequential objective in the shots -> infinity limit
    f, leak = opt.evaluate_batch(np.array([[12.0, 1.0], [20.0, -2.0]]))
    for (sigma, beta), fb in zip([(12.0, 1.0), (20.0, -2.0)], f):
        r = O.casq_solve(om, [O.PulseSpec("drag", "d0", 64)], [dict(amp=0.45, angle=0.0, sigma=sigma, beta=beta)], method="jax_odeint",
                         run_options={"atol": 1e-6, "rtol": 1e-8, "hmax": DT})
        assert abs((1 - r["populations"][-1].get("1", 0.0)) - fb) < 1e-5
