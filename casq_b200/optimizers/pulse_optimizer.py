"""Pulse optimizer (drop-in for src/casq/optimizers/pulse_optimizer.py:60-543, plots excluded) plus the
population API the B200 engine makes worthwhile.

``solve`` keeps casq's semantics: scipy.optimize.minimize over the gate's shape parameters, objective
``1 - hellinger_fidelity(target, counts[-1])`` of ONE ``pulse_backend.solve`` per evaluation
(pulse_optimizer.py:450-505).  ``evaluate_batch`` / ``solve_population`` are additive: a whole
candidate population is one ``solve_batch`` launch, and the objective is the deterministic population
``1 - P("1")`` with casq's leakage folding (plus the true ``1 - |psi_1|^2``) instead of sampled counts.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Any, Callable, Optional, Union

import numpy as np
import numpy.typing as npt
from scipy.optimize import minimize

from ..backends.pulse_backend import PulseBackend
from ..common.exceptions import CasqError
from ..gates import PulseCircuit, PulseGate
from ..pulse import ScalableSymbolicPulse
from ..quantum_info import DensityMatrix, Statevector


def hellinger_fidelity(dist_p: dict, dist_q: dict) -> float:
    """(sum_i sqrt(p_i q_i))^2 on normalised counts (qiskit.quantum_info.hellinger_fidelity, used at
    pulse_optimizer.py:502)."""
    p_sum, q_sum = sum(dist_p.values()), sum(dist_q.values())
    if p_sum <= 0 or q_sum <= 0:
        raise CasqError("hellinger_fidelity needs non-empty distributions.")
    total = 0.0
    for key, value in dist_p.items():
        if key in dist_q:
            total += np.sqrt((value / p_sum) * (dist_q[key] / q_sum))
    return float(total**2)


class PulseOptimizer:
    """PulseOptimizer(pulse_gate, pulse_backend, method, target_measurement, method_options, fidelity_type,
    target_qubit, use_jit) -- same arguments as casq."""

    class PulseType(Enum):
        DRAG = 0
        GAUSSIAN = 1
        GAUSSIAN_SQUARE = 2

    class FidelityType(Enum):
        COUNTS = 0

    class OptimizationMethod(str, Enum):
        SCIPY_BFGS = "BFGS"
        SCIPY_CG = "CG"
        SCIPY_COBYLA = "COBYLA"
        SCIPY_DOGLEG = "dogleg"
        SCIPY_L_BFGS_B = "L-BFGS-B"
        SCIPY_NEWTON_CG = "Newton-CG"
        SCIPY_NELDER_MEAD = "Nelder-Mead"
        SCIPY_POWELL = "Powell"
        SCIPY_SLSQP = "SLSQP"
        SCIPY_TNC = "TNC"
        SCIPY_TRUST_CONSTR = "trust-constr"
        SCIPY_TRUST_EXACT = "trust-exact"
        SCIPY_TRUST_KRYLOV = "trust-krylov"
        SCIPY_TRUST_NCG = "trust-ncg"

    class FiniteDifferenceScheme(str, Enum):
        CS = "cs"
        TWO_POINT = "2-point"
        THREE_POINT = "3-point"

    @dataclass
    class Iteration:
        index: int
        parameters: dict[str, Any]
        result: Any
        objective: float

    @dataclass
    class Solution:
        initial_parameters: dict[str, Any]
        initial_pulse: ScalableSymbolicPulse
        num_iterations: int
        iterations: list["PulseOptimizer.Iteration"]
        parameters: dict[str, Any]
        measurement: Union[dict[str, float], DensityMatrix, Statevector]
        fidelity: float
        pulse: ScalableSymbolicPulse
        gate: PulseGate
        circuit: PulseCircuit
        message: str

    @dataclass
    class PopulationSolution:
        """Result of ``solve_population``."""

        parameters: dict[str, Any]
        infidelity: float
        leakage: float
        generations: int
        evaluations: int
        history: list[float]  # best infidelity per generation
        population: np.ndarray  # last generation [B, P]
        objectives: np.ndarray  # its infidelities [B]

    def __init__(
        self,
        pulse_gate: PulseGate,
        pulse_backend: PulseBackend,
        method: PulseBackend.ODESolverMethod,
        target_measurement: Union[dict[str, float], DensityMatrix, Statevector],
        method_options: Optional[dict[str, Any]] = None,
        fidelity_type: Optional["PulseOptimizer.FidelityType"] = None,
        target_qubit: Optional[int] = None,
        use_jit: bool = False,
    ):
        if not isinstance(pulse_gate, PulseGate):
            raise CasqError("pulse_gate must be a PulseGate.")
        if not isinstance(pulse_backend, PulseBackend):
            raise CasqError("pulse_backend must be a PulseBackend.")
        self.pulse_gate = pulse_gate
        self.pulse_backend = pulse_backend
        self.method = method
        self.target_measurement = target_measurement
        self.method_options = method_options if method_options else {}
        self.fidelity_type = PulseOptimizer.FidelityType.COUNTS if fidelity_type is None else fidelity_type
        self.target_qubit = target_qubit if target_qubit else self.pulse_backend.hamiltonian.extracted_qubits[0]
        # casq force-disables jit (pulse_optimizer.py:355-359) and rejects jax methods when jax is off (:360-367);
        # the device engine needs no jax, so every method of the enum is accepted here.
        self.use_jit = False
        self.objective_function = self._build_objective_function()
        self._counter: int = 0
        self._iterations: list[PulseOptimizer.Iteration] = []

    # ------------------------------------------------------------------ casq's sequential search
    def solve(self, initial_params: npt.NDArray, method: "PulseOptimizer.OptimizationMethod", jac=None, hess=None, hessp=None,
              bounds=None, constraints=None, tol: Optional[float] = None, maxiter: Optional[int] = None,
              verbose: bool = True) -> "PulseOptimizer.Solution":
        """Wrapper around scipy.optimize.minimize, one ``pulse_backend.solve`` per objective evaluation
        (pulse_optimizer.py:371-482)."""
        M = PulseOptimizer.OptimizationMethod
        options: dict = {"disp": verbose}
        if maxiter:
            options.update(maxiter=maxiter)
        if method not in [M.SCIPY_COBYLA, M.SCIPY_L_BFGS_B, M.SCIPY_NELDER_MEAD, M.SCIPY_POWELL, M.SCIPY_SLSQP, M.SCIPY_TNC,
                          M.SCIPY_TRUST_CONSTR]:
            bounds = None
        if method not in [M.SCIPY_COBYLA, M.SCIPY_SLSQP, M.SCIPY_TRUST_CONSTR]:
            constraints = None
        self._counter = 0
        self._iterations = []
        kwargs = dict(fun=self.objective_function, x0=np.asarray(initial_params, dtype=float), method=method.value, jac=jac,
                      hess=hess, hessp=hessp, bounds=bounds, tol=tol, options=options)
        if constraints is not None:
            kwargs["constraints"] = constraints
        opt_results = minimize(**kwargs)
        initial_parameters = self.pulse_gate.to_parameters_dict(initial_params)
        initial_pulse = self.pulse_gate.pulse(initial_parameters)
        parameters = self.pulse_gate.to_parameters_dict(opt_results.x)
        opt_pulse = self.pulse_gate.pulse(parameters)
        opt_circuit = PulseCircuit.from_pulse_gate(self.pulse_gate, parameters)
        counts = self.pulse_backend.solve(circuit=opt_circuit, method=self.method, run_options=dict(self.method_options)).counts[-1]
        return PulseOptimizer.Solution(
            initial_parameters=initial_parameters, initial_pulse=initial_pulse, num_iterations=opt_results.nfev,
            iterations=self._iterations, parameters=parameters, measurement=counts, fidelity=1 - opt_results.fun,
            gate=self.pulse_gate, pulse=opt_pulse, circuit=opt_circuit, message=str(opt_results.message))

    def _build_circuit(self, parameters: dict[str, Any]) -> PulseCircuit:
        return PulseCircuit.from_pulse_gate(self.pulse_gate, parameters)

    def _build_objective_function(self) -> Callable[[npt.NDArray], float]:
        def objective(parameters: npt.NDArray) -> float:
            parameters_dict = self.pulse_gate.to_parameters_dict(parameters)
            circuit = self._build_circuit(parameters_dict)
            solution = self.pulse_backend.solve(circuit=circuit, method=self.method, run_options=dict(self.method_options))
            counts = solution.counts[-1]
            infidelity = 1.0 - hellinger_fidelity(self.target_measurement, counts)
            self._objective_callback(parameters_dict, counts, infidelity)
            return infidelity

        return objective

    def _objective_callback(self, parameters: dict[str, Any], result: Any, objective: float) -> None:
        self._counter += 1
        self._iterations.append(PulseOptimizer.Iteration(index=self._counter, parameters=parameters, result=result,
                                                         objective=objective))

    # ------------------------------------------------------------------ batched (additive) API
    def _target_probability_vector(self) -> np.ndarray:
        tm = self.target_measurement
        if not isinstance(tm, dict):
            raise CasqError("Batched objectives need a counts/probability dictionary as target_measurement.")
        total = float(sum(tm.values()))
        return np.array([tm.get("0", 0.0) / total, tm.get("1", 0.0) / total])

    def evaluate_batch(self, parameters: npt.NDArray, amplitude=None, angle=None):
        """Objective for B candidates at once: [B, P] array in ``to_parameters_dict`` order.

        Returns (infidelity [B], leakage [B]) as numpy arrays: infidelity = 1 - Hellinger fidelity between the target
        distribution and the EXACT populations {"0","1"} (leakage folded into "1" like casq, helpers.py:129-134), i.e.
        the shots -> infinity limit of casq's objective; leakage = population outside levels 0 and 1."""
        params = np.atleast_2d(np.asarray(parameters, dtype=float))
        sol = self.pulse_backend.solve_batch(self.pulse_gate, params, self.method, qubit=self.target_qubit, amplitude=amplitude,
                                             angle=angle, run_options=dict(self.method_options))
        pops = sol.populations[:, -1, :].cpu().numpy()
        probs = sol.probabilities[:, -1, :].cpu().numpy()
        target = self._target_probability_vector()
        fid = (np.sqrt(np.clip(pops, 0, None) * target[None, :]).sum(axis=1)) ** 2
        bad = np.zeros(len(params), dtype=bool)
        if sol.status is not None:
            bad = sol.status.cpu().numpy() != 0
        infid = np.where(bad, 1.0, 1.0 - fid)
        leakage = probs[:, 2:].sum(axis=1) if probs.shape[1] > 2 else np.zeros(len(params))
        return infid, leakage

    def solve_population(self, lower: npt.NDArray, upper: npt.NDArray, population: int = 1024, generations: int = 8,
                         elite_fraction: float = 0.1, seed: Optional[int] = None,
                         initial_population: Optional[npt.NDArray] = None) -> "PulseOptimizer.PopulationSolution":
        """Cross-entropy search inside the box [lower, upper]: every generation is ONE batched solve of ``population``
        candidates; the next generation is sampled from a Gaussian fitted to the elite."""
        lower, upper = np.asarray(lower, dtype=float), np.asarray(upper, dtype=float)
        if lower.shape != upper.shape or np.any(upper <= lower):
            raise CasqError("solve_population needs lower < upper with equal shapes.")
        rng = np.random.default_rng(seed)
        P = lower.size
        pop = rng.uniform(lower, upper, size=(population, P)) if initial_population is None else np.asarray(initial_population, dtype=float)
        n_elite = max(2, int(round(elite_fraction * len(pop))))
        history, evaluations = [], 0
        best_x, best_f, best_leak = None, np.inf, np.nan
        for _ in range(generations):
            f, leak = self.evaluate_batch(pop)
            evaluations += len(pop)
            order = np.argsort(f)
            if f[order[0]] < best_f:
                best_x, best_f, best_leak = pop[order[0]].copy(), float(f[order[0]]), float(leak[order[0]])
            history.append(best_f)
            elite = pop[order[:n_elite]]
            mean, std = elite.mean(axis=0), elite.std(axis=0) + 1e-6 * (upper - lower)
            last_pop, last_f = pop, f
            pop = np.clip(rng.normal(mean, std, size=(len(pop), P)), lower, upper)
            pop[0] = best_x  # keep the incumbent
        return PulseOptimizer.PopulationSolution(
            parameters=self.pulse_gate.to_parameters_dict(best_x), infidelity=best_f, leakage=best_leak, generations=generations,
            evaluations=evaluations, history=history, population=last_pop, objectives=last_f)
