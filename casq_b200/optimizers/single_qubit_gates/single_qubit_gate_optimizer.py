"""Single-qubit gate optimizer (drop-in for src/casq/optimizers/single_qubit_gates/single_qubit_gate_optimizer.py)."""
from __future__ import annotations

from typing import Any, Optional, Union

from ...backends.pulse_backend import PulseBackend
from ...gates import PulseCircuit, PulseGate
from ...quantum_info import DensityMatrix, Statevector
from ..pulse_optimizer import PulseOptimizer


class SingleQubitGateOptimizer(PulseOptimizer):
    """Optimizes the shape parameters of one pulse gate on one qubit followed by a measurement (:74-85)."""

    def __init__(self, pulse_gate: PulseGate, pulse_backend: PulseBackend, method: PulseBackend.ODESolverMethod,
                 target_measurement: Union[dict[str, float], DensityMatrix, Statevector],
                 method_options: Optional[dict[str, Any]] = None, fidelity_type: Optional[PulseOptimizer.FidelityType] = None,
                 target_qubit: Optional[int] = None, use_jit: bool = False):
        super().__init__(pulse_gate=pulse_gate, pulse_backend=pulse_backend, method=method, target_measurement=target_measurement,
                         method_options=method_options, fidelity_type=fidelity_type, target_qubit=target_qubit, use_jit=use_jit)

    def _build_circuit(self, parameters: dict[str, Any]) -> PulseCircuit:
        return PulseCircuit.from_pulse_gate(self.pulse_gate, parameters)
