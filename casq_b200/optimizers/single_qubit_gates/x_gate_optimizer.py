"""X-gate optimizer (drop-in for src/casq/optimizers/single_qubit_gates/x_gate_optimizer.py:36-68)."""
from __future__ import annotations

from typing import Any, Optional

from ...backends.pulse_backend import PulseBackend
from ...gates import PulseGate
from ..pulse_optimizer import PulseOptimizer
from .single_qubit_gate_optimizer import SingleQubitGateOptimizer


class XGateOptimizer(SingleQubitGateOptimizer):
    """Target measurement fixed to {"0": 0, "1": 1024} (x_gate_optimizer.py:59)."""

    def __init__(self, pulse_gate: PulseGate, pulse_backend: PulseBackend, method: PulseBackend.ODESolverMethod,
                 method_options: Optional[dict[str, Any]] = None, fidelity_type: Optional[PulseOptimizer.FidelityType] = None,
                 target_qubit: Optional[int] = None, use_jit: bool = False):
        super().__init__(pulse_gate=pulse_gate, pulse_backend=pulse_backend, method=method,
                         target_measurement={"0": 0, "1": 1024}, method_options=method_options, fidelity_type=fidelity_type,
                         target_qubit=target_qubit, use_jit=use_jit)
