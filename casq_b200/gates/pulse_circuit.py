"""Pulse circuit (drop-in for src/casq/gates/pulse_circuit.py:36-98, without the qiskit QuantumCircuit base)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

from ..common.exceptions import CasqError
from ..common.helpers import dbid, ufid
from ..pulse import Schedule
from .pulse_gate import PulseGate


@dataclass
class CircuitInstruction:
    operation: Any  # a PulseGate or the string "measure"
    qubits: tuple
    clbits: tuple = ()


class PulseCircuit:
    """A register of qubits/clbits, an ordered instruction list, and per-(gate, qubits) pulse calibrations."""

    @staticmethod
    def from_pulse_gate(gate: PulseGate, parameters: dict[str, Any]) -> "PulseCircuit":
        """1 qubit, 1 clbit: the gate followed by a measurement (pulse_circuit.py:48-64)."""
        circuit = PulseCircuit(1, 1)
        circuit.pulse_gate(gate, parameters, 0)
        circuit.measure(0, 0)
        return circuit

    def __init__(self, num_qubits: int = 1, num_clbits: int = 0, name: Optional[str] = None, metadata: Optional[dict] = None):
        self.dbid = dbid()
        self.ufid = ufid(self)
        self.name = self.ufid if name is None else name
        self.num_qubits = int(num_qubits)
        self.num_clbits = int(num_clbits)
        self.metadata = metadata
        self.data: list[CircuitInstruction] = []
        self.calibrations: dict = {}

    def append(self, gate: PulseGate, qargs) -> CircuitInstruction:
        qargs = tuple(int(q) for q in qargs)
        for q in qargs:
            if not 0 <= q < self.num_qubits:
                raise CasqError(f"Qubit index {q} out of range for a {self.num_qubits}-qubit circuit.")
        inst = CircuitInstruction(gate, qargs)
        self.data.append(inst)
        return inst

    def add_calibration(self, gate_name: str, qubits, schedule: Schedule, params=None) -> None:
        self.calibrations.setdefault(gate_name, {})[(tuple(qubits), tuple(params or ()))] = schedule

    def pulse_gate(self, gate: PulseGate, parameters: dict[str, Any], qubit=0) -> CircuitInstruction:
        """Append the gate and attach its schedule as the calibration (pulse_circuit.py:81-98).  ``qubit`` is an int for
        casq's single-qubit gates; a multi-qubit gate (CrossResonancePulseGate) takes a sequence (control, target)."""
        qubits = [qubit] if isinstance(qubit, int) else [int(q) for q in qubit]
        if len(qubits) != gate.num_qubits:
            raise CasqError(f"Gate '{gate.name}' acts on {gate.num_qubits} qubit(s), got {qubits}.")
        inst = self.append(gate, qubits)
        self.add_calibration(gate.name, qubits, gate.schedule(parameters, qubits[0]), [])
        return inst

    def measure(self, qubit: int, clbit: int) -> CircuitInstruction:
        inst = CircuitInstruction("measure", (int(qubit),), (int(clbit),))
        self.data.append(inst)
        return inst

    def measured_qubits(self) -> list:
        return [i.qubits[0] for i in self.data if i.operation == "measure"]
