"""Pulse gate base class (drop-in for src/casq/gates/pulse_gate.py:39-112)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional

import numpy.typing as npt

from ..common.helpers import dbid, ufid
from ..pulse import DriveChannel, Play, ScalableSymbolicPulse, Schedule


class PulseGate(ABC):
    """Abstract base class for all pulse gates (single-qubit, played on ``DriveChannel(qubit)``).

    Args mirror casq: num_qubits, duration (samples), amplitude, angle, limit_amplitude, name.
    """

    def __init__(self, num_qubits: int, duration: int, amplitude: float, angle: Optional[float] = 0,
                 limit_amplitude: bool = True, name: Optional[str] = None) -> None:
        self.dbid = dbid()
        self.ufid = name if name else ufid(self)
        self.name = self.ufid
        self.label = self.ufid
        self.num_qubits = num_qubits
        self.params: list = []
        self.duration = duration
        self.amplitude = amplitude
        self.angle = angle
        self.limit_amplitude = limit_amplitude

    @abstractmethod
    def pulse(self, params: dict[str, Any]) -> ScalableSymbolicPulse:
        """Builds the library pulse for this gate."""

    @abstractmethod
    def to_parameters_dict(self, parameters: npt.NDArray) -> dict[str, Any]:
        """Builds the parameter dictionary from an array of parameter values."""

    def channel(self, qubit: int):
        """The channel the gate plays on: DriveChannel(qubit) for every casq gate (pulse_gate.py:111); multi-qubit gates
        (CrossResonancePulseGate) override it with their ControlChannel."""
        return DriveChannel(qubit)

    def channel_name(self, qubit: int) -> str:
        return self.channel(qubit).name

    def schedule(self, parameters: dict[str, Any], qubit: int) -> Schedule:
        """One-instruction schedule playing the gate's pulse on DriveChannel(qubit) (pulse_gate.py:97-112)."""
        sched = Schedule(name=f"{self.name}Schedule")
        sched.insert(0, Play(self.pulse(parameters), self.channel(qubit)))
        return sched
