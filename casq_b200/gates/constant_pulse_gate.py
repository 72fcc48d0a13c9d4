"""Constant pulse gate (drop-in for src/casq/gates/constant_pulse_gate.py)."""
from __future__ import annotations

from typing import Any, Optional

import numpy.typing as npt

from ..pulse import Constant, ScalableSymbolicPulse
from .pulse_gate import PulseGate


class ConstantPulseGate(PulseGate):
    """ConstantPulseGate: forwards to the ``Constant`` envelope family; optimisable parameters [].

    Args: duration (samples), amplitude, angle (default 0), limit_amplitude, name.
    """

    def __init__(self, duration: int, amplitude: float, angle: Optional[float] = 0,
                 limit_amplitude: bool = True, name: Optional[str] = None) -> None:
        super().__init__(1, duration, amplitude, angle, limit_amplitude, name)

    def pulse(self, parameters: Optional[dict[str, Any]] = None) -> ScalableSymbolicPulse:
        return Constant(duration=self.duration, amp=self.amplitude, angle=self.angle,
                     limit_amplitude=self.limit_amplitude, name=self.name, **(parameters or {}))

    def to_parameters_dict(self, parameters: npt.NDArray) -> dict[str, Any]:
        """Array in order [] -> dictionary."""
        return {}
