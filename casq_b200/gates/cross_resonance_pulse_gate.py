"""Cross-resonance pulse gate: a GaussianSquare envelope played on a ControlChannel.

NOT in the reference: casq's PulseGate is single-qubit and always plays on DriveChannel(qubit)
(src/casq/gates/pulse_gate.py:43,111), although TransmonModel emits the cross-drive terms ``...||U{k}``
(src/casq/models/transmon_model.py:160-161) and the roadmap lists multi-qubit gates
(docs/source/roadmap.rst:36).  This is SURVEY section 8(f) rank 4: the gate that makes BASELINE cfg 3 (cross-resonance
GaussianSquare sweeps on two coupled transmons) reachable through the gate API instead of raw slot tuples.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy.typing as npt

from ..pulse import ControlChannel, GaussianSquare, ScalableSymbolicPulse
from .pulse_gate import PulseGate


class CrossResonancePulseGate(PulseGate):
    """Two-qubit gate (control, target): GaussianSquare on ControlChannel(control_channel), the channel whose LO is the
    TARGET qubit's drive frequency and whose operator acts on the CONTROL qubit.  Optimisable parameters ['sigma', 'width'].

    Args: control_channel (index k of ``u{k}``), duration (samples), amplitude, angle, limit_amplitude, name.
    """

    def __init__(self, control_channel: int, duration: int, amplitude: float, angle: Optional[float] = 0,
                 limit_amplitude: bool = True, name: Optional[str] = None) -> None:
        super().__init__(2, duration, amplitude, angle, limit_amplitude, name)
        self.control_channel = int(control_channel)

    def channel(self, qubit: int):  # noqa: ARG002 - the control channel is fixed by the coupling, not by the qubit argument
        return ControlChannel(self.control_channel)

    def pulse(self, parameters: Optional[dict[str, Any]] = None) -> ScalableSymbolicPulse:
        return GaussianSquare(duration=self.duration, amp=self.amplitude, angle=self.angle,
                              limit_amplitude=self.limit_amplitude, name=self.name, **(parameters or {}))

    def to_parameters_dict(self, parameters: npt.NDArray) -> dict[str, Any]:
        """Array in order ['sigma', 'width'] -> dictionary."""
        return {"sigma": parameters[0], "width": parameters[1]}
