"""Minimal pulse-schedule objects standing in for the slice of ``qiskit.pulse`` that casq touches.

qiskit is not a dependency of casq_b200.  casq's gates only ever build a one-instruction schedule
``play(<library pulse>, DriveChannel(qubit))`` (src/casq/gates/pulse_gate.py:97-112), so this module has
the five library pulse classes with qiskit's constructor arguments and validation rules (SURVEY Appendix
A.2), channels, ``Play`` and ``Schedule``.  Envelope *samples* are never computed here: the device
sampler (casq_envelope_sample / the fused steppers) does that.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from .common.exceptions import CasqError


@dataclass(frozen=True)
class Channel:
    index: int
    prefix = "ch"

    @property
    def name(self) -> str:
        return f"{self.prefix}{self.index}"


class DriveChannel(Channel):
    prefix = "d"


class ControlChannel(Channel):
    prefix = "u"


class MeasureChannel(Channel):
    prefix = "m"


class ScalableSymbolicPulse:
    """A parametrised envelope family; parameters may be scalars or equal-length arrays (a sweep)."""

    kind = "constant"
    _param_names: tuple = ()

    def __init__(self, duration: int, amp, angle=0.0, limit_amplitude: bool = True, name: Optional[str] = None, **params: Any):
        if int(duration) != duration or duration < 1:
            raise CasqError(f"Pulse duration must be a positive integer, got {duration}.")
        self.duration = int(duration)
        self.amp = amp
        self.angle = 0.0 if angle is None else angle
        self.limit_amplitude = limit_amplitude
        self.name = name
        unknown = set(params) - set(self._param_names)
        if unknown:
            raise CasqError(f"{type(self).__name__} got unexpected parameter(s) {sorted(unknown)}.")
        missing = [p for p in self._param_names if p not in params]
        if missing:
            raise CasqError(f"{type(self).__name__} needs parameter(s) {missing}.")
        self.params = dict(params)
        self.validate_parameters()

    @property
    def pulse_type(self) -> str:
        return type(self).__name__

    def validate_parameters(self) -> None:
        amp = np.asarray(self.amp, dtype=float)
        if self.limit_amplitude and np.any(np.abs(amp) > 1.0):
            raise CasqError("The amplitude norm must be <= 1, found a larger value (limit_amplitude=True).")
        if "sigma" in self.params and np.any(np.asarray(self.params["sigma"], dtype=float) <= 0):
            raise CasqError("Pulse parameter sigma must be > 0.")
        if "width" in self.params:
            w = np.asarray(self.params["width"], dtype=float)
            if np.any(w < 0) or np.any(w > self.duration):
                raise CasqError("Pulse parameter width must lie in [0, duration].")

    def engine_params(self) -> dict:
        """The (amp, angle, sigma, width, beta) block the C ABI takes for this pulse."""
        out = {"amp": self.amp, "angle": self.angle}
        out.update(self.params)
        return out

    def __repr__(self) -> str:
        extra = ", ".join(f"{k}={v}" for k, v in self.params.items())
        return f"{self.pulse_type}(duration={self.duration}, amp={self.amp}, angle={self.angle}{', ' + extra if extra else ''})"


class Constant(ScalableSymbolicPulse):
    kind = "constant"


class Gaussian(ScalableSymbolicPulse):
    kind = "gaussian"
    _param_names = ("sigma",)


class Drag(ScalableSymbolicPulse):
    kind = "drag"
    _param_names = ("sigma", "beta")


class GaussianSquare(ScalableSymbolicPulse):
    kind = "gaussian_square"
    _param_names = ("sigma", "width")


class GaussianSquareDrag(ScalableSymbolicPulse):
    kind = "gaussian_square_drag"
    _param_names = ("sigma", "width", "beta")


@dataclass
class Play:
    pulse: ScalableSymbolicPulse
    channel: Channel

    @property
    def duration(self) -> int:
        return self.pulse.duration


@dataclass
class ShiftFrequency:
    """qiskit.pulse.ShiftFrequency: later Plays on `channel` are detuned by `frequency` Hz (accumulates).  casq's own gates
    never emit it; it is what a per-point detuning of ``solve_batch`` means for a single circuit."""

    frequency: float
    channel: Channel
    duration: int = 0


@dataclass
class SetFrequency:
    """qiskit.pulse.SetFrequency: later Plays on `channel` run at `frequency` Hz (shift = frequency - carrier)."""

    frequency: float
    channel: Channel
    duration: int = 0


@dataclass
class ShiftPhase:
    """qiskit.pulse.ShiftPhase: later Plays on `channel` are multiplied by exp(i phase) (accumulates)."""

    phase: float
    channel: Channel
    duration: int = 0


@dataclass
class SetPhase:
    phase: float
    channel: Channel
    duration: int = 0


@dataclass
class Acquire:
    duration: int
    qubit: int


@dataclass
class Schedule:
    """An ordered list of (start_sample, instruction)."""

    name: str = "schedule"
    instructions: list = field(default_factory=list)

    def insert(self, start: int, inst) -> "Schedule":
        self.instructions.append((int(start), inst))
        return self

    def append(self, inst) -> "Schedule":
        """Append after everything already scheduled on the same channel."""
        ch = getattr(inst, "channel", None)
        start = 0
        for s, i in self.instructions:
            if getattr(i, "channel", None) == ch:
                start = max(start, s + i.duration)
        return self.insert(start, inst)

    @property
    def duration(self) -> int:
        return max((s + i.duration for s, i in self.instructions), default=0)

    @property
    def channels(self) -> list:
        return sorted({i.channel.name for _, i in self.instructions if hasattr(i, "channel")})

    def plays(self) -> list:
        return [(s, i) for s, i in self.instructions if isinstance(i, Play)]
