"""Identifier helpers (mirrors dbid/ufid of src/casq/common/helpers.py:50-72; no wonderwords dependency)."""
from __future__ import annotations

import random
from typing import Any
from uuid import uuid4

_ADJECTIVES = ("amber", "brisk", "calm", "deft", "eager", "fleet", "gentle", "hardy", "ivory", "jolly", "keen",
               "lucid", "merry", "noble", "opal", "proud", "quick", "rapid", "sunny", "tidy", "umber", "vivid",
               "warm", "young", "zesty")


def dbid() -> str:
    """Database identifier (UUID4 string)."""
    return str(uuid4())


def ufid(obj: Any) -> str:
    """User-friendly identifier: <adjective><ClassName>."""
    name = obj.__class__.__name__
    if not name[0].isupper():
        name = name.capitalize()
    return f"{random.choice(_ADJECTIVES)}{name}"


# ---------------------------------------------------------------------------------------------------------------
# discretize: schedule -> per-channel sampled signals (mirrors src/casq/common/helpers.py:39-48,90-144).  The samples
# come from the device envelope sampler (casq_envelope_sample), i.e. the same code the steppers evaluate on the fly.
from dataclasses import dataclass  # noqa: E402
from enum import Enum  # noqa: E402
from typing import Optional  # noqa: E402


class TimeUnit(Enum):
    """Time units."""

    SAMPLE = 0
    PICO_SEC = 1
    NANO_SEC = 2
    MICRO_SEC = 3
    MILLI_SEC = 4
    SEC = 5


@dataclass
class DiscreteSignal:
    """Piecewise-constant complex envelope on a carrier: s(t) = Re[samples[floor(t/dt)] e^{i(2 pi f t + phase)}]."""

    dt: float
    samples: Any  # numpy complex128 [N]
    carrier_freq: float
    phase: float = 0.0
    name: str = ""

    @property
    def duration(self) -> int:
        return len(self.samples)


@dataclass
class SignalData:
    """Same fields as casq's SignalData (helpers.py:90-101)."""

    name: str
    dt: float
    duration: float
    signal: DiscreteSignal
    i_signal: DiscreteSignal
    q_signal: DiscreteSignal
    carrier: float


def discretize(schedule, dt: float, channel_frequencies: dict, carrier_frequency: Optional[float] = None) -> list:
    """One SignalData per channel of ``channel_frequencies`` (silent channels carry zeros), sampled at the midpoints
    n + 1/2 on the GPU; i/q components as upstream ``get_awg_signals`` builds them (I = samples, Q = Im - i Re, carrier
    shifted by the IF modulation)."""
    import numpy as np

    from .. import engine

    n_total = schedule.duration
    out = []
    for name, freq in channel_frequencies.items():
        samples = np.zeros(n_total, dtype=np.complex128)
        for start, inst in schedule.plays():
            if inst.channel.name != name:
                continue
            params = engine.pack_params([inst.pulse.engine_params()], device="cuda")[0]
            env = engine.envelope_sample(inst.pulse.kind, inst.pulse.duration, params)[0].cpu().numpy()
            samples[start:start + inst.pulse.duration] += env
        carrier = carrier_frequency if carrier_frequency else freq
        sig = DiscreteSignal(dt, samples, freq, 0.0, name)
        sig_i = DiscreteSignal(dt, samples.copy(), freq + carrier, 0.0, f"{name}_i")
        sig_q = DiscreteSignal(dt, samples.imag - 1j * samples.real, freq + carrier, 0.0, f"{name}_q")
        out.append(SignalData(name, dt, n_total, sig, sig_i, sig_q, carrier))
    return out
