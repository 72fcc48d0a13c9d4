"""Control model (drop-in for src/casq/models/control_model.py:30-50)."""
from __future__ import annotations

from typing import Optional


class ControlModel:
    """dt (sampling interval, s), channel carrier frequencies (Hz) and the control-channel map."""

    def __init__(self, dt: float, channel_carrier_freqs: dict, control_channel_map: Optional[dict] = None) -> None:
        self.dt = dt
        self.channel_carrier_freqs = channel_carrier_freqs
        self.control_channel_map = control_channel_map
