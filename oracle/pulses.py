"""Oracle: pulse envelopes (midpoint sampling) and DiscreteSignal evaluation.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: restates
  * qiskit-terra 0.24.2 ``qiskit.pulse.library.symbolic_pulses`` {Constant, Gaussian, Drag,
    GaussianSquare, GaussianSquareDrag}, reached from casq at src/casq/gates/*_pulse_gate.py
    (e.g. drag_pulse_gate.py:75-82, gaussian_square_pulse_gate.py:76-83)  -- SURVEY Appendix A.2
  * qiskit-dynamics 0.4.1 ``InstructionToSignals`` / ``DiscreteSignal`` reached from
    src/casq/common/helpers.py:124-127 and inside Solver.solve                -- SURVEY Appendix A.3
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

KINDS = ("constant", "gaussian", "drag", "gaussian_square", "gaussian_square_drag")


def _lifted_gaussian(t, center, t_zero, sigma):
    g = np.exp(-(((t - center) / sigma) ** 2) / 2)
    off = np.exp(-(((t_zero - center) / sigma) ** 2) / 2)
    return (g - off) / (1 - off)


def envelope_samples(kind: str, duration: int, amp, angle=0.0, sigma=None, width=None, beta=None):
    """samples[..., n] = envelope(n + 1/2), n = 0..duration-1; leading dims broadcast over parameters."""
    amp = np.asarray(amp, dtype=float)[..., None]
    angle = np.asarray(0.0 if angle is None else angle, dtype=float)[..., None]
    A = amp * np.exp(1j * angle)
    t = np.arange(duration, dtype=float) + 0.5
    c = duration / 2
    if kind == "constant":
        return A * np.ones_like(t)
    sigma = np.asarray(sigma, dtype=float)[..., None]
    if np.any(sigma <= 0):
        raise ValueError("sigma must be > 0")
    if kind == "gaussian":
        return A * _lifted_gaussian(t, c, duration + 1, sigma)
    if kind == "drag":
        beta = np.asarray(beta, dtype=float)[..., None]
        g = _lifted_gaussian(t, c, duration + 1, sigma)
        return A * (g + 1j * beta * (-(t - c) / sigma**2) * g)
    width = np.asarray(width, dtype=float)[..., None]
    if np.any(width < 0) or np.any(width > duration):
        raise ValueError("width must be in [0, duration]")
    t_l = c - width / 2
    t_r = c + width / 2
    led = _lifted_gaussian(t, t_l, -1, sigma)
    red = _lifted_gaussian(t, t_r, duration + 1, sigma)
    if kind == "gaussian_square":
        env = np.where(t <= t_l, led, np.where(t >= t_r, red, 1.0))
        return A * env
    if kind == "gaussian_square_drag":
        beta = np.asarray(beta, dtype=float)[..., None]
        dl = -(t - t_l) / sigma**2 * led
        dr = -(t - t_r) / sigma**2 * red
        env = np.where(t <= t_l, led + 1j * beta * dl, np.where(t >= t_r, red + 1j * beta * dr, 1.0 + 0j))
        return A * env
    raise ValueError(f"unknown pulse kind {kind}")


@dataclass
class PulseSpec:
    """One Play instruction: which family on which Hamiltonian channel, starting at which sample."""

    kind: str
    channel: str
    duration: int
    start: int = 0


def channel_samples(model_channels, specs, params, total_duration=None, dt=None):
    """InstructionToSignals.get_signals restated: one complex sample array per Hamiltonian channel.

    specs: list[PulseSpec]; params: list of dicts (amp, angle, sigma, width, beta, freq_shift), each value a
    scalar or [B] array.  Returns samples [B, K, N] (zeros where nothing is played).
    freq_shift (Hz, needs ``dt``): the frequency shift accumulated on the channel before the Play
    (ShiftFrequency / SetFrequency); upstream multiplies the played samples by
    exp(2j pi freq_shift * dt * (start_sample + arange(n_samples))) -- SURVEY Appendix A.3.
    """
    B = 1
    for p in params:
        for v in p.values():
            if v is not None and np.ndim(v) > 0:
                B = max(B, len(v))
    N = max(s.start + s.duration for s in specs) if total_duration is None else total_duration
    out = np.zeros((B, len(model_channels), N), dtype=complex)
    for s, p in zip(specs, params):
        k = list(model_channels).index(s.channel)
        env = envelope_samples(s.kind, s.duration, **{kk: vv for kk, vv in p.items() if kk != "freq_shift"})
        env = np.broadcast_to(env, (B, s.duration)) if env.ndim > 1 else np.broadcast_to(env[None], (B, s.duration))
        fs = p.get("freq_shift")
        if fs is not None and np.any(np.asarray(fs) != 0):
            if dt is None:
                raise ValueError("freq_shift needs dt")
            times = dt * (s.start + np.arange(s.duration, dtype=float))
            env = env * np.exp(2j * np.pi * np.asarray(fs, dtype=float).reshape(-1, 1) * times)
        out[:, k, s.start : s.start + s.duration] += env
    return out


def floor_divide_index(t, dt, n_samples):
    """DiscreteSignal.envelope index: clip(int((t - 0) // dt), -1, N); -1 and N address a zero pad."""
    idx = np.asarray(np.floor_divide(np.asarray(t, dtype=float), dt)).astype(np.int64)
    return np.clip(idx, -1, n_samples)


def signal_values(samples, nu, t, dt):
    """Complex channel values z_k(t) = samples_k[idx(t)] * exp(i 2 pi nu_k t).  samples [B,K,N] -> [B,K].

    The real signal of qiskit-dynamics is Re z_k(t).
    """
    N = samples.shape[-1]
    idx = int(floor_divide_index(t, dt, N))
    if idx < 0 or idx >= N:
        env = np.zeros(samples.shape[:-1], dtype=complex)
    else:
        env = samples[..., idx]
    return env * np.exp(1j * 2 * np.pi * nu * t)
