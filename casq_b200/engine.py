"""Batched device solves: thin torch-tensor wrappers over the C ABI (include/casq_b200.h).

PyTorch is the host container only (device memory, streams); all arithmetic happens in libcasq_b200.so.
Every function raises CasqError when CUDA or the library is unavailable -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _native
from .common.exceptions import CasqError

PARAM_ORDER = ("amp", "angle", "sigma", "width", "beta", "freq_shift")


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise CasqError("casq_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", device if isinstance(device, int) else (device.index or 0))


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def pack_params(slot_params: Sequence[dict], batch: Optional[int] = None, device=None, pin: bool = False) -> torch.Tensor:
    """Build the SoA parameter block [n_slots, 6, B] (float64) from per-slot dicts.

    Each dict maps amp/angle/sigma/width/beta/freq_shift (Hz) to a scalar or a length-B array; missing entries are 0
    (angle None -> 0, like qiskit's symbolic pulses).  Returned on `device` when given, else on the host
    (pinned when `pin`).
    """
    B = batch
    for p in slot_params:
        for v in p.values():
            if v is not None and np.ndim(v) > 0:
                n = len(v)
                if B is not None and B != n and B != 1:
                    raise CasqError(f"inconsistent batch sizes in pulse parameters: {B} vs {n}")
                B = n if (B is None or B == 1) else B
    B = 1 if B is None else B
    host = torch.zeros((max(1, len(slot_params)), _native.NPARAM, B), dtype=torch.float64, pin_memory=pin)
    hv = host.numpy()
    for s, p in enumerate(slot_params):
        unknown = set(p) - set(PARAM_ORDER)
        if unknown:
            raise CasqError(f"unknown pulse parameter(s): {sorted(unknown)}")
        for j, name in enumerate(PARAM_ORDER):
            v = p.get(name)
            if v is None:
                continue
            hv[s, j, :] = np.asarray(v, dtype=np.float64)
    return host if device is None else host.to(device, non_blocking=True)


def envelope_sample(kind, duration: int, params: torch.Tensor, start: int = 0, dt: float = 0.0) -> torch.Tensor:
    """samples[b, n] = envelope(n + 1/2) on the device.  params: CUDA float64 [6, B] (one slot); `start` and `dt` only
    matter for points with a frequency shift (phase exp(i 2 pi freq_shift dt (start + n)))."""
    if not params.is_cuda:
        raise CasqError("envelope_sample: params must be a CUDA tensor")
    params = params.contiguous()
    if params.dim() != 2 or params.shape[0] != _native.NPARAM or params.dtype != torch.float64:
        raise CasqError("envelope_sample: params must be float64 [6, B]")
    B = params.shape[1]
    out = torch.empty((B, duration), dtype=torch.complex128, device=params.device)
    k = _native.PULSE_KINDS[kind] if isinstance(kind, str) else int(kind)
    with torch.cuda.device(params.device):
        _native.check(
            _native.load().casq_envelope_sample(k, int(duration), B, params.data_ptr(), int(start), float(dt), out.data_ptr(),
                                                _stream_ptr(params.device)),
            "casq_envelope_sample",
        )
    return out


def solve_sv_rk4(model: _native.NativeModel, slots, params: torch.Tensor, y0, t_span, max_dt: float,
                 t_eval=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fixed-step RK4 for a batch of pulse points.  Returns CUDA complex128 [B, T, d]: the states
    Solver.solve returns (rotating frame, lab basis) at t_eval (or at the two ends of t_span).

    slots: [(kind, channel_index, start, duration)]; params: CUDA float64 [n_slots, 6, B];
    y0: [d] or [B, d] (numpy or torch, lab basis).
    """
    dev, params, slots, B, d, y0_t, T, te = _prep_solve(model, slots, params, y0, t_eval, "solve_sv_rk4")
    te_ptr = None if te is None else te.ctypes.data
    if out is None:
        out = torch.empty((B, T, d), dtype=torch.complex128, device=dev)
    elif out.shape != (B, T, d) or out.dtype != torch.complex128 or not out.is_cuda or not out.is_contiguous():
        raise CasqError("solve_sv_rk4: bad `out` tensor")
    with torch.cuda.device(dev):
        _native.check(
            _native.load().casq_solve_sv_rk4(
                model.handle, len(slots), _native.make_slots(slots), B, params.data_ptr(), y0_t.data_ptr(),
                1 if y0_t.dim() == 2 else 0, float(t_span[0]), float(t_span[1]), float(max_dt), T, te_ptr,
                out.data_ptr(), _stream_ptr(dev)),
            "casq_solve_sv_rk4",
        )
    return out


def _check_params(model, slots, params, what):
    """Shared argument check of the batched solves -> (device, contiguous params [n_slots, 6, B], slot list, B)."""
    dev = _require_cuda(params.device if isinstance(params, torch.Tensor) and params.is_cuda else model.device)
    if not (isinstance(params, torch.Tensor) and params.is_cuda and params.dtype == torch.float64):
        raise CasqError(f"{what}: params must be a CUDA float64 tensor [n_slots, 6, B]")
    params = params.contiguous()
    slots = list(slots)
    if params.dim() != 3 or params.shape[1] != _native.NPARAM or params.shape[0] != max(1, len(slots)):
        raise CasqError(f"{what}: params shape {tuple(params.shape)} does not match {len(slots)} slot(s)")
    return dev, params, slots, params.shape[2]


def _prep_solve(model, slots, params, y0, t_eval, what):
    dev, params, slots, B = _check_params(model, slots, params, what)
    d = model.dim
    y0_t = torch.as_tensor(np.asarray(y0) if not isinstance(y0, torch.Tensor) else y0).to(torch.complex128)
    if y0_t.shape not in ((d,), (B, d)):
        raise CasqError(f"{what}: y0 must be [{d}] or [{B},{d}], got {tuple(y0_t.shape)}")
    y0_t = y0_t.to(dev).contiguous()
    if t_eval is None:
        T, te = 2, None
    else:
        te = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64))
        T = len(te)
    return dev, params, slots, B, d, y0_t, T, te


def solve_sv_dopri5(model: _native.NativeModel, slots, params: torch.Tensor, y0, t_span, rtol: float = 1.4e-8,
                    atol: float = 1.4e-8, hmax: float = 0.0, mxstep: int = 0, t_eval=None, return_stats: bool = False):
    """Adaptive Dopri5 (jax.experimental.ode.odeint controller) for a batch of pulse points.

    Returns CUDA complex128 [B, T, d]; with return_stats also (nsteps int64 [B,2], status int32 [B])."""
    dev, params, slots, B, d, y0_t, T, te = _prep_solve(model, slots, params, y0, t_eval, "solve_sv_dopri5")
    out = torch.empty((B, T, d), dtype=torch.complex128, device=dev)
    nsteps = torch.zeros((B, 2), dtype=torch.int64, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    hmax = 0.0 if hmax is None or not np.isfinite(hmax) else float(hmax)
    mxstep = 0 if mxstep is None or not np.isfinite(mxstep) else int(mxstep)
    with torch.cuda.device(dev):
        _native.check(
            _native.load().casq_solve_sv_dopri5(
                model.handle, len(slots), _native.make_slots(slots), B, params.data_ptr(), y0_t.data_ptr(),
                1 if y0_t.dim() == 2 else 0, float(t_span[0]), float(t_span[1]), float(rtol), float(atol), hmax, mxstep,
                T, te.ctypes.data if te is not None else None, out.data_ptr(), nsteps.data_ptr(), status.data_ptr(),
                _stream_ptr(dev)),
            "casq_solve_sv_dopri5",
        )
    return (out, nsteps, status) if return_stats else out


def propagators_expm(model: _native.NativeModel, slots, params: torch.Tensor, t_span, max_dt: float) -> torch.Tensor:
    """Exponential-midpoint propagators for a batch of pulse points: CUDA complex128 [B, d, d]
    (rotating frame, lab basis).  One expm (scaling-and-squaring Pade) per step of the fixed-step grid."""
    dev, params, slots, B = _check_params(model, slots, params, "propagators_expm")
    d = model.dim
    out = torch.empty((B, d, d), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        _native.check(
            _native.load().casq_propagators_expm(model.handle, len(slots), _native.make_slots(slots), B, params.data_ptr(),
                                                 float(t_span[0]), float(t_span[1]), float(max_dt), out.data_ptr(),
                                                 _stream_ptr(dev)),
            "casq_propagators_expm",
        )
    return out


def solve_lindblad_rk4(model: _native.NativeModel, slots, params: torch.Tensor, rho0, t_span, max_dt: float, t_eval=None,
                       want_rho: bool = True, want_pops: bool = True):
    """Fixed-step RK4 of the Lindblad equation for a batch of pulse points (dissipators: model.set_dissipators).

    rho0: [d] state vector, [d,d] or [B,d,d] density matrix (lab frame).  Returns (rho [B,T,d,d] in the rotating
    frame or None, pops [B,T,d] dressed-basis populations after frame removal or None)."""
    dev, params, slots, B = _check_params(model, slots, params, "solve_lindblad_rk4")
    d = model.dim
    r0 = torch.as_tensor(np.asarray(rho0) if not isinstance(rho0, torch.Tensor) else rho0).to(torch.complex128)
    if r0.shape == (d,):
        r0 = torch.outer(r0, r0.conj())
    if r0.shape not in ((d, d), (B, d, d)):
        raise CasqError(f"solve_lindblad_rk4: rho0 must be [{d}], [{d},{d}] or [{B},{d},{d}], got {tuple(r0.shape)}")
    r0 = r0.to(dev).contiguous()
    if t_eval is None:
        T, te = 2, None
    else:
        te = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64))
        T = len(te)
    rho = torch.empty((B, T, d, d), dtype=torch.complex128, device=dev) if want_rho else None
    pops = torch.empty((B, T, d), dtype=torch.float64, device=dev) if want_pops else None
    with torch.cuda.device(dev):
        _native.check(
            _native.load().casq_solve_lindblad_rk4(
                model.handle, len(slots), _native.make_slots(slots), B, params.data_ptr(), r0.data_ptr(),
                1 if r0.dim() == 3 else 0, float(t_span[0]), float(t_span[1]), float(max_dt), T,
                te.ctypes.data if te is not None else None, rho.data_ptr() if rho is not None else None,
                pops.data_ptr() if pops is not None else None, _stream_ptr(dev)),
            "casq_solve_lindblad_rk4",
        )
    return rho, pops


def postprocess_sv(model: _native.NativeModel, times, raw: torch.Tensor, normalize: bool = True,
                   want_states: bool = True, want_probs: bool = True):
    """raw [B,T,d] (as returned by the solves) -> (states [B,T,d] lab/dressed/normalised, probs [B,T,d],
    pops [B,T,2] = casq's {"0","1"} with every level >= 1 folded into "1")."""
    if not (raw.is_cuda and raw.dtype == torch.complex128 and raw.dim() == 3 and raw.shape[2] == model.dim):
        raise CasqError("postprocess_sv: raw must be a CUDA complex128 tensor [B, T, d]")
    raw = raw.contiguous()
    B, T, d = raw.shape
    times = np.ascontiguousarray(np.asarray(times, dtype=np.float64))
    if times.shape != (T,):
        raise CasqError(f"postprocess_sv: need {T} times, got {times.shape}")
    dev = raw.device
    states = torch.empty_like(raw) if want_states else None
    probs = torch.empty((B, T, d), dtype=torch.float64, device=dev) if want_probs else None
    pops = torch.empty((B, T, 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _native.check(
            _native.load().casq_postprocess_sv(
                model.handle, B, T, times.ctypes.data, raw.data_ptr(), 1 if normalize else 0,
                states.data_ptr() if states is not None else None, probs.data_ptr() if probs is not None else None,
                pops.data_ptr(), _stream_ptr(dev)),
            "casq_postprocess_sv",
        )
    return states, probs, pops


def fp64_fma_probe(device=0, iters: int = 4096, reps: int = 5) -> float:
    """Measured fp64 FMA throughput (TFLOP/s) of this GPU: the roofline denominator for the RK4 kernel."""
    dev = _require_cuda(device)
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    props = torch.cuda.get_device_properties(dev)
    grid, block = props.multi_processor_count * 8, 256
    lib = _native.load()
    best = 0.0
    with torch.cuda.device(dev):
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(lib.casq_fp64_fma_probe(grid, block, iters, sink.data_ptr(), _stream_ptr(dev)), "casq_fp64_fma_probe")
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = max(best, 2.0 * 16 * iters * grid * block / (ms * 1e-3) / 1e12)
    return best
