"""ctypes binding of libcasq_b200.so (include/casq_b200.h).  There is no CPU fallback: if the library is
missing or a call fails, a CasqError is raised (src/casq/common/exceptions.py:27-45 is the convention)."""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np

from .common.exceptions import CasqError

import os

# CASQ_B200_LIB: tuning knob -- load another build of the same library (kernel variants compared in one GPU session)
_LIB_PATH = Path(os.environ.get("CASQ_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libcasq_b200.so")
_lib = None

NPARAM = 6
MAX_SLOTS = 4
PULSE_KINDS = {"constant": 0, "gaussian": 1, "drag": 2, "gaussian_square": 3, "gaussian_square_drag": 4}


class PulseSlot(C.Structure):
    _fields_ = [("kind", C.c_int32), ("channel", C.c_int32), ("start", C.c_int32), ("duration", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    "casq_last_error": (C.c_char_p, []),
    "casq_version": (C.c_int, []),
    "casq_model_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, _P, _P, C.c_double, C.c_int, _P, C.c_double, C.c_int]),
    "casq_model_destroy": (None, [_P]),
    "casq_model_set_dissipators": (C.c_int, [_P, C.c_int, _P]),
    "casq_model_get": (C.c_int, [_P, _P, _P, _P, _P]),
    "casq_model_dim": (C.c_int, [_P]),
    "casq_model_invalidate_cache": (C.c_int, [_P]),
    "casq_envelope_sample": (C.c_int, [C.c_int, C.c_int, C.c_int64, _P, C.c_int, C.c_double, _P, _P]),
    "casq_solve_sv_rk4": (C.c_int, [_P, C.c_int, C.POINTER(PulseSlot), C.c_int64, _P, _P, C.c_int, C.c_double, C.c_double,
                                    C.c_double, C.c_int, _P, _P, _P]),
    "casq_solve_sv_dopri5": (C.c_int, [_P, C.c_int, C.POINTER(PulseSlot), C.c_int64, _P, _P, C.c_int, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int, _P, _P, _P, _P, _P]),
    "casq_propagators_expm": (C.c_int, [_P, C.c_int, C.POINTER(PulseSlot), C.c_int64, _P, C.c_double, C.c_double, C.c_double, _P, _P]),
    "casq_solve_lindblad_rk4": (C.c_int, [_P, C.c_int, C.POINTER(PulseSlot), C.c_int64, _P, _P, C.c_int, C.c_double, C.c_double,
                                          C.c_double, C.c_int, _P, _P, _P, _P]),
    "casq_postprocess_sv": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P]),
    "casq_fp64_fma_probe": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises CasqError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise CasqError(
            f"{_LIB_PATH} is missing: build it with `python -m casq_b200._build` (nvcc, sm_100a). "
            "casq_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().casq_last_error().decode(errors="replace")
        raise CasqError(f"{what} failed (status {status}): {msg}")


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.complex128))


class NativeModel:
    """Owns one casq_model handle (the device-engine counterpart of qiskit-dynamics' Solver + the dressed
    basis of DynamicsBackend; built where QiskitPulseBackend._get_native_backend builds those,
    src/casq/backends/qiskit/qiskit_pulse_backend.py:166-189)."""

    def __init__(self, static, operators, carrier_hz, dt, rotating_frame=None, rwa_cutoff_freq=None, device=0):
        lib = load()
        static = _c128(static)
        d = static.shape[0]
        operators = _c128(operators).reshape(-1, d, d)
        K = operators.shape[0]
        carrier = np.ascontiguousarray(np.asarray(carrier_hz, dtype=np.float64))
        if carrier.shape != (K,):
            raise CasqError(f"need one carrier frequency per channel ({K}), got shape {carrier.shape}")
        if rotating_frame is None:
            kind, frame = 0, None
        else:
            rf = np.asarray(rotating_frame)
            if rf.ndim == 1:
                kind, frame = 2, np.ascontiguousarray(rf.real.astype(np.float64))
            elif rf.shape == (d, d):
                kind, frame = 1, _c128(rf)
            else:
                raise CasqError(f"rotating_frame must be [d] or [d,d], got {rf.shape}")
        self._keep = (static, operators, carrier, frame)
        h = _P()
        check(
            lib.casq_model_create(C.byref(h), d, K, static.ctypes.data, operators.ctypes.data if K else None,
                                  carrier.ctypes.data if K else None, float(dt), kind,
                                  frame.ctypes.data if frame is not None else None,
                                  math.nan if rwa_cutoff_freq is None else float(rwa_cutoff_freq), int(device)),
            "casq_model_create",
        )
        self.handle = h
        self.dim, self.n_channels, self.dt, self.device = d, K, float(dt), int(device)
        fe = np.zeros(d)
        fb = np.zeros((d, d), dtype=np.complex128)
        de = np.zeros(d)
        ds = np.zeros((d, d), dtype=np.complex128)
        check(lib.casq_model_get(h, fe.ctypes.data, fb.ctypes.data, de.ctypes.data, ds.ctypes.data), "casq_model_get")
        self.frame_evals, self.frame_basis, self.dressed_evals, self.dressed_states = fe, fb, de, ds

    def set_dissipators(self, ops) -> None:
        ops = _c128(ops).reshape(-1, self.dim, self.dim)
        check(load().casq_model_set_dissipators(self.handle, ops.shape[0], ops.ctypes.data), "casq_model_set_dissipators")

    def invalidate_cache(self) -> None:
        check(load().casq_model_invalidate_cache(self.handle), "casq_model_invalidate_cache")

    def close(self) -> None:
        if getattr(self, "handle", None):
            load().casq_model_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


def make_slots(slots) -> "C.Array":
    """slots: iterable of (kind:str|int, channel_index, start, duration)."""
    slots = list(slots)
    arr = (PulseSlot * max(1, len(slots)))()
    for i, (kind, ch, start, dur) in enumerate(slots):
        arr[i] = PulseSlot(PULSE_KINDS[kind] if isinstance(kind, str) else int(kind), int(ch), int(start), int(dur))
    return arr
