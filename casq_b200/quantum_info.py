"""Tiny stand-ins for qiskit.quantum_info.Statevector / DensityMatrix (the types casq's Solution.states
holds, src/casq/backends/pulse_backend.py:98).  Only what the drop-in surface needs."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


class _State:
    def __init__(self, data, dims: Optional[Sequence[int]] = None):
        self.data = np.asarray(data, dtype=complex)
        n = self.data.shape[0]
        self._dims = tuple(dims) if dims is not None else (n,)
        if int(np.prod(self._dims)) != n:
            raise ValueError(f"dims {self._dims} do not match dimension {n}")

    def dims(self) -> tuple:
        return self._dims

    @property
    def dim(self) -> int:
        return self.data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def probabilities_dict(self) -> dict:
        """Keys are level strings, subsystem 0 right-most (qiskit little-endian)."""
        out = {}
        for idx, p in enumerate(self.probabilities()):
            digits, rem = [], idx
            for dim in self._dims:
                digits.append(str(rem % dim))
                rem //= dim
            out["".join(reversed(digits))] = float(p)
        return out


class Statevector(_State):
    def probabilities(self) -> np.ndarray:
        return np.abs(self.data) ** 2

    def __truediv__(self, s):
        return Statevector(self.data / s, self._dims)

    def __repr__(self):
        return f"Statevector({np.array2string(self.data, precision=6)}, dims={self._dims})"


class DensityMatrix(_State):
    def probabilities(self) -> np.ndarray:
        return np.real(np.diag(self.data))

    def __truediv__(self, s):
        return DensityMatrix(self.data / s, self._dims)

    def __repr__(self):
        return f"DensityMatrix(dim={self.dim}, dims={self._dims})"


def populations_by_qubit(probabilities, dims: Sequence[int], measured: Optional[Sequence[int]] = None) -> dict:
    """Outcome populations with every level >= 1 of each measured subsystem clipped to "1" -- what
    ``_get_memory_slot_probabilities(..., max_outcome_value=1)`` yields when ``subsystem_dims`` IS forwarded (casq does
    not forward it, src/casq/backends/qiskit/qiskit_pulse_backend.py:184-188, so its multi-qubit populations fold
    everything into one pseudo-qubit: SURVEY §5 quirk 2).  Bit-strings are big-endian in the measured list (slot 0
    right-most).  ``probabilities`` is one level-probability vector [prod(dims)] (numpy or torch)."""
    p = np.asarray(probabilities.cpu() if hasattr(probabilities, "cpu") else probabilities, dtype=float)
    dims = list(dims)
    if p.shape != (int(np.prod(dims)),):
        raise ValueError(f"need {int(np.prod(dims))} probabilities, got {p.shape}")
    measured = list(range(len(dims))) if measured is None else list(measured)
    out: dict = {}
    for idx, val in enumerate(p):
        levels, rem = [], idx
        for dim in dims:  # subsystem 0 is the least-significant digit
            levels.append(rem % dim)
            rem //= dim
        key = "".join(str(min(levels[q], 1)) for q in reversed(measured))
        out[key] = out.get(key, 0.0) + float(val)
    return out
