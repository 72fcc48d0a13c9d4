"""Noise model: T1 / T2 Lindblad operators for transmon qubits.

casq names a NoiseModel in its object model (docs/source/_static/images/casq-object-model.gv:9) and exposes
T1/T2 values (src/casq/backends/qiskit/backend_characteristics.py:107-123) but never turns them into operators
(the Solver it builds gets no dissipators, src/casq/backends/qiskit/qiskit_pulse_backend.py:174-183; roadmap item,
docs/source/roadmap.rst:32).  This class is that missing piece, for the Lindblad path of the B200 engine.
Per qubit q (SURVEY §8(d) cfg 4):  L1 = sqrt(1/T1) a_q,  Lphi = sqrt(2 gamma_phi) a_q^+ a_q,
gamma_phi = 1/T2 - 1/(2 T1)  (must be >= 0, i.e. T2 <= 2 T1).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from ..common.exceptions import CasqError


class NoiseModel:
    """NoiseModel(qubit_map={q: NoiseModel.QubitNoise(t1, t2)}) -> static Lindblad operators on the model's space."""

    @dataclass
    class QubitNoise:
        t1: Optional[float] = None  # seconds; None = no relaxation
        t2: Optional[float] = None  # seconds; None = no extra dephasing

    def __init__(self, qubit_map: dict[int, "NoiseModel.QubitNoise"]):
        for q, n in qubit_map.items():
            if n.t1 is not None and n.t1 <= 0:
                raise CasqError(f"T1 of qubit {q} must be positive.")
            if n.t2 is not None:
                if n.t2 <= 0:
                    raise CasqError(f"T2 of qubit {q} must be positive.")
                if n.t1 is not None and n.t2 > 2 * n.t1 * (1 + 1e-12):
                    raise CasqError(f"T2 of qubit {q} exceeds 2*T1: the pure-dephasing rate would be negative.")
        self.qubit_map = dict(qubit_map)

    def lindblad_operators(self, qubit_dims: dict[int, int]) -> np.ndarray:
        """[J, d, d] operators on the space spanned by ``qubit_dims`` (HamiltonianModel.qubit_dims; subsystem 0 is
        the least-significant Kronecker factor)."""
        order = sorted(qubit_dims)
        total = int(np.prod([qubit_dims[q] for q in order]))
        ops = []
        for q, noise in sorted(self.qubit_map.items()):
            if q not in qubit_dims:
                continue
            dim = qubit_dims[q]
            a = np.diag(np.sqrt(np.arange(1, dim)), k=1).astype(complex)
            num = a.conj().T @ a

            def embed(local):
                out = np.ones((1, 1), dtype=complex)
                for s in order:
                    out = np.kron(local if s == q else np.eye(qubit_dims[s]), out)
                return out

            g1 = 0.0 if noise.t1 is None else 1.0 / noise.t1
            gphi = 0.0 if noise.t2 is None else 1.0 / noise.t2 - 0.5 * g1
            if g1 > 0:
                ops.append(np.sqrt(g1) * embed(a))
            if gphi > 0:
                ops.append(np.sqrt(2.0 * gphi) * embed(num))
        return np.array(ops, dtype=complex).reshape(-1, total, total)
