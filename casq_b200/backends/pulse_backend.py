"""Pulse backend API (drop-in for src/casq/backends/pulse_backend.py:51-366).

Same abstract interface, ``ODESolverMethod`` enum and ``Solution`` dataclass fields as casq.  The
matplotlib/qutip plot helpers of casq's Solution (pulse_backend.py:108-318) are presentation code outside
the accelerated path and are not provided; the Bloch-vector maths they wrap is (``bloch_vectors``).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from datetime import datetime
from enum import Enum
from typing import Any, Optional, Union

import numpy as np

from ..gates.pulse_circuit import PulseCircuit
from ..models.control_model import ControlModel
from ..models.hamiltonian_model import HamiltonianModel
from ..quantum_info import DensityMatrix, Statevector


class PulseBackend(ABC):
    """PulseBackend(hamiltonian, control, seed=None); builds its native backend eagerly (pulse_backend.py:320-330)."""

    class ODESolverMethod(str, Enum):
        """ODE solver methods (same members and values as casq, pulse_backend.py:60-71)."""

        QISKIT_DYNAMICS_RK4 = "RK4"
        QISKIT_DYNAMICS_JAX_RK4 = "jax_RK4"
        QISKIT_DYNAMICS_JAX_ODEINT = "jax_odeint"
        SCIPY_BDF = "BDF"
        SCIPY_DOP853 = "DOP853"
        SCIPY_LSODA = "LSODA"
        SCIPY_RADAU = "Radau"
        SCIPY_RK23 = "RK23"
        SCIPY_RK45 = "RK45"

    @dataclass
    class Solution:
        """Same fields as casq's PulseBackend.Solution (pulse_backend.py:73-106)."""

        circuit_name: str
        qubits: list[int]
        times: list[float]
        samples: list[list[int]]
        counts: list[dict[str, int]]
        populations: list[dict[str, float]]
        states: list[Union[DensityMatrix, Statevector]]
        iq_data: list[list[tuple[float, float]]]
        avg_iq_data: list[tuple[float, float]]
        shots: int = 1024
        seed: Optional[int] = None
        is_success: bool = True
        timestamp: float = field(default_factory=lambda: datetime.timestamp(datetime.now()))

        def bloch_vectors(self, qubit: int = 0) -> np.ndarray:
            """<X>, <Y>, <Z> of the qubit subspace (levels 0, 1) per time point: [T, 3].

            The maths behind casq's plot_bloch_trajectory/_xyz helpers (pulse_backend.py:289-318) for the
            single-subsystem states casq produces."""
            out = []
            for s in self.states:
                if isinstance(s, Statevector):
                    a, b = s.data[0], s.data[1]
                    rho = np.array([[a * np.conj(a), a * np.conj(b)], [b * np.conj(a), b * np.conj(b)]])
                else:
                    rho = s.data[:2, :2]
                out.append([2 * rho[0, 1].real, 2 * rho[1, 0].imag, (rho[0, 0] - rho[1, 1]).real])
            return np.asarray(out, dtype=float)

    def __init__(self, hamiltonian: HamiltonianModel, control: ControlModel, seed: Optional[int] = None):
        self.hamiltonian = hamiltonian
        self.control = control
        self._seed = seed
        self._native_backend = self._get_native_backend()

    @abstractmethod
    def solve(
        self,
        circuit: PulseCircuit,
        method: "PulseBackend.ODESolverMethod",
        initial_state: Optional[Union[DensityMatrix, Statevector]] = None,
        shots: int = 1024,
        steps: Optional[int] = None,
        run_options: Optional[dict[str, Any]] = None,
    ) -> "PulseBackend.Solution":
        """Simulate one pulse circuit (pulse_backend.py:332-358)."""

    @abstractmethod
    def _get_native_backend(self) -> Any:
        """Build the engine-specific backend object (pulse_backend.py:360-366)."""
