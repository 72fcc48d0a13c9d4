"""Backend factory (drop-in for src/casq/backends/helpers.py:39-120).

``BackendLibrary`` keeps casq's members and adds ``B200``; ``build`` is the dispatch point where the new
engine registers.  QISKIT / C3 / QUTIP raise the same ``ValueError`` casq raises for libraries it does not
implement (helpers.py:71-72): this package deliberately has no CPU backend.
"""
from __future__ import annotations

from enum import Enum
from typing import Any, Optional

from ..models.control_model import ControlModel
from ..models.hamiltonian_model import HamiltonianModel
from .b200.b200_pulse_backend import B200PulseBackend


class BackendLibrary(Enum):
    """Backend library."""

    C3 = 0
    QISKIT = 1
    QUTIP = 2
    B200 = 3


def build(backend_library: BackendLibrary, hamiltonian: HamiltonianModel, control: ControlModel,
          seed: Optional[int] = None, device: int = 0, noise: Any = None) -> B200PulseBackend:
    """Build a PulseBackend (same signature as casq's ``build`` plus ``device`` and an optional ``NoiseModel``)."""
    if backend_library is BackendLibrary.B200:
        return B200PulseBackend(hamiltonian=hamiltonian, control=control, seed=seed, device=device, noise=noise)
    raise ValueError(f"Unknown backend library: {backend_library.name}.")


def build_from_backend(backend: Any, *args: Any, **kwargs: Any) -> B200PulseBackend:
    """casq's ``build_from_backend`` scrapes a live qiskit ``Backend`` object (helpers.py:75-120,
    backend_characteristics.py:86-105).  qiskit is not a dependency here, so any object is an unknown backend;
    build the models from ``TransmonModel.MANILA`` / ``TransmonModel.manila()`` and call ``build`` instead."""
    raise ValueError(f"Unknown backend: {repr(backend)}.")
