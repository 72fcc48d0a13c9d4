"""casq_b200.backends (mirrors src/casq/backends/__init__.py)."""
from .b200.b200_pulse_backend import B200PulseBackend, BatchSolution
from .helpers import BackendLibrary, build, build_from_backend
from .pulse_backend import PulseBackend

__all__ = ["build", "build_from_backend", "PulseBackend", "B200PulseBackend", "BatchSolution", "BackendLibrary"]
