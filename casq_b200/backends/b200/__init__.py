from .b200_pulse_backend import B200PulseBackend, BatchSolution

__all__ = ["B200PulseBackend", "BatchSolution"]
