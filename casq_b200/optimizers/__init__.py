"""casq_b200.optimizers (mirrors src/casq/optimizers/__init__.py)."""
from .pulse_optimizer import PulseOptimizer, hellinger_fidelity
from .single_qubit_gates import SingleQubitGateOptimizer, XGateOptimizer

__all__ = ["PulseOptimizer", "SingleQubitGateOptimizer", "XGateOptimizer", "hellinger_fidelity"]
