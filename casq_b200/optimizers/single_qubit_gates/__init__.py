from .single_qubit_gate_optimizer import SingleQubitGateOptimizer
from .x_gate_optimizer import XGateOptimizer

__all__ = ["SingleQubitGateOptimizer", "XGateOptimizer"]
