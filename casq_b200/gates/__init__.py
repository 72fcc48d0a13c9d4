"""casq_b200.gates (mirrors src/casq/gates/__init__.py)."""
from .constant_pulse_gate import ConstantPulseGate
from .cross_resonance_pulse_gate import CrossResonancePulseGate
from .drag_pulse_gate import DragPulseGate
from .gaussian_pulse_gate import GaussianPulseGate
from .gaussian_square_drag_pulse_gate import GaussianSquareDragPulseGate
from .gaussian_square_pulse_gate import GaussianSquarePulseGate
from .pulse_circuit import PulseCircuit
from .pulse_gate import PulseGate

__all__ = ["PulseGate", "PulseCircuit", "ConstantPulseGate", "DragPulseGate", "GaussianPulseGate",
           "GaussianSquarePulseGate", "GaussianSquareDragPulseGate", "CrossResonancePulseGate"]
