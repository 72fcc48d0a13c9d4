"""casq_b200.models (mirrors src/casq/models/__init__.py)."""
from .control_model import ControlModel
from .hamiltonian_model import HamiltonianModel
from .noise_model import NoiseModel
from .transmon_model import TransmonModel

__all__ = ["HamiltonianModel", "ControlModel", "TransmonModel", "NoiseModel"]
