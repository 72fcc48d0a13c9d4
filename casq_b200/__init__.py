"""casq_b200: B200-native pulse-simulation engine behind casq's PulseBackend API.

Package layout follows casq (backends / gates / models / optimizers / common); the arithmetic lives in
``csrc/`` (hand-written sm_100a CUDA behind the C ABI of ``include/casq_b200.h``).
"""
__version__ = "0.1.0"
