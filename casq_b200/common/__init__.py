"""casq_b200.common (mirrors src/casq/common/__init__.py; plotting is out of scope)."""
from .decorators import timer, trace
from .exceptions import CasqError
from .helpers import DiscreteSignal, SignalData, TimeUnit, dbid, discretize, ufid

__all__ = ["CasqError", "trace", "timer", "dbid", "ufid", "discretize", "SignalData", "DiscreteSignal", "TimeUnit"]
