"""Logging decorators (mirrors src/casq/common/decorators.py:34-107 on top of the standard ``logging`` module;
loguru is not a dependency).  ``timer(unit="msec")`` reports real milliseconds (the reference multiplies by 1e6,
decorators.py:101)."""
from __future__ import annotations

import functools
import inspect
import logging
import time
from typing import Any, Callable

logger = logging.getLogger("casq_b200")
TRACE = 5
logging.addLevelName(TRACE, "TRACE")


def _level(level) -> int:
    if isinstance(level, int):
        return level
    return {"TRACE": TRACE}.get(level.upper(), logging.getLevelName(level.upper()) if isinstance(logging.getLevelName(level.upper()), int) else logging.DEBUG)


def trace(*, log_entry: bool = True, log_exit: bool = True, level: str = "TRACE") -> Any:
    """Log entry to and exit from the decorated function."""

    def wrapper(func: Callable) -> Any:
        @functools.wraps(func)
        def wrapped(*args: Any, **kwargs: Any) -> Any:
            name = func.__name__
            if log_entry:
                logger.log(_level(level), "Entering [%s(args=%s, kwargs=%s)].", name, inspect.getfullargspec(func)[0],
                           [f"{k}={v!r}" for k, v in kwargs.items()])
            result = func(*args, **kwargs)
            if log_exit:
                logger.log(_level(level), "Exiting [%s] with result [%r].", name, result)
            return result

        return wrapped

    return wrapper


def timer(*, level: str = "DEBUG", unit: str = "msec") -> Any:
    """Log the wall time of the decorated function ('msec' or 'sec').  Not a benchmarking tool: GPU work is
    asynchronous; bench.py times with CUDA events."""

    def wrapper(func: Callable) -> Any:
        @functools.wraps(func)
        def wrapped(*args: Any, **kwargs: Any) -> Any:
            start = time.perf_counter()
            result = func(*args, **kwargs)
            elapsed = time.perf_counter() - start
            if unit.lower() == "sec":
                logger.log(_level(level), "Executed [%s] in %.3f seconds.", func.__name__, elapsed)
            else:
                logger.log(_level(level), "Executed [%s] in %.3f milliseconds.", func.__name__, elapsed * 1e3)
            return result

        return wrapped

    return wrapper
