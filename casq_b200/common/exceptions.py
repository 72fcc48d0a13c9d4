"""Exceptions (mirrors src/casq/common/exceptions.py:27-45: ``CasqError(*message)`` joins its arguments)."""
from __future__ import annotations

from typing import Any


class CasqError(Exception):
    """Error raised by casq_b200, including every non-zero status returned by the C ABI."""

    def __init__(self, *message: Any) -> None:
        super().__init__(" ".join(map(str, message)))
        self.message = " ".join(map(str, message))

    def __str__(self) -> str:
        return repr(self.message)
