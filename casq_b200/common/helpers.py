"""Identifier helpers (mirrors dbid/ufid of src/casq/common/helpers.py:50-72; no wonderwords dependency)."""
from __future__ import annotations

import random
from typing import Any
from uuid import uuid4

_ADJECTIVES = ("amber", "brisk", "calm", "deft", "eager", "fleet", "gentle", "hardy", "ivory", "jolly", "keen",
               "lucid", "merry", "noble", "opal", "proud", "quick", "rapid", "sunny", "tidy", "umber", "vivid",
               "warm", "young", "zesty")


def dbid() -> str:
    """Database identifier (UUID4 string)."""
    return str(uuid4())


def ufid(obj: Any) -> str:
    """User-friendly identifier: <adjective><ClassName>."""
    name = obj.__class__.__name__
    if not name[0].isupper():
        name = name.capitalize()
    return f"{random.choice(_ADJECTIVES)}{name}"
