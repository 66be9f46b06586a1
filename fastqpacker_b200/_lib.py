"""Loads the in-tree CUDA library.  No CPU fallback: a missing library or missing GPU raises."""
from __future__ import annotations

import os

from ._binding import FqzContext, FqzError, FqzLibrary

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfqzgpu.so")
_library = None
_contexts: dict[int, FqzContext] = {}


def library() -> FqzLibrary:
    global _library
    if _library is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(fastqpacker_b200 has no CPU fallback)"
            )
        _library = FqzLibrary(LIB_PATH)
    return _library


def context(device: int = 0) -> FqzContext:
    """Process-wide context of one GPU (created on first use)."""
    if device not in _contexts:
        _contexts[device] = library().context(device)
    return _contexts[device]


__all__ = ["library", "context", "FqzError", "LIB_PATH"]
