"""fastqpacker_b200 — B200-native (sm_100a) block codec for fqpack's .fqz format.

Host-side mirror of the reference's `internal/compress` API over the C-ABI in include/fqzgpu.h.
Everything runs on the GPU through fastqpacker_b200/libfqzgpu.so; there is no CPU fallback.
"""
from ._lib import FqzError, context, library  # noqa: F401
