"""TEST-ONLY loader of the CPU-emulated build of the CUDA sources (tests/emu/Makefile).
Used by tests/test_emu_*.py to debug kernel logic without a GPU; never used by the product."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_ctx = None


def emu_context():
    global _ctx
    if _ctx is None:
        subprocess.check_call(["make", "-C", HERE, "-s", "-j8"])
        from fastqpacker_b200._binding import FqzLibrary

        # FQZ_EMU_LIB selects the ASan/UBSan variant (make SAN=... B=_build_asan), run under LD_PRELOAD=libasan
        _ctx = FqzLibrary(os.environ.get("FQZ_EMU_LIB", os.path.join(HERE, "_build", "libfqzgpu_emu.so"))).context(0)
    return _ctx
