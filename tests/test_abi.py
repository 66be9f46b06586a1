"""The C-ABI library loads without a GPU and exports every symbol include/fqzgpu.h declares; without
a CUDA device every compute entry point refuses to run (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fqzgpu.h")
LIB = os.path.join(ROOT, "fastqpacker_b200", "libfqzgpu.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fqz_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g

        g.build()
    return C.CDLL(LIB)


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for name in ("fqz_init", "fqz_compress", "fqz_decompress", "fqz_compress_feed", "fqz_decompress_feed", "fqz_compress_shard",
                 "fqz_encode_streams", "fqz_decode_streams", "fqz_zstd_compress", "fqz_zstd_decompress", "fqz_compress_device",
                 "fqz_decompress_device", "fqz_is_gzip", "fqz_gunzip", "fqz_gunzip_device", "fqz_compress_gz", "fqz_gunzip_stats", "fqz_info",
                 "fqz_check", "fqz_block_index", "fqz_decompress_blocks"):
        assert name in syms


@pytest.mark.parametrize("name", declared_symbols())
def test_library_exports(lib, name):
    assert getattr(lib, name) is not None


def test_abi_version_and_error_texts(lib):
    lib.fqz_abi_version.restype = C.c_int
    assert lib.fqz_abi_version() == 1
    lib.fqz_strerror.restype = C.c_char_p
    lib.fqz_strerror.argtypes = [C.c_int]
    # texts follow the reference (parser.go:143,164,180; container.go:54; compress.go:979-1057)
    assert lib.fqz_strerror(-1) == b"invalid FASTQ: header line must start with @"
    assert lib.fqz_strerror(-2) == b"invalid FASTQ: separator line must start with +"
    assert lib.fqz_strerror(-3) == b"invalid FASTQ: sequence and quality lengths must match"
    assert lib.fqz_strerror(-5) == b"invalid magic bytes: not an FQZ file"
    assert lib.fqz_strerror(-9) == b"truncated header data"
    assert lib.fqz_strerror(-14) == b"truncated N position data"
    # Go's compress/gzip and compress/flate texts for a gzipped input (cmd/fqpack/main.go:142-174)
    assert lib.fqz_strerror(-18) == b"gzip: invalid header"
    assert lib.fqz_strerror(-19) == b"gzip: invalid checksum"
    assert lib.fqz_strerror(-20).startswith(b"flate: corrupt input")
    assert lib.fqz_strerror(-21) == b"unexpected EOF"


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    lib.fqz_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    assert lib.fqz_init(0, C.byref(ctx)) == -33  # FQZ_E_NO_DEVICE
    assert not ctx.value


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "fastqpacker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "fqz_oracle" not in src and "libfqzoracle" not in src and "from oracle" not in src, f


def test_block_index_runs_without_a_gpu(lib, oracle):
    """fqz_block_index is header arithmetic on the host: it needs neither a context nor a device, so the real library's
    version is checked here against the Python header walk and the oracle."""
    from fastqpacker_b200._binding import FqzLibrary
    from tests.decode_cases import check_block_index

    check_block_index(FqzLibrary(LIB), oracle)
